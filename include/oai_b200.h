/* liboai_b200 -- C ABI of the B200-native per-knee inference hot path of OAI Analysis 2.
 *
 * The reference (uncbiag/OAI_analysis_2) is pure Python and has no FFI of its own: its hot path calls
 * torch / ITK / icon_registration library ops.  Each entry point below names the reference call it replaces
 * (paths relative to the reference repository root).  INTEGRATION.md shows the ctypes stub a maintainer adds on the
 * reference side.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; oai_last_error() returns a message for the calling
 *     thread's last failure;
 *   - the caller owns every device buffer and the CUDA stream (passed as void* == cudaStream_t); nothing here
 *     allocates caller-visible memory, synchronises the device, or keeps a pointer after returning;
 *   - "act16" tensors are channels-last NDHWC, 16-bit (fp16 when ab_format == 0, bf16 when 1);
 *   - volumes are dense z,y,x (numpy order of the reference, image_transforms.py:376-379) float32.
 */
#ifndef OAI_B200_H_
#define OAI_B200_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define OAI_API __attribute__((visibility("default")))
#else
#define OAI_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

OAI_API const char* oai_last_error(void);
OAI_API int oai_version(void);
/* number of kernels launched through this library since load (bench.py reports it as gpu_launches) */
OAI_API long long oai_launch_count(void);

/* ------------------------------------------------------------------------------------------------------------
 * Segmentation UNet layers  (reference: oai_analysis/segmentation/networks.py:38-149)
 * ------------------------------------------------------------------------------------------------------------ */

/* 3x3x3 stride-1 "same" convolution or pointwise convolution as a tcgen05 implicit GEMM.
 * Replaces nn.Conv3d(k=3,p=1) / nn.ConvTranspose3d(k=3,s=1,p=1) [+BatchNorm3d(eval)+ReLU folded]
 * (networks.py:80-107) and, with pointwise=1, each of the 8 sub-filters of nn.ConvTranspose3d(k=2,s=2).
 *
 *   src0/src1 : act16 [NT, D, H, W, c0] / [NT, D, H, W, c1]  (src1 may be NULL with c1 == 0; the K loop runs over
 *               src0's channels then src1's, which is torch.cat((src0, src1), dim=1) of networks.py:127,134,141)
 *   wpack     : weights packed by oai_pack_conv_weights() for the same geometry
 *   bias      : float32 [cout]
 *   out       : act16, element offset of voxel (n,d,h,w) channel c is
 *               obase + n*osN + d*osD + h*osH + w*osW + c   (lets a k2s2 transposed conv scatter its 8 sub-grids)
 */
OAI_API int oai_conv3d_igemm(const void* src0, int c0, const void* src1, int c1, int NT, int D, int H, int W,
                     const void* wpack, size_t wpack_bytes, const float* bias, int cout, int pointwise, int relu,
                     int ab_format, void* out, long long obase, long long osN, long long osD, long long osH,
                     long long osW, int flags, void* stream);

/* Geometry the kernel will use for (D,H,W,cin,cout,pointwise): fills plan[8] =
 * {mode, kd_per_block, R, nhalf, cout_per_half, nblk, wblock_bytes, nchunks}.  Pure host arithmetic (no GPU). */
OAI_API int oai_conv3d_igemm_plan(int D, int H, int W, int c0, int c1, int cout, int pointwise, int flags, int* plan);

/* Pack float32 conv weights [cout][cin][3][3][3] (or [cout][cin] when pointwise) into the pre-swizzled
 * shared-memory block images the kernel streams with cp.async.bulk.  Host function (no GPU).
 * dst must hold oai_conv3d_igemm_plan()'s nhalf*nblk*wblock_bytes bytes. */
OAI_API int oai_pack_conv_weights(const float* w, int cout, int c0, int c1, int D, int H, int W, int pointwise, int ab_format,
                          int flags, void* dst, size_t dst_bytes);

/* Tiling geometry arrays used below (all z,y,x): geom[12] = {tile[3], effective[3], overlap[3], grid[3]} exactly as
 * Partition computes them (image_transforms.py:389-391,404-406); vol_dims[3] = image size. */

/* Partition.__call__ (image_transforms.py:408-446: reflect pad + overlap tiling) fused with the first UNet layer
 * ec0 = Conv3d(1->c0,k3,p1)[+BN]+ReLU (networks.py:43,84-86).  Reads the float32 volume in place (tiles are never
 * materialised), writes act16 [ntiles, td, th, tw, c0] for tiles tile0 .. tile0+ntiles-1 (i-major tile order).
 * w27c: folded weights [27][c0] (tap = (kd*3+kh)*3+kw), bias [c0]. */
OAI_API int oai_seg_stem(const float* vol, const int* vol_dims, const int* geom, int tile0, int ntiles,
                         const float* w27c, const float* bias, int c0, void* out, int ab_format, void* stream);

/* nn.MaxPool3d(2) (networks.py:52-54) on act16 [N,D,H,W,C] -> [N,D/2,H/2,W/2,C]. */
OAI_API int oai_maxpool3d_2(const void* in, void* out, int N, int D, int H, int W, int C, int ab_format, void* stream);

/* dc0 = Conv3d(C->ncls,k1) (networks.py:66,148) + torch.sigmoid (segmenter.py:121) [+ ">0.5" when out_mode==1,
 * segmenter.py:123-124; out_mode==2 writes raw logits] + Partition.assemble (image_transforms.py:492-513): each tile's interior is written to its
 * place in out[ncls][VD][VH][VW] (float32), trimmed to the image, with the crop_zyx border shell set to 0. */
OAI_API int oai_seg_head(const void* act, int C, int ncls, const float* w, const float* b, float* out,
                         const int* vol_dims, const int* geom, int tile0, int ntiles, const int* crop_zyx, int out_mode,
                         int ab_format, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OAI_B200_H_ */
