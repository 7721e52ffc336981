// C-ABI entry points of the segmentation stage's bandwidth-bound kernels (stem / max-pool / head).
#include "../../include/oai_b200.h"
#include "api_common.h"
#include "seg_misc.cuh"


using namespace oai;

extern "C" int oai_seg_stem(const float* vol, const int* vol_dims, const int* geom, int tile0, int ntiles,
                            const float* w27c, const float* bias, int c0, void* out, int ab_format, void* stream) {
  return oai_seg_stem_ex(vol, vol_dims, geom, tile0, ntiles, w27c, bias, c0, out, ab_format, 0, stream);
}

extern "C" int oai_seg_stem_ex(const float* vol, const int* vol_dims, const int* geom, int tile0, int ntiles,
                               const float* w27c, const float* bias, int c0, void* out, int ab_format, int out_split,
                               void* stream) {
  OAI_REQUIRE(vol && vol_dims && geom && w27c && bias && out, "seg_stem: null pointer");
  OAI_REQUIRE(c0 % 8 == 0 && c0 > 0 && c0 <= 64, "seg_stem: c0=%d must be a multiple of 8 in (0,64]", c0);
  for (int a = 0; a < 3; ++a) OAI_REQUIRE(vol_dims[a] >= 1 && geom[a] >= 1, "seg_stem: empty axis %d", a);
  StemParams p;
  p.vol = vol; p.VD = vol_dims[0]; p.VH = vol_dims[1]; p.VW = vol_dims[2];
  p.td = geom[0]; p.th = geom[1]; p.tw = geom[2];
  p.ed = geom[3]; p.eh = geom[4]; p.ew = geom[5];
  p.od = geom[6]; p.oh = geom[7]; p.ow = geom[8];
  p.gh = geom[10]; p.gw = geom[11];
  p.tile0 = tile0; p.ntiles = ntiles; p.c0 = c0; p.w = w27c; p.b = bias; p.out = out; p.fmt = ab_format;
  p.out_split = out_split;
  return stem_launch(p, static_cast<cudaStream_t>(stream));
}

extern "C" int oai_maxpool3d_2(const void* in, void* out, int N, int D, int H, int W, int C, int ab_format,
                               void* stream) {
  OAI_REQUIRE(in && out, "maxpool: null pointer");
  OAI_REQUIRE(C % 8 == 0 && D % 2 == 0 && H % 2 == 0 && W % 2 == 0, "maxpool: need C%%8==0 and even D,H,W");
  return maxpool2_launch(in, out, N, D, H, W, C, 0, 0, ab_format, static_cast<cudaStream_t>(stream));
}

extern "C" int oai_maxpool3d_2_ex(const void* in, void* out, int N, int D, int H, int W, int C, int in_split,
                                  int out_split, int ab_format, void* stream) {
  OAI_REQUIRE(in && out, "maxpool: null pointer");
  OAI_REQUIRE(C % 8 == 0 && D % 2 == 0 && H % 2 == 0 && W % 2 == 0, "maxpool: need C%%8==0 and even D,H,W");
  OAI_REQUIRE(in_split || !out_split, "maxpool: a split output needs a split input");
  return maxpool2_launch(in, out, N, D, H, W, C, in_split, out_split, ab_format, static_cast<cudaStream_t>(stream));
}

extern "C" int oai_seg_head(const void* act, int C, int ncls, const float* w, const float* b, float* out,
                            const int* vol_dims, const int* geom, int tile0, int ntiles, const int* crop_zyx,
                            int out_mode, int ab_format, void* stream) {
  OAI_REQUIRE(act && w && b && out && vol_dims && geom && crop_zyx, "seg_head: null pointer");
  OAI_REQUIRE(C % 8 == 0 && C <= 64 && ncls >= 1 && ncls <= 8, "seg_head: C=%d ncls=%d unsupported", C, ncls);
  HeadParams p;
  p.act = act; p.C = C; p.ncls = ncls; p.w = w; p.b = b; p.out = out;
  p.VD = vol_dims[0]; p.VH = vol_dims[1]; p.VW = vol_dims[2];
  p.td = geom[0]; p.th = geom[1]; p.tw = geom[2];
  p.ed = geom[3]; p.eh = geom[4]; p.ew = geom[5];
  p.od = geom[6]; p.oh = geom[7]; p.ow = geom[8];
  p.gh = geom[10]; p.gw = geom[11];
  p.tile0 = tile0; p.ntiles = ntiles;
  p.cz = crop_zyx[0]; p.cy = crop_zyx[1]; p.cx = crop_zyx[2];
  p.out_mode = out_mode; p.fmt = ab_format;
  return head_launch(p, static_cast<cudaStream_t>(stream));
}
