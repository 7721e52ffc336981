// Bandwidth-bound pieces of the segmentation stage that sit around the tcgen05 convolutions:
//   stem  : reflect-pad + overlap-tile gather (Partition.__call__, image_transforms.py:408-446) fused with the first
//           conv ec0 (Conv3d 1->C0 k3 p1 + folded BN + ReLU, networks.py:43) -> NDHWC 16-bit
//   pool  : MaxPool3d(2) (networks.py:52-54) on NDHWC 16-bit
//   head  : dc0 (Conv3d 64->n_classes k1, networks.py:66,148) + torch.sigmoid (segmenter.py:121) [+ >0.5,
//           segmenter.py:123-124] + Partition.assemble crop-and-place with the zeroed border shell
//           (image_transforms.py:492-513), writing float32 class volumes
#include "api_common.h"
#include "seg_misc.cuh"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace oai {

__device__ __forceinline__ uint32_t pack16(float a, float b, int fmt) {
  if (fmt == 0) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack16(uint32_t u, int fmt) {
  if (fmt == 0) return __half22float2(*reinterpret_cast<__half2*>(&u));
  return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
}

__device__ __forceinline__ uint32_t pack16_residual(float a, float b, uint32_t hi, int fmt) {
  const float2 h = unpack16(hi, fmt);
  return pack16(a - h.x, b - h.y, fmt);
}

// numpy 'reflect' (edge sample not repeated), valid for pads smaller than the axis length
__device__ __forceinline__ int reflect_idx(int i, int n) {
  if (n == 1) return 0;
  const int period = 2 * (n - 1);
  i %= period;
  if (i < 0) i += period;
  return i < n ? i : period - i;
}


constexpr int kStemTW = 128, kStemTH = 4, kStemXS = 4;  // block: 4 rows x 128 x, walks all d-slices; thread: 4 x-voxels

// One block owns a (4-row x 128-x) column of a tile and walks its d-slices with a 3-slice rolling window in shared
// memory, so every input sample is fetched from L2 once per block column (1.5x halo) instead of 4.6x.
__global__ void __launch_bounds__(128) stem_kernel(const StemParams p) {
  extern __shared__ __align__(16) float s_stem[];
  constexpr int SW = kStemTW + 2, SH = kStemTH + 2;
  float* s_w = s_stem;                 // 27*c0
  float* s_b = s_w + 27 * p.c0;        // c0
  float* s_in = s_b + p.c0;            // [3 slots][SH][SW]
  const int tid = threadIdx.x;
  for (int i = tid; i < 27 * p.c0; i += blockDim.x) s_w[i] = p.w[i];
  for (int i = tid; i < p.c0; i += blockDim.x) s_b[i] = p.b[i];

  const int nbx = (p.tw + kStemTW - 1) / kStemTW, nby = (p.th + kStemTH - 1) / kStemTH;
  int blk = blockIdx.x;
  const int bx = blk % nbx; blk /= nbx;
  const int by = blk % nby; blk /= nby;
  const int t = blk;  // tile within batch
  const int tile = p.tile0 + t;
  const int tk = tile % p.gw, tj = (tile / p.gw) % p.gh, ti = tile / (p.gw * p.gh);
  // tile origin in padded coordinates minus the leading pad == image coordinates of tile voxel (0,0,0)
  const int oz = ti * p.ed - p.od, oy = tj * p.eh - p.oh, ox = tk * p.ew - p.ow;
  const int x0 = bx * kStemTW, y0 = by * kStemTH;

  // slice lz (tile-local; may be -1 or td: conv zero padding at the TILE border) -> slot (lz + 3) % 3
  auto load_slice = [&](int lz) {
    float* dst = s_in + ((lz + 3) % 3) * SH * SW;
    const bool zin = lz >= 0 && lz < p.td;
    const int gz = zin ? reflect_idx(oz + lz, p.VD) : 0;
    for (int i = tid; i < SH * SW; i += blockDim.x) {
      const int xx = i % SW, yy = i / SW;
      const int ly = y0 + yy - 1, lx = x0 + xx - 1;
      float v = 0.f;
      if (zin && ly >= 0 && ly < p.th && lx >= 0 && lx < p.tw) {
        int gy = oy + ly, gx = ox + lx;
        if (gy < 0 || gy >= p.VH) gy = reflect_idx(gy, p.VH);
        if (gx < 0 || gx >= p.VW) gx = reflect_idx(gx, p.VW);
        v = __ldg(p.vol + (static_cast<size_t>(gz) * p.VH + gy) * p.VW + gx);
      }
      dst[i] = v;
    }
  };
  load_slice(-1);
  load_slice(0);

  const int lx = (tid % (kStemTW / kStemXS)) * kStemXS, ly = tid / (kStemTW / kStemXS);
  const int x = x0 + lx, y = y0 + ly;
  const bool active = x < p.tw && y < p.th;
  for (int d = 0; d < p.td; ++d) {
    load_slice(d + 1);
    __syncthreads();
    if (active) {
      float in[9][kStemXS + 2];
#pragma unroll
      for (int kd = 0; kd < 3; ++kd) {
        const float* sl = s_in + ((d + kd + 2) % 3) * SH * SW;  // slice d + kd - 1
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
          for (int i = 0; i < kStemXS + 2; ++i) in[kd * 3 + kh][i] = sl[(ly + kh) * SW + lx + i];
      }
      const int CP = p.out_split ? 2 * p.c0 : p.c0;
      uint16_t* dst = reinterpret_cast<uint16_t*>(p.out) +
                      ((((static_cast<size_t>(t) * p.td + d) * p.th + y) * p.tw + x) * CP);
      for (int c8 = 0; c8 < p.c0; c8 += 8) {
        float acc[kStemXS][8];
#pragma unroll
        for (int v = 0; v < kStemXS; ++v)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[v][j] = s_b[c8 + j];
#pragma unroll
        for (int r = 0; r < 9; ++r)
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const float4 wa = *reinterpret_cast<const float4*>(s_w + (r * 3 + kw) * p.c0 + c8);
            const float4 wb = *reinterpret_cast<const float4*>(s_w + (r * 3 + kw) * p.c0 + c8 + 4);
            const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
            for (int v = 0; v < kStemXS; ++v)
#pragma unroll
              for (int j = 0; j < 8; ++j) acc[v][j] = fmaf(in[r][v + kw], wv[j], acc[v][j]);
          }
#pragma unroll
        for (int v = 0; v < kStemXS; ++v) {
          if (x + v >= p.tw) break;
          uint4 o;
          o.x = pack16(fmaxf(acc[v][0], 0.f), fmaxf(acc[v][1], 0.f), p.fmt);
          o.y = pack16(fmaxf(acc[v][2], 0.f), fmaxf(acc[v][3], 0.f), p.fmt);
          o.z = pack16(fmaxf(acc[v][4], 0.f), fmaxf(acc[v][5], 0.f), p.fmt);
          o.w = pack16(fmaxf(acc[v][6], 0.f), fmaxf(acc[v][7], 0.f), p.fmt);
          *reinterpret_cast<uint4*>(dst + static_cast<size_t>(v) * CP + c8) = o;
          if (p.out_split) {
            uint4 l;
            l.x = pack16_residual(fmaxf(acc[v][0], 0.f), fmaxf(acc[v][1], 0.f), o.x, p.fmt);
            l.y = pack16_residual(fmaxf(acc[v][2], 0.f), fmaxf(acc[v][3], 0.f), o.y, p.fmt);
            l.z = pack16_residual(fmaxf(acc[v][4], 0.f), fmaxf(acc[v][5], 0.f), o.z, p.fmt);
            l.w = pack16_residual(fmaxf(acc[v][6], 0.f), fmaxf(acc[v][7], 0.f), o.w, p.fmt);
            *reinterpret_cast<uint4*>(dst + static_cast<size_t>(v) * CP + p.c0 + c8) = l;
          }
        }
      }
    }
    __syncthreads();  // everyone is done with slice d-1 before its slot is overwritten by slice d+2
  }
}

// ---- tensor-path variant for C0 = 32 --------------------------------------------------------------------------------
// Same block structure (4 rows x 128 x, rolling 3-slice window), but the 27-tap x 32-channel contraction of a row runs
// on mma.sync.m16n8k16: M = 16 x-adjacent voxels, K = 27 taps padded to 32, N = 32 channels.  Input samples and
// weights are split into fp16 hi + lo pairs and hi*hi + lo*hi + hi*lo is accumulated in fp32, so the result matches the
// fp32 FMA kernel to ~2^-22 (the output is rounded to 16 bits anyway).  The window holds one word per sample,
// (hi | lo << 16), split once when the slice is loaded; an A-fragment register is two such words re-paired with PRMT.
// A warp owns one row: 8 m-tiles per slice, 24 MMAs each.  The n index of the B fragments is permuted (column n of
// n-tile nt = channel 8 (n / 2) + 2 nt + n % 2) so that a lane's C fragments are 8 consecutive channels: one 16-byte
// store per voxel, a warp store covers 8 voxels x 64 bytes contiguously.
__device__ __forceinline__ void stem_mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void stem_split(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(x0, x1);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ uint32_t stem_split1(float x) {  // (fp16(x) | fp16(x - fp16(x)) << 16)
  const __half h = __float2half_rn(x);
  const __half l = __float2half_rn(x - __half2float(h));
  return static_cast<uint32_t>(__half_as_ushort(h)) | (static_cast<uint32_t>(__half_as_ushort(l)) << 16);
}

__global__ void __launch_bounds__(128) stem_mma_kernel(const StemParams p) {
  constexpr int SW = kStemTW + 2, SH = kStemTH + 2, C0 = 32;
  __shared__ __align__(16) uint32_t s_in[3 * SH * SW];  // [3 slots][SH][SW] words (hi | lo << 16)
  const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5, g = lane >> 2, t = lane & 3;

  const int nbx = (p.tw + kStemTW - 1) / kStemTW, nby = (p.th + kStemTH - 1) / kStemTH;
  int blk = blockIdx.x;
  const int bx = blk % nbx; blk /= nbx;
  const int by = blk % nby; blk /= nby;
  const int tl = blk;  // tile within batch
  const int tile = p.tile0 + tl;
  const int tk = tile % p.gw, tj = (tile / p.gw) % p.gh, ti = tile / (p.gw * p.gh);
  const int oz = ti * p.ed - p.od, oy = tj * p.eh - p.oh, ox = tk * p.ew - p.ow;
  const int x0 = bx * kStemTW, y0 = by * kStemTH;

  // a thread fills the same (row, x) positions of the window for every slice: their in-slice offsets (reflect padding,
  // tile-border zeros) are computed once
  constexpr int NE = (SH * SW + 127) / 128;
  int eoff[NE];
#pragma unroll
  for (int e = 0; e < NE; ++e) {
    const int i = tid + 128 * e;
    eoff[e] = -1;
    if (i < SH * SW) {
      const int xx = i % SW, yy = i / SW;
      const int ly = y0 + yy - 1, lx = x0 + xx - 1;
      if (ly >= 0 && ly < p.th && lx >= 0 && lx < p.tw) {
        int gy = oy + ly, gx = ox + lx;
        if (gy < 0 || gy >= p.VH) gy = reflect_idx(gy, p.VH);
        if (gx < 0 || gx >= p.VW) gx = reflect_idx(gx, p.VW);
        eoff[e] = gy * p.VW + gx;
      }
    }
  }
  // slice lz (tile-local; -1 and td are the conv's zero padding at the TILE border) -> slot (lz + 3) % 3.  fetch() only
  // issues the global loads; commit() splits and stores them, one slice later, so the latency hides behind the MMAs
  auto fetch = [&](int lz, float (&v)[NE]) {
    const bool zin = lz >= 0 && lz < p.td;
    const float* src = p.vol + static_cast<size_t>(zin ? reflect_idx(oz + lz, p.VD) : 0) * p.VH * p.VW;
#pragma unroll
    for (int e = 0; e < NE; ++e) v[e] = (zin && eoff[e] >= 0) ? __ldg(src + eoff[e]) : 0.f;
  };
  auto commit = [&](int lz, const float (&v)[NE]) {
    uint32_t* dst = s_in + ((lz + 3) % 3) * SH * SW;
#pragma unroll
    for (int e = 0; e < NE; ++e)
      if (tid + 128 * e < SH * SW) dst[tid + 128 * e] = stem_split1(v[e]);
  };
  float pv[NE];
  fetch(-1, pv); commit(-1, pv);
  fetch(0, pv);  commit(0, pv);
  fetch(1, pv);

  // B fragments (weights) and the bias stay in registers for the whole block: k = tap (27 real + 5 zero),
  // column n = g of n-tile nt = channel 8 (g / 2) + 2 nt + g % 2
  uint32_t bh[2][4][2], bl[2][4][2];
  float bias[4][2];
#pragma unroll
  for (int s = 0; s < 2; ++s)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int k = 16 * s + 8 * h + 2 * t, co = 8 * (g >> 1) + 2 * nt + (g & 1);
        const float w0 = k < 27 ? __ldg(p.w + k * C0 + co) : 0.f, w1 = k + 1 < 27 ? __ldg(p.w + (k + 1) * C0 + co) : 0.f;
        stem_split(w0, w1, bh[s][nt][h], bl[s][nt][h]);
      }
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {   // C columns 2t, 2t+1 of n-tile nt = channels 8t + 2nt, 8t + 2nt + 1
    bias[nt][0] = __ldg(p.b + 8 * t + 2 * nt);
    bias[nt][1] = __ldg(p.b + 8 * t + 2 * nt + 1);
  }
  // this thread's 8 taps (A columns 2t, 2t+1, 2t+8, 2t+9 of both k-steps): slice selector and in-slice offset
  int tkd[8], trel[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int k = 16 * (i >> 2) + 8 * ((i >> 1) & 1) + 2 * t + (i & 1);
    const int kk = k < 27 ? k : 0;   // padded taps read a valid sample; their weights are zero
    tkd[i] = kk / 9;
    trel[i] = ((kk / 3) % 3) * SW + kk % 3;
  }
  const int ly = wrp, y = y0 + ly;
  int slot0 = 2;   // slot of slice d - 1 = (d + 2) % 3
  for (int d = 0; d < p.td; ++d) {
    commit(d + 1, pv);
    __syncthreads();
    fetch(d + 2, pv);   // in flight while slice d computes
    if (y < p.th) {
      int off[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        int sl = slot0 + tkd[i];
        sl = sl >= 3 ? sl - 3 : sl;
        off[i] = sl * SH * SW + trel[i] + ly * SW + g;
      }
      const int CP = p.out_split ? 2 * C0 : C0;   // channels per voxel row of the output tensor
      uint16_t* row_out = reinterpret_cast<uint16_t*>(p.out) +
                          (((static_cast<size_t>(tl) * p.td + d) * p.th + y) * p.tw + x0) * CP;
#pragma unroll 2
      for (int j = 0; j < kStemTW / 16; ++j) {
        if (x0 + 16 * j >= p.tw) break;
        // A fragments: rows g / g+8 = voxels 16j+g / 16j+g+8, columns = taps; words (hi | lo << 16) re-paired
        uint32_t ah[2][4], al[2][4];
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
          for (int h = 0; h < 2; ++h) {       // h: column half (taps 2t.. / 2t+8..)
            const int i0 = 4 * s + 2 * h;
            const uint32_t a0 = s_in[off[i0] + 16 * j], a1 = s_in[off[i0 + 1] + 16 * j];
            const uint32_t c0 = s_in[off[i0] + 16 * j + 8], c1 = s_in[off[i0 + 1] + 16 * j + 8];
            ah[s][2 * h] = __byte_perm(a0, a1, 0x5410);     al[s][2 * h] = __byte_perm(a0, a1, 0x7632);      // row g
            ah[s][2 * h + 1] = __byte_perm(c0, c1, 0x5410); al[s][2 * h + 1] = __byte_perm(c0, c1, 0x7632);  // row g+8
          }
        float acc[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          acc[nt][0] = acc[nt][2] = bias[nt][0];
          acc[nt][1] = acc[nt][3] = bias[nt][1];
        }
#pragma unroll
        for (int term = 0; term < 3; ++term)
#pragma unroll
          for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
              stem_mma16816(acc[nt], term == 1 ? al[s] : ah[s], term == 2 ? bl[s][nt][0] : bh[s][nt][0],
                            term == 2 ? bl[s][nt][1] : bh[s][nt][1]);
        // ReLU + 16-bit pack: this lane holds channels 8t .. 8t+7 of voxels g (acc[.][0..1]) and g + 8 (acc[.][2..3])
        uint4 o0, o1;
        o0.x = pack16(fmaxf(acc[0][0], 0.f), fmaxf(acc[0][1], 0.f), p.fmt);
        o0.y = pack16(fmaxf(acc[1][0], 0.f), fmaxf(acc[1][1], 0.f), p.fmt);
        o0.z = pack16(fmaxf(acc[2][0], 0.f), fmaxf(acc[2][1], 0.f), p.fmt);
        o0.w = pack16(fmaxf(acc[3][0], 0.f), fmaxf(acc[3][1], 0.f), p.fmt);
        o1.x = pack16(fmaxf(acc[0][2], 0.f), fmaxf(acc[0][3], 0.f), p.fmt);
        o1.y = pack16(fmaxf(acc[1][2], 0.f), fmaxf(acc[1][3], 0.f), p.fmt);
        o1.z = pack16(fmaxf(acc[2][2], 0.f), fmaxf(acc[2][3], 0.f), p.fmt);
        o1.w = pack16(fmaxf(acc[3][2], 0.f), fmaxf(acc[3][3], 0.f), p.fmt);
        const int xa = x0 + 16 * j + g, xb = xa + 8;
        if (xa < p.tw) *reinterpret_cast<uint4*>(row_out + static_cast<size_t>(16 * j + g) * CP + 8 * t) = o0;
        if (xb < p.tw) *reinterpret_cast<uint4*>(row_out + static_cast<size_t>(16 * j + g + 8) * CP + 8 * t) = o1;
        if (p.out_split) {   // lo plane: rounding residuals of the hi plane
          uint4 l0, l1;
          l0.x = pack16_residual(fmaxf(acc[0][0], 0.f), fmaxf(acc[0][1], 0.f), o0.x, p.fmt);
          l0.y = pack16_residual(fmaxf(acc[1][0], 0.f), fmaxf(acc[1][1], 0.f), o0.y, p.fmt);
          l0.z = pack16_residual(fmaxf(acc[2][0], 0.f), fmaxf(acc[2][1], 0.f), o0.z, p.fmt);
          l0.w = pack16_residual(fmaxf(acc[3][0], 0.f), fmaxf(acc[3][1], 0.f), o0.w, p.fmt);
          l1.x = pack16_residual(fmaxf(acc[0][2], 0.f), fmaxf(acc[0][3], 0.f), o1.x, p.fmt);
          l1.y = pack16_residual(fmaxf(acc[1][2], 0.f), fmaxf(acc[1][3], 0.f), o1.y, p.fmt);
          l1.z = pack16_residual(fmaxf(acc[2][2], 0.f), fmaxf(acc[2][3], 0.f), o1.z, p.fmt);
          l1.w = pack16_residual(fmaxf(acc[3][2], 0.f), fmaxf(acc[3][3], 0.f), o1.w, p.fmt);
          if (xa < p.tw) *reinterpret_cast<uint4*>(row_out + static_cast<size_t>(16 * j + g) * CP + C0 + 8 * t) = l0;
          if (xb < p.tw) *reinterpret_cast<uint4*>(row_out + static_cast<size_t>(16 * j + g + 8) * CP + C0 + 8 * t) = l1;
        }
      }
    }
    __syncthreads();  // everyone is done with slice d-1 before its slot is overwritten by slice d+2
    slot0 = slot0 == 2 ? 0 : slot0 + 1;
  }
}

int stem_launch(const StemParams& p, cudaStream_t st) {
  const int nbx = (p.tw + kStemTW - 1) / kStemTW, nby = (p.th + kStemTH - 1) / kStemTH;
  const long long blocks = static_cast<long long>(p.ntiles) * nby * nbx;
  static const bool no_mma = getenv("OAI_STEM_FP32") != nullptr;   // A/B switch: the fp32 FMA kernel
  if (p.c0 == 32 && !no_mma) {
    stem_mma_kernel<<<static_cast<unsigned>(blocks), 128, 0, st>>>(p);
    return launched("stem_mma_kernel");
  }
  const size_t smem = (27 * p.c0 + p.c0 + 3 * (kStemTH + 2) * (kStemTW + 2)) * sizeof(float);
  stem_kernel<<<static_cast<unsigned>(blocks), 128, smem, st>>>(p);
  return launched("stem_kernel");
}

// ------------------------------------------------------------------------------------------------ max pool 2x2x2
// in: [N,D,H,W,Cin8*8] of which the first C8*8 channels are pooled (a [hi | lo] tensor pooled through its hi plane:
// rounding is monotonic, so max(hi) is the hi of the max).  With split != 0 the tensor is [hi | lo] on both sides and
// the winner is chosen by hi + lo; its two halves are copied unchanged.
// packed 2 x 16-bit maximum without leaving the storage format (HMNMX2)
__device__ __forceinline__ uint32_t max16x2(uint32_t a, uint32_t b, int fmt) {
  if (fmt == 0) {
    const __half2 r = __hmax2(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
    return *reinterpret_cast<const uint32_t*>(&r);
  }
  const __nv_bfloat162 r = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
  return *reinterpret_cast<const uint32_t*>(&r);
}

// The common case (one 16-bit plane in and out): the eight 16-byte loads of a window are issued back to back and
// reduced with packed maxima -- ~100 instructions and 40 registers per thread.  The first version converted every
// value to fp32 (500 instructions, 64 registers, half occupancy) and reached 60 % of the HBM rate
// (profiles/r02_ncu_pool_stem.txt).
__global__ void __launch_bounds__(256) maxpool2_packed_kernel(const uint4* __restrict__ in, uint4* __restrict__ out,
                                                              int N, int D, int H, int W, int C8, int Cin8, int fmt) {
  const int Do = D / 2, Ho = H / 2, Wo = W / 2;
  const long long total = static_cast<long long>(N) * Do * Ho * Wo * C8;
  const size_t sW = Cin8, sH = static_cast<size_t>(W) * Cin8, sD = static_cast<size_t>(H) * W * Cin8;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    long long r = i;
    const int c = r % C8; r /= C8;
    const int w = r % Wo; r /= Wo;
    const int h = r % Ho; r /= Ho;
    const int d = r % Do; r /= Do;
    const uint4* p = in + (((static_cast<size_t>(r) * D + 2 * d) * H + 2 * h) * W + 2 * w) * Cin8 + c;
    const uint4 v0 = __ldg(p), v1 = __ldg(p + sW), v2 = __ldg(p + sH), v3 = __ldg(p + sH + sW);
    const uint4 v4 = __ldg(p + sD), v5 = __ldg(p + sD + sW), v6 = __ldg(p + sD + sH), v7 = __ldg(p + sD + sH + sW);
    uint4 o;
    o.x = max16x2(max16x2(max16x2(v0.x, v1.x, fmt), max16x2(v2.x, v3.x, fmt), fmt),
                  max16x2(max16x2(v4.x, v5.x, fmt), max16x2(v6.x, v7.x, fmt), fmt), fmt);
    o.y = max16x2(max16x2(max16x2(v0.y, v1.y, fmt), max16x2(v2.y, v3.y, fmt), fmt),
                  max16x2(max16x2(v4.y, v5.y, fmt), max16x2(v6.y, v7.y, fmt), fmt), fmt);
    o.z = max16x2(max16x2(max16x2(v0.z, v1.z, fmt), max16x2(v2.z, v3.z, fmt), fmt),
                  max16x2(max16x2(v4.z, v5.z, fmt), max16x2(v6.z, v7.z, fmt), fmt), fmt);
    o.w = max16x2(max16x2(max16x2(v0.w, v1.w, fmt), max16x2(v2.w, v3.w, fmt), fmt),
                  max16x2(max16x2(v4.w, v5.w, fmt), max16x2(v6.w, v7.w, fmt), fmt), fmt);
    out[i] = o;
  }
}

__global__ void __launch_bounds__(256) maxpool2_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int N,
                                                       int D, int H, int W, int C8, int Cin8, int split, int fmt) {
  const int Do = D / 2, Ho = H / 2, Wo = W / 2;
  const long long total = static_cast<long long>(N) * Do * Ho * Wo * C8;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    long long r = i;
    const int c = r % C8; r /= C8;
    const int w = r % Wo; r /= Wo;
    const int h = r % Ho; r /= Ho;
    const int d = r % Do; r /= Do;
    const int n = static_cast<int>(r);
    float m[8], mh[8], ml[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { m[j] = -INFINITY; mh[j] = 0.f; ml[j] = 0.f; }
#pragma unroll
    for (int dz = 0; dz < 2; ++dz)
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          const size_t off =
              ((((static_cast<size_t>(n) * D + 2 * d + dz) * H + 2 * h + dy) * W + 2 * w + dx) * Cin8) + c;
          const uint4 v = __ldg(in + off);
          const float2 a = unpack16(v.x, fmt), b = unpack16(v.y, fmt), cc = unpack16(v.z, fmt), e = unpack16(v.w, fmt);
          const float f[8] = {a.x, a.y, b.x, b.y, cc.x, cc.y, e.x, e.y};
          if (!split) {
#pragma unroll
            for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], f[j]);
          } else {
            const uint4 u = __ldg(in + off + C8);
            const float2 a2 = unpack16(u.x, fmt), b2 = unpack16(u.y, fmt), c2 = unpack16(u.z, fmt),
                         e2 = unpack16(u.w, fmt);
            const float g[8] = {a2.x, a2.y, b2.x, b2.y, c2.x, c2.y, e2.x, e2.y};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float t = f[j] + g[j];
              if (t > m[j]) { m[j] = t; mh[j] = f[j]; ml[j] = g[j]; }
            }
          }
        }
    uint4 o;
    if (!split) {
      o.x = pack16(m[0], m[1], fmt); o.y = pack16(m[2], m[3], fmt);
      o.z = pack16(m[4], m[5], fmt); o.w = pack16(m[6], m[7], fmt);
      out[i] = o;
    } else {
      const size_t ob = (i / C8) * (2 * C8) + c;
      o.x = pack16(mh[0], mh[1], fmt); o.y = pack16(mh[2], mh[3], fmt);
      o.z = pack16(mh[4], mh[5], fmt); o.w = pack16(mh[6], mh[7], fmt);
      out[ob] = o;
      o.x = pack16(ml[0], ml[1], fmt); o.y = pack16(ml[2], ml[3], fmt);
      o.z = pack16(ml[4], ml[5], fmt); o.w = pack16(ml[6], ml[7], fmt);
      out[ob + C8] = o;
    }
  }
}

int maxpool2_launch(const void* in, void* out, int N, int D, int H, int W, int C, int in_split, int out_split, int fmt,
                    cudaStream_t st) {
  const long long total = static_cast<long long>(N) * (D / 2) * (H / 2) * (W / 2) * (C / 8);
  long long blocks = (total + 255) / 256;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  if (!out_split) {
    maxpool2_packed_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(
        static_cast<const uint4*>(in), static_cast<uint4*>(out), N, D, H, W, C / 8, (in_split ? 2 : 1) * (C / 8), fmt);
    return launched("maxpool2_packed_kernel");
  }
  maxpool2_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(static_cast<const uint4*>(in),
                                                                   static_cast<uint4*>(out), N, D, H, W, C / 8,
                                                                   (in_split ? 2 : 1) * (C / 8), out_split, fmt);
  return launched("maxpool2_kernel");
}

// ------------------------------------------------------------------------------------------------ head

__global__ void __launch_bounds__(128) head_kernel(const HeadParams p) {
  __shared__ float s_w[8 * 64 + 8];
  for (int i = threadIdx.x; i < p.ncls * p.C; i += blockDim.x) s_w[i] = p.w[i];
  for (int i = threadIdx.x; i < p.ncls; i += blockDim.x) s_w[p.ncls * p.C + i] = p.b[i];
  __syncthreads();
  // one thread per interior voxel of a tile: interior = [od, od+ed) x [oh, oh+eh) x [ow, ow+ew)
  const long long per_tile = static_cast<long long>(p.ed) * p.eh * p.ew;
  const long long total = per_tile * p.ntiles;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    long long r = i;
    const int x = r % p.ew; r /= p.ew;
    const int y = r % p.eh; r /= p.eh;
    const int z = r % p.ed; r /= p.ed;
    const int t = static_cast<int>(r);
    const int tile = p.tile0 + t;
    const int tk = tile % p.gw, tj = (tile / p.gw) % p.gh, ti = tile / (p.gw * p.gh);
    const int gz = ti * p.ed + z, gy = tj * p.eh + y, gx = tk * p.ew + x;  // image coordinates
    if (gz >= p.VD || gy >= p.VH || gx >= p.VW) continue;                 // trimmed (image_transforms.py:504)
    const bool shell = gz < p.cz || gz >= p.VD - p.cz || gy < p.cy || gy >= p.VH - p.cy || gx < p.cx ||
                       gx >= p.VW - p.cx;
    const size_t vox = (((static_cast<size_t>(t) * p.td + z + p.od) * p.th + y + p.oh) * p.tw + x + p.ow);
    const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.act) + vox * p.C);
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    if (!shell) {
      for (int c8 = 0; c8 < p.C / 8; ++c8) {
        const uint4 v = __ldg(src + c8);
        const float2 a = unpack16(v.x, p.fmt), b = unpack16(v.y, p.fmt), c = unpack16(v.z, p.fmt),
                     e = unpack16(v.w, p.fmt);
        const float f[8] = {a.x, a.y, b.x, b.y, c.x, c.y, e.x, e.y};
        for (int k = 0; k < p.ncls; ++k) {
          const float* wk = s_w + k * p.C + c8 * 8;
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[k] = fmaf(f[j], wk[j], acc[k]);
        }
      }
    }
    const size_t o = (static_cast<size_t>(gz) * p.VH + gy) * p.VW + gx;
    const size_t cls_stride = static_cast<size_t>(p.VD) * p.VH * p.VW;
    for (int k = 0; k < p.ncls; ++k) {
      float v = 0.f;
      if (!shell) {
        const float logit = acc[k] + s_w[p.ncls * p.C + k];
        v = p.out_mode == 2 ? logit : 1.f / (1.f + __expf(-logit));
        if (p.out_mode == 1) v = v > 0.5f ? 1.f : 0.f;
      }
      p.out[k * cls_stride + o] = v;
    }
  }
}

int head_launch(const HeadParams& p, cudaStream_t st) {
  const long long total = static_cast<long long>(p.ed) * p.eh * p.ew * p.ntiles;
  long long blocks = (total + 127) / 128;
  const long long cap = static_cast<long long>(num_sms()) * 32;
  if (blocks > cap) blocks = cap;
  head_kernel<<<static_cast<unsigned>(blocks), 128, 0, st>>>(p);
  return launched("head_kernel");
}

}  // namespace oai
