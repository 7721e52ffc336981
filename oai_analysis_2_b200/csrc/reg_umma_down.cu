// tallUNet2 down step on the 5th-generation tensor cores: Conv3d(k3, s2, p1) on leaky_relu(x) + the avg-pool residual
// (icon_registration networks.UNet2.forward, down path) as a tcgen05 implicit GEMM with split-fp16 operands
// (hi*hi + lo*hi + hi*lo, fp32 accumulation in TMEM: fp32-level accuracy), for the levels with 16+ input channels.
//
//   * out[o] = sum_k in[2o - 1 + k] * w[k]: per axis tap k reads input parity r = (k + 1) & 1 at lattice offset
//     s = (k == 0 ? -1 : 0).  A pre-pass (reg_split_s2d_kernel) rewrites the layer input once as eight parity planes,
//     channels-last fp16 hi / lo (space-to-depth; positions past an odd extent are zeros), so every tap's A operand is
//     a plain TMA box of one parity plane -- [128 output lattice points x 16 channels], 32-byte rows, SWIZZLE_32B --
//     and the conv's zero padding is the box's out-of-bounds fill;
//   * M tile = 128 output points of one z-slice (16 rows x 8 columns), N = Cn output channels (32 or 64), K loop =
//     (16-channel chunk, tap): one stage = the tap's hi and lo boxes + the tap's weight rows (hi, lo), three MMAs;
//   * a unit's accumulator is Cn columns of one TMEM half; consecutive units of a CTA alternate halves so the epilogue
//     (2^-wexp, bias, avg_pool3d(2, ceil_mode) residual on the trailing channels, planar fp32 stores) runs under the
//     next unit's MMAs;
//   * warp roles: 0 = activation producer (TMA), 1 = weight producer, 2 = MMA issuer + TMEM owner, 4..7 = epilogue.
#include "api_common.h"
#include "ptx.cuh"
#include "reg_kernels.cuh"

#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdlib>

namespace oai {

namespace {

constexpr int kDTX = 8, kDTY = 16;               // output tile: 128 lattice points of one slice
constexpr int kDRowB = 32;                       // 16 fp16 channels
constexpr int kDBox = kDTX * kDTY * kDRowB;      // 4096 bytes per box
constexpr int kDStages = 8;
constexpr int kDThreads = 256;
constexpr uint32_t kDLayoutSw32 = 6;

struct DownParams {
  const float* in;              // raw layer input (planar fp32): residual source
  long long in_nstride, in_cstride;
  int cin, Di, Hi, Wi, N;
  const uint8_t* wumma;         // [nsplit][nchunks][27] blocks of 2 x (Cn rows x 32 B), pre-swizzled
  const float* bias;
  float* out;
  long long out_nstride, out_cstride;
  int cout, Do, Ho, Wo;
  int Cn, nsplit, nchunks;
  int nx, ny, nunits;
  uint32_t bblock;              // bytes of one weight block (2 * Cn * 32)
  uint32_t stage_bytes;         // 2 boxes + weight block, 1024-aligned
  float inv, out_scale;
};

struct DUnit {
  int n, h, z, y0, x0;
};

__device__ __forceinline__ DUnit decode_dunit(const DownParams& p, int u) {
  DUnit ui;
  ui.x0 = (u % p.nx) * kDTX; u /= p.nx;
  ui.y0 = (u % p.ny) * kDTY; u /= p.ny;
  ui.z = u % p.Do; u /= p.Do;
  ui.h = u % p.nsplit;
  ui.n = u / p.nsplit;
  return ui;
}

__device__ __forceinline__ float leaky_d(float v) { return v > 0.f ? v : 0.01f * v; }

__device__ __forceinline__ void st_global_256d(void* p, const uint32_t* v) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
               "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}

}  // namespace

// Layer input [N][cin] planes (explicit strides) -> xs [N][8 parity classes][cin / 16][2 (hi, lo)][Dc][Hc][Wc][16] fp16
// (Dc = ceil(D / 2) ...; class = rz*4 + ry*2 + rx holds in[2zc + rz][2yc + ry][2xc + rx], zeros past the extent):
// leaky_relu, then the hi / lo split; one thread moves 16 channels of one lattice point.
__global__ void __launch_bounds__(256) reg_split_s2d_kernel(const float* __restrict__ in, long long in_nstride,
                                                            long long in_cstride, int N, int cin, int Di, int Hi,
                                                            int Wi, uint4* __restrict__ xs) {
  const int ng = cin / 16;
  const int Dc = (Di + 1) / 2, Hc = (Hi + 1) / 2, Wc = (Wi + 1) / 2;
  const long long cvol = static_cast<long long>(Dc) * Hc * Wc;
  const long long total = static_cast<long long>(N) * 8 * cvol * ng;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    long long r = i;
    const int xc = static_cast<int>(r % Wc); r /= Wc;
    const int yc = static_cast<int>(r % Hc); r /= Hc;
    const int zc = static_cast<int>(r % Dc); r /= Dc;
    const int j = static_cast<int>(r % ng); r /= ng;
    const int cls = static_cast<int>(r % 8);
    const long long n = r / 8;
    const int z = 2 * zc + (cls >> 2), y = 2 * yc + ((cls >> 1) & 1), x = 2 * xc + (cls & 1);
    uint32_t hi[8], lo[8];
    if (z < Di && y < Hi && x < Wi) {
      const float* src = in + n * in_nstride + (16ll * j) * in_cstride + (static_cast<long long>(z) * Hi + y) * Wi + x;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float x0 = leaky_d(__ldg(src + (2 * k) * in_cstride)), x1 = leaky_d(__ldg(src + (2 * k + 1) * in_cstride));
        const __half2 h = __floats2half2_rn(x0, x1);
        const float2 hf = __half22float2(h);
        const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
        hi[k] = *reinterpret_cast<const uint32_t*>(&h);
        lo[k] = *reinterpret_cast<const uint32_t*>(&l);
      }
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) hi[k] = lo[k] = 0u;
    }
    // chunk-planar, the hi and the lo plane of a (sample, class, chunk) adjacent: one TMA box fetches both
    const long long o = ((((((n * 8 + cls) * ng + j) * 2) * Dc + zc) * Hc + yc) * Wc + xc) * 2;
    st_global_256d(xs + o, hi);
    st_global_256d(xs + o + cvol * 2, lo);
  }
}

// w [cin][27][cout_pad] fp32 (tap = (kd*3+kh)*3+kw) -> weight blocks: block (h, c, t) = co split h, 16-channel chunk c,
// tap t; inside, plane 0 = rn16(w * 2^wexp), plane 1 = rn16 of the remainder; row n = output channel h*Cn + n, element
// k = channel within the chunk, SWIZZLE_32B pattern.
__global__ void reg_pack_conv3_umma_kernel(const float* __restrict__ w, int cin, int cout, int cout_pad, int Cn,
                                           int wexp, __half* __restrict__ dst) {
  const int nchunks = cin / 16, nsplit = cout / Cn;
  const long long total = static_cast<long long>(nsplit) * nchunks * 27 * Cn * 16;
  const float sc = exp2f(static_cast<float>(wexp));
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    long long r = i;
    const int k = static_cast<int>(r % 16); r /= 16;
    const int n = static_cast<int>(r % Cn); r /= Cn;
    const int t = static_cast<int>(r % 27); r /= 27;
    const int c = static_cast<int>(r % nchunks);
    const int h = static_cast<int>(r / nchunks);
    const float v = w[(static_cast<size_t>(c * 16 + k) * 27 + t) * cout_pad + h * Cn + n] * sc;
    const __half hi = __float2half_rn(v);
    const __half lo = __float2half_rn(v - __half2float(hi));
    const size_t block = (static_cast<size_t>(h) * nchunks + c) * 27 + t;
    const size_t off = static_cast<size_t>(n) * 16 + ((((k >> 3) ^ ((n >> 2) & 1))) << 3) + (k & 7);
    __half* b = dst + block * (2 * static_cast<size_t>(Cn) * 16);
    b[off] = hi;
    b[static_cast<size_t>(Cn) * 16 + off] = lo;
  }
}

__global__ void __launch_bounds__(kDThreads, 1)
conv3s2_umma_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ DownParams p) {
  extern __shared__ uint8_t down_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(down_smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(kDStages) * p.stage_bytes);
  uint64_t* full = bars;               // [kDStages]
  uint64_t* empty = bars + kDStages;   // [kDStages]
  uint64_t* acc_full = bars + 2 * kDStages;       // [2]
  uint64_t* acc_empty = bars + 2 * kDStages + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kDStages + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Cn = p.Cn;
  const int Dc = (p.Di + 1) / 2;
  (void)Dc;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kDStages; ++i) {
      mbar_init(&full[i], 2);    // the activation and the weight producer
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 4);
    }
    fence_mbar_init();
    tma_prefetch_desc(&tm_x);
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int nst = p.nchunks * 27;   // stages per unit

  if (warp == 0) {
    // ------------------------------------------------------------ activation producer: per (chunk, tap) one box holding
    // the tap's hi and lo tiles.  (One thread issuing two boxes and the weight rows per stage was the kernel's pace: a
    // bulk-copy instruction takes a few hundred nanoseconds to issue, the stages are 96 tensor clocks long.)
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0;
      for (int u = blockIdx.x; u < p.nunits; u += gridDim.x) {
        const DUnit ui = decode_dunit(p, u);
        for (int c = 0; c < p.nchunks; ++c)
          for (int t = 0; t < 27; ++t) {
            const int kz = t / 9, ky = (t / 3) % 3, kx = t % 3;
            // tap k reads parity (k + 1) & 1 at lattice offset (k == 0 ? -1 : 0)
            const int cls = (((kz + 1) & 1) << 2) | (((ky + 1) & 1) << 1) | ((kx + 1) & 1);
            const int cz = ui.z - (kz == 0), cy = ui.y0 - (ky == 0), cx = ui.x0 - (kx == 0);
            mbar_wait(&empty[st], ph ^ 1u, 100 + st);
            mbar_arrive_expect_tx(&full[st], 2u * kDBox);
            tma_load_5d(smem + static_cast<size_t>(st) * p.stage_bytes, &tm_x, &full[st], 0, cx, cy, cz,
                        ((ui.n * 8 + cls) * p.nchunks + c) * 2);
            if (++st == kDStages) { st = 0; ph ^= 1u; }
          }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ weight producer: the tap's rows (hi, lo)
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0;
      for (int u = blockIdx.x; u < p.nunits; u += gridDim.x) {
        const DUnit ui = decode_dunit(p, u);
        const uint8_t* wsrc = p.wumma + static_cast<size_t>(ui.h) * nst * p.bblock;
        for (int s = 0; s < nst; ++s) {
          mbar_wait(&empty[st], ph ^ 1u, 200 + st);
          mbar_arrive_expect_tx(&full[st], p.bblock);
          bulk_load(smem + static_cast<size_t>(st) * p.stage_bytes + 2 * kDBox, wsrc + static_cast<size_t>(s) * p.bblock,
                    p.bblock, &full[st]);
          if (++st == kDStages) { st = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------ MMA issuer
    const bool leader = elect_one();
    const uint32_t smem_lo = ((smem_u32(smem) & 0x3FFFF) >> 4) | (1u << 16);
    const uint32_t desc_hi = ((8 * kDRowB) >> 4) | (1u << 14) | (kDLayoutSw32 << 29);   // SBO = 8 contiguous rows
    const uint32_t idesc = umma_idesc_f16(128, Cn, 0);
    const uint32_t stage_step = p.stage_bytes >> 4;
    const uint32_t a_lo_off = kDBox >> 4, b_off = (2 * kDBox) >> 4, b_lo_off = (static_cast<uint32_t>(Cn) * kDRowB) >> 4;
    int st = 0, it = 0;
    uint32_t ph = 0;
    for (int u = blockIdx.x; u < p.nunits; u += gridDim.x, ++it) {
      const int half = it & 1;
      const uint32_t use = (it >> 1) & 1u;
      const uint32_t d = tmem_base + static_cast<uint32_t>(half * 256);
      mbar_wait(&acc_empty[half], use ^ 1u, 500 + half);
      tc_fence_after();
      for (int s = 0; s < nst; ++s) {
        mbar_wait(&full[st], ph, 300 + st);
        tc_fence_after();
        const uint32_t a0 = smem_lo + static_cast<uint32_t>(st) * stage_step, b0 = a0 + b_off;
        if (leader) {
          umma_f16_ss_lohi(d, a0, b0, desc_hi, idesc, s ? 1u : 0u);            // hi * w_hi (first stage overwrites)
          umma_f16_ss_lohi(d, a0 + a_lo_off, b0, desc_hi, idesc, 1u);          // lo * w_hi
          umma_f16_ss_lohi(d, a0, b0 + b_lo_off, desc_hi, idesc, 1u);          // hi * w_lo
          umma_commit(&empty[st]);
        }
        if (++st == kDStages) { st = 0; ph ^= 1u; }
      }
      if (leader) umma_commit(&acc_full[half]);
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue
    const int q = warp & 3, m = q * 32 + lane;
    const int ty = m >> 3, tx = m & 7;
    const int cfront = p.cout - p.cin;   // the pooled channels are zero-padded in front
    int it = 0;
    for (int u = blockIdx.x; u < p.nunits; u += gridDim.x, ++it) {
      const DUnit ui = decode_dunit(p, u);
      const int half = it & 1;
      const uint32_t use = (it >> 1) & 1u;
      mbar_wait(&acc_full[half], use, 600 + half);
      tc_fence_after();
      const int y = ui.y0 + ty, x = ui.x0 + tx, z = ui.z;
      const bool inside = y < p.Ho && x < p.Wo;
      // avg_pool3d(2, ceil_mode=True): the window is clipped to the volume and divided by the samples it holds.  The
      // eight taps are read unconditionally (offsets clamped into the window, weights zero for the clipped ones) so the
      // loads of a channel -- and of the following channels -- are in flight together.
      const int nz = min(2, p.Di - 2 * z), ny = min(2, p.Hi - 2 * y), nxw = min(2, p.Wi - 2 * x);
      const float rcnt = inside ? 1.f / static_cast<float>(nz * ny * nxw) : 0.f;
      int roff[8];
      float rwgt[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int dz = k >> 2, dy = (k >> 1) & 1, dx = k & 1;
        const bool ok = inside && dz < nz && dy < ny && dx < nxw;
        roff[k] = ok ? (dz * p.Hi + dy) * p.Wi + dx : 0;
        rwgt[k] = ok ? rcnt : 0.f;
      }
      const float* raw_n = p.in + ui.n * p.in_nstride +
                           (inside ? (static_cast<long long>(2 * z) * p.Hi + 2 * y) * p.Wi + 2 * x : 0);
      float* out_n = p.out + ui.n * p.out_nstride + (static_cast<long long>(z) * p.Ho + y) * p.Wo + x;
      const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(half * 256);
      for (int cb = 0; cb < Cn; cb += 32) {
        uint32_t v[32];
        tmem_ld_32x32(t_lane + static_cast<uint32_t>(cb), v);
        tmem_ld_wait();
        if (cb + 32 >= Cn) {   // the accumulator is in registers: hand the TMEM half back before the stores
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[half]);
        }
        if (!inside) continue;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int co = ui.h * Cn + cb + i;
          float r = __uint_as_float(v[i]) * p.inv + __ldg(p.bias + co);
          const int cs = co - cfront;
          if (cs >= 0) {
            const float* pl = raw_n + cs * p.in_cstride;
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) s = fmaf(__ldg(pl + roff[k]), rwgt[k], s);
            r += s;
          }
          out_n[co * p.out_cstride] = r * p.out_scale;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------- host side
namespace {

int down_cn(int cout) { return cout >= 64 ? 64 : 32; }

typedef CUresult (*EncodeTiledFnD)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_s2d_tmap(CUtensorMap* tm, void* base, int cin, int Wc, int Hc, int Dc, int planes) {
  static EncodeTiledFnD fn = nullptr;
  if (!fn) {
    void* q = nullptr;
    cudaDriverEntryPointQueryResult r;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &r) == cudaSuccess &&
        r == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFnD>(q);
  }
  if (!fn) return fail("conv3s2_umma: cuTensorMapEncodeTiled not available from the driver");
  (void)cin;   // every plane is one 16-channel chunk of one parity class
  cuuint64_t dims[5] = {16, (cuuint64_t)Wc, (cuuint64_t)Hc, (cuuint64_t)Dc, (cuuint64_t)planes};
  cuuint64_t strides[4] = {32, (cuuint64_t)Wc * 32, (cuuint64_t)Hc * Wc * 32, (cuuint64_t)Dc * Hc * Wc * 32};
  cuuint32_t box[5] = {16, kDTX, kDTY, 1, 2};   // the hi and the lo plane in one box
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, base, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail("conv3s2_umma: cuTensorMapEncodeTiled failed with %d (cin=%d W=%d H=%d D=%d planes=%d)", (int)r, cin, Wc,
                Hc, Dc, planes);
  return 0;
}

bool down_umma_disabled() {
  const char* e = getenv("OAI_B200_CONV3_UMMA");
  return e && e[0] == '0';
}

}  // namespace

bool conv3_umma_eligible(const Conv3Params& p) {
  return p.stride == 2 && p.leaky_in && p.residual && p.cin % 16 == 0 && p.cin >= 16 &&
         (p.cout == 32 || p.cout % 64 == 0) && p.cout >= p.cin && p.cout <= 512 && p.Wo >= 8 && p.Ho >= 8 &&
         !down_umma_disabled();
}

size_t conv3_umma_wbytes(int cin, int cout) {
  if (cin % 16 || !(cout == 32 || cout % 64 == 0)) return 0;
  return static_cast<size_t>(cin / 16) * 27 * 2 * cout * kDRowB;
}

size_t conv3_umma_workspace(const Conv3Params& p) {
  const size_t cvol = static_cast<size_t>((p.Di + 1) / 2) * ((p.Hi + 1) / 2) * ((p.Wi + 1) / 2);
  return static_cast<size_t>(p.N) * 8 * cvol * p.cin * 4;
}

int reg_pack_conv3_umma_launch(const float* w, int cin, int cout, int cout_pad, int wexp, void* dst, cudaStream_t st) {
  const int Cn = down_cn(cout);
  const long long total = static_cast<long long>(cout / Cn) * (cin / 16) * 27 * Cn * 16;
  const unsigned blocks = static_cast<unsigned>(std::min<long long>((total + 255) / 256, 148 * 16));
  reg_pack_conv3_umma_kernel<<<blocks, 256, 0, st>>>(w, cin, cout, cout_pad, Cn, wexp, static_cast<__half*>(dst));
  return launched("reg_pack_conv3_umma_kernel");
}

int conv3_umma_launch(const Conv3Params& p, cudaStream_t st) {
  const int Dc = (p.Di + 1) / 2, Hc = (p.Hi + 1) / 2, Wc = (p.Wi + 1) / 2;
  if (p.xsplit_bytes < conv3_umma_workspace(p) || (reinterpret_cast<uintptr_t>(p.xsplit) & 127))
    return fail("conv3s2_umma: workspace of %zu bytes, 128-byte aligned, required", conv3_umma_workspace(p));
  if (p.Do != Dc || p.Ho != Hc || p.Wo != Wc) return fail("conv3s2_umma: output dims must be ceil(input / 2)");
  {
    const long long total = static_cast<long long>(p.N) * 8 * Dc * Hc * Wc * (p.cin / 16);
    const unsigned blocks = static_cast<unsigned>(std::min<long long>((total + 255) / 256, 148 * 32));
    reg_split_s2d_kernel<<<blocks, 256, 0, st>>>(p.in, p.in_nstride, p.in_cstride, p.N, p.cin, p.Di, p.Hi, p.Wi,
                                                 static_cast<uint4*>(p.xsplit));
    if (int rc = launched("reg_split_s2d_kernel")) return rc;
  }
  DownParams q{};
  q.in = p.in; q.in_nstride = p.in_nstride; q.in_cstride = p.in_cstride;
  q.cin = p.cin; q.Di = p.Di; q.Hi = p.Hi; q.Wi = p.Wi; q.N = p.N;
  q.wumma = static_cast<const uint8_t*>(p.wumma);
  q.bias = p.bias;
  q.out = p.out; q.out_nstride = p.out_nstride; q.out_cstride = p.out_cstride;
  q.cout = p.cout; q.Do = p.Do; q.Ho = p.Ho; q.Wo = p.Wo;
  q.Cn = down_cn(p.cout);
  q.nsplit = p.cout / q.Cn;
  q.nchunks = p.cin / 16;
  q.bblock = 2u * q.Cn * kDRowB;
  q.stage_bytes = (2u * kDBox + q.bblock + 1023u) & ~1023u;
  q.nx = (p.Wo + kDTX - 1) / kDTX; q.ny = (p.Ho + kDTY - 1) / kDTY;
  const long long nunits = static_cast<long long>(p.N) * q.nsplit * p.Do * q.ny * q.nx;
  if (nunits > 0x7fffffffLL) return fail("conv3s2_umma: too many units");
  q.nunits = static_cast<int>(nunits);
  q.inv = exp2f(static_cast<float>(-p.wexp));
  q.out_scale = p.out_scale;
  CUtensorMap tm;
  if (int rc = make_s2d_tmap(&tm, p.xsplit, p.cin, Wc, Hc, Dc, p.N * 8 * (p.cin / 16) * 2)) return rc;
  const size_t smem = 1024 + static_cast<size_t>(kDStages) * q.stage_bytes + 256;
  static size_t configured[64] = {0};   // the attribute is per device
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || smem > configured[dev]) {
    if (int rc = check_cuda(cudaFuncSetAttribute(conv3s2_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 static_cast<int>(smem)),
                            "conv3s2_umma: cudaFuncSetAttribute"))
      return rc;
    if (dev >= 0 && dev < 64) configured[dev] = smem;
  }
  const int grid = std::min(q.nunits, num_sms());
  conv3s2_umma_kernel<<<grid, kDThreads, smem, st>>>(tm, q);
  return launched("conv3s2_umma_kernel");
}

}  // namespace oai
