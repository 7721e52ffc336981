// tallUNet2 up step on the 5th-generation tensor cores: ConvTranspose3d(k4, s2, p1) + trilinear-upsample residual +
// BatchNorm(eval) + crop (icon_registration networks.UNet2.forward, up path) as a tcgen05 implicit GEMM with
// split-fp16 operands (hi*hi + lo*hi + hi*lo, fp32 accumulation in TMEM: fp32-level accuracy).
//
// Formulation (input-lattice stationary, parity classes and d-shifts stacked along N):
//   * out[2i + r] (r in {0,1}^3, the 8 parity classes of lattice point i) = sum over input shifts s in {-1,0,1}^3 of
//     in[i + s] * w[k], k = r + 1 - 2s per axis (valid when 0 <= k <= 3): a class uses 2 of the 3 shifts per axis;
//   * M tile = 128 lattice points of one z-slice (16 rows x 8 columns); the A operand of an in-plane shift (sy, sx) is
//     the SAME halo'd 18 x 10 box of 32-byte channel rows (16 fp16 channels, SWIZZLE_32B) addressed (sy*10 + sx) rows
//     further -- 8-row groups are the 8 x-adjacent points of one lattice row, SBO = 10 rows;
//   * accumulators of one lattice slice = 8 classes x Cn output channels (columns class*Cn + co); consecutive lattice
//     slices are adjacent in TMEM, so input slice zi updates [upper half of slice zi-1 | slice zi | lower half of
//     slice zi+1] with ONE MMA of N = 16*Cn columns against weight rows stored in that order (invalid (class, shift)
//     pairs are zero rows);
//   * a unit = R lattice slices of one tile (R * 8 * Cn = 256 columns = half of TMEM); consecutive units of a CTA
//     alternate between the two halves, so the epilogue of unit i overlaps the MMAs of unit i+1;
//   * K loop: 16-channel chunks; per chunk the unit's R+2 input slices (hi and lo boxes) arrive by TMA and stay while
//     the nine in-plane shifts stream their weight blocks (cp.async.bulk, pre-swizzled) through a ring;
//   * warp roles: 0 = TMA producer, 1 = weight producer, 2 = MMA issuer + TMEM owner, 4..11 = two epilogue groups
//     (TMEM -> registers -> 2^-wexp, bias, upsampled residual, BatchNorm -> planar fp32 stores, x-parity pairs as
//     8-byte stores).
#include "api_common.h"
#include "ptx.cuh"
#include "reg_kernels.cuh"

#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdlib>

namespace oai {

namespace {

constexpr int kTX = 8, kTY = 16;                 // lattice tile: 128 points of one slice
constexpr int kBX = kTX + 2, kBY = kTY + 2;      // halo'd TMA box
constexpr int kRowB = 32;                        // 16 fp16 channels
constexpr int kBoxBytes = kBX * kBY * kRowB;     // 5760
constexpr int kSlotBytes = 6144;                 // box slot (1024-aligned)
constexpr int kHalfCols = 256;
constexpr int kUmmaThreads = 384;               // warps 0-2: producers + MMA issuer, 3: idle, 4-11: two epilogue groups
constexpr uint32_t kLayoutSw32 = 6;              // UMMA shared-memory descriptor layout type: SWIZZLE_32B
constexpr int kRawX = 16, kRawBytes = 8 * 3 * kBY * kRawX * 4;   // raw residual box: 8 channels x 3 x 18 x 16 floats

struct UmmaParams {
  const float* in;              // raw layer input (planar fp32): residual source
  long long in_nstride, in_cstride;
  int cin, Di, Hi, Wi, N;
  const uint8_t* wumma;         // [nsplit][nchunks][9] blocks of 2 x (16*Cn rows x 32 B), pre-swizzled
  const float *bias, *bn_scale, *bn_shift;
  float* out;
  long long out_nstride, out_cstride;
  int cout, Do, Ho, Wo;
  int Cn, R, nsplit, nchunks, nb;
  int nx, ny, nzg, nunits;
  uint32_t bblock;              // bytes of one weight block
  uint32_t abuf_bytes;          // (R + 2) * 2 * kSlotBytes
  float inv;                    // 2^-wexp
  int stage_raw;                // the residual's raw neighbourhood arrives by TMA (planar input with dense planes)
  int raw_ratio;                // planes per sample of the raw input (in_nstride / in_cstride)
  int debug;                    // development switches (OAI_CONVT4_DEBUG): 1 no residual, 2 no MMAs, 8 no stores
};

struct Unit {
  int n, h, z0, rd, y0, x0;
};

__device__ __forceinline__ Unit decode_unit(const UmmaParams& p, int u) {
  Unit ui;
  ui.x0 = (u % p.nx) * kTX; u /= p.nx;
  ui.y0 = (u % p.ny) * kTY; u /= p.ny;
  ui.z0 = (u % p.nzg) * p.R; u /= p.nzg;
  ui.h = u % p.nsplit;
  ui.n = u / p.nsplit;
  ui.rd = min(p.R, p.Di - ui.z0);
  return ui;
}

__device__ __forceinline__ float leaky_f(float v) { return v > 0.f ? v : 0.01f * v; }

// one 32-byte global store (sm_100: STG.256); p must be 32-byte aligned
__device__ __forceinline__ void st_global_256(void* p, const uint32_t* v) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
               "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}

}  // namespace

// Layer input [N][cin] planes (explicit strides) -> xcl [2 (hi, lo)][N][cin / 16][vol][16] fp16 (channels-last inside a
// 16-channel chunk, chunks planar: the x-adjacent 32-byte rows a TMA box fetches are contiguous in memory, so a box is
// a few dozen line requests instead of 180 sector requests): leaky_relu, then the hi / lo split; one thread moves the 16
// channels of one voxel and chunk (sixteen coalesced plane reads, one whole 32-byte sector written per precision plane).
__global__ void __launch_bounds__(256) reg_split_cl_kernel(const float* __restrict__ in, long long in_nstride,
                                                           long long in_cstride, int N, int cin, long long vol,
                                                           uint4* __restrict__ xcl) {
  const int ng = cin / 16;
  const long long total = static_cast<long long>(N) * ng * vol;
  const long long plane = static_cast<long long>(N) * vol * ng * 2;   // uint4 per precision plane
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long v = i % vol;
    const int j = static_cast<int>((i / vol) % ng);
    const long long n = i / (vol * ng);
    const float* src = in + n * in_nstride + (16ll * j) * in_cstride + v;
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float x0 = leaky_f(__ldg(src + (2 * k) * in_cstride)), x1 = leaky_f(__ldg(src + (2 * k + 1) * in_cstride));
      const __half2 h = __floats2half2_rn(x0, x1);
      const float2 hf = __half22float2(h);
      const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
      hi[k] = *reinterpret_cast<const uint32_t*>(&h);
      lo[k] = *reinterpret_cast<const uint32_t*>(&l);
    }
    const long long o = ((n * ng + j) * vol + v) * 2;   // chunk-planar: x-adjacent rows of a chunk are contiguous
    st_global_256(xcl + o, hi);           // one 32-byte store per plane: the sector is written whole
    st_global_256(xcl + plane + o, lo);
  }
}

// w [cin][64][cout] fp32 (tap = (kd*4+kh)*4+kw) -> the kernel's weight blocks: block (h, c, j) = co split h, 16-channel
// chunk c, in-plane shift j = (sy+1)*3 + (sx+1); inside, plane 0 = rn16(w * 2^wexp), plane 1 = rn16 of the remainder;
// row n = slot * Cn + co' with slot 0..3 = (s_d = +1, classes 4..7), 4..11 = (s_d = 0, classes 0..7), 12..15 =
// (s_d = -1, classes 0..3); element k = channel within the chunk, stored with the SWIZZLE_32B pattern (16-byte halves
// of a 32-byte row swapped in rows 4..7 of every 8).
__global__ void reg_pack_convt4_umma_kernel(const float* __restrict__ w, int cin, int cout, int Cn, int wexp,
                                            __half* __restrict__ dst) {
  const int nchunks = cin / 16, nsplit = cout / Cn, rows = 16 * Cn;
  const long long total = static_cast<long long>(nsplit) * nchunks * 9 * rows * 16;
  const float sc = exp2f(static_cast<float>(wexp));
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    long long r = i;
    const int k = static_cast<int>(r % 16); r /= 16;
    const int n = static_cast<int>(r % rows); r /= rows;
    const int j = static_cast<int>(r % 9); r /= 9;
    const int c = static_cast<int>(r % nchunks);
    const int h = static_cast<int>(r / nchunks);
    const int slot = n / Cn, co = h * Cn + n % Cn;
    int sd, cls;
    if (slot < 4) { sd = 1; cls = 4 + slot; }
    else if (slot < 12) { sd = 0; cls = slot - 4; }
    else { sd = -1; cls = slot - 12; }
    const int sy = j / 3 - 1, sx = j % 3 - 1;
    const int kd = (cls >> 2) + 1 - 2 * sd, kh = ((cls >> 1) & 1) + 1 - 2 * sy, kw = (cls & 1) + 1 - 2 * sx;
    float v = 0.f;
    if (kd >= 0 && kd < 4 && kh >= 0 && kh < 4 && kw >= 0 && kw < 4)
      v = w[(static_cast<size_t>(c * 16 + k) * 64 + (kd * 4 + kh) * 4 + kw) * cout + co] * sc;
    const __half hi = __float2half_rn(v);
    const __half lo = __float2half_rn(v - __half2float(hi));
    const size_t block = (static_cast<size_t>(h) * nchunks + c) * 9 + j;
    const size_t off = static_cast<size_t>(n) * 16 + ((((k >> 3) ^ ((n >> 2) & 1))) << 3) + (k & 7);
    __half* b = dst + block * (2 * static_cast<size_t>(rows) * 16);
    b[off] = hi;
    b[static_cast<size_t>(rows) * 16 + off] = lo;
  }
}

__global__ void __launch_bounds__(kUmmaThreads, 1)
convt4_umma_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_raw,
                   const __grid_constant__ UmmaParams p) {
  extern __shared__ uint8_t umma_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(umma_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* bbuf = smem;                                                   // nb weight blocks
  uint8_t* abuf = bbuf + static_cast<size_t>(p.nb) * p.bblock;            // 2 activation buffers
  uint8_t* rawbuf = abuf + 2 * static_cast<size_t>(p.abuf_bytes);        // one raw box per epilogue group
  uint64_t* bars = reinterpret_cast<uint64_t*>(rawbuf + (p.stage_raw ? 2 * kRawBytes : 0));
  uint64_t* full_a = bars;         // [2]
  uint64_t* empty_a = bars + 2;    // [2]
  uint64_t* full_b = bars + 4;     // [8]
  uint64_t* empty_b = bars + 12;   // [8]
  uint64_t* acc_full = bars + 20;  // [2]
  uint64_t* acc_empty = bars + 22; // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);
  uint64_t* raw_full = bars + 25;  // [2]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Cn = p.Cn;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&full_a[i], 1);
      mbar_init(&empty_a[i], 1);
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 8);
      mbar_init(&raw_full[i], 1);
    }
    for (int i = 0; i < 8; ++i) {
      mbar_init(&full_b[i], 1);
      mbar_init(&empty_b[i], 1);
    }
    fence_mbar_init();
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_raw);
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ activation producer: R+2 slices x (hi, lo) per chunk
    if (lane == 0) {
      int ab = 0;
      uint32_t ph = 0;
      for (int u = blockIdx.x; u < p.nunits; u += gridDim.x) {
        const Unit ui = decode_unit(p, u);
        const int zlo = max(0, ui.z0 - 1), zhi = min(p.Di - 1, ui.z0 + ui.rd);
        for (int c = 0; c < p.nchunks; ++c) {
          mbar_wait(&empty_a[ab], ph ^ 1u, 100 + ab);
          mbar_arrive_expect_tx(&full_a[ab], static_cast<uint32_t>(zhi - zlo + 1) * 2u * kBoxBytes);
          uint8_t* base = abuf + static_cast<size_t>(ab) * p.abuf_bytes;
          for (int zi = zlo; zi <= zhi; ++zi) {
            const int slot = zi - (ui.z0 - 1);
            tma_load_5d(base + (slot * 2) * kSlotBytes, &tm_x, &full_a[ab], 0, ui.x0 - 1, ui.y0 - 1, zi,
                        ui.n * p.nchunks + c);
            tma_load_5d(base + (slot * 2 + 1) * kSlotBytes, &tm_x, &full_a[ab], 0, ui.x0 - 1, ui.y0 - 1, zi,
                        (p.N + ui.n) * p.nchunks + c);
          }
          if (++ab == 2) { ab = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ weight producer: nine blocks per chunk
    if (lane == 0) {
      int bb = 0;
      uint32_t ph = 0;
      for (int u = blockIdx.x; u < p.nunits; u += gridDim.x) {
        const Unit ui = decode_unit(p, u);
        const uint8_t* src = p.wumma + static_cast<size_t>(ui.h) * p.nchunks * 9 * p.bblock;
        for (int b = 0; b < p.nchunks * 9; ++b) {
          mbar_wait(&empty_b[bb], ph ^ 1u, 200 + bb);
          mbar_arrive_expect_tx(&full_b[bb], p.bblock);
          bulk_load(bbuf + static_cast<size_t>(bb) * p.bblock, src + static_cast<size_t>(b) * p.bblock, p.bblock,
                    &full_b[bb]);
          if (++bb == p.nb) { bb = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------ MMA issuer
    // One thread feeds the tensor pipe and a lone warp retires an instruction only every few clocks, so the loop is
    // kept to descriptor adds: per unit the column span / weight-row offset / box offset of each input slice are
    // tabulated once, the descriptors' high words (SBO, version, swizzle mode) are constants, and only the unit's very
    // first weight block (where fresh columns are overwritten instead of accumulated) takes the general path.
    const bool leader = elect_one();
    const uint32_t abuf_lo = ((smem_u32(abuf) & 0x3FFFF) >> 4) | (1u << 16);
    const uint32_t bbuf_lo = ((smem_u32(bbuf) & 0x3FFFF) >> 4) | (1u << 16);
    const uint32_t a_hi = ((kBX * kRowB) >> 4) | (1u << 14) | (kLayoutSw32 << 29);   // SBO = box pitch (10 rows)
    const uint32_t b_hi = ((8 * kRowB) >> 4) | (1u << 14) | (kLayoutSw32 << 29);     // SBO = 8 contiguous rows
    const uint32_t wlo_off = (static_cast<uint32_t>(16 * Cn) * kRowB) >> 4;          // lo weight plane inside a block
    const uint32_t alo_off = kSlotBytes >> 4;                                        // lo box of a slice
    const uint32_t idesc0 = umma_idesc_f16(128, 0, 0);
    const uint32_t abuf_step = p.abuf_bytes >> 4, bblk_step = p.bblock >> 4;
    int ab = 0, bb = 0, it = 0;
    uint32_t aph = 0, bph = 0;
    for (int u = blockIdx.x; u < p.nunits; u += gridDim.x, ++it) {
      const Unit ui = decode_unit(p, u);
      const int half = it & 1;
      const uint32_t use = (it >> 1) & 1u;
      const uint32_t tbase = tmem_base + static_cast<uint32_t>(half * kHalfCols);
      const int zlo = max(0, ui.z0 - 1), zhi = min(p.Di - 1, ui.z0 + ui.rd);
      const int ncols = ui.rd * 8 * Cn;
      // per input slice: TMEM column of its span, MMA shape, weight-row offset, box offset (descriptor units)
      // (entries indexed at compile time so the tables stay in registers; R + 2 <= 4 slices)
      uint32_t s_col[4], s_idesc[4], s_brow[4], s_abox[4];
      int s_c0[4], s_c1[4];
      bool s_on[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int zi = zlo + k;
        const int start = (zi - 1 - ui.z0) * 8 * Cn + 4 * Cn;
        const int c0 = max(start, 0), c1 = min(start + 16 * Cn, ncols);
        s_on[k] = zi <= zhi && c1 > c0;
        s_c0[k] = c0; s_c1[k] = c1;
        s_col[k] = tbase + static_cast<uint32_t>(c0);
        s_idesc[k] = idesc0 + (static_cast<uint32_t>((c1 - c0) >> 3) << 17);
        s_brow[k] = (static_cast<uint32_t>(c0 - start) * kRowB) >> 4;
        s_abox[k] = (static_cast<uint32_t>((zi - (ui.z0 - 1)) * 2) * kSlotBytes) >> 4;
      }
      mbar_wait(&acc_empty[half], use ^ 1u, 500 + half);
      tc_fence_after();
      for (int c = 0; c < p.nchunks; ++c) {
        mbar_wait(&full_a[ab], aph, 400 + ab);
        tc_fence_after();
        const uint32_t a_chunk = abuf_lo + static_cast<uint32_t>(ab) * abuf_step;
        uint32_t a_shift = 0;   // (sy * 10 + sx) rows of 32 bytes, in descriptor units
        for (int j = 0; j < 9; ++j) {
          mbar_wait(&full_b[bb], bph, 300 + bb);
          tc_fence_after();
          const uint32_t b_blk = bbuf_lo + static_cast<uint32_t>(bb) * bblk_step;
          const bool mma_on = !(p.debug & 2);
          if (c == 0 && j == 0) {
            // first block of the unit: columns from hw on are overwritten, the ones below accumulate
            if (leader && mma_on) {
              int hw = 0;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (!s_on[k]) continue;
                const uint32_t a0 = a_chunk + s_abox[k], b0 = b_blk + s_brow[k];
                const int c0 = s_c0[k], c1 = s_c1[k];
                const int mid = min(max(hw, c0), c1);
                if (mid > c0)
                  umma_f16_ss_lohi2(tbase + c0, a0, a_hi, b0, b_hi,
                                    idesc0 + (static_cast<uint32_t>((mid - c0) >> 3) << 17), 1u);
                if (c1 > mid)
                  umma_f16_ss_lohi2(tbase + mid, a0, a_hi, b0 + ((static_cast<uint32_t>(mid - c0) * kRowB) >> 4), b_hi,
                                    idesc0 + (static_cast<uint32_t>((c1 - mid) >> 3) << 17), 0u);
                hw = max(hw, c1);
                umma_f16_ss_lohi2(s_col[k], a0 + alo_off, a_hi, b0, b_hi, s_idesc[k], 1u);
                umma_f16_ss_lohi2(s_col[k], a0, a_hi, b0 + wlo_off, b_hi, s_idesc[k], 1u);
              }
            }
          } else {
            // the whole warp computes the (warp-uniform) operands so they live in uniform registers; one lane issues
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (!s_on[k]) continue;
              const uint32_t a0 = a_chunk + s_abox[k] + a_shift, b0 = b_blk + s_brow[k];
              const uint32_t d = s_col[k], id = s_idesc[k];
              if (leader && mma_on) {
                umma_f16_ss_lohi2(d, a0, a_hi, b0, b_hi, id, 1u);             // hi * w_hi
                umma_f16_ss_lohi2(d, a0 + alo_off, a_hi, b0, b_hi, id, 1u);   // lo * w_hi
                umma_f16_ss_lohi2(d, a0, a_hi, b0 + wlo_off, b_hi, id, 1u);   // hi * w_lo
              }
            }
          }
          if (leader) umma_commit(&empty_b[bb]);
          if (++bb == p.nb) { bb = 0; bph ^= 1u; }
          a_shift += (j % 3 == 2) ? static_cast<uint32_t>((kBX - 2) * kRowB) >> 4 : static_cast<uint32_t>(kRowB >> 4);
        }
        if (leader) umma_commit(&empty_a[ab]);
        if (++ab == 2) { ab = 0; aph ^= 1u; }
      }
      if (leader) umma_commit(&acc_full[half]);
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue: two groups of four warps
    // A unit's accumulators are cut into four items of 8 output channels x 8 classes of one lattice slice
    // (Cn = 16: 2 slices x 2 channel halves; Cn = 32: 1 slice x 4 channel quarters); group eg takes the items
    // eg, eg + 2.  Per channel a thread reads the 3 x 3 x 3 raw neighbourhood of its lattice point once (27 loads,
    // replicate-clamped) and interpolates all eight classes from it separably.
    const int q = warp & 3, eg = (warp - 4) >> 2, m = q * 32 + lane;
    const int ty = m >> 3, tx = m & 7;
    const int per_slice = Cn / 8;   // items per lattice slice
    // The raw neighbourhoods of an item (8 channels x 3 slices x 18 rows x 16 columns, fp32) are staged in shared memory
    // by one TMA box per item and group: the gather through L1 / L2 left the two warps per scheduler waiting on load
    // latency for most of the epilogue.  The box of the next item is requested as soon as the group has read this one.
    const bool staged = p.stage_raw != 0;
    const bool rleader = q == 0 && lane == 0;
    float* const sraw = reinterpret_cast<float*>(rawbuf + eg * kRawBytes);
    uint64_t* const rbar = &raw_full[eg];
    const CUtensorMap* const ptm_raw = &tm_raw;   // parameter-space address, taken outside the lambda
    uint32_t rph = 0;
    auto issue_raw = [&](const Unit& uu, int item) {
      const int a = item / per_slice, c8 = (item % per_slice) * 8;
      mbar_arrive_expect_tx(rbar, kRawBytes);
      tma_load_5d(sraw, ptm_raw, rbar, uu.x0 - 4, uu.y0 - 1, uu.z0 + a - 1, uu.n * p.raw_ratio + uu.h * Cn + c8, 0);
    };
    if (staged && rleader && static_cast<int>(blockIdx.x) < p.nunits) issue_raw(decode_unit(p, blockIdx.x), eg);
    int it = 0;
    for (int u = blockIdx.x; u < p.nunits; u += gridDim.x, ++it) {
      const Unit ui = decode_unit(p, u);
      const int half = it & 1;
      const uint32_t use = (it >> 1) & 1u;
      mbar_wait(&acc_full[half], use, 600 + half);
      tc_fence_after();
      const int y = ui.y0 + ty, x = ui.x0 + tx;
      const bool inside = y < p.Hi && x < p.Wi && 2 * x < p.Wo;
      const int yc = min(y, p.Hi - 1), xc = min(x, p.Wi - 1);
      const int xg[3] = {max(xc - 1, 0), xc, min(xc + 1, p.Wi - 1)};
      const int yg[3] = {max(yc - 1, 0), yc, min(yc + 1, p.Hi - 1)};
      int xo3[3], yr[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        xo3[k] = staged ? xg[k] - (ui.x0 - 4) : xg[k];
        yr[k] = staged ? (yg[k] - (ui.y0 - 1)) * kRawX : yg[k] * p.Wi;
      }
      const bool pair_ok = 2 * x + 1 < p.Wo;
      const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(half * kHalfCols);
      const int nitems = ui.rd * per_slice;
      for (int item = eg; item < nitems; item += 2) {
        const int a = item / per_slice, c8 = (item % per_slice) * 8;
        const int z = ui.z0 + a;
        uint32_t v[8][8];   // [class][channel]
#pragma unroll
        for (int cls = 0; cls < 8; ++cls) tmem_ld_32x8(t_lane + static_cast<uint32_t>(a * 8 * Cn + cls * Cn + c8), v[cls]);
        tmem_ld_wait();
        if (staged) {
          mbar_wait(rbar, rph, 700 + eg);
          rph ^= 1u;
        }
        if (inside && !(p.debug & 8)) {
          const int zg[3] = {max(z - 1, 0), z, min(z + 1, p.Di - 1)};
          int zr[3];
#pragma unroll
          for (int k = 0; k < 3; ++k) zr[k] = staged ? (zg[k] - (z - 1)) * (kBY * kRawX) : zg[k] * p.Hi * p.Wi;
          const int cog0 = ui.h * Cn + c8;
          const float* raw_c = staged ? sraw
                                      : p.in + ui.n * p.in_nstride + static_cast<long long>(cog0) * p.in_cstride;
          const long long raw_cs = staged ? 3 * kBY * kRawX : p.in_cstride;
          float* out_c = p.out + ui.n * p.out_nstride + static_cast<long long>(cog0) * p.out_cstride;
          long long orow[4];
          bool ook[4];
#pragma unroll
          for (int rp = 0; rp < 4; ++rp) {
            const int zo = 2 * z + (rp >> 1), yo = 2 * y + (rp & 1);
            ook[rp] = zo < p.Do && yo < p.Ho;
            orow[rp] = (static_cast<long long>(zo) * p.Ho + yo) * p.Wo + 2 * x;
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float* rc = raw_c + i * raw_cs;
            // upsample2x (trilinear, align_corners=False) = per axis 0.75 * own sample + 0.25 * the neighbour on the
            // class's side: x first (two parities per row), then y, then z
            float ex[3][3][2];
            if (!(p.debug & 1)) {
              float raw[3][3][3];
#pragma unroll
              for (int dz = 0; dz < 3; ++dz)
#pragma unroll
                for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                  for (int dx = 0; dx < 3; ++dx) raw[dz][dy][dx] = rc[zr[dz] + yr[dy] + xo3[dx]];
#pragma unroll
              for (int dz = 0; dz < 3; ++dz)
#pragma unroll
                for (int dy = 0; dy < 3; ++dy) {
                  const float c75 = 0.75f * raw[dz][dy][1];
                  ex[dz][dy][0] = fmaf(0.25f, raw[dz][dy][0], c75);
                  ex[dz][dy][1] = fmaf(0.25f, raw[dz][dy][2], c75);
                }
            } else {
#pragma unroll
              for (int dz = 0; dz < 3; ++dz)
#pragma unroll
                for (int dy = 0; dy < 3; ++dy) ex[dz][dy][0] = ex[dz][dy][1] = 0.f;
            }
            float ey[3][2][2];   // [dz][rh][rw]
#pragma unroll
            for (int dz = 0; dz < 3; ++dz)
#pragma unroll
              for (int rw = 0; rw < 2; ++rw) {
                const float c75 = 0.75f * ex[dz][1][rw];
                ey[dz][0][rw] = fmaf(0.25f, ex[dz][0][rw], c75);
                ey[dz][1][rw] = fmaf(0.25f, ex[dz][2][rw], c75);
              }
            const float bias = __ldg(p.bias + cog0 + i), bs = __ldg(p.bn_scale + cog0 + i),
                        bt = __ldg(p.bn_shift + cog0 + i);
            float* dst_c = out_c + i * p.out_cstride;
#pragma unroll
            for (int rp = 0; rp < 4; ++rp) {
              const int rd = rp >> 1, rh = rp & 1;
              float o[2];
#pragma unroll
              for (int rw = 0; rw < 2; ++rw) {
                const float r = fmaf(0.25f, ey[rd ? 2 : 0][rh][rw], 0.75f * ey[1][rh][rw]);
                o[rw] = (fmaf(__uint_as_float(v[rp * 2 + rw][i]), p.inv, bias) + r) * bs + bt;
              }
              if (!ook[rp]) continue;
              float* dst = dst_c + orow[rp];
              if (pair_ok && (reinterpret_cast<uintptr_t>(dst) & 7) == 0) {
                *reinterpret_cast<float2*>(dst) = make_float2(o[0], o[1]);
              } else {
                dst[0] = o[0];
                if (pair_ok) dst[1] = o[1];
              }
            }
          }
        }
        if (staged) {
          // the group is done with the box: request the next item's (this unit's, or the next unit's first)
          fence_proxy_async_smem();
          asm volatile("bar.sync %0, 128;" ::"r"(1 + eg) : "memory");
          if (rleader) {
            if (item + 2 < nitems) {
              issue_raw(ui, item + 2);
            } else if (u + static_cast<int>(gridDim.x) < p.nunits) {
              issue_raw(decode_unit(p, u + static_cast<int>(gridDim.x)), eg);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[half]);
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

// Output channels per unit: 32 where the layer has them (measured 5 - 15 % faster than 16 on every level although the
// weight stream per launch is proportional to Cn: fewer, longer units); OAI_CONVT4_CN=16 is the A/B switch.
int umma_cn(int cout) {
  static const int forced = [] {
    const char* e = getenv("OAI_CONVT4_CN");
    return e ? atoi(e) : 0;
  }();
  return (cout >= 32 && forced != 16) ? 32 : 16;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* q = nullptr;
    cudaDriverEntryPointQueryResult r;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &r) == cudaSuccess &&
        r == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(q);
  }
  return fn;
}

int make_xcl_tmap(CUtensorMap* tm, void* base, int cin, int Wi, int Hi, int Di, int planes) {
  EncodeTiledFn fn = encode_tiled();
  if (!fn) return fail("convt4_umma: cuTensorMapEncodeTiled not available from the driver");
  (void)cin;   // every plane is one 16-channel chunk
  cuuint64_t dims[5] = {16, (cuuint64_t)Wi, (cuuint64_t)Hi, (cuuint64_t)Di, (cuuint64_t)planes};
  cuuint64_t strides[4] = {32, (cuuint64_t)Wi * 32, (cuuint64_t)Hi * Wi * 32, (cuuint64_t)Di * Hi * Wi * 32};
  cuuint32_t box[5] = {16, kBX, kBY, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, base, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail("convt4_umma: cuTensorMapEncodeTiled failed with %d (cin=%d W=%d H=%d D=%d planes=%d)", (int)r, cin, Wi,
                Hi, Di, planes);
  return 0;
}

// `planes` dense fp32 planes of Di x Hi x Wi as a rank-5 TMA tensor (W, H, D, plane, 1); box = 16 x 18 x 3 x 8 x 1
int make_raw_tmap(CUtensorMap* tm, float* base, int Wi, int Hi, int Di, long long planes) {
  EncodeTiledFn fn = encode_tiled();
  if (!fn) return fail("convt4_umma: cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[5] = {(cuuint64_t)Wi, (cuuint64_t)Hi, (cuuint64_t)Di, (cuuint64_t)planes, 1};
  cuuint64_t strides[4] = {(cuuint64_t)Wi * 4, (cuuint64_t)Hi * Wi * 4, (cuuint64_t)Di * Hi * Wi * 4,
                           (cuuint64_t)planes * Di * Hi * Wi * 4};
  cuuint32_t box[5] = {kRawX, kBY, 3, 8, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, base, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail("convt4_umma: cuTensorMapEncodeTiled (raw planes) failed with %d (W=%d H=%d D=%d planes=%lld)", (int)r,
                Wi, Hi, Di, planes);
  return 0;
}

// A/B and test switch: OAI_B200_CONVT4_GATHER=1 reads the residual through L1 instead of the staged boxes
bool umma_gather_forced() {
  const char* e = getenv("OAI_B200_CONVT4_GATHER");
  return e && e[0] == '1';
}

}  // namespace

bool convt4_umma_eligible(const ConvT4Params& p) {
  return (p.cout == 16 || p.cout == 32 || p.cout == 64 || p.cout == 128) && p.cin % 16 == 0 && p.cin >= 16 && p.Wi >= kTX &&
         p.Hi >= kTX && p.cout <= p.cin;
}

size_t convt4_umma_wbytes(int cin, int cout) {
  if (cin % 16 || cout % 16) return 0;
  const int Cn = umma_cn(cout);
  return static_cast<size_t>(cout / Cn) * (cin / 16) * 9 * 2 * (16 * Cn) * kRowB;
}

int reg_pack_convt4_umma_launch(const float* w, int cin, int cout, int wexp, void* dst, cudaStream_t st) {
  const int Cn = umma_cn(cout);
  const long long total = static_cast<long long>(cout / Cn) * (cin / 16) * 9 * (16 * Cn) * 16;
  const unsigned blocks = static_cast<unsigned>(std::min<long long>((total + 255) / 256, 148 * 16));
  reg_pack_convt4_umma_kernel<<<blocks, 256, 0, st>>>(w, cin, cout, Cn, wexp, static_cast<__half*>(dst));
  return launched("reg_pack_convt4_umma_kernel");
}

int convt4_umma_launch(const ConvT4Params& p, cudaStream_t st) {
  const long long vol = static_cast<long long>(p.Di) * p.Hi * p.Wi;
  if (p.xsplit_bytes < static_cast<size_t>(p.N) * p.cin * vol * 4 || (reinterpret_cast<uintptr_t>(p.xsplit) & 127))
    return fail("convt4_umma: workspace of %zu bytes, 128-byte aligned, required",
                static_cast<size_t>(p.N) * p.cin * vol * 4);
  {
    const long long total = static_cast<long long>(p.N) * (p.cin / 16) * vol;
    const unsigned blocks = static_cast<unsigned>(std::min<long long>((total + 255) / 256, 148 * 32));
    reg_split_cl_kernel<<<blocks, 256, 0, st>>>(p.in, p.in_nstride, p.in_cstride, p.N, p.cin, vol,
                                                reinterpret_cast<uint4*>(p.xsplit));
    if (int rc = launched("reg_split_cl_kernel")) return rc;
  }
  UmmaParams q{};
  q.in = p.in; q.in_nstride = p.in_nstride; q.in_cstride = p.in_cstride;
  q.cin = p.cin; q.Di = p.Di; q.Hi = p.Hi; q.Wi = p.Wi; q.N = p.N;
  q.wumma = static_cast<const uint8_t*>(p.wumma);
  q.bias = p.bias; q.bn_scale = p.bn_scale; q.bn_shift = p.bn_shift;
  q.out = p.out; q.out_nstride = p.out_nstride; q.out_cstride = p.out_cstride;
  q.cout = p.cout; q.Do = p.Do; q.Ho = p.Ho; q.Wo = p.Wo;
  q.Cn = umma_cn(p.cout);
  q.R = kHalfCols / (8 * q.Cn);
  q.nsplit = p.cout / q.Cn;
  q.nchunks = p.cin / 16;
  q.bblock = 2u * 16u * q.Cn * kRowB;
  q.nb = q.Cn == 16 ? 4 : 3;   // weight blocks in flight (16 / 32 KB each)
  if (const char* e = getenv("OAI_CONVT4_NB")) q.nb = std::max(2, std::min(8, atoi(e)));   // development switch
  q.nx = (p.Wi + kTX - 1) / kTX; q.ny = (p.Hi + kTY - 1) / kTY; q.nzg = (p.Di + q.R - 1) / q.R;
  const long long nunits = static_cast<long long>(p.N) * q.nsplit * q.nzg * q.ny * q.nx;
  if (nunits > 0x7fffffffLL) return fail("convt4_umma: too many units");
  q.nunits = static_cast<int>(nunits);
  q.abuf_bytes = static_cast<uint32_t>((q.R + 2) * 2 * kSlotBytes);
  q.inv = exp2f(static_cast<float>(-p.wexp));
  {
    const char* e = getenv("OAI_CONVT4_DEBUG");
    q.debug = e ? atoi(e) : 0;
  }
  CUtensorMap tm, tm_raw;
  if (int rc = make_xcl_tmap(&tm, p.xsplit, p.cin, p.Wi, p.Hi, p.Di, 2 * p.N * (p.cin / 16))) return rc;
  // residual source by TMA when the raw input is a stack of dense, 16-byte aligned planes (every tallUNet2 level is)
  q.stage_raw = p.in_cstride == vol && p.in_nstride % p.in_cstride == 0 && p.Wi % 4 == 0 &&
                (reinterpret_cast<uintptr_t>(p.in) & 15) == 0 && !umma_gather_forced();
  q.raw_ratio = q.stage_raw ? static_cast<int>(p.in_nstride / p.in_cstride) : 0;
  if (q.stage_raw) {
    if (int rc = make_raw_tmap(&tm_raw, const_cast<float*>(p.in), p.Wi, p.Hi, p.Di,
                               static_cast<long long>(p.N - 1) * q.raw_ratio + p.cin))
      return rc;
  } else {
    tm_raw = tm;
  }
  const size_t smem = 1024 + static_cast<size_t>(q.nb) * q.bblock + 2 * static_cast<size_t>(q.abuf_bytes) +
                      (q.stage_raw ? 2 * kRawBytes : 0) + 256;
  static size_t configured[64] = {0};   // the attribute is per device
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || smem > configured[dev]) {
    if (int rc = check_cuda(cudaFuncSetAttribute(convt4_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 static_cast<int>(smem)),
                            "convt4_umma: cudaFuncSetAttribute"))
      return rc;
    if (dev >= 0 && dev < 64) configured[dev] = smem;
  }
  const int grid = std::min(q.nunits, num_sms());
  convt4_umma_kernel<<<grid, kUmmaThreads, smem, st>>>(tm, tm_raw, q);
  return launched("convt4_umma_kernel");
}

}  // namespace oai
