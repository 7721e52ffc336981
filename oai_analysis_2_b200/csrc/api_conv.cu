// C-ABI entry points of the tcgen05 implicit-GEMM conv: geometry planner, weight packer, launcher.
#include "../../include/oai_b200.h"
#include "api_common.h"
#include "conv_igemm.cuh"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstring>
#include <vector>

namespace oai {

cudaError_t conv_igemm_launch(const ConvIgemmParams& p, const CUtensorMap& tm0, const CUtensorMap& tm1, int num_sms,
                              cudaStream_t stream);

namespace {

constexpr int kFlagForcePerTap = 1;
constexpr int kFlagBaseOffFormula = 2;  // debug: descriptor base_offset = (addr>>7)&7 (measured WRONG on B200)
constexpr int kFlagForceKd1 = 4;
constexpr int kFlagNoFastPath = 8;
constexpr int kFlagWideN = 16;
constexpr int kFlagWideRows = 32;  // A/B: keep 128-byte rows (zero-filled upper half) for a 32-channel source  // A/B: keep 256-wide N tiles for 3x3x3 layers with cout >= 256 (one kd per weight block)
constexpr size_t kSmemBudget = 232448 - 1024 - 48 * 8 - 2112 - 2048;  // 227 KB minus alignment slack, barriers, head weights, bias

struct Plan {
  int mode, kd_per_block, R, Rd, up_groups, nhalf, cph, nblk, nchunk0, nchunk1, k16, row_bytes, TW, TH, n_wbuf, n_astage;
  uint32_t wblock_bytes, astage_bytes, astage_stride;
};

int make_plan(int D, int H, int W, int c0, int c1, int cout, int pointwise, int flags, Plan* pl, int d_cnt = 0) {
  OAI_REQUIRE(D > 0 && H > 0 && W > 0 && c0 > 0 && c1 >= 0 && cout > 0, "conv plan: bad dims");
  OAI_REQUIRE(c0 % 8 == 0 && c1 % 8 == 0, "conv plan: channel counts must be multiples of 8 (got %d,%d)", c0, c1);
  pl->TW = W < 128 ? W : 128;
  OAI_REQUIRE(128 % pl->TW == 0 && W % pl->TW == 0, "conv plan: W=%d must divide or be a multiple of 128", W);
  pl->TH = 128 / pl->TW;
  OAI_REQUIRE(H % pl->TH == 0, "conv plan: H=%d not a multiple of the %d-row M tile", H, pl->TH);
  // 3x3x3 layers split cout into 128-wide N tiles: each tile then stacks the three kd taps of an input slice (N = 256 +
  // 128) and keeps R = 4 accumulators, which halves the weight bytes streamed per output compared with a 256-wide
  // tile limited to R = 2 by the 512 TMEM columns.  Pointwise layers keep 256-wide tiles.
  const int ntile = (!pointwise && cout > 128 && !(flags & kFlagWideN)) ? 128 : 256;
  pl->nhalf = (cout + ntile - 1) / ntile;
  OAI_REQUIRE(cout % pl->nhalf == 0, "conv plan: cout=%d not divisible into %d N splits", cout, pl->nhalf);
  pl->cph = cout / pl->nhalf;
  OAI_REQUIRE(pl->cph % 32 == 0, "conv plan: cout per split (%d) must be a multiple of 32", pl->cph);
  if (d_cnt <= 0) d_cnt = D;
  // as many accumulators as TMEM holds; the last d-group of a region may be partial
  int R = 512 / pl->cph < 8 ? 512 / pl->cph : 8;
  if (R > d_cnt) R = d_cnt;
  pl->R = R;
  pl->Rd = R;
  pl->up_groups = 1;
  pl->nchunk0 = (c0 + 63) / 64;
  pl->nchunk1 = (c1 + 63) / 64;
  pl->k16 = (c1 == 0 && c0 <= 32) ? (c0 <= 16 ? 1 : 2) : 4;
  // a single 32-channel source keeps 64-byte rows (SWIZZLE_64B): half the TMA / shared-memory bytes per voxel
  pl->row_bytes = (c1 == 0 && c0 == 32 && !(flags & kFlagWideRows)) ? 64 : 128;
  const int rb = pl->row_bytes;
  const int nch = pl->nchunk0 + pl->nchunk1;
  if (pointwise == 2) {
    // ConvTranspose3d(k2,s2): the unit's accumulators are taps of one M tile
    pl->mode = kModeUp2;
    pl->kd_per_block = 1;
    pl->R = 512 / pl->cph < 8 ? 512 / pl->cph : 8;
    pl->Rd = 1;
    pl->up_groups = 8 / pl->R;
    pl->wblock_bytes = pl->R * pl->cph * rb;
    pl->nblk = nch;
    pl->n_wbuf = 2;
  } else if (pointwise) {
    pl->mode = kModePointwise;
    pl->kd_per_block = 1;
    pl->wblock_bytes = pl->cph * rb;
    pl->nblk = nch;
    pl->n_wbuf = 3;
  } else if (pl->TW == 128 && pl->cph <= 64 && !(flags & kFlagForcePerTap)) {
    pl->mode = kModeRowShared;
    pl->kd_per_block = 3;
    pl->wblock_bytes = 9 * pl->cph * rb;
    pl->nblk = nch * 3;
    pl->n_wbuf = 2;
  } else {
    pl->mode = kModePerTap;
    pl->kd_per_block = (pl->cph <= 128 && !(flags & kFlagForceKd1)) ? 3 : 1;
    pl->wblock_bytes = pl->kd_per_block * pl->cph * rb;
    pl->nblk = nch * (pl->kd_per_block == 3 ? 9 : 27);
    pl->n_wbuf = pl->kd_per_block == 3 ? 2 : 3;
  }
  pl->astage_bytes = (pl->mode == kModeRowShared ? 130 : 128) * rb;
  pl->astage_stride = (pl->astage_bytes + 1023u) & ~1023u;
  const size_t wstride = (pl->wblock_bytes + 1023u) & ~size_t(1023);
  const size_t left = kSmemBudget - pl->n_wbuf * wstride;
  int ns = static_cast<int>(left / pl->astage_stride);
  if (ns > 8) ns = 8;
  OAI_REQUIRE(ns >= 2, "conv plan: shared memory budget leaves %d A stages", ns);
  pl->n_astage = ns;
  return 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// NDHWC 16-bit activation tensor viewed as a rank-5 TMA tensor (C, W, H, D, N); box = (64, bw, bh, 1, 1).
int make_act_tmap(CUtensorMap* tm, const void* base, int C, int W, int H, int D, int N, int bw, int bh, int fmt,
                  int row_bytes = 128) {
  EncodeTiledFn fn = encode_fn();
  OAI_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
  cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2,
                           (cuuint64_t)D * H * W * C * 2};
  cuuint32_t box[5] = {(cuuint32_t)(row_bytes / 2), (cuuint32_t)bw, (cuuint32_t)bh, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(tm, fmt == 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5,
                  const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  OAI_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with %d (C=%d W=%d H=%d D=%d N=%d box %dx%d)", (int)r,
              C, W, H, D, N, bw, bh);
  return 0;
}

inline uint16_t to16(float x, int fmt) {
  if (fmt == 0) {
    __half h = __float2half_rn(x);
    uint16_t u;
    memcpy(&u, &h, 2);
    return u;
  }
  __nv_bfloat16 h = __float2bfloat16_rn(x);
  uint16_t u;
  memcpy(&u, &h, 2);
  return u;
}

// byte offset of element j (16-bit) of row r in a K-major swizzled block: 16-byte chunks are XOR-ed with the row index
// (SWIZZLE_128B: chunk ^= r & 7 over 128-byte rows; SWIZZLE_64B: chunk ^= (r >> 1) & 3 over 64-byte rows)
inline size_t swz_off(int r, int j, int row_bytes) {
  const int chunk = j >> 3;
  const int sw = row_bytes == 128 ? (chunk ^ (r & 7)) : (chunk ^ ((r >> 1) & 3));
  return static_cast<size_t>(r) * row_bytes + (static_cast<size_t>(sw) << 4) + (j & 7) * 2;
}

inline float from16(uint16_t u, int fmt) {
  if (fmt == 0) {
    __half h;
    memcpy(&h, &u, 2);
    return __half2float(h);
  }
  __nv_bfloat16 h;
  memcpy(&h, &u, 2);
  return __bfloat162float(h);
}

}  // namespace
}  // namespace oai

using namespace oai;

namespace {
// Optional per-launch timing of the conv kernel (CUDA events on the launch stream) for bench.py's roofline line.
struct ProfEntry {
  cudaEvent_t a, b;
  double flops;       // algorithmic: every MAC the reference executes for this layer
  double exec_flops;  // MACs actually issued (dead-halo rows skipped)
};

// Inside a stream capture (the per-knee CUDA graph) the profiling events become external event-record nodes: every
// replay re-records them, so oai_profile_end reports the conv launches of the most recent replay.
inline void prof_record(cudaEvent_t ev, cudaStream_t st) {
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(st, &cs);
  cudaEventRecordWithFlags(ev, st, cs == cudaStreamCaptureStatusActive ? cudaEventRecordExternal : cudaEventRecordDefault);
}
std::vector<ProfEntry> g_prof;
size_t g_prof_used = 0;
bool g_prof_on = false;
}  // namespace

extern "C" int oai_conv3d_igemm_plan(int D, int H, int W, int c0, int c1, int cout, int pointwise, int flags,
                                     int* plan) {
  Plan pl;
  if (make_plan(D, H, W, c0, c1, cout, pointwise, flags, &pl)) return 1;
  plan[0] = pl.mode;
  plan[1] = pl.mode == kModeUp2 ? pl.up_groups : pl.kd_per_block;
  plan[2] = pl.R;
  plan[3] = pl.nhalf;
  plan[4] = pl.cph;
  plan[5] = pl.nblk;
  plan[6] = static_cast<int>(pl.wblock_bytes);
  plan[7] = pl.nchunk0 + pl.nchunk1;
  plan[8] = pl.row_bytes;
  return 0;
}

extern "C" int oai_pack_conv_weights(const float* w, int cout, int c0, int c1, int D, int H, int W, int pointwise,
                                     int ab_format, int flags, void* dst, size_t dst_bytes) {
  Plan pl;
  if (make_plan(D, H, W, c0, c1, cout, pointwise, flags, &pl)) return 1;
  const size_t need = static_cast<size_t>(pl.nhalf) * pl.nblk * pl.wblock_bytes;
  OAI_REQUIRE(dst_bytes >= need, "pack: dst holds %zu bytes, need %zu", dst_bytes, need);
  const int cin = c0 + c1;
  const int ktaps = pointwise ? 1 : 27;
  uint8_t* out = static_cast<uint8_t*>(dst);
  memset(out, 0, need);
  // Round to 16 bits with error feedback along the taps of each (co, ci) filter: the rounding residual of one tap is
  // carried into the next, so the 27 rounding errors of a filter sum to (almost) zero.  On smooth inputs -- where all
  // taps see nearly the same activation -- this cancels the systematic part of the weight-quantisation error.
  std::vector<uint16_t> q(static_cast<size_t>(cout) * cin * ktaps);
  for (size_t f = 0; f < static_cast<size_t>(cout) * cin; ++f) {
    double carry = 0.0;
    for (int t = 0; t < ktaps; ++t) {
      const double v = static_cast<double>(w[f * ktaps + t]) + carry;
      const uint16_t h = to16(static_cast<float>(v), ab_format);
      q[f * ktaps + t] = h;
      carry = v - static_cast<double>(from16(h, ab_format));
    }
  }
  for (int nh = 0; nh < pl.nhalf; ++nh) {
    for (int b = 0; b < pl.nblk; ++b) {
      uint8_t* blk = out + (static_cast<size_t>(nh) * pl.nblk + b) * pl.wblock_bytes;
      // decode block exactly as the kernel does
      int c, kh, kw0, nkw, kdlo, nkd;
      if (pl.mode == kModeRowShared) {
        c = b / 3; kh = b % 3; kw0 = 0; nkw = 3; kdlo = 0; nkd = 3;
      } else if (pl.mode == kModePerTap) {
        if (pl.kd_per_block == 3) {
          c = b / 9; const int r = b % 9; kh = r / 3; kw0 = r % 3; nkw = 1; kdlo = 0; nkd = 3;
        } else {
          c = b / 27; const int r = b % 27; kh = r / 9; kw0 = (r / 3) % 3; nkw = 1; kdlo = r % 3; nkd = 1;
        }
      } else {
        c = b; kh = 0; kw0 = 0; nkw = 1; kdlo = 0; nkd = 1;
      }
      const bool s0 = c < pl.nchunk0;
      const int cbase = s0 ? c * 64 : c0 + (c - pl.nchunk0) * 64;
      const int climit = s0 ? c0 : cin;
      for (int kwi = 0; kwi < nkw; ++kwi) {
        for (int ti = 0; ti < nkd; ++ti) {
          const int kd = kdlo + nkd - 1 - ti;
          const int kw = kw0 + kwi;
          const int tap = pointwise ? 0 : (kd * 3 + kh) * 3 + kw;
          for (int co = 0; co < pl.cph; ++co) {
            const int r = (kwi * nkd + ti) * pl.cph + co;
            const uint16_t* wrow = q.data() + (static_cast<size_t>(nh * pl.cph + co) * cin) * ktaps;
            for (int j = 0; j < 64; ++j) {
              const int ci = cbase + j;
              if (ci >= climit) break;
              if (j * 2 >= pl.row_bytes) break;
              const uint16_t h = wrow[static_cast<size_t>(ci) * ktaps + tap];
              const size_t off = swz_off(r, j, pl.row_bytes);
              memcpy(blk + off, &h, 2);
            }
          }
        }
      }
    }
  }
  return 0;
}

static int conv_common(const void* src0, int c0, const void* src1, int c1, int NT, int D, int H, int W,
                       const void* wpack, size_t wpack_bytes, const float* bias, int cout, int pointwise, int relu,
                       int ab_format, void* out, long long obase, long long osN, long long osD, long long osH,
                       long long osW, int flags, const HeadFuse* head, const int* region, void* stream);

extern "C" int oai_conv3d_igemm(const void* src0, int c0, const void* src1, int c1, int NT, int D, int H, int W,
                                const void* wpack, size_t wpack_bytes, const float* bias, int cout, int pointwise,
                                int relu, int ab_format, void* out, long long obase, long long osN, long long osD,
                                long long osH, long long osW, int flags, void* stream) {
  OAI_REQUIRE(out != nullptr, "conv: null output");
  return conv_common(src0, c0, src1, c1, NT, D, H, W, wpack, wpack_bytes, bias, cout, pointwise, relu, ab_format, out,
                     obase, osN, osD, osH, osW, flags, nullptr, nullptr, stream);
}

extern "C" int oai_conv3d_igemm_region(const void* src0, int c0, const void* src1, int c1, int NT, int D, int H, int W,
                                       const void* wpack, size_t wpack_bytes, const float* bias, int cout,
                                       int pointwise, int relu, int ab_format, void* out, long long obase,
                                       long long osN, long long osD, long long osH, long long osW, int flags,
                                       const int* region, void* stream) {
  OAI_REQUIRE(out != nullptr && region != nullptr, "conv region: null pointer");
  return conv_common(src0, c0, src1, c1, NT, D, H, W, wpack, wpack_bytes, bias, cout, pointwise, relu, ab_format, out,
                     obase, osN, osD, osH, osW, flags, nullptr, region, stream);
}

extern "C" int oai_conv3d_igemm_head(const void* src0, int c0, const void* src1, int c1, int NT, int D, int H, int W,
                                     const void* wpack, size_t wpack_bytes, const float* bias, int ab_format,
                                     int ncls, const float* head_w, const float* head_b, float* out,
                                     const int* vol_dims, const int* geom, int tile0, const int* crop_zyx,
                                     int out_mode, int flags, void* stream) {
  // only the tile interior is ever written, so only the interior rows are computed
  const int region[4] = {geom[6], geom[3], geom[7], geom[4]};
  OAI_REQUIRE(head_w && head_b && out && vol_dims && geom && crop_zyx, "conv head: null pointer");
  OAI_REQUIRE(ncls >= 1 && ncls <= 8, "conv head: ncls=%d unsupported", ncls);
  OAI_REQUIRE(geom[0] == D && geom[1] == H && geom[2] == W, "conv head: tile geometry does not match the layer");
  HeadFuse hd;
  hd.enabled = 1; hd.ncls = ncls; hd.out_mode = out_mode; hd.w = head_w; hd.b = head_b; hd.out = out;
  hd.VD = vol_dims[0]; hd.VH = vol_dims[1]; hd.VW = vol_dims[2];
  hd.ed = geom[3]; hd.eh = geom[4]; hd.ew = geom[5];
  hd.od = geom[6]; hd.oh = geom[7]; hd.ow = geom[8];
  hd.gh = geom[10]; hd.gw = geom[11]; hd.tile0 = tile0;
  hd.cz = crop_zyx[0]; hd.cy = crop_zyx[1]; hd.cx = crop_zyx[2];
  return conv_common(src0, c0, src1, c1, NT, D, H, W, wpack, wpack_bytes, bias, 64, 0, 1, ab_format, nullptr, 0, 0, 0,
                     0, 0, flags, &hd, region, stream);
}

static int conv_common(const void* src0, int c0, const void* src1, int c1, int NT, int D, int H, int W,
                       const void* wpack, size_t wpack_bytes, const float* bias, int cout, int pointwise, int relu,
                       int ab_format, void* out, long long obase, long long osN, long long osD, long long osH,
                       long long osW, int flags, const HeadFuse* head, const int* region, void* stream) {
  // region = {d_lo, d_cnt, h_lo, h_cnt}: the output sub-box to compute (full rows in w); h range is widened to whole
  // M-tile rows.  Everything outside is dead halo the caller never reads.
  int d_lo = 0, d_cnt = D, h_lo = 0, h_cnt = H;
  if (region) {
    d_lo = region[0]; d_cnt = region[1]; h_lo = region[2]; h_cnt = region[3];
    OAI_REQUIRE(d_lo >= 0 && d_cnt >= 1 && d_lo + d_cnt <= D && h_lo >= 0 && h_cnt >= 1 && h_lo + h_cnt <= H,
                "conv: region [%d,+%d)x[%d,+%d) outside %dx%d", d_lo, d_cnt, h_lo, h_cnt, D, H);
  }
  Plan pl;
  if (make_plan(D, H, W, c0, c1, cout, pointwise, flags, &pl, d_cnt)) return 1;
  const int hp_lo = h_lo / pl.TH;
  const int hp_cnt = (h_lo + h_cnt + pl.TH - 1) / pl.TH - hp_lo;
  OAI_REQUIRE(src0 && wpack && bias, "conv: null pointer");
  OAI_REQUIRE(!head || (pl.cph == 64 && pl.nhalf == 1), "conv head: the fused head needs a 64-channel layer");
  OAI_REQUIRE((c1 == 0) == (src1 == nullptr), "conv: src1/c1 mismatch");
  const size_t need = static_cast<size_t>(pl.nhalf) * pl.up_groups * pl.nblk * pl.wblock_bytes;
  OAI_REQUIRE(wpack_bytes == need, "conv: packed weights are %zu bytes, geometry needs %zu", wpack_bytes, need);
  OAI_REQUIRE(obase % 8 == 0 && osN % 8 == 0 && osD % 8 == 0 && osH % 8 == 0 && osW % 8 == 0,
              "conv: output strides must keep 16-byte alignment");

  ConvIgemmParams p;
  memset(&p, 0, sizeof(p));
  p.NT = NT; p.D = D; p.H = H; p.W = W;
  p.TW = pl.TW; p.TH = pl.TH; p.R = pl.R;
  p.cout = pl.cph; p.nhalf = pl.nhalf;
  p.nchunk0 = pl.nchunk0; p.nchunk1 = pl.nchunk1; p.k16_steps = pl.k16; p.row_bytes = pl.row_bytes;
  p.mode = pl.mode; p.kd_per_block = pl.kd_per_block; p.nblk = pl.nblk;
  p.wblock_bytes = pl.wblock_bytes; p.n_wbuf = pl.n_wbuf; p.n_astage = pl.n_astage;
  p.astage_bytes = pl.astage_bytes; p.astage_stride = pl.astage_stride;
  p.ab_format = ab_format; p.relu = relu;
  p.base_off_mode = (flags & kFlagBaseOffFormula) ? 1 : 0;
  p.no_fast_path = (flags & kFlagNoFastPath) ? 1 : 0;
  p.wpack = static_cast<const uint8_t*>(wpack);
  p.bias = bias;
  p.out = out;
  p.obase = obase; p.osN = osN; p.osD = osD; p.osH = osH; p.osW = osW;
  p.d_lo = d_lo; p.d_cnt = d_cnt; p.hp_lo = hp_lo; p.hp_cnt = hp_cnt;
  p.Rd = pl.Rd; p.up_groups = pl.up_groups;
  p.nunits = NT * ((d_cnt + pl.Rd - 1) / pl.Rd) * (hp_cnt * (W / pl.TW)) * pl.up_groups * pl.nhalf;
  if (pl.mode == kModeUp2) {
    for (int t = 0; t < 8; ++t)
      p.tap_off[t] = ((static_cast<long long>(t >> 2) * 2 * H + ((t >> 1) & 1)) * 2 * W + (t & 1)) * cout;
  }
  if (head) p.head = *head;

  OAI_REQUIRE(pl.mode != kModeUp2 || (c1 == 0 && !head), "conv: up2 mode takes one source and no fused head");
  const int bw = pl.mode == kModeRowShared ? 130 : pl.TW;
  const int bh = pl.TH;
  CUtensorMap tm0, tm1;
  if (make_act_tmap(&tm0, src0, c0, W, H, D, NT, bw, bh, ab_format, pl.row_bytes)) return 1;
  if (src1) {
    if (make_act_tmap(&tm1, src1, c1, W, H, D, NT, bw, bh, ab_format)) return 1;
  } else {
    tm1 = tm0;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfEntry* pe = nullptr;
  if (g_prof_on) {
    if (g_prof_used == g_prof.size()) {
      ProfEntry ne;
      if (cudaEventCreate(&ne.a) != cudaSuccess || cudaEventCreate(&ne.b) != cudaSuccess)
        return fail("conv profile: cannot create events");
      g_prof.push_back(ne);
    }
    pe = &g_prof[g_prof_used++];
    pe->flops = 2.0 * NT * D * H * W * static_cast<double>(cout) * (c0 + c1) * (pointwise ? 1 : 27);
    pe->exec_flops = pe->flops * (static_cast<double>(d_cnt) * hp_cnt * pl.TH) / (static_cast<double>(D) * H);
    prof_record(pe->a, st);
  }
  cudaError_t e = conv_igemm_launch(p, tm0, tm1, num_sms(), st);
  if (pe) prof_record(pe->b, st);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return check_cuda(e, "conv_igemm launch");
}

extern "C" int oai_profile_begin(void) {
  g_prof_used = 0;
  g_prof_on = true;
  return 0;
}

extern "C" int oai_profile_end(double* conv_ms, long long* conv_launches, double* conv_flops,
                               double* conv_exec_flops) {
  g_prof_on = false;
  double ms = 0, fl = 0, xfl = 0;
  for (size_t i = 0; i < g_prof_used; ++i) {
    if (int rc = check_cuda(cudaEventSynchronize(g_prof[i].b), "conv profile: event sync")) return rc;
    float t = 0;
    if (int rc = check_cuda(cudaEventElapsedTime(&t, g_prof[i].a, g_prof[i].b), "conv profile: elapsed")) return rc;
    ms += t;
    fl += g_prof[i].flops;
    xfl += g_prof[i].exec_flops;
  }
  if (conv_ms) *conv_ms = ms;
  if (conv_launches) *conv_launches = static_cast<long long>(g_prof_used);
  if (conv_flops) *conv_flops = fl;
  if (conv_exec_flops) *conv_exec_flops = xfl;
  g_prof_used = 0;
  return 0;
}


// ---------------------------------------------------------------------------------------------- ConvTranspose3d k2 s2
extern "C" int oai_pack_convt2_weights(const float* w, int cout, int cin, int D, int H, int W, int ab_format,
                                       void* dst, size_t dst_bytes) {
  Plan pl;
  if (make_plan(D, H, W, cin, 0, cout, 2, 0, &pl)) return 1;
  const size_t need = static_cast<size_t>(pl.nhalf) * pl.up_groups * pl.nblk * pl.wblock_bytes;
  OAI_REQUIRE(dst_bytes >= need, "pack up2: dst holds %zu bytes, need %zu", dst_bytes, need);
  uint8_t* out = static_cast<uint8_t*>(dst);
  memset(out, 0, need);
  for (int nh = 0; nh < pl.nhalf; ++nh)
    for (int tg = 0; tg < pl.up_groups; ++tg)
      for (int b = 0; b < pl.nblk; ++b) {
        uint8_t* blk = out + ((static_cast<size_t>(nh) * pl.up_groups + tg) * pl.nblk + b) * pl.wblock_bytes;
        for (int ti = 0; ti < pl.R; ++ti) {
          const int tap = tg * pl.R + ti;
          for (int co = 0; co < pl.cph; ++co) {
            const int r = ti * pl.cph + co;
            const float* wrow = w + static_cast<size_t>(nh * pl.cph + co) * cin * 8;
            for (int j = 0; j < 64; ++j) {
              const int ci = b * 64 + j;
              if (ci >= cin) break;
              const uint16_t h = to16(wrow[static_cast<size_t>(ci) * 8 + tap], ab_format);
              const size_t off = swz_off(r, j, 128);
              memcpy(blk + off, &h, 2);
            }
          }
        }
      }
  return 0;
}

extern "C" int oai_convt2_igemm(const void* src, int cin, int NT, int D, int H, int W, const void* wpack,
                                size_t wpack_bytes, const float* bias, int cout, int relu, int ab_format, void* out,
                                const int* region, void* stream) {
  OAI_REQUIRE(out != nullptr, "convt2: null output");
  const long long c = cout;
  return conv_common(src, cin, nullptr, 0, NT, D, H, W, wpack, wpack_bytes, bias, cout, 2, relu, ab_format, out, 0,
                     8ll * D * H * W * c, 8ll * H * W * c, 4ll * W * c, 2 * c, 0, nullptr, region, stream);
}
