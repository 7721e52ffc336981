// Intensity pre-normalisation on the device: oai_analysis/dask_processing.py:10-26 (image_normalize) =
//   window_min/max = np.percentile(volume, p_lo / p_hi)          (numpy "linear" method: lerp of two order statistics)
//   itk.IntensityWindowingImageFilter: x < wmin -> out_min, x > wmax -> out_max, else x * factor + offset (in double)
// The reference sorts (partitions) 23.6 M voxels on the host; here the four order statistics come from an exact
// three-pass radix select (11 + 11 + 10 bits of the order-preserving integer image of the float keys), with no host
// round trip: histogram -> pick bin -> histogram of the surviving prefix ... -> window -> apply.  HBM-bound: the volume
// is read four times and written once (20 B / voxel).
#include "../../include/oai_b200.h"
#include "api_common.h"

namespace oai {
namespace {

constexpr int kTargets = 4;  // floor / ceil order statistics of the two percentiles
struct SelectState {
  unsigned long long rank[kTargets];   // remaining rank inside the surviving prefix
  unsigned int prefix[kTargets];       // selected high bits so far
  unsigned int hist1[2048];
  unsigned int hist2[kTargets][2048];
  unsigned int hist3[kTargets][1024];
  float frac[2];                       // interpolation weights (numpy's gamma) of the two percentiles
  double window[2];                    // result: window_min, window_max
};

__device__ __forceinline__ unsigned int float_key(float v) {
  const unsigned int u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // monotone: key order == float order
}
__device__ __forceinline__ float key_float(unsigned int k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__global__ void __launch_bounds__(256) window_init_kernel(SelectState* s, long long n, double p_lo, double p_hi) {
  const int t = threadIdx.x;
  for (int i = t; i < 2048; i += 256) {
    s->hist1[i] = 0;
    for (int k = 0; k < kTargets; ++k) s->hist2[k][i] = 0;
    if (i < 1024)
      for (int k = 0; k < kTargets; ++k) s->hist3[k][i] = 0;
  }
  if (t < 2) {
    // numpy >= 2 on a float32 array (NEP 50: the Python-float percentile and the integer count are "weak" and take the
    // array's dtype): q = float32(p) / float32(100); virtual index = float32(n - 1) * q; previous = floor, next =
    // previous + 1, both the last element once the virtual index reaches n - 1; gamma = virtual - previous, in float32
    // (numpy/lib/_function_base_impl.py::_quantile, _get_indexes, _get_gamma).
    const float q = __fdiv_rn(static_cast<float>(t == 0 ? p_lo : p_hi), 100.f);
    const float vi = __fmul_rn(static_cast<float>(n - 1), q);
    unsigned long long k0, k1;
    float gamma;
    if (vi >= static_cast<float>(n - 1)) {
      k0 = k1 = static_cast<unsigned long long>(n - 1);
      gamma = 0.f;
    } else {
      const float f = fmaxf(floorf(vi), 0.f);
      k0 = static_cast<unsigned long long>(f);
      k1 = k0 + 1;
      gamma = __fsub_rn(vi, f);
    }
    s->rank[2 * t] = k0;
    s->rank[2 * t + 1] = k1;
    s->frac[t] = gamma;
    s->prefix[2 * t] = s->prefix[2 * t + 1] = 0;
  }
}

// PASS 0: top 11 bits of every key; PASS 1 / 2: next 11 / last 10 bits of the keys whose higher bits match a target
template <int PASS>
__global__ void __launch_bounds__(512) window_hist_kernel(const float* __restrict__ in, long long n, SelectState* s) {
  constexpr int BINS = PASS == 2 ? 1024 : 2048;
  constexpr int NH = PASS == 0 ? 1 : kTargets;
  __shared__ unsigned int sh[NH][BINS];
  for (int i = threadIdx.x; i < NH * BINS; i += blockDim.x) (&sh[0][0])[i] = 0;
  unsigned int pre[kTargets] = {0, 0, 0, 0};
  if (PASS > 0) {
#pragma unroll
    for (int k = 0; k < kTargets; ++k) pre[k] = s->prefix[k];
  }
  __syncthreads();
  const long long n4 = n / 4;
  auto visit = [&](float v) {
    const unsigned int key = float_key(v);
    if (PASS == 0) {
      // smooth volumes put most keys of a warp into a handful of bins: one atomic per distinct bin of the warp
      const unsigned int bin = key >> 21, act = __activemask();
      const unsigned int same = __match_any_sync(act, bin);
      if ((__ffs(same) - 1) == static_cast<int>(threadIdx.x & 31)) atomicAdd(&sh[0][bin], __popc(same));
    } else {
      const unsigned int hi = PASS == 1 ? key >> 21 : key >> 10;
      const unsigned int bin = PASS == 1 ? (key >> 10) & 2047u : key & 1023u;
#pragma unroll
      for (int k = 0; k < kTargets; ++k)
        if (hi == pre[k]) atomicAdd(&sh[PASS == 0 ? 0 : k][bin], 1u);
    }
  };
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(in) + i);
    visit(v.x); visit(v.y); visit(v.z); visit(v.w);
  }
  if (blockIdx.x == 0)
    for (long long i = 4 * n4 + threadIdx.x; i < n; i += blockDim.x) visit(__ldg(in + i));
  __syncthreads();
  unsigned int* g = PASS == 0 ? s->hist1 : (PASS == 1 ? &s->hist2[0][0] : &s->hist3[0][0]);
  for (int i = threadIdx.x; i < NH * BINS; i += blockDim.x) {
    const unsigned int c = (&sh[0][0])[i];
    if (c) atomicAdd(g + i, c);
  }
}

// one warp per target: find the bin holding the remaining rank, extend the prefix
template <int PASS>
__global__ void __launch_bounds__(32 * kTargets) window_pick_kernel(SelectState* s, float out_min, float out_max) {
  constexpr int BINS = PASS == 2 ? 1024 : 2048;
  constexpr int BITS = PASS == 2 ? 10 : 11;
  const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned int* h = PASS == 0 ? s->hist1 : (PASS == 1 ? s->hist2[k] : s->hist3[k]);
  unsigned long long rank = s->rank[k], base = 0;
  int found = -1;
  for (int b0 = 0; b0 < BINS && found < 0; b0 += 32) {
    const unsigned int c = h[b0 + lane];
    unsigned long long incl = c;  // inclusive scan over the warp
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned long long o = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += o;
    }
    const bool here = base + incl > rank;
    const unsigned int m = __ballot_sync(0xffffffffu, here);
    if (m) {
      const int l = __ffs(m) - 1;
      const unsigned long long before = base + __shfl_sync(0xffffffffu, incl, l) - __shfl_sync(0xffffffffu, (unsigned long long)c, l);
      found = b0 + l;
      rank -= before;
    } else {
      base += __shfl_sync(0xffffffffu, incl, 31);
    }
  }
  if (found < 0) found = BINS - 1;  // unreachable for rank < n
  if (lane == 0) {
    s->rank[k] = rank;
    s->prefix[k] = (s->prefix[k] << BITS) | static_cast<unsigned int>(found);
  }
  if (PASS == 2) {
    __syncthreads();
    if (threadIdx.x < 2) {  // numpy _lerp in float32: a + (b - a) * t, taken from the other end for t >= 0.5
      const float a = key_float(s->prefix[2 * threadIdx.x]), b = key_float(s->prefix[2 * threadIdx.x + 1]);
      const float t = s->frac[threadIdx.x], d = __fsub_rn(b, a);
      const float r = t >= 0.5f ? __fsub_rn(b, __fmul_rn(d, __fsub_rn(1.f, t))) : __fadd_rn(a, __fmul_rn(d, t));
      s->window[threadIdx.x] = static_cast<double>(r);
    }
  }
}

__global__ void __launch_bounds__(256) window_apply_kernel(const float* __restrict__ in, long long n,
                                                           const SelectState* __restrict__ s, float out_min,
                                                           float out_max, float* __restrict__ out) {
  const float wmin = static_cast<float>(s->window[0]), wmax = static_cast<float>(s->window[1]);
  // itk::Functor::IntensityWindowingTransform: factor / offset in RealType (double)
  const double factor = (static_cast<double>(out_max) - static_cast<double>(out_min)) /
                        (static_cast<double>(wmax) - static_cast<double>(wmin));
  const double offset = static_cast<double>(out_min) - static_cast<double>(wmin) * factor;
  auto f = [&](float x) {
    if (x < wmin) return out_min;
    if (x > wmax) return out_max;
    return static_cast<float>(static_cast<double>(x) * factor + offset);
  };
  const long long n4 = n / 4;
  const bool vec = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  if (vec) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(in) + i);
      reinterpret_cast<float4*>(out)[i] = make_float4(f(v.x), f(v.y), f(v.z), f(v.w));
    }
    if (blockIdx.x == 0)
      for (long long i = 4 * n4 + threadIdx.x; i < n; i += blockDim.x) out[i] = f(__ldg(in + i));
  } else {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
      out[i] = f(__ldg(in + i));
  }
}

}  // namespace
}  // namespace oai

using namespace oai;

extern "C" size_t oai_intensity_window_workspace(void) { return sizeof(SelectState); }

extern "C" int oai_intensity_window(const float* in, long long n, double perc_lo, double perc_hi, float out_min,
                                    float out_max, float* out, void* workspace, size_t workspace_bytes,
                                    void* stream) {
  OAI_REQUIRE(in && out && workspace, "intensity_window: null pointer");
  OAI_REQUIRE(n > 0, "intensity_window: empty volume");
  OAI_REQUIRE(perc_lo >= 0.0 && perc_hi <= 100.0 && perc_lo <= perc_hi,
              "intensity_window: percentiles must satisfy 0 <= lo <= hi <= 100 (got %g, %g)", perc_lo, perc_hi);
  OAI_REQUIRE(workspace_bytes >= sizeof(SelectState) && (reinterpret_cast<uintptr_t>(workspace) & 7) == 0,
              "intensity_window: workspace of %zu bytes (8-byte aligned) required", sizeof(SelectState));
  OAI_REQUIRE((reinterpret_cast<uintptr_t>(in) & 15) == 0, "intensity_window: input must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SelectState* s = static_cast<SelectState*>(workspace);
  const unsigned grid = static_cast<unsigned>(num_sms()) * 4;
  window_init_kernel<<<1, 256, 0, st>>>(s, n, perc_lo, perc_hi);
  if (int rc = launched("window_init_kernel")) return rc;
  window_hist_kernel<0><<<grid, 512, 0, st>>>(in, n, s);
  if (int rc = launched("window_hist_kernel<0>")) return rc;
  window_pick_kernel<0><<<1, 32 * kTargets, 0, st>>>(s, out_min, out_max);
  if (int rc = launched("window_pick_kernel<0>")) return rc;
  window_hist_kernel<1><<<grid, 512, 0, st>>>(in, n, s);
  if (int rc = launched("window_hist_kernel<1>")) return rc;
  window_pick_kernel<1><<<1, 32 * kTargets, 0, st>>>(s, out_min, out_max);
  if (int rc = launched("window_pick_kernel<1>")) return rc;
  window_hist_kernel<2><<<grid, 512, 0, st>>>(in, n, s);
  if (int rc = launched("window_hist_kernel<2>")) return rc;
  window_pick_kernel<2><<<1, 32 * kTargets, 0, st>>>(s, out_min, out_max);
  if (int rc = launched("window_pick_kernel<2>")) return rc;
  window_apply_kernel<<<grid * 2, 256, 0, st>>>(in, n, s, out_min, out_max, out);
  return launched("window_apply_kernel");
}

extern "C" int oai_intensity_window_result(const void* workspace, double* window_host, void* stream) {
  OAI_REQUIRE(workspace && window_host, "intensity_window_result: null pointer");
  const SelectState* s = static_cast<const SelectState*>(workspace);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (int rc = check_cuda(cudaMemcpyAsync(window_host, s->window, 2 * sizeof(double), cudaMemcpyDeviceToHost, st),
                          "intensity_window_result: copy"))
    return rc;
  return check_cuda(cudaStreamSynchronize(st), "intensity_window_result: sync");
}
