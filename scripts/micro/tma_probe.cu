// Probe: which cuTensorMapEncodeTiled configurations of a 32-bit (W,H,D,plane,1) tensor can cp.async.bulk.tensor.5d
// load with a box whose innermost start coordinate is not 16-byte aligned?  One configuration per process.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I oai_analysis_2_b200/csrc scripts/micro/tma_probe.cu -o scripts/micro/tma_probe
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "ptx.cuh"
using namespace oai;

__global__ void probe(const __grid_constant__ CUtensorMap tm, int bytes, int cx, int cy, int cz, uint32_t* out, int nwords) {
  extern __shared__ uint8_t raw[];
  uint32_t* s = reinterpret_cast<uint32_t*>((reinterpret_cast<uintptr_t>(raw) + 127) & ~uintptr_t(127));
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar, bytes);
    tma_load_5d(s, &tm, &bar, cx, cy, cz, 0, 0);
  }
  mbar_wait(&bar, 0, 1);
  for (int i = threadIdx.x; i < nwords; i += blockDim.x) out[i] = s[i];
}

int main(int argc, char** argv) {
  const int bx = argc > 1 ? atoi(argv[1]) : 36, l2 = argc > 2 ? atoi(argv[2]) : 2, dt = argc > 3 ? atoi(argv[3]) : 0;
  const int cx = argc > 4 ? atoi(argv[4]) : -1, W = argc > 5 ? atoi(argv[5]) : 40;
  const int H = 8, D = 3, P = 16, by = 6, bz = 3, bp = 8;
  std::vector<uint32_t> h((size_t)P * D * H * W);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (uint32_t)i + 1;
  uint32_t *d, *o;
  cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  const int nwords = bx * by * bz * bp;
  cudaMalloc(&o, nwords * 4);
  typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                         const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                         CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* q; cudaDriverEntryPointQueryResult r;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &r);
  Fn fn = (Fn)q;
  CUtensorMap tm;
  cuuint64_t dims[5] = {(cuuint64_t)W, H, D, P, 1};
  cuuint64_t strides[4] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)D * H * W * 4, (cuuint64_t)P * D * H * W * 4};
  cuuint32_t box[5] = {(cuuint32_t)bx, by, bz, bp, 1}, es[5] = {1, 1, 1, 1, 1};
  CUresult e = fn(&tm, dt == 0 ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, d, dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)l2,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("bx=%d l2=%d dt=%d cx=%d W=%d encode=%d ", bx, l2, dt, cx, W, (int)e);
  if (e) { printf("\n"); return 0; }
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, nwords * 4 + 256);
  probe<<<1, 128, nwords * 4 + 256>>>(tm, nwords * 4, cx, -1, -1, o, nwords);
  cudaError_t ce = cudaDeviceSynchronize();
  printf("run=%s ", cudaGetErrorString(ce));
  if (ce == cudaSuccess) {
    std::vector<uint32_t> g(nwords);
    cudaMemcpy(g.data(), o, nwords * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int p = 0; p < bp; ++p) for (int z = 0; z < bz; ++z) for (int y = 0; y < by; ++y) for (int x = 0; x < bx; ++x) {
      const int gx = cx + x, gy = -1 + y, gz = -1 + z;
      const uint32_t want = (gx < 0 || gx >= W || gy < 0 || gy >= H || gz < 0 || gz >= D) ? 0u : h[(((size_t)p * D + gz) * H + gy) * W + gx];
      if (g[((p * bz + z) * by + y) * bx + x] != want) ++bad;
    }
    printf("mismatches=%d", bad);
  }
  printf("\n");
  return 0;
}
