// Micro-benchmark: pacing of back-to-back tcgen05.mma (kind::f16, cta_group::1, M=128) as a function of N and of
// whether consecutive MMAs hit the same accumulator.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
// -I oai_analysis_2_b200/csrc scripts/micro/mma_pace.cu -o /tmp/mma_pace ; run on a B200.
#include <cstdio>
#include <cstdlib>
#include "ptx.cuh"
using namespace oai;

__global__ void __launch_bounds__(128, 1) pace_kernel(int N, int iters, int alt, int kstep, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (64 + 32) * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x < 32) {
    const uint32_t a = smem_u32(smem), b = smem_u32(smem + 32 * 1024);
    const uint32_t idesc = umma_idesc_f16(128, N, 0);
    long long t0 = 0, t1 = 0;
    for (int rep = 0; rep < 2; ++rep) {
      t0 = clock64();
      if (elect_one()) {
        for (int i = 0; i < iters; ++i) {
          const uint32_t d = tm + ((alt && (i & 1)) ? 256u : 0u);
          const uint32_t ko = (kstep ? (i & 3) * 32 : 0);
          umma_f16_ss(d, umma_desc_sw128(a + ko, 1024, 0), umma_desc_sw128(b + ko, 1024, 0), idesc, 1u);
        }
        umma_commit(&bar);
      }
      __syncwarp();
      mbar_wait(&bar, rep & 1);
      t1 = clock64();
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) *out = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  cudaFuncSetAttribute(pace_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 2048;
  for (int grid : {1, 148}) {
    for (int N : {64, 128, 192, 256}) {
      for (int alt : {0, 1}) {
        if (alt && N > 256) continue;
        for (int kstep : {0, 1}) {
          pace_kernel<<<grid, 128, 100 * 1024>>>(N, iters, alt, kstep, d);
          long long h = 0;
          cudaError_t e = cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          printf("grid %3d N %3d alt %d kstep %d : %.1f clk/MMA (floor %d)\n", grid, N, alt, kstep, double(h) / iters,
                 N / 2);
        }
      }
    }
  }
  return 0;
}
