// Micro-benchmark: issue rate of the legacy warp-level mma.sync on sm_100a (m16n8k8 tf32 and m16n8k16 f16, fp32
// accumulate), as a function of warps per SM and independent accumulators per warp.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 scripts/micro/mma_sync_rate.cu -o scripts/micro/mma_sync_rate
#include <cstdio>
#include <cstdint>

template <int KIND, int NACC>
__global__ void rate_kernel(int iters, float* out, long long* cyc) {
  float c[NACC][4];
#pragma unroll
  for (int j = 0; j < NACC; ++j) c[j][0] = c[j][1] = c[j][2] = c[j][3] = 0.f;
  uint32_t a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = 7, a3 = 11, b0 = 5, b1 = 9;
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < NACC; ++j) {
      if (KIND == 0) {
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3])
                     : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      } else {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3])
                     : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < NACC; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int KIND, int NACC>
void run(int warps, float* out, long long* cyc) {
  const int iters = 4096;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  rate_kernel<KIND, NACC><<<148, warps * 32>>>(iters, out, cyc);
  cudaEventRecord(e0);
  rate_kernel<KIND, NACC><<<148, warps * 32>>>(iters, out, cyc);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  long long h = 0; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double mmas = 148.0 * warps * iters * NACC;
  const double macs = mmas * 16 * 8 * (KIND == 0 ? 8 : 16);
  printf("%s warps/SM=%2d acc=%d : %.3f ms, %.1f dense TFLOP/s, %.2f clk per MMA per SMSP\n", KIND == 0 ? "tf32 m16n8k8 " : "f16  m16n8k16",
         warps, NACC, ms, 2 * macs / ms / 1e9, double(h) / (double(iters) * NACC * (warps > 4 ? warps / 4.0 : 1.0)));
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  run<0, 4>(4, out, cyc); run<0, 8>(4, out, cyc); run<0, 8>(8, out, cyc); run<0, 8>(16, out, cyc);
  run<1, 4>(4, out, cyc); run<1, 8>(4, out, cyc); run<1, 8>(8, out, cyc); run<1, 8>(16, out, cyc);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
