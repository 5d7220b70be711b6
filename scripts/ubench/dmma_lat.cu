// dmma_lat.cu -- dependent-issue latency of DMMA (mma.sync.m8n8k4.f64) against the SHFL.BFLY + DADD step of a warp
// butterfly, one warp alone on an SM and with 16 warps per SM running the same chain.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_lat dmma_lat.cu && ./dmma_lat
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void k(long long *out, int iters, int mode) {
  double c0 = threadIdx.x * 1e-9, c1 = 0.0, x = 1.0 + threadIdx.x * 1e-12;
  const long long t0 = clock64();
  if (mode == 0) {
    for (int i = 0; i < iters; ++i) dmma(c0, c1, c0, x);  // the A operand depends on the previous accumulator
  } else if (mode == 1) {
    for (int i = 0; i < iters; ++i) c0 += __shfl_xor_sync(0xffffffffu, c0, 1 << (i % 5));
  } else {
    for (int i = 0; i < iters; ++i) c0 = fma(c0, x, c1);
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  if (c0 + c1 == 1234.5) out[1] = 1;
}
int main() {
  long long *d, h;
  cudaMalloc(&d, 16);
  const char *names[] = {"DMMA -> DMMA (A operand from the accumulator)", "SHFL.BFLY + DADD", "DFMA -> DFMA"};
  for (int warps : {1, 16})
    for (int mode = 0; mode < 3; ++mode) {
      k<<<148, warps * 32>>>(d, 4096, mode);
      cudaDeviceSynchronize();
      cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      printf("%2d warps/SM  %-48s %.1f cycles per step\n", warps, names[mode], h / 4096.0);
    }
  return 0;
}
