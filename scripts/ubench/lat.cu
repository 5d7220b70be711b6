// micro-benchmarks: dependent-issue latency of the FP64 building blocks of the Jacobi kernel on B200
#include <cstdio>
#include <cuda_runtime.h>
#define N 2048
template <int OP> __global__ void k(double *out, long long *cyc, double x0, float f0) {
  double x = x0 + threadIdx.x * 1e-9, y = 1.0000001, z = 0.5;
  float f = f0 + threadIdx.x * 1e-6f;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) {
    if (OP == 0) x = fma(x, y, z);
    if (OP == 1) x = rsqrt(x) + 1.0;
    if (OP == 2) x = sqrt(x) + 1.0;
    if (OP == 3) x = 1.0 / x + 1.0;
    if (OP == 4) x += __shfl_xor_sync(0xffffffffu, x, 1);
    if (OP == 5) { f = rsqrtf(f) + 1.0f; }
    if (OP == 6) { f = (float)x; x = (double)f + 1.0; }
    if (OP == 7) x = x * y;
    if (OP == 8) x = x + z;
    if (OP == 9) { f = __frcp_rn(f) + 1.0f; }
    if (OP == 10) { f = sqrtf(f) + 1.0f; }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = x + f;
}
// throughput: 8 independent DFMA chains per thread, W warps per block
__global__ void thr(double *out, long long *cyc) {
  double a[8];
  for (int j = 0; j < 8; ++j) a[j] = j + threadIdx.x * 1e-9;
  long long t0 = clock64();
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = fma(a[j], 1.0000001, 0.5);
  long long t1 = clock64();
  double s = 0; for (int j = 0; j < 8; ++j) s += a[j];
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  double *out; long long *cyc, h;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
  const char *names[] = {"DFMA dep", "rsqrt(double)+add dep", "sqrt(double)+add dep", "1/x double +add dep", "SHFL+DADD dep",
                         "rsqrtf+add dep", "cvt f64->f32->f64 + add dep", "DMUL dep", "DADD dep", "__frcp_rn+add", "sqrtf+add"};
#define RUN(OP) k<OP><<<1, 32>>>(out, cyc, 1.5, 1.5f); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-32s %7.1f clk/iter (1 warp)\n", names[OP], (double)h / N);
  RUN(0) RUN(7) RUN(8) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(9) RUN(10)
  for (int w = 1; w <= 16; w *= 2) {
    thr<<<1, 32 * w>>>(out, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DFMA 8-way ILP, %2d warps/SM: %6.2f clk per warp-DFMA per warp; SM rate %.1f DFMA lanes/clk\n", w, (double)h / (N * 8), 32.0 * w * N * 8 / h);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
