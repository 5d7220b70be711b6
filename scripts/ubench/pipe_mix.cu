// pipe_mix.cu -- do the FP64 FMA pipe (DFMA) and the FP64 tensor path (DMMA, mma.sync.m8n8k4.f64) of sm_100a share one
// throughput limit, or can a kernel that issues both exceed either peak?  Each CTA has nf warps of independent DFMA
// chains and nd warps of independent DMMA chains; the table gives TFLOP/s of each kind and their sum.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_mix pipe_mix.cu && ./pipe_mix
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(1024) mix_kernel(double *out, int iters, int nf, int nd) {
  const int warp = threadIdx.x >> 5;
  const double x = 1.0000001, y = 1e-9 * threadIdx.x;
  double s = 0.0;
  if (warp < nf) {
    double a[8];
    for (int k = 0; k < 8; ++k) a[k] = k + y;
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int k = 0; k < 8; ++k) a[k] = fma(a[k], x, y);
    for (int k = 0; k < 8; ++k) s += a[k];
  } else if (warp < nf + nd) {
    double c[8][2];
    for (int k = 0; k < 8; ++k) c[k][0] = c[k][1] = 0.0;
    // one DMMA = 256 FMA = 8 warp-wide DFMA instructions: iters / 8 keeps the two kinds of warps busy equally long
    // when the pipes are independent
    for (int it = 0; it < iters / 8; ++it)
#pragma unroll
      for (int k = 0; k < 8; ++k) dmma(c[k][0], c[k][1], x, y);
    for (int k = 0; k < 8; ++k) s += c[k][0] + c[k][1];
  }
  if (s == 123.456) out[0] = s;
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  double *out;
  cudaMalloc(&out, 8);
  const int iters = 40000;
  const int cfg[][2] = {{16, 0}, {0, 16}, {8, 8}, {16, 16}, {32, 0}, {0, 32}, {12, 4}, {4, 12}, {16, 4}, {16, 8}, {8, 16}, {4, 4}, {8, 0}, {0, 8}};
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  printf("%s, %d SMs\n nf nd   ms    DFMA TF  DMMA TF  sum TF\n", prop.name, prop.multiProcessorCount);
  for (auto &c : cfg) {
    const int nf = c[0], nd = c[1], nt = (nf + nd) * 32;
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
      cudaEventRecord(e0);
      mix_kernel<<<prop.multiProcessorCount, nt>>>(out, iters, nf, nd);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep > 0 && ms < best) best = ms;
    }
    const double blocks = prop.multiProcessorCount;
    const double f_fma = blocks * nf * 32.0 * iters * 8 * 2.0, f_dmma = blocks * nd * (iters / 8) * 8 * 512.0;
    printf("%3d %3d %7.3f %8.2f %8.2f %8.2f\n", nf, nd, best, f_fma / best * 1e-9, f_dmma / best * 1e-9,
           (f_fma + f_dmma) / best * 1e-9);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
