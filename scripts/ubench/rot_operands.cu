// Micro-benchmark: does the DFMA rate of the row rotation depend on where the rotation coefficients live?
//   src 0: kernel parameters (constant bank operands)      src 1: per-thread registers (loaded from global)
//   src 2: registers broadcast with __shfl_sync(x, 0) (provably warp-uniform -> uniform registers?)
// usage: rot_operands <warps per CTA>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef double2 cplx;
template <int SRC>
__global__ void k(double *out, const double *coef, int iters, double pcs, double psr, double psi) {
  cplx u[8], v[8];
  for (int e = 0; e < 8; ++e) {
    u[e] = make_double2(1.0 + threadIdx.x * 1e-3 + e, 0.5 + e);
    v[e] = make_double2(0.25 + threadIdx.x * 1e-3 - e, 1.5 - e);
  }
  double cs = pcs, sr = psr, si = psi;
  if (SRC >= 1) {
    cs = coef[threadIdx.x & 3];
    sr = coef[4 + (threadIdx.x & 3)];
    si = coef[8 + (threadIdx.x & 3)];
  }
  if (SRC == 2) {
    cs = __shfl_sync(0xffffffffu, cs, 0);
    sr = __shfl_sync(0xffffffffu, sr, 0);
    si = __shfl_sync(0xffffffffu, si, 0);
  }
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      cplx un, vn;
      un.x = fma(cs, u[e].x, fma(-sr, v[e].x, si * v[e].y));
      un.y = fma(cs, u[e].y, -fma(sr, v[e].y, si * v[e].x));
      vn.x = fma(cs, v[e].x, fma(sr, u[e].x, si * u[e].y));
      vn.y = fma(cs, v[e].y, fma(sr, u[e].y, -si * u[e].x));
      u[e] = un;
      v[e] = vn;
    }
  }
  long long t1 = clock64();
  double s = 0;
  for (int e = 0; e < 8; ++e) s += u[e].x + u[e].y + v[e].x + v[e].y;
  if (s == 123.456) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[1] = (double)(t1 - t0);
}
int main(int argc, char **argv) {
  int warps = argc > 1 ? atoi(argv[1]) : 8;
  int iters = 20000;
  double *out, *coef;
  cudaMalloc(&out, 64);
  cudaMalloc(&coef, 12 * 8);
  double h_c[12] = {0.999, 0.999, 0.999, 0.999, 1e-3, 1e-3, 1e-3, 1e-3, 2e-3, 2e-3, 2e-3, 2e-3};
  cudaMemcpy(coef, h_c, sizeof(h_c), cudaMemcpyHostToDevice);
  int sms;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (int src = 0; src < 3; ++src) {
    for (int rep = 0; rep < 2; ++rep) {
      if (src == 0) k<0><<<sms, warps * 32>>>(out, coef, iters, 0.999, 1e-3, 2e-3);
      if (src == 1) k<1><<<sms, warps * 32>>>(out, coef, iters, 0.999, 1e-3, 2e-3);
      if (src == 2) k<2><<<sms, warps * 32>>>(out, coef, iters, 0.999, 1e-3, 2e-3);
      cudaDeviceSynchronize();
    }
    double h[2];
    cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    double instr = 96.0 * iters * warps;
    printf("warps %2d coefficients from %s: %.3f FP64 warp-instr/clk/SM (%.1f %% of 1.87)\n", warps,
           src == 0 ? "kernel params" : src == 1 ? "registers    " : "shfl-uniform ", instr / h[1], 100.0 * instr / h[1] / 1.87);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
