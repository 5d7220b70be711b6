// Micro-benchmark: how fast can one SM stream the Jacobi inner loops when everything is in registers?
//   mode 0: pure DFMA chains a = fma(a, x, y), ILP 8            (the peak the roofline quotes)
//   mode 1: rotation of two complex rows (12 FP64 ops per element pair, 8 elements per lane)
//   mode 2: dot product of two complex rows (4 DFMA per element pair)
//   mode 3: rotation fused with the next dot (the pipelined `dense`)
// Launch: rot_rate <warps per CTA> ; one CTA per SM, prints FP64 instructions per clock per SM (peak 0.5 x 32 lanes).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef double2 cplx;
template <int MODE>
__global__ void k(double *out, int iters, double cs, double sr, double si) {
  cplx u[8], v[8], w[8];
  for (int e = 0; e < 8; ++e) {
    u[e] = make_double2(1.0 + threadIdx.x * 1e-3 + e, 0.5 + e);
    v[e] = make_double2(0.25 + threadIdx.x * 1e-3 - e, 1.5 - e);
    w[e] = make_double2(0.75 + e, 0.1 * e);
  }
  double g0 = 0, g1 = 0, h0 = 0, h1 = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        u[e].x = fma(u[e].x, cs, sr);
        u[e].y = fma(u[e].y, cs, si);
        v[e].x = fma(v[e].x, cs, sr);
        v[e].y = fma(v[e].y, cs, si);
      }
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        if (MODE == 1 || MODE == 3) {
          cplx un, vn;
          un.x = fma(cs, u[e].x, fma(-sr, v[e].x, si * v[e].y));
          un.y = fma(cs, u[e].y, -fma(sr, v[e].y, si * v[e].x));
          vn.x = fma(cs, v[e].x, fma(sr, u[e].x, si * u[e].y));
          vn.y = fma(cs, v[e].y, fma(sr, u[e].y, -si * u[e].x));
          u[e] = un;
          v[e] = vn;
        }
        if (MODE == 2 || MODE == 3) {
          g0 = fma(w[e].x, v[e].x, g0);
          g1 = fma(w[e].y, v[e].y, g1);
          h0 = fma(w[e].y, v[e].x, h0);
          h1 = fma(-w[e].x, v[e].y, h1);
        }
      }
      if (MODE == 2) {  // keep the loop from being hoisted: perturb one operand
        w[it & 7].x += g0 * 1e-300;
      }
    }
  }
  long long t1 = clock64();
  double s = g0 + g1 + h0 + h1;
  for (int e = 0; e < 8; ++e) s += u[e].x + u[e].y + v[e].x + v[e].y;
  if (s == 123.456) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[1] = (double)(t1 - t0);
}
int main(int argc, char **argv) {
  int warps = argc > 1 ? atoi(argv[1]) : 16;
  int iters = 20000;
  double *out;
  cudaMalloc(&out, 64);
  int sms;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int per_iter[4] = {32, 96, 32, 128};
  for (int mode = 0; mode < 4; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      if (mode == 0) k<0><<<sms, warps * 32>>>(out, iters, 0.999, 1e-3, 2e-3);
      if (mode == 1) k<1><<<sms, warps * 32>>>(out, iters, 0.999, 1e-3, 2e-3);
      if (mode == 2) k<2><<<sms, warps * 32>>>(out, iters, 0.999, 1e-3, 2e-3);
      if (mode == 3) k<3><<<sms, warps * 32>>>(out, iters, 0.999, 1e-3, 2e-3);
      cudaDeviceSynchronize();
    }
    double h[2];
    cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    double instr = (double)per_iter[mode] * iters * warps;
    printf("warps %2d mode %d: %.0f clk, %.3f FP64 warp-instr/clk/SM (%.1f %% of 2/clk... peak 1.87)\n", warps, mode, h[1],
           instr / h[1], 100.0 * instr / h[1] / 1.87);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
