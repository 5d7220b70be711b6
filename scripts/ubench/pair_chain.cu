// Micro-benchmark: cost of the Jacobi pair primitives with every row in registers (no shared memory, no hand-over).
//   mode 0: tcr::pair2 (lock-step dual pair: dots, transposed reduction, lane-packed set-up, two rotations)
//   mode 1: tcr::setup1 + tcr::dense for two chains in the pipelined order (A.setup || B.dense, B.setup || A.dense)
//   mode 2: dots + reductions only      mode 3: set-up only (after a reduction)   mode 4: two rotations only
// usage: pair_chain <warps per CTA>; prints cycles per dual step (= 2 row pairs) per warp and FP64 pipe share
#include <cstdio>
#include <cstdlib>
#include "../../time_crystal_tensor_network_b200/csrc/tc_jacobi_rb.cuh"
using namespace tcr;
template <int MODE>
__global__ void __launch_bounds__(256, 1) k(double *out, int iters, double seed) {
  cplx uA[8], vA[8], uB[8], vB[8];
  const int lane = threadIdx.x & 31;
  for (int e = 0; e < 8; ++e) {
    uA[e] = make_double2(1.0 + lane * 1e-3 + e * seed, 0.5 + e);
    vA[e] = make_double2(0.25 + lane * 1e-3 - e, 1.5 - e * seed);
    uB[e] = make_double2(0.75 + e + lane * 1e-2, 0.1 * e);
    vB[e] = make_double2(0.33 - e, 0.7 + e * seed + lane * 1e-2);
  }
  double aA = 0, bA = 0, aB = 0, bB = 0;
  for (int e = 0; e < 8; ++e) {
    aA += cabs2(uA[e]); bA += cabs2(vA[e]); aB += cabs2(uB[e]); bB += cabs2(vB[e]);
  }
  aA = tcj::warp_sum(aA); bA = tcj::warp_sum(bA); aB = tcj::warp_sum(aB); bB = tcj::warp_sum(bB);
  double gAr = 0, gAi = 0, gBr = 0, gBi = 0;
  dot_rows<8>(uA, vA, gAr, gAi);
  dot_rows<8>(uB, vB, gBr, gBi);
  RotP rA, rB;
  rA.cs = rB.cs = 0.8; rA.sr = rB.sr = 0.36; rA.si = rB.si = 0.48;
  int acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
      acc += pair2<8>(uA, vA, aA, bA, true, uB, vB, aB, bB, true, 0.0, 1e-40, lane);
    } else if (MODE == 1) {
      acc += setup1(gAr, gAi, aA, bA, true, 0.0, 1e-40, rA);
      dense<8>(uB, vB, rB, uA, gBr, gBi);
      acc += setup1(gBr, gBi, aB, bB, true, 0.0, 1e-40, rB);
      dense<8>(uA, vA, rA, uB, gAr, gAi);
    } else if (MODE == 2) {
      double a, b, c, d;
      dot_rows<8>(uA, vA, a, b);
      dot_rows<8>(uB, vB, c, d);
      for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(FULLM, a, o); b += __shfl_xor_sync(FULLM, b, o);
        c += __shfl_xor_sync(FULLM, c, o); d += __shfl_xor_sync(FULLM, d, o);
      }
      uA[it & 7].x += 1e-300 * (a + b + c + d);
    } else if (MODE == 3) {
      double a = gAr * 1e-20, b = gAi * 1e-20;   // no shuffles: plain set-up chain
      const double g2 = fma(a, a, b * b), dd = bA - aA;
      const double rinv = rsqrt_nb(fma(dd, dd, 4.0 * g2));
      const double c2 = fma(0.5 * fabs(dd), rinv, 0.5);
      const double cinv = rsqrt_nb(c2);
      const double ks = copysign(rinv * cinv, dd);
      gAr = ks * a + c2 * cinv; gAi = ks * b; aA += 1e-30 * g2 * ks * cinv;
    } else {
      rot_rows<8>(uA, vA, rA.cs, rA.sr, rA.si);
      rot_rows<8>(uB, vB, rB.cs, rB.sr, rB.si);
    }
  }
  long long t1 = clock64();
  double s = gAr + gAi + gBr + gBi + aA + bA + aB + bB + acc;
  for (int e = 0; e < 8; ++e) s += uA[e].x + uA[e].y + vA[e].x + vA[e].y + uB[e].x + uB[e].y + vB[e].x + vB[e].y;
  if (s == 123.456) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[1] = (double)(t1 - t0);
}
int main(int argc, char **argv) {
  int warps = argc > 1 ? atoi(argv[1]) : 8;
  int iters = 4000;
  double *out;
  cudaMalloc(&out, 64);
  int sms;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (int mode = 0; mode < 5; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      if (mode == 0) k<0><<<sms, warps * 32>>>(out, iters, 0.37);
      if (mode == 1) k<1><<<sms, warps * 32>>>(out, iters, 0.37);
      if (mode == 2) k<2><<<sms, warps * 32>>>(out, iters, 0.37);
      if (mode == 3) k<3><<<sms, warps * 32>>>(out, iters, 0.37);
      if (mode == 4) k<4><<<sms, warps * 32>>>(out, iters, 0.37);
      cudaDeviceSynchronize();
    }
    double h[2];
    cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("warps %2d mode %d: %.0f clk per dual step per warp; SM-wide %.1f clk per row pair\n", warps, mode, h[1] / iters,
           h[1] / iters / 2.0 / warps);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
