// Micro-benchmark: cost of the Jacobi pair primitives with every row in registers (no shared memory, no hand-over).
//   mode 0: tcr::pair2 (lock-step dual pair: dots, transposed reduction, lane-packed set-up, two rotations)
//   mode 1: tcr::setup1 + tcr::dense for two chains in the pipelined order (A.setup || B.dense, B.setup || A.dense)
//   mode 2: dots + reductions only      mode 3: set-up only (after a reduction)   mode 4: two rotations only
//   mode 5: ONE pair per step as in the 16-warp kernel (dot, butterfly reduction, tcb::make_rot, rotation); build with
//           -DWARPS16 (128 registers) and run with 16 warps
// usage: pair_chain <warps per CTA>; prints cycles per dual step (= 2 row pairs) per warp and FP64 pipe share
#include <cstdio>
#include <cstdlib>
#include "../experiments/tc_jacobi_rb.cuh"
using namespace tcr;
// primitives of the software-pipelined visit that was measured and dropped (see tc_jacobi_rb.cuh header)
struct RotP {
  double cs, sr, si;
};
__device__ __forceinline__ int setup1(double gr, double gi, double &ai, double &aj, bool act, double dead, double tol2,
                                      RotP &r) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    gr += __shfl_xor_sync(FULLM, gr, o);
    gi += __shfl_xor_sync(FULLM, gi, o);
  }
  const double g2 = fma(gr, gr, gi * gi), thr = ai * aj;
  const bool rot = act && ai > dead && aj > dead && (g2 > tol2 * thr);
  const double dd = aj - ai;
  const double rinv = rsqrt_nb(rot ? fma(dd, dd, 4.0 * g2) : 1.0);
  const double c2 = fma(0.5 * fabs(dd), rinv, 0.5);
  const double cinv = rsqrt_nb(c2);
  const double ks = copysign(rinv * cinv, dd);
  r.cs = rot ? c2 * cinv : 1.0;
  r.sr = rot ? ks * gr : 0.0;
  r.si = rot ? ks * gi : 0.0;
  const double tg = rot ? g2 * ks * cinv : 0.0;
  ai -= tg;
  aj += tg;
  return (int)rot;
}
template <int NPL>
__device__ __forceinline__ void dense(cplx (&u)[NPL], cplx (&v)[NPL], const RotP &r, const cplx (&w)[NPL], double &gr,
                                      double &gi) {
  double g0 = 0.0, g1 = 0.0, h0 = 0.0, h1 = 0.0;
#pragma unroll
  for (int e = 0; e < NPL; ++e) {
    cplx un, vn;
    un.x = fma(r.cs, u[e].x, fma(-r.sr, v[e].x, r.si * v[e].y));
    un.y = fma(r.cs, u[e].y, -fma(r.sr, v[e].y, r.si * v[e].x));
    vn.x = fma(r.cs, v[e].x, fma(r.sr, u[e].x, r.si * u[e].y));
    vn.y = fma(r.cs, v[e].y, fma(r.sr, u[e].y, -r.si * u[e].x));
    u[e] = un;
    v[e] = vn;
    g0 = fma(w[e].x, vn.x, g0);
    g1 = fma(w[e].y, vn.y, g1);
    h0 = fma(w[e].y, vn.x, h0);
    h1 = fma(-w[e].x, vn.y, h1);
  }
  gr = g0 + g1;
  gi = h0 + h1;
}
#ifdef WARPS16
#define LB 512
#else
#define LB 256
#endif
template <int MODE>
__global__ void __launch_bounds__(LB, 1) k(double *out, int iters, double seed) {
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
#ifdef STAGGER
  {  // de-phase the warps: warp w starts w * STAGGER cycles late
    const long long tw = clock64() + (long long)(threadIdx.x >> 5) * STAGGER;
    while (clock64() < tw) {}
  }
#endif
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
    } else if (MODE == 5) {
      double gr, gi;
      dot_rows<8>(uA, vA, gr, gi);
      tcb::warp_sum2(gr, gi);
      tcb::Rot r;
      if (tcb::make_rot(true, aA, bA, gr, gi, 1e-40, r)) {
        rot_rows<8>(uA, vA, r.cs, r.sr, r.si);
        aA = r.ni;
        bA = r.nj;
      }
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
#ifdef WARPS16
  for (int mode = 5; mode < 6; ++mode) {
#else
  for (int mode = 0; mode < 6; ++mode) {
#endif
    for (int rep = 0; rep < 2; ++rep) {
      if (mode == 0) k<0><<<sms, warps * 32>>>(out, iters, 0.37);
      if (mode == 1) k<1><<<sms, warps * 32>>>(out, iters, 0.37);
      if (mode == 2) k<2><<<sms, warps * 32>>>(out, iters, 0.37);
      if (mode == 3) k<3><<<sms, warps * 32>>>(out, iters, 0.37);
      if (mode == 4) k<4><<<sms, warps * 32>>>(out, iters, 0.37);
      if (mode == 5) k<5><<<sms, warps * 32>>>(out, iters, 0.37);
      cudaDeviceSynchronize();
    }
    double h[2];
    cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    const double pairs = mode == 5 ? 1.0 : 2.0;
    printf("warps %2d mode %d: %.0f clk per step per warp; SM-wide %.1f clk per row pair\n", warps, mode, h[1] / iters,
           h[1] / iters / pairs / warps);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
