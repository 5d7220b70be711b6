// tc_theta.cuh -- K1: C = gate . (kick? B_i)(kick? B_{i+1}), theta = S_i C for every bond of a layer
// (reference op: src/models/kicked_ising.py:128-160 -> _apply_two_site_gate / _apply_pi_pulse).
//
// Batched ragged complex-FP64 GEMM on the FP64 tensor pipe (DMMA, mma.sync.m8n8k4.f64), specialised for the layout of
// the site tensors so that the 2x2 kick costs no extra global traffic:
//   A operand  (a, p0) x m : B_i is stored [chi_l][2][chi_m], the rows (a,0), (a,1) are neighbours.  A thread stages a
//              ROW PAIR (4 consecutive m of both rows, 64 B each), applies the kick to the pair in registers and writes
//              both rows of the shared-memory tile.
//   B operand  m x (p1, b) : B_{i+1} is stored [chi_m][2][chi_r]; for one m the two p1 halves are chi_r apart.  A CTA
//              tile therefore covers 32 values of b for BOTH p1 (64 columns): a thread loads 4 consecutive b of both
//              halves once, applies the kick and writes the p1 = 0 and p1 = 1 columns of the tile.
// The next k-slab is fetched into registers while the current one is multiplied (one __syncthreads pair per slab, global
// latency hidden behind 64 DMMAs per warp and slab).  Epilogue: diagonal Ising phase and S_i scaling fused, theta
// written with interleaved columns (2 b + p1); a general 4x4 gate leaves the raw product for gate_mix_kernel.
// CTA tile 64 x 64 x 16, 4 warps (2 x 2), warp tile 32 x 32 = 4 x 4 DMMA tiles, complex product = 4 real DMMAs.
#pragma once
#include "tc_common.cuh"
#include "tc_gemm.cuh"

namespace tch {
constexpr int BM = 64, BNB = 32, BN = 2 * BNB, BK = 16, LDA = BK + 4, LDB = BN + 4, NT = 128;

__global__ void __launch_bounds__(NT) theta_gemm_kernel(TcDev d, LayerArgs a) {
  Bond b;
  if (!get_bond(d, a, blockIdx.y, blockIdx.z, b)) return;
  const int M = b.M, K = b.chiM, chiR = b.chiR;
  const int tiles_n = (chiR + BNB - 1) / BNB, tiles_m = (M + BM - 1) / BM;
  const int t = blockIdx.x;
  if (t >= tiles_m * tiles_n) return;
  const int row0 = (t / tiles_n) * BM, b0 = (t % tiles_n) * BNB;
  const cplx *Bi = site_ptr(d, b.r, b.i), *Bn = site_ptr(d, b.r, b.i + 1);
  const bool kickL = (a.kick_mode & 1) != 0;
  const bool kickR = kickL || ((a.kick_mode & 2) && b.i == d.L - 2);
  const cplx *kk = d.kick + (size_t)b.r * 4;
  const cplx k00 = kk[0], k01 = kk[1], k10 = kk[2], k11 = kk[3];

  __shared__ __align__(16) cplx As[BM * LDA];  // [row][k]
  __shared__ __align__(16) cplx Bs[BK * LDB];  // [k][col], col = p1 * 32 + (b - b0)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 1, wn = warp & 1;
  const int fr = lane >> 2, fk = lane & 3;
  // staging roles: A -- row pair apr (rows 2 apr, 2 apr + 1 of the tile), k quad akq;  B -- k row bk, b quad bbq
  const int apr = tid >> 2, akq = tid & 3;
  const int bk = tid >> 3, bbq = tid & 7;
  const int arow = row0 + 2 * apr;  // even: (a, p0 = 0); M is even, so the pair is inside or outside together
  const cplx zero = cmake(0.0, 0.0);

  cplx ra0[4], ra1[4], rb0[4], rb1[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + 4 * akq + j;
      const bool ok = arow < M && k < K;
      ra0[j] = ok ? Bi[(size_t)arow * K + k] : zero;
      ra1[j] = ok ? Bi[(size_t)(arow + 1) * K + k] : zero;
    }
    const int km = k0 + bk;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int bb = b0 + 4 * bbq + j;
      const bool ok = km < K && bb < chiR;
      rb0[j] = ok ? Bn[(size_t)(2 * km) * chiR + bb] : zero;
      rb1[j] = ok ? Bn[(size_t)(2 * km + 1) * chiR + bb] : zero;
    }
  };
  auto stage = [&]() {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      cplx y0 = ra0[j], y1 = ra1[j];
      if (kickL) {
        y0 = cmul(k00, ra0[j]);
        cfma(y0, k01, ra1[j]);
        y1 = cmul(k10, ra0[j]);
        cfma(y1, k11, ra1[j]);
      }
      As[(2 * apr) * LDA + 4 * akq + j] = y0;
      As[(2 * apr + 1) * LDA + 4 * akq + j] = y1;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      cplx y0 = rb0[j], y1 = rb1[j];
      if (kickR) {
        y0 = cmul(k00, rb0[j]);
        cfma(y0, k01, rb1[j]);
        y1 = cmul(k10, rb0[j]);
        cfma(y1, k11, rb1[j]);
      }
      Bs[bk * LDB + 4 * bbq + j] = y0;
      Bs[bk * LDB + BNB + 4 * bbq + j] = y1;
    }
  };

  double cre[4][4][2], cim[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) cre[i][j][0] = cre[i][j][1] = cim[i][j][0] = cim[i][j][1] = 0.0;

  fetch(0);
  for (int k0 = 0; k0 < K; k0 += BK) {
    stage();
    __syncthreads();
    if (k0 + BK < K) fetch(k0 + BK);  // in flight while this slab is multiplied
#pragma unroll
    for (int ks = 0; ks < BK; ks += 4) {
      double are[4], aim[4], bre[4], bim[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const cplx v = As[(wm * 32 + i * 8 + fr) * LDA + ks + fk];
        are[i] = v.x;
        aim[i] = v.y;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const cplx v = Bs[(ks + fk) * LDB + wn * 32 + j * 8 + fr];
        bre[j] = v.x;
        bim[j] = v.y;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          tcg::dmma(cre[i][j][0], cre[i][j][1], are[i], bre[j]);
          tcg::dmma(cim[i][j][0], cim[i][j][1], are[i], bim[j]);
          tcg::dmma(cre[i][j][0], cre[i][j][1], -aim[i], bim[j]);
          tcg::dmma(cim[i][j][0], cim[i][j][1], aim[i], bre[j]);
        }
    }
    __syncthreads();
  }
  // ---- epilogue: warp column half wn is p1 (tile columns 0..31 are p1 = 0, 32..63 are p1 = 1)
  const int N = b.N;
  cplx *C = d.Cw + b.slot * d.slot_stride, *X = d.Xw + b.slot * d.slot_stride;
  const double *S = S_ptr(d, b.r, b.i);
  const bool diag = a.diag != 0;
  const cplx *g = a.gate_override ? a.gate_override : d.gates + ((size_t)b.r * (d.L - 1) + b.i) * 16;
  const int p1 = wn;
  const cplx ph0 = diag ? g[(0 * 2 + p1) * 5] : zero, ph1 = diag ? g[(1 * 2 + p1) * 5] : zero;  // diagonal entries (p0, p1)
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = row0 + wm * 32 + i * 8 + fr;
    if (row >= M) continue;
    const cplx ph = (row & 1) ? ph1 : ph0;
    const double s = S[row >> 1];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int bb = b0 + j * 8 + 2 * fk + e;
        if (bb >= chiR) continue;
        const cplx v = cmake(cre[i][j][e], cim[i][j][e]);
        const size_t o = (size_t)row * N + p1 * chiR + bb;
        if (diag) {
          const cplx c = cmul(ph, v);
          C[o] = c;
          X[(size_t)row * N + 2 * bb + p1] = cscale(c, s);  // interleaved columns
        } else {
          C[o] = v;
        }
      }
  }
}
}  // namespace tch
