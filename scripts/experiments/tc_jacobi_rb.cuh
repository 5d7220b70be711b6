// NOTE (round 2): this experiment was written against the round-1 helpers of tc_jacobi_blocked.cuh (standard rotations:
// tcb::make_rot(alive, ai, aj, gr, gi, ...), Rot{cs, sr, si, ni, nj}); the product kernel now uses fast scaled rotations with a
// different Rot / make_rot, so this file and scripts/ubench/pair_chain.cu build only against commit 4557cd6 (round-1 final).
// Kept for the record of the negative result (second tuning pass).
// tc_jacobi_rb.cuh -- K2b, register-blocked variant (TC_JACOBI=rb; the 16-warp kernel of tc_jacobi_blocked.cuh is
// the default): one-sided Jacobi on the rows of the triangular factor with FOUR stationary rows per warp in registers
// and every streamed row reused four times per shared-memory load.
//
// Design (vs ncu of the 16-warp kernel, profiles/r01_*: one stationary row per warp costs 4 KB LDS + 4 KB STS per row
// pair -- shared-memory pipe 52 %, FP64 pipe 45 %, and every pair pays its own shuffle reduction and set-up chain):
//   * a CTA of 8 warps (255 registers) holds a P block of 32 rows in registers, 4 per warp; a streamed q row is
//     loaded once, rotated against the warp's 4 rows and stored once: 2 KB of shared-memory traffic per pair;
//   * every warp keeps TWO q rows in flight as a wavefront, (p_t, q_a) next to (p_{t-1}, q_b): the two dot
//     products are reduced together ("transposed": the low half-warp ends up with pair A, the high half with
//     pair B, 10 instead of 20 shuffles) and ONE lane-packed rotation set-up serves both pairs;
//   * q rows travel from warp to warp through per-row version counters (acquire/release in shared memory), the
//     16-row q blocks through two TMA-filled stages (cp.async.bulk + mbarrier) exactly as in the 16-warp kernel.
// Sweep = for every P block: internal pairs (two 16-row tournaments in shared memory, then half against half
// with 2 rows per warp in registers), then all later rows stream through in blocks of 16.
// Same rotation formulas, thresholds and stopping rule as tc_jacobi_blocked.cuh; only the pair ORDER differs
// (any order that meets every pair once per sweep is a cyclic Jacobi ordering).
//
// Measured (B200, L = 32, chi = 128, 32 chains, one parity layer): 44.9 ms against 42.6 ms for the 16-warp kernel, i.e.
// the 4x lower shared-memory traffic does not pay: with 2 warps per scheduler the serial chain dot -> shuffle
// reduction -> set-up -> rotate of a warp is not hidden (scripts/ubench/pair_chain.cu: 120 cycles per row pair
// per SM with every row in registers, FP64 bound 92).  A software-pipelined visit (A's set-up next to B's rotations
// fused with the next dot, branch-free) was built and measured slower still (51.7 ms): ptxas keeps the two chains
// back to back inside the basic block.  Lessons kept here: control values that come from memory go through a REDUX
// (`uni`) and spin loops exit on a vote, otherwise every later shuffle carries a BRA.DIV divergence check; rsqrt(double)
// hides a branch, `rsqrt_nb` does not.
#pragma once
#include "../../time_crystal_tensor_network_b200/csrc/tc_common.cuh"
#include "../../time_crystal_tensor_network_b200/csrc/tc_jacobi.cuh"
#include "../../time_crystal_tensor_network_b200/csrc/tc_jacobi_blocked.cuh"

namespace tcr {
using tcb::bulk_load;
using tcb::bulk_store;
using tcb::bulk_wait_all;
using tcb::fence_async_smem;
using tcb::mbar_expect_tx;
using tcb::mbar_init;
using tcb::smem_u32;

constexpr int NW = 8, NT = NW * 32;
#ifndef TCR_PR
#define TCR_PR 4
#endif
constexpr int PR = TCR_PR;    // stationary rows per warp in the streaming phase (2 from H1 + PR - 2 from H2)
constexpr int QB = 2 * NW;    // rows of a q stage (16): two per warp and slot
constexpr int R2 = PR - 2;    // H2 rows per warp
constexpr int PB = QB + NW * R2;  // rows of a P block: halves H1 (16 rows) and H2 (8 R2 rows)
constexpr int MAX_N = 256;
constexpr unsigned FULLM = 0xffffffffu;

#ifdef TCB_TIMING
// diagnostic build: per-phase clock64 sums of warp 3 of the 256 x 256 matrices into tcb::g_tcb_timing
// [0] internal phase, [1] streaming visits, [2] hand-over waits inside visits, [3] stage waits + end-of-visit barrier,
// [4] P block load wait, [5] tournaments (part of [0]), [6] visits, [7] whole kernel
#define TCR_ARG , long long (&tacc)[8]
#define TCR_PASS , tacc
#define TCR_T(v) const long long v = clock64()
#define TCR_ACC(k, a, b) tacc[k] += (b) - (a)
#else
#define TCR_ARG
#define TCR_PASS
#define TCR_T(v)
#define TCR_ACC(k, a, b)
#endif

__host__ __device__ inline size_t smem_bytes(int n2) {
  // two q stages (= one P block) + row norms + 3 mbarriers + version counters
  return (size_t)2 * QB * n2 * sizeof(cplx) + (size_t)n2 * sizeof(double) + 64 + 2 * QB * sizeof(int);
}

template <int NPL, bool FULL>
__device__ __forceinline__ void ld_row(cplx (&v)[NPL], const cplx *row, int N, int lane) {
#pragma unroll
  for (int e = 0; e < NPL; ++e) {
    const int c = lane + 32 * e;
    v[e] = (FULL || c < N) ? row[c] : cmake(0.0, 0.0);
  }
}
template <int NPL, bool FULL>
__device__ __forceinline__ void st_row(const cplx (&v)[NPL], cplx *row, int N, int lane) {
#pragma unroll
  for (int e = 0; e < NPL; ++e) {
    const int c = lane + 32 * e;
    if (FULL || c < N) row[c] = v[e];
  }
}

// Control values that come from memory (matrix sizes from the chi table, rotation counters from shared memory) are the
// same in every lane, but the compiler cannot know: a branch on them counts as divergent and every later shuffle gets
// a BRA.DIV divergence check, which ends the basic block.  A warp reduction (REDUX) returns a provably uniform value.
__device__ __forceinline__ int uni(int x) { return __reduce_max_sync(FULLM, x); }

// Spin loops exit on a warp vote: the branch is warp-uniform for the compiler, so the code after it is known to be
// converged and the shuffles there need no BRA.DIV divergence check (which would end the basic block and with it
// the overlap of a rotation set-up with the other chain's rotations).
__device__ __forceinline__ void ver_wait(uint32_t addr, int want) {
  int v;
  unsigned spins = 0;
  do {
    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    if (++spins > (1u << 24)) __trap();  // a lost hand-over must fail loudly, never hang the GPU
  } while (__any_sync(FULLM, v < want));
}
__device__ __forceinline__ void mbar_wait_u(uint64_t *bar, uint32_t parity) {
  uint32_t ok = 0;
  const uint32_t addr = smem_u32(bar);
  unsigned spins = 0;
  do {
    if (++spins > (1u << 26)) __trap();  // a lost bulk copy must fail loudly, never hang the GPU
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (__any_sync(FULLM, ok == 0));
}
using tcb::rsqrt_nb;

template <int NPL>
__device__ __forceinline__ void dot_rows(const cplx (&u)[NPL], const cplx (&v)[NPL], double &gr, double &gi) {
  // g = sum u conj(v), four independent accumulation chains
  double g0 = 0.0, g1 = 0.0, h0 = 0.0, h1 = 0.0;
#pragma unroll
  for (int e = 0; e < NPL; ++e) {
    g0 = fma(u[e].x, v[e].x, g0);
    g1 = fma(u[e].y, v[e].y, g1);
    h0 = fma(u[e].y, v[e].x, h0);
    h1 = fma(-u[e].x, v[e].y, h1);
  }
  gr = g0 + g1;
  gi = h0 + h1;
}

template <int NPL>
__device__ __forceinline__ void rot_rows(cplx (&u)[NPL], cplx (&v)[NPL], double cs, double sr, double si) {
  // u' = c u - (s e) v ;  v' = conj(s e) u + c v
#pragma unroll
  for (int e = 0; e < NPL; ++e) {
    cplx un, vn;
    un.x = fma(cs, u[e].x, fma(-sr, v[e].x, si * v[e].y));
    un.y = fma(cs, u[e].y, -fma(sr, v[e].y, si * v[e].x));
    vn.x = fma(cs, v[e].x, fma(sr, u[e].x, si * u[e].y));
    vn.y = fma(cs, v[e].y, fma(sr, u[e].y, -si * u[e].x));
    u[e] = un;
    v[e] = vn;
  }
}

// Two independent row pairs at once: A = (uA, vA) with squared norms (aA, bA), B likewise.  Returns flag bits:
// 1 / 2 = pair A / B was rotated, 4 / 8 = that rotation was not yet small (convergence bookkeeping); the norms are
// updated in place (all lanes hold them).
template <int NPL>
__device__ __forceinline__ int pair2(cplx (&uA)[NPL], cplx (&vA)[NPL], double &aA, double &bA, bool actA,
                                     cplx (&uB)[NPL], cplx (&vB)[NPL], double &aB, double &bB, bool actB,
                                     double dead, double tol2, int lane) {
  actA = __all_sync(FULLM, actA && aA > dead && bA > dead);  // votes: warp-uniform for the compiler as well
  actB = __all_sync(FULLM, actB && aB > dead && bB > dead);
  if (!actA && !actB) return 0;
  double grA = 0.0, giA = 0.0, grB = 0.0, giB = 0.0;
  if (actA) dot_rows<NPL>(uA, vA, grA, giA);
  if (actB) dot_rows<NPL>(uB, vB, grB, giB);
  // transposed reduction: lanes 0..15 end up with the sums of pair A, lanes 16..31 with those of pair B
  const bool hi = lane >= 16;
  double kx = hi ? grB : grA, ky = hi ? giB : giA;
  {
    const double sx = hi ? grA : grB, sy = hi ? giA : giB;
    kx += __shfl_xor_sync(FULLM, sx, 16);
    ky += __shfl_xor_sync(FULLM, sy, 16);
  }
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) {
    kx += __shfl_xor_sync(FULLM, kx, o);
    ky += __shfl_xor_sync(FULLM, ky, o);
  }
  // lane-packed rotation set-up (formulas of tcb::make_rot, FP64 branch): dd = aj - ai, 2r = sqrt(dd^2 + 4|g|^2),
  // c^2 = 1/2 + |dd|/(4r), s e = sign(dd) g / (2 r c), moved squared norm t|g| = sign(dd) |g|^2 / (2 r c^2)
  const double ai = hi ? aB : aA, aj = hi ? bB : bA;
  const bool act = hi ? actB : actA;
  const double g2 = fma(kx, kx, ky * ky);
  const double thr = ai * aj;
  const bool rot = act && (g2 > tol2 * thr);
  const bool big = rot && (g2 > tcj::SMALL_REL2 * thr);
  const unsigned brot = __ballot_sync(FULLM, rot);
  const bool rotA = (brot & 1u) != 0, rotB = (brot & 0x10000u) != 0;
  if (!rotA && !rotB) return 0;
  const unsigned bbig = __ballot_sync(FULLM, big);
  double cs, sr, si, tg;
  {
    const double dd = aj - ai;
    const double q = fma(dd, dd, 4.0 * g2);
    const double rinv = rsqrt_nb(rot ? q : 1.0);  // 1 / (2r)
    const double c2 = fma(0.5 * fabs(dd), rinv, 0.5);
    const double cinv = rsqrt_nb(c2);
    cs = c2 * cinv;
    const double ks = copysign(rinv * cinv, dd);
    sr = ks * kx;
    si = ks * ky;
    tg = g2 * ks * cinv;
  }
  if (rotA) {
    const double c = __shfl_sync(FULLM, cs, 0), s0 = __shfl_sync(FULLM, sr, 0), s1 = __shfl_sync(FULLM, si, 0);
    const double t = __shfl_sync(FULLM, tg, 0);
    rot_rows<NPL>(uA, vA, c, s0, s1);
    aA -= t;
    bA += t;
  }
  if (rotB) {
    const double c = __shfl_sync(FULLM, cs, 16), s0 = __shfl_sync(FULLM, sr, 16), s1 = __shfl_sync(FULLM, si, 16);
    const double t = __shfl_sync(FULLM, tg, 16);
    rot_rows<NPL>(uB, vB, c, s0, s1);
    aB -= t;
    bB += t;
  }
  return (int)rotA | ((int)rotB << 1) | (int)((bbig & 1u) << 2) | (int)(((bbig >> 16) & 1u) << 3);
}
__device__ __forceinline__ int nbig(int flags) { return ((flags >> 2) & 1) + ((flags >> 3) & 1); }

__device__ __forceinline__ void ver_release(uint32_t addr, int val, int lane) {
  __syncwarp();
  if (lane == 0) asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(addr), "r"(val) : "memory");
}

// One q stage (rowsQ <= 16 rows in shared memory) against the NP stationary rows of every warp.
// Slot s of warp w handles the stage rows ja = 2 ((w + s) mod 8) and ja + 1; row j is the s-th user's once its
// version counter reads base + s.  Wavefront inside the warp: step t pairs (p_t, q_a) with (p_{t-1}, q_b); the last
// pair of q_b overlaps the first pair of the next slot's q_a.
template <int NPL, bool FULL, int NP, class Mid>
__device__ __forceinline__ int visit(cplx (&u)[PR][NPL], double (&pn)[PR], unsigned pvalid, cplx *Q, double *qn,
                                     int rowsQ, int N, int *ver, int base, int warp, int lane, double dead,
                                     double tol2, Mid &&mid TCR_ARG) {
  int nrot = 0;
  cplx vA[NPL], vB[NPL];
  double nA = 0.0, nB = 0.0;
  bool okA = false, okB = false;
  int jb_prev = 0;
  const uint32_t vaddr = smem_u32(ver);
#pragma unroll 1
  for (int s = 0; s < NW; ++s) {
    const int ja = 2 * ((warp + s) & (NW - 1)), jb = ja + 1;
    if (s == NW / 2) mid();
    // ---- q_a of this slot
    TCR_T(tw0);
    if (s > 0) ver_wait(vaddr + 4 * ja, base + s);
    TCR_T(tw1);
    TCR_ACC(2, tw0, tw1);
    okA = ja < rowsQ;
    if (okA) {
      ld_row<NPL, FULL>(vA, Q + (size_t)ja * N, N, lane);
      nA = qn[ja];
    }
    // step 0: (p_0, q_a) with the last pair of the previous slot's q_b
    nrot += nbig(pair2<NPL>(u[0], vA, pn[0], nA, okA && (pvalid & 1u), u[NP - 1], vB, pn[NP - 1], nB,
                       s > 0 && okB && ((pvalid >> (NP - 1)) & 1u), dead, tol2, lane));
    if (s > 0) {
      if (okB) {
        st_row<NPL, FULL>(vB, Q + (size_t)jb_prev * N, N, lane);
        if (lane == 0) qn[jb_prev] = nB;
      }
      ver_release(vaddr + 4 * jb_prev, base + s, lane);  // use number s-1 of that row is over
    }
    // ---- q_b of this slot
    TCR_T(tw2);
    if (s > 0) ver_wait(vaddr + 4 * jb, base + s);
    TCR_T(tw3);
    TCR_ACC(2, tw2, tw3);
    okB = jb < rowsQ;
    if (okB) {
      ld_row<NPL, FULL>(vB, Q + (size_t)jb * N, N, lane);
      nB = qn[jb];
    }
#pragma unroll
    for (int t = 1; t < NP; ++t)
      nrot += nbig(pair2<NPL>(u[t], vA, pn[t], nA, okA && ((pvalid >> t) & 1u), u[t - 1], vB, pn[t - 1], nB,
                         okB && ((pvalid >> (t - 1)) & 1u), dead, tol2, lane));
    if (okA) {
      st_row<NPL, FULL>(vA, Q + (size_t)ja * N, N, lane);
      if (lane == 0) qn[ja] = nA;
    }
    ver_release(vaddr + 4 * ja, base + s + 1, lane);
    jb_prev = jb;
  }
  // drain: the last pair of the last q_b
  {
    double dumA = 0.0, dumB = 0.0;
    nrot += nbig(pair2<NPL>(u[0], vA, dumA, dumB, false, u[NP - 1], vB, pn[NP - 1], nB,
                       okB && ((pvalid >> (NP - 1)) & 1u), dead, tol2, lane));
    if (okB) {
      st_row<NPL, FULL>(vB, Q + (size_t)jb_prev * N, N, lane);
      if (lane == 0) qn[jb_prev] = nB;
    }
    ver_release(vaddr + 4 * jb_prev, base + NW, lane);
  }
  return nrot;
}

// Internal pairs of a P block that sits in shared memory (H1 = rows 0..rows1-1 in the stage-0 area, H2 = rows2 rows
// in the stage-1 area): (1) circle-method tournaments inside H1 and inside H2, warp w plays pair w of each side by
// side; (2) H1 against H2 with the H1 rows 2w, 2w+1 of warp w in registers and H2 as the q stage.  All rows are back
// in shared memory when it returns.
template <int NPL, bool FULL>
__device__ __forceinline__ int internal_phase(int n2, int row0, int rows1, int rows2, int N, int verBase1, double dead,
                                           double tol2 TCR_ARG) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  cplx *const H1 = reinterpret_cast<cplx *>(smem_raw);
  cplx *const H2 = H1 + (size_t)QB * N;
  unsigned char *const tail = smem_raw + (size_t)2 * QB * n2 * sizeof(cplx);
  double *const nP0 = reinterpret_cast<double *>(tail) + row0;
  int *const s_ver = reinterpret_cast<int *>(tail + n2 * sizeof(double) + 4 * sizeof(uint64_t));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int nrot = 0;
  const int rmax = max(rows1, rows2) - 1;
  TCR_T(ti0);
  for (int r = 0; r < rmax; ++r) {
    const bool actA = r < rows1 - 1 && warp < rows1 / 2;
    const bool actB = r < rows2 - 1 && warp < rows2 / 2;
    if (actA || actB) {
      int iA = 0, jA = 1, iB = 0, jB = 1;  // an inactive side loads rows 0, 1 of H1 (always present) and drops them
      if (actA) tcj::rr_pair(rows1, r, warp, iA, jA);
      if (actB) tcj::rr_pair(rows2, r, warp, iB, jB);
      cplx *const hB = actB ? H2 : H1;
      double *const nB = actB ? nP0 + QB : nP0;
      cplx uA[NPL], vA[NPL], uB[NPL], vB[NPL];
      ld_row<NPL, FULL>(uA, H1 + (size_t)iA * N, N, lane);
      ld_row<NPL, FULL>(vA, H1 + (size_t)jA * N, N, lane);
      ld_row<NPL, FULL>(uB, hB + (size_t)iB * N, N, lane);
      ld_row<NPL, FULL>(vB, hB + (size_t)jB * N, N, lane);
      double aA = nP0[iA], bA = nP0[jA], aB = nB[iB], bB = nB[jB];
      const int fl = pair2<NPL>(uA, vA, aA, bA, actA, uB, vB, aB, bB, actB, dead, tol2, lane);
      nrot += nbig(fl);
      if (fl & 1) {
        st_row<NPL, FULL>(uA, H1 + (size_t)iA * N, N, lane);
        st_row<NPL, FULL>(vA, H1 + (size_t)jA * N, N, lane);
        if (lane == 0) {
          nP0[iA] = aA;
          nP0[jA] = bA;
        }
      }
      if (fl & 2) {
        st_row<NPL, FULL>(uB, H2 + (size_t)iB * N, N, lane);
        st_row<NPL, FULL>(vB, H2 + (size_t)jB * N, N, lane);
        if (lane == 0) {
          nP0[QB + iB] = aB;
          nP0[QB + jB] = bB;
        }
      }
    }
    __syncthreads();
  }
  TCR_T(ti1);
  TCR_ACC(5, ti0, ti1);
  if (rows2 > 0) {
    cplx u[PR][NPL];
    double pn[PR];
    unsigned pvalid = 0;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int r = 2 * warp + k;
      const bool ok = r < rows1;
      ld_row<NPL, FULL>(u[k], H1 + (size_t)(ok ? r : 0) * N, N, lane);
      pn[k] = ok ? nP0[r] : 0.0;
      pvalid |= (ok ? 1u : 0u) << k;
    }
    nrot += visit<NPL, FULL, 2>(u, pn, pvalid, H2, nP0 + QB, rows2, N, s_ver + QB, verBase1, warp, lane, dead, tol2,
                                [] {} TCR_PASS);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int r = 2 * warp + k;
      if (r < rows1) {
        st_row<NPL, FULL>(u[k], H1 + (size_t)r * N, N, lane);
        if (lane == 0) nP0[r] = pn[k];
      }
    }
  }
  return nrot;
}

template <int NPL, bool FULL>
__device__ void sweeps(const TcDev &d, const Bond &b, cplx *X, int K, int N, int *s_rot, double *red) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  cplx *const sQ = reinterpret_cast<cplx *>(smem_raw);  // stage k = sQ + k * QB * N; a P block fills both
  unsigned char *const tail = smem_raw + (size_t)2 * QB * d.n2 * sizeof(cplx);
  double *const s_nrm2 = reinterpret_cast<double *>(tail);
  uint64_t *const barP = reinterpret_cast<uint64_t *>(tail + d.n2 * sizeof(double));
  uint64_t *const barQ = barP + 1;
  int *const s_ver = reinterpret_cast<int *>(barP + 4);  // [2][QB]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nP = (K + PB - 1) / PB;
  const double tol = 2.0 * sqrt((double)N) * 2.220446049250313e-16;
  const double tol2 = tol * tol;
  const uint32_t row_bytes = (uint32_t)N * sizeof(cplx);
  uint32_t phP = 0, phQ0 = 0, phQ1 = 0;
  int verBase0 = 0, verBase1 = 0;
  int sweep = 0;
#ifdef TCB_TIMING
  long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const long long tk0 = clock64();
#endif
  double dead = 0.0;
  for (; sweep < tcj::MAX_SWEEPS; ++sweep) {
    // all stores of the previous sweep have landed before rows are re-read
    if (tid == 0) bulk_wait_all();
    __syncthreads();
    for (int r = warp; r < K; r += NW) {
      const cplx *row = X + (size_t)r * N;
      double s = 0.0;
      for (int c = lane; c < N; c += 32) s += cabs2(__ldcg(reinterpret_cast<const double2 *>(row + c)));
      s = tcj::warp_sum(s);
      if (lane == 0) s_nrm2[r] = s;
    }
    if (tid == 0) *s_rot = 0;
    __syncthreads();
    if (sweep == 0) {
      double p = 0.0;
      for (int r = tid; r < K; r += NT) p += s_nrm2[r];
      dead = tcj::DEAD_REL2 * block_sum(p, red);
    }
    int nrot = 0;  // rotations that were not yet small
    for (int a = 0; a < nP; ++a) {
      const int row0 = a * PB;
      const int rowsP = min(PB, K - row0);
      const int rows1 = min(QB, rowsP), rows2 = rowsP - rows1;  // halves H1 (stage 0) and H2 (stage 1)
      cplx *gP = X + (size_t)row0 * N;
      double *nP0 = s_nrm2 + row0;
      if (tid == 0) {
        bulk_wait_all();  // earlier stores out of the stages have finished reading shared memory
        mbar_expect_tx(barP, rowsP * row_bytes);
        bulk_load(sQ, gP, rowsP * row_bytes, barP);
      }
      TCR_T(tp0);
      mbar_wait_u(barP, phP);
      phP ^= 1;
      TCR_T(tp1);
      TCR_ACC(4, tp0, tp1);
      // ---- internal pairs of the P block
      nrot += internal_phase<NPL, FULL>(d.n2, row0, rows1, rows2, N, verBase1, dead, tol2 TCR_PASS);
      if (rows2 > 0) verBase1 += NW;
      __syncthreads();
      TCR_T(tp2);
      TCR_ACC(0, tp1, tp2);
      // ---- the stationary rows of this warp: H1 rows 2w, 2w+1 and H2 rows R2 w ..
      cplx *H1 = sQ, *H2 = sQ + (size_t)QB * N;
      cplx u[PR][NPL];
      double pn[PR];
      unsigned pvalid = 0;
#pragma unroll
      for (int k = 0; k < PR; ++k) {
        const int r = (k < 2) ? (2 * warp + k) : (R2 * warp + (k - 2));
        const bool ok = (k < 2) ? (r < rows1) : (r < rows2);
        const cplx *src = ((k < 2) ? H1 : H2) + (size_t)(ok ? r : 0) * N;
        ld_row<NPL, FULL>(u[k], src, N, lane);
        if (!ok) {
#pragma unroll
          for (int e = 0; e < NPL; ++e) u[k][e] = cmake(0.0, 0.0);
        }
        pn[k] = ok ? nP0[(k < 2 ? 0 : QB) + r] : 0.0;
        pvalid |= (ok ? 1u : 0u) << k;
      }
      __syncthreads();  // both stage areas are free from here
      // ---- every later row streams through in blocks of QB rows
      const int qrow0 = row0 + PB;
      const int nQ = qrow0 < K ? (K - qrow0 + QB - 1) / QB : 0;
      if (tid == 0) {
        for (int k = 0; k < 2 && k < nQ; ++k) {
          const int rq = min(QB, K - (qrow0 + k * QB));
          mbar_expect_tx(&barQ[k], rq * row_bytes);
          bulk_load(sQ + (size_t)k * QB * N, X + (size_t)(qrow0 + k * QB) * N, rq * row_bytes, &barQ[k]);
        }
      }
      for (int qb = 0; qb < nQ; ++qb) {
        const int buf = qb & 1;
        const int rq0 = qrow0 + qb * QB;
        const int rowsQ = min(QB, K - rq0);
        TCR_T(tv0);
        mbar_wait_u(&barQ[buf], buf ? phQ1 : phQ0);
        if (buf)
          phQ1 ^= 1;
        else
          phQ0 ^= 1;
        TCR_T(tv1);
        cplx *Q = sQ + (size_t)buf * QB * N;
        // mid-visit: refill the other stage with block qb + 1; its store (end of visit qb - 1) is long over by then
        auto refill = [&]() {
          if (tid == 0 && qb >= 1 && qb + 1 < nQ) {
            bulk_wait_all();
            const int rq = min(QB, K - (rq0 + QB));
            mbar_expect_tx(&barQ[buf ^ 1], rq * row_bytes);
            bulk_load(sQ + (size_t)(buf ^ 1) * QB * N, X + (size_t)(rq0 + QB) * N, rq * row_bytes, &barQ[buf ^ 1]);
          }
        };
        nrot += visit<NPL, FULL, PR>(u, pn, pvalid, Q, s_nrm2 + rq0, rowsQ, N, s_ver + buf * QB,
                                     buf ? verBase1 : verBase0, warp, lane, dead, tol2, refill TCR_PASS);
        if (buf)
          verBase1 += NW;
        else
          verBase0 += NW;
        TCR_T(tv2);
        fence_async_smem();
        __syncthreads();
        TCR_T(tv3);
        TCR_ACC(1, tv1, tv2);
        TCR_ACC(3, tv0, tv1);
        TCR_ACC(3, tv2, tv3);
#ifdef TCB_TIMING
        tacc[6] += 1;
#endif
        if (tid == 0) bulk_store(X + (size_t)rq0 * N, Q, rowsQ * row_bytes);
      }
      // ---- the P block back to global, straight from the registers
#pragma unroll
      for (int k = 0; k < PR; ++k) {
        if ((pvalid >> k) & 1u) {
          const int r = (k < 2) ? (2 * warp + k) : (QB + R2 * warp + (k - 2));
          st_row<NPL, FULL>(u[k], gP + (size_t)r * N, N, lane);
          if (lane == 0) nP0[r] = pn[k];
        }
      }
      asm volatile("fence.proxy.async;" ::: "memory");  // generic-proxy stores before later bulk (async-proxy) loads
      __syncthreads();
    }
    if (lane == 0 && nrot) atomicAdd(s_rot, nrot);
    __syncthreads();
    const int tot = uni(*s_rot);
    __syncthreads();
    if (tot == 0) break;
  }
#ifdef TCB_TIMING
  if (K == 256 && N == 256 && warp == 3 && lane == 0) {
    tacc[7] = clock64() - tk0;
    for (int k = 0; k < 8; ++k) atomicAdd(&tcb::g_tcb_timing[k], (unsigned long long)tacc[k]);
  }
#endif
  if (tid == 0) {
    bulk_wait_all();
    if (sweep >= tcj::MAX_SWEEPS) atomicAdd(&d.flags[1], 1);
    atomicMax(&d.flags[2], sweep + 1);
    if (K >= 128) {  // sweep statistics of the large matrices (diagnostics)
      atomicAdd(&d.flags[3], sweep + 1);
      atomicAdd(&d.flags[4], 1);
    }
  }
  __syncthreads();
  double *w = d.ww + b.slot * d.n2;
  for (int r = warp; r < K; r += NW) {
    const cplx *row = X + (size_t)r * N;
    double s = 0.0;
    for (int c = lane; c < N; c += 32) s += cabs2(__ldcg(reinterpret_cast<const double2 *>(row + c)));
    s = tcj::warp_sum(s);
    if (lane == 0) w[r] = sqrt(s);
  }
}

__global__ void __launch_bounds__(NT, 1) jacobi_rb_kernel(TcDev d, LayerArgs a) {
  Bond b;
  // blockIdx.x = chain, blockIdx.y = rank of the bond in centre-out order
  if (!get_bond(d, a, centre_out(blockIdx.y, a.nb), blockIdx.x, b)) return;
  const int N = uni(b.N), K = uni(b.M < b.N ? b.M : b.N);
  cplx *X = d.Xw + b.slot * d.slot_stride;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + (size_t)2 * QB * d.n2 * sizeof(cplx) + d.n2 * sizeof(double));
  __shared__ double red[32];
  __shared__ int s_rot;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    fence_async_smem();
  }
  if (threadIdx.x < 2 * QB) reinterpret_cast<int *>(bars + 4)[threadIdx.x] = 0;
  __syncthreads();
  const int npl = (N + 31) / 32;
  if (N == 256)
    sweeps<8, true>(d, b, X, K, N, &s_rot, red);
  else if (npl <= 1)
    sweeps<1, false>(d, b, X, K, N, &s_rot, red);
  else if (npl <= 2)
    sweeps<2, false>(d, b, X, K, N, &s_rot, red);
  else if (npl <= 4)
    sweeps<4, false>(d, b, X, K, N, &s_rot, red);
  else
    sweeps<8, false>(d, b, X, K, N, &s_rot, red);
}
}  // namespace tcr
