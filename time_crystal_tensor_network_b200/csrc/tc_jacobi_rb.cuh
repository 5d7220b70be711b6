// tc_jacobi_rb.cuh -- K2b fast path, register-blocked: one-sided Jacobi on the rows of the triangular
// factor with FOUR stationary rows per warp in registers and every streamed row reused four times per
// shared-memory load.
//
// Why (ncu of the 16-warp kernel tc_jacobi_blocked.cuh, profiles/r01_*): one stationary row per warp costs
// 4 KB LDS + 4 KB STS per row pair, which puts the shared-memory pipe at 52 % while the FP64 pipe sits at 45 %:
// the two co-limit, and every pair pays its own shuffle reduction and scalar rotation set-up chain.  Here
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
#pragma once
#include "tc_common.cuh"
#include "tc_jacobi.cuh"
#include "tc_jacobi_blocked.cuh"

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
#define TCR_ARGP , long long *taccp
#define TCR_PASSP , tacc
#define TCR_TACC                                  \
  long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define TCR_TSAVE \
  for (int k_ = 0; k_ < 8; ++k_) taccp[k_] += tacc[k_]
#define TCR_T(v) const long long v = clock64()
#define TCR_ACC(k, a, b) tacc[k] += (b) - (a)
#else
#define TCR_ARG
#define TCR_PASS
#define TCR_ARGP
#define TCR_PASSP
#define TCR_TACC
#define TCR_TSAVE
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
// 1/sqrt(x) for normal positive x without the special-case branch of rsqrt(double): hardware seed (MUFU.RSQ64H,
// ~2^-22) and two Newton steps; the branch would split the basic block just like BRA.DIV.
__device__ __forceinline__ double rsqrt_nb(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double h = 0.5 * x;
  double e = fma(-h * y, y, 0.5);
  y = fma(y, e, y);
  e = fma(-h * y, y, 0.5);
  y = fma(y, e, y);
  return y;
}
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
// visits return two counters in one int: bits 0..15 rotations, bits 16..31 rotations that were not yet small
__device__ __forceinline__ int packfl(int flags) { return (flags & 1) + ((flags >> 1) & 1) + (nbig(flags) << 16); }

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
    nrot += packfl(pair2<NPL>(u[0], vA, pn[0], nA, okA && (pvalid & 1u), u[NP - 1], vB, pn[NP - 1], nB,
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
      nrot += packfl(pair2<NPL>(u[t], vA, pn[t], nA, okA && ((pvalid >> t) & 1u), u[t - 1], vB, pn[t - 1], nB,
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
    nrot += packfl(pair2<NPL>(u[0], vA, dumA, dumB, false, u[NP - 1], vB, pn[NP - 1], nB,
                       okB && ((pvalid >> (NP - 1)) & 1u), dead, tol2, lane));
    if (okB) {
      st_row<NPL, FULL>(vB, Q + (size_t)jb_prev * N, N, lane);
      if (lane == 0) qn[jb_prev] = nB;
    }
    ver_release(vaddr + 4 * jb_prev, base + NW, lane);
  }
  return nrot;
}

// ------------------------------------------------------------------------------------------------
// Software-pipelined visit (NP >= 3).  The lock-step visit above leaves the FP64 pipe idle while a warp walks
// through its serial chain dot -> shuffle reduction -> rotation set-up (~400 cycles of latency per step, 2 warps per
// scheduler cannot hide it).  Here the two q rows of a warp run half a step apart:
//     time 2t   : A.setup(t)   ||  B.dense(t-1) = rotate (p_{t-2}, q_b), fused with the NEXT dot (p_{t-1}', q_b')
//     time 2t+1 : B.setup(t)   ||  A.dense(t)   = rotate (p_t, q_a),     fused with the NEXT dot (p_{t+1}, q_a')
// so every latency chain sits in the same basic block as 128 independent DFMAs of the other chain.  Everything is
// branch-free: a pair below the threshold gets the identity rotation (c = 1, s = 0: exact), which is why the sweep
// loop falls back to the skipping lock-step visit once most pairs of a sweep no longer rotate.
// No dot is ever stale: B touches p_{t-1} only after A.dense(t-1), and for NP >= 3 the tail of the previous slot's
// q_b (rows p_{NP-2}, p_{NP-1} at times 0 and 2) is over before A's dots with those rows are formed (times >= 2NP-5).
// ------------------------------------------------------------------------------------------------
struct RotP {
  double cs, sr, si;
};

// butterfly-reduce the lane partials of g = p . conj(q), then the rotation of tcb::make_rot (FP64 branch) for the rows
// with squared norms (ai, aj); identity when the pair is inactive or below the threshold.  Norms updated in place.
// Returns bit0 = rotated, bit1 = rotation not yet small.
__device__ __forceinline__ int setup1(double gr, double gi, double &ai, double &aj, bool act, double dead, double tol2,
                                      RotP &r) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    gr += __shfl_xor_sync(FULLM, gr, o);
    gi += __shfl_xor_sync(FULLM, gi, o);
  }
  const double g2 = fma(gr, gr, gi * gi);
  const double thr = ai * aj;
  const bool rot = act && ai > dead && aj > dead && (g2 > tol2 * thr);
  const bool big = rot && (g2 > tcj::SMALL_REL2 * thr);
  const double dd = aj - ai;
  const double rinv = rsqrt_nb(rot ? fma(dd, dd, 4.0 * g2) : 1.0);  // 1 / (2r)
  const double c2 = fma(0.5 * fabs(dd), rinv, 0.5);
  const double cinv = rsqrt_nb(c2);
  const double ks = copysign(rinv * cinv, dd);
  r.cs = rot ? c2 * cinv : 1.0;
  r.sr = rot ? ks * gr : 0.0;
  r.si = rot ? ks * gi : 0.0;
  const double tg = rot ? g2 * ks * cinv : 0.0;
  ai -= tg;
  aj += tg;
  return (int)rot | ((int)big << 1);
}

// rotate (u, v), then the lane partials of the next dot  w . conj(v')
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

// One slot of the pipelined visit.  HAS_A: this slot brings a new pair of q rows (q_a = row ja, q_b = row ja + 1);
// HAS_B: the previous slot's q_b (row jbp) is still finishing.  First slot <true, false>, drain <false, true>.
template <int NPL, bool FULL, int NP, bool HAS_A, bool HAS_B>
__device__ __forceinline__ int slot_sp(cplx (&u)[PR][NPL], double (&pn)[PR], unsigned pvalid, cplx (&vA)[NPL],
                                       cplx (&vB)[NPL], double &nB, bool &okB, double &gBr, double &gBi, RotP &rB,
                                       cplx *Q, double *qn, int rowsQ, int N, uint32_t vaddr, int base, int s, int ja,
                                       int jbp, int lane, double dead, double tol2 TCR_ARG) {
  int fl = 0;  // bits 0..15: rotations, 16..31: rotations not yet small
  RotP rA;
  double nA = 0.0, gAr = 0.0, gAi = 0.0;
  bool okA = false;
  if (HAS_A) {
    TCR_T(tw0);
    if (s > 0) ver_wait(vaddr + 4 * ja, base + s);
    TCR_T(tw1);
    TCR_ACC(2, tw0, tw1);
    okA = ja < rowsQ;
    ld_row<NPL, FULL>(vA, Q + (size_t)(okA ? ja : 0) * N, N, lane);
    nA = okA ? qn[ja] : 0.0;
    dot_rows<NPL>(u[0], vA, gAr, gAi);
  }
  auto acc = [&](int f) { fl += (f & 1) + ((f & 2) << 15); };
  // ---- time 0: A.setup(0) || B.dense(NP-1) of the previous slot
  if (HAS_A) acc(setup1(gAr, gAi, pn[0], nA, okA && (pvalid & 1u), dead, tol2, rA));
  if (HAS_B) dense<NPL>(u[NP - 2], vB, rB, u[NP - 1], gBr, gBi);
  // ---- time 1: B.setup(NP) || A.dense(0)
  if (HAS_B) acc(setup1(gBr, gBi, pn[NP - 1], nB, okB && ((pvalid >> (NP - 1)) & 1u), dead, tol2, rB));
  if (HAS_A) dense<NPL>(u[0], vA, rA, u[1], gAr, gAi);
  // ---- time 2: A.setup(1) || last rotation of the previous q_b, its store, the new q_b and its first dot
  if (HAS_A) acc(setup1(gAr, gAi, pn[1], nA, okA && ((pvalid >> 1) & 1u), dead, tol2, rA));
  if (HAS_B) {
    rot_rows<NPL>(u[NP - 1], vB, rB.cs, rB.sr, rB.si);
    if (okB) {
      st_row<NPL, FULL>(vB, Q + (size_t)jbp * N, N, lane);
      if (lane == 0) qn[jbp] = nB;
    }
    ver_release(vaddr + 4 * jbp, base + s, lane);  // use number s-1 of that row is over
  }
  if (HAS_A) {
    const int jb = ja + 1;
    TCR_T(tw2);
    if (s > 0) ver_wait(vaddr + 4 * jb, base + s);
    TCR_T(tw3);
    TCR_ACC(2, tw2, tw3);
    okB = jb < rowsQ;
    ld_row<NPL, FULL>(vB, Q + (size_t)(okB ? jb : 0) * N, N, lane);
    nB = okB ? qn[jb] : 0.0;
    dot_rows<NPL>(u[0], vB, gBr, gBi);
    // ---- times 2t+1, 2t+2
#pragma unroll
    for (int t = 1; t < NP; ++t) {
      acc(setup1(gBr, gBi, pn[t - 1], nB, okB && ((pvalid >> (t - 1)) & 1u), dead, tol2, rB));  // B.setup(t)
      if (t < NP - 1) {
        dense<NPL>(u[t], vA, rA, u[t + 1], gAr, gAi);                                               // A.dense(t)
        acc(setup1(gAr, gAi, pn[t + 1], nA, okA && ((pvalid >> (t + 1)) & 1u), dead, tol2, rA));  // A.setup(t+1)
        dense<NPL>(u[t - 1], vB, rB, u[t], gBr, gBi);                                               // B.dense(t)
      } else {
        rot_rows<NPL>(u[t], vA, rA.cs, rA.sr, rA.si);  // last rotation of q_a; B.dense(NP-1) opens the next slot
      }
    }
    if (okA) {
      st_row<NPL, FULL>(vA, Q + (size_t)ja * N, N, lane);
      if (lane == 0) qn[ja] = nA;
    }
    ver_release(vaddr + 4 * ja, base + s + 1, lane);
  }
  return fl;
}

template <int NPL, bool FULL, int NP, class Mid>
__device__ __forceinline__ int visit_sp(cplx (&u)[PR][NPL], double (&pn)[PR], unsigned pvalid, cplx *Q, double *qn,
                                        int rowsQ, int N, int *ver, int base, int warp, int lane, double dead,
                                        double tol2, Mid &&mid TCR_ARG) {
  static_assert(NP >= 3, "the pipelined visit needs three stationary rows per warp (stale-dot hazard otherwise)");
  cplx vA[NPL], vB[NPL];
  double nB = 0.0, gBr = 0.0, gBi = 0.0;
  bool okB = false;
  RotP rB;
  rB.cs = 1.0;
  rB.sr = rB.si = 0.0;
  const uint32_t vaddr = smem_u32(ver);
  int fl = slot_sp<NPL, FULL, NP, true, false>(u, pn, pvalid, vA, vB, nB, okB, gBr, gBi, rB, Q, qn, rowsQ, N, vaddr, base,
                                               0, 2 * warp, 0, lane, dead, tol2 TCR_PASS);
#pragma unroll 1
  for (int s = 1; s < NW; ++s) {
    if (s == NW / 2) mid();
    const int ja = 2 * ((warp + s) & (NW - 1)), jbp = 2 * ((warp + s - 1) & (NW - 1)) + 1;
    fl += slot_sp<NPL, FULL, NP, true, true>(u, pn, pvalid, vA, vB, nB, okB, gBr, gBi, rB, Q, qn, rowsQ, N, vaddr, base, s,
                                             ja, jbp, lane, dead, tol2 TCR_PASS);
  }
  fl += slot_sp<NPL, FULL, NP, false, true>(u, pn, pvalid, vA, vB, nB, okB, gBr, gBi, rB, Q, qn, rowsQ, N, vaddr, base, NW,
                                            0, 2 * ((warp + NW - 1) & (NW - 1)) + 1, lane, dead, tol2 TCR_PASS);
  return fl;
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
                                [] {} TCR_PASS) >> 16;
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

// Sweeps of one matrix.  SP = true: pipelined streaming visits, runs while at least half of the streaming pairs of a
// sweep still rotate (identity rotations cost the pipelined visit as much as real ones), then hands the matrix over;
// SP = false: skipping lock-step visits until convergence.  The two live in two kernels (launched back to back) so
// that each streaming loop gets its own register allocation: side by side in one kernel they spill ~750 instructions
// per slot, as __noinline__ functions ~200, alone 10-80.
// `state` (one int per matrix, the knew slot finalize_kernel overwrites later): sweeps done | CONVERGED.
constexpr int CONVERGED = 1 << 20;

template <int NPL, bool FULL, bool SP>
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
  int *const state = d.knew + b.slot;
  // streaming pairs per sweep (all pairs minus the P-block internal ones) for the pipelined / skipping switch
  int stream_pairs = K * (K - 1) / 2;
  for (int a = 0; a < nP; ++a) {
    const int rp = min(PB, K - a * PB);
    stream_pairs -= rp * (rp - 1) / 2;
  }
  // d.rot64 bit 2: always the skipping lock-step visit, bit 3: always the pipelined visit (A/B testing)
  const int force = (d.rot64 >> 2) & 3;
  int sweep = 0;
  bool converged = false;
  if (SP) {
    if (force == 1 || stream_pairs == 0) {
      if (tid == 0) *state = 0;
      return;
    }
  } else {
    const int st = uni(*state);
    if (st & CONVERGED) return;
    sweep = st;
  }
#ifdef TCB_TIMING
  long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const long long tk0 = clock64();
#endif
  double dead = 0.0;
  bool first = true;
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
    if (tid == 0) s_rot[0] = s_rot[1] = 0;
    __syncthreads();
    if (first) {
      // |theta|_F^2 is invariant under the rotations: the hand-over kernel recomputes the same threshold
      double p = 0.0;
      for (int r = tid; r < K; r += NT) p += s_nrm2[r];
      dead = tcj::DEAD_REL2 * block_sum(p, red);
      first = false;
    }
    int nrot = 0, nstream = 0;  // rotations not yet small (all phases); rotations of the streaming phase
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
        int fl;
        if (SP)
          fl = visit_sp<NPL, FULL, PR>(u, pn, pvalid, Q, s_nrm2 + rq0, rowsQ, N, s_ver + buf * QB,
                                       buf ? verBase1 : verBase0, warp, lane, dead, tol2, refill TCR_PASS);
        else
          fl = visit<NPL, FULL, PR>(u, pn, pvalid, Q, s_nrm2 + rq0, rowsQ, N, s_ver + buf * QB,
                                    buf ? verBase1 : verBase0, warp, lane, dead, tol2, refill TCR_PASS);
        nrot += fl >> 16;
        nstream += fl & 0xffff;
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
    if (lane == 0 && nrot) atomicAdd(&s_rot[0], nrot);
    if (lane == 0 && nstream) atomicAdd(&s_rot[1], nstream);
    __syncthreads();
    const int tot = uni(s_rot[0]), tot_stream = uni(s_rot[1]);
    __syncthreads();
    if (tot == 0) {
      converged = true;
      break;
    }
    if (SP && force != 2 && 2 * tot_stream <= stream_pairs) {
      ++sweep;
      break;  // hand over to the skipping kernel
    }
  }
#ifdef TCB_TIMING
  if (K == 256 && N == 256 && warp == 3 && lane == 0) {
    tacc[7] = clock64() - tk0;
    for (int k = 0; k < 8; ++k) atomicAdd(&tcb::g_tcb_timing[k], (unsigned long long)tacc[k]);
  }
#endif
  if (tid == 0) bulk_wait_all();
  if (SP && !converged && sweep < tcj::MAX_SWEEPS) {
    if (tid == 0) *state = sweep;
    return;
  }
  if (tid == 0) {
    *state = CONVERGED;
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

// Largest matrices first: CTA x of chain y works on the bonds from the centre of the chain outwards, so the
// full-size updates start in the first wave and the small edge bonds fill the tail.
__device__ __forceinline__ int centre_out(int x, int nb) {
  const int c = nb / 2;
  const int k = (x + 1) / 2;
  return (x & 1) ? c - k : c + k;  // x = 0 -> c, 1 -> c-1, 2 -> c+1, ...: a bijection of [0, nb) for odd and even nb
}

template <bool SP>
__global__ void __launch_bounds__(NT, 1) jacobi_rb_kernel(TcDev d, LayerArgs a) {
  Bond b;
  // blockIdx.x = chain, blockIdx.y = rank of the bond in centre-out order
  if (!get_bond(d, a, centre_out(blockIdx.y, a.nb), blockIdx.x, b)) return;
  const int N = uni(b.N), K = uni(b.M < b.N ? b.M : b.N);
  cplx *X = d.Xw + b.slot * d.slot_stride;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + (size_t)2 * QB * d.n2 * sizeof(cplx) + d.n2 * sizeof(double));
  __shared__ double red[32];
  __shared__ int s_rot[2];
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
    sweeps<8, true, SP>(d, b, X, K, N, s_rot, red);
  else if (npl <= 1)
    sweeps<1, false, SP>(d, b, X, K, N, s_rot, red);
  else if (npl <= 2)
    sweeps<2, false, SP>(d, b, X, K, N, s_rot, red);
  else if (npl <= 4)
    sweeps<4, false, SP>(d, b, X, K, N, s_rot, red);
  else
    sweeps<8, false, SP>(d, b, X, K, N, s_rot, red);
}
}  // namespace tcr
