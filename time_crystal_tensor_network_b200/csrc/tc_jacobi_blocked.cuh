// tc_jacobi_blocked.cuh -- K2b fast path: one-sided Jacobi on the rows of the triangular factor with
// the working set staged in shared memory by TMA bulk copies and the stationary rows in registers.
//
// One CTA (16 warps) per matrix, K rows x N columns (N <= 256), rows grouped in blocks of 16 (one
// contiguous 16*N*16 B chunk in the row-major workspace, so a block moves with ONE cp.async.bulk).
// Sweep = for every block p: stage p, rotate its internal pairs, then keep row p_w in the registers of
// warp w while every later block q streams through a double-buffered shared-memory stage:
//   round s of visit (p, q): warp w rotates (p_w, q_{(w+s) mod 16}); q rows are read from and written
//   back to shared memory, p rows never leave registers; rows pass from warp to warp through per-row
//   version counters (acquire / release), a barrier per round only with TC_ROT64 bit 1.
// While visit (p, q) computes, the bulk load of q+1 and the bulk store of q-1 are in flight.
// Rotations are FAST (scaled) rotations: every row carries a squared scale w (true row = sqrt(w) x stored row), the
// factor cos of a rotation goes into w instead of into the 2 N elements, and the update is u -= A g v, v += B conj(g) u:
// 8 instead of 12 FMAs per complex element.  A row's scale is folded back into its elements once per sweep, when its
// block leaves the P stage.  Per pair: 4 KB LDS + 4 KB STS and ~141 FP64 instructions per lane (32 dot, 64 rotation,
// ~45 reduction and set-up); L2 traffic per sweep ~ (K/16)^2/2 blocks.
// ncu of the standard-rotation version (profiles/r01d_ncu_kernels.txt): FP64 pipe 51 % busy, shared-memory
// wavefronts 49 %, issue slots 46 %.
#pragma once
#include "tc_common.cuh"
#include "tc_jacobi.cuh"

namespace tcb {
// Rows per block = warps per CTA (one warp per row of a block), a template parameter of the kernel:
//   16: the default -- 16 warps, 128 registers, three 64 KB stages at 256 columns, one CTA per SM;
//    8: narrow contexts (widest matrix 128 columns, chi_cap <= 64: BASELINE configs 2 and 3) -- 8 warps, three 16 KB
//       stages, TWO CTAs per SM at the same 128 registers: two independent matrices per SM, half as many internal pairs
//       (+4.5 % at config 2; at 256 columns the same split costs 7 %: twice the block visits and L2 traffic).
#ifndef TCB_NARROW_CTAS
#define TCB_NARROW_CTAS 2  // resident CTAs per SM of the narrow instance (3: 80 registers, 16 B of spills, measured equal at config 2)
#endif
constexpr int BR_WIDE = 16, BR_NARROW = 8;
constexpr int MAX_N = 256, MAX_N_NARROW = 128;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok = 0;
  const uint32_t addr = smem_u32(bar);
  unsigned long long spins = 0;
  do {  // exit on a warp vote: a uniform branch for the compiler (see the note on REDUX in the kernel)
    if (++spins > (1ull << 26)) __trap();  // a lost bulk copy must fail loudly, never hang the GPU
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (__any_sync(0xffffffffu, ok == 0));
}
__device__ __forceinline__ void bulk_load(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_store(void *gmem_dst, const void *smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
               "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// the committed bulk stores of this thread have finished READING shared memory (their source may be overwritten)
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

#ifdef TCB_TIMING
// phase timers (diagnostic build only): [0] load+dot, [1] warp reduction, [2] rotation set-up, [3] rotate+store,
// [4] hand-over wait / barrier, [5] pairs rotated, [6] pairs visited, [7] whole kernel (warp 3 of K=256 matrices).
// With -DTCB_COARSE the same slots hold the coarse phases of a sweep instead: [0] row norms at the sweep start,
// [1] wait for the P block, [2] internal pairs, [3] wait for a q block, [4] the 16 rounds of a visit,
// [5] end-of-visit fence + barrier + store, [6] P block back to global.
__device__ unsigned long long g_tcb_timing[8];
#ifdef TCB_COARSE
#define TCB_T(var)
#define TCB_ACC(k, a, b)
#define TCC_T(var) const long long var = clock64()
#define TCC_ACC(k, a, b) tacc[k] += (b) - (a)
#else
#define TCB_T(var) const long long var = clock64()
#define TCB_ACC(k, a, b) tacc[k] += (b) - (a)
#define TCC_T(var)
#define TCC_ACC(k, a, b)
#endif
#else
#define TCB_T(var)
#define TCB_ACC(k, a, b)
#define TCC_T(var)
#define TCC_ACC(k, a, b)
#endif

struct Rot {
  double ar, ai, br, bi, c2, ni, nj;  // A g and B g (see make_rot), cos^2, new squared norms
};

// 1/sqrt(x) for normal positive x without the special-case branch of rsqrt(double): hardware seed (MUFU.RSQ64H,
// ~2^-22) and two Newton steps (error 1.5 e^2 per step: 9e-14, then rounding).  Both arguments in make_rot are
// normal and positive: dd^2 + 4|g|^2 > 0 once the pair passed the threshold, and c^2 in [1/2, 1].
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

// Fast rotation that orthogonalises two rows X_i = sqrt(wi) u, X_j = sqrt(wj) v with true squared norms ai, aj and
// stored-row product (gr, gi) = u . conj(v); false when the pair is below the threshold or one of the rows is
// numerically zero (`alive` false).  True quantities: g = sqrt(wi wj) (gr, gi), dd = aj - ai, 2r = sqrt(dd^2 + 4|g|^2),
// c^2 = 1/2 + |dd|/(4r), s e = sign(dd) g / (2 r c).  The standard update X_i' = c X_i - (s e) X_j,
// X_j' = conj(s e) X_i + c X_j becomes, with the factor c moved into the scales (wi' = c^2 wi, wj' = c^2 wj),
//   u' = u - A (gr, gi) v,   v' = v + B conj(gr, gi) u,   A = wj k,  B = wi k,  k = sign(dd) / (2 r c^2):
// FP64 set-up with two rsqrt and no division / sqrt / |g|; t|g| = |g|^2 k is the squared norm that moves between the rows.
__device__ __forceinline__ bool make_rot(bool alive, double ai, double aj, double wi, double wj, double gr, double gi,
                                         double thr2, double small2, Rot &r, int &big) {
  const double g2 = (wi * wj) * fma(gr, gr, gi * gi);
  const double aa = ai * aj;
  // `big` = the pair keeps the iteration going (stopping rule), whether or not this sweep rotates it
  big = alive && g2 > small2 * aa;
  if (!(alive && g2 > thr2 * aa)) return false;
  const double dd = aj - ai;
  const double rinv = rsqrt_nb(fma(dd, dd, 4.0 * g2));  // 1 / (2r)
  const double c2 = fma(0.5 * fabs(dd), rinv, 0.5);
  const double cinv = rsqrt_nb(c2);
  const double k = copysign(rinv * cinv, dd) * cinv;
  const double A = wj * k, B = wi * k;
  r.ar = A * gr;
  r.ai = A * gi;
  r.br = B * gr;
  r.bi = B * gi;
  r.c2 = c2;
  const double tg = g2 * k;
  r.ni = ai - tg;
  r.nj = aj + tg;
  return true;
}

__device__ __forceinline__ void rot_apply(cplx &u, cplx &v, const Rot &r) {
  // u' = u - (ar + i ai) v ;  v' = v + (br - i bi) u
  cplx un, vn;
  un.x = fma(-r.ar, v.x, fma(r.ai, v.y, u.x));
  un.y = fma(-r.ar, v.y, fma(-r.ai, v.x, u.y));
  vn.x = fma(r.br, u.x, fma(r.bi, u.y, v.x));
  vn.y = fma(r.br, u.y, fma(-r.bi, u.x, v.y));
  u = un;
  v = vn;
}

// sums of a and b over the warp, in every lane.  Packed butterfly: after the first exchange the low half-warp carries
// the partial sums of a and the high half those of b, so rounds 2..5 move one double instead of two; a last exchange
// hands each half the other total (6 shuffles + 5 adds instead of 10 + 10).
__device__ __forceinline__ void warp_sum2(double &a, double &b) {
  const bool hi = (threadIdx.x & 16) != 0;
  double k = hi ? b : a;
  k += __shfl_xor_sync(0xffffffffu, hi ? a : b, 16);
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) k += __shfl_xor_sync(0xffffffffu, k, o);
  const double other = __shfl_xor_sync(0xffffffffu, k, 16);
  a = hi ? other : k;
  b = hi ? k : other;
}

// both rows in shared memory (internal pairs of a block); ni / nj point at the rows' {squared norm, squared scale}
template <int NPL, bool FULL>
__device__ __forceinline__ int pair_smem(cplx *xi, cplx *xj, int N, int lane, double2 *ni, double2 *nj, double dead,
                                         double tol2, double small2) {
  cplx u[NPL], v[NPL];
  double g0 = 0.0, g1 = 0.0, h0 = 0.0, h1 = 0.0;
#pragma unroll
  for (int e = 0; e < NPL; ++e) {
    const int c = lane + 32 * e;
    u[e] = (FULL || c < N) ? xi[c] : cmake(0.0, 0.0);
    v[e] = (FULL || c < N) ? xj[c] : cmake(0.0, 0.0);
    g0 = fma(u[e].x, v[e].x, g0);
    g1 = fma(u[e].y, v[e].y, g1);
    h0 = fma(u[e].y, v[e].x, h0);
    h1 = fma(-u[e].x, v[e].y, h1);
  }
  const double2 si = *ni, sj = *nj;
  double gr = g0 + g1, gi = h0 + h1;
  warp_sum2(gr, gi);
  Rot r;
  int big;
  if (!make_rot(si.x > dead && sj.x > dead, si.x, sj.x, si.y, sj.y, gr, gi, tol2, small2, r, big)) return big;
#pragma unroll
  for (int e = 0; e < NPL; ++e) {
    const int c = lane + 32 * e;
    rot_apply(u[e], v[e], r);
    if (FULL || c < N) {
      xi[c] = u[e];
      xj[c] = v[e];
    }
  }
  if (lane == 0) {
    *ni = make_double2(r.ni, si.y * r.c2);
    *nj = make_double2(r.nj, sj.y * r.c2);
  }
  return big | (1 << 16);  // bits 0..15 count the pairs that keep the iteration going, bits 16.. the rotations made
}

// row i in registers (u, its squared norm ai and squared scale wi too), row j in shared memory.  The row is loaded
// before anything is decided: its LDS latency overlaps the norm load, and a numerically zero row (rare) just costs its
// dot product.
template <int NPL, bool FULL>
__device__ __forceinline__ int pair_reg(cplx (&u)[NPL], cplx *xj, int N, int lane, double &ai, double &wi, double2 *nj,
                                        double dead, double tol2, double small2
#ifdef TCB_TIMING
                                        , long long (&tacc)[8]
#endif
) {
  TCB_T(t0);
  cplx v[NPL];
  double g0 = 0.0, g1 = 0.0, h0 = 0.0, h1 = 0.0;
#pragma unroll
  for (int e = 0; e < NPL; ++e) {
    const int c = lane + 32 * e;
    v[e] = (FULL || c < N) ? xj[c] : cmake(0.0, 0.0);
    g0 = fma(u[e].x, v[e].x, g0);
    g1 = fma(u[e].y, v[e].y, g1);
    h0 = fma(u[e].y, v[e].x, h0);
    h1 = fma(-u[e].x, v[e].y, h1);
  }
  const double2 sj = *nj;
  double gr = g0 + g1, gi = h0 + h1;
  TCB_T(t1);
  warp_sum2(gr, gi);
  TCB_T(t2);
  TCB_ACC(0, t0, t1);
  TCB_ACC(1, t1, t2);
#ifdef TCB_TIMING
  tacc[6] += 1;
#endif
  Rot r;
  int big;
  if (!make_rot(ai > dead && sj.x > dead, ai, sj.x, wi, sj.y, gr, gi, tol2, small2, r, big)) return big;
  TCB_T(t3);
#pragma unroll
  for (int e = 0; e < NPL; ++e) {
    const int c = lane + 32 * e;
    rot_apply(u[e], v[e], r);
    if (FULL || c < N) xj[c] = v[e];
  }
  ai = r.ni;
  wi *= r.c2;
  if (lane == 0) *nj = make_double2(r.nj, sj.y * r.c2);
  TCB_T(t4);
  TCB_ACC(2, t2, t3);
  TCB_ACC(3, t3, t4);
#ifdef TCB_TIMING
  tacc[5] += 1;
#endif
  return big | (1 << 16);
}

template <int NPL, bool FULL, int BR>
__device__ void sweeps(const TcDev &d, const Bond &b, cplx *X, int K, int N, int *s_rot, double *red) {
  constexpr int NW = BR, NT = BR * 32;
  // carve the stage out of dynamic shared memory here so that the compiler keeps the shared address space
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // plain pointers computed from the shared array (no struct / array of pointers: an indexed pointer
  // array degrades the accesses to generic LD/ST)
  cplx *const sP = reinterpret_cast<cplx *>(smem_raw);
  cplx *const sQ = sP + (size_t)BR * N;  // Q[buf] = sQ + buf * BR * N
  unsigned char *const tail = smem_raw + (size_t)3 * BR * d.n2 * sizeof(cplx);
  double2 *const s_nrm2 = reinterpret_cast<double2 *>(tail);  // per row: {true squared norm, squared scale w}
  uint64_t *const barP = reinterpret_cast<uint64_t *>(tail + d.n2 * sizeof(double2));
  uint64_t *const barQ = barP + 1;
  int *const s_ver = reinterpret_cast<int *>(barP + 4);
  int *const s_fin = reinterpret_cast<int *>(barP + 3);  // [2]: warps that have left the current visit of each Q stage
  const int tid = threadIdx.x, lane = tid & 31, warp = __reduce_max_sync(0xffffffffu, tid >> 5);  // provably uniform
#ifdef TCB_TIMING
  long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const long long tk0 = clock64();
#endif
  const int nblk = (K + BR - 1) / BR;
  const double tol = 2.0 * sqrt((double)N) * 2.220446049250313e-16;
  const double tol2_final = tol * tol;
  bool thr_off = false;
  const uint32_t row_bytes = (uint32_t)N * sizeof(cplx);
  uint32_t phP = 0, phQ0 = 0, phQ1 = 0;  // scalars, not arrays: dynamic indexing would put them in local memory
  int verBase0 = 0, verBase1 = 0;
  int finBase0 = 0, finBase1 = 0;
  const bool lockstep = (d.rot64 & 2) != 0;
  // TC_ROT64 bit 2 (opt-in, A/B): visits follow each other without a CTA-wide barrier -- the warp that leaves a visit
  // last stores the block and reloads the stage for the visit after next.  Measured on the B200: 36.3 -> 36.0 ms per
  // layer launch (0.7 %), parity and the LAPACK stress test green; not worth making the default.
  const bool nobar = !lockstep && (d.rot64 & 4) != 0;
  double dead = 0.0;
  int sweep = 0;
  for (; sweep < tcj::MAX_SWEEPS; ++sweep) {
    // all bulk stores of the previous sweep must have landed before rows are re-read
    TCC_T(c0);
    if (tid == 0) bulk_wait_all();
    __syncthreads();
    for (int r = warp; r < K; r += NW) {
      const cplx *row = X + (size_t)r * N;
      double s = 0.0;
      for (int c = lane; c < N; c += 32) s += cabs2(__ldcg(reinterpret_cast<const double2 *>(row + c)));
      s = tcj::warp_sum(s);
      if (lane == 0) s_nrm2[r] = make_double2(s, 1.0);  // every row was folded when its block left the P stage
    }
    if (tid == 0) *s_rot = 0;
    __syncthreads();
    if (sweep == 0) {
      double p = 0.0;
      for (int r = tid; r < K; r += NT) p += s_nrm2[r].x;
      dead = tcj::DEAD_REL2 * block_sum(p, red);
    }
    // threshold Jacobi: the early sweeps rotate only the pairs that are far from orthogonal (|g|^2 / (a_i a_j) above
    // 3e-3, 3e-4, 3e-5, 3e-6 in sweeps 0..3, TcDev::thr_sched); a small rotation made now is undone by the large ones around it and has
    // to be made again.  Same final accuracy and sweep count, a fifth fewer rotations (NumPy model on TEBD matrices).
    // The parameter of the pair functions called tol2 is this sweep's rotation threshold from here on.
    const double tol2 = (sweep < 6 && !thr_off) ? fmax(tol2_final, d.thr_sched[sweep]) : tol2_final;
    // stopping rule: a sweep that rotated every pair above the final tolerance and found them all below 1e-8 ends the
    // iteration (quadratic convergence: what is left is below 1e-16); a threshold sweep has skipped pairs, so there
    // anything above the final tolerance keeps the iteration going
    const double small2 = tol2 > tol2_final ? tol2_final : fmax(tol2_final, d.small_rel2);
    int nrot = 0;
    TCC_T(c1);
    TCC_ACC(0, c0, c1);
    for (int p = 0; p < nblk; ++p) {
      TCC_T(c2);
      const int rowsP = min(BR, K - p * BR);
      cplx *gP = X + (size_t)p * BR * N;
      if (tid == 0) {
        bulk_wait_all();  // the previous stores out of P / Q have finished reading shared memory
        mbar_expect_tx(barP, rowsP * row_bytes);
        bulk_load(sP, gP, rowsP * row_bytes, barP);
        if (p + 1 < nblk) {
          const int rq = min(BR, K - (p + 1) * BR);
          mbar_expect_tx(&barQ[0], rq * row_bytes);
          bulk_load(sQ, X + (size_t)(p + 1) * BR * N, rq * row_bytes, &barQ[0]);
        }
        if (nobar && p + 2 < nblk) {
          const int rq = min(BR, K - (p + 2) * BR);
          mbar_expect_tx(&barQ[1], rq * row_bytes);
          bulk_load(sQ + (size_t)BR * N, X + (size_t)(p + 2) * BR * N, rq * row_bytes, &barQ[1]);
        }
      }
      mbar_wait(barP, phP);
      phP ^= 1;
      TCC_T(c3);
      TCC_ACC(1, c2, c3);
      // ---- internal pairs: circle method on the rows of a block, one pair per warp, BR/2 warps per block.
      // Blocks are handled two at a time (p even: block p in P on warps 0..BR/2-1 and block p+1, already
      // prefetched into stage 0, on warps BR/2..BR-1) so that no warp idles; any order of the pairs within a
      // sweep is a valid cyclic Jacobi ordering.
      bool q0_ready = false;
      if ((p & 1) == 0) {
        const bool pairUp = p + 1 < nblk;
        const int rowsN = pairUp ? min(BR, K - (p + 1) * BR) : 0;
        if (pairUp) {
          mbar_wait(&barQ[0], phQ0);
          phQ0 ^= 1;
          q0_ready = true;
        }
        const int half = warp >= BR / 2;
        const int wl = warp - half * (BR / 2);
        const int rowsB = half ? rowsN : rowsP;
        cplx *blk = half ? sQ : sP;
        double2 *nb = s_nrm2 + (p + half) * BR;
        const int rmax = max(rowsP, rowsN) - 1;
        for (int r = 0; r < rmax; ++r) {
          if (r < rowsB - 1 && wl < rowsB / 2) {
            int i, j;
            tcj::rr_pair(rowsB, r, wl, i, j);
            nrot += pair_smem<NPL, FULL>(blk + (size_t)i * N, blk + (size_t)j * N, N, lane, nb + i, nb + j, dead, tol2,
                                         small2);
          }
          __syncthreads();
        }
      }
      TCC_T(c4);
      TCC_ACC(2, c3, c4);
      // ---- row p_w into registers
      cplx u[NPL];
      const bool haveP = warp < rowsP;
#pragma unroll
      for (int e = 0; e < NPL; ++e) {
        const int c = lane + 32 * e;
        u[e] = (haveP && (FULL || c < N)) ? sP[(size_t)warp * N + c] : cmake(0.0, 0.0);
      }
      // squared norm and squared scale of the stationary row, in registers for the visits
      const double2 sP0 = haveP ? s_nrm2[p * BR + warp] : make_double2(0.0, 1.0);
      double aP = sP0.x, wP = sP0.y;
      // ---- every later block streams through Q
      for (int q = p + 1; q < nblk; ++q) {
        const int buf = (q - p - 1) & 1;
        const int rowsQ = min(BR, K - q * BR);
        // the prefetch of block q+1 is issued in the middle of this visit (round BR/2): by then the store
        // of block q-1 out of the other stage has long completed, so its wait costs nothing
        auto prefetch_next = [&]() {
          if (tid == 0 && q + 1 < nblk) {
            bulk_wait_all();
            const int rq = min(BR, K - (q + 1) * BR);
            mbar_expect_tx(&barQ[buf ^ 1], rq * row_bytes);
            bulk_load((sQ + (size_t)(buf ^ 1) * BR * N), X + (size_t)(q + 1) * BR * N, rq * row_bytes,
                      &barQ[buf ^ 1]);
          }
        };
        TCC_T(c5);
        if (!(q == p + 1 && q0_ready)) {  // stage 0 of the first visit may already have been consumed above
          mbar_wait(&barQ[buf], buf ? phQ1 : phQ0);
          if (buf)
            phQ1 ^= 1;
          else
            phQ0 ^= 1;
        }
        cplx *Q = (sQ + (size_t)buf * BR * N);
        TCC_T(c6);
        TCC_ACC(3, c5, c6);
        if (lockstep) {
          for (int s = 0; s < BR; ++s) {
            if (s == BR / 2) prefetch_next();
            const int jq = (warp + s) & (BR - 1);
            if (haveP && jq < rowsQ)
              nrot += pair_reg<NPL, FULL>(u, Q + (size_t)jq * N, N, lane, aP, wP, s_nrm2 + q * BR + jq, dead, tol2, small2
#ifdef TCB_TIMING
                                          , tacc
#endif
              );
            TCB_T(tb0);
            __syncthreads();
            TCB_T(tb1);
            TCB_ACC(4, tb0, tb1);
          }
        } else {
          // point-to-point hand-over: row jq carries a version counter; round s of this visit may touch it
          // once rounds 0..s-1 have released it (the previous holder is warp w+1).  No CTA-wide barrier
          // inside the visit, so the warps drift apart and one warp's scalar rotation set-up overlaps the
          // FP64-heavy dot / rotate phases of the others.
          const uint32_t vaddr = smem_u32(s_ver + buf * BR);
          const int base = buf ? verBase1 : verBase0;
          for (int s = 0; s < BR; ++s) {
            if (!nobar && s == BR / 2) prefetch_next();
            const int jq = (warp + s) & (BR - 1);
            TCB_T(tw0);
            if (s > 0) {
              int v;
              unsigned long long spins = 0;
              do {
                asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(vaddr + 4 * jq) : "memory");
                if (++spins > (1ull << 24)) __trap();
              } while (__any_sync(0xffffffffu, v < base + s));
            }
            TCB_T(tw1);
            TCB_ACC(4, tw0, tw1);
            if (haveP && jq < rowsQ)
              nrot += pair_reg<NPL, FULL>(u, Q + (size_t)jq * N, N, lane, aP, wP, s_nrm2 + q * BR + jq, dead, tol2, small2
#ifdef TCB_TIMING
                                          , tacc
#endif
              );
            __syncwarp();
            if (lane == 0)
              asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(vaddr + 4 * jq), "r"(base + s + 1) : "memory");
          }
          if (buf)
            verBase1 = base + BR;
          else
            verBase0 = base + BR;
        }
        TCC_T(c7);
        TCC_ACC(4, c6, c7);
        fence_async_smem();
        if (nobar) {
          // no barrier: every warp signs off on this stage; the last one stores the block and, once the store has read
          // the stage, loads the block of the visit after next into it.  The others are already in the next visit.
          __syncwarp();
          if (lane == 0) {
            const int fb = buf ? finBase1 : finBase0;
            int old;
            asm volatile("atom.acq_rel.cta.shared.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(smem_u32(s_fin + buf)) : "memory");
            if (old == fb + NW - 1) {
              fence_async_smem();
              bulk_store(X + (size_t)q * BR * N, Q, rowsQ * row_bytes);
              if (q + 2 < nblk) {
                bulk_wait_read();
                const int rq = min(BR, K - (q + 2) * BR);
                mbar_expect_tx(&barQ[buf], rq * row_bytes);
                bulk_load(Q, X + (size_t)(q + 2) * BR * N, rq * row_bytes, &barQ[buf]);
              }
            }
          }
          if (buf)
            finBase1 += NW;
          else
            finBase0 += NW;
          __syncwarp();
        } else {
          __syncthreads();
          if (tid == 0) bulk_store(X + (size_t)q * BR * N, Q, rowsQ * row_bytes);
        }
        TCC_T(c8);
        TCC_ACC(5, c7, c8);
      }
      TCC_T(c9);
      // ---- block p back to global
      if (haveP) {
        // the row's scale goes back into its elements here, once per sweep (every block is the P block once)
        const double sc = sqrt(wP);
        if (lane == 0) s_nrm2[p * BR + warp] = make_double2(aP, 1.0);
#pragma unroll
        for (int e = 0; e < NPL; ++e) {
          const int c = lane + 32 * e;
          if (FULL || c < N) sP[(size_t)warp * N + c] = cscale(u[e], sc);
        }
      }
      fence_async_smem();
      if (nobar && lane == 0) bulk_wait_all();  // the block stores this warp issued as the last one out of a visit
      __syncthreads();
      if (tid == 0) bulk_store(gP, sP, rowsP * row_bytes);
      TCC_T(c10);
      TCC_ACC(6, c9, c10);
    }
    if (lane == 0 && nrot) atomicAdd(s_rot, nrot);
    __syncthreads();
    const int both = __reduce_max_sync(0xffffffffu, *s_rot);
    const int tot = both & 0xffff;
    __syncthreads();
    if (tot == 0) break;
    // a threshold sweep that found nothing to rotate (a nearly orthogonal matrix: weak gates, small chi) would be
    // followed by more of the same: go straight to the final tolerance
    if (((unsigned)both >> 16) == 0) thr_off = true;
  }
#ifdef TCB_TIMING
  if (K == 256 && N == 256 && warp == 3 && lane == 0) {
    tacc[7] = clock64() - tk0;
    for (int k = 0; k < 8; ++k) atomicAdd(&g_tcb_timing[k], (unsigned long long)tacc[k]);
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

// dynamic smem: 3 * BR * n2 cplx (P, Q0, Q1) + n2 double2 + 3 mbarriers + counters.
// <MAXNPL = 8, BR = 16>: any matrix up to 256 columns, one CTA per SM.  <MAXNPL = 4, BR = 8>: contexts whose widest matrix
// has 128 columns, two CTAs per SM.  Both 128 registers.
template <int MAXNPL, int BR>
__global__ void __launch_bounds__(BR * 32, BR == 16 ? 1 : TCB_NARROW_CTAS) jacobi_blocked_kernel(TcDev d, LayerArgs a) {
  Bond b;
  // blockIdx.x = chain, blockIdx.y = rank of the bond in centre-out order (largest matrices first)
  if (!get_bond(d, a, centre_out(blockIdx.y, a.nb), blockIdx.x, b)) return;
  // sizes come from the chi table in memory: the same in every lane, but only a warp reduction (REDUX) makes them
  // provably uniform for the compiler -- a branch on a loaded value counts as divergent and every later shuffle then
  // carries a BRA.DIV divergence check
  const int N = __reduce_max_sync(0xffffffffu, b.N), K = __reduce_max_sync(0xffffffffu, b.M < b.N ? b.M : b.N);
  cplx *X = d.Xw + b.slot * d.slot_stride;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + (size_t)3 * BR * d.n2 * sizeof(cplx) + d.n2 * sizeof(double2));
  __shared__ double red[32];
  __shared__ int s_rot;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    fence_async_smem();
  }
  if (threadIdx.x < 2 * BR) reinterpret_cast<int *>(bars + 4)[threadIdx.x] = 0;
  if (threadIdx.x < 2) reinterpret_cast<int *>(bars + 3)[threadIdx.x] = 0;  // visit sign-off counters
  __syncthreads();
  const int npl = (N + 31) / 32;
  if (MAXNPL >= 8 && N == 256)
    sweeps<8, true, BR>(d, b, X, K, N, &s_rot, red);
  else if (npl <= 1)
    sweeps<1, false, BR>(d, b, X, K, N, &s_rot, red);
  else if (npl <= 2)
    sweeps<2, false, BR>(d, b, X, K, N, &s_rot, red);
  else if (MAXNPL <= 4 && N == 128)
    sweeps<4, true, BR>(d, b, X, K, N, &s_rot, red);
  else if (npl <= 4)
    sweeps<4, false, BR>(d, b, X, K, N, &s_rot, red);
  else if (MAXNPL >= 8)
    sweeps<8, false, BR>(d, b, X, K, N, &s_rot, red);
}
}  // namespace tcb
