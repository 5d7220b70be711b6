// tc_jacobi_team.cuh -- K2b for wide matrices (chi_cap > 128: theta up to 512 x 512, BASELINE config 4) and, as an
// experiment switch (TC_JACOBI=team), for the 256-column matrices of the metric shape.
//
// Same algorithm as tc_jacobi_blocked.cuh (one-sided Jacobi on the rows of the triangular factor, fast scaled
// rotations, threshold sweeps, quadratic-convergence stopping rule, row blocks staged by cp.async.bulk, stationary rows
// in registers, q rows handed from holder to holder through version counters), with two changes of structure:
//
//  * a row is worked on by a TEAM of two warps, each owning one half of the columns (32 NPL columns per warp): a
//    512-column row is 2 x 8 complex per lane, the register budget of the 256-column kernel, and a CTA of 16 warps holds
//    8 teams = blocks of 8 rows (64 KB at N = 512: P + two Q stages = 192 KB).  The two halves of a dot product meet
//    through a double-buffered shared-memory slot and a named barrier of 64 threads; both warps then compute the same
//    rotation.  Each warp only ever touches its own half of any row, so the hand-over counters are per (row, half).
//  * a matrix can be shared by a thread-block CLUSTER of CS CTAs (config 4 is one chain: ~31 matrices per layer on 148
//    SMs).  The row blocks are dealt into 2 CS groups; a sweep is (a) every CTA sweeps two groups on its own (all pairs
//    inside a group), (b) 2 CS - 1 rounds of a round-robin tournament between groups, CS disjoint group pairs per round,
//    one per CTA: for every block p of the first group, every block q of the second streams through the Q stages.
//    A cluster barrier separates the rounds; rows live in global memory / L2 between tasks (folded: scale 1), their
//    squared norms in the slot's `ww` row, the rotation counter of a sweep in its `knew` entry.
//    L2 traffic per pair: 2 x 64 KB per 64 pairs = 2 KB, against 32 KB for the warp-per-pair cluster kernel this replaces.
#pragma once
#include "tc_common.cuh"
#include "tc_jacobi.cuh"
#include "tc_jacobi_blocked.cuh"

namespace tct {
constexpr int BRW = 8;                    // rows per block = teams per CTA
constexpr int NW = 2 * BRW, NT = NW * 32;  // two warps per team
using tcb::bulk_load;
using tcb::bulk_store;
using tcb::bulk_wait_all;
using tcb::fence_async_smem;
using tcb::make_rot;
using tcb::mbar_expect_tx;
using tcb::mbar_init;
using tcb::mbar_wait;
using tcb::Rot;
using tcb::rot_apply;
using tcb::smem_u32;
using tcb::warp_sum2;

// named barrier of the 64 threads of a team (identifiers 1..BRW; 0 is __syncthreads)
__device__ __forceinline__ void team_bar(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// dynamic shared memory: P, Q0, Q1 (BRW rows of n2 columns each), then per-row {norm^2, scale^2}, 4 mbarrier slots,
// hand-over counters [2 stages][BRW rows][2 halves], dot-product exchange slots [BRW teams][2 parities][2 halves]
__host__ __device__ inline size_t smem_bytes(int n2) {
  return (size_t)3 * BRW * n2 * sizeof(cplx) + (size_t)n2 * sizeof(double2) + 4 * sizeof(uint64_t) +
         2 * BRW * 2 * sizeof(int) + (size_t)BRW * 4 * sizeof(double2);
}

struct Sm {
  cplx *sP, *sQ;
  double2 *nw;
  uint64_t *barP, *barQ;
  int *ver;
  double2 *xch;
};

// sum of (a, b) over the 64 lanes of a team, in every lane of both warps (bitwise the same in both: x + y = y + x)
__device__ __forceinline__ void team_sum2(double &a, double &b, double2 *xch, int half, int &par, int barid, int lane) {
  warp_sum2(a, b);
  if (lane == 0) xch[par * 2 + half] = make_double2(a, b);
  team_bar(barid);
  const double2 o = xch[par * 2 + (half ^ 1)];
  a += o.x;
  b += o.y;
  par ^= 1;  // the next exchange uses the other slot: a warp one exchange ahead cannot overwrite what its partner still reads
}

// both rows in shared memory (internal pairs of a block); xi / xj already point at this warp's half
template <int NPL, bool FULL>
__device__ __forceinline__ int pair_smem(cplx *xi, cplx *xj, int ncol, int lane, double2 *ni, double2 *nj, double dead,
                                         double tol2, double small2, double2 *xch, int half, int &par, int barid) {
  cplx u[NPL], v[NPL];
  double g0 = 0.0, g1 = 0.0, h0 = 0.0, h1 = 0.0;
#pragma unroll
  for (int e = 0; e < NPL; ++e) {
    const int c = lane + 32 * e;
    u[e] = (FULL || c < ncol) ? xi[c] : cmake(0.0, 0.0);
    v[e] = (FULL || c < ncol) ? xj[c] : cmake(0.0, 0.0);
    g0 = fma(u[e].x, v[e].x, g0);
    g1 = fma(u[e].y, v[e].y, g1);
    h0 = fma(u[e].y, v[e].x, h0);
    h1 = fma(-u[e].x, v[e].y, h1);
  }
  const double2 si = *ni, sj = *nj;
  double gr = g0 + g1, gi = h0 + h1;
  team_sum2(gr, gi, xch, half, par, barid, lane);
  Rot r;
  int big;
  if (!make_rot(si.x > dead && sj.x > dead, si.x, sj.x, si.y, sj.y, gr, gi, tol2, small2, r, big)) return half ? 0 : big;
#pragma unroll
  for (int e = 0; e < NPL; ++e) {
    const int c = lane + 32 * e;
    rot_apply(u[e], v[e], r);
    if (FULL || c < ncol) {
      xi[c] = u[e];
      xj[c] = v[e];
    }
  }
  // both warps computed the same numbers; the rows' {norm, scale} are read again only after the round's barrier
  if (lane == 0 && half == 0) {
    *ni = make_double2(r.ni, si.y * r.c2);
    *nj = make_double2(r.nj, sj.y * r.c2);
  }
  return half ? 0 : (big | (1 << 16));
}

// this warp's half of row i in registers, its half of row j in shared memory
template <int NPL, bool FULL>
__device__ __forceinline__ int pair_reg(cplx (&u)[NPL], cplx *xj, int ncol, int lane, double &ai, double &wi, double2 *nj,
                                        double dead, double tol2, double small2, double2 *xch, int half, int &par, int barid) {
  cplx v[NPL];
  double g0 = 0.0, g1 = 0.0, h0 = 0.0, h1 = 0.0;
#pragma unroll
  for (int e = 0; e < NPL; ++e) {
    const int c = lane + 32 * e;
    v[e] = (FULL || c < ncol) ? xj[c] : cmake(0.0, 0.0);
    g0 = fma(u[e].x, v[e].x, g0);
    g1 = fma(u[e].y, v[e].y, g1);
    h0 = fma(u[e].y, v[e].x, h0);
    h1 = fma(-u[e].x, v[e].y, h1);
  }
  const double2 sj = *nj;
  double gr = g0 + g1, gi = h0 + h1;
  team_sum2(gr, gi, xch, half, par, barid, lane);
  Rot r;
  int big;
  if (!make_rot(ai > dead && sj.x > dead, ai, sj.x, wi, sj.y, gr, gi, tol2, small2, r, big)) return half ? 0 : big;
#pragma unroll
  for (int e = 0; e < NPL; ++e) {
    const int c = lane + 32 * e;
    rot_apply(u[e], v[e], r);
    if (FULL || c < ncol) xj[c] = v[e];
  }
  ai = r.ni;
  wi *= r.c2;
  // both warps store the same 16 bytes: the next holder's warp of either half acquires from the warp of its own half
  if (lane == 0) *nj = make_double2(r.nj, sj.y * r.c2);
  return half ? 0 : (big | (1 << 16));
}

struct Ctx {  // per-thread loop state that survives across tasks
  uint32_t phP, phQ0, phQ1;
  int verBase0, verBase1, par;
  int nrot;
};

// One pass with block p in the P stage: optional internal pairs (internal = 1: of block p alone; 2: of p and, side by
// side, of the first q block), then every block q in [q_lo, q_hi) streams through the Q stages.  fold_q: the scales of
// the q rows are folded into their elements before the block is stored (last pass of a task over these q blocks).
template <int NPL, bool FULL>
__device__ void p_pass(const Sm &sm, Ctx &cx, cplx *X, int K, int N, int p, int q_lo, int q_hi, int internal, bool fold_q,
                       double dead, double tol2, double small2) {
  const int tid = threadIdx.x, lane = tid & 31, warp = __reduce_max_sync(0xffffffffu, tid >> 5);
  const int team = warp >> 1, half = warp & 1, barid = 1 + team;
  const int colbase = half * 32 * NPL;
  const int ncol = N - colbase;  // columns of this half (may be <= 0 for narrow matrices: every access is masked)
  const uint32_t row_bytes = (uint32_t)N * sizeof(cplx);
  double2 *xch = sm.xch + team * 4;
  const int rowsP = min(BRW, K - p * BRW);
  cplx *gP = X + (size_t)p * BRW * N;
  cplx *sP = sm.sP, *sQ = sm.sQ;
  if (tid == 0) {
    bulk_wait_all();  // earlier stores out of P / Q have finished reading shared memory
    mbar_expect_tx(sm.barP, rowsP * row_bytes);
    bulk_load(sP, gP, rowsP * row_bytes, sm.barP);
    if (q_lo < q_hi) {
      const int rq = min(BRW, K - q_lo * BRW);
      mbar_expect_tx(&sm.barQ[0], rq * row_bytes);
      bulk_load(sQ, X + (size_t)q_lo * BRW * N, rq * row_bytes, &sm.barQ[0]);
    }
  }
  mbar_wait(sm.barP, cx.phP);
  cx.phP ^= 1;
  bool q0_ready = false;
  if (internal) {
    // circle method on the rows of a block, one pair per team, BRW / 2 teams per block; block p on teams 0..BRW/2-1
    // and, when the first q block is p + 1 (sweep inside a group), that block on the other teams
    const bool pairUp = internal == 2 && q_lo < q_hi;
    const int rowsN = pairUp ? min(BRW, K - q_lo * BRW) : 0;
    if (pairUp) {
      mbar_wait(&sm.barQ[0], cx.phQ0);
      cx.phQ0 ^= 1;
      q0_ready = true;
    }
    const int hb = team >= BRW / 2;
    const int tl = team - hb * (BRW / 2);
    const int rowsB = hb ? rowsN : rowsP;
    cplx *blk = hb ? sQ : sP;
    double2 *nb = sm.nw + (hb ? q_lo : p) * BRW;
    const int rmax = max(rowsP, rowsN) - 1;
    for (int r = 0; r < rmax; ++r) {
      if (r < rowsB - 1 && tl < rowsB / 2) {
        int i, j;
        tcj::rr_pair(rowsB, r, tl, i, j);
        cx.nrot += pair_smem<NPL, FULL>(blk + (size_t)i * N + colbase, blk + (size_t)j * N + colbase, ncol, lane, nb + i,
                                        nb + j, dead, tol2, small2, xch, half, cx.par, barid);
      }
      __syncthreads();
    }
  }
  // ---- this warp's half of row p_team into registers
  cplx u[NPL];
  const bool haveP = team < rowsP;
#pragma unroll
  for (int e = 0; e < NPL; ++e) {
    const int c = lane + 32 * e;
    u[e] = (haveP && (FULL || c < ncol)) ? sP[(size_t)team * N + colbase + c] : cmake(0.0, 0.0);
  }
  const double2 sP0 = haveP ? sm.nw[p * BRW + team] : make_double2(0.0, 1.0);
  double aP = sP0.x, wP = sP0.y;
  for (int q = q_lo; q < q_hi; ++q) {
    const int buf = (q - q_lo) & 1;
    const int rowsQ = min(BRW, K - q * BRW);
    if (!(q == q_lo && q0_ready)) {
      mbar_wait(&sm.barQ[buf], buf ? cx.phQ1 : cx.phQ0);
      if (buf)
        cx.phQ1 ^= 1;
      else
        cx.phQ0 ^= 1;
    }
    cplx *Q = sQ + (size_t)buf * BRW * N;
    const uint32_t vaddr = smem_u32(sm.ver + (buf * BRW) * 2 + half);
    const int base = buf ? cx.verBase1 : cx.verBase0;
    for (int s = 0; s < BRW; ++s) {
      if (s == BRW / 2 && tid == 0 && q + 1 < q_hi) {
        // prefetch of block q + 1 in the middle of this visit: the store of block q - 1 out of the other stage has
        // long completed by now
        bulk_wait_all();
        const int rq = min(BRW, K - (q + 1) * BRW);
        mbar_expect_tx(&sm.barQ[buf ^ 1], rq * row_bytes);
        bulk_load(sQ + (size_t)(buf ^ 1) * BRW * N, X + (size_t)(q + 1) * BRW * N, rq * row_bytes, &sm.barQ[buf ^ 1]);
      }
      const int jq = (team + s) & (BRW - 1);
      if (s > 0) {  // this half of row jq was released by the warp of the same half of the previous holder
        int v;
        unsigned long long spins = 0;
        do {
          asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(vaddr + 8 * jq) : "memory");
          if (++spins > (1ull << 24)) __trap();
        } while (__any_sync(0xffffffffu, v < base + s));
      }
      // haveP and jq < rowsQ are the same in both warps of a team: the team barrier inside pair_reg is safe
      if (haveP && jq < rowsQ)
        cx.nrot += pair_reg<NPL, FULL>(u, Q + (size_t)jq * N + colbase, ncol, lane, aP, wP, sm.nw + q * BRW + jq, dead,
                                       tol2, small2, xch, half, cx.par, barid);
      __syncwarp();
      if (lane == 0)
        asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(vaddr + 8 * jq), "r"(base + s + 1) : "memory");
    }
    if (buf)
      cx.verBase1 = base + BRW;
    else
      cx.verBase0 = base + BRW;
    if (fold_q) {
      __syncthreads();  // every rotation of this visit is done
      for (int r = warp; r < rowsQ; r += NW) {
        const double2 s0 = sm.nw[q * BRW + r];
        const double sc = sqrt(s0.y);
        for (int c = lane; c < N; c += 32) Q[(size_t)r * N + c] = cscale(Q[(size_t)r * N + c], sc);
      }
      __syncthreads();
      for (int r = tid; r < rowsQ; r += NT) sm.nw[q * BRW + r].y = 1.0;
    }
    fence_async_smem();
    __syncthreads();
    if (tid == 0) bulk_store(X + (size_t)q * BRW * N, Q, rowsQ * row_bytes);
  }
  // ---- block p back to global, scale folded into the elements
  if (haveP) {
    const double sc = sqrt(wP);
#pragma unroll
    for (int e = 0; e < NPL; ++e) {
      const int c = lane + 32 * e;
      if (FULL || c < ncol) sP[(size_t)team * N + colbase + c] = cscale(u[e], sc);
    }
  }
  fence_async_smem();
  __syncthreads();
  // the row's {norm, scale} is reset only now, behind the barrier: in a pass without visits nothing else separates this
  // write from the partner warp's read of the scale at the top of the pass (a partner that read 1 would leave its half
  // of the row unscaled: the race behind the occasionally too large smallest singular values of the first version)
  if (haveP && lane == 0 && half == 0) sm.nw[p * BRW + team] = make_double2(aP, 1.0);
  if (tid == 0) bulk_store(gP, sP, rowsP * row_bytes);
}

// norms of the blocks [b_lo, b_hi) between shared memory and the slot's global row (tasks of different CTAs touch
// disjoint blocks; the cluster barrier between rounds orders the accesses)
__device__ __forceinline__ void norms_in(const Sm &sm, const double *gn, int K, int b_lo, int b_hi) {
  for (int r = b_lo * BRW + threadIdx.x; r < min(K, b_hi * BRW); r += NT) sm.nw[r] = make_double2(__ldcg(gn + r), 1.0);
}
__device__ __forceinline__ void norms_out(const Sm &sm, double *gn, int K, int b_lo, int b_hi) {
  for (int r = b_lo * BRW + threadIdx.x; r < min(K, b_hi * BRW); r += NT) __stcg(gn + r, sm.nw[r].x);
}

template <int NPL, bool FULL>
__device__ void sweeps(const TcDev &d, const Bond &b, cplx *X, int K, int N, int CS, int crank, const Sm &sm, double *red) {
  const int tid = threadIdx.x, lane = tid & 31, warp = __reduce_max_sync(0xffffffffu, tid >> 5);
  double *gn = d.ww + b.slot * d.n2;  // squared norms of the folded rows (the singular values at the end)
  int *gcnt = d.knew + b.slot;        // rotations that keep the iteration going, summed over the cluster per sweep
  const int nblk = (K + BRW - 1) / BRW;
  // groups of blocks: one when a CTA has the matrix to itself, 2 CS otherwise (a round-robin tournament between an even
  // number of groups keeps every CTA busy in every round)
  const int ng = CS > 1 ? 2 * CS : 1;
  const int gs = (nblk + ng - 1) / ng;  // blocks per group (the last groups may be short or empty)
  const double tol = 2.0 * sqrt((double)N) * 2.220446049250313e-16;
  const double tol2_final = tol * tol;
  const int gw = crank * NW + warp, nwc = CS * NW;  // this warp among the warps of the cluster
  Ctx cx{0, 0, 0, 0, 0, 0, 0};
  bool thr_off = false;
  double dead = 0.0;
  int sweep = 0;
  for (; sweep < tcj::MAX_SWEEPS; ++sweep) {
    // fresh norms of the (folded) rows at the start of every sweep
    if (tid == 0) bulk_wait_all();
    __syncthreads();
    fence_proxy_async_all();
    for (int r = gw; r < K; r += nwc) {
      const cplx *row = X + (size_t)r * N;
      double s = 0.0;
      for (int c = lane; c < N; c += 32) s += cabs2(__ldcg(reinterpret_cast<const double2 *>(row + c)));
      s = tcj::warp_sum(s);
      if (lane == 0) __stcg(gn + r, s);
    }
    if (crank == 0 && tid == 0) __stcg(gcnt, 0);
    tcj::cluster_sync_all();
    if (sweep == 0) {
      double p = 0.0;
      for (int r = tid; r < K; r += NT) p += __ldcg(gn + r);
      dead = tcj::DEAD_REL2 * block_sum(p, red);
    }
    const double tol2 = (sweep < 6 && !thr_off) ? fmax(tol2_final, d.thr_sched[sweep]) : tol2_final;
    const double small2 = tol2 > tol2_final ? tol2_final : fmax(tol2_final, d.small_rel2);
    cx.nrot = 0;
    // ---- (a) all pairs inside a group: groups crank and crank + CS.  The internal pairs of two neighbouring blocks are
    // rotated side by side, the blocks paired from the END of the group ((b1-2, b1-1), (b1-4, b1-3), ...; the first block
    // of a group with an odd number of blocks is on its own), so that the last block of the group gets its internal
    // pairs and its only visit in the pass of block b1-2, is folded there and needs no pass of its own: a block is never
    // stored from a Q stage and then, with nothing else in between, loaded, re-scaled and stored again from the P stage.
    for (int g = crank; g < ng; g += CS) {
      const int b0 = min(nblk, g * gs), b1 = min(nblk, (g + 1) * gs);
      if (b0 >= b1) continue;
      norms_in(sm, gn, K, b0, b1);
      __syncthreads();
      if (b1 - b0 == 1) {
        p_pass<NPL, FULL>(sm, cx, X, K, N, b0, b1, b1, 1, false, dead, tol2, small2);
      } else {
        for (int p = b0; p < b1 - 1; ++p) {
          const int internal = ((b1 - 1 - p) & 1) ? 2 : ((p == b0) ? 1 : 0);
          p_pass<NPL, FULL>(sm, cx, X, K, N, p, p + 1, b1, internal, p == b1 - 2, dead, tol2, small2);
        }
      }
      __syncthreads();
      norms_out(sm, gn, K, b0, b1);
    }
    // ---- (b) pairs between groups: ng - 1 rounds, CS disjoint group pairs per round
    for (int rd = 0; rd + 1 < ng; ++rd) {
      if (tid == 0) bulk_wait_all();
      __syncthreads();
      fence_proxy_async_all();
      tcj::cluster_sync_all();
      fence_proxy_async_all();  // the bulk loads below (async proxy) read what other CTAs stored before the barrier
      int ga, gb;
      tcj::rr_pair(ng, rd, crank, ga, gb);
      const int a0 = min(nblk, ga * gs), a1 = min(nblk, (ga + 1) * gs);
      const int c0 = min(nblk, gb * gs), c1 = min(nblk, (gb + 1) * gs);
      if (a0 >= a1 || c0 >= c1) continue;
      norms_in(sm, gn, K, a0, a1);
      norms_in(sm, gn, K, c0, c1);
      __syncthreads();
      for (int p = a0; p < a1; ++p)
        p_pass<NPL, FULL>(sm, cx, X, K, N, p, c0, c1, 0, p == a1 - 1, dead, tol2, small2);
      __syncthreads();
      norms_out(sm, gn, K, a0, a1);
      norms_out(sm, gn, K, c0, c1);
    }
    // only "any pair left" and "any rotation made" matter: one count per warp, so that the packed sum cannot overflow
    // its 16-bit fields (a 512-row matrix has 130 816 pairs per sweep)
    if (lane == 0 && cx.nrot) atomicAdd(gcnt, ((cx.nrot & 0xffff) ? 1 : 0) | ((cx.nrot >> 16) ? (1 << 16) : 0));
    if (tid == 0) bulk_wait_all();
    __syncthreads();
    fence_proxy_async_all();
    tcj::cluster_sync_all();
    const int both = __reduce_max_sync(0xffffffffu, __ldcg(gcnt));
    tcj::cluster_sync_all();  // everybody has read the counter before the next sweep clears it
    if ((both & 0xffff) == 0) break;
    if (((unsigned)both >> 16) == 0) thr_off = true;  // a threshold sweep that rotated nothing: go to the final tolerance
  }
  if (crank == 0 && tid == 0) {
    if (sweep >= tcj::MAX_SWEEPS) atomicAdd(&d.flags[1], 1);
    atomicMax(&d.flags[2], sweep + 1);
    if (K >= 128) {
      atomicAdd(&d.flags[3], sweep + 1);
      atomicAdd(&d.flags[4], 1);
    }
  }
  // singular values = final row norms
  for (int r = gw; r < K; r += nwc) {
    const cplx *row = X + (size_t)r * N;
    double s = 0.0;
    for (int c = lane; c < N; c += 32) s += cabs2(__ldcg(reinterpret_cast<const double2 *>(row + c)));
    s = tcj::warp_sum(s);
    if (lane == 0) gn[r] = sqrt(s);
  }
}

// MAXNPL = 8: matrices up to 512 columns (chi_cap <= 256), 128 registers, one CTA per SM.
// MAXNPL = 4: up to 256 columns with 64 registers and 96 KB of shared memory, two CTAs = 32 warps per SM (TC_JACOBI=team).
// grid (nb * CS, chains), cluster (CS, 1, 1)
template <int MAXNPL>
__global__ void __launch_bounds__(NT, MAXNPL <= 4 ? 2 : 1) jacobi_team_kernel(TcDev d, LayerArgs a, int CS) {
  Bond b;
  // every CTA of a cluster sees the same bond, so the cluster leaves or stays as a whole
  if (!get_bond(d, a, centre_out(blockIdx.x / CS, a.nb), blockIdx.y, b)) return;
  const int N = __reduce_max_sync(0xffffffffu, b.N), K = __reduce_max_sync(0xffffffffu, b.M < b.N ? b.M : b.N);
  cplx *X = d.Xw + b.slot * d.slot_stride;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Sm sm;
  sm.sP = reinterpret_cast<cplx *>(smem_raw);
  sm.sQ = sm.sP + (size_t)BRW * d.n2;
  unsigned char *tail = smem_raw + (size_t)3 * BRW * d.n2 * sizeof(cplx);
  sm.nw = reinterpret_cast<double2 *>(tail);
  sm.barP = reinterpret_cast<uint64_t *>(tail + (size_t)d.n2 * sizeof(double2));
  sm.barQ = sm.barP + 1;
  sm.ver = reinterpret_cast<int *>(sm.barP + 4);
  sm.xch = reinterpret_cast<double2 *>(sm.ver + 2 * BRW * 2);
  __shared__ double red[32];
  if (threadIdx.x == 0) {
    mbar_init(&sm.barP[0], 1);
    mbar_init(&sm.barQ[0], 1);
    mbar_init(&sm.barQ[1], 1);
    fence_async_smem();
  }
  if (threadIdx.x < 2 * BRW * 2) sm.ver[threadIdx.x] = 0;
  __syncthreads();
  const int crank = CS > 1 ? (int)tcj::cluster_rank() : 0;
  // the stage rows are laid out with the matrix's own N, so N = 64 MAXNPL is the only FULL case
  if (N == 64 * MAXNPL)
    sweeps<MAXNPL, true>(d, b, X, K, N, CS, crank, sm, red);
  else if (2 * N <= 64 * MAXNPL && MAXNPL >= 8)
    sweeps<4, false>(d, b, X, K, N, CS, crank, sm, red);  // half-width instance for the narrower bonds of a wide context
  else
    sweeps<MAXNPL, false>(d, b, X, K, N, CS, crank, sm, red);
}
}  // namespace tct
