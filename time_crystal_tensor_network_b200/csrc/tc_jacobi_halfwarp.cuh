// tc_jacobi_halfwarp.cuh -- K2b for narrow matrices (N <= 128 columns: contexts with chi_cap <= 64, BASELINE configs 2
// and 3): the blocked one-sided Jacobi of tc_jacobi_blocked.cuh with a row on HALF a warp.
//
// At 128 columns a row is 2 KB; on a full warp that is 4 complex per lane and the pair visit is dominated by what does
// not shrink with the row: the butterfly reduction, the rotation set-up, the hand-over (45 of 93 FP64 instructions per
// lane).  Here a row lives on 16 lanes (8 complex per lane at 128 columns -- the register budget of the 256-column kernel),
// so one warp instruction stream carries TWO independent pairs: one butterfly level less, one set-up, one hand-over
// wait and one release per two pairs, and 32 pairs in flight per SM instead of 16.
//   CTA = 8 warps = 16 half-warps, row blocks of 16 (one cp.async.bulk per block), three stages (P, Q0, Q1) = 96 KB at
//   128 columns: two CTAs per SM.  Half-warp x = 2 warp + (lane >> 4) keeps row x of the P block in registers; in round s
//   of a visit it rotates (p_x, q_{(x+s) mod 16}).  Row q_j is passed x -> x-1 each round: from the odd to the even half of
//   a warp (program order, no synchronisation) and from the even half of warp w to the odd half of warp w-1 (version
//   counter, acquire / release) -- one wait and one release per warp and round.
// Same rotations (fast, scaled), thresholds, stopping rule and staging as the 16-warp kernel; only the pair ORDER inside a
// block visit is the same ring, on half-warps.  A half whose pair is below the threshold while the other half rotates
// applies the identity; a half without a pair (ragged last block) goes through the motions on row 0 and stores nothing.
#pragma once
#include "tc_common.cuh"
#include "tc_jacobi.cuh"
#include "tc_jacobi_blocked.cuh"

namespace tchw {
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

// warps per CTA, a template parameter: 8 = 16 half-warps, row blocks of 16, two CTAs per SM (ensembles: throughput);
// 16 = 32 half-warps, row blocks of 32, one CTA per SM (a single chain: all 32 pairs in flight work on ONE matrix)
constexpr int NW_MANY = 8, NW_ONE = 16;
constexpr int MAX_N = 128;
constexpr unsigned FULLM = 0xffffffffu;

// sums of a and b over each half-warp, in every lane of the half (packed butterfly: 5 shuffles + 4 adds)
__device__ __forceinline__ void half_sum2(double &a, double &b) {
  const bool hi = (threadIdx.x & 8) != 0;
  double k = hi ? b : a;
  k += __shfl_xor_sync(FULLM, hi ? a : b, 8);
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) k += __shfl_xor_sync(FULLM, k, o);
  const double other = __shfl_xor_sync(FULLM, k, 8);
  a = hi ? other : k;
  b = hi ? k : other;
}
__device__ __forceinline__ double half_sum(double v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(FULLM, v, o);
  return v;
}

// both rows in shared memory, one pair per half-warp (internal pairs of a block).  `active` false: this half has no pair
// in this round (xi / xj point at valid rows, nothing is stored).
template <int NPL, bool FULL>
__device__ __forceinline__ int pair_smem_h(bool active, cplx *xi, cplx *xj, int N, int hl, double2 *ni, double2 *nj,
                                           double dead, double tol2, double small2) {
  cplx u[NPL], v[NPL];
  double g0 = 0.0, g1 = 0.0, h0 = 0.0, h1 = 0.0;
#pragma unroll
  for (int e = 0; e < NPL; ++e) {
    const int c = hl + 16 * e;
    u[e] = (FULL || c < N) ? xi[c] : cmake(0.0, 0.0);
    v[e] = (FULL || c < N) ? xj[c] : cmake(0.0, 0.0);
    g0 = fma(u[e].x, v[e].x, g0);
    g1 = fma(u[e].y, v[e].y, g1);
    h0 = fma(u[e].y, v[e].x, h0);
    h1 = fma(-u[e].x, v[e].y, h1);
  }
  const double2 si = *ni, sj = *nj;
  double gr = g0 + g1, gi = h0 + h1;
  half_sum2(gr, gi);
  Rot r;
  int big;
  const bool rot = make_rot(active && si.x > dead && sj.x > dead, si.x, sj.x, si.y, sj.y, gr, gi, tol2, small2, r, big);
  if (!__any_sync(FULLM, rot)) return big;
  if (rot) {
#pragma unroll
    for (int e = 0; e < NPL; ++e) {
      const int c = hl + 16 * e;
      rot_apply(u[e], v[e], r);
      if (FULL || c < N) {
        xi[c] = u[e];
        xj[c] = v[e];
      }
    }
    if (hl == 0) {
      *ni = make_double2(r.ni, si.y * r.c2);
      *nj = make_double2(r.nj, sj.y * r.c2);
    }
  }
  return big | ((int)rot << 16);  // bits 0..15: pairs that keep the iteration going, bits 16..: rotations made
}

// row i of this half in registers (u, squared norm ai, squared scale wi), row j in shared memory
template <int NPL, bool FULL>
__device__ __forceinline__ int pair_reg_h(bool active, cplx (&u)[NPL], cplx *xj, int N, int hl, double &ai, double &wi,
                                          double2 *nj, double dead, double tol2, double small2) {
  cplx v[NPL];
  double g0 = 0.0, g1 = 0.0, h0 = 0.0, h1 = 0.0;
#pragma unroll
  for (int e = 0; e < NPL; ++e) {
    const int c = hl + 16 * e;
    v[e] = (FULL || c < N) ? xj[c] : cmake(0.0, 0.0);
    g0 = fma(u[e].x, v[e].x, g0);
    g1 = fma(u[e].y, v[e].y, g1);
    h0 = fma(u[e].y, v[e].x, h0);
    h1 = fma(-u[e].x, v[e].y, h1);
  }
  const double2 sj = *nj;
  double gr = g0 + g1, gi = h0 + h1;
  half_sum2(gr, gi);
  Rot r;
  int big;
  const bool rot = make_rot(active && ai > dead && sj.x > dead, ai, sj.x, wi, sj.y, gr, gi, tol2, small2, r, big);
  if (!__any_sync(FULLM, rot)) return big;
  if (rot) {
#pragma unroll
    for (int e = 0; e < NPL; ++e) {
      const int c = hl + 16 * e;
      rot_apply(u[e], v[e], r);
      if (FULL || c < N) xj[c] = v[e];
    }
    ai = r.ni;
    wi *= r.c2;
    if (hl == 0) *nj = make_double2(r.nj, sj.y * r.c2);
  }
  return big | ((int)rot << 16);
}

template <int NPL, bool FULL, int NW>
__device__ void sweeps(const TcDev &d, const Bond &b, cplx *X, int K, int N, int *s_rot, double *red) {
  constexpr int NT = NW * 32, BR = 2 * NW;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  cplx *const sP = reinterpret_cast<cplx *>(smem_raw);
  cplx *const sQ = sP + (size_t)BR * N;  // Q[buf] = sQ + buf * BR * N
  unsigned char *const tail = smem_raw + (size_t)3 * BR * d.n2 * sizeof(cplx);
  double2 *const s_nrm2 = reinterpret_cast<double2 *>(tail);  // per row: {true squared norm, squared scale w}
  uint64_t *const barP = reinterpret_cast<uint64_t *>(tail + d.n2 * sizeof(double2));
  uint64_t *const barQ = barP + 1;
  int *const s_ver = reinterpret_cast<int *>(barP + 4);
  const int tid = threadIdx.x, lane = tid & 31, warp = __reduce_max_sync(FULLM, tid >> 5);  // provably uniform
  const int hl = lane & 15, hf = lane >> 4;
  const int px = 2 * warp + hf;  // this half-warp's row of the P block
  const int nblk = (K + BR - 1) / BR;
  const double tol = 2.0 * sqrt((double)N) * 2.220446049250313e-16;
  const double tol2_final = tol * tol;
  bool thr_off = false;
  const uint32_t row_bytes = (uint32_t)N * sizeof(cplx);
  uint32_t phP = 0, phQ0 = 0, phQ1 = 0;
  int verBase0 = 0, verBase1 = 0;
  double dead = 0.0;
  int sweep = 0;
  for (; sweep < tcj::MAX_SWEEPS; ++sweep) {
    if (tid == 0) bulk_wait_all();  // all bulk stores of the previous sweep have landed before rows are re-read
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
    const double tol2 = (sweep < 6 && !thr_off) ? fmax(tol2_final, d.thr_sched[sweep]) : tol2_final;
    const double small2 = tol2 > tol2_final ? tol2_final : fmax(tol2_final, d.small_rel2);
    int nrot = 0;
    for (int p = 0; p < nblk; ++p) {
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
      }
      mbar_wait(barP, phP);
      phP ^= 1;
      // ---- internal pairs: circle method on the rows of a block, one pair per half-warp, BR/2 pairs per round = 4 warps
      // per block; blocks two at a time (p even: block p on warps 0..3, block p+1, prefetched into stage 0, on warps 4..7)
      bool q0_ready = false;
      if ((p & 1) == 0) {
        const bool pairUp = p + 1 < nblk;
        const int rowsN = pairUp ? min(BR, K - (p + 1) * BR) : 0;
        if (pairUp) {
          mbar_wait(&barQ[0], phQ0);
          phQ0 ^= 1;
          q0_ready = true;
        }
        const int half = warp >= NW / 2;
        const int k = 2 * (warp - half * (NW / 2)) + hf;  // pair of the round this half-warp takes
        const int rowsB = half ? rowsN : rowsP;
        cplx *blk = half ? sQ : sP;
        double2 *nb = s_nrm2 + (p + half) * BR;
        const int rmax = max(rowsP, rowsN) - 1;
        for (int r = 0; r < rmax; ++r) {
          const bool act = r < rowsB - 1 && k < rowsB / 2;
          int i = 0, j = 0;
          if (act) tcj::rr_pair(rowsB, r, k, i, j);
          if (__any_sync(FULLM, act))
            nrot += pair_smem_h<NPL, FULL>(act, blk + (size_t)i * N, blk + (size_t)j * N, N, hl, nb + i, nb + j, dead, tol2,
                                           small2);
          __syncthreads();
        }
      }
      // ---- row p_x into registers
      cplx u[NPL];
      const bool haveP = px < rowsP;
#pragma unroll
      for (int e = 0; e < NPL; ++e) {
        const int c = hl + 16 * e;
        u[e] = (haveP && (FULL || c < N)) ? sP[(size_t)px * N + c] : cmake(0.0, 0.0);
      }
      const double2 sP0 = haveP ? s_nrm2[p * BR + px] : make_double2(0.0, 1.0);
      double aP = sP0.x, wP = sP0.y;
      // ---- every later block streams through Q
      for (int q = p + 1; q < nblk; ++q) {
        const int buf = (q - p - 1) & 1;
        const int rowsQ = min(BR, K - q * BR);
        if (!(q == p + 1 && q0_ready)) {  // stage 0 of the first visit may already have been consumed above
          mbar_wait(&barQ[buf], buf ? phQ1 : phQ0);
          if (buf)
            phQ1 ^= 1;
          else
            phQ0 ^= 1;
        }
        cplx *Q = sQ + (size_t)buf * BR * N;
        const uint32_t vaddr = smem_u32(s_ver + buf * BR);
        const int base = buf ? verBase1 : verBase0;
        for (int s = 0; s < BR; ++s) {
          if (s == BR / 2 && tid == 0 && q + 1 < nblk) {  // prefetch of block q+1 in the middle of this visit
            bulk_wait_all();
            const int rq = min(BR, K - (q + 1) * BR);
            mbar_expect_tx(&barQ[buf ^ 1], rq * row_bytes);
            bulk_load(sQ + (size_t)(buf ^ 1) * BR * N, X + (size_t)(q + 1) * BR * N, rq * row_bytes, &barQ[buf ^ 1]);
          }
          const int jq = (px + s) & (BR - 1);
          // the odd half's row comes from the even half of warp w+1 (counter); the even half's row was the odd half's in
          // the previous round (same warp: program order)
          const int jwait = (2 * warp + 1 + s) & (BR - 1);
          if (s > 0) {
            int v;
            unsigned long long spins = 0;
            do {
              asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(vaddr + 4 * jwait) : "memory");
              if (++spins > (1ull << 24)) __trap();
            } while (__any_sync(FULLM, v < base + s));
          }
          const bool act = haveP && jq < rowsQ;
          const int jrow = act ? jq : 0;
          if (__any_sync(FULLM, act))
            nrot += pair_reg_h<NPL, FULL>(act, u, Q + (size_t)jrow * N, N, hl, aP, wP, s_nrm2 + q * BR + jrow, dead, tol2, small2);
          __syncwarp();
          // the even half's row goes on to the odd half of warp w-1: it has now been through s+1 half-warps ... but the
          // counter only ever needs the parity of the cross-warp steps: publish (base + s + 1) for the row of the even half
          if (lane == 0) {
            const int jrel = (2 * warp + s) & (BR - 1);
            asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(vaddr + 4 * jrel), "r"(base + s + 1) : "memory");
          }
        }
        if (buf)
          verBase1 = base + BR;
        else
          verBase0 = base + BR;
        fence_async_smem();
        __syncthreads();
        if (tid == 0) bulk_store(X + (size_t)q * BR * N, Q, rowsQ * row_bytes);
      }
      // ---- block p back to global, the rows' scales folded into their elements (once per sweep)
      if (haveP) {
        const double sc = sqrt(wP);
        if (hl == 0) s_nrm2[p * BR + px] = make_double2(aP, 1.0);
#pragma unroll
        for (int e = 0; e < NPL; ++e) {
          const int c = hl + 16 * e;
          if (FULL || c < N) sP[(size_t)px * N + c] = cscale(u[e], sc);
        }
      }
      fence_async_smem();
      __syncthreads();
      if (tid == 0) bulk_store(gP, sP, rowsP * row_bytes);
    }
    if (hl == 0 && nrot) atomicAdd(s_rot, nrot);
    __syncthreads();
    const int both = __reduce_max_sync(FULLM, *s_rot);
    __syncthreads();
    if ((both & 0xffff) == 0) break;
    if (((unsigned)both >> 16) == 0) thr_off = true;  // a threshold sweep that rotated nothing: go to the final tolerance
  }
  if (tid == 0) {
    bulk_wait_all();
    if (sweep >= tcj::MAX_SWEEPS) atomicAdd(&d.flags[1], 1);
    atomicMax(&d.flags[2], sweep + 1);
    if (K >= 128) {
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

__host__ __device__ inline size_t smem_bytes(int n2, int nw) {
  return (size_t)3 * 2 * nw * n2 * sizeof(cplx) + (size_t)n2 * sizeof(double2) + 64 + 2 * 2 * nw * sizeof(int);
}

template <int NW>
__global__ void __launch_bounds__(NW * 32, NW == 8 ? 2 : 1) jacobi_halfwarp_kernel(TcDev d, LayerArgs a) {
  constexpr int BR = 2 * NW;
  Bond b;
  // blockIdx.x = chain, blockIdx.y = rank of the bond in centre-out order (largest matrices first)
  if (!get_bond(d, a, centre_out(blockIdx.y, a.nb), blockIdx.x, b)) return;
  const int N = __reduce_max_sync(FULLM, b.N), K = __reduce_max_sync(FULLM, b.M < b.N ? b.M : b.N);
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
  __syncthreads();
  const int npl = (N + 15) / 16;
  if (N == 128)
    sweeps<8, true, NW>(d, b, X, K, N, &s_rot, red);
  else if (npl <= 1)
    sweeps<1, false, NW>(d, b, X, K, N, &s_rot, red);
  else if (npl <= 2)
    sweeps<2, false, NW>(d, b, X, K, N, &s_rot, red);
  else if (npl <= 4)
    sweeps<4, false, NW>(d, b, X, K, N, &s_rot, red);
  else
    sweeps<8, false, NW>(d, b, X, K, N, &s_rot, red);
}
}  // namespace tchw
