// tc_jacobi_blocked.cuh -- K2b fast path: one-sided Jacobi on the rows of the triangular factor with
// the working set staged in shared memory by TMA bulk copies and the stationary rows in registers.
//
// One CTA (16 warps) per matrix, K rows x N columns (N <= 256), rows grouped in blocks of 16 (one
// contiguous 16*N*16 B chunk in the row-major workspace, so a block moves with ONE cp.async.bulk).
// Sweep = for every block p: stage p, rotate its internal pairs, then keep row p_w in the registers of
// warp w while every later block q streams through a double-buffered shared-memory stage:
//   round s of visit (p, q): warp w rotates (p_w, q_{(w+s) mod 16}); q rows are read from and written
//   back to shared memory, p rows never leave registers; one __syncthreads per round.
// While visit (p, q) computes, the bulk load of q+1 and the bulk store of q-1 are in flight.
// Per pair: 4 KB LDS + 4 KB STS, ~100 DFMA per lane; L2 traffic per sweep ~ (K/16)^2/2 blocks.
#pragma once
#include "tc_common.cuh"
#include "tc_jacobi.cuh"

namespace tcb {
constexpr int NW = 16, NT = NW * 32, BR = 16;
constexpr int MAX_N = 256;

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
  while (!ok) {
    if (++spins > (1ull << 26)) __trap();  // a lost bulk copy must fail loudly, never hang the GPU
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
  }
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
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct Rot {
  double cs, sr, si, ni, nj;  // cos, s*e (complex), new squared norms
};

// Rotation that orthogonalises two rows with squared norms ai, aj and g = x_i . conj(x_j).
// The rotation ANGLE is computed in FP32 (SFU rsqrt/rcp instead of ~60 dependent FP64 instructions):
// any tau gives an exactly unitary transformation c [[1, -tau], [conj(tau), 1]] as long as
// c = 1/sqrt(1 + |tau|^2) is formed in FP64 from the tau that is actually applied; an FP32-accurate
// angle only leaves a residual ~1e-7 |g|, which the quadratically convergent sweeps absorb.
// The norm updates use the exact identities for the applied tau.
__device__ __forceinline__ bool make_rot(double ai, double aj, double gr, double gi, double tol2, Rot &r,
                                         int rot64) {
  const double g2 = gr * gr + gi * gi;
  if (!(g2 > tol2 * ai * aj)) return false;
  if (rot64) {  // all-FP64 angle (A/B switch TC_ROT64=1)
    const double rg = rsqrt(g2), ga = g2 * rg;
    const double dd = aj - ai;
    const double h = sqrt(fma(dd, dd, 4.0 * g2));
    const double t = copysign(2.0 * ga / (fabs(dd) + h), dd);
    r.cs = rsqrt(fma(t, t, 1.0));
    const double sn = r.cs * t;
    r.sr = sn * gr * rg;
    r.si = sn * gi * rg;
    r.ni = ai - t * ga;
    r.nj = aj + t * ga;
    return true;
  }
  // scale by an exact power of two so that max(ai, aj) is O(1) in float
  const double m = fmax(ai, aj);
  const int hi = __double2hiint(m);
  const double sc = __hiloint2double((2046 << 20) - (hi & 0x7ff00000), 0);  // 2^-exponent(m)
  const float fd = (float)((aj - ai) * sc);
  const float fr = (float)(gr * sc), fi = (float)(gi * sc);
  const float mx = fmaxf(fabsf(fr), fabsf(fi)), mn = fminf(fabsf(fr), fabsf(fi));
  const float q = __fdividef(mn, mx);
  const float iga = __frcp_rn(mx * sqrtf(fmaf(q, q, 1.0f)));  // 1 / |g| (scaled)
  const float az = 0.5f * fabsf(fd) * iga;                     // |zeta| = |aj - ai| / (2 |g|)
  float t = az > 1e15f ? __fdividef(0.5f, az) : __frcp_rn(az + sqrtf(fmaf(az, az, 1.0f)));
  t = copysignf(t, fd);
  const double tr = (double)(t * (fr * iga)), ti = (double)(t * (fi * iga));  // tau = t (g / |g|): unit phase first, t alone can be 1e-30
  const double t2 = fma(tr, tr, ti * ti);
  r.cs = rsqrt(1.0 + t2);
  r.sr = r.cs * tr;
  r.si = r.cs * ti;
  const double c2 = r.cs * r.cs;
  const double cross = 2.0 * fma(tr, gr, ti * gi);  // 2 Re(conj(tau) g)
  r.ni = c2 * (ai - cross + t2 * aj);
  r.nj = c2 * (aj + cross + t2 * ai);
  return true;
}

__device__ __forceinline__ void rot_apply(cplx &u, cplx &v, const Rot &r) {
  // u' = c u - (s e) v ;  v' = conj(s e) u + c v
  cplx un, vn;
  un.x = fma(r.cs, u.x, fma(-r.sr, v.x, r.si * v.y));
  un.y = fma(r.cs, u.y, -fma(r.sr, v.y, r.si * v.x));
  vn.x = fma(r.cs, v.x, fma(r.sr, u.x, r.si * u.y));
  vn.y = fma(r.cs, v.y, fma(r.sr, u.y, -r.si * u.x));
  u = un;
  v = vn;
}

__device__ __forceinline__ void warp_sum2(double &a, double &b) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
}

// both rows in shared memory (internal pairs of a block)
template <int NPL, bool FULL>
__device__ __forceinline__ int pair_smem(cplx *xi, cplx *xj, int N, int lane, double *ni, double *nj, double dead,
                                         double tol2, int rot64) {
  const double ai = *ni, aj = *nj;
  if (ai <= dead || aj <= dead) return 0;
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
  double gr = g0 + g1, gi = h0 + h1;
  warp_sum2(gr, gi);
  Rot r;
  if (!make_rot(ai, aj, gr, gi, tol2, r, rot64)) return 0;
  const int big = (gr * gr + gi * gi) > tcj::SMALL_REL2 * ai * aj;
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
    *ni = r.ni;
    *nj = r.nj;
  }
  return big;
}

// row i in registers (u), row j in shared memory
template <int NPL, bool FULL>
__device__ __forceinline__ int pair_reg(cplx (&u)[NPL], cplx *xj, int N, int lane, double *ni, double *nj, double dead,
                                        double tol2, int rot64) {
  const double ai = *ni, aj = *nj;
  if (ai <= dead || aj <= dead) return 0;
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
  double gr = g0 + g1, gi = h0 + h1;
  warp_sum2(gr, gi);
  Rot r;
  if (!make_rot(ai, aj, gr, gi, tol2, r, rot64)) return 0;
  const int big = (gr * gr + gi * gi) > tcj::SMALL_REL2 * ai * aj;
#pragma unroll
  for (int e = 0; e < NPL; ++e) {
    const int c = lane + 32 * e;
    rot_apply(u[e], v[e], r);
    if (FULL || c < N) xj[c] = v[e];
  }
  if (lane == 0) {
    *ni = r.ni;
    *nj = r.nj;
  }
  return big;
}

struct Stage {
  cplx *P, *Q[2];
  double *nrm2;
  int *ver;  // [2][BR] hand-over counters of the rows in Q[0], Q[1]
  uint64_t *barP, *barQ;  // barQ[2]
};

template <int NPL, bool FULL>
__device__ void sweeps(const TcDev &d, const Bond &b, cplx *X, int K, int N, int *s_rot, double *red) {
  // carve the stage out of dynamic shared memory here so that the compiler keeps the shared address space
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Stage st;
  st.P = reinterpret_cast<cplx *>(smem_raw);
  st.Q[0] = st.P + (size_t)BR * N;
  st.Q[1] = st.Q[0] + (size_t)BR * N;
  {
    unsigned char *tail = smem_raw + (size_t)3 * BR * d.n2 * sizeof(cplx);
    st.nrm2 = reinterpret_cast<double *>(tail);
    st.barP = reinterpret_cast<uint64_t *>(tail + d.n2 * sizeof(double));
    st.barQ = st.barP + 1;
    st.ver = reinterpret_cast<int *>(st.barP + 4);
  }
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nblk = (K + BR - 1) / BR;
  const double tol = 2.0 * sqrt((double)N) * 2.220446049250313e-16;
  const double tol2 = tol * tol;
  const uint32_t row_bytes = (uint32_t)N * sizeof(cplx);
  uint32_t phP = 0, phQ[2] = {0, 0};
  int verBase[2] = {0, 0};
  const int rot64 = d.rot64 & 1;
  const bool lockstep = (d.rot64 & 2) != 0;
  double dead = 0.0;
  int sweep = 0;
  for (; sweep < tcj::MAX_SWEEPS; ++sweep) {
    // all bulk stores of the previous sweep must have landed before rows are re-read
    if (tid == 0) bulk_wait_all();
    __syncthreads();
    for (int r = warp; r < K; r += NW) {
      const cplx *row = X + (size_t)r * N;
      double s = 0.0;
      for (int c = lane; c < N; c += 32) s += cabs2(__ldcg(reinterpret_cast<const double2 *>(row + c)));
      s = tcj::warp_sum(s);
      if (lane == 0) st.nrm2[r] = s;
    }
    if (tid == 0) *s_rot = 0;
    __syncthreads();
    if (sweep == 0) {
      double p = 0.0;
      for (int r = tid; r < K; r += NT) p += st.nrm2[r];
      dead = tcj::DEAD_REL2 * block_sum(p, red);
    }
    int nrot = 0;
    for (int p = 0; p < nblk; ++p) {
      const int rowsP = min(BR, K - p * BR);
      cplx *gP = X + (size_t)p * BR * N;
      if (tid == 0) {
        bulk_wait_all();  // the previous stores out of P / Q have finished reading shared memory
        mbar_expect_tx(st.barP, rowsP * row_bytes);
        bulk_load(st.P, gP, rowsP * row_bytes, st.barP);
        if (p + 1 < nblk) {
          const int rq = min(BR, K - (p + 1) * BR);
          mbar_expect_tx(&st.barQ[0], rq * row_bytes);
          bulk_load(st.Q[0], X + (size_t)(p + 1) * BR * N, rq * row_bytes, &st.barQ[0]);
        }
      }
      mbar_wait(st.barP, phP);
      phP ^= 1;
      // ---- internal pairs of block p: circle method on rowsP (even) rows, warps 0 .. rowsP/2-1
      for (int r = 0; r < rowsP - 1; ++r) {
        if (warp < rowsP / 2) {
          int i, j;
          tcj::rr_pair(rowsP, r, warp, i, j);
          nrot += pair_smem<NPL, FULL>(st.P + (size_t)i * N, st.P + (size_t)j * N, N, lane, st.nrm2 + p * BR + i,
                                 st.nrm2 + p * BR + j, dead, tol2, rot64);
        }
        __syncthreads();
      }
      // ---- row p_w into registers
      cplx u[NPL];
      const bool haveP = warp < rowsP;
#pragma unroll
      for (int e = 0; e < NPL; ++e) {
        const int c = lane + 32 * e;
        u[e] = (haveP && (FULL || c < N)) ? st.P[(size_t)warp * N + c] : cmake(0.0, 0.0);
      }
      // ---- every later block streams through Q
      for (int q = p + 1; q < nblk; ++q) {
        const int buf = (q - p - 1) & 1;
        const int rowsQ = min(BR, K - q * BR);
        if (tid == 0 && q + 1 < nblk) {
          bulk_wait_all();  // store of block q-1 (out of Q[buf^1]) complete
          const int rq = min(BR, K - (q + 1) * BR);
          mbar_expect_tx(&st.barQ[buf ^ 1], rq * row_bytes);
          bulk_load(st.Q[buf ^ 1], X + (size_t)(q + 1) * BR * N, rq * row_bytes, &st.barQ[buf ^ 1]);
        }
        mbar_wait(&st.barQ[buf], phQ[buf]);
        phQ[buf] ^= 1;
        cplx *Q = st.Q[buf];
        if (lockstep) {
          for (int s = 0; s < BR; ++s) {
            const int jq = (warp + s) & (BR - 1);
            if (haveP && jq < rowsQ)
              nrot += pair_reg<NPL, FULL>(u, Q + (size_t)jq * N, N, lane, st.nrm2 + p * BR + warp,
                                          st.nrm2 + q * BR + jq, dead, tol2, rot64);
            __syncthreads();
          }
        } else {
          // point-to-point hand-over: row jq carries a version counter; round s of this visit may touch it
          // once rounds 0..s-1 have released it (the previous holder is warp w+1).  No CTA-wide barrier
          // inside the visit, so the warps drift apart and one warp's scalar rotation set-up overlaps the
          // FP64-heavy dot / rotate phases of the others.
          const uint32_t vaddr = smem_u32(st.ver + buf * BR);
          const int base = verBase[buf];
          for (int s = 0; s < BR; ++s) {
            const int jq = (warp + s) & (BR - 1);
            if (s > 0) {
              int v;
              unsigned long long spins = 0;
              do {
                asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(vaddr + 4 * jq) : "memory");
                if (++spins > (1ull << 24)) __trap();
              } while (v < base + s);
            }
            if (haveP && jq < rowsQ)
              nrot += pair_reg<NPL, FULL>(u, Q + (size_t)jq * N, N, lane, st.nrm2 + p * BR + warp,
                                          st.nrm2 + q * BR + jq, dead, tol2, rot64);
            __syncwarp();
            if (lane == 0)
              asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(vaddr + 4 * jq), "r"(base + s + 1) : "memory");
          }
          verBase[buf] = base + BR;
        }
        fence_async_smem();
        __syncthreads();
        if (tid == 0) bulk_store(X + (size_t)q * BR * N, Q, rowsQ * row_bytes);
      }
      // ---- block p back to global
      if (haveP) {
#pragma unroll
        for (int e = 0; e < NPL; ++e) {
          const int c = lane + 32 * e;
          if (FULL || c < N) st.P[(size_t)warp * N + c] = u[e];
        }
      }
      fence_async_smem();
      __syncthreads();
      if (tid == 0) bulk_store(gP, st.P, rowsP * row_bytes);
    }
    if (lane == 0 && nrot) atomicAdd(s_rot, nrot);
    __syncthreads();
    const int tot = *s_rot;
    __syncthreads();
    if (tot == 0) break;
  }
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

// dynamic smem: 3 * BR * n2 cplx (P, Q0, Q1) + n2 doubles + 3 mbarriers (n2 <= MAX_N)
__global__ void __launch_bounds__(NT, 1) jacobi_blocked_kernel(TcDev d, LayerArgs a) {
  Bond b;
  if (!get_bond(d, a, blockIdx.x, blockIdx.y, b)) return;
  const int N = b.N, K = b.M < b.N ? b.M : b.N;
  cplx *X = d.Xw + b.slot * d.slot_stride;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + (size_t)3 * BR * d.n2 * sizeof(cplx) + d.n2 * sizeof(double));
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
}  // namespace tcb
