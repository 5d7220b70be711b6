// tc_jacobi.cuh -- batched truncated SVD of the two-site tensor theta (complex FP64) by one-sided
// Jacobi (Hestenes) on the ROWS of theta, one CTA per matrix, followed by sort + truncation +
// renormalisation in-kernel.
//
// Only right singular vectors are needed by the inverse-free TEBD update (B_{i+1} = V_k^H,
// B_i = C V_k), so the rotations act from the left:  J^H theta = Sigma V^H.  When the
// rows of theta are mutually orthogonal, row k IS sigma_k v_k^H: no rotation accumulation, no U,
// and the normalised rows drop straight into the (k, p1, chi_r) layout of site i+1.  One-sided
// Jacobi keeps high relative accuracy for the small singular values the reference retains
// (sigma > 1e-13, kicked_ising.py:186 -> TeNPy apply_local_op default cutoff).
//
// Preconditioner: theta's columns (p1, b) carry the Schmidt values S_{i+2}[b] of the right bond
// (site i+1 is in 'B' form), so in the interleaved order col' = 2 b + p1 they are graded like a
// pivoted QR would arrange them.  A Householder QR from the left, theta P = Q R with Q discarded,
// leaves R = Q^H theta P with the same right singular vectors; Jacobi on the rows of the triangular
// factor converges in 6-9 sweeps instead of 16-25 on theta itself (Drmac-Veselic preconditioning;
// sweep counts measured on TEBD matrices).
#pragma once
#include "tc_common.cuh"

namespace tcj {
constexpr int NT = 256;
constexpr int MAX_SWEEPS = 48;
// A sweep whose rotations all had relative off-diagonal |g|/sqrt(a_i a_j) below 1e-8 ends the
// iteration: by quadratic convergence what is left is below 1e-16 (checked on TEBD matrices: the final
// orthogonality is identical to running the extra verification sweep).
constexpr double SMALL_REL2 = 1e-16;
constexpr double DEAD_REL2 = 1e-30;  // rows with |x|^2 < DEAD_REL2 * |theta|_F^2 are numerically zero

// round-robin tournament (circle method): pair k of round r among M players (M even)
__device__ __forceinline__ void rr_pair(int M, int r, int k, int &i, int &j) {
  const int m1 = M - 1;
  if (k == 0) {
    i = m1;
    j = r;
  } else {
    i = r + k;
    if (i >= m1) i -= m1;
    j = r - k;
    if (j < 0) j += m1;
  }
  if (i > j) {
    int t = i;
    i = j;
    j = t;
  }
}

__device__ __forceinline__ double warp_sum(double v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}


// zlarfg on (alpha, |x|^2): H = I - tau [1;v][1;v]^H with H^H [alpha; x] = [beta; 0], beta real;
// `scale` = 1/(alpha - beta) turns the tail x into v.
__device__ __forceinline__ void larfg(cplx alpha, double xnorm2, double &beta, cplx &tau, cplx &scale) {
  if (xnorm2 == 0.0 && alpha.y == 0.0) {
    beta = alpha.x;
    tau = cmake(0.0, 0.0);
    scale = cmake(0.0, 0.0);
    return;
  }
  const double nrm = sqrt(alpha.x * alpha.x + alpha.y * alpha.y + xnorm2);
  beta = signbit(alpha.x) ? nrm : -nrm;
  tau = cmake((beta - alpha.x) / beta, -alpha.y / beta);
  scale = crecip(cmake(alpha.x - beta, alpha.y));
}

// ------------------------------------------------------------------------------------------------
// K2a: in-place Householder QR of X (M x N, row-major, columns already in interleaved order):
// X <- R (upper trapezoidal, K = min(M, N) rows; rows K.. and everything below the diagonal are zero).
// The reflectors are applied on the fly and discarded.  One CTA per matrix, thread per column for the
// trailing update (coalesced across the row).  dynamic smem: n2 cplx + 32 doubles
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) qr_kernel(TcDev d, LayerArgs a) {
  Bond b;
  if (!get_bond(d, a, blockIdx.x, blockIdx.y, b)) return;
  const int M = b.M, N = b.N;
  cplx *X = d.Xw + b.slot * d.slot_stride;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx *v = reinterpret_cast<cplx *>(smem_raw);
  double *red = reinterpret_cast<double *>(v + d.n2);
  __shared__ cplx s_tau, s_scale;
  __shared__ double s_beta;
  const int tid = threadIdx.x;
  const int steps = (M - 1) < N ? (M - 1) : N;
  for (int k = 0; k < steps; ++k) {
    const int len = M - k;
    double part = 0.0;
    for (int r = tid; r < len; r += NT) {
      const cplx x = X[(size_t)(k + r) * N + k];
      v[r] = x;
      if (r > 0) part += cabs2(x);
    }
    const double xn2 = block_sum(part, red);
    if (tid == 0) {
      double beta;
      cplx tau, sc;
      larfg(v[0], xn2, beta, tau, sc);
      s_beta = beta;
      s_tau = tau;
      s_scale = sc;
    }
    __syncthreads();
    const cplx tau = s_tau;
    if (tau.x == 0.0 && tau.y == 0.0) {
      __syncthreads();
      continue;
    }
    {
      const cplx sc = s_scale;
      const double beta = s_beta;
      for (int r = tid; r < len; r += NT) {
        v[r] = (r == 0) ? cmake(1.0, 0.0) : cmul(v[r], sc);
        X[(size_t)(k + r) * N + k] = (r == 0) ? cmake(beta, 0.0) : cmake(0.0, 0.0);
      }
    }
    __syncthreads();
    const cplx tauc = cconj(tau);
    for (int j = k + 1 + tid; j < N; j += NT) {
      cplx *col = X + (size_t)k * N + j;
      cplx acc0 = cmake(0.0, 0.0), acc1 = cmake(0.0, 0.0);
      int r = 0;
      for (; r + 1 < len; r += 2) {
        cfmac(acc0, v[r], col[(size_t)r * N]);
        cfmac(acc1, v[r + 1], col[(size_t)(r + 1) * N]);
      }
      if (r < len) cfmac(acc0, v[r], col[(size_t)r * N]);
      const cplx w = cmul(tauc, cadd(acc0, acc1));
      for (r = 0; r < len; ++r) {
        cplx x = col[(size_t)r * N];
        const cplx m = cmul(v[r], w);
        col[(size_t)r * N] = csub(x, m);
      }
    }
    __syncthreads();
  }
}


// ------------------------------------------------------------------------------------------------
// K2a fast path (M <= 256 rows): blocked Householder QR, panels of QB columns.
//   panel:    one row per thread, the panel row in registers; per column one batched block reduction
//             (|x|^2 and the v^H a products for the remaining panel columns); reflectors V to smem
//   T:        compact WY factor, H_0 ... H_{QB-1} = I - V T V^H (larft forward/columnwise)
//   trailing: warp per tile of 8 columns on the FP64 tensor pipe (DMMA): W = V^H A (one pass), W <- T^H W,
//             A -= V W (second pass): two passes over the trailing matrix per PANEL instead of per column.
// static smem: V 32 KB + reduction scratch
// ------------------------------------------------------------------------------------------------
// FP64 tensor-pipe MMA (8 x 4) x (4 x 8): A fragment element [lane >> 2][lane & 3], B fragment [lane & 3][lane >> 2],
// accumulator [lane >> 2][2 (lane & 3), + 1]
__device__ __forceinline__ void tcg_dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned cluster_rank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

constexpr int QNT = 256, QMAXM = 256;    // fast-path instance: M <= 256 rows, 256 threads
constexpr int QNTN = 128, QMAXMN = 128;  // narrow instance: M <= 128 rows, 128 threads (ensembles with chi_cap <= 64); same bits
constexpr int QNTW = 512, QMAXMW = 512;  // wide instance (chi_cap <= 256): M <= 512 rows, 512 threads
#ifndef TC_QR_U1
#define TC_QR_U1 4  // rows-of-4 steps of pass 1 of the trailing update in flight per warp
#endif
#ifndef TC_QR_U2
#define TC_QR_U2 2  // rows-of-8 steps of pass 2 in flight per warp
#endif
#ifndef TC_QB
#define TC_QB 8
#endif
constexpr int QRU1 = TC_QR_U1, QRU2 = TC_QR_U2;
constexpr int QBDEF = TC_QB;  // panel width (8 or 16 columns).  16 halves the passes over the trailing matrix but was measured
                              // slower on the B200 (QR share of the step 6.9 % -> 10.0 %, r02d): the kernel is bound by the panel
                              // factorisation (block reductions, V^H V, T) and not by the trailing update

// block-wide sums of NV doubles per thread; result in out[0..NV) (all threads), scratch [QT/32][NV]
template <int NV, int QT>
__device__ __forceinline__ void block_sum_vec(double (&v)[NV], double *scratch, int nv_used) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k)
    if (k < nv_used)
      for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  __syncthreads();
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < NV; ++k)
      if (k < nv_used) scratch[warp * NV + k] = v[k];
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k)
    if (k < nv_used) {
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < QT / 32; ++w) t += scratch[w * NV + k];
      v[k] = t;
    }
}

// QT threads = max rows, QB = panel width (multiple of 8); dynamic smem: V panel, QT * QB cplx.
// CS > 1: a thread-block cluster of CS CTAs shares one matrix (launches with too few matrices to fill the GPU: a single
// chain, BASELINE config 4).  Every CTA factorises the panel redundantly (same arithmetic, same bits; only rank 0 writes
// the R part back) and the 8-column tiles of the trailing matrix are dealt round-robin BY ABSOLUTE COLUMN BLOCK to the
// CS * QT / 32 warps of the cluster, so a tile is read and written by the same SM for the whole factorisation (L1 is not
// coherent between SMs); the only data that crosses SMs is the next panel, read with ld.cg after a cluster barrier.
template <int QT, int QB>
__global__ void __launch_bounds__(QT) qr_blocked_kernel(TcDev d, LayerArgs a, int CS) {
  Bond b;
  // blockIdx.x / CS = chain, blockIdx.y = rank of the bond in centre-out order (largest matrices first); every CTA of a
  // cluster sees the same bond, so the cluster leaves or stays as a whole
  if (!get_bond(d, a, centre_out(blockIdx.y, a.nb), blockIdx.x / CS, b)) return;
  const int crank = CS > 1 ? (int)cluster_rank() : 0;
  const int M = b.M, N = b.N;
  cplx *X = d.Xw + b.slot * d.slot_stride;
  extern __shared__ __align__(16) unsigned char qr_smem_raw[];
  cplx *const V = reinterpret_cast<cplx *>(qr_smem_raw);  // [QT][QB]
  __shared__ __align__(16) cplx Tm[QB * QB];
  __shared__ double scratch[(QT / 32) * 2 * QB];
  __shared__ cplx s_alpha;
  __shared__ cplx s_tau[QB];
  const int tid = threadIdx.x;
  const int steps = (M - 1) < N ? (M - 1) : N;  // columns that have something to annihilate
  for (int k0 = 0; k0 < steps; k0 += QB) {
    const int nbw = (steps - k0) < QB ? (steps - k0) : QB;  // reflectors in this panel
    const int pw = (N - k0) < QB ? (N - k0) : QB;           // panel columns present
    const int rows = M - k0;                                // panel rows, one per thread
    const bool have = tid < rows;
    cplx p[QB];
#pragma unroll
    for (int j = 0; j < QB; ++j)
      p[j] = (have && j < pw) ? __ldcg(reinterpret_cast<const double2 *>(X + (size_t)(k0 + tid) * N + k0 + j)) : cmake(0.0, 0.0);
#pragma unroll
    for (int j = 0; j < QB; ++j) {
      cplx tauj = cmake(0.0, 0.0);
      if (j < nbw) {
        // |x|^2 below the diagonal of column j
        double red[2 * QB];
        red[0] = (have && tid > j) ? cabs2(p[j]) : 0.0;
        if (tid == j) s_alpha = p[j];
        block_sum_vec<2 * QB, QT>(red, scratch, 1);
        double beta;
        cplx sc;
        larfg(s_alpha, red[0], beta, tauj, sc);
        const cplx v = (tid == j) ? cmake(1.0, 0.0) : ((have && tid > j) ? cmul(p[j], sc) : cmake(0.0, 0.0));
        V[tid * QB + j] = v;
        if (tauj.x != 0.0 || tauj.y != 0.0) {
          // w_c = conj(tau) sum_r conj(v_r) p_r[c] for the remaining panel columns
#pragma unroll
          for (int c = 0; c < QB; ++c) {
            cplx t = cmake(0.0, 0.0);
            if (c > j) cfmac(t, v, p[c]);
            red[2 * c] = t.x;
            red[2 * c + 1] = t.y;
          }
          block_sum_vec<2 * QB, QT>(red, scratch, 2 * QB);
          const cplx tc = cconj(tauj);
#pragma unroll
          for (int c = 0; c < QB; ++c)
            if (c > j) {
              const cplx w = cmul(tc, cmake(red[2 * c], red[2 * c + 1]));
              const cplx m = cmul(v, w);
              p[c] = csub(p[c], m);
            }
          if (tid == j) p[j] = cmake(beta, 0.0);
          if (tid > j) p[j] = cmake(0.0, 0.0);
        } else {
          __syncthreads();  // keep the barrier count uniform with the branch above
          __syncthreads();
        }
      } else {
        V[tid * QB + j] = cmake(0.0, 0.0);
      }
      if (tid == 0) s_tau[j] = tauj;
    }
    // R part of the panel back to global
    if (have && crank == 0)
#pragma unroll
      for (int j = 0; j < QB; ++j)
        if (j < pw) X[(size_t)(k0 + tid) * N + k0 + j] = p[j];
    __syncthreads();
    // ---- T: T[j][j] = tau_j, T[0:j, j] = -tau_j T[0:j,0:j] (V[:,0:j]^H v_j)
    {
      const int warp = tid >> 5, lane = tid & 31;
      __shared__ cplx G[QB * QB];
      for (int pr = warp; pr < QB * QB; pr += QT / 32) {
        const int i = pr / QB, j = pr % QB;
        if (i < j) {
          cplx acc = cmake(0.0, 0.0);
          for (int r = lane; r < rows; r += 32) cfmac(acc, V[r * QB + i], V[r * QB + j]);
          acc.x = warp_sum(acc.x);
          acc.y = warp_sum(acc.y);
          if (lane == 0) G[pr] = acc;
        }
      }
      __syncthreads();
      if (warp == 0) {  // column after column, lane i computes row i of the column (rows are independent)
        for (int j = 0; j < QB; ++j) {
          const cplx tj = s_tau[j];
          if (lane < QB) {
            cplx t = cmake(0.0, 0.0);
            if (lane == j) t = tj;
            if (lane < j) {
              cplx acc = cmake(0.0, 0.0);
              for (int l = lane; l < j; ++l) cfma(acc, Tm[lane * QB + l], G[l * QB + j]);
              t = cmul(cmake(-tj.x, -tj.y), acc);
            }
            Tm[lane * QB + j] = t;
          }
          __syncwarp();
        }
      }
      __syncthreads();
    }
    // ---- trailing columns: A <- (I - V T^H V^H) A on the FP64 tensor pipe (mma.sync m8n8k4.f64 = DMMA; same FP64
    // peak as DFMA on sm_100a, but one warp instruction per 256 FMAs and no shared-memory broadcast per FMA).
    // A warp owns tiles of 8 trailing columns.  Pass 1, W = V^H A: M = QB reflectors (QB / 8 tiles), K = rows, N = 8
    // columns (A operand conj(V)^T from shared memory, B operand the trailing matrix straight from L2).  W2 = T^H W per
    // lane from a per-warp scratch, directly in the B-operand layout of pass 2, A -= V W2: M = 8 rows, K = QB
    // reflectors, N = 8 columns.
    {
      constexpr int MT = QB / 8, KC = QB / 4;
      __shared__ __align__(16) cplx Wsm[(QT / 32) * QB * 8];
      const int lane = tid & 31, warp = tid >> 5, fr = lane >> 2, fk = lane & 3;
      const int ncols = N - (k0 + QB);
      const int ntiles = (ncols + 7) / 8;
      cplx *At = X + (size_t)k0 * N + k0 + QB;  // trailing block: row r (relative to k0), column cc (relative to k0 + QB)
      cplx *Ws = Wsm + warp * QB * 8;
      // tile ct is the absolute 8-column block (k0 + QB) / 8 + ct: owner = block index modulo the warps of the cluster
      constexpr int NWQ = QT / 32;
      const int nwc = CS * NWQ, gw = crank * NWQ + warp;
      const int ct0 = ((gw - (k0 + QB) / 8) % nwc + nwc) % nwc;
      for (int ct = ct0; ct < ntiles; ct += nwc) {
        const int c0 = ct * 8;
        const bool cok = c0 + fr < ncols;
        double wre[MT][2], wim[MT][2];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) wre[mt][0] = wre[mt][1] = wim[mt][0] = wim[mt][1] = 0.0;
#pragma unroll QRU1
        for (int r0 = 0; r0 < rows; r0 += 4) {
          const int r = r0 + fk;
          const bool rok = r < rows;
          const cplx av = (cok && rok) ? At[(size_t)r * N + c0 + fr] : cmake(0.0, 0.0);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const cplx v = rok ? V[r * QB + mt * 8 + fr] : cmake(0.0, 0.0);
            tcg_dmma(wre[mt][0], wre[mt][1], v.x, av.x);  // conj(v) a = (vx ax + vy ay) + i (vx ay - vy ax)
            tcg_dmma(wre[mt][0], wre[mt][1], v.y, av.y);
            tcg_dmma(wim[mt][0], wim[mt][1], v.x, av.y);
            tcg_dmma(wim[mt][0], wim[mt][1], -v.y, av.x);
          }
        }
        // C layout: this lane holds W[j = 8 mt + fr][c = 2 fk, 2 fk + 1]
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          Ws[(mt * 8 + fr) * 8 + 2 * fk] = cmake(wre[mt][0], wim[mt][0]);
          Ws[(mt * 8 + fr) * 8 + 2 * fk + 1] = cmake(wre[mt][1], wim[mt][1]);
        }
        __syncwarp();
        cplx w2[KC];  // W2[i = 4 kc + fk][c = fr] = sum_{j <= i} conj(T[j][i]) W[j][c]
#pragma unroll
        for (int kc = 0; kc < KC; ++kc) {
          const int i = 4 * kc + fk;
          cplx acc = cmake(0.0, 0.0);
#pragma unroll
          for (int j = 0; j < QB; ++j)
            if (j <= i) cfmac(acc, Tm[j * QB + i], Ws[j * 8 + fr]);
          w2[kc] = acc;
        }
        __syncwarp();
        const int cc = c0 + 2 * fk;
#pragma unroll QRU2
        for (int r0 = 0; r0 < rows; r0 += 8) {
          const int r = r0 + fr;
          const bool rok = r < rows;
          cplx *ap = At + (size_t)r * N + cc;
          const cplx a0 = (rok && cc < ncols) ? ap[0] : cmake(0.0, 0.0);
          const cplx a1 = (rok && cc + 1 < ncols) ? ap[1] : cmake(0.0, 0.0);
          double cre[2] = {a0.x, a1.x}, cim[2] = {a0.y, a1.y};
#pragma unroll
          for (int kc = 0; kc < KC; ++kc) {
            const cplx v = rok ? V[r * QB + 4 * kc + fk] : cmake(0.0, 0.0);
            tcg_dmma(cre[0], cre[1], -v.x, w2[kc].x);  // a - v w2
            tcg_dmma(cre[0], cre[1], v.y, w2[kc].y);
            tcg_dmma(cim[0], cim[1], -v.x, w2[kc].y);
            tcg_dmma(cim[0], cim[1], -v.y, w2[kc].x);
          }
          if (rok && cc < ncols) ap[0] = cmake(cre[0], cim[0]);
          if (rok && cc + 1 < ncols) ap[1] = cmake(cre[1], cim[1]);
        }
      }
    }
    if (CS > 1)
      cluster_sync_all();  // release / acquire at cluster scope: the next panel's columns are visible to every CTA
    else
      __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// v1: warp per row pair, rows streamed from L1/L2.  dynamic smem: n2 doubles (row norms^2)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) jacobi_rows_kernel(TcDev d, LayerArgs a) {
  Bond b;
  if (!get_bond(d, a, blockIdx.x, blockIdx.y, b)) return;
  const int N = b.N, M = b.M < b.N ? b.M : b.N;  // rows of the triangular factor
  cplx *X = d.Xw + b.slot * d.slot_stride;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *nrm2 = reinterpret_cast<double *>(smem_raw);
  __shared__ double red[32];
  __shared__ int s_rot;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = NT / 32;
  const double tol = 2.0 * sqrt((double)N) * 2.220446049250313e-16;
  const double tol2 = tol * tol;

  double dead = 0.0;
  int sweep = 0;
  for (; sweep < MAX_SWEEPS; ++sweep) {
    // fresh row norms at the start of every sweep
    for (int r = warp; r < M; r += NW) {
      const cplx *row = X + (size_t)r * N;
      double s = 0.0;
      for (int c = lane; c < N; c += 32) s += cabs2(row[c]);
      s = warp_sum(s);
      if (lane == 0) nrm2[r] = s;
    }
    if (tid == 0) s_rot = 0;
    __syncthreads();
    if (sweep == 0) {
      double p = 0.0;
      for (int r = tid; r < M; r += NT) p += nrm2[r];
      dead = DEAD_REL2 * block_sum(p, red);
    }
    int nrot = 0;
    for (int r = 0; r < M - 1; ++r) {
      for (int k = warp; k < M / 2; k += NW) {
        int i, j;
        rr_pair(M, r, k, i, j);
        const double ai = nrm2[i], aj = nrm2[j];
        if (ai <= dead || aj <= dead) continue;
        cplx *xi = X + (size_t)i * N, *xj = X + (size_t)j * N;
        double gr = 0.0, gi = 0.0;  // g = sum x_i conj(x_j)
        for (int c = lane; c < N; c += 32) {
          const cplx u = xi[c], v = xj[c];
          gr = fma(u.x, v.x, gr);
          gr = fma(u.y, v.y, gr);
          gi = fma(u.y, v.x, gi);
          gi = fma(-u.x, v.y, gi);
        }
        gr = warp_sum(gr);
        gi = warp_sum(gi);
        const double g2 = gr * gr + gi * gi;
        if (!(g2 > tol2 * ai * aj)) continue;
        nrot += g2 > fmax(tol2, d.small_rel2) * ai * aj;
        const double ga = sqrt(g2);
        const double zeta = (aj - ai) / (2.0 * ga);
        const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double cs = 1.0 / sqrt(1.0 + t * t);
        const double sn = cs * t;
        const double er = gr / ga, ei = gi / ga;  // e = g / |g|
        const double sr = sn * er, si = sn * ei;  // s e
        // x_i' = c x_i - (s e) x_j ;  x_j' = conj(s e) x_i + c x_j
        for (int c = lane; c < N; c += 32) {
          const cplx u = xi[c], v = xj[c];
          cplx un, vn;
          un.x = cs * u.x - (sr * v.x - si * v.y);
          un.y = cs * u.y - (sr * v.y + si * v.x);
          vn.x = cs * v.x + (sr * u.x + si * u.y);
          vn.y = cs * v.y + (sr * u.y - si * u.x);
          xi[c] = un;
          xj[c] = vn;
        }
        if (lane == 0) {
          nrm2[i] = ai - t * ga;
          nrm2[j] = aj + t * ga;
        }
      }
      __syncthreads();
    }
    if (lane == 0 && nrot) atomicAdd(&s_rot, nrot);
    __syncthreads();
    const int tot = s_rot;
    __syncthreads();
    if (tot == 0) break;
  }
  if (tid == 0) {
    if (sweep >= MAX_SWEEPS) atomicAdd(&d.flags[1], 1);
    atomicMax(&d.flags[2], sweep + 1);
  }
  // singular values = final row norms
  double *w = d.ww + b.slot * d.n2;
  for (int r = warp; r < M; r += NW) {
    const cplx *row = X + (size_t)r * N;
    double s = 0.0;
    for (int c = lane; c < N; c += 32) s += cabs2(row[c]);
    s = warp_sum(s);
    if (lane == 0) w[r] = sqrt(s);
  }
}

// ------------------------------------------------------------------------------------------------
// v1c: the warp-per-pair kernel on a thread-block CLUSTER (matrices wider than the register-resident fast path,
// chi_cap > 128, where a context holds few chains: BASELINE config 4 is ONE chain of 63 bonds, i.e. ~31 matrices of
// 512 x 512 per layer -- one CTA per matrix would leave 4/5 of the SMs idle).  CS CTAs share a matrix: the M/2
// disjoint pairs of a tournament round are dealt to the CS * 8 warps of the cluster, rows and row norms live in
// global memory / L2 (ld.cg / st.cg: L1 is not coherent between the SMs of a cluster), one cluster barrier
// (release / acquire, ~0.2 us) per round.  Scratch: the slot's `ww` row (squared norms, overwritten by the singular
// values at the end) and its `knew` entry (rotation counter; finalize_kernel overwrites it).
// ------------------------------------------------------------------------------------------------
constexpr int WNPL = 16;  // complex elements per lane and row in the cluster kernel: matrices up to 512 columns

__global__ void __launch_bounds__(NT) jacobi_rows_cluster_kernel(TcDev d, LayerArgs a, int CS) {
  Bond b;
  // every CTA of a cluster sees the same bond, so the cluster leaves or stays as a whole
  if (!get_bond(d, a, centre_out(blockIdx.x / CS, a.nb), blockIdx.y, b)) return;
  const int N = b.N, M = b.M < b.N ? b.M : b.N;  // rows of the triangular factor
  cplx *X = d.Xw + b.slot * d.slot_stride;
  double *nrm2 = d.ww + b.slot * d.n2;
  int *cnt = d.knew + b.slot;
  __shared__ double red[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = NT / 32;
  const int rank = (int)cluster_rank();
  const int gw = rank * NW + warp, nw = CS * NW;  // this warp among the warps of the cluster
  const double tol = 2.0 * sqrt((double)N) * 2.220446049250313e-16;
  const double tol2_final = tol * tol;
  double dead = 0.0;
  int sweep = 0;
  for (; sweep < MAX_SWEEPS; ++sweep) {
    for (int r = gw; r < M; r += nw) {
      const double2 *row = reinterpret_cast<const double2 *>(X + (size_t)r * N);
      double s = 0.0;
      for (int c = lane; c < N; c += 32) s += cabs2(__ldcg(row + c));
      s = warp_sum(s);
      if (lane == 0) __stcg(nrm2 + r, s);
    }
    if (gw == 0 && lane == 0) __stcg(cnt, 0);
    cluster_sync_all();
    if (sweep == 0) {
      double p = 0.0;
      for (int r = tid; r < M; r += NT) p += __ldcg(nrm2 + r);
      dead = DEAD_REL2 * block_sum(p, red);
    }
    // threshold sweeps and their stopping rule as in tc_jacobi_blocked.cuh (here a skipped rotation also saves the
    // write-back of the two rows, half of the L2 traffic of the pair)
    const double tol2 = sweep < 6 ? fmax(tol2_final, d.thr_sched[sweep]) : tol2_final;
    const double small2 = tol2 > tol2_final ? tol2_final : fmax(tol2_final, d.small_rel2);
    int nrot = 0;
    for (int r = 0; r < M - 1; ++r) {
      for (int k = gw; k < M / 2; k += nw) {
        int i, j;
        rr_pair(M, r, k, i, j);
        const double ai = __ldcg(nrm2 + i), aj = __ldcg(nrm2 + j);
        if (ai <= dead || aj <= dead) continue;
        double2 *xi = reinterpret_cast<double2 *>(X + (size_t)i * N), *xj = reinterpret_cast<double2 *>(X + (size_t)j * N);
        // both rows into registers with every load in flight at once (N <= 32 WNPL = 512 columns: 16 complex per lane
        // and row); they stay there between the dot product and the rotation, so a pair costs one read and one
        // write of its rows in L2 instead of two reads and a write
        cplx u[WNPL], v[WNPL];
#pragma unroll
        for (int e = 0; e < WNPL; ++e) {
          const int c = lane + 32 * e;
          u[e] = c < N ? __ldcg(xi + c) : cmake(0.0, 0.0);
          v[e] = c < N ? __ldcg(xj + c) : cmake(0.0, 0.0);
        }
        double g0 = 0.0, g1 = 0.0, h0 = 0.0, h1 = 0.0;  // g = sum x_i conj(x_j)
#pragma unroll
        for (int e = 0; e < WNPL; ++e) {
          g0 = fma(u[e].x, v[e].x, g0);
          g1 = fma(u[e].y, v[e].y, g1);
          h0 = fma(u[e].y, v[e].x, h0);
          h1 = fma(-u[e].x, v[e].y, h1);
        }
        double gr = warp_sum(g0 + g1);
        double gi = warp_sum(h0 + h1);
        const double g2 = gr * gr + gi * gi;
        nrot += g2 > small2 * ai * aj;
        if (!(g2 > tol2 * ai * aj)) continue;
        const double ga = sqrt(g2);
        const double zeta = (aj - ai) / (2.0 * ga);
        const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double cs = 1.0 / sqrt(1.0 + t * t);
        const double sn = cs * t;
        const double er = gr / ga, ei = gi / ga;  // e = g / |g|
        const double sr = sn * er, si = sn * ei;  // s e
        // x_i' = c x_i - (s e) x_j ;  x_j' = conj(s e) x_i + c x_j
#pragma unroll
        for (int e = 0; e < WNPL; ++e) {
          const int c = lane + 32 * e;
          cplx un, vn;
          un.x = cs * u[e].x - (sr * v[e].x - si * v[e].y);
          un.y = cs * u[e].y - (sr * v[e].y + si * v[e].x);
          vn.x = cs * v[e].x + (sr * u[e].x + si * u[e].y);
          vn.y = cs * v[e].y + (sr * u[e].y - si * u[e].x);
          if (c < N) {
            __stcg(xi + c, un);
            __stcg(xj + c, vn);
          }
        }
        if (lane == 0) {
          __stcg(nrm2 + i, ai - t * ga);
          __stcg(nrm2 + j, aj + t * ga);
        }
      }
      cluster_sync_all();
    }
    if (lane == 0 && nrot) atomicAdd(cnt, nrot);
    cluster_sync_all();
    const int tot = __ldcg(cnt);
    cluster_sync_all();  // everybody has read the counter before the next sweep clears it
    if (tot == 0) break;
  }
  if (gw == 0 && lane == 0) {
    if (sweep >= MAX_SWEEPS) atomicAdd(&d.flags[1], 1);
    atomicMax(&d.flags[2], sweep + 1);
  }
  // singular values = final row norms (into the same ww row the squared norms lived in)
  for (int r = gw; r < M; r += nw) {
    const double2 *row = reinterpret_cast<const double2 *>(X + (size_t)r * N);
    double s = 0.0;
    for (int c = lane; c < N; c += 32) s += cabs2(__ldcg(row + c));
    s = warp_sum(s);
    if (lane == 0) nrm2[r] = sqrt(s);
  }
}

// ------------------------------------------------------------------------------------------------
// sort + truncation + renormalisation + write-back of S_{i+1}, chi_{i+1} and B_{i+1} = V_k^H.
// Truncation rules: mode 0 = TeNPy apply_local_op default (keep sigma > cutoff, absolute);
// mode 1 = TeNPy truncate() (chi_max, svd_min, trunc_cut on the normalised spectrum); SURVEY A.2.3/A.2.5.
// dynamic smem: n2 doubles (sorted values)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) finalize_kernel(TcDev d, LayerArgs a) {
  Bond b;
  if (!get_bond(d, a, blockIdx.x, blockIdx.y, b)) return;
  const int N = b.N, M = b.M < b.N ? b.M : b.N;  // rows of the triangular factor
  const int chiR = b.chiR;
  const cplx *X = d.Xw + b.slot * d.slot_stride;
  const double *w = d.ww + b.slot * d.n2;
  int *perm = d.perm + b.slot * d.n2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *sorted = reinterpret_cast<double *>(smem_raw);
  __shared__ int s_k;
  __shared__ double s_renorm;
  const int tid = threadIdx.x;

  // descending rank sort (stable)
  for (int t = tid; t < M; t += NT) {
    const double wt = w[t];
    int rank = 0;
    for (int j = 0; j < M; ++j) {
      const double wj = w[j];
      rank += (wj > wt) || (wj == wt && j < t);
    }
    perm[rank] = t;
    sorted[rank] = wt;
  }
  __syncthreads();
  if (tid == 0) {
    double tot2 = 0.0;
    for (int j = M - 1; j >= 0; --j) tot2 += sorted[j] * sorted[j];
    const double tot = sqrt(tot2);
    int kk;
    if (d.mode == 0) {
      kk = 0;
      while (kk < M && sorted[kk] > d.cutoff) ++kk;
    } else {
      kk = M;
      if (d.chi_max > 0 && d.chi_max < kk) kk = d.chi_max;
      int ksvd = 0;
      const double smin = d.svd_min * tot;  // s / tot > svd_min
      while (ksvd < M && sorted[ksvd] > smin) ++ksvd;
      if (ksvd >= 1 && ksvd < kk) kk = ksvd;
      if (d.trunc_cut > 0.0) {
        const double lim = d.trunc_cut * d.trunc_cut * tot2;
        double c = 0.0;
        int first_good = -1;
        for (int j = M - 1; j >= 0; --j) {
          c += sorted[j] * sorted[j];
          if (c > lim) {
            first_good = j;
            break;
          }
        }
        if (first_good >= 0 && first_good + 1 < kk) kk = first_good + 1;
      }
    }
    if (kk < 1) kk = 1;
    if (kk > d.chi_cap) {
      kk = d.chi_cap;
      atomicAdd(&d.flags[0], 1);
    }
    double kept2 = 0.0;
    for (int j = kk - 1; j >= 0; --j) kept2 += sorted[j] * sorted[j];
    double ren = sqrt(kept2);
    if (!(ren > 0.0)) ren = 1.0;
    s_renorm = ren;
    s_k = kk;
    d.knew[b.slot] = kk;
    d.renorm[b.slot] = ren;
    d.chi[(size_t)b.r * (d.L + 1) + b.i + 1] = kk;
    double disc = 0.0;
    for (int j = M - 1; j >= kk; --j) disc += sorted[j] * sorted[j];
    if (tot2 > 0.0) d.trunc_err[(size_t)b.r * (d.L + 1) + b.i + 1] += disc / tot2;
  }
  __syncthreads();
  const int kk = s_k;
  const double inv = 1.0 / s_renorm;
  double *Sout = S_ptr(d, b.r, b.i + 1);
  for (int j = tid; j < kk; j += NT) Sout[j] = sorted[j] * inv;
  // B_{i+1}[k][p1][b] = row perm[k] of X / sigma_k, columns back from the interleaved order 2 b + p1
  cplx *Bn = site_ptr(d, b.r, b.i + 1);
  for (int e = tid; e < kk * N; e += NT) {
    const int k = e / N, c = e - k * N;  // c = interleaved column
    const double s = sorted[k];
    const cplx v = X[(size_t)perm[k] * N + c];
    Bn[(size_t)k * N + (c & 1) * chiR + (c >> 1)] = s > 0.0 ? cscale(v, 1.0 / s) : cmake(0.0, 0.0);
  }
}
}  // namespace tcj
