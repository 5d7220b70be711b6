// tc_observe.cuh -- state set-up, single-site gates and the observable kernels (HBM-bound passes).
//
// Every site tensor is kept in right-canonical 'B' form with the Schmidt values of the bond on its
// left in S, so theta_i = S_i B_i and
//   <O_i>     = sum_{a,b,p,q} S_a^2 conj(B[a,p,b]) O[p,q] B[a,q,b]        (MPS.expectation_value, observables.py:62)
//   S_ent(b)  = -sum_k s_k^2 ln s_k^2, s_k^2 > 1e-30                      (MPS.entanglement_entropy, tensor_utils.py:180)
//   <psi0|psi> with psi0 a product state = B_0[idx_0] B_1[idx_1] ...      (MPS.overlap, observables.py:25)
#pragma once
#include "tc_common.cuh"

namespace tco {
constexpr int NT = 256;

// MPS.from_product_state (tensor_utils.py:60): chi = 1 everywhere, one-hot tensors, S = 1.
__global__ void product_state_kernel(TcDev d) {
  const int r = blockIdx.x;
  for (int i = threadIdx.x; i <= d.L; i += blockDim.x) {
    d.chi[(size_t)r * (d.L + 1) + i] = 1;
    S_ptr(d, r, i)[0] = 1.0;
    d.trunc_err[(size_t)r * (d.L + 1) + i] = 0.0;
    if (i < d.L) {
      cplx *B = site_ptr(d, r, i);
      const int idx = d.init_idx[(size_t)r * d.L + i];
      B[0] = cmake(idx == 0 ? 1.0 : 0.0, 0.0);
      B[1] = cmake(idx == 1 ? 1.0 : 0.0, 0.0);
    }
  }
}

// single-site operator on the physical leg: B[a,p,b] <- sum_q op[p,q] B[a,q,b]
// (MPS.apply_local_op one-site, unitary=True: no re-canonicalisation; kicked_ising.py:206).
// grid (L or 1, R or 1); op_fixed == nullptr -> the model's kick of chain r
__global__ void one_site_kernel(TcDev d, int r0, int site0, const cplx *op_fixed) {
  const int r = r0 + blockIdx.y, site = site0 + blockIdx.x;
  const cplx *op = op_fixed ? op_fixed : d.kick + (size_t)r * 4;
  const cplx o00 = op[0], o01 = op[1], o10 = op[2], o11 = op[3];
  const int *c = d.chi + (size_t)r * (d.L + 1);
  const int chiL = c[site], chiR = c[site + 1];
  cplx *B = site_ptr(d, r, site);
  for (int e = threadIdx.x; e < chiL * chiR; e += blockDim.x) {
    const int a = e / chiR, b = e - a * chiR;
    cplx *p0 = B + (size_t)(2 * a) * chiR + b, *p1 = p0 + chiR;
    const cplx x0 = *p0, x1 = *p1;
    cplx y0 = cmul(o00, x0), y1 = cmul(o10, x0);
    cfma(y0, o01, x1);
    cfma(y1, o11, x1);
    *p0 = y0;
    *p1 = y1;
  }
}

// one CTA per (site, chain): single-site reduced density matrix, <Z>, and the entropy of the bond
// to the right of the site.  rdm[R][L][4] = (rho00, rho11, Re rho01, Im rho01), rho_pq = <p|rho|q>.
// Output rows are indexed by the chain's position in the launch (r - r0) plus out_r0.
__global__ void __launch_bounds__(NT) measure_kernel(TcDev d, int r0, double *rdm, double *Z, double *ent) {
  const int site = blockIdx.x, r = r0 + blockIdx.y;  // grid (L, chains of this launch), chains r0 ..
  const int *c = d.chi + (size_t)r * (d.L + 1);
  const int chiL = c[site], chiR = c[site + 1];
  const cplx *B = site_ptr(d, r, site);
  const double *S = S_ptr(d, r, site);
  __shared__ double red[32];
  double r00 = 0.0, r11 = 0.0, re01 = 0.0, im01 = 0.0;
  for (int e = threadIdx.x; e < chiL * chiR; e += NT) {
    const int a = e / chiR, b = e - a * chiR;
    const double s2 = S[a] * S[a];
    const cplx x0 = B[(size_t)(2 * a) * chiR + b], x1 = B[(size_t)(2 * a + 1) * chiR + b];
    r00 = fma(s2, cabs2(x0), r00);
    r11 = fma(s2, cabs2(x1), r11);
    // rho01 = sum theta[a,0,b] conj(theta[a,1,b])
    re01 = fma(s2, x0.x * x1.x + x0.y * x1.y, re01);
    im01 = fma(s2, x0.y * x1.x - x0.x * x1.y, im01);
  }
  r00 = block_sum(r00, red);
  r11 = block_sum(r11, red);
  if (rdm) {
    re01 = block_sum(re01, red);
    im01 = block_sum(im01, red);
  }
  if (threadIdx.x == 0) {
    const size_t o = (size_t)r * d.L + site;
    if (rdm) {
      rdm[o * 4 + 0] = r00;
      rdm[o * 4 + 1] = r11;
      rdm[o * 4 + 2] = re01;
      rdm[o * 4 + 3] = im01;
    }
    if (Z) Z[o] = r00 - r11;
  }
  if (ent && site < d.L - 1) {
    const double *Sr = S_ptr(d, r, site + 1);
    double h = 0.0;
    for (int k = threadIdx.x; k < chiR; k += NT) {
      const double p = Sr[k] * Sr[k];
      if (p > 1e-30) h -= p * log(p);
    }
    h = block_sum(h, red);
    if (threadIdx.x == 0) ent[(size_t)r * (d.L - 1) + site] = h;
  }
}

// <psi0|psi> for the product state psi0 given to tc_set_product_state (reference op: MPS.overlap behind
// calculate_loschmidt_echo, src/core/observables.py:11-38): the contraction is a chain of (chi_l x chi_r) slices
// M_i = B_i[:, idx_i, :] of the site tensors, <psi0|psi> = e^T M_0 M_1 ... M_{L-1} e.  The chain is cut in the middle
// and the two halves run as TWO CTAs per chain: blockIdx.y = 0 carries a row vector from the left end (v <- v M_i),
// blockIdx.y = 1 a column vector from the right end (w <- M_i w); whichever CTA finishes second (a counter per chain)
// forms v . w.  32 warps per CTA so that every row of a 128-row slice is in flight at once: a site costs one round trip
// to memory instead of four.
//   left half:  warp w takes rows w, w + 32, ...; every lane accumulates its columns, the 32 partial vectors are added
//               through shared memory (two levels);
//   right half: warp w takes rows w, w + 32, ...; a row is one dot product with w (coalesced read, warp reduction).
// scratch: ovs[R][2][chi_cap] complex (the two half vectors), ovc[R] arrival counters (left at zero for the next call).
// dynamic smem: (2 chi_cap + 32 * OVC) cplx
constexpr int OVC = 128;    // columns per pass of the left half
constexpr int NT_OV = 1024;
// the slice of site i (chi_l rows of chi_r complex, row a at (2 a + idx) chi_r) into L2 ahead of its use: the loads of a
// site do not depend on the vector, only the FMAs do, so the next site's lines travel while this site is reduced
__device__ __forceinline__ void ov_prefetch(const cplx *B, int chiL, int chiR, int tid) {
  const int lines_per_row = (chiR * (int)sizeof(cplx) + 127) / 128;
  for (int e = tid; e < chiL * lines_per_row; e += 1024) {
    const int a2 = e / lines_per_row, l = e - a2 * lines_per_row;
    const char *p = reinterpret_cast<const char *>(B + (size_t)(2 * a2) * chiR) + 128 * l;
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
  }
}

__global__ void __launch_bounds__(NT_OV) overlap_product_kernel(TcDev d, int r0, double *ov, cplx *ovs, int *ovc) {
  const int r = r0 + blockIdx.x, dir = blockIdx.y;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx *v = reinterpret_cast<cplx *>(smem_raw);
  cplx *vn = v + d.chi_cap;
  cplx *part = vn + d.chi_cap;  // [32][OVC]
  __shared__ int s_chi[1026];
  __shared__ int8_t s_idx[1025];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = NT_OV / 32;
  const int L = d.L, Lh = L / 2;  // the left half covers sites 0 .. Lh-1, the right half sites Lh .. L-1
  // bond dimensions and basis indices of the chain into shared memory once (the slice addresses depend on them: two
  // dependent global loads per site otherwise); chains longer than 1024 sites read them from global memory
  const bool cached = L <= 1024;
  const int *c = d.chi + (size_t)r * (L + 1);
  if (cached) {
    for (int k = tid; k <= L; k += NT_OV) s_chi[k] = c[k];
    for (int k = tid; k < L; k += NT_OV) s_idx[k] = d.init_idx[(size_t)r * L + k];
  }
  if (tid == 0) v[0] = cmake(1.0, 0.0);
  __syncthreads();
  if (dir == 0) {
    for (int i = 0; i < Lh; ++i) {
      const int chiL = cached ? s_chi[i] : c[i], chiR = cached ? s_chi[i + 1] : c[i + 1];
      const int idx = cached ? s_idx[i] : d.init_idx[(size_t)r * L + i];
      const cplx *B = site_ptr(d, r, i) + (size_t)idx * chiR;
      for (int j = (i == 0 ? 1 : 2); cached && j <= 2 && i + j < Lh; ++j)  // two sites ahead (site 1 too at the start)
        ov_prefetch(site_ptr(d, r, i + j) + (size_t)s_idx[i + j] * s_chi[i + j + 1], s_chi[i + j], s_chi[i + j + 1], tid);
      for (int c0 = 0; c0 < chiR; c0 += OVC) {
        cplx acc[OVC / 32];
#pragma unroll
        for (int k = 0; k < OVC / 32; ++k) acc[k] = cmake(0.0, 0.0);
#pragma unroll 4
        for (int a2 = warp; a2 < chiL; a2 += NW) {
          const cplx va = v[a2];
          const cplx *row = B + (size_t)(2 * a2) * chiR + c0;
#pragma unroll
          for (int k = 0; k < OVC / 32; ++k) {
            const int b = lane + 32 * k;
            if (c0 + b < chiR) cfma(acc[k], va, row[b]);
          }
        }
#pragma unroll
        for (int k = 0; k < OVC / 32; ++k) part[warp * OVC + lane + 32 * k] = acc[k];
        __syncthreads();
        {  // 32 partial vectors -> 8 (thread t: column t % OVC, warps 4 g .. 4 g + 3 with g = t / OVC)
          const int b = tid % OVC, g = tid / OVC;
          cplx t = part[(4 * g) * OVC + b];
#pragma unroll
          for (int w = 1; w < 4; ++w) t = cadd(t, part[(4 * g + w) * OVC + b]);
          __syncthreads();
          part[g * OVC + b] = t;
        }
        __syncthreads();
        for (int b = tid; b < OVC && c0 + b < chiR; b += NT_OV) {
          cplx t = part[b];
#pragma unroll
          for (int g = 1; g < NT_OV / OVC; ++g) t = cadd(t, part[g * OVC + b]);
          vn[c0 + b] = t;
        }
        __syncthreads();
      }
      cplx *t = v;
      v = vn;
      vn = t;
    }
  } else {
    for (int i = L - 1; i >= Lh; --i) {
      const int chiL = cached ? s_chi[i] : c[i], chiR = cached ? s_chi[i + 1] : c[i + 1];
      const int idx = cached ? s_idx[i] : d.init_idx[(size_t)r * L + i];
      const cplx *B = site_ptr(d, r, i) + (size_t)idx * chiR;
      for (int j = (i == L - 1 ? 1 : 2); cached && j <= 2 && i - j >= Lh; ++j)
        ov_prefetch(site_ptr(d, r, i - j) + (size_t)s_idx[i - j] * s_chi[i - j + 1], s_chi[i - j], s_chi[i - j + 1], tid);
#pragma unroll 4
      for (int a2 = warp; a2 < chiL; a2 += NW) {
        const cplx *row = B + (size_t)(2 * a2) * chiR;
        cplx acc = cmake(0.0, 0.0);
        for (int b = lane; b < chiR; b += 32) cfma(acc, row[b], v[b]);
        for (int o = 16; o > 0; o >>= 1) {
          acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
          acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
        }
        if (lane == 0) vn[a2] = acc;
      }
      __syncthreads();
      cplx *t = v;
      v = vn;
      vn = t;
    }
  }
  // ---- hand the half vector over; the CTA that arrives second closes the contraction
  const int n = c[Lh];
  cplx *mine = ovs + ((size_t)r * 2 + dir) * d.chi_cap;
  for (int k = tid; k < n; k += NT_OV) mine[k] = v[k];
  __threadfence();
  __syncthreads();
  __shared__ int s_last;
  if (tid == 0) s_last = atomicAdd(ovc + r, 1);
  __syncthreads();
  if (s_last == 0) return;
  __threadfence();
  // both halves are read back from the scratch in a fixed order (left . right), so that the result does not depend on
  // which CTA happens to arrive second (the records of a chain are bit-identical from run to run)
  const cplx *left = ovs + ((size_t)r * 2 + 0) * d.chi_cap, *right = left + d.chi_cap;
  if (warp == 0) {
    cplx acc = cmake(0.0, 0.0);
    for (int k = lane; k < n; k += 32) {
      const double2 a = __ldcg(reinterpret_cast<const double2 *>(left + k));
      const double2 b = __ldcg(reinterpret_cast<const double2 *>(right + k));
      cfma(acc, cmake(a.x, a.y), cmake(b.x, b.y));
    }
    for (int o = 16; o > 0; o >>= 1) {
      acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
      acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
    }
    if (lane == 0) {
      ov[(size_t)r * 2 + 0] = acc.x;
      ov[(size_t)r * 2 + 1] = acc.y;
      ovc[r] = 0;  // ready for the next snapshot
    }
  }
}

// chains r0 .. r0 + nr - 1
__global__ void chi_record_kernel(TcDev d, int r0, int nr, int32_t *out) {
  const size_t lo = (size_t)r0 * (d.L + 1), n = (size_t)nr * (d.L + 1);
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x)
    out[lo + e] = d.chi[lo + e];
}

// General transfer-matrix contraction between chain ra of context da (bra, conjugated) and chain rb
// of context db (ket) over sites lo..hi, with optional single-site operators at lo and hi:
//   E'[a',b'] = sum_{a,b,p,q} conj(A[a,p,a']) O[p,q] E[a,b] K[b,q,b']
// init_mode 0: E = [[1]] (overlap from site 0);  1: E = diag(S_lo^2) of the ket (correlators in one
// state, whose left environment in canonical form is the squared Schmidt spectrum).
// Result = trace(E) after site hi.  One CTA; E0/E1: chi_cap_a*chi_cap_b cplx, T: chi_cap_a*2*chi_cap_b.
__global__ void __launch_bounds__(NT) transfer_kernel(TcDev da, int ra, TcDev db, int rb, int lo, int hi,
                                                      const cplx *op_lo, const cplx *op_hi, int init_mode,
                                                      cplx *E0, cplx *E1, cplx *T, double *out) {
  const int tid = threadIdx.x;
  const int *ca = da.chi + (size_t)ra * (da.L + 1), *cb = db.chi + (size_t)rb * (db.L + 1);
  cplx *E = E0, *En = E1;
  {
    const int na = ca[lo], nb = cb[lo];
    const double *S = S_ptr(db, rb, lo);
    for (int e = tid; e < na * nb; e += NT) {
      const int x = e / nb, y = e - x * nb;
      double v = 0.0;
      if (x == y) v = init_mode == 0 ? 1.0 : S[x] * S[x];
      E[e] = cmake(v, 0.0);
    }
  }
  __syncthreads();
  for (int i = lo; i <= hi; ++i) {
    const int na = ca[i], nb = cb[i], na2 = ca[i + 1], nb2 = cb[i + 1];
    const cplx *A = site_ptr(da, ra, i), *K = site_ptr(db, rb, i);
    const cplx *op = (i == lo && op_lo) ? op_lo : ((i == hi && op_hi) ? op_hi : nullptr);
    cplx o00 = cmake(1, 0), o01 = cmake(0, 0), o10 = cmake(0, 0), o11 = cmake(1, 0);
    if (op) {
      o00 = op[0];
      o01 = op[1];
      o10 = op[2];
      o11 = op[3];
      if (i == lo && i == hi && op_lo && op_hi) {  // both operators on one site: O = op_lo op_hi
        const cplx h00 = op_hi[0], h01 = op_hi[1], h10 = op_hi[2], h11 = op_hi[3];
        const cplx l00 = op_lo[0], l01 = op_lo[1], l10 = op_lo[2], l11 = op_lo[3];
        o00 = cadd(cmul(l00, h00), cmul(l01, h10));
        o01 = cadd(cmul(l00, h01), cmul(l01, h11));
        o10 = cadd(cmul(l10, h00), cmul(l11, h10));
        o11 = cadd(cmul(l10, h01), cmul(l11, h11));
      }
    }
    // T[a][p][b'] = sum_q O[p,q] sum_b E[a,b] K[b,q,b']
    for (int e = tid; e < na * nb2; e += NT) {
      const int x = e / nb2, y = e - x * nb2;
      cplx t0 = cmake(0, 0), t1 = cmake(0, 0);
      for (int b = 0; b < nb; ++b) {
        const cplx ev = E[x * nb + b];
        cfma(t0, ev, K[(size_t)(2 * b) * nb2 + y]);
        cfma(t1, ev, K[(size_t)(2 * b + 1) * nb2 + y]);
      }
      cplx u0 = cmul(o00, t0), u1 = cmul(o10, t0);
      cfma(u0, o01, t1);
      cfma(u1, o11, t1);
      T[(size_t)(2 * x) * nb2 + y] = u0;
      T[(size_t)(2 * x + 1) * nb2 + y] = u1;
    }
    __syncthreads();
    // E'[a'][b'] = sum_{a,p} conj(A[a,p,a']) T[a,p,b']
    for (int e = tid; e < na2 * nb2; e += NT) {
      const int x = e / nb2, y = e - x * nb2;
      cplx acc = cmake(0, 0);
      for (int ap = 0; ap < 2 * na; ++ap) cfmac(acc, A[(size_t)ap * na2 + x], T[(size_t)ap * nb2 + y]);
      En[e] = acc;
    }
    __syncthreads();
    cplx *t = E;
    E = En;
    En = t;
  }
  if (tid == 0) {
    const int na = ca[hi + 1], nb = cb[hi + 1];
    const int n = na < nb ? na : nb;
    cplx tr = cmake(0, 0);
    for (int k = 0; k < n; ++k) tr = cadd(tr, E[k * nb + k]);
    out[0] = tr.x;
    out[1] = tr.y;
  }
}

// copy chain rs of src into chain rd of dst (compact site layout is chi_cap independent).
__global__ void copy_chain_kernel(TcDev dst, int rd, TcDev src, int rs) {
  const int site = blockIdx.x;  // 0..L (L = the last bond only)
  const int *c = src.chi + (size_t)rs * (src.L + 1);
  const int n = c[site];
  if (threadIdx.x == 0) {
    dst.chi[(size_t)rd * (dst.L + 1) + site] = n;
    dst.trunc_err[(size_t)rd * (dst.L + 1) + site] = src.trunc_err[(size_t)rs * (src.L + 1) + site];
    if (site < src.L) dst.init_idx[(size_t)rd * dst.L + site] = src.init_idx[(size_t)rs * src.L + site];
  }
  const double *Ss = S_ptr(src, rs, site);
  double *Sd = S_ptr(dst, rd, site);
  for (int k = threadIdx.x; k < n; k += blockDim.x) Sd[k] = Ss[k];
  if (site < src.L) {
    const int tot = n * 2 * c[site + 1];
    const cplx *Bs = site_ptr(src, rs, site);
    cplx *Bd = site_ptr(dst, rd, site);
    for (int e = threadIdx.x; e < tot; e += blockDim.x) Bd[e] = Bs[e];
  }
}
}  // namespace tco
