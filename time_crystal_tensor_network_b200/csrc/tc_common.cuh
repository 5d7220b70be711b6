// tc_common.cuh -- shared device-side definitions for the B200 TEBD engine (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

typedef double2 cplx;

__host__ __device__ __forceinline__ cplx cmake(double re, double im) { return make_double2(re, im); }
__host__ __device__ __forceinline__ cplx cadd(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ cplx csub(cplx a, cplx b) { return make_double2(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ cplx cmul(cplx a, cplx b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// conj(a) * b
__host__ __device__ __forceinline__ cplx cmulc(cplx a, cplx b) {
  return make_double2(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x);
}
__host__ __device__ __forceinline__ cplx cconj(cplx a) { return make_double2(a.x, -a.y); }
__host__ __device__ __forceinline__ cplx cscale(cplx a, double s) { return make_double2(a.x * s, a.y * s); }
__host__ __device__ __forceinline__ double cabs2(cplx a) { return a.x * a.x + a.y * a.y; }
// a += b * c
__host__ __device__ __forceinline__ void cfma(cplx &a, cplx b, cplx c) {
  a.x = fma(b.x, c.x, a.x);
  a.x = fma(-b.y, c.y, a.x);
  a.y = fma(b.x, c.y, a.y);
  a.y = fma(b.y, c.x, a.y);
}
// a += conj(b) * c
__host__ __device__ __forceinline__ void cfmac(cplx &a, cplx b, cplx c) {
  a.x = fma(b.x, c.x, a.x);
  a.x = fma(b.y, c.y, a.x);
  a.y = fma(b.x, c.y, a.y);
  a.y = fma(-b.y, c.x, a.y);
}
__host__ __device__ __forceinline__ cplx crecip(cplx a) {
  // Smith's algorithm (as LAPACK zladiv) to avoid overflow/underflow in |a|^2
  if (fabs(a.x) >= fabs(a.y)) {
    double r = a.y / a.x, den = a.x + a.y * r;
    return make_double2(1.0 / den, -r / den);
  } else {
    double r = a.x / a.y, den = a.y + a.x * r;
    return make_double2(r / den, -1.0 / den);
  }
}

// Device view of a context: everything a kernel needs, passed by value.
struct TcDev {
  int L, chi_cap, R, n2, nbmax, ws_chains;
  size_t site_stride;  // cplx elements per site tensor slot  (chi_cap*2*chi_cap)
  size_t slot_stride;  // cplx elements per workspace matrix  (n2*n2)
  // state
  cplx *B;          // [R][L][site_stride], site (r,i) compact row-major [chi_l][2][chi_r]
  double *S;        // [R][L+1][chi_cap]
  int *chi;         // [R][L+1]
  int8_t *init_idx; // [R][L]
  // model
  const cplx *gates;  // [R][L-1][16]
  const cplx *kick;   // [R][4]
  int rot64;          // TC_ROT64 A/B switches of the Jacobi kernel: bit 1 = barrier per round instead of the row hand-over,
                      // bit 2 = no barrier between the visits of a pass (the last warp out stores and reloads the stage)
  int gates_diag;     // every gate of the model is diagonal (fused phase epilogue)
  double thr_sched[6];  // threshold Jacobi: sweeps 0..5 rotate only pairs with |g|^2 / (a_i a_j) above these
  double small_rel2;    // stopping rule: a sweep whose rotations all had |g|^2 / (a_i a_j) below this ends the iteration
                        // (tcj::SMALL_REL2; 0 with TC_EARLY_STOP=0: iterate until a sweep rotates nothing)
  double *trunc_err;  // [R][L+1] discarded weight accumulated per bond (single writer, deterministic)
  int *flags;         // [0]: chi_cap overflow count, [1]: Jacobi non-convergence count, [2]: sweeps max
  // workspace (one layer of ws_chains chains): slot = (r - r0)*nbmax + jb
  cplx *Cw;              // gate-applied two-site tensor C, [M][N]
  cplx *Xw;              // theta = S_l C, rows orthogonalised in place by the Jacobi kernel
  double *ww;            // [slots][n2] row norms (singular values, unsorted)
  int *perm;             // [slots][n2] descending order
  int *knew;             // [slots]
  double *renorm;        // [slots]
  // truncation
  int mode;
  double cutoff;
  int chi_max;
  double svd_min, trunc_cut;
};

// Which bonds a launch works on.
struct LayerArgs {
  int first_site;  // left site of bond jb = 0
  int site_step;   // 2 for an even/odd layer
  int nb;          // number of bonds
  int r0;          // first chain of this launch (grid.y / grid.z index ry -> chain r0 + ry)
  int nr;          // number of chains in this launch
  int kick_mode;   // bit0: kick both sites of every bond; bit1: kick right site of bond i == L-2
  const cplx *gate_override;  // != nullptr: this 4x4 gate for every bond instead of d.gates
  int diag;                   // the gates of this launch are diagonal: phase fused into the GEMM epilogue
  int slot0;                  // first workspace chain slot of this launch (chain groups own disjoint slot ranges)
};

struct Bond {
  int r, jb, i;
  int chiL, chiM, chiR;
  int M, N;  // theta is M x N (M = 2 chiL, N = 2 chiR)
  size_t slot;
};

__device__ __forceinline__ bool get_bond(const TcDev &d, const LayerArgs &a, int jb, int ry, Bond &b) {
  b.r = a.r0 + ry;
  b.jb = jb;
  b.i = a.first_site + a.site_step * jb;
  if (jb >= a.nb || b.i > d.L - 2 || ry >= a.nr || b.r >= d.R) return false;
  const int *c = d.chi + (size_t)b.r * (d.L + 1);
  b.chiL = c[b.i];
  b.chiM = c[b.i + 1];
  b.chiR = c[b.i + 2];
  b.M = 2 * b.chiL;
  b.N = 2 * b.chiR;
  b.slot = (size_t)(a.slot0 + ry) * d.nbmax + jb;
  return true;
}

// Largest matrices first (long SVD kernels): with grid (chains, bonds) the CTAs of bond rank y are scheduled before
// those of rank y+1; rank 0 is the centre of the chain, so the full-size updates start in the first wave and the
// small edge bonds fill the tail.
__device__ __forceinline__ int centre_out(int x, int nb) {
  const int c = nb / 2;
  const int k = (x + 1) / 2;
  return (x & 1) ? c - k : c + k;  // x = 0 -> c, 1 -> c-1, 2 -> c+1, ...: a bijection of [0, nb) for odd and even nb
}

__device__ __forceinline__ cplx *site_ptr(const TcDev &d, int r, int site) {
  return d.B + ((size_t)r * d.L + site) * d.site_stride;
}
__device__ __forceinline__ double *S_ptr(const TcDev &d, int r, int bond) {
  return d.S + ((size_t)r * (d.L + 1) + bond) * d.chi_cap;
}

// block-wide sum of one double over blockDim.x threads (blockDim.x multiple of 32, <= 1024).
// `scratch` must hold 32 doubles.  All threads get the result.
__device__ __forceinline__ double block_sum(double v, double *scratch) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (l == 0) scratch[w] = v;
  __syncthreads();
  double t = (l < nw) ? scratch[l] : 0.0;
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  return t;
}
