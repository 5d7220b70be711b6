// tc_gemm.cuh -- batched ragged complex-FP64 GEMM on the FP64 tensor pipe (DMMA, mma.sync m8n8k4.f64).
//
// One launch covers every (chain, bond) matrix of a layer; each matrix has its own (M, N, K), read
// from the device-side bond-dimension table by the problem policy `P`.  tcgen05 has no f64 kind, so
// the FP64 tensor path on sm_100a is the warp-level mma.sync (SASS: DMMA).
//
// CTA tile 64 x 64 x 16, 4 warps (2 x 2), warp tile 32 x 32 = 4 x 4 DMMA tiles, complex product as
// four real DMMAs (re/im accumulators kept separately).  Operand tiles are staged K-contiguous in
// shared memory as interleaved complex with a row stride of 20 elements (320 B): the fragment read
// of a quarter warp (2 rows x 4 k, 16 B each) then covers all 32 banks exactly once.
//
// Policy interface:
//   bool   init(const TcDev&, const LayerArgs&, int jb, int ry)   -> false: nothing to do
//   int    M, N, K
//   cplx   A(int row, int k)      A operand element (0 outside the matrix)
//   cplx   B(int col, int k)      B operand element, i.e. the product is sum_k A(row,k) * B(col,k)
//   void   store(int row, int col, cplx v)
//   static constexpr bool A_KCONTIG, B_KCONTIG   which index is contiguous in global memory
#pragma once
#include "tc_common.cuh"

namespace tcg {
constexpr int BM = 64, BN = 64, BK = 16, LDS = BK + 4, NT = 128;

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <class P>
__global__ void __launch_bounds__(NT) gemm_kernel(TcDev d, LayerArgs a) {
  P p;
  if (!p.init(d, a, blockIdx.y, blockIdx.z)) return;
  const int tiles_n = (p.N + BN - 1) / BN, tiles_m = (p.M + BM - 1) / BM;
  const int t = blockIdx.x;
  if (t >= tiles_m * tiles_n) return;
  const int row0 = (t / tiles_n) * BM, col0 = (t % tiles_n) * BN;

  __shared__ cplx As[BM * LDS];
  __shared__ cplx Bs[BN * LDS];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 1, wn = warp & 1;
  const int fr = lane >> 2, fk = lane & 3;

  double cre[4][4][2], cim[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) cre[i][j][0] = cre[i][j][1] = cim[i][j][0] = cim[i][j][1] = 0.0;

  for (int k0 = 0; k0 < p.K; k0 += BK) {
    // ---- stage the two operand tiles (with the policy's on-load transform)
#pragma unroll
    for (int q = 0; q < (BM * BK) / NT; ++q) {
      const int e = tid + NT * q;
      int x, kk;
      if (P::A_KCONTIG) { kk = e % BK; x = e / BK; } else { x = e % BM; kk = e / BM; }
      As[x * LDS + kk] = p.A(row0 + x, k0 + kk);
    }
#pragma unroll
    for (int q = 0; q < (BN * BK) / NT; ++q) {
      const int e = tid + NT * q;
      int x, kk;
      if (P::B_KCONTIG) { kk = e % BK; x = e / BK; } else { x = e % BN; kk = e / BN; }
      Bs[x * LDS + kk] = p.B(col0 + x, k0 + kk);
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; kk += 4) {
      double are[4], aim[4], bre[4], bim[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        cplx v = As[(wm * 32 + i * 8 + fr) * LDS + kk + fk];
        are[i] = v.x;
        aim[i] = v.y;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        cplx v = Bs[(wn * 32 + j * 8 + fr) * LDS + kk + fk];
        bre[j] = v.x;
        bim[j] = v.y;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          dmma(cre[i][j][0], cre[i][j][1], are[i], bre[j]);
          dmma(cim[i][j][0], cim[i][j][1], are[i], bim[j]);
          dmma(cre[i][j][0], cre[i][j][1], -aim[i], bim[j]);
          dmma(cim[i][j][0], cim[i][j][1], aim[i], bre[j]);
        }
    }
    __syncthreads();
  }
  // ---- epilogue
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int row = row0 + wm * 32 + i * 8 + fr;
      const int col = col0 + wn * 32 + j * 8 + 2 * fk;
      if (row < p.M) {
        if (col < p.N) p.store(row, col, cmake(cre[i][j][0], cim[i][j][0]));
        if (col + 1 < p.N) p.store(row, col + 1, cmake(cre[i][j][1], cim[i][j][1]));
      }
    }
}
}  // namespace tcg
