// pair_visit.cu -- what would a HALF-WARP row layout buy the blocked Jacobi kernel?  The product kernel keeps one
// 256-column stationary row per warp (8 complex per lane) and has 16 row pairs in flight per SM; its pair visit is a
// latency chain (LDS -> dot -> butterfly -> set-up -> rotate -> STS).  With a row on 16 lanes (16 complex per lane) a
// warp carries TWO independent pairs through one instruction stream: 32 pairs in flight, one butterfly level less, the
// set-up shared by two pairs -- at 64 data registers per thread.  This benchmark times the bare visit (rows of the
// partner block in shared memory, no hand-over, no staging) for both layouts with the product's own primitives.
//   mode 0: full-warp rows, 16 warps (tcb::pair_reg<8, true>), each warp alternates between two private partner rows
//   mode 1: half-warp rows, 16 warps, each half-warp revisits one private partner row
//   mode 2: mode 0 revisiting ONE private row (the same store -> load dependence as mode 1)
//   mode 3: half-warp rows, 8 warps (255 registers)
//   modes 4, 5, 6: full-warp rows with 24 / 32 / 12 warps (80 / 64 / 168 registers): sensitivity to the pairs in flight
// each in two flavours: every visit rotates / no visit rotates (threshold never met).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o pair_visit pair_visit.cu && ./pair_visit
#include <cstdio>
#include <cstdlib>
#include "../../time_crystal_tensor_network_b200/csrc/tc_jacobi_blocked.cuh"

using tcb::Rot;
__device__ __forceinline__ double frac(double x) { return x - floor(x); }

__device__ __forceinline__ void half_sum2(double &a, double &b) {
  const bool hi = (threadIdx.x & 8) != 0;
  double k = hi ? b : a;
  k += __shfl_xor_sync(0xffffffffu, hi ? a : b, 8);
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) k += __shfl_xor_sync(0xffffffffu, k, o);
  const double other = __shfl_xor_sync(0xffffffffu, k, 8);
  a = hi ? other : k;
  b = hi ? k : other;
}

// one pair per half-warp: row i on the 16 lanes of this half (u, 16 complex per lane), row j in shared memory
__device__ __forceinline__ int pair_half(cplx (&u)[16], cplx *xj, int hl, double &ai, double &wi, double2 *nj, double dead,
                                         double tol2, double small2) {
  cplx v[16];
  double g0 = 0.0, g1 = 0.0, h0 = 0.0, h1 = 0.0;
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    v[e] = xj[hl + 16 * e];
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
  const bool rot = tcb::make_rot(ai > dead && sj.x > dead, ai, sj.x, wi, sj.y, gr, gi, tol2, small2, r, big);
  if (!__any_sync(0xffffffffu, rot)) return big;
  if (!rot) {  // the other half of the warp rotates: this half applies the identity
    r.ar = r.ai = r.br = r.bi = 0.0;
    r.c2 = 1.0;
    r.ni = ai;
    r.nj = sj.x;
  }
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    tcb::rot_apply(u[e], v[e], r);
    xj[hl + 16 * e] = v[e];
  }
  ai = r.ni;
  wi *= r.c2;
  if (hl == 0) *nj = make_double2(r.nj, sj.y * r.c2);
  return big | ((int)rot << 16);
}

template <int MODE>
__global__ void __launch_bounds__(MODE == 3 ? 256 : MODE == 4 ? 768 : MODE == 5 ? 1024 : MODE == 6 ? 384 : 512, 1) k(double *out, int iters, double tol2) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int N = 256;
  cplx *rows = reinterpret_cast<cplx *>(smem_raw);                     // 32 rows
  double2 *nrm = reinterpret_cast<double2 *>(smem_raw + 32 * N * 16);  // 32 x {norm^2, scale^2}
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < 32 * N; i += blockDim.x) {
    const int r = i / N, c = i % N;
    rows[i] = make_double2(frac(0.37 * r + 0.011 * c) + (c == r ? 3.0 : 0.0), frac(0.21 * r + 0.017 * c) - 0.5);
  }
  __syncthreads();
  if (tid < 32) {
    double s = 0.0;
    for (int c = 0; c < N; ++c) s += cabs2(rows[tid * N + c]);
    nrm[tid] = make_double2(s, 1.0);
  }
  __syncthreads();
  int acc = 0;
  long long t0, t1;
  if (MODE == 0 || MODE == 2 || MODE >= 4) {
    cplx u[8];
    double ai = 0.0, wi = 1.0;
    for (int e = 0; e < 8; ++e) {
      u[e] = make_double2(frac(0.3 * warp + 0.02 * (lane + 32 * e)) - 0.5, frac(0.1 * warp + 0.03 * (lane + 32 * e)) + (e == 0 && lane == warp ? 2.0 : 0.0));
      ai += cabs2(u[e]);
    }
    ai = tcj::warp_sum(ai);
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int j = MODE == 0 ? 2 * warp + (it & 1) : warp;
      acc += tcb::pair_reg<8, true>(u, rows + (size_t)j * N, N, lane, ai, wi, nrm + j, 0.0, tol2, 1e-30);
      __syncwarp();
    }
    t1 = clock64();
    double s = ai + wi;
    for (int e = 0; e < 8; ++e) s += u[e].x + u[e].y;
    if (s == 123.456) out[2] = s;
  } else {
    cplx u[16];
    const int hl = lane & 15, half = 2 * warp + (lane >> 4);
    double ai = 0.0, wi = 1.0;
    for (int e = 0; e < 16; ++e) {
      u[e] = make_double2(frac(0.3 * half + 0.02 * (hl + 16 * e)) - 0.5, frac(0.1 * half + 0.03 * (hl + 16 * e)) + (e == 0 && hl == (half & 15) ? 2.0 : 0.0));
      ai += cabs2(u[e]);
    }
    for (int o = 8; o > 0; o >>= 1) ai += __shfl_xor_sync(0xffffffffu, ai, o);
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      acc += pair_half(u, rows + (size_t)half * N, hl, ai, wi, nrm + half, 0.0, tol2, 1e-30);
      __syncwarp();
    }
    t1 = clock64();
    double s = ai + wi;
    for (int e = 0; e < 16; ++e) s += u[e].x + u[e].y;
    if (s == 123.456) out[2] = s;
  }
  if (tid == 0 && blockIdx.x == 0) {
    out[0] = (double)(t1 - t0);
    out[1] = (double)(acc >> 16);
  }
}

template <int MODE>
static void run(const char *name, int warps, int pairs_per_warp, double *out, int sms) {
  const int iters = 4000;
  const size_t smem = 32 * 256 * 16 + 32 * 16;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (double tol2 : {-1.0, 1e30}) {
    for (int rep = 0; rep < 2; ++rep) {
      k<MODE><<<sms, warps * 32, smem>>>(out, iters, tol2);
      cudaDeviceSynchronize();
    }
    double h[2];
    cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("%-46s %-9s %6.0f clk per visit per warp, %6.1f clk per row pair per SM  (rotations by warp 0: %.0f)\n", name,
           tol2 < 0 ? "rotate" : "no-rotate", h[0] / iters, h[0] / iters / (warps * pairs_per_warp), h[1]);
  }
}

int main() {
  double *out;
  cudaMalloc(&out, 64);
  int sms;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  run<0>("full-warp rows, 16 warps, two partner rows", 16, 1, out, sms);
  run<2>("full-warp rows, 16 warps, one partner row", 16, 1, out, sms);
  run<1>("half-warp rows, 16 warps (2 pairs per warp)", 16, 2, out, sms);
  run<3>("half-warp rows,  8 warps (2 pairs per warp)", 8, 2, out, sms);
  run<6>("full-warp rows, 12 warps, one partner row", 12, 1, out, sms);
  run<4>("full-warp rows, 24 warps, one partner row", 24, 1, out, sms);
  run<5>("full-warp rows, 32 warps, one partner row", 32, 1, out, sms);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
