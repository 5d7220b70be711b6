// tc_engine.cu -- C ABI (include/tc_b200.h) of the B200-native kicked-Ising Floquet/TEBD engine.
//
// One context = an ensemble of R independent chains on one GPU.  Per two-site update (the unit the
// reference executes through MPS.apply_local_op, src/models/kicked_ising.py:162-188) the device runs
//   K1  C = gate . (kick? B_i)(kick? B_{i+1})          DMMA GEMM, kick fused into the operand loads,
//       theta = S_i C                                  diagonal Ising phase fused into the epilogue
//   K2  theta P = Q R (Q discarded), J^H R = Sigma V^H  Householder QR preconditioner, then one-sided
//                                                      Jacobi on the rows of R (tc_jacobi.cuh)
//       sort, truncate, renormalise, B_{i+1} = V_k^H   in-kernel
//   K3  B_i = C V_k / |Sigma_k|                        DMMA GEMM (inverse-free update)
// for every bond of a parity class and every chain in one launch each.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <string>
#include <vector>

#include "../../include/tc_b200.h"
#include "tc_common.cuh"
#include "tc_gemm.cuh"
#include "tc_theta.cuh"
#include "tc_jacobi.cuh"
#include "tc_jacobi_blocked.cuh"
#include "tc_jacobi_halfwarp.cuh"
#include "tc_jacobi_team.cuh"
#include "tc_observe.cuh"

// ------------------------------------------------------------------------------------------------
// errors / bookkeeping
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static std::atomic<long long> g_launches{0};

static int fail(const char *fmt, const char *a = "", const char *b = "") {
  char buf[512];
  snprintf(buf, sizeof(buf), fmt, a, b);
  g_err = buf;
  return 1;
}
#define CK(call)                                                                          \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess) return fail("CUDA error: %s at %s", cudaGetErrorString(e_), #call); \
  } while (0)
#define LAUNCHED()                                                                   \
  do {                                                                               \
    g_launches.fetch_add(1, std::memory_order_relaxed);                              \
    cudaError_t e_ = cudaGetLastError();                                             \
    if (e_ != cudaSuccess) return fail("kernel launch failed: %s", cudaGetErrorString(e_)); \
  } while (0)

struct tc_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false, own_arena = false;
  void *arena = nullptr;
  size_t arena_bytes = 0;
  TcDev d{};
  cplx *gates_dev = nullptr, *kick_dev = nullptr;  // non-const aliases of d.gates / d.kick
  cplx *op_scratch = nullptr;                      // 32 cplx: caller-supplied operators
  cplx *E0 = nullptr, *E1 = nullptr, *Tt = nullptr;  // transfer contraction scratch
  double *small_out = nullptr;                       // 8 doubles
  cplx *ovs = nullptr;                               // [R][2][chi_cap] half vectors of overlap_product_kernel
  int *ovc = nullptr;                                // [R] arrival counters of its two CTAs per chain
  bool have_model = false;
  // TC_SMALL_KERNEL=0: the full-warp 16-warp instance of the Jacobi kernel also for contexts whose widest matrix has 128
  // columns.  Default: such contexts (chi_cap <= 64, BASELINE configs 2 and 3) run the half-warp-row kernel -- ensembles
  // its 8-warp instance, two CTAs per SM (config 2's shape: 2410 -> 2914 chain-steps/s), single chains its 16-warp instance.
  // =2: the full-warp instance with row blocks of 8 (2504), =3 / =4: force the 8-warp / 16-warp half-warp instance (tests).
  // History: round 1 had a 16-warp, 64-register, two-CTAs-per-SM instance here (+8 % then); with the fast rotations that
  // one spilled (312 B) and lost 6 %.
  bool no_small_kernel = false;
  bool force_small_kernel = false;   // TC_SMALL_KERNEL=2: the narrow instance for every context with chi_cap <= 64 (tests)
  int halfwarp_kernel = 0;           // TC_SMALL_KERNEL=3: the half-warp-row kernel for every narrow context (A/B, tests)
  bool force_simple_jacobi = false;  // TC_JACOBI=simple: the warp-per-pair kernel for every size (A/B testing)
  bool old_theta = false;            // TC_THETA=v1: the policy-functor GEMM of round 1 for K1 (A/B testing)
  bool team_jacobi = false;          // TC_JACOBI=team: the two-warps-per-row kernel for the narrow matrices too (A/B testing)
  bool old_wide = false;             // TC_JACOBI=wide_v1: the warp-per-pair cluster kernel for chi_cap > 128 (A/B testing)
  // chain groups: the chains never interact, so G groups run their periods on G streams and the tail of one group's
  // layer (fewer CTAs than SMs left) overlaps the next kernels of the others.  TC_GROUPS, default 4.
  int ngroups = 1;
  int sm_count = 148;
  int wide_cluster = 0;  // TC_WIDE_CLUSTER: CTAs per matrix of the wide (chi_cap > 128) Jacobi kernel, 0 = automatic
  int qr_cluster = 0;    // TC_QR_CLUSTER: CTAs per matrix of the blocked QR kernel, 0 = automatic
  bool no_narrow_qr = false;  // TC_QR_NARROW=0: the 256-thread QR instance also for ensembles of narrow contexts (A/B)
  std::vector<cudaStream_t> gstreams;
  std::vector<cudaEvent_t> gjoin;
  cudaEvent_t gfork = nullptr;
  // record buffers for tc_floquet_run_host
  void *rec = nullptr;
  size_t rec_bytes = 0;
  // per-kernel-class CUDA-event timing (tc_profile): class, start, stop
  bool profile = false;
  struct Span {
    int cls;
    cudaEvent_t e0, e1;
  };
  std::vector<Span> spans;
  std::vector<cudaEvent_t> free_events;
};

static cudaEvent_t prof_event(tc_ctx *c) {
  cudaEvent_t e;
  if (!c->free_events.empty()) {
    e = c->free_events.back();
    c->free_events.pop_back();
  } else {
    cudaEventCreate(&e);
  }
  return e;
}
struct ProfScope {  // brackets the launches of one kernel class with events on the context's stream
  tc_ctx *c;
  tc_ctx::Span sp;
  ProfScope(tc_ctx *ctx, int cls) : c(ctx) {
    if (!c->profile) return;
    sp.cls = cls;
    sp.e0 = prof_event(c);
    sp.e1 = prof_event(c);
    cudaEventRecord(sp.e0, c->stream);
  }
  ~ProfScope() {
    if (!c->profile) return;
    cudaEventRecord(sp.e1, c->stream);
    c->spans.push_back(sp);
  }
};

// ------------------------------------------------------------------------------------------------
// arena layout
// ------------------------------------------------------------------------------------------------
static size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

struct Layout {
  size_t B, S, chi, init_idx, gates, kick, trunc_err, flags, Cw, Xw, ww, perm, knew, renorm, op, E0, E1, Tt,
      small_out, ovs, ovc, total;
  int ws_chains, nbmax, n2;
};

static Layout make_layout(int L, int chi_cap, int R, bool storage_only = false) {
  Layout o{};
  const size_t cs = sizeof(cplx);
  o.n2 = 2 * chi_cap;
  o.nbmax = L / 2 > 0 ? L / 2 : 1;
  const size_t site = (size_t)chi_cap * 2 * chi_cap, slot = (size_t)o.n2 * o.n2;
  size_t budget = (size_t)64 << 30;
  if (const char *e = getenv("TC_WS_BYTES")) budget = strtoull(e, nullptr, 10);
  const size_t per_chain = (size_t)o.nbmax * (2 * slot * cs + o.n2 * (sizeof(double) + sizeof(int)) + 16);
  long long wc = (long long)(budget / per_chain);
  if (wc < 1) wc = 1;
  if (wc > R) wc = R;
  if (storage_only) wc = 0;  // snapshots: state + model + observable scratch, no SVD workspace
  o.ws_chains = (int)wc;
  const size_t slots = (size_t)o.ws_chains * o.nbmax;
  size_t p = 0;
  auto take = [&](size_t bytes) {
    size_t at = p;
    p = align_up(p + bytes);
    return at;
  };
  o.B = take((size_t)R * L * site * cs);
  o.S = take((size_t)R * (L + 1) * chi_cap * sizeof(double));
  o.chi = take((size_t)R * (L + 1) * sizeof(int));
  o.init_idx = take((size_t)R * L);
  o.gates = take((size_t)R * (L > 1 ? L - 1 : 1) * 16 * cs);
  o.kick = take((size_t)R * 4 * cs);
  o.trunc_err = take((size_t)R * (L + 1) * sizeof(double));
  o.flags = take(8 * sizeof(int));
  o.Cw = take(slots * slot * cs);
  o.Xw = take(slots * slot * cs);
  o.ww = take(slots * o.n2 * sizeof(double));
  o.perm = take(slots * o.n2 * sizeof(int));
  o.knew = take(slots * sizeof(int));
  o.renorm = take(slots * sizeof(double));
  o.op = take(32 * cs);
  o.E0 = take((size_t)chi_cap * chi_cap * cs);
  o.E1 = take((size_t)chi_cap * chi_cap * cs);
  o.Tt = take((size_t)chi_cap * 2 * chi_cap * cs);
  o.small_out = take(8 * sizeof(double));
  o.ovs = take((size_t)R * 2 * chi_cap * cs);  // half vectors of the product-state overlap
  o.ovc = take((size_t)R * sizeof(int));
  o.total = p;
  return o;
}

// ------------------------------------------------------------------------------------------------
// GEMM policies
// ------------------------------------------------------------------------------------------------
// K1: T[(a,p0),(p1,b)] = sum_m (kick B_i)[a,p0,m] (kick B_{i+1})[m,p1,b]; diagonal gate: C = phase T,
// theta = S_i C written in the epilogue; general gate: raw T is stored and gate_mix_kernel follows.
struct ThetaPolicy {
  int M, N, K, chiM, chiR;
  const cplx *Bi, *Bn;
  cplx *C, *X;
  const double *S;
  cplx k00, k01, k10, k11, g0, g1, g2, g3;
  bool kickL, kickR, diag;
  static constexpr bool A_KCONTIG = true, B_KCONTIG = false;
  __device__ bool init(const TcDev &d, const LayerArgs &a, int jb, int ry) {
    Bond b;
    if (!get_bond(d, a, jb, ry, b)) return false;
    M = b.M;
    N = b.N;
    K = b.chiM;
    chiM = b.chiM;
    chiR = b.chiR;
    Bi = site_ptr(d, b.r, b.i);
    Bn = site_ptr(d, b.r, b.i + 1);
    C = d.Cw + b.slot * d.slot_stride;
    X = d.Xw + b.slot * d.slot_stride;
    S = S_ptr(d, b.r, b.i);
    kickL = (a.kick_mode & 1) != 0;
    kickR = kickL || ((a.kick_mode & 2) && b.i == d.L - 2);
    const cplx *kk = d.kick + (size_t)b.r * 4;
    k00 = kk[0];
    k01 = kk[1];
    k10 = kk[2];
    k11 = kk[3];
    diag = a.diag != 0;
    const cplx *g = a.gate_override ? a.gate_override : d.gates + ((size_t)b.r * (d.L - 1) + b.i) * 16;
    g0 = g[0];
    g1 = g[5];
    g2 = g[10];
    g3 = g[15];
    return true;
  }
  __device__ __forceinline__ cplx A(int row, int k) const {
    if (row >= M || k >= K) return cmake(0.0, 0.0);
    if (!kickL) return Bi[(size_t)row * chiM + k];
    const size_t base = (size_t)(row & ~1) * chiM + k;
    const cplx x0 = Bi[base], x1 = Bi[base + chiM];
    cplx y = cmul((row & 1) ? k10 : k00, x0);
    cfma(y, (row & 1) ? k11 : k01, x1);
    return y;
  }
  __device__ __forceinline__ cplx B(int col, int k) const {
    if (col >= N || k >= K) return cmake(0.0, 0.0);
    const int p = col >= chiR, bb = col - p * chiR;
    if (!kickR) return Bn[(size_t)(2 * k + p) * chiR + bb];
    const size_t base = (size_t)(2 * k) * chiR + bb;
    const cplx x0 = Bn[base], x1 = Bn[base + chiR];
    cplx y = cmul(p ? k10 : k00, x0);
    cfma(y, p ? k11 : k01, x1);
    return y;
  }
  __device__ __forceinline__ void store(int row, int col, cplx v) const {
    const size_t o = (size_t)row * N + col;
    if (diag) {
      const int p0 = row & 1, p1 = col >= chiR;
      const cplx ph = p0 ? (p1 ? g3 : g2) : (p1 ? g1 : g0);
      const cplx c = cmul(ph, v);
      C[o] = c;
      X[(size_t)row * N + 2 * (col - p1 * chiR) + p1] = cscale(c, S[row >> 1]);  // interleaved columns
    } else {
      C[o] = v;
    }
  }
};

// general 4x4 gate on the raw two-site tensor, then theta = S_i C.  grid (ceil(chiL*chiR/256), nb, nr)
__global__ void __launch_bounds__(256) gate_mix_kernel(TcDev d, LayerArgs a) {
  Bond b;
  if (!get_bond(d, a, blockIdx.y, blockIdx.z, b)) return;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= b.chiL * b.chiR) return;
  const int al = e / b.chiR, be = e - al * b.chiR;
  const cplx *g = a.gate_override ? a.gate_override : d.gates + ((size_t)b.r * (d.L - 1) + b.i) * 16;
  cplx *C = d.Cw + b.slot * d.slot_stride, *X = d.Xw + b.slot * d.slot_stride;
  const double s = S_ptr(d, b.r, b.i)[al];
  size_t o[4];
  cplx t[4];
  for (int q = 0; q < 4; ++q) {
    o[q] = (size_t)(2 * al + (q >> 1)) * b.N + (q & 1) * b.chiR + be;
    t[q] = C[o[q]];
  }
  for (int p = 0; p < 4; ++p) {
    cplx acc = cmake(0.0, 0.0);
    for (int q = 0; q < 4; ++q) cfma(acc, g[p * 4 + q], t[q]);
    C[o[p]] = acc;
    X[(size_t)(2 * al + (p >> 1)) * b.N + 2 * be + (p & 1)] = cscale(acc, s);  // interleaved columns
  }
}

// K3: B_i[(a,p0),k] = sum_col C[(a,p0),col] conj(B_{i+1}[k,col]) / renorm
struct BLPolicy {
  int M, N, K;
  const cplx *C, *Bn;
  cplx *Bi;
  double inv;
  static constexpr bool A_KCONTIG = true, B_KCONTIG = true;
  __device__ bool init(const TcDev &d, const LayerArgs &a, int jb, int ry) {
    Bond b;
    if (!get_bond(d, a, jb, ry, b)) return false;
    M = b.M;
    N = b.chiM;  // already the truncated bond dimension
    K = b.N;
    C = d.Cw + b.slot * d.slot_stride;
    Bn = site_ptr(d, b.r, b.i + 1);
    Bi = site_ptr(d, b.r, b.i);
    inv = 1.0 / d.renorm[b.slot];
    return true;
  }
  __device__ __forceinline__ cplx A(int row, int k) const {
    return (row < M && k < K) ? C[(size_t)row * K + k] : cmake(0.0, 0.0);
  }
  __device__ __forceinline__ cplx B(int col, int k) const {
    return (col < N && k < K) ? cconj(Bn[(size_t)col * K + k]) : cmake(0.0, 0.0);
  }
  __device__ __forceinline__ void store(int row, int col, cplx v) const {
    Bi[(size_t)row * N + col] = cscale(v, inv);
  }
};

// ------------------------------------------------------------------------------------------------
// FP64 throughput probes (roofline denominator)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) probe_fma_kernel(double *out, int iters) {
  double a[8];
  const double x = 1.0000001, y = 1e-9 * threadIdx.x;
  for (int k = 0; k < 8; ++k) a[k] = k + y;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = fma(a[k], x, y);
  double s = 0.0;
  for (int k = 0; k < 8; ++k) s += a[k];
  if (s == 123.456) out[0] = s;
}
__global__ void __launch_bounds__(256) probe_dmma_kernel(double *out, int iters) {
  double c[8][2];
  const double x = 1.0 + 1e-9 * threadIdx.x, y = 1e-9;
  for (int k = 0; k < 8; ++k) c[k][0] = c[k][1] = 0.0;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int k = 0; k < 8; ++k) tcg::dmma(c[k][0], c[k][1], x, y);
  double s = 0.0;
  for (int k = 0; k < 8; ++k) s += c[k][0] + c[k][1];
  if (s == 123.456) out[0] = s;
}

// ------------------------------------------------------------------------------------------------
// dynamic shared-memory limits of the kernels that need more than 48 KB: set once per device to the largest size any
// context can ask for.  (The attribute is the allowed MAXIMUM and belongs to the function, not to a context: setting it
// per context to that context's own size lets a small context lower it under a large one created earlier.)
// ------------------------------------------------------------------------------------------------
static constexpr size_t OV_SMEM_MAX = (2 * (size_t)1024 + (tco::NT_OV / 32) * tco::OVC) * sizeof(cplx);  // chi_cap <= 1024
static size_t blocked_smem(int n2, int br) {
  return (size_t)3 * br * n2 * sizeof(cplx) + (size_t)n2 * sizeof(double2) + 64 + 2 * br * sizeof(int);
}
static int ensure_kernel_attributes(int device) {
  static std::atomic<unsigned long long> done{0};
  if (device < 64 && (done.load() >> device) & 1ull) return 0;
  CK(cudaFuncSetAttribute(tcj::qr_blocked_kernel<tcj::QNT, tcj::QBDEF>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          tcj::QNT * tcj::QBDEF * (int)sizeof(cplx)));
  CK(cudaFuncSetAttribute(tcj::qr_blocked_kernel<tcj::QNTN, tcj::QBDEF>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          tcj::QNTN * tcj::QBDEF * (int)sizeof(cplx)));
  CK(cudaFuncSetAttribute(tcj::qr_blocked_kernel<tcj::QNTW, tcj::QBDEF>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          tcj::QNTW * tcj::QBDEF * (int)sizeof(cplx)));
  CK(cudaFuncSetAttribute(tct::jacobi_team_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tct::smem_bytes(512)));
  CK(cudaFuncSetAttribute(tct::jacobi_team_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tct::smem_bytes(256)));
  CK(cudaFuncSetAttribute(tcb::jacobi_blocked_kernel<8, tcb::BR_WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          (int)blocked_smem(tcb::MAX_N, tcb::BR_WIDE)));
  CK(cudaFuncSetAttribute(tcb::jacobi_blocked_kernel<4, tcb::BR_NARROW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          (int)blocked_smem(tcb::MAX_N_NARROW, tcb::BR_NARROW)));
  CK(cudaFuncSetAttribute(tchw::jacobi_halfwarp_kernel<tchw::NW_MANY>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          (int)tchw::smem_bytes(tchw::MAX_N, tchw::NW_MANY)));
  CK(cudaFuncSetAttribute(tchw::jacobi_halfwarp_kernel<tchw::NW_ONE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          (int)tchw::smem_bytes(tchw::MAX_N, tchw::NW_ONE)));
  CK(cudaFuncSetAttribute(tco::overlap_product_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)OV_SMEM_MAX));
  if (device < 64) done.fetch_or(1ull << device);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------------
static int check_ctx(tc_ctx *c) {
  if (!c) return fail("null context");
  CK(cudaSetDevice(c->device));
  return 0;
}
#define CTX(c)                 \
  do {                         \
    if (check_ctx(c)) return 1; \
  } while (0)

static bool gates_all_diagonal(const double *g, size_t n_gates) {
  for (size_t k = 0; k < n_gates; ++k)
    for (int p = 0; p < 4; ++p)
      for (int q = 0; q < 4; ++q)
        if (p != q && (g[(k * 16 + p * 4 + q) * 2] != 0.0 || g[(k * 16 + p * 4 + q) * 2 + 1] != 0.0)) return false;
  return true;
}

// one group of two-site updates: bonds first_site + 2 jb, jb < nb, on chains [r_lo, r_hi)
// `st`, `ws_lo`, `ws_n`: the stream to launch on and the range of workspace chain slots this call may use
static int run_bonds(tc_ctx *c, int first_site, int nb, int r_lo, int r_hi, int kick_mode, const cplx *gate_override,
                     int diag, cudaStream_t st, int ws_lo, int ws_n) {
  if (nb <= 0 || r_hi <= r_lo) return 0;
  const TcDev &d = c->d;
  if (d.ws_chains < 1 || ws_n < 1)
    return fail("storage-only context (tc_ctx_create2): no SVD workspace, copy the chain into a full context to evolve it");
  const int tiles1 = (d.n2 + tcg::BM - 1) / tcg::BM;
  const int tiles = tiles1 * tiles1;
  for (int r0 = r_lo; r0 < r_hi; r0 += ws_n) {
    const int nr = (r_hi - r0) < ws_n ? (r_hi - r0) : ws_n;
    LayerArgs a{first_site, 2, nb, r0, nr, kick_mode, gate_override, diag, ws_lo};
    {
      ProfScope ps(c, TC_PROF_THETA);
      if (c->old_theta) {
        tcg::gemm_kernel<ThetaPolicy><<<dim3(tiles, nb, nr), tcg::NT, 0, st>>>(d, a);
      } else {
        const int tm = (d.n2 + tch::BM - 1) / tch::BM, tn = (d.chi_cap + tch::BNB - 1) / tch::BNB;
        tch::theta_gemm_kernel<<<dim3(tm * tn, nb, nr), tch::NT, 0, st>>>(d, a);
      }
      LAUNCHED();
      if (!diag) {
        const int gx = (d.chi_cap * d.chi_cap + 255) / 256;
        gate_mix_kernel<<<dim3(gx, nb, nr), 256, 0, st>>>(d, a);
        LAUNCHED();
      }
    }
    {
      ProfScope ps(c, TC_PROF_QR);
      // a cluster of CS CTAs per matrix when the launch has too few matrices to fill the GPU (a single chain, config 4)
      int CS = 1;
      while (CS < 8 && (long long)nb * nr * CS * 2 <= (long long)c->sm_count) CS *= 2;
      if (c->qr_cluster > 0) CS = c->qr_cluster < 8 ? c->qr_cluster : 8;  // 8 = the portable cluster size limit
      auto launch_qr = [&](auto kern, int threads) -> int {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(nr * CS, nb);
        cfg.blockDim = dim3(threads);
        cfg.dynamicSmemBytes = (size_t)threads * tcj::QBDEF * sizeof(cplx);  // V panel
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = CS;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        CK(cudaLaunchKernelEx(&cfg, kern, d, a, CS));
        return 0;
      };
      if (d.n2 <= tcj::QMAXMN && !c->force_simple_jacobi && !c->no_narrow_qr && (long long)d.R * (d.L / 2) > c->sm_count) {
        // ensembles of narrow contexts: one thread per row still, 4 warps per CTA, more CTAs per SM; the same arithmetic
        if (launch_qr(tcj::qr_blocked_kernel<tcj::QNTN, tcj::QBDEF>, tcj::QNTN)) return 1;
      } else if (d.n2 <= tcj::QMAXM && !c->force_simple_jacobi) {
        if (launch_qr(tcj::qr_blocked_kernel<tcj::QNT, tcj::QBDEF>, tcj::QNT)) return 1;
      } else if (d.n2 <= tcj::QMAXMW && !c->force_simple_jacobi) {
        if (launch_qr(tcj::qr_blocked_kernel<tcj::QNTW, tcj::QBDEF>, tcj::QNTW)) return 1;
      } else {
        tcj::qr_kernel<<<dim3(nb, nr), tcj::NT, (d.n2 + 64) * sizeof(cplx), st>>>(d, a);
      }
      LAUNCHED();
    }
    {
      ProfScope ps(c, TC_PROF_JACOBI);
      auto launch_team = [&](int maxnpl) -> int {
        // a cluster of CS CTAs per matrix when the launch has too few matrices to fill the GPU
        const int per_sm = maxnpl <= 4 ? 2 : 1;
        int CS = 1;
        // from the context's shape, not from the chain group being launched: the cluster size fixes the pair order, and a
        // context's results must not depend on TC_GROUPS / TC_WS_BYTES
        while (CS < 8 && (long long)d.R * (d.L / 2) * CS * 2 <= (long long)c->sm_count * per_sm) CS *= 2;
        if (c->wide_cluster > 0) CS = c->wide_cluster;
        const size_t smem = tct::smem_bytes(d.n2);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(nb * CS, nr);
        cfg.blockDim = dim3(tct::NT);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = CS;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        if (maxnpl <= 4)
          CK(cudaLaunchKernelEx(&cfg, tct::jacobi_team_kernel<4>, d, a, CS));
        else
          CK(cudaLaunchKernelEx(&cfg, tct::jacobi_team_kernel<8>, d, a, CS));
        return 0;
      };
      if (d.n2 <= tcb::MAX_N && c->team_jacobi) {
        if (launch_team(4)) return 1;
      } else if (d.n2 > tcb::MAX_N && d.n2 <= 512 && !c->force_simple_jacobi && !c->old_wide) {
        if (launch_team(8)) return 1;
      } else if (d.n2 <= tcb::MAX_N && !c->force_simple_jacobi) {
        // Narrow contexts (widest matrix 128 columns) run the half-warp-row kernel (tc_jacobi_halfwarp.cuh): with more
        // matrices per layer than SMs its 8-warp instance, two CTAs per SM (throughput); otherwise (a single chain) the
        // 16-warp instance, 32 half-warps on one matrix (Jacobi of one period at L = 16, chi = 64: 10.9 ms against 12.5 ms
        // on full-warp rows and 15.0 ms on the 8-warp instance).  The choice depends on the context's shape only (not on
        // the chain group or chunk being launched), so a context's results do not depend on TC_GROUPS / TC_WS_BYTES.
        const bool narrow = d.n2 <= tchw::MAX_N && !c->no_small_kernel;
        const bool many = (long long)d.R * (d.L / 2) > c->sm_count;
        const bool automatic = !c->force_small_kernel && !c->halfwarp_kernel;
        if (narrow && (c->halfwarp_kernel == 3 || (automatic && many)))
          tchw::jacobi_halfwarp_kernel<tchw::NW_MANY>  // rows on half-warps, 8 warps, two CTAs per SM: throughput
              <<<dim3(nr, nb), tchw::NW_MANY * 32, tchw::smem_bytes(d.n2, tchw::NW_MANY), st>>>(d, a);
        else if (narrow && (c->halfwarp_kernel == 4 || automatic))
          tchw::jacobi_halfwarp_kernel<tchw::NW_ONE>  // 16 warps = 32 half-warps on one matrix: latency of a single chain
              <<<dim3(nr, nb), tchw::NW_ONE * 32, tchw::smem_bytes(d.n2, tchw::NW_ONE), st>>>(d, a);
        else if (narrow && c->force_small_kernel)  // TC_SMALL_KERNEL=2 (A/B): full-warp rows in blocks of 8
          tcb::jacobi_blocked_kernel<4, tcb::BR_NARROW>
              <<<dim3(nr, nb), tcb::BR_NARROW * 32, blocked_smem(d.n2, tcb::BR_NARROW), st>>>(d, a);
        else
          tcb::jacobi_blocked_kernel<8, tcb::BR_WIDE>
              <<<dim3(nr, nb), tcb::BR_WIDE * 32, blocked_smem(d.n2, tcb::BR_WIDE), st>>>(d, a);
      } else {
        // wide matrices: a cluster of CS CTAs per matrix when the launch has too few matrices to fill the GPU
        int CS = 1;
        while (CS < 8 && (long long)nb * nr * CS * 2 <= c->sm_count) CS *= 2;  // 190 registers: one CTA per SM
        if (c->wide_cluster > 0) CS = c->wide_cluster;
        if (CS > 1 && !c->force_simple_jacobi && d.n2 <= 32 * tcj::WNPL) {
          cudaLaunchConfig_t cfg = {};
          cfg.gridDim = dim3(nb * CS, nr);
          cfg.blockDim = dim3(tcj::NT);
          cfg.dynamicSmemBytes = 0;
          cfg.stream = st;
          cudaLaunchAttribute attr[1];
          attr[0].id = cudaLaunchAttributeClusterDimension;
          attr[0].val.clusterDim.x = CS;
          attr[0].val.clusterDim.y = 1;
          attr[0].val.clusterDim.z = 1;
          cfg.attrs = attr;
          cfg.numAttrs = 1;
          CK(cudaLaunchKernelEx(&cfg, tcj::jacobi_rows_cluster_kernel, d, a, CS));
        } else {
          tcj::jacobi_rows_kernel<<<dim3(nb, nr), tcj::NT, d.n2 * sizeof(double), st>>>(d, a);
        }
      }
      LAUNCHED();
    }
    {
      ProfScope ps(c, TC_PROF_FINALIZE);
      tcj::finalize_kernel<<<dim3(nb, nr), tcj::NT, d.n2 * sizeof(double), st>>>(d, a);
      LAUNCHED();
    }
    {
      ProfScope ps(c, TC_PROF_BLEFT);
      tcg::gemm_kernel<BLPolicy><<<dim3(tiles, nb, nr), tcg::NT, 0, st>>>(d, a);
      LAUNCHED();
    }
  }
  return 0;
}

static int run_layer_on(tc_ctx *c, int parity, int kick_mode, int r_lo, int r_hi, cudaStream_t st, int ws_lo, int ws_n) {
  const int L = c->d.L;
  if (L < 2) return 0;
  const int nb = (L - 1 - parity + 1) / 2;  // bonds parity, parity+2, ... <= L-2
  return run_bonds(c, parity, nb, r_lo, r_hi, kick_mode, nullptr, c->d.gates_diag, st, ws_lo, ws_n);
}
static int run_layer(tc_ctx *c, int parity, int kick_mode) {
  return run_layer_on(c, parity, kick_mode, 0, c->d.R, c->stream, 0, c->d.ws_chains);
}

static int run_kick_all(tc_ctx *c) {
  ProfScope ps(c, TC_PROF_KICK);
  tco::one_site_kernel<<<dim3(c->d.L, c->d.R), 256, 0, c->stream>>>(c->d, 0, 0, nullptr);
  LAUNCHED();
  return 0;
}

// even, odd, kick (fused into the loads of the second Ising layer), even, odd -- for the chains [r_lo, r_hi)
static int run_period_on(tc_ctx *c, int r_lo, int r_hi, cudaStream_t st, int ws_lo, int ws_n) {
  const int L = c->d.L;
  if (run_layer_on(c, 0, 0, r_lo, r_hi, st, ws_lo, ws_n)) return 1;
  if (run_layer_on(c, 1, 0, r_lo, r_hi, st, ws_lo, ws_n)) return 1;
  if (run_layer_on(c, 0, 1, r_lo, r_hi, st, ws_lo, ws_n)) return 1;  // the even bonds cover sites 0 .. 2*floor(L/2)-1
  if (L == 2) return 0;
  // odd L: site L-1 is the right site of the last odd bond
  return run_layer_on(c, 1, (L & 1) ? 2 : 0, r_lo, r_hi, st, ws_lo, ws_n);
}

// n Floquet periods of every chain.  With G > 1 chain groups each group runs all its n periods on its own stream
// (fork after what is already queued on the context's stream, join before what comes next); per-kernel profiling
// needs the launches in one stream and runs the groups one after the other.
// after(t, r_lo, r_hi, st) is called once per period t and chain range, on the stream that ran that range: the records
// of tc_floquet_run_dev are taken there, group by group, so that a record does not make the groups wait for each
// other (a join per period costs 2 % at the metric shape: the slowest group's last layer runs alone).
template <class After>
static int run_periods_with(tc_ctx *c, int n, After &&after) {
  const TcDev &d = c->d;
  if (n <= 0) return 0;
  if (d.L < 2) {
    for (int t = 0; t < n; ++t) {
      if (run_kick_all(c)) return 1;
      if (after(t, 0, d.R, c->stream)) return 1;
    }
    return 0;
  }
  int G = c->ngroups;
  if (G > d.R) G = d.R;
  if (G > d.ws_chains) G = d.ws_chains;
  if (d.ws_chains < 1)
    return fail("storage-only context (tc_ctx_create2): no SVD workspace, copy the chain into a full context to evolve it");
  if (c->profile || G <= 1) {
    for (int t = 0; t < n; ++t) {
      if (run_period_on(c, 0, d.R, c->stream, 0, d.ws_chains)) return 1;
      if (after(t, 0, d.R, c->stream)) return 1;
    }
    return 0;
  }
  while ((int)c->gstreams.size() < G) {
    cudaStream_t s;
    CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    c->gstreams.push_back(s);
    cudaEvent_t e;
    CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    c->gjoin.push_back(e);
  }
  if (!c->gfork) CK(cudaEventCreateWithFlags(&c->gfork, cudaEventDisableTiming));
  CK(cudaEventRecord(c->gfork, c->stream));
  // workspace slots of group g: its own chains when the workspace holds every chain (the usual case; an even split
  // would leave the larger groups of an uneven partition one slot short and make them run a second, almost empty
  // chunk per layer), an even share otherwise
  const bool whole = d.ws_chains >= d.R;
  const int wpg = d.ws_chains / G;
  for (int g = 0; g < G; ++g) CK(cudaStreamWaitEvent(c->gstreams[g], c->gfork, 0));
  // period-major enqueue order: a stream takes only so many pending launches before the host blocks; feeding one
  // group all its periods first would let that group run alone for long calls
  int rc = 0;
  for (int t = 0; t < n && !rc; ++t)
    for (int g = 0; g < G && !rc; ++g) {
      const int r_lo = (int)((long long)d.R * g / G), r_hi = (int)((long long)d.R * (g + 1) / G);
      rc = run_period_on(c, r_lo, r_hi, c->gstreams[g], whole ? r_lo : g * wpg, whole ? r_hi - r_lo : wpg);
      if (!rc) rc = after(t, r_lo, r_hi, c->gstreams[g]);
    }
  // join on the error path too: what was queued on the group streams still uses the arena, and the caller may free
  // it as soon as the context's stream is idle
  for (int g = 0; g < G; ++g) {
    CK(cudaEventRecord(c->gjoin[g], c->gstreams[g]));
    CK(cudaStreamWaitEvent(c->stream, c->gjoin[g], 0));
  }
  return rc;
}
static int run_periods(tc_ctx *c, int n) {
  return run_periods_with(c, n, [](int, int, int, cudaStream_t) { return 0; });
}

// observables of the chains [r_lo, r_hi) on stream st; the output arrays are indexed by the absolute chain number
static int measure_range(tc_ctx *c, double *rdm, double *Z, double *ent, double *ov, int32_t *chi, int r_lo, int r_hi,
                         cudaStream_t st) {
  const TcDev &d = c->d;
  const int nr = r_hi - r_lo;
  if (nr <= 0) return 0;
  if (rdm || Z || ent) {
    tco::measure_kernel<<<dim3(d.L, nr), tco::NT, 0, st>>>(d, r_lo, rdm, Z, ent);
    LAUNCHED();
  }
  if (ov) {
    // (running this latency-bound chain on a side stream next to the bandwidth-bound measure kernel was measured:
    // 0.170 -> 0.167 ms per snapshot, not worth the extra streams and events)
    const size_t smem = (2 * (size_t)d.chi_cap + (tco::NT_OV / 32) * tco::OVC) * sizeof(cplx);
    tco::overlap_product_kernel<<<dim3(nr, 2), tco::NT_OV, smem, st>>>(d, r_lo, ov, c->ovs, c->ovc);
    LAUNCHED();
  }
  if (chi) {
    tco::chi_record_kernel<<<64, 256, 0, st>>>(d, r_lo, nr, chi);
    LAUNCHED();
  }
  return 0;
}
static int measure_into(tc_ctx *c, double *rdm, double *Z, double *ent, double *ov, int32_t *chi) {
  ProfScope ps(c, TC_PROF_MEASURE);
  return measure_range(c, rdm, Z, ent, ov, chi, 0, c->d.R, c->stream);
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int tc_version(void) { return 100; }
const char *tc_last_error(void) { return g_err.c_str(); }
int tc_device_count(int *count) {
  CK(cudaGetDeviceCount(count));
  return 0;
}
long long tc_launch_count(void) { return g_launches.load(); }

size_t tc_ctx_arena_bytes2(int L, int chi_cap, int R, int storage_only) {
  if (L < 1 || chi_cap < 1 || R < 1) return 0;
  return make_layout(L, chi_cap, R, storage_only != 0).total;
}
size_t tc_ctx_arena_bytes(int L, int chi_cap, int R) { return tc_ctx_arena_bytes2(L, chi_cap, R, 0); }

int tc_ctx_create(int device, int L, int chi_cap, int R, void *arena, size_t arena_bytes, void *stream, tc_ctx **out) {
  return tc_ctx_create2(device, L, chi_cap, R, 0, arena, arena_bytes, stream, out);
}

int tc_ctx_create2(int device, int L, int chi_cap, int R, int storage_only, void *arena, size_t arena_bytes,
                   void *stream, tc_ctx **out) {
  if (!out) return fail("tc_ctx_create: out is null");
  if (L < 1 || chi_cap < 1 || R < 1) return fail("tc_ctx_create: L, chi_cap, R must be >= 1");
  if (R > 65535) return fail("tc_ctx_create: R > 65535 chains per context");
  CK(cudaSetDevice(device));
  if (ensure_kernel_attributes(device)) return 1;
  if (chi_cap > 1024) return fail("tc_ctx_create: chi_cap > 1024");
  Layout lo = make_layout(L, chi_cap, R, storage_only != 0);
  tc_ctx *c = new tc_ctx();
  c->device = device;
  if (arena) {
    if (arena_bytes < lo.total) {
      delete c;
      return fail("tc_ctx_create: arena too small");
    }
    if (((uintptr_t)arena & 255) != 0) {
      delete c;
      return fail("tc_ctx_create: arena must be 256-byte aligned");
    }
    c->arena = arena;
  } else {
    cudaError_t e = cudaMalloc(&c->arena, lo.total);
    if (e != cudaSuccess) {
      delete c;
      return fail("tc_ctx_create: cudaMalloc failed: %s", cudaGetErrorString(e));
    }
    c->own_arena = true;
  }
  c->arena_bytes = lo.total;
  c->ngroups = 4;
  if (const char *e = getenv("TC_SMALL_KERNEL")) {
    c->no_small_kernel = atoi(e) == 0;
    c->force_small_kernel = atoi(e) == 2;
    c->halfwarp_kernel = (atoi(e) == 3 || atoi(e) == 4) ? atoi(e) : 0;
  }
  cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, c->device);
  if (const char *e = getenv("TC_WIDE_CLUSTER")) c->wide_cluster = atoi(e);
  if (const char *e = getenv("TC_QR_CLUSTER")) c->qr_cluster = atoi(e);
  if (const char *e = getenv("TC_QR_NARROW")) c->no_narrow_qr = atoi(e) == 0;
  if (const char *e = getenv("TC_THETA")) c->old_theta = strcmp(e, "v1") == 0;
  if (const char *e = getenv("TC_GROUPS")) c->ngroups = atoi(e) > 0 ? atoi(e) : 1;
  if (const char *e = getenv("TC_JACOBI")) {
    c->force_simple_jacobi = strcmp(e, "simple") == 0;
    c->team_jacobi = strcmp(e, "team") == 0;
    c->old_wide = strcmp(e, "wide_v1") == 0;
  }
  if (stream) {
    c->stream = (cudaStream_t)stream;
  } else {
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
      if (c->own_arena) cudaFree(c->arena);
      delete c;
      return fail("tc_ctx_create: stream: %s", cudaGetErrorString(e));
    }
    c->own_stream = true;
  }
  char *base = (char *)c->arena;
  TcDev &d = c->d;
  d.L = L;
  d.chi_cap = chi_cap;
  d.R = R;
  d.n2 = lo.n2;
  d.nbmax = lo.nbmax;
  d.ws_chains = lo.ws_chains;
  d.site_stride = (size_t)chi_cap * 2 * chi_cap;
  d.slot_stride = (size_t)lo.n2 * lo.n2;
  d.B = (cplx *)(base + lo.B);
  d.S = (double *)(base + lo.S);
  d.chi = (int *)(base + lo.chi);
  d.init_idx = (int8_t *)(base + lo.init_idx);
  c->gates_dev = (cplx *)(base + lo.gates);
  c->kick_dev = (cplx *)(base + lo.kick);
  d.gates = c->gates_dev;
  d.kick = c->kick_dev;
  d.gates_diag = 0;
  d.rot64 = 1;  // bit 0: unused (the FP32 angle variant is gone); bit 1: lock-step rounds (barrier instead of hand-over)
  if (const char *e = getenv("TC_ROT64")) d.rot64 = atoi(e);
  {  // threshold schedule of the Jacobi sweeps 0..3; TC_THRESH=0 switches it off, TC_THRESH=a,b,c,d sets it (A/B)
    // (round 1, standard rotations: 1e-2, 1e-3, 1e-4, 1e-6; re-tuned on the B200 after the fast rotations, r02 sweep in
    // profiles/README.md: 204.6 -> 211.7 chain-steps/s, mean sweeps 8.27 -> 7.97)
    double sched[6] = {3e-3, 3e-4, 3e-5, 3e-6, 0.0, 0.0};
    if (const char *e = getenv("TC_THRESH")) {
      double v[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
      sscanf(e, "%lf,%lf,%lf,%lf,%lf,%lf", &v[0], &v[1], &v[2], &v[3], &v[4], &v[5]);
      for (int k = 0; k < 6; ++k) sched[k] = v[k];
    }
    for (int k = 0; k < 6; ++k) d.thr_sched[k] = sched[k];
  }
  d.small_rel2 = tcj::SMALL_REL2;
  if (const char *e = getenv("TC_EARLY_STOP"))
    if (atoi(e) == 0) d.small_rel2 = 0.0;  // A/B: run until a sweep rotates nothing (the verification sweep)
  d.trunc_err = (double *)(base + lo.trunc_err);
  d.flags = (int *)(base + lo.flags);
  d.Cw = (cplx *)(base + lo.Cw);
  d.Xw = (cplx *)(base + lo.Xw);
  d.ww = (double *)(base + lo.ww);
  d.perm = (int *)(base + lo.perm);
  d.knew = (int *)(base + lo.knew);
  d.renorm = (double *)(base + lo.renorm);
  c->op_scratch = (cplx *)(base + lo.op);
  c->E0 = (cplx *)(base + lo.E0);
  c->E1 = (cplx *)(base + lo.E1);
  c->Tt = (cplx *)(base + lo.Tt);
  c->small_out = (double *)(base + lo.small_out);
  c->ovs = (cplx *)(base + lo.ovs);
  c->ovc = (int *)(base + lo.ovc);
  cudaMemsetAsync(c->ovc, 0, (size_t)R * sizeof(int), c->stream);
  d.mode = TC_TRUNC_REFERENCE;
  d.cutoff = 1e-13;
  d.chi_max = 0;
  d.svd_min = 0.0;
  d.trunc_cut = 0.0;
  cudaMemsetAsync(d.flags, 0, 8 * sizeof(int), c->stream);
  cudaMemsetAsync(d.trunc_err, 0, (size_t)R * (L + 1) * sizeof(double), c->stream);
  cudaMemsetAsync(d.init_idx, 0, (size_t)R * L, c->stream);
  // default state: |0...0>
  tco::product_state_kernel<<<R, 64, 0, c->stream>>>(d);
  g_launches.fetch_add(1);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    tc_ctx_destroy(c);
    return fail("tc_ctx_create: init kernel: %s", cudaGetErrorString(e));
  }
  *out = c;
  return 0;
}

int tc_ctx_destroy(tc_ctx *c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  for (auto &sp : c->spans) {
    cudaEventDestroy(sp.e0);
    cudaEventDestroy(sp.e1);
  }
  for (auto e : c->free_events) cudaEventDestroy(e);
  if (c->rec) cudaFree(c->rec);
  if (c->own_arena) cudaFree(c->arena);
  for (cudaStream_t gs : c->gstreams) cudaStreamDestroy(gs);
  for (cudaEvent_t e : c->gjoin) cudaEventDestroy(e);
  if (c->gfork) cudaEventDestroy(c->gfork);
  if (c->own_stream) cudaStreamDestroy(c->stream);
  delete c;
  return 0;
}

int tc_sync(tc_ctx *c) {
  CTX(c);
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}

int tc_ctx_info(tc_ctx *c, int *L, int *chi_cap, int *R, int *device) {
  if (!c) return fail("null context");
  if (L) *L = c->d.L;
  if (chi_cap) *chi_cap = c->d.chi_cap;
  if (R) *R = c->d.R;
  if (device) *device = c->device;
  return 0;
}

int tc_get_flags(tc_ctx *c, int32_t *out4) {
  CTX(c);
  int tmp[8];
  CK(cudaMemcpyAsync(tmp, c->d.flags, 8 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  out4[0] = tmp[0];
  out4[1] = tmp[1];
  out4[2] = tmp[2];
  out4[3] = tmp[4] > 0 ? (int)(100.0 * tmp[3] / tmp[4]) : 0;  // mean sweeps x 100 over the SVDs with >= 128 rows
  return 0;
}
int tc_profile(tc_ctx *c, int enable) {
  if (!c) return fail("null context");
  c->profile = enable != 0;
  return 0;
}

int tc_profile_read(tc_ctx *c, double *ms_out, long long *count_out, int reset) {
  CTX(c);
  CK(cudaStreamSynchronize(c->stream));
  for (int k = 0; k < TC_PROF_NCLASS; ++k) {
    ms_out[k] = 0.0;
    if (count_out) count_out[k] = 0;
  }
  for (auto &sp : c->spans) {
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, sp.e0, sp.e1));
    ms_out[sp.cls] += ms;
    if (count_out) count_out[sp.cls] += 1;
  }
  if (reset) {
    for (auto &sp : c->spans) {
      c->free_events.push_back(sp.e0);
      c->free_events.push_back(sp.e1);
    }
    c->spans.clear();
  }
  return 0;
}

// ---- state ----------------------------------------------------------------------------------
int tc_set_product_state(tc_ctx *c, const int8_t *idx_host) {
  CTX(c);
  const size_t n = (size_t)c->d.R * c->d.L;
  for (size_t k = 0; k < n; ++k)
    if (idx_host[k] != 0 && idx_host[k] != 1) return fail("tc_set_product_state: basis index must be 0 or 1");
  CK(cudaMemcpyAsync(c->d.init_idx, idx_host, n, cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));  // idx_host may be pageable and reused by the caller
  tco::product_state_kernel<<<c->d.R, 64, 0, c->stream>>>(c->d);
  LAUNCHED();
  return 0;
}

static int read_chi(tc_ctx *c, int r, std::vector<int> &chi) {
  chi.resize(c->d.L + 1);
  CK(cudaMemcpyAsync(chi.data(), c->d.chi + (size_t)r * (c->d.L + 1), chi.size() * sizeof(int),
                     cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}

int tc_set_site(tc_ctx *c, int r, int site, const double *data_host, int chi_l, int chi_r) {
  CTX(c);
  const TcDev &d = c->d;
  if (r < 0 || r >= d.R || site < 0 || site >= d.L) return fail("tc_set_site: index out of range");
  if (chi_l < 1 || chi_r < 1 || chi_l > d.chi_cap || chi_r > d.chi_cap) return fail("tc_set_site: chi out of range");
  cplx *B = d.B + ((size_t)r * d.L + site) * d.site_stride;
  CK(cudaMemcpyAsync(B, data_host, (size_t)chi_l * 2 * chi_r * sizeof(cplx), cudaMemcpyHostToDevice, c->stream));
  int dims[2] = {chi_l, chi_r};
  CK(cudaMemcpyAsync(d.chi + (size_t)r * (d.L + 1) + site, dims, 2 * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}

int tc_get_site(tc_ctx *c, int r, int site, double *out_host, int *chi_l, int *chi_r) {
  CTX(c);
  const TcDev &d = c->d;
  if (r < 0 || r >= d.R || site < 0 || site >= d.L) return fail("tc_get_site: index out of range");
  std::vector<int> chi;
  if (read_chi(c, r, chi)) return 1;
  if (chi_l) *chi_l = chi[site];
  if (chi_r) *chi_r = chi[site + 1];
  if (out_host) {
    const cplx *B = d.B + ((size_t)r * d.L + site) * d.site_stride;
    CK(cudaMemcpyAsync(out_host, B, (size_t)chi[site] * 2 * chi[site + 1] * sizeof(cplx), cudaMemcpyDeviceToHost,
                       c->stream));
    CK(cudaStreamSynchronize(c->stream));
  }
  return 0;
}

int tc_set_S(tc_ctx *c, int r, int bond, const double *S_host, int n) {
  CTX(c);
  const TcDev &d = c->d;
  if (r < 0 || r >= d.R || bond < 0 || bond > d.L) return fail("tc_set_S: index out of range");
  if (n < 1 || n > d.chi_cap) return fail("tc_set_S: n out of range");
  CK(cudaMemcpyAsync(d.S + ((size_t)r * (d.L + 1) + bond) * d.chi_cap, S_host, n * sizeof(double),
                     cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(d.chi + (size_t)r * (d.L + 1) + bond, &n, sizeof(int), cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}

int tc_get_S(tc_ctx *c, int r, int bond, double *out_host, int *n) {
  CTX(c);
  const TcDev &d = c->d;
  if (r < 0 || r >= d.R || bond < 0 || bond > d.L) return fail("tc_get_S: index out of range");
  std::vector<int> chi;
  if (read_chi(c, r, chi)) return 1;
  if (n) *n = chi[bond];
  if (out_host) {
    CK(cudaMemcpyAsync(out_host, d.S + ((size_t)r * (d.L + 1) + bond) * d.chi_cap, chi[bond] * sizeof(double),
                       cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
  }
  return 0;
}

int tc_get_chi(tc_ctx *c, int32_t *out_host) {
  CTX(c);
  CK(cudaMemcpyAsync(out_host, c->d.chi, (size_t)c->d.R * (c->d.L + 1) * sizeof(int), cudaMemcpyDeviceToHost,
                     c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}

int tc_copy_chain(tc_ctx *dst, int r_dst, tc_ctx *src, int r_src) {
  CTX(src);
  if (!dst) return fail("null context");
  if (dst->device != src->device) return fail("tc_copy_chain: contexts live on different devices");
  if (dst->d.L != src->d.L) return fail("tc_copy_chain: chain lengths differ");
  if (r_dst < 0 || r_dst >= dst->d.R || r_src < 0 || r_src >= src->d.R) return fail("tc_copy_chain: chain out of range");
  std::vector<int> chi;
  if (read_chi(src, r_src, chi)) return 1;
  for (int v : chi)
    if (v > dst->d.chi_cap) return fail("tc_copy_chain: destination chi_cap too small");
  if (dst->stream != src->stream) CK(cudaStreamSynchronize(dst->stream));
  tco::copy_chain_kernel<<<src->d.L + 1, 256, 0, src->stream>>>(dst->d, r_dst, src->d, r_src);
  LAUNCHED();
  if (dst->stream != src->stream) CK(cudaStreamSynchronize(src->stream));
  return 0;
}

int tc_get_trunc_err(tc_ctx *c, double *out_host, int reset) {
  CTX(c);
  const TcDev &d = c->d;
  std::vector<double> tmp((size_t)d.R * (d.L + 1));
  CK(cudaMemcpyAsync(tmp.data(), d.trunc_err, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (reset) CK(cudaMemsetAsync(d.trunc_err, 0, tmp.size() * sizeof(double), c->stream));
  CK(cudaStreamSynchronize(c->stream));
  for (int r = 0; r < d.R; ++r) {
    double s = 0.0;
    for (int b = 0; b <= d.L; ++b) s += tmp[(size_t)r * (d.L + 1) + b];
    out_host[r] = s;
  }
  return 0;
}

// ---- model ----------------------------------------------------------------------------------
int tc_set_model(tc_ctx *c, const double *gates_host, const double *kick_host) {
  CTX(c);
  const TcDev &d = c->d;
  if (gates_host && d.L > 1) {
    const size_t ng = (size_t)d.R * (d.L - 1);
    c->d.gates_diag = gates_all_diagonal(gates_host, ng) ? 1 : 0;
    CK(cudaMemcpyAsync(c->gates_dev, gates_host, ng * 16 * sizeof(cplx), cudaMemcpyHostToDevice, c->stream));
  }
  if (kick_host) CK(cudaMemcpyAsync(c->kick_dev, kick_host, (size_t)d.R * 4 * sizeof(cplx), cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  if (gates_host || d.L == 1) c->have_model = true;
  return 0;
}

int tc_set_trunc(tc_ctx *c, int mode, double cutoff, int chi_max, double svd_min, double trunc_cut) {
  if (!c) return fail("null context");
  if (mode != TC_TRUNC_REFERENCE && mode != TC_TRUNC_TEBD) return fail("tc_set_trunc: unknown mode");
  if (trunc_cut >= 1.0) return fail("tc_set_trunc: trunc_cut >= 1");
  c->d.mode = mode;
  c->d.cutoff = cutoff;
  c->d.chi_max = chi_max;
  c->d.svd_min = svd_min;
  c->d.trunc_cut = trunc_cut;
  return 0;
}

// ---- gates ----------------------------------------------------------------------------------
int tc_apply_layer(tc_ctx *c, int parity, int kick_mode) {
  CTX(c);
  if (!c->have_model) return fail("tc_apply_layer: no model set");
  if (parity != 0 && parity != 1) return fail("tc_apply_layer: parity must be 0 or 1");
  return run_layer(c, parity, kick_mode);
}

int tc_apply_kick(tc_ctx *c) {
  CTX(c);
  return run_kick_all(c);
}

int tc_floquet_step(tc_ctx *c, int n_steps) {
  CTX(c);
  if (!c->have_model) return fail("tc_floquet_step: no model set");
  return run_periods(c, n_steps);
}

int tc_apply_two_site(tc_ctx *c, int r, int site, const double *gate_host) {
  CTX(c);
  const TcDev &d = c->d;
  if (r < 0 || r >= d.R) return fail("tc_apply_two_site: chain out of range");
  if (site < 0 || site + 1 >= d.L) return fail("tc_apply_two_site: site out of range");
  CK(cudaMemcpyAsync(c->op_scratch, gate_host, 16 * sizeof(cplx), cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  const int diag = gates_all_diagonal(gate_host, 1) ? 1 : 0;
  return run_bonds(c, site, 1, r, r + 1, 0, c->op_scratch, diag, c->stream, 0, c->d.ws_chains);
}

int tc_apply_one_site(tc_ctx *c, int r, int site, const double *op_host) {
  CTX(c);
  const TcDev &d = c->d;
  if (r < 0 || r >= d.R || site < 0 || site >= d.L) return fail("tc_apply_one_site: index out of range");
  CK(cudaMemcpyAsync(c->op_scratch + 16, op_host, 4 * sizeof(cplx), cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  tco::one_site_kernel<<<dim3(1, 1), 256, 0, c->stream>>>(d, r, site, c->op_scratch + 16);
  LAUNCHED();
  return 0;
}

// ---- observables ----------------------------------------------------------------------------
int tc_measure_dev(tc_ctx *c, double *rdm_dev, double *ent_dev) {
  CTX(c);
  return measure_into(c, rdm_dev, nullptr, ent_dev, nullptr, nullptr);
}

static int ensure_rec(tc_ctx *c, size_t bytes) {
  if (c->rec_bytes >= bytes) return 0;
  if (c->rec) cudaFree(c->rec);
  c->rec = nullptr;
  c->rec_bytes = 0;
  CK(cudaMalloc(&c->rec, bytes));
  c->rec_bytes = bytes;
  return 0;
}

int tc_measure(tc_ctx *c, double *rdm_host, double *ent_host) {
  CTX(c);
  const TcDev &d = c->d;
  const size_t nr = (size_t)d.R * d.L * 4, ne = (size_t)d.R * (d.L > 1 ? d.L - 1 : 0);
  if (ensure_rec(c, (nr + ne + 1) * sizeof(double))) return 1;
  double *rd = (double *)c->rec, *ed = rd + nr;
  if (measure_into(c, rdm_host ? rd : nullptr, nullptr, (ent_host && ne) ? ed : nullptr, nullptr, nullptr)) return 1;
  if (rdm_host) CK(cudaMemcpyAsync(rdm_host, rd, nr * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (ent_host && ne) CK(cudaMemcpyAsync(ent_host, ed, ne * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}

static int transfer(tc_ctx *bra, int rb, tc_ctx *ket, int rk, int lo, int hi, const cplx *op_lo, const cplx *op_hi,
                    int init_mode, double *out_host) {
  if (bra->d.chi_cap > ket->d.chi_cap) {
    // scratch is taken from the ket context and sized chi_cap_ket^2; the bra may be larger
    std::vector<int> chi;
    if (read_chi(bra, rb, chi)) return 1;
    for (int v : chi)
      if (v > ket->d.chi_cap) return fail("tc_overlap: bra bond dimension exceeds the ket context's chi_cap");
  }
  if (bra->stream != ket->stream) CK(cudaStreamSynchronize(bra->stream));
  tco::transfer_kernel<<<1, tco::NT, 0, ket->stream>>>(bra->d, rb, ket->d, rk, lo, hi, op_lo, op_hi, init_mode,
                                                        ket->E0, ket->E1, ket->Tt, ket->small_out);
  LAUNCHED();
  CK(cudaMemcpyAsync(out_host, ket->small_out, 2 * sizeof(double), cudaMemcpyDeviceToHost, ket->stream));
  CK(cudaStreamSynchronize(ket->stream));
  return 0;
}

int tc_overlap(tc_ctx *bra, int r_bra, tc_ctx *ket, int r_ket, double *out_host) {
  CTX(ket);
  if (!bra) return fail("null context");
  if (bra->device != ket->device) return fail("tc_overlap: contexts live on different devices");
  if (bra->d.L != ket->d.L) return fail("tc_overlap: chain lengths differ");
  if (r_bra < 0 || r_bra >= bra->d.R || r_ket < 0 || r_ket >= ket->d.R) return fail("tc_overlap: chain out of range");
  return transfer(bra, r_bra, ket, r_ket, 0, ket->d.L - 1, nullptr, nullptr, 0, out_host);
}

int tc_correlation(tc_ctx *c, int r, int i, int j, const double *op1_host, const double *op2_host, double *out_host) {
  CTX(c);
  const TcDev &d = c->d;
  if (r < 0 || r >= d.R || i < 0 || j < 0 || i >= d.L || j >= d.L) return fail("tc_correlation: index out of range");
  const double *lo_op = op1_host, *hi_op = op2_host;
  int lo = i, hi = j;
  if (i > j) {
    lo = j;
    hi = i;
    lo_op = op2_host;
    hi_op = op1_host;
  }
  CK(cudaMemcpyAsync(c->op_scratch + 20, lo_op, 4 * sizeof(cplx), cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->op_scratch + 24, hi_op, 4 * sizeof(cplx), cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  // for i == j the kernel multiplies op_lo op_hi = op1 op2 (no swap happened)
  return transfer(c, r, c, r, lo, hi, c->op_scratch + 20, c->op_scratch + 24, 1, out_host);
}

// ---- fused time loop ------------------------------------------------------------------------
int tc_floquet_run_dev(tc_ctx *c, int n_steps, int measure_every, int rec0, int measure_now, double *Z_dev,
                       double *ent_dev, double *ov_dev, int32_t *chi_dev) {
  CTX(c);
  const TcDev &d = c->d;
  if (n_steps > 0 && !c->have_model) return fail("tc_floquet_run_dev: no model set");
  if (measure_every < 1) measure_every = 1;
  const size_t zs = (size_t)d.R * d.L, es = (size_t)d.R * (d.L > 1 ? d.L - 1 : 0), os = (size_t)d.R * 2,
               cs = (size_t)d.R * (d.L + 1);
  auto rec = [&](int k) {
    return measure_into(c, nullptr, Z_dev ? Z_dev + zs * k : nullptr, (ent_dev && es) ? ent_dev + es * k : nullptr,
                        ov_dev ? ov_dev + os * k : nullptr, chi_dev ? chi_dev + cs * k : nullptr);
  };
  int k = rec0;
  if (measure_now) {
    if (rec(k)) return 1;
    ++k;
  }
  // periods t = 0 .. n_steps-1, a record after every period with t % measure_every == 0, taken chain range by chain
  // range on the stream that ran the range (see run_periods_with); the call joins the groups once, at its end
  const int k0 = k;
  return run_periods_with(c, n_steps, [&](int t, int r_lo, int r_hi, cudaStream_t st) {
    if (t % measure_every) return 0;
    const int kk = k0 + t / measure_every;
    ProfScope ps(c, TC_PROF_MEASURE);  // profile mode runs everything on the context's stream
    return measure_range(c, nullptr, Z_dev ? Z_dev + zs * kk : nullptr, (ent_dev && es) ? ent_dev + es * kk : nullptr,
                         ov_dev ? ov_dev + os * kk : nullptr, chi_dev ? chi_dev + cs * kk : nullptr, r_lo, r_hi, st);
  });
}

int tc_floquet_run_host(tc_ctx *c, const double *gates_host, const double *kick_host, int n_steps, int measure_every,
                        int measure_now, double *Z_host, double *ent_host, double *ov_host, int32_t *chi_host) {
  CTX(c);
  const TcDev &d = c->d;
  if (gates_host || kick_host)
    if (tc_set_model(c, gates_host, kick_host)) return 1;
  if (measure_every < 1) measure_every = 1;
  const int n_rec = (measure_now ? 1 : 0) + (n_steps > 0 ? (n_steps - 1) / measure_every + 1 : 0);
  const size_t zs = (size_t)d.R * d.L, es = (size_t)d.R * (d.L > 1 ? d.L - 1 : 0), os = (size_t)d.R * 2,
               cs = (size_t)d.R * (d.L + 1);
  const size_t nz = Z_host ? zs * n_rec : 0, ne = ent_host ? es * n_rec : 0, no = ov_host ? os * n_rec : 0,
               nc = chi_host ? cs * n_rec : 0;
  if (ensure_rec(c, (nz + ne + no + 1) * sizeof(double) + nc * sizeof(int32_t))) return 1;
  double *Zd = (double *)c->rec, *Ed = Zd + nz, *Od = Ed + ne;
  int32_t *Cd = (int32_t *)(Od + no);
  if (tc_floquet_run_dev(c, n_steps, measure_every, 0, measure_now, nz ? Zd : nullptr, ne ? Ed : nullptr,
                         no ? Od : nullptr, nc ? Cd : nullptr))
    return 1;
  if (nz) CK(cudaMemcpyAsync(Z_host, Zd, nz * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (ne) CK(cudaMemcpyAsync(ent_host, Ed, ne * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (no) CK(cudaMemcpyAsync(ov_host, Od, no * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (nc) CK(cudaMemcpyAsync(chi_host, Cd, nc * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}

// ---- diagnostics ----------------------------------------------------------------------------
int tc_dbg_get(tc_ctx *c, int which, int r, int jb, void *out_host, size_t bytes) {
  CTX(c);
  const TcDev &d = c->d;
  if (r < 0 || r >= d.ws_chains || jb < 0 || jb >= d.nbmax) return fail("tc_dbg_get: slot out of range");
  const size_t slot = (size_t)r * d.nbmax + jb;
  const void *src = nullptr;
  size_t maxb = 0;
  switch (which) {
    case TC_DBG_C: src = d.Cw + slot * d.slot_stride; maxb = d.slot_stride * sizeof(cplx); break;
    case TC_DBG_X: src = d.Xw + slot * d.slot_stride; maxb = d.slot_stride * sizeof(cplx); break;
    case TC_DBG_W: src = d.ww + slot * d.n2; maxb = d.n2 * sizeof(double); break;
    case TC_DBG_PERM: src = d.perm + slot * d.n2; maxb = d.n2 * sizeof(int); break;
    default: return fail("tc_dbg_get: unknown buffer");
  }
  if (bytes > maxb) return fail("tc_dbg_get: too many bytes");
  CK(cudaMemcpyAsync(out_host, src, bytes, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}

#ifdef TCB_TIMING
int tc_dbg_timing(unsigned long long *out8, int reset) {
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpyFromSymbol(out8, tcb::g_tcb_timing, 8 * sizeof(unsigned long long)));
  if (reset) {
    unsigned long long z[8] = {0};
    CK(cudaMemcpyToSymbol(tcb::g_tcb_timing, z, sizeof(z)));
  }
  return 0;
}
#endif

int tc_probe_fp64(int device, int use_dmma, double *gflops_out) {
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  double *out;
  CK(cudaMalloc(&out, 8));
  const int iters = 20000, blocks = prop.multiProcessorCount * 8;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaEventRecord(e0));
    if (use_dmma)
      probe_dmma_kernel<<<blocks, 256>>>(out, iters);
    else
      probe_fma_kernel<<<blocks, 256>>>(out, iters);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  g_launches.fetch_add(4);
  // per thread per iteration: 8 FMA (2 flop) or, per warp, 8 DMMA m8n8k4 (512 flop)
  const double flop = use_dmma ? (double)blocks * 8 * iters * 8 * 512.0 : (double)blocks * 256 * iters * 8 * 2.0;
  *gflops_out = flop / (best * 1e-3) * 1e-9;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  return 0;
}

}  // extern "C"
