/*
 * tc_b200.h -- C ABI of the B200-native kicked-Ising Floquet/TEBD engine.
 *
 * The reference (connor-a-casey/time-crystal-tensor-network) has no FFI of its own: its
 * boundary is the Python API of src/models/kicked_ising.py, src/dynamics/tebd_evolution.py,
 * src/core/observables.py and src/core/tensor_utils.py, all of which bottom out in TeNPy's
 * MPS methods.  Each entry point below names the reference call site it stands in for
 * (paths relative to the reference root).  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; tc_last_error() gives the text
 *   - complex numbers are interleaved (re, im) doubles, i.e. numpy complex128 memory
 *   - "host" pointers are ordinary (ideally pinned) host memory; "dev" pointers are device memory
 *   - one context = one ensemble of R independent chains (disorder realisations / phase points)
 *     of L sites on one GPU; all work is issued on the context's stream; nothing synchronises
 *     unless the function moves data to the host or is tc_sync()
 *   - site tensor (r, i): compact row-major [chi_l][2][chi_r]; Schmidt values S[r][b], bond b is
 *     left of site b (b = 0..L); chi[r][0] = chi[r][L] = 1
 *   - all site tensors are kept in right-canonical ("B") form
 */
#ifndef TC_B200_H
#define TC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tc_ctx tc_ctx;

/* truncation modes for tc_set_trunc */
#define TC_TRUNC_REFERENCE 0 /* TeNPy MPS.apply_local_op default: keep sigma > cutoff (absolute), renormalise;
                                what src/models/kicked_ising.py:186 actually does (trunc_params ignored)   */
#define TC_TRUNC_TEBD 1      /* TeNPy truncate(): chi_max, svd_min, trunc_cut on the normalised spectrum      */

/* workspace selectors for tc_dbg_get (stage-wise kernel tests) */
#define TC_DBG_C 0    /* gate-applied two-site tensor C, complex [M][N]                            */
#define TC_DBG_X 1    /* theta = S_l C after the Jacobi sweeps: rows are sigma_k v_k^H, complex [M][N] */
#define TC_DBG_W 2    /* row norms = singular values, unsorted, real [M]                            */
#define TC_DBG_PERM 3 /* descending order permutation, int32 [M]                                   */

int tc_version(void);
const char *tc_last_error(void);
int tc_device_count(int *count);

/* ---- context ------------------------------------------------------------------------------ */
/* bytes of device memory a context needs (state + one layer of workspace) */
size_t tc_ctx_arena_bytes(int L, int chi_cap, int R);
/* arena: device memory of at least tc_ctx_arena_bytes (e.g. a torch uint8 tensor); NULL = cudaMalloc.
   stream: a cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream); NULL = a new stream.      */
int tc_ctx_create(int device, int L, int chi_cap, int R, void *arena, size_t arena_bytes,
                  void *stream, tc_ctx **out);
/* Storage-only variant (storage_only != 0): the arena holds the state, the model and the observable scratch but no
   SVD workspace -- for snapshots that are only measured, copied or overlapped (the states list
   CustomFloquet.evolve_floquet returns, src/dynamics/tebd_evolution.py:236-241).  Gate and Floquet calls on such a
   context fail with an error; copy the chain into a full context (tc_copy_chain) to evolve it. */
size_t tc_ctx_arena_bytes2(int L, int chi_cap, int R, int storage_only);
int tc_ctx_create2(int device, int L, int chi_cap, int R, int storage_only, void *arena, size_t arena_bytes,
                   void *stream, tc_ctx **out);
int tc_ctx_destroy(tc_ctx *ctx);
int tc_sync(tc_ctx *ctx);
int tc_ctx_info(tc_ctx *ctx, int *L, int *chi_cap, int *R, int *device);
/* health counters since context creation: out[0] = updates truncated by chi_cap (overflow), out[1] = SVDs
   that hit the sweep limit, out[2] = largest number of Jacobi sweeps any SVD needed, out[3] = 100 x mean sweeps of the SVDs with
   at least 128 rows (fast path only)                                                              */
int tc_get_flags(tc_ctx *ctx, int32_t *out4);

/* ---- state --------------------------------------------------------------------------------- */
/* MPS.from_product_state (src/core/tensor_utils.py:60): idx[R][L] internal basis indices (host).
   Also remembered as the reference state for the Loschmidt echo inside tc_floquet_steps.         */
int tc_set_product_state(tc_ctx *ctx, const int8_t *idx_host);
/* raw tensor access (host complex128 [chi_l][2][chi_r]) -- MPS.get_B / set_B                    */
int tc_set_site(tc_ctx *ctx, int r, int site, const double *data_host, int chi_l, int chi_r);
int tc_get_site(tc_ctx *ctx, int r, int site, double *out_host, int *chi_l, int *chi_r);
/* Schmidt values on bond b -- MPS.get_SL (src/core/observables.py:250)                          */
int tc_set_S(tc_ctx *ctx, int r, int bond, const double *S_host, int n);
int tc_get_S(tc_ctx *ctx, int r, int bond, double *out_host, int *n);
/* MPS.chi (src/dynamics/tebd_evolution.py:233,247): out[R][L+1] int32 (host)                     */
int tc_get_chi(tc_ctx *ctx, int32_t *out_host);
/* MPS.copy (src/models/kicked_ising.py:115,...): copy chain r_src of src into chain r_dst of dst */
int tc_copy_chain(tc_ctx *dst, int r_dst, tc_ctx *src, int r_src);
/* accumulated truncation error (discarded weight) per chain, out[R] (host); reset with reset!=0  */
int tc_get_trunc_err(tc_ctx *ctx, double *out_host, int reset);

/* ---- model --------------------------------------------------------------------------------- */
/* KickedIsingModel._prepare_gates (src/models/kicked_ising.py:73-98): gates[R][L-1][4][4] complex,
   row index (p0 p1), column index (q0 q1); kick[R][2][2] complex.  Host pointers.                */
int tc_set_model(tc_ctx *ctx, const double *gates_host, const double *kick_host);
int tc_set_trunc(tc_ctx *ctx, int mode, double cutoff, int chi_max, double svd_min, double trunc_cut);

/* ---- gates --------------------------------------------------------------------------------- */
/* KickedIsingModel._apply_ising_evolution, one parity class (kicked_ising.py:133-146): two-site
   gate + SVD update on bonds (i, i+1), i = parity, parity+2, ... for every chain.
   kick_mode bit0: apply the kick to both sites of every bond before the gate;
   kick_mode bit1: apply the kick to the right site of the last bond (i = L-2) only.              */
int tc_apply_layer(tc_ctx *ctx, int parity, int kick_mode);
/* KickedIsingModel._apply_pi_pulse (kicked_ising.py:150-160): kick on every site of every chain  */
int tc_apply_kick(tc_ctx *ctx);
/* KickedIsingModel.floquet_step (kicked_ising.py:100-126) on every chain: even, odd, kick (fused
   into the next layer's loads), even, odd.  n_steps periods, no measurement.                     */
int tc_floquet_step(tc_ctx *ctx, int n_steps);
/* MPS.apply_local_op with a caller-supplied operator on one chain (kicked_ising.py:186,206;
   src/core/tensor_utils.py:103): op is [4][4] (two-site, sites site,site+1) or [2][2] (one-site) */
int tc_apply_two_site(tc_ctx *ctx, int r, int site, const double *gate_host);
int tc_apply_one_site(tc_ctx *ctx, int r, int site, const double *op_host);

/* ---- observables --------------------------------------------------------------------------- */
/* single-site reduced density matrices and bond entropies of every chain:
   rdm[R][L][4] = (rho00, rho11, Re rho01, Im rho01)  -> MPS.expectation_value (observables.py:62)
   ent[R][L-1]  = -sum s^2 ln s^2, s^2 > 1e-30        -> MPS.entanglement_entropy (tensor_utils.py:180)
   Device pointers; either may be NULL.                                                            */
int tc_measure_dev(tc_ctx *ctx, double *rdm_dev, double *ent_dev);
int tc_measure(tc_ctx *ctx, double *rdm_host, double *ent_host);
/* <bra_r|ket_r'> -- MPS.overlap (observables.py:25; tensor_utils.py:192); out[2] = (re, im), host */
int tc_overlap(tc_ctx *bra, int r_bra, tc_ctx *ket, int r_ket, double *out_host);
/* two-point function <op1_i op2_j> on chain r -- MPS.correlation_function (observables.py:121);
   op1, op2 are [2][2] complex (host), out[2] = (re, im)                                           */
int tc_correlation(tc_ctx *ctx, int r, int i, int j, const double *op1_host, const double *op2_host,
                   double *out_host);

/* ---- fused time loop ----------------------------------------------------------------------- */
/* CustomFloquet.evolve_floquet (src/dynamics/tebd_evolution.py:218-259) for the whole ensemble,
   entirely on the device: n_steps Floquet periods; after period t (0-based) with
   t % measure_every == 0 the observables are written to record k = t / measure_every + rec0.
   Output device buffers (any may be NULL):
     Z_dev  [n_rec][R][L]    <sigma^z_i>
     ent_dev[n_rec][R][L-1]  bond entropies
     ov_dev [n_rec][R][2]    <psi_0|psi(t)> with psi_0 the product state given to tc_set_product_state
     chi_dev[n_rec][R][L+1]  int32 bond dimensions
   If measure_now != 0 the current state is recorded first, into record rec0, and the records of
   the periods follow from rec0 + 1 (n_steps = 0, measure_now = 1 records the current state only).
   Asynchronous: the chain groups take their records on their own streams and are joined to the
   context's stream when the call returns, so work enqueued on that stream afterwards (and the
   host after tc_sync) sees every record.                                                          */
int tc_floquet_run_dev(tc_ctx *ctx, int n_steps, int measure_every, int rec0, int measure_now,
                       double *Z_dev, double *ent_dev, double *ov_dev, int32_t *chi_dev);
/* same with HOST buffers for inputs and outputs (the call the end-to-end benchmark times): uploads
   gates/kick (may be NULL = keep), runs, downloads the records.  n_rec records are written starting
   at record 0: if measure_now, record 0 is the state before the first step.                       */
int tc_floquet_run_host(tc_ctx *ctx, const double *gates_host, const double *kick_host, int n_steps,
                        int measure_every, int measure_now, double *Z_host, double *ent_host,
                        double *ov_host, int32_t *chi_host);

/* ---- diagnostics --------------------------------------------------------------------------- */
/* per-kernel-class device timing with CUDA events on the context's stream (bench.py's roofline leg).
   tc_profile_read synchronises, then writes the summed milliseconds and the number of timed launch
   groups per class into ms_out[TC_PROF_NCLASS] / count_out[TC_PROF_NCLASS] (count_out may be NULL).   */
#define TC_PROF_THETA 0    /* K1 theta GEMM (+ general-gate mix)      */
#define TC_PROF_QR 1       /* K2a Householder QR preconditioner        */
#define TC_PROF_JACOBI 2   /* K2b one-sided Jacobi sweeps              */
#define TC_PROF_FINALIZE 3 /* K2c sort / truncate / renormalise / V^H  */
#define TC_PROF_BLEFT 4    /* K3 B_i = C V_k GEMM                      */
#define TC_PROF_MEASURE 5  /* observable kernels                       */
#define TC_PROF_KICK 6     /* stand-alone kick                         */
#define TC_PROF_NCLASS 8
int tc_profile(tc_ctx *ctx, int enable);
int tc_profile_read(tc_ctx *ctx, double *ms_out, long long *count_out, int reset);
/* copy a workspace buffer of slot (r, bond position jb in the last layer) to the host             */
int tc_dbg_get(tc_ctx *ctx, int which, int r, int jb, void *out_host, size_t bytes);
/* count of kernel launches issued through this library since load                                */
long long tc_launch_count(void);
/* FP64 throughput probes (for the roofline denominator): returns achieved GFLOP/s                 */
int tc_probe_fp64(int device, int use_dmma, double *gflops_out);

#ifdef __cplusplus
}
#endif
#endif /* TC_B200_H */
