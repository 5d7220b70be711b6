"""Host-side handle of one engine context (an ensemble of R chains on one GPU) and the batched
disorder-ensemble driver built on it.

PyTorch is used only to own device memory and the CUDA stream; every computation is a call into
libtc_b200.so through the C ABI of include/tc_b200.h.
"""
import ctypes as C
import os
import sys

import numpy as np

from . import _lib
from ._lib import EngineError, check, dptr, iptr

SIGMA_X = np.array([[0, 1], [1, 0]], dtype=complex)


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise EngineError('no CUDA device visible: the B200 engine has no CPU fallback')
    return torch


def _echo(ov):
    """|overlap|^2 with a rounding overshoot above 1 (a product state back on itself: 1 + 9e-16) returned as exactly 1,
    as core.observables.calculate_loschmidt_echo does."""
    le = np.abs(ov) ** 2
    le[(le > 1.0) & (le < 1.0 + 1e-12)] = 1.0
    return le


def _torch_owns_memory():
    """Who owns a context's arena and stream.  PyTorch does when the process uses PyTorch anyway (bench.py, the sharded
    drivers, anything that hands torch tensors to run_dev); a process that never imported it -- ``python main.py``, the
    reference's scripts on the drop-in modules -- lets the library allocate (cudaMalloc / its own stream, the C ABI's
    arena = NULL form) and saves the six seconds ``import torch`` costs.  TC_ARENA=torch | engine forces either."""
    mode = os.environ.get('TC_ARENA', '')
    if mode in ('torch', 'engine'):
        return mode == 'torch'
    return 'torch' in sys.modules


class Context:
    """R independent chains of L sites, bond dimension at most chi_cap, on one GPU."""

    def __init__(self, L, chi_cap, R=1, device=0, storage_only=False):
        """storage_only: a context without SVD workspace (snapshots that are only measured, copied or overlapped);
        gate and Floquet calls on it raise."""
        self.lib = _lib.load()
        self.L, self.chi_cap, self.R, self.device = int(L), int(chi_cap), int(R), int(device)
        self.storage_only = bool(storage_only)
        nbytes = self.lib.tc_ctx_arena_bytes2(self.L, self.chi_cap, self.R, int(self.storage_only))
        if nbytes == 0:
            raise ValueError(f'invalid context shape L={L}, chi_cap={chi_cap}, R={R}')
        handle = C.c_void_p()
        if _torch_owns_memory():
            torch = _torch()
            self.stream = torch.cuda.Stream(device=self.device)
            with torch.cuda.device(self.device):
                self._arena = torch.empty(nbytes, dtype=torch.uint8, device=f'cuda:{self.device}')
            check(self.lib.tc_ctx_create2(self.device, self.L, self.chi_cap, self.R, int(self.storage_only),
                                          self._arena.data_ptr(), nbytes, self.stream.cuda_stream, C.byref(handle)),
                  'tc_ctx_create2')
        else:
            # arena and stream owned by the library (no CUDA device -> the call fails: there is no CPU fallback)
            self.stream = self._arena = None
            check(self.lib.tc_ctx_create2(self.device, self.L, self.chi_cap, self.R, int(self.storage_only),
                                          None, 0, None, C.byref(handle)), 'tc_ctx_create2')
        self._h = handle
        self.arena_bytes = nbytes

    # ------------------------------------------------------------------ lifetime
    def close(self):
        h, self._h = getattr(self, '_h', None), None
        if h:
            self.lib.tc_ctx_destroy(h)
        self._arena = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        check(self.lib.tc_sync(self._h), 'tc_sync')

    # ------------------------------------------------------------------ state
    def set_product_state(self, idx):
        idx = np.ascontiguousarray(np.broadcast_to(np.asarray(idx, dtype=np.int8), (self.R, self.L)))
        check(self.lib.tc_set_product_state(self._h, idx.ctypes.data_as(C.POINTER(C.c_int8))), 'tc_set_product_state')

    def get_site(self, r, site):
        cl, cr = C.c_int(), C.c_int()
        check(self.lib.tc_get_site(self._h, r, site, None, C.byref(cl), C.byref(cr)), 'tc_get_site')
        out = np.empty((cl.value, 2, cr.value), dtype=np.complex128)
        check(self.lib.tc_get_site(self._h, r, site, dptr(out), C.byref(cl), C.byref(cr)), 'tc_get_site')
        return out

    def set_site(self, r, site, tensor):
        t = np.ascontiguousarray(tensor, dtype=np.complex128)
        if t.ndim != 3 or t.shape[1] != 2:
            raise ValueError('site tensor must have shape (chi_l, 2, chi_r)')
        check(self.lib.tc_set_site(self._h, r, site, dptr(t), t.shape[0], t.shape[2]), 'tc_set_site')

    def get_S(self, r, bond):
        n = C.c_int()
        check(self.lib.tc_get_S(self._h, r, bond, None, C.byref(n)), 'tc_get_S')
        out = np.empty(n.value, dtype=np.float64)
        check(self.lib.tc_get_S(self._h, r, bond, dptr(out), C.byref(n)), 'tc_get_S')
        return out

    def set_S(self, r, bond, S):
        s = np.ascontiguousarray(S, dtype=np.float64)
        check(self.lib.tc_set_S(self._h, r, bond, dptr(s), len(s)), 'tc_set_S')

    def chi(self):
        """int32 [R][L+1]; chi[:, 0] = chi[:, L] = 1."""
        out = np.empty((self.R, self.L + 1), dtype=np.int32)
        check(self.lib.tc_get_chi(self._h, iptr(out)), 'tc_get_chi')
        return out

    def copy_chain_from(self, r_dst, src, r_src):
        check(self.lib.tc_copy_chain(self._h, r_dst, src._h, r_src), 'tc_copy_chain')

    def trunc_err(self, reset=False):
        out = np.empty(self.R, dtype=np.float64)
        check(self.lib.tc_get_trunc_err(self._h, dptr(out), int(reset)), 'tc_get_trunc_err')
        return out

    def flags(self):
        out = np.zeros(4, dtype=np.int32)
        check(self.lib.tc_get_flags(self._h, iptr(out)), 'tc_get_flags')
        return {'chi_cap_overflow': int(out[0]), 'svd_not_converged': int(out[1]), 'max_sweeps': int(out[2]),
                'mean_sweeps_large': out[3] / 100.0}

    # ------------------------------------------------------------------ model
    def set_model(self, gates, kick):
        """gates: complex [R][L-1][4][4] (or broadcastable [L-1][4][4]); kick: complex [R][2][2] (or [2][2])."""
        g = k = None
        if gates is not None and self.L > 1:
            g = np.ascontiguousarray(np.broadcast_to(np.asarray(gates, dtype=np.complex128).reshape(
                (-1, self.L - 1, 4, 4)), (self.R, self.L - 1, 4, 4)))
        if kick is not None:
            k = np.ascontiguousarray(np.broadcast_to(np.asarray(kick, dtype=np.complex128).reshape((-1, 2, 2)),
                                                     (self.R, 2, 2)))
        if g is None and self.L > 1 and gates is not None:
            raise ValueError('gates required')
        if g is None and self.L == 1:
            g = np.zeros((self.R, 1, 4, 4), dtype=np.complex128)   # marks the model as set
        check(self.lib.tc_set_model(self._h, dptr(g), dptr(k)), 'tc_set_model')

    def set_trunc(self, mode='reference', cutoff=1e-13, chi_max=0, svd_min=0.0, trunc_cut=0.0):
        m = {'reference': _lib.TRUNC_REFERENCE, 'tebd': _lib.TRUNC_TEBD}[mode]
        check(self.lib.tc_set_trunc(self._h, m, float(cutoff), int(chi_max or 0), float(svd_min or 0.0),
                                    float(trunc_cut or 0.0)), 'tc_set_trunc')

    # ------------------------------------------------------------------ gates
    def apply_layer(self, parity, kick_mode=0):
        check(self.lib.tc_apply_layer(self._h, parity, kick_mode), 'tc_apply_layer')

    def apply_kick(self):
        check(self.lib.tc_apply_kick(self._h), 'tc_apply_kick')

    def floquet_step(self, n_steps=1):
        check(self.lib.tc_floquet_step(self._h, n_steps), 'tc_floquet_step')

    def apply_two_site(self, r, site, gate):
        g = np.ascontiguousarray(np.asarray(gate, dtype=np.complex128).reshape(4, 4))
        check(self.lib.tc_apply_two_site(self._h, r, site, dptr(g)), 'tc_apply_two_site')

    def apply_one_site(self, r, site, op):
        o = np.ascontiguousarray(np.asarray(op, dtype=np.complex128).reshape(2, 2))
        check(self.lib.tc_apply_one_site(self._h, r, site, dptr(o)), 'tc_apply_one_site')

    # ------------------------------------------------------------------ observables
    def measure(self, entropies=True):
        """Returns (rdm[R][L][4], ent[R][L-1]); rdm = (rho00, rho11, Re, Im of sum theta_0 conj(theta_1))."""
        rdm = np.empty((self.R, self.L, 4), dtype=np.float64)
        ent = np.empty((self.R, max(self.L - 1, 0)), dtype=np.float64) if entropies else None
        check(self.lib.tc_measure(self._h, dptr(rdm), dptr(ent) if (ent is not None and ent.size) else None),
              'tc_measure')
        return rdm, ent

    def overlap(self, r_bra, ket, r_ket):
        """<self[r_bra] | ket[r_ket]>."""
        out = np.empty(2, dtype=np.float64)
        check(self.lib.tc_overlap(self._h, r_bra, ket._h, r_ket, dptr(out)), 'tc_overlap')
        return complex(out[0], out[1])

    def correlation(self, r, i, j, op1, op2):
        a = np.ascontiguousarray(np.asarray(op1, dtype=np.complex128).reshape(2, 2))
        b = np.ascontiguousarray(np.asarray(op2, dtype=np.complex128).reshape(2, 2))
        out = np.empty(2, dtype=np.float64)
        check(self.lib.tc_correlation(self._h, r, i, j, dptr(a), dptr(b), dptr(out)), 'tc_correlation')
        return complex(out[0], out[1])

    # ------------------------------------------------------------------ fused loop
    def n_records(self, n_steps, measure_every=1, measure_now=True):
        return (1 if measure_now else 0) + ((n_steps - 1) // measure_every + 1 if n_steps > 0 else 0)

    def run_host(self, n_steps, measure_every=1, measure_now=True, gates=None, kick=None,
                 want=('Z', 'ent', 'ov', 'chi')):
        """tc_floquet_run_host: upload the model (optional), run, download the records (host arrays)."""
        n = self.n_records(n_steps, measure_every, measure_now)
        Z = np.empty((n, self.R, self.L)) if 'Z' in want else None
        ent = np.empty((n, self.R, max(self.L - 1, 0))) if 'ent' in want else None
        ov = np.empty((n, self.R, 2)) if 'ov' in want else None
        chi = np.empty((n, self.R, self.L + 1), dtype=np.int32) if 'chi' in want else None
        g = k = None
        if gates is not None:
            g = np.ascontiguousarray(np.asarray(gates, dtype=np.complex128).reshape(self.R, max(self.L - 1, 1), 4, 4))
        if kick is not None:
            k = np.ascontiguousarray(np.asarray(kick, dtype=np.complex128).reshape(self.R, 2, 2))
        check(self.lib.tc_floquet_run_host(self._h, dptr(g), dptr(k), n_steps, measure_every, int(measure_now),
                                           dptr(Z), dptr(ent) if (ent is not None and ent.size) else None,
                                           dptr(ov), iptr(chi)), 'tc_floquet_run_host')
        return {'Z': Z, 'ent': ent, 'ov': ov, 'chi': chi}

    def run_dev(self, n_steps, measure_every=1, rec0=0, measure_now=False, Z=None, ent=None, ov=None, chi=None):
        """tc_floquet_run_dev with torch device tensors as record buffers; asynchronous."""
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
        check(self.lib.tc_floquet_run_dev(self._h, n_steps, measure_every, rec0, int(measure_now),
                                          p(Z), p(ent), p(ov), p(chi)), 'tc_floquet_run_dev')

    def profile(self, enable=True):
        check(self.lib.tc_profile(self._h, int(enable)), 'tc_profile')

    def profile_read(self, reset=True):
        """{kernel class: (milliseconds, timed launch groups)} since the last reset (synchronises)."""
        ms = np.zeros(8, dtype=np.float64)
        cnt = np.zeros(8, dtype=np.int64)
        check(self.lib.tc_profile_read(self._h, dptr(ms), cnt.ctypes.data_as(C.POINTER(C.c_longlong)), int(reset)),
              'tc_profile_read')
        return {name: (float(ms[k]), int(cnt[k])) for k, name in enumerate(_lib.PROF_CLASSES)}

    def dbg_get(self, which, r, jb, shape, dtype):
        out = np.empty(shape, dtype=dtype)
        check(self.lib.tc_dbg_get(self._h, which, r, jb, out.ctypes.data_as(C.c_void_p), out.nbytes), 'tc_dbg_get')
        return out


def launch_count():
    return int(_lib.load().tc_launch_count())


def probe_fp64(device=0, dmma=False):
    """Measured FP64 throughput in GFLOP/s (FMA pipe or mma.sync DMMA)."""
    out = C.c_double()
    check(_lib.load().tc_probe_fp64(device, int(dmma), C.byref(out)), 'tc_probe_fp64')
    return out.value


# ----------------------------------------------------------------------------------------------
# model helpers (host, vectorised over chains)
# ----------------------------------------------------------------------------------------------
def kick_matrix(epsilon=0.0):
    """expm(-i (pi/2)(1-eps) sigma_x) = cos(a) I - i sin(a) sigma_x; eps = 0 is the reference's exact
    pi-pulse -i sigma_x (src/models/kicked_ising.py:76)."""
    eps = np.asarray(epsilon, dtype=float)
    a = np.pi / 2 * (1.0 - eps)
    c, s = np.where(eps == 0.0, 0.0, np.cos(a)), np.sin(a)
    k = np.zeros(np.shape(a) + (2, 2), dtype=complex)
    k[..., 0, 0] = k[..., 1, 1] = c
    k[..., 0, 1] = k[..., 1, 0] = -1j * s
    return k


def ising_gates(J, h_fields, tau):
    """Diagonal bond gates exp(-i tau/2 (J zz' + h_i z + h_{i+1} z')) (kicked_ising.py:83-88) for
    h_fields [R][L] (J, tau scalars or [R]).  Returns complex [R][L-1][4][4]."""
    h = np.atleast_2d(np.asarray(h_fields, dtype=float))
    R, L = h.shape
    J = np.broadcast_to(np.asarray(J, dtype=float), (R,))[:, None]
    tau = np.broadcast_to(np.asarray(tau, dtype=float), (R,))[:, None]
    g = np.zeros((R, max(L - 1, 0), 4, 4), dtype=complex)
    z = (1.0, -1.0)
    for p0 in range(2):
        for p1 in range(2):
            e = J * z[p0] * z[p1] + h[:, :-1] * z[p0] + h[:, 1:] * z[p1]
            g[:, :, 2 * p0 + p1, 2 * p0 + p1] = np.exp(-1j * tau / 2 * e)
    return g


def disorder_fields(L, W, seed):
    """h_fields exactly as the reference draws them (kicked_ising.py:55-59): legacy global RNG, reseeded."""
    np.random.seed(seed)
    return np.random.uniform(-W, W, L)


def product_indices(L, state='neel', up_index=1):
    """Internal basis indices of the product states of src/core/tensor_utils.py:44-55."""
    if state == 'all_up':
        lab = [1] * L
    elif state == 'all_down':
        lab = [0] * L
    elif state == 'neel':
        lab = [1 if i % 2 == 0 else 0 for i in range(L)]
    else:
        raise ValueError(f'Unknown state type: {state}')
    return np.array([up_index if u else 1 - up_index for u in lab], dtype=np.int8)


class FloquetEnsemble:
    """Batched kicked-Ising Floquet evolution of R chains (disorder realisations, initial states or
    (eps, W, J) phase-diagram points) on one GPU, observables recorded on the device.

    The per-chain sequence is the reference's (src/models/kicked_ising.py:100-160): even bonds, odd
    bonds, kick on every site, even bonds, odd bonds, each two-site gate followed by its own SVD.
    """

    def __init__(self, L, J, tau, h_fields, epsilon=0.0, chi_max=64, mode='tebd', svd_min=1e-12,
                 trunc_cut=1e-7, cutoff=1e-13, state='neel', up_index=1, device=0, chi_cap=None):
        h = np.atleast_2d(np.asarray(h_fields, dtype=float))
        self.R, self.L = h.shape
        if self.L != L:
            raise ValueError('h_fields must have shape [R][L]')
        self.tau = np.broadcast_to(np.asarray(tau, dtype=float), (self.R,)).copy()
        self.J = np.broadcast_to(np.asarray(J, dtype=float), (self.R,)).copy()
        self.h_fields = h
        self.gates = ising_gates(self.J, h, self.tau)
        self.kick = np.ascontiguousarray(np.broadcast_to(kick_matrix(epsilon), (self.R, 2, 2)))
        cap = chi_cap or min(int(chi_max), 2 ** (L // 2))
        self.ctx = Context(L, max(cap, 1), self.R, device)
        self.ctx.set_trunc(mode, cutoff=cutoff, chi_max=chi_max, svd_min=svd_min, trunc_cut=trunc_cut)
        if isinstance(state, str):
            idx = product_indices(L, state, up_index)
        else:
            idx = np.asarray(state, dtype=np.int8)
        self.ctx.set_product_state(idx)
        self.ctx.set_model(self.gates, self.kick)
        self.periods_done = 0

    def run(self, n_periods, measure_every=1, measure_now=None, upload_model=False):
        """Evolve n_periods; returns host arrays Z[T][R][L], S_ent[T][R][L-1], overlap[T][R] (complex),
        LE[T][R], chi[T][R][L+1], periods[T]."""
        if measure_now is None:
            measure_now = self.periods_done == 0
        rec = self.ctx.run_host(n_periods, measure_every, measure_now,
                                gates=self.gates if upload_model else None,
                                kick=self.kick if upload_model else None)
        periods = ([self.periods_done] if measure_now else []) + \
            [self.periods_done + t + 1 for t in range(n_periods) if t % measure_every == 0]
        self.periods_done += n_periods
        ov = rec['ov'][..., 0] + 1j * rec['ov'][..., 1]
        fl = self.ctx.flags()
        if fl['chi_cap_overflow']:
            raise EngineError(f"bond dimension exceeded chi_cap={self.ctx.chi_cap} in {fl['chi_cap_overflow']} updates")
        if fl['svd_not_converged']:
            raise EngineError(f"{fl['svd_not_converged']} SVDs did not converge")
        return {'Z': rec['Z'], 'S_ent': rec['ent'], 'overlap': ov, 'LE': _echo(ov), 'chi': rec['chi'],
                'periods': np.array(periods), 'flags': fl}

    def close(self):
        self.ctx.close()
