"""Diagnostic: per-phase cycle counters of the Jacobi kernels (library built with -DTCB_TIMING as
libtc_b200_timing.so).  TC_JACOBI=rb (default) reports the register-blocked kernel's phases, TC_JACOBI=blocked the
16-warp kernel's."""
import ctypes as C, sys, os
import numpy as np
sys.path.insert(0, '.')
from time_crystal_tensor_network_b200 import _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), 'libtc_b200_timing.so')
from time_crystal_tensor_network_b200 import engine as eng
R, L, chi = 8, 32, 128
hs = np.array([eng.disorder_fields(L, 0.3, 1000 + r) for r in range(R)])
ens = eng.FloquetEnsemble(L, 1.0, 1.0, hs, epsilon=0.3, chi_max=chi, mode='tebd', svd_min=1e-12, trunc_cut=1e-7)
ens.ctx.floquet_step(9)
ens.ctx.set_model(ens.gates, np.broadcast_to(eng.kick_matrix(0.1), (R, 2, 2)).copy())
ens.ctx.floquet_step(1)
lib = _lib.load()
out = (C.c_ulonglong * 8)()
lib.tc_dbg_timing(out, 1)
ens.ctx.floquet_step(1)
lib.tc_dbg_timing(out, 1)
v = np.array(list(out), dtype=float)
if os.environ.get('TCB_COARSE'):
    # library built with -DTCB_TIMING -DTCB_COARSE: coarse phases of a sweep of the 16-warp kernel (warp 3, K = N = 256)
    names = ['row norms at the sweep start', 'wait for the P block', 'internal pairs', 'wait for a q block',
             'the 16 rounds of a visit', 'end-of-visit fence + barrier + store', 'P block back to global', 'kernel total']
    for n, x in zip(names, v):
        print(f'{n:40s} {x:.4g}  ({100 * x / v[7]:.1f} % of kernel)')
    print('sum of phases / kernel total = %.3f' % (v[:7].sum() / v[7]))
elif os.environ.get('TC_JACOBI', 'rb') == 'blocked':
    names = ['load+dot', 'warp reduce', 'rotation set-up', 'rotate+store', 'wait/barrier', 'pairs rotated', 'pairs visited', 'kernel total']
    for n, x in zip(names, v):
        print(f'{n:18s} {x:.4g}')
    print('per visited pair: load+dot %.0f  reduce %.0f  | per rotated pair: setup %.0f  rotate %.0f | wait per visited %.0f' % (
        v[0] / v[6], v[1] / v[6], v[2] / v[5], v[3] / v[5], v[4] / v[6]))
    print('sum of phases / kernel total = %.2f ; rotated fraction %.2f' % (v[:5].sum() / v[7], v[5] / v[6]))
else:
    names = ['internal phase', 'streaming visits', '  hand-over waits', 'stage wait + barrier', 'P block load wait',
             '  tournaments', 'visits', 'kernel total']
    for n, x in zip(names, v):
        print(f'{n:22s} {x:.4g}  ({100 * x / v[7]:.1f} % of kernel)' if n != 'visits' else f'{n:22s} {x:.4g}')
    print('cycles per visit %.0f (= %d pairs per warp)' % (v[1] / v[6], 64))
print(ens.ctx.flags())
