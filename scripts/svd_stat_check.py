"""Failure statistics of a Jacobi variant on one layer from a saturated state.
usage: python scripts/svd_stat_check.py L chi prep reps VAR=VAL[,VAR=VAL] ..."""
import os
import sys

import numpy as np

os.environ.setdefault('TC_ARENA', 'torch')   # the state is cloned arena to arena

sys.path.insert(0, '.')
from time_crystal_tensor_network_b200 import engine as eng
from time_crystal_tensor_network_b200 import _lib

L, chi, prep, reps = (int(x) for x in sys.argv[1:5])
PAR = int(os.environ.get('DBG_PARITY', 0))
hs = np.array([eng.disorder_fields(L, 0.3, 11)])
kw = dict(epsilon=0.3, chi_max=chi, mode='tebd', svd_min=1e-12, trunc_cut=1e-7)
os.environ['TC_GROUPS'] = '1'
os.environ['TC_JACOBI'] = 'wide_v1'
base = eng.FloquetEnsemble(L, 1.0, 1.0, hs, **kw)
base.ctx.floquet_step(prep)
base.ctx.sync()
chi0 = base.ctx.chi()[0]
Sl = [base.ctx.get_S(0, min(L, 2 * jb + PAR)) for jb in range(L // 2)]
for var in sys.argv[5:]:
    for k in ('TC_JACOBI', 'TC_WIDE_CLUSTER', 'TC_ROT64'):
        os.environ.pop(k, None)
    for kv in var.split(','):
        k, v = kv.split('=')
        os.environ[k] = v
    e = eng.FloquetEnsemble(L, 1.0, 1.0, hs, **kw)
    bad = {}
    worst = 0.0
    for rep in range(reps):
        e.ctx._arena.copy_(base.ctx._arena)
        e.ctx.apply_layer(PAR, 0)
        e.ctx.sync()
        if os.environ.get('TC_DBG_FLAGS'):
            print('rep', rep, end=' ', flush=True)
            e.ctx.flags()
        for jb in range(L // 2):
            i = 2 * jb + PAR
            if i + 1 >= L:
                continue
            M, N = 2 * int(chi0[i]), 2 * int(chi0[i + 2])
            if min(M, N) < 64:
                continue
            C = e.ctx.dbg_get(_lib.DBG_C, 0, jb, (M, N), np.complex128)
            if rep == 0:
                sv = np.linalg.svd(C * np.repeat(Sl[jb], 2)[:, None], compute_uv=False)
                bad.setdefault(('sv', jb), sv)
            sv = bad[('sv', jb)]
            w = np.sort(e.ctx.dbg_get(_lib.DBG_W, 0, jb, (min(M, N),), np.float64))[::-1]
            err = float(np.max(np.abs(w - sv)) / sv[0])
            if 'TC_ROT64' in os.environ and int(os.environ['TC_ROT64']) & 2048 and min(M, N) + 64 <= 2 * chi:
                dd = e.ctx.dbg_get(_lib.DBG_W, 0, jb, (2 * chi,), np.float64)[min(M, N) + 32:min(M, N) + 48]
                if dd[:8].any() or err > 1e-12:
                    k = int(np.argmax(np.abs(w - sv)))
                    print('   sigma index', k, 'of', len(sv), 'device %.6e lapack %.6e; neighbours lapack' % (w[k], sv[k]), sv[max(0, k - 2):k + 3], 'device', w[max(0, k - 2):k + 3])
                    print('   rep', rep, 'bond', jb, 'err %.1e' % err, 'counts[code 1=P load, 2=P after internal, 3=Q0 after internal, 4=Q load, 5=after rounds]', dd[:8], 'first: code %d sweep %d p %d q %d row %d tracked %.6e actual %.6e w %.3e' % tuple(dd[8:16]))
            if err > 1e-12:
                print('   rep', rep, 'bond', jb, 'err %.1e' % err)
                bad[(jb, M, N)] = bad.get((jb, M, N), 0) + 1
                worst = max(worst, err)
            if 'TC_ROT64' in os.environ and int(os.environ['TC_ROT64']) & 256 and min(M, N) + 14 < 2 * chi and jb == 3 and rep < 12:
                fro = e.ctx.dbg_get(_lib.DBG_W, 0, jb, (2 * chi,), np.float64)[min(M, N) + 1:min(M, N) + 13]
                print('   rep', rep, 'err %.1e' % err, 'Frobenius drift per sweep:', ' '.join('%.1e' % (x / fro[0] - 1) for x in fro[1:10]))
    print(var, 'failures per bond:', {k: v for k, v in bad.items() if len(k) == 3}, f'of {reps}; worst {worst:.1e}', e.ctx.flags(), flush=True)
    e.close()
