"""Config 4's regime (one chain, L = 64, chi_max = 256): time per period and Jacobi sweep statistics of the wide-matrix
SVD kernels after the bond dimension has saturated.  usage: python scripts/c4_sweeps.py [prep_periods] [timed_periods]
(kernel variant through the environment: TC_WIDE_CLUSTER, TC_JACOBI=wide_v1, TC_QR_CLUSTER, TC_THRESH)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from time_crystal_tensor_network_b200 import engine as eng  # noqa: E402

prep = int(sys.argv[1]) if len(sys.argv) > 1 else 60
timed = int(sys.argv[2]) if len(sys.argv) > 2 else 10
L = 64
hs = np.array([eng.disorder_fields(L, 0.3, 1000)])
ens = eng.FloquetEnsemble(L, 1.0, 1.0, hs, epsilon=0.1, chi_max=256, mode='tebd', svd_min=1e-12, trunc_cut=1e-7)
ens.ctx.floquet_step(prep)
ens.ctx.sync()
f0 = ens.ctx.flags()
t0 = time.time()
ens.ctx.floquet_step(timed)
ens.ctx.sync()
dt = time.time() - t0
f1 = ens.ctx.flags()
print({'s_per_period': round(dt / timed, 4), 'chi_mid': int(ens.ctx.chi()[0][L // 2]), 'mean_sweeps_large_cumulative': (f0['mean_sweeps_large'], f1['mean_sweeps_large']),
       'max_sweeps': f1['max_sweeps'], 'variant': {k: v for k, v in os.environ.items() if k.startswith('TC_')}})
