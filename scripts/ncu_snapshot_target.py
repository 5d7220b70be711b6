"""Target of the ncu capture of the observable kernels: 32 chains at the metric shape, saturated bonds, five snapshots.
usage: python scripts/ncu_snapshot_target.py"""
import sys

import numpy as np

sys.path.insert(0, '.')
from time_crystal_tensor_network_b200.engine import FloquetEnsemble, disorder_fields

R, L, chi = 32, 32, 128
hs = np.array([disorder_fields(L, 0.3, 1000 + r) for r in range(R)])
ens = FloquetEnsemble(L, 1.0, 1.0, hs, epsilon=0.3, chi_max=chi, mode='tebd', svd_min=1e-12, trunc_cut=1e-7)
ens.ctx.floquet_step(10)
for _ in range(5):
    rec = ens.ctx.run_host(0, 1, True)
print('LE', (rec['ov'][0, :3] ** 2).sum(axis=1), 'chi mid', ens.ctx.chi()[:3, L // 2])
