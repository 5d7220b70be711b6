"""Per-kernel-class time of one Floquet period at a given shape (CUDA events, chain groups off).
usage: python scripts/profile_config.py R L chi prep_periods [eps]"""
import sys
import time

import numpy as np

sys.path.insert(0, '.')
from time_crystal_tensor_network_b200.engine import FloquetEnsemble, disorder_fields

R, L, chi, nprep = (int(x) for x in sys.argv[1:5])
eps = float(sys.argv[5]) if len(sys.argv) > 5 else 0.3
hs = np.array([disorder_fields(L, 0.3, 1000 + r) for r in range(R)])
ens = FloquetEnsemble(L, 1.0, 1.0, hs, epsilon=eps, chi_max=chi, mode='tebd', svd_min=1e-12, trunc_cut=1e-7)
ens.ctx.floquet_step(nprep)
ens.ctx.sync()
t0 = time.time()
ens.ctx.floquet_step(1)
ens.ctx.sync()
print(f'R={R} L={L} chi_max={chi}: {1e3 * (time.time() - t0):.1f} ms per period (product path), chi mid {ens.ctx.chi()[:, L // 2]}')
ens.ctx.profile(True)
ens.ctx.profile_read(reset=True)
ens.ctx.floquet_step(1)
prof = ens.ctx.profile_read(reset=True)
ens.ctx.profile(False)
tot = sum(v[0] for v in prof.values())
for k, v in prof.items():
    if v[1]:
        print(f'  {k:12s} {v[0]:10.2f} ms  {100 * v[0] / tot:5.1f} %  ({v[1]} launches)')
print(ens.ctx.flags())
