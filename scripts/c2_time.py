import sys, time, numpy as np
sys.path.insert(0, '.')
from time_crystal_tensor_network_b200 import engine as eng
L, R, n = 20, 256, 60
hs = np.array([eng.disorder_fields(L, 0.3, 1000 + r) for r in range(R)])
ens = eng.FloquetEnsemble(L, 1.0, 1.0, hs, epsilon=0.1, chi_max=64, mode='tebd', svd_min=1e-12, trunc_cut=1e-7)
ens.run(40)            # saturate
ens.ctx.sync()
t0 = time.time(); out = ens.run(n); ens.ctx.sync(); dt = time.time() - t0
print(f'C2 shape, {n} periods with a record per period: {dt:.2f} s, {R * n / dt:.0f} chain-steps/s, checksum {out["Z"].sum():.12f} {out["LE"].sum():.12f}')
