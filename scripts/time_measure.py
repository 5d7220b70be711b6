import sys, time, numpy as np
sys.path.insert(0, '.')
from time_crystal_tensor_network_b200.engine import FloquetEnsemble, disorder_fields
R, L, chi = 32, 32, 128
hs = np.array([disorder_fields(L, 0.3, 1000 + r) for r in range(R)])
ens = FloquetEnsemble(L, 1.0, 1.0, hs, epsilon=0.3, chi_max=chi, mode='tebd', svd_min=1e-12, trunc_cut=1e-7)
ens.ctx.floquet_step(9)
ens.ctx.profile(True); ens.ctx.profile_read(reset=True)
for _ in range(5): ens.ctx.run_host(0, 1, True)
prof = ens.ctx.profile_read(reset=True)
print('measure class (measure + overlap + chi_record):', prof['measure'][0] / 5, 'ms per snapshot')
