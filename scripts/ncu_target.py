"""Target of the ncu captures (profiles/README.md): R chains of the metric shape, `prep` periods at eps = 0.3 until the
central bonds are saturated, then `n` more periods.  usage: python scripts/ncu_target.py [R=14] [prep=10] [n=1] [L=32] [chi=128]"""
import sys

import numpy as np

sys.path.insert(0, '.')
from time_crystal_tensor_network_b200.engine import FloquetEnsemble, disorder_fields

arg = [int(x) for x in sys.argv[1:]]
R, prep, n, L, chi = (arg + [14, 10, 1, 32, 128][len(arg):])[:5]
hs = np.array([disorder_fields(L, 0.3, 1000 + r) for r in range(R)])
ens = FloquetEnsemble(L, 1.0, 1.0, hs, epsilon=0.3, chi_max=chi, mode='tebd', svd_min=1e-12, trunc_cut=1e-7)
ens.ctx.floquet_step(prep)
ens.ctx.sync()
rec = ens.ctx.run_host(n, 1, False)
print('chi mid', ens.ctx.chi()[:, L // 2], 'flags', ens.ctx.flags(), 'Z[0,0,:4]', rec['Z'][0, 0, :4])
