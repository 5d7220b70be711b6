import sys, numpy as np
sys.path.insert(0, '.')
from time_crystal_tensor_network_b200.engine import FloquetEnsemble, disorder_fields
L = 10
h = disorder_fields(L, 0.3, 42)
ens = FloquetEnsemble(L, 1.0, 1.0, h[None, :], epsilon=0.1, mode='reference', chi_max=32)
for t in range(16):
    ens.ctx.floquet_step(1)
    print(t, ens.ctx.flags(), ens.ctx.chi()[0].tolist(), flush=True)
