"""Exploration on the GPU box: FP64 peaks and per-period time of the L=32, chi=128 workload."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, '.')
from time_crystal_tensor_network_b200.engine import FloquetEnsemble, probe_fp64, launch_count, disorder_fields

R = int(sys.argv[1]) if len(sys.argv) > 1 else 4
L = int(sys.argv[2]) if len(sys.argv) > 2 else 32
chi = int(sys.argv[3]) if len(sys.argv) > 3 else 128
nper = int(sys.argv[4]) if len(sys.argv) > 4 else 16
eps = float(sys.argv[5]) if len(sys.argv) > 5 else 0.1

print('fp64 fma GF/s', probe_fp64(0, False), 'dmma GF/s', probe_fp64(0, True), flush=True)
hs = np.array([disorder_fields(L, 0.3, 1000 + r) for r in range(R)])
ens = FloquetEnsemble(L, 1.0, 1.0, hs, epsilon=eps, chi_max=chi, mode='tebd', svd_min=1e-12, trunc_cut=1e-7)
for t in range(nper):
    t0 = time.time()
    out = ens.run(1, measure_now=False)
    dt = time.time() - t0
    c = out['chi'][-1]
    print(f'period {t + 1}: {dt * 1e3:9.1f} ms  chi max {c.max()} mean-mid {c[:, L // 2].mean():.1f} '
          f'S_mid {out["S_ent"][-1][:, L // 2 - 1].mean():.4f} LE {out["LE"][-1].mean():.3e} flags {out["flags"]}',
          flush=True)
print('launches', launch_count())
