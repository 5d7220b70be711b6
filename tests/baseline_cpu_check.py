"""Checker used by scripts/run_baseline_configs.py --cpu (TEST INFRASTRUCTURE: the only place that script reaches the
oracle): one more Floquet period of chain 0 on the GPU and on the CPU oracle, both from the GPU's own final state."""
import time

import numpy as np


def one_period_against_oracle(ens, L, chi):
    """One more period of chain 0 on the GPU and on the CPU oracle from the same state."""
    from oracle import tebd_ref   # checker and CPU baseline only
    ctx = ens.ctx
    Bs = [ctx.get_site(0, i) for i in range(L)]
    Ss = [ctx.get_S(0, b) for b in range(L + 1)]
    psi = tebd_ref.MPS([None] * L, Bs, Ss, [(0.0, 1.0)] * L)
    kick = np.asarray(ens.kick[0])
    gates = [ens.gates[0, i] for i in range(L - 1)]
    trunc = dict(chi_max=chi, svd_min=1e-12, trunc_cut=1e-7)
    t0 = time.time()
    psi2, _ = tebd_ref.floquet_step(psi, kick, gates, mode='tebd', trunc=trunc)
    t_cpu = time.time() - t0
    t0 = time.time()
    ctx.floquet_step(1)
    ctx.sync()
    t_gpu = time.time() - t0
    rdm, ent = ctx.measure()
    z = rdm[0, :, 0] - rdm[0, :, 1]
    return {'cpu_oracle_s_per_period_chain0': round(t_cpu, 3), 'gpu_s_per_period_whole_ensemble': round(t_gpu, 4),
            'max_abs_dZ': float(np.max(np.abs(z - tebd_ref.site_z(psi2)))),
            'max_abs_dS': float(np.max(np.abs(ent[0] - psi2.entanglement_entropy()))),
            'chi_equal': bool(list(ctx.chi()[0][1:-1]) == list(psi2.chi))}
