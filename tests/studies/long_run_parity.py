"""How the GPU / oracle deviation grows over hundreds of periods (needs the GPU; not a pytest test): three small chains,
300-400 periods, max |dZ|, |dS|, |dLE| after 10, 50, 100, 200 and all periods.  B200, round 2: 1e-14 ... 7e-14 after 400
periods where the truncation is mild; 8e-11 after 300 periods at L = 12, chi_max = 32, trunc_cut 1e-7 (a truncating
TEBD iteration amplifies rounding differences: 1e-14 at 50 periods, 5e-14 at 100, 2e-12 at 200).
usage: python tests/studies/long_run_parity.py"""
import sys, time, numpy as np
sys.path.insert(0, '.')
from oracle import tebd_ref
from time_crystal_tensor_network_b200.engine import FloquetEnsemble
for (L, chi, eps, W, n, cut) in ((10, 16, 0.1, 0.3, 400, 1e-10), (12, 32, 0.15, 0.5, 300, 1e-7), (10, 32, 0.1, 0.3, 400, 1e-10)):
    h = tebd_ref.disorder_fields(L, W, 4242)
    trunc = dict(chi_max=chi, svd_min=1e-12, trunc_cut=cut)
    t0 = time.time()
    ref = tebd_ref.run(L, 1.0, h, 1.0, n, epsilon=eps, mode='tebd', trunc=trunc)
    t1 = time.time()
    ens = FloquetEnsemble(L, 1.0, 1.0, h[None, :], epsilon=eps, mode='tebd', **trunc)
    out = ens.run(n)
    ens.close()
    dz = np.abs(out['Z'][:, 0] - ref['Z']).max(axis=1)
    ds = np.abs(out['S_ent'][:, 0] - ref['S_ent']).max(axis=1)
    dl = np.abs(out['LE'][:, 0] - ref['LE'])
    chi_eq = np.array_equal(out['chi'][:, 0, 1:-1], ref['chi'])
    print(f'L={L} chi_max={chi} eps={eps} W={W} cut={cut:g} {n} periods (oracle {t1-t0:.0f} s): chi tables equal {chi_eq}, chi reached {ref["chi"].max()}')
    for t in (10, 50, 100, 200, n):
        print(f'   after {t:3d} periods: max dZ {dz[:t+1].max():.1e}  dS {ds[:t+1].max():.1e}  dLE {dl[:t+1].max():.1e}')
