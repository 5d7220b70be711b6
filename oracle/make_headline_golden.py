#!/usr/bin/env python3
"""
Golden records of the LAPACK TEBD oracle (oracle/tebd_ref.py) at the sizes the benchmark and the BASELINE
configurations are quoted on, where the oracle takes tens of seconds per chain: generated here once, committed
under tests/golden/headline_*.npz, compared with the CUDA path by tests/test_gpu_headline.py (1e-8 on <Z_i>(t),
bond entropies and the Loschmidt echo; bond-dimension tables equal).

TEST INFRASTRUCTURE.  Needs only NumPy/SciPy (the oracle), not /root/reference:

    python oracle/make_headline_golden.py [case ...]

The GPU tests re-run the oracle instead of reading the files when TC_GOLDEN_LIVE=1.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import tebd_ref  # noqa: E402

# name -> parameters.  `schedule` = [(epsilon, periods), ...]; fields = np.random.seed(seed); uniform(-W, W, L)
CASES = {
    # bench.py's workload (BASELINE metric shape): entangle at eps = 0.3 until the central bonds sit at chi_max = 128
    # (period 8), then the target eps = 0.1; theta is 256 x 256 with truncation active on the central bonds
    'headline_L32_chi128': dict(L=32, W=0.3, seeds=[1000, 1001], schedule=[(0.3, 9), (0.1, 3)],
                                trunc=dict(chi_max=128, svd_min=1e-12, trunc_cut=1e-7)),
    # BASELINE config 4 regime: chi_max = 256 binding (theta 512 x 512, the wide cluster kernels) on a chain the
    # oracle can still afford
    'wide_L18_chi256': dict(L=18, W=0.3, seeds=[11], schedule=[(0.3, 11)],
                            trunc=dict(chi_max=256, svd_min=1e-12, trunc_cut=1e-7)),
    # BASELINE config 3 shape: L = 24, chi_max = 64, one grid point, 20 periods at eps = 0.1
    'c3_L24_chi64': dict(L=24, W=0.3, seeds=[7], schedule=[(0.1, 20)],
                         trunc=dict(chi_max=64, svd_min=1e-12, trunc_cut=1e-7)),
    # sub-harmonic response of an entangling run: 64 periods at eps = 0.1, L = 12, chi_max = 32 (FFT bin index test)
    'dtc_L12_chi32': dict(L=12, W=0.3, seeds=[42, 43], schedule=[(0.1, 64)],
                          trunc=dict(chi_max=32, svd_min=1e-12, trunc_cut=1e-10)),
}


def generate(name):
    c = CASES[name]
    out = {}
    t0 = time.time()
    for k, seed in enumerate(c['seeds']):
        h = tebd_ref.disorder_fields(c['L'], c['W'], seed)
        r = tebd_ref.run_schedule(c['L'], 1.0, h, 1.0, c['schedule'], state='neel', up_index=1, mode='tebd',
                                  trunc=c['trunc'])
        for key in ('Z', 'S_ent', 'LE', 'chi'):
            out[f'{key}_{k}'] = r[key]
        out[f'h_{k}'] = h
        # smallest relative gap of the Schmidt spectrum at the cut over the run is what limits the agreement of two
        # correct SVDs once chi_max binds; recorded for the tolerance discussion in DESIGN.md
        print(f'  {name} seed {seed}: chi max {r["chi"].max()}, discarded weight {r["trunc_err"]:.3e}, '
              f'{time.time() - t0:.1f} s', flush=True)
    path = os.path.join(ROOT, 'tests', 'golden', name + '.npz')
    np.savez_compressed(path, **out)
    print(f'{path}: {os.path.getsize(path)} bytes')


if __name__ == '__main__':
    for n in (sys.argv[1:] or list(CASES)):
        generate(n)
