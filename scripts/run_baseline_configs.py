"""Runs the BASELINE.json configurations at their named sizes on one GPU and prints one JSON summary per config
(wall time, throughput, bond dimensions, a physics fingerprint).  These are completeness runs, not bench lines
(bench.py measures the headline metric); the parity of every code path they use is in tests/.

    python scripts/run_baseline_configs.py [2 3 4 5] [--quick] [--cpu]

Every ensemble line carries the algorithmic FP64 work of the run (SURVEY 8d flop model on the recorded bond dimensions)
as TFLOP/s and as a fraction of the FP64 peak measured in the same process (tc_probe_fp64).  With --cpu the final state
of chain 0 is handed to the CPU oracle (oracle/tebd_ref.py, the checker: one single-threaded process), which evolves it
one more period; the GPU evolves the same period, the two are compared (max |delta| of <Z_i>, entropies) and timed.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from time_crystal_tensor_network_b200 import engine as eng  # noqa: E402

import bench  # noqa: E402  (flop model)

QUICK = '--quick' in sys.argv
CPU = '--cpu' in sys.argv
PEAK = None


def fp64_peak():
    global PEAK
    if PEAK is None:
        PEAK = eng.probe_fp64(0, False) * 1e-3
    return PEAK


def run_flops(chi_rec, periods):
    """Algorithmic flop of a run from its bond-dimension records chi_rec[T][R][L+1] (T records spread evenly over
    `periods` periods): trapezoid over the records of the per-period count of bench.update_flops."""
    T = chi_rec.shape[0]
    per = np.array([sum(sum(bench.update_flops(chi_rec[t, r])) for r in range(chi_rec.shape[1])) for t in range(T)])
    if T == 1:
        return float(per[0] * periods)
    return float(np.sum(0.5 * (per[1:] + per[:-1])) * periods / (T - 1))


def cpu_check(ens, L, chi, eps):
    """One more period of chain 0 on the GPU and on the CPU oracle from the same state (the checker lives with the tests:
    tests/baseline_cpu_check.py -- nothing outside tests/, smoke() and bench.py's CPU legs touches oracle/)."""
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from baseline_cpu_check import one_period_against_oracle
    return one_period_against_oracle(ens, L, chi)


WHICH = [int(a) for a in sys.argv[1:] if a.isdigit()] or [1, 2, 3, 4, 5]


def subharmonic_bin(series):
    """Index of the largest positive-frequency Fourier component of a mean-removed, Hann-windowed series and the bin of
    half the drive frequency (period doubling <=> they coincide)."""
    x = np.asarray(series, dtype=float)
    x = (x - x.mean()) * np.hanning(len(x))
    p = np.abs(np.fft.rfft(x))
    return int(np.argmax(p[1:]) + 1), int(round(len(x) / 2))


def run_ensemble(tag, L, chi, hs, eps, n_periods, measure_every=1, **kw):
    t0 = time.time()
    ens = eng.FloquetEnsemble(L, 1.0, 1.0, hs, epsilon=eps, chi_max=chi, mode='tebd', svd_min=1e-12, trunc_cut=1e-7, **kw)
    t1 = time.time()
    out = ens.run(n_periods, measure_every=measure_every)
    dt = time.time() - t0
    run_s = time.time() - t1
    R = hs.shape[0]
    stag = (out['Z'] * ((-1.0) ** np.arange(L))).mean(axis=2)          # [T][R] staggered magnetisation
    res = {'config': tag, 'L': L, 'chi_max': chi, 'chains': R, 'periods': n_periods, 'wall_s': round(dt, 2),
           'setup_s': round(t1 - t0, 2), 'run_s': round(run_s, 3),   # setup: context, arena, model upload (the first one also CUDA start-up)
           'chain_steps_per_s': round(R * n_periods / dt, 1), 'chi_reached': int(out['chi'].max()),
           'S_mid_final_mean': float(out['S_ent'][-1][:, L // 2 - 1].mean()),
           'LE_final_mean': float(out['LE'][-1].mean()), 'flags': {k: float(v) for k, v in out['flags'].items()}}
    fl = run_flops(out['chi'], n_periods)
    res['algorithmic_tflop'] = round(fl * 1e-12, 3)
    res['tflops'] = round(fl / run_s * 1e-12, 3) if run_s > 0 else None
    res['frac_of_fp64_peak'] = round(fl / run_s * 1e-12 / fp64_peak(), 4) if run_s > 0 else None
    res['fp64_peak_tflops'] = round(fp64_peak(), 2)
    if CPU and chi > 1 and out['chi'].max() > 1:
        res['cpu_check'] = cpu_check(ens, L, chi, eps)
    ens.close()
    return res, out, stag


results = []


def emit(res):
    results.append(res)
    print(json.dumps(res), flush=True)


if 1 in WHICH:
    # clean kicked Ising, L = 10, chi = 32, 100 periods (main.py's perfect time crystal): chi stays 1, exact alternation
    res, out, stag = run_ensemble('1: perfect time crystal', 10, 32, np.zeros((1, 10)), 0.0, 100)
    res['stag_alternates_exactly'] = bool(np.allclose(stag[1:, 0], -stag[:-1, 0], atol=1e-12) and
                                          np.allclose(np.abs(stag[:, 0]), 1.0, atol=1e-12))
    res['subharmonic_bin,half_drive_bin'] = subharmonic_bin(stag[:, 0])
    emit(res)
    # the same once more: the first line carries the process's CUDA start-up and module load
    res2, _, _ = run_ensemble('1: perfect time crystal (second run in the process)', 10, 32, np.zeros((1, 10)), 0.0, 100)
    emit(res2)
if 2 in WHICH:
    L, R, n = 20, 256, (40 if QUICK else 200)
    hs = np.array([eng.disorder_fields(L, 0.3, 1000 + r) for r in range(R)])
    res, out, stag = run_ensemble('2: disordered DTC', L, 64, hs, 0.1, n)
    res['disorder_avg_stag_first,last'] = [float(stag[0].mean()), float(stag[-1].mean())]
    res['subharmonic_bin,half_drive_bin'] = subharmonic_bin(stag.mean(axis=1))
    emit(res)
if 3 in WHICH:
    L, n, G = 24, (20 if QUICK else 80), (8 if QUICK else 32)
    eps_grid, w_grid = np.linspace(0.0, 0.3, G), np.linspace(0.0, 0.8, G)
    pts = [(e, w) for e in eps_grid for w in w_grid]
    hs = np.array([eng.disorder_fields(L, w, 42) for _, w in pts])
    eps = np.array([e for e, _ in pts])
    res, out, stag = run_ensemble(f'3: phase diagram {G}x{G}', L, 64, hs, eps, n)
    peak = np.array([subharmonic_bin(out['LE'][:, r])[0] for r in range(len(pts))]).reshape(G, G)
    res['LE_final_grid_corners'] = [float(out['LE'][-1][i]) for i in (0, G - 1, G * (G - 1), G * G - 1)]
    res['LE_peak_bin_histogram'] = {int(k): int(v) for k, v in zip(*np.unique(peak, return_counts=True))}
    emit(res)
if 4 in WHICH:
    L, n = 64, (30 if QUICK else 500)
    hs = np.array([eng.disorder_fields(L, 0.3, 1000)])
    res, out, stag = run_ensemble('4: long chain', L, 256, hs, 0.1, n, measure_every=10)
    res['chi_mid_trajectory'] = [int(c) for c in out['chi'][:, 0, L // 2][:: max(1, len(out['chi']) // 10)]]
    emit(res)
if 5 in WHICH:
    from time_crystal_tensor_network_b200.dynamics.tebd_evolution import TEBDEvolution, CustomFloquet
    from time_crystal_tensor_network_b200.models.kicked_ising import KickedIsingModel
    from time_crystal_tensor_network_b200.core.tensor_utils import create_initial_state
    import scipy.linalg as sl
    L, g = (40 if QUICK else 100), 0.7
    X, Z, I2 = np.array([[0, 1], [1, 0]], dtype=complex), np.diag([1.0, -1.0]).astype(complex), np.eye(2)

    class TFIM:
        H_bond = [-np.kron(Z, Z) - g * ((1.0 if i == 0 else 0.5) * np.kron(X, I2) + (1.0 if i == L - 2 else 0.5) * np.kron(I2, X))
                  for i in range(L - 1)]
    t0 = time.time()
    te = TEBDEvolution(TFIM(), max_chi=128)
    psi0 = create_initial_state(L, 'all_up')
    rot = np.array([[np.cos(0.3), -np.sin(0.3)], [np.sin(0.3), np.cos(0.3)]])
    for i in range(L):
        psi0.apply_local_op(i, rot, unitary=True)
    gs, info = te.imaginary_time_evolution(psi0, dts=(0.1, 0.02), steps_per_dt=(20 if QUICK else 60))
    t_gs = time.time() - t0
    model = KickedIsingModel(L, 1.0, 0.3, 1.0, disorder_seed=5)
    model.pi_pulse_gate = sl.expm(-1j * np.pi / 2 * 0.9 * model.sigma_x)
    model.truncation = 'tebd'          # make chi_max / svd_min / trunc_cut effective (the reference ignores trunc_params)
    t1 = time.time()
    n_q = 6 if QUICK else 20
    states, times, finfo = CustomFloquet(model, dict(chi_max=128, svd_min=1e-12, trunc_cut=1e-7)).evolve_floquet(gs, n_q)
    t_q = time.time() - t1
    S = np.array([s.entanglement_entropy() for s in states])
    e_exact = -0.5 * np.sum(np.linalg.svd(2 * g * np.eye(L) - 2 * np.eye(L, k=1), compute_uv=False))   # free fermions
    emit({'config': '5: imaginary-time ground state + Floquet quench', 'L': L, 'chi_max': 128,
                    'ground_state_s': round(t_gs, 2), 'energy_per_site': float(info['energies'][-1] / L),
                    'exact_energy_per_site': float(e_exact / L),
                    'energy_decreasing': bool(info['energies'][0] >= info['energies'][-1]),
                    'gs_max_chi': int(max(gs.chi)), 'quench_periods': n_q, 'quench_s': round(t_q, 2),
                    'S_mid_trajectory': [float(x) for x in S[:, L // 2 - 1]],
                    'max_chi_after_quench': int(finfo['final_bond_dim'])})
