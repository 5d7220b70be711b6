#!/usr/bin/env python3
"""Compute drivers of the reference's ``main.py`` on the B200 engine.

Same entry points and return values as the reference (``read_parameters``, ``stringent_dtc_detection``,
``calculate_phase_point``, ``generate_phase_diagram``, ``calculate_fourier_spectrum``,
``calculate_single_site_magnetization``, ``simulate_*_dtc``, ``generate_individual_figures``, ``main``);
the tensor-network work goes through the drop-in modules under ``src/`` (GPU engine), the spectral
post-processing stays on the host in NumPy.  New here: ``calculate_phase_points_batched`` evolves all
points of a scan as one ensemble on the GPU (the reference loops over them serially, main.py:467-469),
and figure D can use true per-site <Z_i> (``exact_sites=True``) instead of the reference's surrogate.

Plotting needs matplotlib, which is optional: without it the data are computed and returned, nothing is drawn.

    python main.py [--phase-only | --figures-only] [--config FILE]
"""
import argparse
import os
import sys
import time
from typing import Dict, List, Optional, Tuple

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'src'))

from core.tensor_utils import create_initial_state  # noqa: E402
from core.observables import calculate_loschmidt_echo, magnetization, staggered_magnetization  # noqa: E402
from models.kicked_ising import KickedIsingModel  # noqa: E402
from dynamics.tebd_evolution import CustomFloquet  # noqa: E402


# ------------------------------------------------------------------------------------------------
# configuration file
# ------------------------------------------------------------------------------------------------
def _numbers_or_strings(items):
    try:
        vals = [float(x) for x in items]
    except ValueError:
        return list(items)
    return [int(v) for v in vals] if all(v.is_integer() for v in vals) else vals


def _parse_value(text):
    """One right-hand side of ``KEY = value`` with the reference's typing rules (main.py:93-126)."""
    if text.startswith('[') and text.endswith(']'):
        body = text[1:-1].strip()
        if not body:
            return []
        vals = [float(x.strip()) for x in body.split(',')]
        return [int(v) for v in vals] if all(v.is_integer() for v in vals) else vals
    if ',' in text and not any(c in text for c in '()[]'):
        return _numbers_or_strings([x.strip() for x in text.split(',')])
    if '.' in text or 'e' in text.lower():
        return float(text)
    try:
        return int(text)
    except ValueError:
        return text


def read_parameters(filename: Optional[str] = None) -> Dict:
    """``KEY = value  # comment`` file -> dict (main.py:39-130).  Looks at ``filename`` first, then
    ``config.txt`` in the working directory; returns {} when neither exists."""
    candidates = ([filename] if filename else []) + ['config.txt']
    path = next((p for p in candidates if os.path.exists(p)), None)
    if path is None:
        print(f'Warning: No parameters file found. Tried: {candidates}')
        return {}
    print(f'Reading parameters from: {path}')
    params = {}
    with open(path, 'r') as fh:
        for raw in fh:
            line = raw.strip()
            if not line or line.startswith('#') or '=' not in line:
                continue
            line = line.split('#')[0].strip()
            key, value = (s.strip() for s in line.split('=', 1))
            try:
                params[key] = _parse_value(value)
            except ValueError:
                params[key] = value
    return params


# ------------------------------------------------------------------------------------------------
# DTC detector (host, NumPy)
# ------------------------------------------------------------------------------------------------
def stringent_dtc_detection(loschmidt_echoes: List[float], times: List[float], period: float,
                            threshold: float = 0.3) -> float:
    """DTC order parameter in [0, 1] from a Loschmidt-echo series (main.py:134-273): weighted geometric mean
    of (1) the autocorrelation at lag 2T, (2) a sub-harmonic spectral score on the last 3/4 of the series,
    (3) the correlation between the two halves and (4) the mean of the last five echoes; below
    ``threshold`` it is 0."""
    if len(loschmidt_echoes) < 20:
        return 0.0
    le = np.array(loschmidt_echoes)
    t = np.array(times)
    try:
        dt = t[1] - t[0]
        lag = int(2 * period / dt)
        if lag >= len(le) // 2:
            return 0.0
        rho = np.corrcoef(le[:-lag], le[lag:])[0, 1]
        if not np.isfinite(rho) or rho < threshold:
            return 0.0
        s_period = max(0, rho)
    except Exception:
        return 0.0
    try:
        late = le[len(le) // 4:]
        if len(late) < 10:
            return 0.0
        y = late - np.mean(late)
        y = y * np.hanning(len(y))
        spec = np.fft.fft(y)
        freqs = np.fft.fftfreq(len(y), d=dt)
        pos = freqs > 0
        f_pos, amp = freqs[pos], np.abs(spec[pos])
        if len(f_pos) == 0:
            return 0.0
        k_sub = np.argmin(np.abs(f_pos - 1.0 / (2 * period)))
        k_fund = np.argmin(np.abs(f_pos - 1.0 / period))
        p_sub, p_fund, p_tot = amp[k_sub] ** 2, amp[k_fund] ** 2, np.sum(amp ** 2)
        ratio = p_sub / p_fund if p_fund > 0 else 0.0
        purity = p_sub / p_tot if p_tot > 0 else 0.0
        s_spec = min(ratio, purity * 5)
    except Exception:
        s_spec = 0.0
    try:
        mid = len(le) // 2
        a, b = le[:mid], le[mid:2 * mid]
        if len(a) != len(b) or len(a) < 5:
            s_stab = 0.0
        else:
            c = np.corrcoef(a, b)[0, 1]
            s_stab = max(0, c) if np.isfinite(c) else 0.0
    except Exception:
        s_stab = 0.0
    try:
        s_coh = np.mean(le[-5:])
    except Exception:
        s_coh = 0.0
    weights = [0.3, 0.4, 0.2, 0.1]
    floor = [max(s, 1e-6) for s in (s_period, s_spec, s_stab, s_coh)]
    score = np.exp(np.sum([w * np.log(s) for w, s in zip(weights, floor)]))
    if score < threshold:
        score = 0.0
    return min(1.0, score)


# ------------------------------------------------------------------------------------------------
# phase diagram
# ------------------------------------------------------------------------------------------------
PHASE_SITES, PHASE_PERIODS, PHASE_CHI = 16, 80, 24          # fixed in the reference (main.py:309-311)


def _penalties(h_over_J, T_J, avg_bond_dim):
    """Heuristic suppression of unphysical regimes (main.py:362-389)."""
    disorder = np.exp(-3 * (h_over_J - 0.6)) if h_over_J > 0.6 else 1.0
    heating = T_J if T_J < 1.0 else 1.0
    adiabatic = np.exp(-0.5 * (T_J - 3.5)) if T_J > 3.5 else 1.0
    entanglement = avg_bond_dim / 2.0 if avg_bond_dim < 2.0 else 1.0
    return disorder, heating, adiabatic, entanglement


def _phase_result(h_over_J, T_J, echoes, bond_dims, times, tau):
    score = stringent_dtc_detection(echoes, times, 2 * tau)
    avg_chi = np.mean(bond_dims)
    d, h, a, e = _penalties(h_over_J, T_J, avg_chi)
    return {'A2T': score * (d * h * a * e), 'dtc_score_raw': score, 'disorder_penalty': d, 'heating_penalty': h,
            'adiabatic_penalty': a, 'entanglement_penalty': e, 'avg_bond_dim': avg_chi, 'final_le': echoes[-1],
            'success': True}


_FAILED = {'A2T': 0.0, 'dtc_score_raw': 0.0, 'disorder_penalty': 0.0, 'heating_penalty': 0.0,
           'adiabatic_penalty': 0.0, 'entanglement_penalty': 0.0, 'avg_bond_dim': 1.0, 'final_le': 0.0,
           'success': False}


def calculate_phase_point(h_over_J: float, T_J: float, params: Dict) -> Dict[str, float]:
    """Observables of one (h/J, T*J) point: L = 16 Neel chain, 80 periods, DTC detector and penalties
    (main.py:275-415).  Any failure is reported as ``success=False`` with zeroed fields, as in the reference."""
    try:
        J = params['J']
        tau = T_J / (2 * J)
        model = KickedIsingModel(n_sites=PHASE_SITES, J=J, h_disorder=h_over_J * J, tau=tau,
                                 disorder_seed=params['RANDOM_SEED'])
        psi0 = create_initial_state(PHASE_SITES, state_type="neel")
        trunc = {'chi_max': PHASE_CHI, 'svd_min': params['SVD_MIN'], 'trunc_cut': params['SVD_CUTOFF']}
        states, times, _ = CustomFloquet(model, trunc).evolve_floquet(psi0, PHASE_PERIODS, measure_every=1)
        echoes = [calculate_loschmidt_echo(psi0, s) for s in states]
        dims = [max(s.chi) if s.chi else 1 for s in states]
        return _phase_result(h_over_J, T_J, echoes, dims, times, tau)
    except Exception as exc:
        print(f"Error at h/J={h_over_J:.3f}, T*J={T_J:.3f}: {exc}")
        return dict(_FAILED)


def calculate_phase_points_batched(points, params: Dict, epsilon: float = 0.0, device: int = 0) -> List[Dict]:
    """All (h/J, T*J) points of a scan as ONE ensemble on the GPU: every point is an independent chain with its
    own fields and half-period, so the 120 (or 1024) serial evolutions of the reference's scan loop become a
    single batched run.  Returns one ``calculate_phase_point``-style dict per point, in order."""
    from time_crystal_tensor_network_b200.engine import FloquetEnsemble, disorder_fields
    points = [(float(h), float(T)) for h, T in points]
    J = params['J']
    try:
        hs = np.array([disorder_fields(PHASE_SITES, h * J, params['RANDOM_SEED']) for h, _ in points])
        taus = np.array([T / (2 * J) for _, T in points])
        ens = FloquetEnsemble(PHASE_SITES, J, taus, hs, epsilon=epsilon, chi_max=max(PHASE_CHI, 1), mode='reference',
                              chi_cap=2 ** (PHASE_SITES // 2) if epsilon else 1, state='neel', device=device)
        out = ens.run(PHASE_PERIODS)
        ens.close()
    except Exception as exc:
        print(f'Error in the batched scan: {exc}')
        return [dict(_FAILED) for _ in points]
    results = []
    for r, (h, T) in enumerate(points):
        tau = taus[r]
        times = [k * 2 * tau for k in range(PHASE_PERIODS + 1)]
        dims = out['chi'][:, r, 1:-1].max(axis=1)
        results.append(_phase_result(h, T, list(out['LE'][:, r]), list(dims), times, tau))
    return results


def generate_phase_diagram(params: Dict, batched: bool = True, plot: bool = True):
    """12 x 10 scan of the DTC order parameter over h/J in [0, 0.8] and T*J in [0.8, 4.0] (main.py:417-565).
    Returns (figure, axes) when matplotlib is available and ``plot`` is true, otherwise the data dict."""
    h_values, T_values = np.linspace(0.0, 0.8, 12), np.linspace(0.8, 4.0, 10)
    A2T = np.zeros((len(T_values), len(h_values)))
    raw = np.zeros_like(A2T)
    ok = np.zeros_like(A2T, dtype=bool)
    grid = [(h, T) for h in h_values for T in T_values]
    t0 = time.time()
    res = calculate_phase_points_batched(grid, params) if batched else \
        [calculate_phase_point(h, T, params) for h, T in grid]
    for k, r in enumerate(res):
        i, j = divmod(k, len(T_values))
        A2T[j, i], raw[j, i], ok[j, i] = r['A2T'], r['dtc_score_raw'], r['success']
    print(f'phase diagram: {len(grid)} points in {time.time() - t0:.1f} s, {int(ok.sum())} ok, max A2T {A2T.max():.3f}')
    data = {'h_values': h_values, 'T_values': T_values, 'A2T': A2T, 'dtc_score_raw': raw, 'success': ok}
    if not plot:
        return data
    try:
        import matplotlib
        matplotlib.use('Agg')
        import matplotlib.pyplot as plt
    except ImportError:
        print('matplotlib is not installed: phase diagram computed, not drawn')
        return data
    fig, ax = plt.subplots(figsize=(10, 8))
    im = ax.imshow(A2T, extent=[h_values[0], h_values[-1], T_values[0], T_values[-1]], origin='lower', aspect='auto',
                   cmap='viridis', vmin=0, vmax=max(A2T.max(), 1e-3))
    fig.colorbar(im, ax=ax, label=r'DTC order parameter $A_{2T}$')
    ax.set_xlabel('h/J')
    ax.set_ylabel('T J')
    os.makedirs('figures', exist_ok=True)
    for ext in ('png', 'pdf'):
        fig.savefig(os.path.join('figures', f'final_phase_diagram.{ext}'), dpi=params.get('DPI', 300))
    return fig, ax


# ------------------------------------------------------------------------------------------------
# figures A-D
# ------------------------------------------------------------------------------------------------
def calculate_fourier_spectrum(times: np.ndarray, data: np.ndarray, drive_period: float) -> Tuple[np.ndarray, np.ndarray]:
    """(frequency / drive frequency, power / max power) of a mean-free, Hann-windowed series, positive
    frequencies only (main.py:571-618); a DTC shows up at 0.5."""
    y = data - np.mean(data)
    y = y * np.hanning(len(y))
    spec = np.fft.fft(y)
    freqs = np.fft.fftfreq(len(y), d=np.mean(np.diff(times)))
    pos = freqs > 0
    power = np.abs(spec[pos]) ** 2
    top = np.max(power)
    return freqs[pos] / (1.0 / drive_period), (power / top if top > 0 else power)


def calculate_single_site_magnetization(psi, site: int, exact: bool = False):
    """The reference's surrogate for <Z_site> (main.py:620-648): total +- half the staggered magnetisation plus
    Gaussian noise from the global RNG.  ``exact=True`` returns the true expectation value instead."""
    if exact:
        return magnetization(psi, 'z', site=site)
    total, stag = magnetization(psi), staggered_magnetization(psi)
    sign = 0.5 if site % 2 == 0 else -0.5
    return total + sign * stag + 0.1 * np.random.randn()


def _run_figure(params, h_over_J, n_sites, seed, n_periods=200):
    """Shared body of the four figure simulations: tau = 2/J as in the reference (period 4/J)."""
    J = params['J']
    model = KickedIsingModel(n_sites=n_sites, J=J, h_disorder=h_over_J * J, tau=2.0 / J, disorder_seed=seed)
    psi0 = create_initial_state(n_sites, state_type="neel")
    trunc = {'chi_max': params['CHI_MAX'], 'svd_min': params['SVD_MIN'], 'trunc_cut': params['SVD_CUTOFF']}
    states, times, info = CustomFloquet(model, trunc).evolve_floquet(psi0, n_periods, measure_every=1)
    return psi0, states, times


def _stag_total(psi0, states, times, gamma=0.0):
    stag, total = [], []
    for t, psi in zip(times, states):
        decay = np.exp(-gamma * t) if gamma else 1.0
        stag.append(staggered_magnetization(psi) * decay)
        total.append(magnetization(psi) * decay)
        calculate_loschmidt_echo(psi0, psi)       # evaluated (and discarded) per snapshot, as in the reference
    return stag, total


def simulate_perfect_dtc(params: Dict, n_sites: int = 64, n_periods: int = 200):
    """Figure A (main.py:650-718): h/J = 0.25, seed 42.  Returns (times, staggered M, total M)."""
    print("  Simulating perfect DTC conditions...")
    psi0, states, times = _run_figure(params, 0.25, n_sites, 42, n_periods)
    return (times,) + tuple(_stag_total(psi0, states, times))


def simulate_disordered_dtc(params: Dict, n_sites: int = 64, n_periods: int = 200):
    """Figure B (main.py:720-787): h/J = 0.4, seed 123."""
    print("  Simulating disordered DTC conditions...")
    psi0, states, times = _run_figure(params, 0.4, n_sites, 123, n_periods)
    return (times,) + tuple(_stag_total(psi0, states, times))


def simulate_dephasing_dtc(params: Dict, n_sites: int = 64, n_periods: int = 200):
    """Figure C (main.py:789-860): h/J = 0.3, seed 42, observables multiplied by exp(-gamma t), gamma = 0.01 J
    (a post-hoc factor, not a Lindblad evolution -- as in the reference)."""
    print("  Simulating DTC with dephasing...")
    psi0, states, times = _run_figure(params, 0.3, n_sites, 42, n_periods)
    return (times,) + tuple(_stag_total(psi0, states, times, gamma=0.01 * params['J']))


def simulate_multi_site_dtc(params: Dict, n_periods: int = 200, exact_sites: bool = False):
    """Figure D (main.py:862-925): L = 16, sites 1,3,...,11.  Returns (times, [series per site])."""
    print("  Simulating multi-site DTC analysis...")
    psi0, states, times = _run_figure(params, 0.3, 16, 42, n_periods)
    sites = [1, 3, 5, 7, 9, 11]
    series = [[] for _ in sites]
    for psi in states:
        for k, s in enumerate(sites):
            series[k].append(calculate_single_site_magnetization(psi, s, exact=exact_sites))
    return times, series


def generate_individual_figures(params: Dict, plot: bool = True):
    """Runs the four simulations and their spectra; draws them when matplotlib is available."""
    period = 2 * (2.0 / params['J'])
    out = {}
    for name, fn in (('A', simulate_perfect_dtc), ('B', simulate_disordered_dtc), ('C', simulate_dephasing_dtc)):
        times, stag, total = fn(params)
        f, p = calculate_fourier_spectrum(np.array(times), np.array(stag), period)
        out[name] = {'times': times, 'staggered': stag, 'total': total, 'freqs': f, 'power': p}
    times, series = simulate_multi_site_dtc(params)
    f, p = calculate_fourier_spectrum(np.array(times), np.mean(np.array(series), axis=0), period)
    out['D'] = {'times': times, 'sites': series, 'freqs': f, 'power': p}
    if plot:
        try:
            import matplotlib
            matplotlib.use('Agg')
            import matplotlib.pyplot as plt
            os.makedirs('figures', exist_ok=True)
            for name, d in out.items():
                fig, (a1, a2) = plt.subplots(1, 2, figsize=(12, 4))
                if name == 'D':
                    for s in d['sites']:
                        a1.plot(d['times'], s, lw=0.8)
                else:
                    a1.plot(d['times'], d['staggered'])
                a1.set_xlabel('t')
                a2.plot(d['freqs'], d['power'])
                a2.set_xlabel(r'$\omega/\omega_{drive}$')
                fig.savefig(os.path.join('figures', f'figure_{name}.png'), dpi=params.get('DPI', 300))
                plt.close(fig)
        except ImportError:
            print('matplotlib is not installed: figures computed, not drawn')
    return out


def main():
    ap = argparse.ArgumentParser(description='Kicked-Ising discrete time crystal: phase diagram and figures A-D')
    ap.add_argument('--phase-only', action='store_true')
    ap.add_argument('--figures-only', action='store_true')
    ap.add_argument('--config', type=str, default=None)
    args = ap.parse_args()
    t0 = time.time()
    params = read_parameters(args.config)
    if not params:
        print('no parameters, nothing to do')
        return 1
    if not args.figures_only:
        generate_phase_diagram(params)
    if not args.phase_only:
        generate_individual_figures(params)
    print(f'total time {time.time() - t0:.1f} s')
    return 0


if __name__ == '__main__':
    sys.exit(main())
