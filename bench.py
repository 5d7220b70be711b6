#!/usr/bin/env python3
"""Benchmark of the hot path: disorder-averaged Floquet steps/sec at L=32, chi=128, FP64.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...        # CPU arm: the oracle port on the box's host cores

One "step" = one Floquet period (62 two-site updates with their SVDs + 32 kicks, the reference's
sequence, src/models/kicked_ising.py:100-160) of every chain of the ensemble.  Weak scaling (default): every
GPU evolves --chains independent disorder realisations (32 per GPU = 256 on 8 GPUs, BASELINE config);
--scaling strong: --total-chains (256) realisations split over the GPUs.
Both arms run the same schedule per chain: --prep-periods periods at --prep-eps from the Neel state (the central
bonds reach chi_max), --warmup periods at --eps, then exactly --steps timed periods at --eps.
Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'disorder-avg Floquet steps/sec (L=32, chi=128, FP64)'
UNIT = 'chain-steps/s'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=8)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--chains', type=int, default=32, help='chains per GPU')
    ap.add_argument('--L', type=int, default=32)
    ap.add_argument('--chi', type=int, default=128)
    ap.add_argument('--eps', type=float, default=0.1)
    ap.add_argument('--prep-eps', type=float, default=0.3, help='kick imperfection used to entangle the state')
    ap.add_argument('--prep-periods', type=int, default=10, help='preparation periods at --prep-eps (both arms)')
    ap.add_argument('--cpu-periods', type=int, default=2, help='timed periods of the cpu_baseline sample (GPU arm)')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'])
    ap.add_argument('--total-chains', type=int, default=256, help='ensemble size with --scaling strong')
    ap.add_argument('--no-extras', action='store_true', help='skip the extra measurements of the N = 1 line')
    return ap.parse_args()


WORK = dict(J=1.0, tau=1.0, W=0.3, svd_min=1e-12, trunc_cut=1e-7, state='neel', seed0=1000)


def chains_of(a, world):
    """(chains per GPU, total chains) of this run."""
    if a.scaling == 'strong':
        return -(-a.total_chains // world), a.total_chains
    return a.chains, a.chains * world


def make_config(a, world):
    """The workload both arms run; the same keys and values in the GPU line and in the reference line."""
    per, total = chains_of(a, world)
    return {
        'workload': (f'L{a.L}_chi{a.chi}_neel_W{WORK["W"]}_eps{a.eps}_tebd_svdmin1e-12_trunccut1e-7_'
                     f'prep{a.prep_periods}x_eps{a.prep_eps}'),
        'L': a.L, 'chi_max': a.chi, 'eps': a.eps, 'prep_eps': a.prep_eps, 'prep_periods': a.prep_periods,
        'chains_total': total, 'chains_per_gpu': per, 'svds_per_step_per_chain': 2 * (a.L - 1),
        'seeds': f'{WORK["seed0"]}..{WORK["seed0"] + total - 1}',
        'l2': 'working set (17 MB of site tensors per chain + 2 x 1 MiB workspaces per update) far larger than the '
              '126 MB L2, no flush',
        'parallelism': f'independent chains sharded over {world} GPU(s), final gather only',
    }


# ------------------------------------------------------------------------------------------------
# algorithmic work: per update with theta m x n (m >= n), inner dim k, kept chi'
# ------------------------------------------------------------------------------------------------
def update_flops(chi_row):
    """chi_row: int [L+1] bond dimensions.  Returns (F_theta, F_svd, F_B) summed over the 2(L-1) updates
    of one period (every bond is updated twice)."""
    ft = fs = fb = 0.0
    L = len(chi_row) - 1
    for i in range(L - 1):
        cl, cm, cr = int(chi_row[i]), int(chi_row[i + 1]), int(chi_row[i + 2])
        M, N = 2 * cl, 2 * cr
        m, n = max(M, N), min(M, N)
        ft += 8.0 * M * cm * N
        fs += 4.0 * (4.0 * m * m * n + 8.0 * m * n * n + 9.0 * n ** 3)
        fb += 8.0 * M * N * cm
    return 2 * ft, 2 * fs, 2 * fb


# ------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------------
class Clocks:
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '200', '-i', str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) > 8 and r[1].replace('.', '').isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) > 8 and r[2].replace('.', '').isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) > 8:
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[5:9]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (NumPy/LAPACK restatement of the reference's TeNPy path)
# ------------------------------------------------------------------------------------------------
def _cpu_worker(task):
    """Prepare one chain on the CPU (oracle O2, TEBD truncation) and time `n_timed` periods."""
    seed, L, chi, eps, prep_eps, n_prep, n_warm, n_timed, threads = task
    sys.path.insert(0, ROOT)
    from oracle import tebd_ref
    h = tebd_ref.disorder_fields(L, WORK['W'], seed)
    trunc = dict(chi_max=chi, svd_min=WORK['svd_min'], trunc_cut=WORK['trunc_cut'])
    kick_p, gates = tebd_ref.make_gates(L, WORK['J'], h, WORK['tau'], prep_eps)
    kick, _ = tebd_ref.make_gates(L, WORK['J'], h, WORK['tau'], eps)
    psi = tebd_ref.product_state(L, 'neel', 1)
    for _ in range(n_prep):
        psi, _ = tebd_ref.floquet_step(psi, kick_p, gates, mode='tebd', trunc=trunc)
    for _ in range(n_warm):
        psi, _ = tebd_ref.floquet_step(psi, kick, gates, mode='tebd', trunc=trunc)
    t0 = time.perf_counter()
    for _ in range(n_timed):
        psi, _ = tebd_ref.floquet_step(psi, kick, gates, mode='tebd', trunc=trunc)
    dt = time.perf_counter() - t0
    return dt, n_prep, max(psi.chi)


def cpu_ensemble_rate(a, n_warm, n_timed, procs=None):
    """Best-case CPU throughput for independent chains: one single-threaded process per host core, one chain each
    (the core policy of both the cpu_baseline leg and the reference arm: every core of the box), the chains of seeds
    seed0 .. seed0 + cores - 1 on the same per-chain schedule as the GPU arm."""
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    procs = procs or cores
    tasks = [(WORK['seed0'] + r, a.L, a.chi, a.eps, a.prep_eps, a.prep_periods, n_warm, n_timed, 1) for r in range(procs)]
    # spawn (not fork) with single-threaded BLAS set through the environment: a forked child inherits an
    # OpenBLAS pool sized for all cores and 8 such children oversubscribe the box 8x
    keep = {k: os.environ.get(k) for k in ('OPENBLAS_NUM_THREADS', 'OMP_NUM_THREADS', 'MKL_NUM_THREADS')}
    os.environ.update({k: '1' for k in keep})
    t0 = time.perf_counter()
    try:
        with mp.get_context('spawn').Pool(procs) as pool:
            res = pool.map(_cpu_worker, tasks)
    finally:
        for k, v in keep.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    wall = time.perf_counter() - t0
    slowest = max(r[0] for r in res)
    return {'value': procs * n_timed / slowest, 'cores': procs, 'ms_per_step': slowest / n_timed * 1e3,
            'chi_max': max(r[2] for r in res), 'wall_s': wall}


def run_reference(a):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    world = int(os.environ.get('WORLD_SIZE', 1))
    # exactly --steps timed periods after --warmup untimed ones, per chain, as the GPU arm
    r = cpu_ensemble_rate(a, a.warmup, a.steps)
    sample = (f"{r['cores']} chains (seeds {WORK['seed0']}..{WORK['seed0'] + r['cores'] - 1}, one single-threaded process "
              f"per host core, {os.cpu_count()} cores) x {a.steps} timed periods after {a.prep_periods} preparation and "
              f"{a.warmup} warm-up periods, chi_max reached {r['chi_max']}; oracle/tebd_ref.py (NumPy + LAPACK zgesdd), the "
              f"port of the reference's TeNPy path (TeNPy is not installable here)")
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': r['value'], 'unit': UNIT, 'n_gpus': a.gpus,
        'steps': a.steps, 'warmup': a.warmup, 'ms_per_step': r['ms_per_step'], 'higher_is_better': True,
        'scaling': a.scaling, 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': make_config(a, world),
        'cpu_baseline': {'value': r['value'], 'unit': UNIT, 'cores': r['cores'], 'kind': 'port', 'sample': sample},
        'e2e': {'value': r['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'wall_s': round(r['wall_s'], 1),
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def measure_ensemble(a, R, seed_lo, local, barrier, full=True):
    """Prepare R chains (seeds seed_lo ..) on GPU `local` with the common schedule and measure: the device-resident
    timed region (`ms` for a.steps periods), and with full=True the per-kernel pass and the end-to-end loop, all three
    from the same saved state (the work of a period grows with the entanglement of the state)."""
    import torch
    from time_crystal_tensor_network_b200 import engine as eng
    L = a.L
    hs = np.array([eng.disorder_fields(L, WORK['W'], seed_lo + r) for r in range(R)])
    ens = eng.FloquetEnsemble(L, WORK['J'], WORK['tau'], hs, epsilon=a.prep_eps, chi_max=a.chi, mode='tebd',
                              svd_min=WORK['svd_min'], trunc_cut=WORK['trunc_cut'], state=WORK['state'], device=local)
    ctx = ens.ctx
    # ---- state preparation (untimed): the same fixed number of periods at prep_eps as the CPU arm
    t_prep = time.perf_counter()
    ctx.floquet_step(a.prep_periods)
    ctx.sync()
    t_prep = time.perf_counter() - t_prep
    kick = np.ascontiguousarray(np.broadcast_to(eng.kick_matrix(a.eps), (R, 2, 2)))
    ctx.set_model(ens.gates, kick)
    # ---- device-resident timing: W warm-up steps, then exactly K steps.  L2: one step streams the site tensors
    # (17 MB per chain) plus 2 x 1 MiB workspaces per update, far more than the 126 MB L2; no explicit flush is needed.
    for _ in range(a.warmup):
        ctx.floquet_step(1)
    barrier()
    out = {'prep_s': round(t_prep, 2)}
    snap = ctx._arena.clone() if full else None
    chi_start = ctx.chi()
    launches0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with torch.cuda.stream(ctx.stream):
        e0.record(ctx.stream)
        ctx.floquet_step(a.steps)     # the product path: chain groups on their own streams, joined on ctx.stream
        e1.record(ctx.stream)
    barrier()
    out['ms'] = e0.elapsed_time(e1)
    out['launches'] = eng.launch_count() - launches0
    chi_now = ctx.chi()
    out['flags'] = ctx.flags()
    out['chi_mid_min'] = int(chi_now[:, L // 2].min())
    out['chi_mean'] = float(chi_now[:, 1:-1].mean())
    ft = fs = fb = 0.0
    for r in range(R):
        for chi_r in (chi_start[r], chi_now[r]):       # mean of the first and the last period's bond dimensions
            x = update_flops(chi_r)
            ft, fs, fb = ft + 0.5 * x[0], fs + 0.5 * x[1], fb + 0.5 * x[2]
    out['flops'] = (ft, fs, fb)
    if full:
        # ---- kernel pass: the same K steps once more with per-kernel-class CUDA events.  The events need every launch
        # in one stream, so the engine runs the chain groups one after the other here: these are the durations of each
        # kernel alone on the GPU, which is what the roofline fraction of the dominant kernel is about.
        ctx._arena.copy_(snap)
        torch.cuda.synchronize()
        ctx.profile(True)
        ctx.profile_read(reset=True)
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(ctx.stream):
            p0.record(ctx.stream)
            ctx.floquet_step(a.steps)
            p1.record(ctx.stream)
        torch.cuda.synchronize()
        out['prof_ms'] = p0.elapsed_time(p1)
        out['prof'] = ctx.profile_read(reset=True)
        ctx.profile(False)
        # ---- end to end through the host-buffer C-ABI call: model upload + run + observable download per step
        ctx._arena.copy_(snap)
        torch.cuda.synchronize()
        del snap
        rec = ctx.run_host(0, 1, True, gates=ens.gates, kick=kick)   # untimed, no period: allocates the record buffers
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.steps):                                      # K synchronous calls of one period each
            rec = ctx.run_host(1, 1, False, gates=ens.gates, kick=kick)
        torch.cuda.synchronize()
        out['e2e_ms'] = (time.perf_counter() - t0) / a.steps * 1e3
        out['h2d'] = ens.gates.nbytes + kick.nbytes
        out['d2h'] = sum(v.nbytes for v in rec.values() if v is not None)
        out['rec'] = rec
    ens.close()
    return out


def shard_equivalence(world, rank, local):
    """N ranks evolve their shards of 2N small chains and all-gather the records; rank 0 evolves all of them alone:
    the gathered <Z_i>(t), entropies, Loschmidt echo and bond dimensions must be bit-identical (main.py:467-469 is the
    loop being sharded)."""
    from time_crystal_tensor_network_b200 import engine as eng
    from time_crystal_tensor_network_b200.sharding import run_sharded_ensemble
    L, n = 12, 6
    R = 2 * world + 1                                                  # ragged shards
    hs = np.array([eng.disorder_fields(L, 0.3, 3000 + r) for r in range(R)])
    kw = dict(epsilon=0.12, chi_max=16, mode='tebd', svd_min=1e-12, trunc_cut=1e-10)
    got = run_sharded_ensemble(L, 1.0, 1.0, hs, n, rank=rank, world_size=world, device=local, **kw)
    if rank != 0:
        return None
    one = run_sharded_ensemble(L, 1.0, 1.0, hs, n, rank=0, world_size=1, device=local, **kw)
    same = all(np.array_equal(got[k], one[k]) for k in ('Z', 'S_ent', 'LE', 'chi'))
    if not same:
        raise RuntimeError('sharded run differs from the single-GPU run of the same seeds')
    return f'bit-identical Z, S_ent, LE, chi: {R} chains (L={L}, chi_max=16, {n} periods) on {world} ranks vs 1 GPU'


def dropin_api_line():
    """The reference-shaped entry points, timed through the drop-in modules: (i) main.calculate_phase_point as the
    reference's own performance test calls it (tests/test_performance.py:266-273 of the reference: L = 16, 80 periods,
    chi_max = 24, ceiling 60 s; a perfect pulse, so chi stays 1), (ii) CustomFloquet.evolve_floquet on an entangling
    run (L = 16, eps = 0.1, 30 periods, TEBD truncation chi_max = 64): one working context advanced in place, a
    storage-only snapshot per period (src/dynamics/tebd_evolution.py:218-259)."""
    import scipy.linalg
    import main as m
    from time_crystal_tensor_network_b200.models.kicked_ising import KickedIsingModel
    from time_crystal_tensor_network_b200.dynamics.tebd_evolution import CustomFloquet
    from time_crystal_tensor_network_b200.core.tensor_utils import create_initial_state
    params = m.read_parameters(os.path.join(ROOT, 'config.txt'))
    t0 = time.perf_counter()
    res = m.calculate_phase_point(0.2, 2.0, params)
    dt = time.perf_counter() - t0
    out = {'phase_point': {'call': 'main.calculate_phase_point(h=0.2, T=2.0, config.txt: L=16, 80 periods, chi_max=24)',
                           'seconds': round(dt, 3), 'success': bool(res.get('success')), 'reference_test_ceiling_s': 60.0}}
    L, n = 16, 30
    model = KickedIsingModel(L, 1.0, 0.3, 1.0, disorder_seed=42)
    model.pi_pulse_gate = scipy.linalg.expm(-1j * np.pi / 2 * 0.9 * model.sigma_x)
    model.truncation = 'tebd'
    psi0 = create_initial_state(L, 'neel')
    t0 = time.perf_counter()
    states, times, info = CustomFloquet(model, dict(chi_max=64, svd_min=1e-12, trunc_cut=1e-7)).evolve_floquet(psi0, n)
    dt = time.perf_counter() - t0
    out['evolve_floquet'] = {'call': f'CustomFloquet(model, chi_max=64).evolve_floquet(psi0, {n}) at L={L}, eps=0.1',
                             'seconds': round(dt, 3), 'periods_per_s': round(n / dt, 1),
                             'final_bond_dim': int(info['final_bond_dim']), 'snapshots': len(states)}
    return out


def run_ours(a):
    import torch
    import torch.distributed as dist
    from time_crystal_tensor_network_b200 import engine as eng
    from time_crystal_tensor_network_b200.sharding import gather_records, shard_bounds

    world = int(os.environ.get('WORLD_SIZE', 1))
    rank = int(os.environ.get('RANK', 0))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device(f'cuda:{local}'))
    per, total = chains_of(a, world)
    lo, hi = shard_bounds(total, world, rank)
    R, L = hi - lo, a.L

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = Clocks(local)
    if rank == 0:
        clocks.start()
    m = measure_ensemble(a, R, WORK['seed0'] + lo, local, barrier, full=True)
    clk = clocks.stop() if rank == 0 else None

    # ---- max over ranks, final gather of the observables (the only collective)
    t = torch.tensor([m['ms'], m['e2e_ms']], dtype=torch.float64, device=f'cuda:{local}')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max, e2e_ms = float(t[0]), float(t[1])
    Z_all = gather_records(m['rec']['Z'], total, axis=1, device=f'cuda:{local}')
    equiv = shard_equivalence(world, rank, local) if world > 1 else None

    if rank == 0:
        ft, fs, fb = m['flops']
        prof, prof_ms, ms = m['prof'], m['prof_ms'], m['ms']
        fp64_peak = eng.probe_fp64(local, False) * 1e-3          # TFLOP/s, FMA pipe, measured now
        dmma_peak = eng.probe_fp64(local, True) * 1e-3
        svd_ms = prof['jacobi'][0] + prof['qr'][0] + prof['finalize'][0]
        n_svd_launch = max(prof['jacobi'][1], 1)
        jac_ms = prof['jacobi'][0] / n_svd_launch
        # algorithmic SVD flops of one Jacobi launch = one parity layer of all chains; 4 layers per period
        f_svd_launch = fs * a.steps / n_svd_launch
        achieved = f_svd_launch / (svd_ms / n_svd_launch * 1e-3) * 1e-12 if svd_ms > 0 else None
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            peaks = {}
        value = total * a.steps / (ms_max * 1e-3)
        cfg = make_config(a, world)      # identical, key for key, to the reference arm's `config`
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': a.steps, 'warmup': a.warmup,
            'ms_per_step': ms_max / a.steps, 'higher_is_better': True, 'scaling': a.scaling, 'vs_baseline': None,
            'dtype': 'f64', 'data': 'synthetic',
            'config': cfg,
            'e2e': {'value': total / (e2e_ms * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': int(m['h2d']),
                    'd2h_bytes_per_step': int(m['d2h']),
                    'note': 'tc_floquet_run_host: host gates+kick uploaded, observables (Z, entropies, overlap, chi) '
                            'downloaded every step (K synchronous calls of one period, wall clock, from the same state as the '
                            'timed region); the MPS state stays resident as it does in the reference'},
            'gpu_launches': int(m['launches']),
            'clocks': clk,
            'roofline': {
                'bound': 'fp64', 'kernel': 'qr_blocked_kernel + jacobi_blocked_kernel + finalize_kernel (batched truncated SVD)',
                'achieved': achieved, 'peak': fp64_peak, 'unit': 'TFLOP/s',
                'frac': (achieved / fp64_peak) if achieved else None,
                'traffic': None,
                'traffic_note': 'not measured in this run (ncu cannot wrap the bench); the per-launch DRAM bytes of the last '
                                'ncu --set full capture are in profiles/ (README.md names the file)',
                'peak_source': 'tc_probe_fp64 (DFMA chain, all SMs) measured in this run; MEASURED_PEAKS.json has no '
                               f'FP64 entry (hbm_gbs={peaks.get("hbm_gbs")}); DMMA probe {dmma_peak:.1f} TFLOP/s',
                'flops_model': 'SURVEY 8d: 4(4m^2 n + 8 m n^2 + 9 n^3) per update, summed over the actual bond '
                               'dimensions of every update in the launch',
                'algorithmic_flops_per_launch': f_svd_launch,
                'ms_per_launch': svd_ms / n_svd_launch,
                'jacobi_ms_per_launch': jac_ms,
                'step_share': {k: round(v[0] / prof_ms, 4) for k, v in prof.items() if v[1]},
                'kernel_pass': f'separate pass of the same {a.steps} steps (state restored) with the chain groups run one after '
                               f'the other ({prof_ms / a.steps:.1f} ms per step): per-kernel CUDA events need one stream',
                'whole_step_tflops': (ft + fs + fb) * a.steps / (ms * 1e-3) * 1e-12,
                'chain_groups': int(os.environ.get('TC_GROUPS', 4)),
            },
            'svd_flags': m['flags'],
            'state': {'prep_s': m['prep_s'], 'chi_mid_min': m['chi_mid_min'], 'chi_mean': m['chi_mean']},
            'gathered_Z_shape': list(Z_all.shape),
        }
        if equiv:
            line['shard_equivalence'] = equiv
        extras = {}
        if world == 1 and not a.no_extras:
            # ---- strong-scaling anchor (fixed R = 256): the whole ensemble on this one GPU
            if a.scaling == 'weak' and a.total_chains > R:
                try:
                    torch.cuda.empty_cache()
                    m2 = measure_ensemble(a, a.total_chains, WORK['seed0'], local, barrier, full=False)
                    extras['strong_scaling_anchor'] = {
                        'chains_total': a.total_chains, 'n_gpus': 1,
                        'value': a.total_chains * a.steps / (m2['ms'] * 1e-3), 'unit': UNIT,
                        'ms_per_step': m2['ms'] / a.steps, 'chi_mid_min': m2['chi_mid_min'],
                        'note': 'same schedule and timing as the headline line with all --total-chains realisations on one '
                                'GPU; bench.py --scaling strong --gpus N splits the same ensemble over N GPUs'}
                except Exception as ex:
                    extras['strong_scaling_anchor'] = {'failed': str(ex)}
            try:
                extras['dropin_api'] = dropin_api_line()
            except Exception as ex:
                extras['dropin_api'] = {'failed': str(ex)}
        if extras:
            line['extras'] = extras
        # ---- CPU baseline on this box's host cores (bounded sample)
        if not a.no_cpu:
            try:
                r = cpu_ensemble_rate(a, 0, a.cpu_periods)
                line['cpu_baseline'] = {
                    'value': r['value'], 'unit': UNIT, 'cores': r['cores'], 'kind': 'port',
                    'sample': f"{r['cores']} chains (one single-threaded process per host core, {os.cpu_count()} cores) x "
                              f"{a.cpu_periods} timed periods after the same {a.prep_periods} preparation periods, "
                              f"chi_max reached {r['chi_max']}; oracle/tebd_ref.py (NumPy + LAPACK zgesdd)"}
            except Exception as ex:   # the GPU number must not be lost to a host-side failure
                line['cpu_baseline'] = {'value': None, 'unit': UNIT, 'cores': 0, 'kind': 'port', 'sample': f'failed: {ex}'}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)
