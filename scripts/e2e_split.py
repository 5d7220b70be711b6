"""Where the gap between the device-resident value and the end-to-end value of bench.py comes from (metric shape).
Wall-clock around synchronised calls, median of 3.  python scripts/e2e_split.py"""
import sys, time, numpy as np
sys.path.insert(0, '.')
from time_crystal_tensor_network_b200 import engine as eng
R, L, chi = 32, 32, 128
hs = np.array([eng.disorder_fields(L, 0.3, 1000 + r) for r in range(R)])
ens = eng.FloquetEnsemble(L, 1.0, 1.0, hs, epsilon=0.3, chi_max=chi, mode='tebd', svd_min=1e-12, trunc_cut=1e-7)
ctx = ens.ctx
ctx.floquet_step(9)
kick = np.ascontiguousarray(np.broadcast_to(eng.kick_matrix(0.1), (R, 2, 2)))
ctx.set_model(ens.gates, kick)
ctx.floquet_step(3); ctx.sync()
ctx.run_host(1, 1, False, gates=ens.gates, kick=kick)


def wall(f, n=3):
    ts = []
    for _ in range(n):
        ctx.sync(); t0 = time.perf_counter(); f(); ctx.sync(); ts.append(time.perf_counter() - t0)
    return float(np.median(ts)) * 1e3


def a():
    for _ in range(8):
        ctx.floquet_step(1); ctx.sync()


def b():
    for _ in range(8):
        ctx.run_host(1, 1, False, gates=ens.gates, kick=kick)


def b2():
    for _ in range(8):
        ctx.run_host(1, 1, False)


print('8 x [floquet_step(1); sync]           %8.2f ms / period' % (wall(a) / 8), flush=True)
print('8 x floquet_step(1), no sync between     %8.2f ms / period' % (wall(lambda: [ctx.floquet_step(1) for _ in range(8)]) / 8), flush=True)
print('8 x run_host(1) with model upload      %8.2f ms / period' % (wall(b) / 8), flush=True)
print('8 x run_host(1) no upload              %8.2f ms / period' % (wall(b2) / 8), flush=True)
print('floquet_step(8)                        %8.2f ms / period' % (wall(lambda: ctx.floquet_step(8)) / 8), flush=True)
print('run_host(8, every 1)                   %8.2f ms / period' % (wall(lambda: ctx.run_host(8, 1, False)) / 8), flush=True)
print('run_host(8, every 8)                   %8.2f ms / period' % (wall(lambda: ctx.run_host(8, 8, False)) / 8), flush=True)
print('run_host(0, measure now)               %8.2f ms / snapshot' % wall(lambda: ctx.run_host(0, 1, True)), flush=True)
if '--profile' in sys.argv:
    # per-class times with records every period and with one record (profile mode: one stream, kernels serialised)
    ctx.profile(True)
    for every in (1, 8):
        ctx.profile_read(reset=True)
        ctx.run_host(8, every, False)
        p = ctx.profile_read(reset=True)
        print('profile, run_host(8, every %d): ' % every + '  '.join('%s %.2f ms/%d' % (k, v[0], v[1]) for k, v in p.items()), flush=True)
    ctx.profile(False)
for want in ((), ('Z',), ('ent',), ('ov',), ('chi',), ('Z', 'ent', 'ov', 'chi')):
    print('run_host(8, every 1, want=%-28s %8.2f ms / period' % (str(want) + ')', wall(lambda: ctx.run_host(8, 1, False, want=want)) / 8), flush=True)
