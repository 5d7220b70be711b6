"""Shard equivalence on hardware (last bullet): the observables of an ensemble sharded over N GPUs
(one process per GPU, NCCL all-gather of the records) are bit-identical to one GPU evolving all the chains.  Skipped on
a box with one GPU; the gather plumbing itself is covered on the CPU by the world-size-2 gloo test."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_gpu_shards_equal_single_gpu(engine, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    from time_crystal_tensor_network_b200 import engine as eng
    from time_crystal_tensor_network_b200.sharding import run_sharded_ensemble
    out = str(tmp_path / 'gathered.npz')
    port = 29600 + (os.getpid() % 2000)
    res = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                          '--master-addr', '127.0.0.1', '--master-port', str(port),
                          os.path.join(ROOT, 'tests', '_multi_gpu_worker.py'), out],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    got = np.load(out)
    L, R, n = 14, 7, 8
    hs = np.array([eng.disorder_fields(L, 0.3, 2000 + r) for r in range(R)])
    one = run_sharded_ensemble(L, 1.0, 1.0, hs, n, rank=0, world_size=1, device=0, epsilon=0.12, chi_max=24,
                               mode='tebd', svd_min=1e-12, trunc_cut=1e-10)
    for k in ('Z', 'S_ent', 'LE', 'chi'):
        assert np.array_equal(got[k], one[k]), k
