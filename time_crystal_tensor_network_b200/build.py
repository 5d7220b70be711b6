"""Builds the CUDA engine (csrc/*.cu -> libtc_b200.so, in-tree) for sm_100a with nvcc.

    python -m time_crystal_tensor_network_b200.build
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libtc_b200.so')
SOURCES = ['tc_engine.cu']
HEADERS = ['tc_common.cuh', 'tc_gemm.cuh', 'tc_theta.cuh', 'tc_jacobi.cuh', 'tc_jacobi_blocked.cuh', 'tc_jacobi_halfwarp.cuh',
           'tc_jacobi_team.cuh', 'tc_observe.cuh']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-shared', '-Xcompiler', '-fPIC']


def _nvcc():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found: the B200 engine cannot be built (there is no CPU fallback)')


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    deps.append(os.path.join(os.path.dirname(HERE), 'include', 'tc_b200.h'))
    return any(os.path.getmtime(p) > t for p in deps if os.path.exists(p))


def build(force=False, verbose=False):
    """Compile the shared library if it is missing or older than its sources."""
    if not force and not is_stale():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + \
          ['-o', LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + ' '.join(cmd) + '\n' + res.stdout + res.stderr)
    if verbose:
        print(res.stdout + res.stderr)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
