"""Runs the REFERENCE's own, unmodified test suite (tests/test_basic_functionality.py, test_physics_validation.py,
test_performance.py of connor-a-casey/time-crystal-tensor-network) against this repository's drop-in modules on the GPU.

usage: python scripts/run_reference_testsuite.py /path/to/reference/tests [pytest args]

The reference's tests put ``<tests>/../src`` on sys.path and import ``main`` from the directory above; a scratch
directory is laid out the same way with the test files COPIED there (never into this repository) and ``src``,
``main.py``, ``config.txt`` linked to this repository's.  matplotlib is not part of this image: if it is missing, a
stub package (``use``, ``pyplot.show/savefig/subplots`` as mocks) stands in -- the tests only patch those calls.
Result on the B200 (profiles/r02t_reference_testsuite.txt): 39 passed."""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    if len(sys.argv) < 2 or not os.path.isdir(sys.argv[1]):
        sys.exit(__doc__)
    src_tests = os.path.abspath(sys.argv[1])
    work = tempfile.mkdtemp(prefix='tc_reftests_')
    os.makedirs(os.path.join(work, 'tests'))
    for f in os.listdir(src_tests):
        if f.endswith('.py') and (f.startswith('test_') or f == '__init__.py'):
            shutil.copy(os.path.join(src_tests, f), os.path.join(work, 'tests', f))
    for name in ('src', 'main.py', 'config.txt'):
        os.symlink(os.path.join(ROOT, name), os.path.join(work, name))
    path = [work]
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        stub = os.path.join(work, 'stubs', 'matplotlib')
        os.makedirs(stub)
        with open(os.path.join(stub, '__init__.py'), 'w') as f:
            f.write('rcParams = {}\n\n\ndef use(*a, **k):\n    pass\n')
        with open(os.path.join(stub, 'pyplot.py'), 'w') as f:
            f.write('from unittest.mock import MagicMock\nshow = MagicMock()\nsavefig = MagicMock()\nclose = MagicMock()\n\n\n'
                    'def subplots(*a, **k):\n    return MagicMock(), MagicMock()\n\n\ndef __getattr__(name):\n    return MagicMock()\n')
        path.insert(0, os.path.join(work, 'stubs'))
    env = dict(os.environ, PYTHONPATH=os.pathsep.join(path + [os.environ.get('PYTHONPATH', '')]))
    rc = subprocess.call([sys.executable, '-m', 'pytest', 'tests', '-q', '-p', 'no:cacheprovider'] + sys.argv[2:], cwd=work, env=env)
    shutil.rmtree(work, ignore_errors=True)
    sys.exit(rc)


if __name__ == '__main__':
    main()
