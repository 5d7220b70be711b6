"""
CPU oracles for the kicked-Ising Floquet/TEBD hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``time_crystal_tensor_network_b200`` (the
product) may import from here.  Allowed importers: ``tests/``,
``__graft_entry__.smoke()`` (as the checker) and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs (as the timed CPU baseline).

Contents
--------
statevector.py   O1: exact 2^L state-vector simulator of the reference's gate
                 sequence (``src/models/kicked_ising.py:100-160``).
tebd_ref.py      O2: NumPy/LAPACK restatement of the TeNPy MPS semantics the
                 reference relies on (``apply_local_op``/``from_full``/
                 ``get_theta``/``overlap``/``expectation_value`` ...), plus the
                 TeNPy ``truncate()`` rule for the chi_max-limited TEBD mode.
tenpy_shim/      a minimal ``tenpy`` package façade over tebd_ref.MPS so the
                 UNMODIFIED reference modules under /root/reference/src can be
                 imported in the build container to generate golden vectors
                 (``oracle/make_golden.py`` -> ``tests/golden/``).
device_model.py  NumPy model of the *device* algorithm stages (Gram matrix,
                 Householder tridiagonalisation, implicit QL, back-transform),
                 used by the stage-wise GPU kernel tests.

Parity status: TeNPy itself (physics-tenpy, unpinned ``>=0.10.0`` in the
reference's requirements.txt; evidence points at v1.0.x) is NOT installed and
cannot be installed here, so the TeNPy-internal semantics are restated from
its published algorithm: **parity unpinned at the TeNPy boundary**.  What *is*
pinned: the reference's own Python (gate construction, gate order, observables,
FFT post-processing) is executed unmodified on top of the shim to produce the
golden fixtures, and the MPS oracle is cross-checked against the exact
state-vector oracle O1.
"""
