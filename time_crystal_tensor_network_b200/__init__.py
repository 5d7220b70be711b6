"""B200-native TEBD engine for kicked-Ising discrete-time-crystal simulations.

Drop-in modules (same names and signatures as the reference's ``src/`` tree):
``core.tensor_utils``, ``core.observables``, ``models.kicked_ising``, ``dynamics.tebd_evolution``.
Batched ensemble driver: ``engine.FloquetEnsemble``; multi-GPU sharding: ``sharding``.
"""
from . import _lib
from .engine import Context, FloquetEnsemble, EngineError
from .mps import MPS, SpinHalfSite
from .core import tensor_utils, observables
from .models.kicked_ising import KickedIsingModel
from .dynamics.tebd_evolution import CustomFloquet, TEBDEvolution

__all__ = ['Context', 'FloquetEnsemble', 'EngineError', 'MPS', 'SpinHalfSite', 'KickedIsingModel',
           'CustomFloquet', 'TEBDEvolution', 'tensor_utils', 'observables']
