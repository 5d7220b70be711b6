"""Same module path as the reference's ``src/dynamics/tebd_evolution.py``; implementation in the engine package."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
import _bootstrap  # noqa: F401,E402
from time_crystal_tensor_network_b200.dynamics.tebd_evolution import *  # noqa: F401,F403,E402
from time_crystal_tensor_network_b200.dynamics import tebd_evolution as _impl  # noqa: E402

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith('__')})
