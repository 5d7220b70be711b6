"""Drop-in for the reference's ``src/models`` package: same import path, B200 engine underneath."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
import _bootstrap  # noqa: F401,E402
from time_crystal_tensor_network_b200.models import *  # noqa: F401,F403,E402
from time_crystal_tensor_network_b200.models import __all__  # noqa: F401,E402
