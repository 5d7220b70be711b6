"""Makes the engine package importable when only ``<repo>/src`` is on sys.path (the reference's layout); symbolic links
to ``src`` are resolved, so a checkout of the reference can link its ``src`` to this one."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.realpath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
