"""Makes the engine package importable when only ``<repo>/src`` is on sys.path (the reference's layout)."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
