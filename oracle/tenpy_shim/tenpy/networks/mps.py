"""tenpy.networks.mps facade: the MPS class is oracle.tebd_ref.MPS."""
from oracle.tebd_ref import MPS  # noqa: F401
