"""Drop-in for the reference's ``src/dynamics`` package."""
from .tebd_evolution import TEBDEvolution, CustomFloquet

__all__ = ['TEBDEvolution', 'CustomFloquet']
