"""Drop-in for the reference's ``src/models`` package."""
from .kicked_ising import KickedIsingModel

__all__ = ['KickedIsingModel']
