"""tenpy.linalg.np_conserved facade: a dense array wrapper with labels."""
import numpy as np


class Array:
    def __init__(self, data, labels=None):
        self._data = np.asarray(data)
        self.labels = list(labels) if labels is not None else None
        self.rank = self._data.ndim
        self.dtype = self._data.dtype

    @classmethod
    def from_ndarray(cls, data_flat, legcharges=None, dtype=None, qtotal=None, cutoff=None,
                     labels=None, raise_wrong_sector=True, warn_wrong_sector=True):
        return cls(np.array(data_flat, dtype=dtype), labels)

    def to_ndarray(self):
        return self._data

    @property
    def ndim(self):
        return self._data.ndim
