"""tenpy.networks.site facade: SpinHalfSite with the basis-order rule of TeNPy >= 1.0
(sort_charge=True reorders the parity-conserving basis so that 'up' -> internal index 1;
SURVEY A.1.3).  Set TC_ORACLE_UP_INDEX=0 for the pre-1.0 ordering."""
import os

import numpy as np


class _Leg:
    def __init__(self, conserve, conj=False):
        self.conserve = conserve
        self.qconj = -1 if conj else 1

    def conj(self):
        return _Leg(self.conserve, self.qconj == 1)


class SpinHalfSite:
    dim = 2

    def __init__(self, conserve='Sz', sort_charge=None):
        self.conserve = conserve
        default_up = 1 if conserve == 'parity' else 0
        self._up = int(os.environ.get('TC_ORACLE_UP_INDEX', default_up)) if conserve == 'parity' else 0
        self.leg = _Leg(conserve)
        self.state_labels = {'up': self._up, 'down': 1 - self._up}

    def state_index(self, label):
        if isinstance(label, str):
            return self.state_labels[str(label)]
        return int(label)

    def get_op(self, name):
        sz = np.zeros((2, 2))
        sz[self._up, self._up] = 0.5
        sz[1 - self._up, 1 - self._up] = -0.5
        ops = {'Sz': sz, 'Sigmaz': 2 * sz, 'Id': np.eye(2)}
        return ops[name]
