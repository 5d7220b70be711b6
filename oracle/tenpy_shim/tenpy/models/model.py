"""import-only placeholders (kicked_ising.py:19, tebd_evolution.py:11)."""


class CouplingModel:  # pragma: no cover
    pass


class NearestNeighborModel:  # pragma: no cover
    pass
