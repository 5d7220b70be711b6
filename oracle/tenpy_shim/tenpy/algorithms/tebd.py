"""import-only placeholder: TEBDEvolution.evolve (tebd_evolution.py:51-108) is a dead
path in the reference; constructing the engine raises."""


class TEBDEngine:  # pragma: no cover
    def __init__(self, *a, **k):
        raise NotImplementedError("TEBDEngine is not restated in the shim (dead path in the reference)")
