"""import-only placeholder (kicked_ising.py:18)."""


class Chain:  # pragma: no cover
    pass
