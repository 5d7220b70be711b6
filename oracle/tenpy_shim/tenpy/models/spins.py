"""import-only placeholder (tensor_utils.py:8)."""


class SpinChain:  # pragma: no cover
    pass
