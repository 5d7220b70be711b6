"""tenpy.networks.mpo facade (import-only; src/core/tensor_utils.py:7 never uses it)."""


class MPO:  # pragma: no cover - placeholder
    pass
