"""swim_b200 — B200-native bulk simulator of the SWIM membership protocol behind the
Core/Types API surface of jpfuentes2/swim. The compute path is CUDA (sm_100a) only."""
from . import _abi as abi  # noqa: F401

__all__ = ["abi"]
