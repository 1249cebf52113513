from .registry import ALGORITHMS, Register  # noqa: F401
