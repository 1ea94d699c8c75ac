"""Stand-in for `flax` (see ../README.md)."""
from . import linen  # noqa: F401
