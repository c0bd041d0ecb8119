"""Stand-in for the ``flax`` package (test infrastructure, see ../README.md)."""
from . import linen  # noqa: F401
