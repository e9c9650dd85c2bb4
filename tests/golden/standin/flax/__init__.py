"""flax stand-in: just enough of linen/struct for internal/models.py to run its real code on
numpy (see ../README.md)."""
from . import core, linen, struct, training  # noqa: F401
