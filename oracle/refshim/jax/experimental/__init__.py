"""``jax.experimental`` subset (see ../__init__.py)."""
from . import sparse as sparse  # noqa: F401
