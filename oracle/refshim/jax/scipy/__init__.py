"""``jax.scipy`` subset (see ../__init__.py)."""
import types

import numpy as _np

from ..numpy import _wrap


def _solve(a, b, **_k):
    return _wrap(_np.linalg.solve(_np.asarray(a), _np.asarray(b)))


linalg = types.SimpleNamespace(solve=_solve)
