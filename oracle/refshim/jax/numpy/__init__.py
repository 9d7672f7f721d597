"""``jax.numpy`` on NumPy (see ../__init__.py).  Every array-valued result is a ``JArr``
(``numpy.ndarray`` subclass) so that ``x.at[idx].set(v)`` works."""
from __future__ import annotations

import numpy as _np


class _AtIndexer:
    def __init__(self, arr, idx):
        self.arr, self.idx = arr, idx

    def _idx(self):
        i = self.idx
        if isinstance(i, tuple):
            return tuple(_np.asarray(k) if isinstance(k, _np.ndarray) else k for k in i)
        return _np.asarray(i) if isinstance(i, _np.ndarray) else i

    def set(self, v, **_k):
        out = _np.array(self.arr, copy=True)
        out[self._idx()] = v
        return _wrap(out)

    def add(self, v, **_k):
        out = _np.array(self.arr, copy=True)
        _np.add.at(out, self._idx(), v)
        return _wrap(out)

    def multiply(self, v, **_k):
        out = _np.array(self.arr, copy=True)
        _np.multiply.at(out, self._idx(), v)
        return _wrap(out)

    def max(self, v, **_k):
        out = _np.array(self.arr, copy=True)
        _np.maximum.at(out, self._idx(), v)
        return _wrap(out)

    def min(self, v, **_k):
        out = _np.array(self.arr, copy=True)
        _np.minimum.at(out, self._idx(), v)
        return _wrap(out)

    def get(self, **_k):
        return _wrap(_np.asarray(self.arr)[self._idx()])


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        return _AtIndexer(self.arr, idx)


class JArr(_np.ndarray):
    @property
    def at(self):
        return _At(self)

    def block_until_ready(self):
        return self

    def __hash__(self):
        return id(self)

    def __getitem__(self, idx):
        # jnp indexing never raises for an out-of-range integer: it clamps to the last element
        # (e.g. the generated eval code reads limit_state_in[1] from a 1-element dummy when limiting is off)
        if isinstance(idx, (int, _np.integer)) and self.ndim >= 1 and self.shape[0] > 0:
            n = self.shape[0]
            if idx >= n:
                idx = n - 1
            elif idx < -n:
                idx = 0
        return _np.ndarray.__getitem__(self, idx)

    def __iter__(self):
        # (a clamping __getitem__ never raises IndexError, so the sequence-protocol fallback would not stop)
        if self.ndim == 0:
            raise TypeError("iteration over a 0-d array")
        return (_np.ndarray.__getitem__(self, i) for i in range(self.shape[0]))

    def __bool__(self):
        return bool(_np.asarray(self).all()) if self.size == 1 else _np.ndarray.__bool__(self)

    def __index__(self):
        return int(_np.asarray(self))

    def astype(self, dtype, *a, **k):
        return _wrap(_np.asarray(self).astype(dtype, *a, **k))

    def item(self, *a):
        return _np.asarray(self).item(*a)


def _wrap(x):
    if isinstance(x, JArr):
        return x
    if isinstance(x, _np.ndarray):
        return x.view(JArr)
    if isinstance(x, _np.generic):
        return _np.asarray(x).view(JArr)
    if isinstance(x, tuple):
        return tuple(_wrap(v) for v in x)
    if isinstance(x, list):
        return [_wrap(v) for v in x]
    return x


def _unwrap_dtype(kw):
    return kw


ndarray = JArr
float64, float32, int32, int64, bool_, uint8, uint32, int8 = (_np.float64, _np.float32, _np.int32, _np.int64,
                                                                _np.bool_, _np.uint8, _np.uint32, _np.int8)
pi, inf, nan, e, newaxis = _np.pi, _np.inf, _np.nan, _np.e, None
integer, floating, number = _np.integer, _np.floating, _np.number


def array(x, dtype=None, **_k):
    if isinstance(x, (list, tuple)):
        x = [(_np.asarray(v) if isinstance(v, _np.ndarray) else v) for v in x]
    a = _np.array(x, dtype=dtype)
    if dtype is None and a.dtype == _np.float32:
        a = a.astype(_np.float64)
    if dtype is None and a.dtype == _np.int64 and not isinstance(x, _np.ndarray):
        pass
    return _wrap(a)


def asarray(x, dtype=None, **_k):
    return _wrap(_np.asarray(x, dtype=dtype))


def zeros(shape, dtype=None, **_k):
    return _wrap(_np.zeros(shape, dtype=dtype or _np.float64))


def ones(shape, dtype=None, **_k):
    return _wrap(_np.ones(shape, dtype=dtype or _np.float64))


def full(shape, fill_value, dtype=None, **_k):
    return _wrap(_np.full(shape, fill_value, dtype=dtype))


def empty(shape, dtype=None, **_k):
    return zeros(shape, dtype)


def where(c, a=None, b=None):
    if a is None:
        return _wrap(_np.where(c))
    return _wrap(_np.where(c, a, b))


def interp(x, xp, fp, *a, **k):
    return _wrap(_np.asarray(_np.interp(x, xp, fp, *a, **k)))


def searchsorted(a, v, side="left", **_k):
    return _wrap(_np.asarray(_np.searchsorted(a, v, side=side)))


def _make(name):
    f = getattr(_np, name)

    def g(*a, **k):
        k.pop("precision", None)
        with _np.errstate(all="ignore"):
            return _wrap(f(*a, **k))

    g.__name__ = name
    return g


def __getattr__(name):
    if hasattr(_np, name):
        obj = getattr(_np, name)
        if callable(obj) and not isinstance(obj, type):
            fn = _make(name)
            globals()[name] = fn
            return fn
        return obj
    raise AttributeError(f"jax.numpy shim has no attribute {name!r}")


class _Linalg:
    def __getattr__(self, name):
        f = getattr(_np.linalg, name)
        return lambda *a, **k: _wrap(f(*a, **k))


linalg = _Linalg()
