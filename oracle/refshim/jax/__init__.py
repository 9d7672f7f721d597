"""NumPy-backed stand-in for the small part of the ``jax`` API that the reference's transient hot
path uses.  TEST INFRASTRUCTURE ONLY (oracle/): it exists so that the *reference's own Python
modules* (``/root/reference/vajax/analysis/*.py``, ``/root/reference/openvaf_jax``) can be imported
and executed in the build container, where JAX cannot be installed, to generate the
golden vectors under ``tests/golden/`` (``tests/golden/make_golden.py``).  Nothing in the product,
the tests or bench.py imports this package at run time.

Semantics kept: fp64 everywhere (``jax_enable_x64``), functional ``x.at[idx].set/add/max`` with
NumPy's in-order accumulation for duplicate indices (XLA CPU scatter applies updates in order as
well), ``lax.while_loop`` / ``fori_loop`` / ``scan`` / ``cond`` as eager Python loops, ``vmap`` as an
eager loop over the mapped axis, ``jit`` as identity.
"""
from __future__ import annotations

import types

import numpy as _np

from . import numpy as numpy  # noqa: F401  (jax.numpy)
from .numpy import JArr, _wrap

Array = JArr
__version__ = "0.0-numpy-shim"


def jit(fun=None, **_kw):
    if fun is None:
        return lambda f: f
    return fun


def _tree_map(f, tree, *rest):
    if isinstance(tree, tuple) and hasattr(tree, "_fields"):
        return type(tree)(*[_tree_map(f, t, *[r[i] for r in rest]) for i, t in enumerate(tree)])
    if isinstance(tree, (list, tuple)):
        return type(tree)(_tree_map(f, t, *[r[i] for r in rest]) for i, t in enumerate(tree))
    if isinstance(tree, dict):
        return {k: _tree_map(f, v, *[r[k] for r in rest]) for k, v in tree.items()}
    if tree is None:
        return None
    return f(tree, *rest)


def vmap(fun, in_axes=0, out_axes=0):
    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = None
        for a, ax in zip(args, axes):
            if ax is not None:
                n = _np.shape(a)[ax]
                break
        outs = []
        for i in range(n):
            call = [a if ax is None else _wrap(_np.take(_np.asarray(a), i, axis=ax)) for a, ax in zip(args, axes)]
            outs.append(fun(*call))
        if n == 0:
            raise ValueError("vmap over an empty axis is not supported by the shim")
        return _tree_map(lambda *xs: _wrap(_np.stack([_np.asarray(x) for x in xs], axis=0)), outs[0], *outs[1:]) \
            if len(outs) > 1 else _tree_map(lambda x: _wrap(_np.asarray(x)[None]), outs[0])
    return mapped


class _Config:
    jax_enable_x64 = True

    def update(self, *_a, **_k):
        return None

    def __getattr__(self, name):
        return None


config = _Config()


class Device:
    platform = "cpu"
    device_kind = "cpu"
    id = 0

    def __repr__(self):
        return "CpuDevice(id=0)"


def devices(*_a, **_k):
    return [Device()]


def local_devices(*_a, **_k):
    return [Device()]


def default_backend():
    return "cpu"


def device_put(x, *_a, **_k):
    return x


def device_get(x):
    return x


def block_until_ready(x):
    return x


def clear_caches():
    return None


def named_scope(*_a, **_k):
    import contextlib

    return contextlib.nullcontext()


def default_device(*_a, **_k):
    import contextlib

    return contextlib.nullcontext()


def ensure_compile_time_eval():
    import contextlib

    return contextlib.nullcontext()


debug = types.SimpleNamespace(callback=lambda *a, **k: None, print=lambda *a, **k: None)
tree_util = types.SimpleNamespace(tree_map=_tree_map)
tree = types.SimpleNamespace(map=_tree_map)


def _segment(op):
    def f(data, segment_ids, num_segments=None, **_k):
        data = _np.asarray(data)
        ids = _np.asarray(segment_ids)
        n = int(num_segments if num_segments is not None else (ids.max() + 1 if ids.size else 0))
        if op == "sum":
            out = _np.zeros((n,) + data.shape[1:], dtype=data.dtype)
            ok = (ids >= 0) & (ids < n)
            _np.add.at(out, ids[ok], data[ok])
        else:
            out = _np.full((n,) + data.shape[1:], -_np.inf, dtype=data.dtype)
            ok = (ids >= 0) & (ids < n)
            _np.maximum.at(out, ids[ok], data[ok])
        return _wrap(out)
    return f


ops = types.SimpleNamespace(segment_sum=_segment("sum"), segment_max=_segment("max"))

from . import lax as lax  # noqa: E402,F401
from . import scipy as scipy  # noqa: E402,F401
from . import experimental as experimental  # noqa: E402,F401
from . import profiler as profiler  # noqa: E402,F401
