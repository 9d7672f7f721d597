"""``jax.lax`` control flow as eager Python (see __init__.py)."""
import numpy as _np

from .numpy import _wrap


def while_loop(cond_fun, body_fun, init_val):
    val = init_val
    while bool(_np.asarray(cond_fun(val))):
        val = body_fun(val)
    return val


def fori_loop(lower, upper, body_fun, init_val, **_k):
    val = init_val
    for i in range(int(lower), int(upper)):
        val = body_fun(i, val)
    return val


def scan(f, init, xs=None, length=None, **_k):
    from . import _tree_map

    carry = init
    ys = []
    n = length if xs is None else len(_first_leaf(xs))
    for i in range(int(n)):
        x = None if xs is None else _tree_map(lambda a: _wrap(_np.asarray(a)[i]), xs)
        carry, y = f(carry, x)
        ys.append(y)
    if not ys or ys[0] is None:
        return carry, None
    return carry, _tree_map(lambda *v: _wrap(_np.stack([_np.asarray(t) for t in v])), ys[0], *ys[1:]) if len(ys) > 1 \
        else _tree_map(lambda v: _wrap(_np.asarray(v)[None]), ys[0])


def _first_leaf(t):
    while isinstance(t, (tuple, list, dict)):
        t = next(iter(t.values())) if isinstance(t, dict) else t[0]
    return t


def cond(pred, true_fun, false_fun, *operands):
    return true_fun(*operands) if bool(_np.asarray(pred)) else false_fun(*operands)


def select(pred, a, b):
    return _wrap(_np.where(pred, a, b))


def dynamic_slice_in_dim(operand, start_index, slice_size, axis=0):
    a = _np.asarray(operand)
    n = a.shape[axis]
    s = int(_np.clip(int(_np.asarray(start_index)), 0, max(n - slice_size, 0)))  # XLA clamps
    idx = [slice(None)] * a.ndim
    idx[axis] = slice(s, s + slice_size)
    return _wrap(a[tuple(idx)])


def dynamic_update_slice_in_dim(operand, update, start_index, axis=0):
    a = _np.array(operand, copy=True)
    u = _np.asarray(update)
    n = a.shape[axis]
    s = int(_np.clip(int(_np.asarray(start_index)), 0, max(n - u.shape[axis], 0)))
    idx = [slice(None)] * a.ndim
    idx[axis] = slice(s, s + u.shape[axis])
    a[tuple(idx)] = u
    return _wrap(a)


def stop_gradient(x):
    return x
