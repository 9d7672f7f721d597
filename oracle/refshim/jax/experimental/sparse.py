"""Minimal ``jax.experimental.sparse``: just enough BCOO / BCSR for the reference's CSR pattern
discovery (full_mna.py:423-459 reads ``J.indices`` of a trial BCOO Jacobian)."""
import numpy as _np

from ..numpy import _wrap


class BCOO:
    def __init__(self, args, shape=None, **_k):
        data, indices = args
        self.data = _wrap(_np.asarray(data))
        self.indices = _wrap(_np.asarray(indices))
        self.shape = tuple(shape)

    def todense(self):
        out = _np.zeros(self.shape, dtype=self.data.dtype)
        _np.add.at(out, (self.indices[:, 0], self.indices[:, 1]), self.data)
        return _wrap(out)

    def sum_duplicates(self, nse=None, **_k):
        lin = _np.asarray(self.indices[:, 0]).astype(_np.int64) * self.shape[1] + _np.asarray(self.indices[:, 1])
        u, inv = _np.unique(lin, return_inverse=True)
        d = _np.zeros(len(u), dtype=self.data.dtype)
        _np.add.at(d, inv, _np.asarray(self.data))
        idx = _np.stack([u // self.shape[1], u % self.shape[1]], axis=1).astype(_np.int32)
        return BCOO((d, idx), shape=self.shape)


class BCSR:
    def __init__(self, args, shape=None, **_k):
        data, indices, indptr = args
        self.data, self.indices, self.indptr = _wrap(_np.asarray(data)), _wrap(_np.asarray(indices)), _wrap(_np.asarray(indptr))
        self.shape = tuple(shape)

    @classmethod
    def from_bcoo(cls, m):
        m = m.sum_duplicates()
        rows = _np.asarray(m.indices[:, 0])
        indptr = _np.zeros(m.shape[0] + 1, dtype=_np.int32)
        indptr[1:] = _np.cumsum(_np.bincount(rows, minlength=m.shape[0]))
        return cls((m.data, m.indices[:, 1], indptr), shape=m.shape)

    def __matmul__(self, x):
        import scipy.sparse as sp

        A = sp.csr_matrix((_np.asarray(self.data), _np.asarray(self.indices), _np.asarray(self.indptr)), shape=self.shape)
        return _wrap(A @ _np.asarray(x))
