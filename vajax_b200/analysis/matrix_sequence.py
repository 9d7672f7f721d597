"""Matrix-sequence wire format beside the hot path, and a replay harness for the sparse LU.

The reference captures the linear systems of a transient by wrapping ``_make_nr_solver_common``'s
``linear_solve_fn`` and writes them as ``jacobian_NNNN.mtx`` / ``rhs_NNNN.mtx`` (MatrixMarket
"coordinate real general", 1-based, ``%.17E``, exact zeros skipped; the solve is ``J @ delta = -f``):
``scripts/capture_benchmark_matrices.py:36-99`` (writers), ``:127-145`` (capture), ``:225-308`` (directory
layout).  Its single-matrix variant is ``jacobian.npz = {indptr, indices, data, rhs, shape}``
(``scripts/extract_c6288_jacobian.py:170-190``).  This module reads and writes both, so that a sequence
captured on any machine that can run the reference (real c6288 / PSP103 Jacobians) can be replayed through
seam S1 of the library (``vjx_lu_analyze`` once, ``vjx_lu_solve`` per system) without VAJAX installed.

Host-side file handling is NumPy; the solves run on the GPU through the C-ABI (no CPU fallback).
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import re
import time
from typing import Dict, List, Sequence, Tuple

import numpy as np

from .. import lib as vlib

_HEADER = "%%MatrixMarket matrix coordinate real general\n"


# ---- writers (capture_benchmark_matrices.py:36-99) --------------------------------------------------
def write_mm_csr(path: str, data: np.ndarray, indices: np.ndarray, indptr: np.ndarray, shape: Tuple[int, int],
                 comment: str = "") -> int:
    """CSR -> MatrixMarket coordinate; exact zeros in ``data`` are not written (the reference's rule, so
    the entry count of a file can be smaller than the number of CSR slots).  Returns the count written."""
    data = np.asarray(data, dtype=np.float64)
    indptr = np.asarray(indptr)
    indices = np.asarray(indices)
    rows = np.repeat(np.arange(shape[0]), np.diff(indptr))
    keep = data != 0.0
    with open(path, "w") as f:
        f.write(_HEADER)
        for line in comment.splitlines():
            f.write(f"%{line}\n")
        f.write(f"{shape[0]} {shape[1]} {int(keep.sum())}\n")
        for r, c, v in zip(rows[keep], indices[keep], data[keep]):
            f.write(f"{r + 1} {c + 1} {v:.17E}\n")
    return int(keep.sum())


def write_mm_vector(path: str, vec: np.ndarray, comment: str = "") -> int:
    """n x 1 coordinate matrix holding the non-zeros of ``vec``."""
    vec = np.asarray(vec, dtype=np.float64).ravel()
    nz = np.nonzero(vec)[0]
    with open(path, "w") as f:
        f.write(_HEADER)
        for line in comment.splitlines():
            f.write(f"%{line}\n")
        f.write(f"{len(vec)} 1 {len(nz)}\n")
        for i in nz:
            f.write(f"{i + 1} 1 {vec[i]:.17E}\n")
    return len(nz)


def read_mm(path: str) -> Tuple[Tuple[int, int], np.ndarray, np.ndarray, np.ndarray]:
    """MatrixMarket coordinate real general -> (shape, rows, cols, values), 0-based, file order."""
    with open(path) as f:
        head = f.readline()
        if not head.lower().startswith("%%matrixmarket matrix coordinate real general"):
            raise ValueError(f"{path}: not a 'coordinate real general' MatrixMarket file: {head.strip()!r}")
        line = f.readline()
        while line.startswith("%"):
            line = f.readline()
        nr, nc, nnz = (int(t) for t in line.split())
        body = np.loadtxt(f, dtype=np.float64, ndmin=2) if nnz else np.zeros((0, 3))
    if body.shape[0] != nnz:
        raise ValueError(f"{path}: header announces {nnz} entries, file holds {body.shape[0]}")
    r = body[:, 0].astype(np.int64) - 1
    c = body[:, 1].astype(np.int64) - 1
    if nnz and (r.min() < 0 or r.max() >= nr or c.min() < 0 or c.max() >= nc):
        raise ValueError(f"{path}: index outside the {nr} x {nc} shape")
    return (nr, nc), r, c, body[:, 2].copy()


def read_mm_vector(path: str) -> np.ndarray:
    (n, ncol), r, c, v = read_mm(path)
    if ncol != 1:
        raise ValueError(f"{path}: expected an n x 1 vector, got {n} x {ncol}")
    out = np.zeros(n)
    out[r] = v
    return out


def write_sequence(out_dir: str, indptr: np.ndarray, indices: np.ndarray, systems: Sequence[Tuple[np.ndarray, np.ndarray]],
                   name: str = "circuit") -> List[int]:
    """Directory layout of ``write_systems`` (capture_benchmark_matrices.py:225-308) for a sparse capture."""
    os.makedirs(out_dir, exist_ok=True)
    n = len(indptr) - 1
    counts = []
    for k, (data, rhs) in enumerate(systems):
        nnz = int(np.count_nonzero(data))
        counts.append(write_mm_csr(os.path.join(out_dir, f"jacobian_{k:04d}.mtx"), data, indices, indptr, (n, n),
                                   f" {name} benchmark, NR linear system #{k}\n matrix size: {n}x{n}, nnz={nnz} "
                                   f"(CSR with {len(data)} slots)"))
        write_mm_vector(os.path.join(out_dir, f"rhs_{k:04d}.mtx"), rhs, f" {name} benchmark, RHS vector #{k}")
    return counts


def write_jacobian_npz(path: str, indptr, indices, data, rhs) -> None:
    """extract_c6288_jacobian.py:170-190."""
    n = len(indptr) - 1
    np.savez_compressed(path, indptr=np.asarray(indptr, np.int32), indices=np.asarray(indices, np.int32),
                        data=np.asarray(data, np.float64), rhs=np.asarray(rhs, np.float64), shape=np.array([n, n], np.int32))


# ---- readers -----------------------------------------------------------------------------------------
def load_sequence(directory: str) -> Dict[str, object]:
    """Read every ``jacobian_NNNN.mtx`` / ``rhs_NNNN.mtx`` pair of a capture directory and put the
    matrices on ONE CSR pattern: the union of the entries of all files (a file lists only the non-zeros
    of its system, so a slot that happens to be 0.0 in one NR iteration is missing there), plus the
    diagonal, which the solver requires structurally.  Returns ``{n, indptr, indices, data[k], rhs[k], ids}``."""
    files = sorted(glob.glob(os.path.join(directory, "jacobian_*.mtx")))
    if not files:
        raise FileNotFoundError(f"no jacobian_NNNN.mtx under {directory}")
    ids, mats, rhss = [], [], []
    n = None
    for jf in files:
        m = re.search(r"jacobian_(\d+)\.mtx$", jf)
        rf = os.path.join(directory, f"rhs_{m.group(1)}.mtx")
        (nr, nc), r, c, v = read_mm(jf)
        if nr != nc:
            raise ValueError(f"{jf}: Jacobian is not square")
        if n is None:
            n = nr
        elif nr != n:
            raise ValueError(f"{jf}: size {nr} differs from the first system's {n}")
        ids.append(int(m.group(1)))
        mats.append((r * n + c, v))
        rhss.append(read_mm_vector(rf) if os.path.exists(rf) else np.zeros(n))
        if len(rhss[-1]) != n:
            raise ValueError(f"{rf}: length {len(rhss[-1])} != {n}")
    keys = np.unique(np.concatenate([k for k, _ in mats] + [np.arange(n, dtype=np.int64) * (n + 1)]))
    rows = (keys // n).astype(np.int64)
    indptr = np.zeros(n + 1, dtype=np.int32)
    np.add.at(indptr, rows + 1, 1)
    indptr = np.cumsum(indptr).astype(np.int32)
    indices = (keys % n).astype(np.int32)
    data = []
    for k, v in mats:
        d = np.zeros(len(keys))
        pos = np.searchsorted(keys, k)
        np.add.at(d, pos, v)   # duplicate coordinates sum (MatrixMarket semantics)
        data.append(d)
    return dict(n=n, indptr=indptr, indices=indices, data=data, rhs=rhss, ids=ids)


def load_jacobian_npz(path: str) -> Dict[str, object]:
    z = np.load(path)
    n = int(z["shape"][0])
    return dict(n=n, indptr=z["indptr"].astype(np.int32), indices=z["indices"].astype(np.int32),
                data=[z["data"].astype(np.float64)], rhs=[z["rhs"].astype(np.float64)], ids=[0])


# ---- seam S1 on its own: a pattern-only handle ----------------------------------------------------------
class SparseLUSolver:
    """``linear_solve_fn(csr_data, f)`` of the reference (``solver_factories.py:985-1004``; UMFPACK FFI
    ``umfpack_ffi.cpp:189-326``) on a fixed CSR pattern, without a circuit: symbolic analysis + pivot order
    from the first matrix solved (or ``analyze``), numeric refactorisation + solve per call, all on the GPU."""

    def __init__(self, indptr: np.ndarray, indices: np.ndarray, device: int = 0, pivot_tol: float = 1e-3):
        import torch

        if not torch.cuda.is_available():
            raise vlib.VjxError("no CUDA device: vajax_b200 has no CPU fallback")
        self._torch = torch
        self.device = device
        self.indptr = np.ascontiguousarray(indptr, dtype=np.int32)
        self.indices = np.ascontiguousarray(indices, dtype=np.int32)
        self.n = len(self.indptr) - 1
        rows = np.repeat(np.arange(self.n), np.diff(self.indptr))
        has_diag = np.zeros(self.n, dtype=bool)
        has_diag[rows[rows == self.indices]] = True
        if not has_diag.all():
            raise ValueError("every row needs a structural diagonal entry")
        cd = vlib.CircuitDesc()
        cd.n_total = self.n + 1          # ground + n unknowns, no voltage sources: n_aug = n
        cd.n_external = self.n + 1
        cd.n_vsources = cd.n_isources = cd.n_models = cd.n_pwl = cd.n_noi = cd.n_internal = 0
        cd.nse = int(self.indptr[-1])
        cd.csr_indptr = self.indptr.ctypes.data_as(C.c_void_p)
        cd.csr_indices = self.indices.ctypes.data_as(C.c_void_p)
        o = vlib.Options()
        o.nr_damping, o.reltol, o.abstol, o.vntol, o.gmin, o.max_step = 1.0, 1e-3, 1e-12, 1e-6, 0.0, 1e30
        o.tran_itl, o.max_order, o.integration_method, o.pivot_tol, o.use_device_graph = 20, 2, 1, pivot_tol, 0
        cd.opts = o
        self.handle = C.c_void_p(None)
        vlib.check(vlib.load().vjx_create(C.byref(cd), device, C.byref(self.handle)))

    def _dev(self, a):
        return self._torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64), device=f"cuda:{self.device}")

    def analyze(self, csr_data: np.ndarray) -> dict:
        d = self._dev(csr_data)
        vlib.check(vlib.load().vjx_lu_analyze(self.handle, C.c_void_p(d.data_ptr()), None))
        return self.stats()

    def solve(self, csr_data: np.ndarray, b: np.ndarray) -> np.ndarray:
        d, bb = self._dev(csr_data), self._dev(b)
        x = self._torch.zeros(self.n, dtype=self._torch.float64, device=d.device)
        vlib.check(vlib.load().vjx_lu_solve(self.handle, C.c_void_p(d.data_ptr()), C.c_void_p(bb.data_ptr()),
                                            C.c_void_p(x.data_ptr()), None))
        self._torch.cuda.current_stream(self.device).synchronize()
        return x.cpu().numpy()

    def stats(self) -> dict:
        out = np.zeros(10)
        vlib.check(vlib.load().vjx_lu_stats(self.handle, out.ctypes.data_as(C.c_void_p)))
        return dict(nnz_lu=int(out[0]), flops=float(out[1]), levels_l=int(out[2]), levels_u=int(out[3]), n=int(out[5]),
                    n_singletons=int(out[6]), n_core=int(out[7]))

    def close(self):
        if self.handle:
            vlib.load().vjx_destroy(self.handle)
            self.handle = C.c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def replay(seq: Dict[str, object], device: int = 0, reanalyze_every: int = 0, pivot_tol: float = 1e-3) -> List[dict]:
    """Solve ``J_k delta = -f_k`` for every system of a loaded sequence with the pivot order fixed at the
    first one (north_star item 3) and report, per system, the scaled residual
    ``max|J delta + f| / max(1, max|J| |delta| + |f|)`` and the wall time of the call."""
    import scipy.sparse as sp

    s = SparseLUSolver(seq["indptr"], seq["indices"], device=device, pivot_tol=pivot_tol)
    out = []
    try:
        for k, (d, f) in enumerate(zip(seq["data"], seq["rhs"])):
            if k == 0 or (reanalyze_every and k % reanalyze_every == 0):
                s.analyze(d)
            t0 = time.perf_counter()
            x = s.solve(d, -np.asarray(f))
            dt = time.perf_counter() - t0
            A = sp.csr_matrix((d, seq["indices"], seq["indptr"]), shape=(seq["n"], seq["n"]))
            r = A @ x + f
            scale = max(1.0, float(np.max(abs(A) @ np.abs(x) + np.abs(f))))
            out.append(dict(id=seq["ids"][k], residual=float(np.max(np.abs(r))) / scale, wall_ms=dt * 1e3, x=x))
    finally:
        s.close()
    return out
