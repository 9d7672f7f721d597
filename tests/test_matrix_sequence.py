"""Matrix-sequence wire format (reference scripts/capture_benchmark_matrices.py:36-99,225-308,
scripts/extract_c6288_jacobian.py:170-190): exact 17-digit round trip on the CPU, replay through seam S1
of the library on the GPU."""

import numpy as np
import pytest

from vajax_b200.analysis import matrix_sequence as ms


def _systems(n=40, k=4, seed=5):
    import scipy.sparse as sp

    rng = np.random.default_rng(seed)
    A = sp.random(n, n, density=0.08, random_state=rng.integers(1 << 30), format="csr") + sp.eye(n, format="csr") * 4.0
    A.sort_indices()
    indptr, indices = A.indptr.astype(np.int32), A.indices.astype(np.int32)
    out = []
    for j in range(k):
        d = A.data * rng.uniform(0.5, 1.5, size=A.nnz)
        if j == 1:
            d[rng.integers(0, A.nnz, size=3)] = 0.0   # exact zeros are skipped by the writer
        f = rng.normal(size=n)
        f[rng.integers(0, n, size=2)] = 0.0
        out.append((d, f))
    return indptr, indices, out


def test_round_trip_is_exact(tmp_path):
    indptr, indices, systems = _systems()
    counts = ms.write_sequence(str(tmp_path), indptr, indices, systems, name="unit")
    assert counts[1] == np.count_nonzero(systems[1][0]) < len(systems[1][0])
    head = open(tmp_path / "jacobian_0001.mtx").readline()
    assert head == "%%MatrixMarket matrix coordinate real general\n"   # the reference's header, verbatim
    seq = ms.load_sequence(str(tmp_path))
    n = len(indptr) - 1
    assert seq["n"] == n and seq["ids"] == [0, 1, 2, 3]
    import scipy.sparse as sp

    for (d, f), d2, f2 in zip(systems, seq["data"], seq["rhs"]):
        A = sp.csr_matrix((d, indices, indptr), shape=(n, n)).toarray()
        B = sp.csr_matrix((d2, seq["indices"], seq["indptr"]), shape=(n, n)).toarray()
        assert np.array_equal(A, B) and np.array_equal(f, f2)   # %.17E round-trips every fp64 bit
    # the union pattern holds the structural diagonal and is strictly sorted per row
    for r in range(n):
        c = seq["indices"][seq["indptr"][r]:seq["indptr"][r + 1]]
        assert r in c and np.all(np.diff(c) > 0)


def test_reference_style_file_is_read(tmp_path):
    """A file as scipy / the reference's writer produce it: comments, 1-based, duplicate entries sum."""
    p = tmp_path / "jacobian_0000.mtx"
    p.write_text("%%MatrixMarket matrix coordinate real general\n% c6288 benchmark, NR linear system #0\n"
                 "% matrix size: 3x3, nnz=5 (CSR with 6 slots)\n3 3 5\n1 1 2.00000000000000000E+00\n"
                 "2 2 3.00000000000000000E+00\n3 3 4.5E+00\n3 1 -1.25000000000000000E-01\n3 1 -1.25000000000000000E-01\n")
    (tmp_path / "rhs_0000.mtx").write_text("%%MatrixMarket matrix coordinate real general\n3 1 1\n2 1 7.0E+00\n")
    seq = ms.load_sequence(str(tmp_path))
    assert seq["n"] == 3 and list(seq["indptr"]) == [0, 1, 2, 4] and list(seq["indices"]) == [0, 1, 0, 2]
    assert np.array_equal(seq["data"][0], [2.0, 3.0, -0.25, 4.5]) and np.array_equal(seq["rhs"][0], [0.0, 7.0, 0.0])


def test_malformed_files_raise(tmp_path):
    p = tmp_path / "jacobian_0000.mtx"
    p.write_text("%%MatrixMarket matrix array real general\n2 2\n1\n2\n3\n4\n")
    with pytest.raises(ValueError):
        ms.load_sequence(str(tmp_path))
    p.write_text("%%MatrixMarket matrix coordinate real general\n2 2 2\n1 1 1.0\n")
    with pytest.raises(ValueError):
        ms.load_sequence(str(tmp_path))
    p.write_text("%%MatrixMarket matrix coordinate real general\n2 2 1\n3 1 1.0\n")
    with pytest.raises(ValueError):
        ms.load_sequence(str(tmp_path))
    with pytest.raises(FileNotFoundError):
        ms.load_sequence(str(tmp_path / "nothing"))


def test_npz_variant(tmp_path):
    indptr, indices, systems = _systems(k=1)
    ms.write_jacobian_npz(str(tmp_path / "jacobian.npz"), indptr, indices, *systems[0])
    seq = ms.load_jacobian_npz(str(tmp_path / "jacobian.npz"))
    assert np.array_equal(seq["data"][0], systems[0][0]) and np.array_equal(seq["indptr"], indptr)


@pytest.mark.gpu
def test_replay_random_sequence_on_gpu(tmp_path):
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl

    indptr, indices, systems = _systems(n=300, k=5, seed=11)
    ms.write_sequence(str(tmp_path), indptr, indices, systems)
    seq = ms.load_sequence(str(tmp_path))
    res = ms.replay(seq)
    assert [r["id"] for r in res] == [0, 1, 2, 3, 4]
    for r, d, f in zip(res, seq["data"], seq["rhs"]):
        assert r["residual"] < 1e-12
        A = sp.csc_matrix(sp.csr_matrix((d, seq["indices"], seq["indptr"]), shape=(300, 300)))
        ref = spl.splu(A).solve(-f)
        assert np.max(np.abs(r["x"] - ref)) <= 1e-9 * max(1.0, np.max(np.abs(ref)))


@pytest.mark.gpu
def test_replay_captured_nr_systems_of_a_circuit(tmp_path, bundles):
    """Capture the NR linear systems of a circuit (the engine's own build_system at successive NR states, NOI
    rows enforced as the solver sees them), write them in the reference's format, replay them through the
    pattern-only handle: the deltas equal the ones the in-engine LU produces."""
    from vajax_b200.analysis.engine import B200Engine
    from vajax_b200.benchmarks import circuits

    e = B200Engine(circuits.multiplier_circuit(3), bundles).prepare()
    nu = e.n_unknowns
    rng = np.random.default_rng(2)
    X = np.zeros(e.n_total + e.n_vs)
    X[1:e.n_total] = rng.uniform(0, 1.2, nu)
    X[e.noi_indices] = 0.0
    vs, isv = e.eval_sources(1e-10)
    rows = np.repeat(np.arange(e.n_aug), np.diff(e.csr_indptr))
    noi = np.zeros(e.n_aug, bool)
    noi[e.noi_indices - 1] = True
    systems, deltas = [], []
    for it in range(3):
        out = e.build_system(X, vs, isv, np.zeros(nu), 1e12, 1e-12, 0.0, -1e12, -1.0, None, 0.0, None, None, 1)
        csr, f = out[0].copy(), out[1].copy()
        csr[noi[rows] | ((rows < nu) & noi[e.csr_indices])] = 0.0
        csr[noi[rows] & (rows == e.csr_indices)] = 1.0
        f[noi] = 0.0
        d = e.lu_solve(csr, -f)
        systems.append((csr, f))
        deltas.append(d)
        X[1:e.n_total] += 0.5 * d[:nu]
        X[e.noi_indices] = 0.0
    ms.write_sequence(str(tmp_path), e.csr_indptr, e.csr_indices, systems, name="mul3")
    e.close()
    res = ms.replay(ms.load_sequence(str(tmp_path)))
    for r, d in zip(res, deltas):
        assert r["residual"] < 1e-10
        assert np.max(np.abs(r["x"] - d)) <= 1e-9 * max(1.0, np.max(np.abs(d)))
