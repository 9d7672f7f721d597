"""GPU parity tests: the CUDA path (through the C-ABI) against the NumPy oracle on the same inputs.

Tolerances (BASELINE.json north_star): stamped Jacobians 1e-12 relative, per-NR-solve voltages 1e-9
relative, waveforms 1e-6 RMS relative with the same accepted-step sequence.  Vector parity metric:
max_i |a_i - b_i| / max(|b_i|, floor).
"""
import numpy as np
import pytest

from golden_util import assert_vector_parity
from vajax_b200.benchmarks import circuits

pytestmark = pytest.mark.gpu


def relerr(a, b, floor):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)))


def engines(circ, bundles, limiting=False, graph=False):
    from oracle.engine import OracleEngine
    from vajax_b200.analysis.engine import B200Engine

    o = OracleEngine(circ, bundles, use_device_limiting=limiting)
    o.setup()
    g = B200Engine(circ, bundles, use_device_limiting=limiting, use_device_graph=graph).prepare()
    return o, g


from golden_util import CIRCUITS  # noqa: E402  (rc, clipper, ring, mul3, opcov, srcmix, BE / gear2 variants)


@pytest.mark.parametrize("name", list(CIRCUITS))
@pytest.mark.parametrize("limiting", [False, True])
def test_setup_and_build_system_parity(name, limiting, bundles):
    circ = CIRCUITS[name]()
    o, g = engines(circ, bundles, limiting)
    assert g.n_total == o.n_total and g.nse == o.nse and g.n_aug == o.n_aug
    assert np.array_equal(g.csr_indptr, o.csr_indptr) and np.array_equal(g.csr_indices, o.csr_indices)
    assert g.total_limit_states == o.total_limit_states
    rng = np.random.default_rng(7)
    X = np.zeros(o.n_total + o.n_vs)
    X[1:] = rng.uniform(-0.3, 1.2, X.size - 1)
    Qp, dq, q2 = (rng.normal(0, 1e-13, o.n_unknowns) for _ in range(3))
    vs, isv = o.eval_sources(2.5e-10)
    ls = rng.uniform(0, 0.7, o.total_limit_states)
    for (c0, c1, d1, c2, it) in [(0.0, 0.0, 0.0, 0.0, 0), (2e12, -2e12, -1.0, 0.0, 1), (1.5e9, -2e9, 0.0, 0.5e9, 3)]:
        ro = o.build_system(X, vs, isv, Qp, c0, 1e-12, 1e-9, c1, d1, dq, c2, q2, ls, it)
        rg = g.build_system(X, vs, isv, Qp, c0, 1e-12, 1e-9, c1, d1, dq, c2, q2, ls, it)
        for nm, a, b in zip(["csr", "f", "Q", "Ivs", "limit", "maxres"], rg, ro):
            assert_vector_parity(nm, a, b, (name, limiting, c0))
    g.close()


@pytest.mark.parametrize("name", list(CIRCUITS))
def test_lu_solve_against_superlu(name, bundles):
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla

    circ = CIRCUITS[name]()
    o, g = engines(circ, bundles)
    rng = np.random.default_rng(3)
    X = np.zeros(o.n_total + o.n_vs)
    X[1:o.n_total] = rng.uniform(0, 1.2, o.n_total - 1)
    if len(o.noi_indices):
        X[o.noi_indices] = 0
    vs, isv = o.eval_sources(1e-10)
    csr, f, *_ = o.build_system(X, vs, isv, np.zeros(o.n_unknowns), 1e12)
    if len(o.noi_res_idx):
        csr[o.noi_row_mask] = 0
        csr[o.noi_col_mask] = 0
        csr[o.noi_diag] = 1
    A = sp.csr_matrix((csr, o.csr_indices, o.csr_indptr), shape=(o.n_aug, o.n_aug))
    for k in range(3):
        b = rng.normal(size=o.n_aug)
        x = g.lu_solve(csr * (1.0 + 0.1 * k), b)
        xr = spla.splu((A * (1.0 + 0.1 * k)).tocsc()).solve(b)
        assert relerr(x, xr, 1e-9 * np.max(np.abs(xr))) < 1e-9
        r = A @ x * (1.0 + 0.1 * k) - b
        assert np.max(np.abs(r)) < 1e-9 * max(1.0, np.max(np.abs(b)))
    st = g.lu_stats()
    # the NOI-zero slots of A are structural zeros for the factorisation (csrc ensure_lu), so nnz(L+U) can be < nse
    assert st["n"] == o.n_aug and st["nnz_lu"] >= o.n_aug
    g.close()


@pytest.mark.parametrize("name", list(CIRCUITS))
@pytest.mark.parametrize("limiting", [False, True])
def test_nr_solve_parity(name, limiting, bundles):
    circ = CIRCUITS[name]()
    o, g = engines(circ, bundles, limiting)
    nu = o.n_unknowns
    X0 = np.zeros(o.n_total + o.n_vs)
    X0[:o.n_total] = o.init_mid_rail()
    vs, isv = o.eval_sources(0.0)
    ro = o.nr_solve(X0, vs, isv, np.zeros(nu), 0.0, 1e-12)
    rg = g.nr_solve(X0, vs, isv, np.zeros(nu), 0.0, 1e-12)
    assert rg[1] == ro[1] and rg[2] == ro[2], (rg[1], ro[1], rg[2], ro[2])
    c0 = 2.0 / 1e-11
    if ro[2]:
        assert relerr(rg[0], ro[0], 1e-6) < 1e-9
        assert relerr(rg[4], ro[4], max(1e-24, 1e-9 * np.max(np.abs(ro[4])))) < 1e-8
        Xs, Qs, lim = ro[0], ro[4], ro[7]
    else:
        # the mid-rail DC solve diverges in the reference semantics too (the multipliers run with
        # icmode=uic): iterates of a diverging NR are chaotic, so only flags are compared, and the
        # transient-style solve starts from the UIC state (full_mna.py:882-921)
        Xs = np.zeros_like(X0)
        Qs = o.build_system(Xs, vs, isv, np.zeros(nu), 0.0)[2]
        lim = None
    # a transient-style solve (trapezoidal coefficients, dt = 10 ps)
    r2o = o.nr_solve(Xs, vs, isv, Qs, c0, 1e-12, 0.0, -c0, -1.0, np.zeros(nu), 0.0, None, lim, None)
    r2g = g.nr_solve(Xs, vs, isv, Qs, c0, 1e-12, 0.0, -c0, -1.0, np.zeros(nu), 0.0, None, lim, None)
    assert r2g[1] == r2o[1] and r2g[2] == r2o[2]
    assert r2o[2], "transient-style solve must converge in the oracle"
    assert relerr(r2g[0], r2o[0], 1e-6) < 1e-9
    assert relerr(r2g[4], r2o[4], max(1e-24, 1e-9 * np.max(np.abs(r2o[4])))) < 1e-8
    g.close()


@pytest.mark.parametrize("name", ["rc", "clipper", "ring", "mul3", "opcov", "ringw"])
@pytest.mark.parametrize("limiting", [False, True])
def test_nr_solve_stream_ordered_equals_host_polled(name, limiting, bundles):
    """vjx_nr_solve_async (the XLA-FFI form: NR loop as a graph WHILE node, scalars written by the device, no
    host read) returns bit-for-bit what the host-polled vjx_nr_solve returns, call after call on one handle."""
    circ = CIRCUITS[name]()
    o, g = engines(circ, bundles, limiting)
    nu = o.n_unknowns
    X0 = np.zeros(o.n_total + o.n_vs)
    X0[:o.n_total] = o.init_mid_rail()
    vs, isv = o.eval_sources(0.0)
    c0 = 2.0 / 1e-11
    rng = np.random.default_rng(3)
    for k, args in enumerate([(np.zeros(nu), 0.0, 1e-12), (rng.normal(0, 1e-15, nu), c0, 1e-12, 0.0, -c0, -1.0, np.zeros(nu)),
                              (np.zeros(nu), 0.0, 1e-12)]):
        Xs = X0 if k != 1 else np.zeros_like(X0)
        a = g.nr_solve(Xs, vs, isv, *args)
        b = g.nr_solve(Xs, vs, isv, *args, stream_ordered=True)
        assert a[1] == b[1] and a[2] == b[2], (k, a[1], b[1], a[2], b[2])
        assert (a[3] == b[3]) or (np.isnan(a[3]) and np.isnan(b[3])), (k, a[3], b[3])
        for i in (0, 4, 5, 6, 7, 8):
            assert np.array_equal(a[i], b[i], equal_nan=True), (k, i)
    # max_iterations = 0: only the final build
    a = g.nr_solve(X0, vs, isv, np.zeros(nu), 0.0, max_iterations=0)
    b = g.nr_solve(X0, vs, isv, np.zeros(nu), 0.0, max_iterations=0, stream_ordered=True)
    assert a[1] == b[1] == 0 and np.array_equal(a[4], b[4]) and np.array_equal(a[0], b[0])
    g.close()


@pytest.mark.parametrize("graph", [False, True])
@pytest.mark.parametrize("name", list(CIRCUITS))
def test_transient_parity(name, graph, bundles):
    circ = CIRCUITS[name]()
    o, g = engines(circ, bundles, False, graph)
    rec = []
    to, Vo, so = o.run_transient(record=rec)
    tg, Vg, sg = g.run_transient()
    assert sg["error_flags"] == 0
    assert sg["n_steps"] == so["n_steps"], (sg["n_steps"], so["n_steps"])
    assert sg["rejected_steps"] == so["rejected_steps"]
    assert sg["total_nr_iterations"] == so["total_nr_iterations"]
    n = so["n_steps"]
    # accepted-step sequence: same (step, t, dt, nr_iters, accept) records
    gr = g.step_records()
    assert len(gr) == len(rec)
    for a, b in zip(gr, rec):
        assert int(a[0]) == b["step"] and int(a[3]) == b["nr_iters"] and bool(a[4]) == b["accept"]
        # dt is a function of the LTE norm, i.e. of (V_new - V_pred): it can only agree to the accuracy
        # of the solves (1e-9 relative, north_star), not to 1 ulp
        assert abs(a[1] - b["t_next"]) <= 1e-9 * abs(b["t_next"]) and abs(a[2] - b["dt"]) <= 1e-9 * b["dt"]
    assert np.max(np.abs(tg[:n] - to[:n])) <= 1e-9 * to[n - 1]
    span = max(np.max(Vo[:n]) - np.min(Vo[:n]), 1e-12)
    rms = np.sqrt(np.mean((Vg[:n] - Vo[:n]) ** 2)) / span
    assert rms < 1e-6, rms
    assert relerr(Vg[:n], Vo[:n], 1e-6) < 1e-7
    g.close()


def test_handle_ids_and_sizes(bundles):
    """What the FFI handlers rely on: a handle resolves from its id until it is destroyed, ids are stable
    and distinct, vjx_sizes reports the shapes the handlers validate operands against."""
    import ctypes as C

    from vajax_b200 import ffi
    from vajax_b200 import lib as vlib
    from vajax_b200.analysis.engine import B200Engine

    lib = vlib.load()
    lib.vjx_handle_from_id.restype = C.c_void_p
    lib.vjx_handle_from_id.argtypes = [C.c_longlong]
    a = B200Engine(CIRCUITS["ring"](), bundles).prepare()
    b = B200Engine(CIRCUITS["rc"](), bundles).prepare()
    ia, ib = ffi.handle_id(a.handle), ffi.handle_id(b.handle)
    assert ia > 0 and ib > 0 and ia != ib and ffi.handle_id(a.handle) == ia
    assert lib.vjx_handle_from_id(ia) == a.handle.value and lib.vjx_handle_from_id(ib) == b.handle.value
    sz = ffi.sizes(a.handle)
    assert sz["n_aug"] == a.n_aug and sz["nse"] == a.nse and sz["n_total"] == a.n_total
    assert sz["n_unknowns"] == a.n_unknowns and sz["n_vsources"] == a.n_vs
    b.close()
    assert lib.vjx_handle_from_id(ib) is None and lib.vjx_handle_from_id(ia) == a.handle.value
    a.close()


@pytest.mark.parametrize("mt", ["resistor", "capacitor", "diode", "psp103", "opcov", "psp103w"])
def test_init_kernel_against_reference(mt, bundles):
    """N1: the emitted init kernel (vjx_model_init) against the reference code generator's init function on the
    same seeded instances (tests/golden/reference_units.npz model_*_cache / model_*_collapse, produced by
    ``OpenVAFToJAX.translate_init_array`` + vmap in make_golden.py)."""
    from golden_util import GOLDEN
    import os

    from vajax_b200.analysis.engine import B200Engine

    units = np.load(os.path.join(GOLDEN, "reference_units.npz"))
    init_in = units[f"model_{mt}_init_in"]
    e = B200Engine(CIRCUITS["rc"](), bundles)
    e._init_plugins = {}
    cache, col = e._run_init(mt, [np.ascontiguousarray(init_in[:, j]) for j in range(init_in.shape[1])], init_in.shape[0])
    ref = units[f"model_{mt}_cache"]
    cache = cache.cpu().numpy()
    assert cache.shape == ref.shape
    if ref.size:
        assert np.max(np.abs(cache - ref) / np.maximum(np.abs(ref), 1e-300)) < 1e-13
    assert np.array_equal(col.cpu().numpy().astype(np.float64), units[f"model_{mt}_collapse"])


def test_model_kernel_variants_are_bit_identical(bundles, monkeypatch):
    """Output slices (1 / 3 / 6), register budgets (launch bounds), lock-step barriers, the LU grid divisor, the residual-only kernel of the build at the converged
    X (vs. the full kernel, VJX_FULL_FINAL_EVAL) and the parallel stamp / residual branches (vs. the chain, VJX_NO_FORK)
    are scheduling choices: every output is computed by the same sequence of fp64 operations, so the transient is the
    same bit for bit."""
    from vajax_b200.analysis.engine import B200Engine

    circ = CIRCUITS["mul2w"]()
    ref = None
    for kw, env in (({}, {}), (dict(eval_slices=1), {}), (dict(eval_slices=3), {}), (dict(eval_slices=6, eval_min_blocks=3), {}),
                    (dict(eval_slices=1, eval_min_blocks=5), {}), ({}, {"VJX_FULL_FINAL_EVAL": "1"}), ({}, {"VJX_NO_FORK": "1"}),
                    ({}, {"VJX_EVAL_LOCKSTEP": "200"}), (dict(eval_slices=1, eval_min_blocks=4, lu_grid_div=4), {"VJX_EVAL_LOCKSTEP": "500"})):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        for graph in (False, True):
            e = B200Engine(circ, bundles, use_device_graph=graph, **kw).prepare()
            t, V, st = e.run_transient()
            rec = e.step_records()
            e.close()
            if ref is None:
                ref = (t, V, st["n_steps"], st["total_nr_iterations"], rec)
            else:
                assert st["n_steps"] == ref[2] and st["total_nr_iterations"] == ref[3], (kw, env, graph)
                assert np.array_equal(t, ref[0]) and np.array_equal(V, ref[1]) and np.array_equal(rec, ref[4]), (kw, env, graph)
        for k in env:
            monkeypatch.delenv(k)
