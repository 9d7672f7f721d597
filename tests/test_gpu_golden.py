"""GPU parity against the committed golden fixtures (outputs of the reference's own modules,
tests/golden/make_golden.py) - the CUDA path through the C-ABI, no oracle in between.
Bars (north_star): stamped Jacobian 1e-12 relative, voltages 1e-9, waveform 1e-6 RMS with the same
accepted-step sequence."""
import numpy as np
import pytest

from golden_util import BS_CASES, BS_NAMES, CIRCUITS, assert_vector_parity, bs_inputs, load, scaled_err, tag, tran_stats

pytestmark = pytest.mark.gpu


def _engine(name, bundles, limiting, graph=False):
    from vajax_b200.analysis.engine import B200Engine

    return B200Engine(CIRCUITS[name](), bundles, use_device_limiting=limiting, use_device_graph=graph).prepare()


@pytest.mark.parametrize("name", list(CIRCUITS))
@pytest.mark.parametrize("limiting", [False, True])
def test_build_system_against_reference(name, limiting, bundles):
    g = load(name)
    t = tag(name, limiting)
    e = _engine(name, bundles, limiting)
    n_total, nu, n_vs, n_lim = g[t + "_sizes"]
    assert (e.n_total, e.n_unknowns, e.n_vs, e.total_limit_states) == (n_total, nu, n_vs, n_lim)
    X, vs, isv, Qp, dq, q2, ls = bs_inputs(g, t, e.n_vs)
    for ci, (c0, c1, d1, c2, it) in enumerate(BS_CASES):
        out = e.build_system(X, vs, isv, Qp, c0, 1e-12, 1e-9, c1, d1, dq, c2, q2, ls, it)
        for nm, a in zip(BS_NAMES, out):
            assert_vector_parity(nm, a, g[f"{t}_bs{ci}_{nm}"], (t, ci))
    # deterministic: a second evaluation is bit-identical (no floating-point atomics)
    again = e.build_system(X, vs, isv, Qp, c0, 1e-12, 1e-9, c1, d1, dq, c2, q2, ls, it)
    for a, b in zip(out, again):
        assert np.array_equal(a, b)
    e.close()


@pytest.mark.parametrize("name", list(CIRCUITS))
@pytest.mark.parametrize("graph", [False, True])
def test_transient_against_reference(name, graph, bundles):
    g = load(name)
    t = tag(name, False)
    e = _engine(name, bundles, False, graph)
    times, V, st = e.run_transient()
    ref_t, ref_V, ref_st = g[t + "_tran_times"], g[t + "_tran_V"], tran_stats(g, t)
    n = st["n_steps"]
    assert st["error_flags"] == 0
    assert n == len(ref_t)
    assert st["rejected_steps"] == int(ref_st["rejected_steps"])
    assert st["total_nr_iterations"] == int(ref_st["total_nr_iterations"])
    assert scaled_err(times[:n], ref_t, rel_floor=1e-30, abs_floor=1e-30) < 1e-9
    span = max(float(np.max(ref_V) - np.min(ref_V)), 1e-12)
    assert np.sqrt(np.mean((V[:n] - ref_V) ** 2)) / span < 1e-6
    assert float(np.max(np.abs(V[:n] - ref_V))) < 1e-7 * max(1.0, float(np.max(np.abs(ref_V))))
    e.close()


def test_transient_with_limiting_against_reference(bundles):
    g = load("clipper")
    t = tag("clipper", True)
    e = _engine("clipper", bundles, True, True)
    times, V, st = e.run_transient()
    ref_t, ref_V, ref_st = g[t + "_tran_times"], g[t + "_tran_V"], tran_stats(g, t)
    n = st["n_steps"]
    assert n == len(ref_t) and st["total_nr_iterations"] == int(ref_st["total_nr_iterations"])
    assert float(np.max(np.abs(V[:n] - ref_V))) < 1e-7
    e.close()


def test_result_object_and_raw_csv_files(bundles, tmp_path):
    """N4 output path end to end: device run -> TransientResult (reference: engine.py:985-1024) -> raw / CSV
    files -> read back; the named waveforms are the reference golden's columns."""
    from vajax_b200.io import read_csv, read_rawfile, write_csv, write_rawfile

    g = load("rc")
    t = tag("rc", False)
    e = _engine("rc", bundles, False, True)
    res = e.run_transient_result()
    ref_t, ref_V = g[t + "_tran_times"], g[t + "_tran_V"]
    assert res.num_steps == len(ref_t) and set(res.voltages) and set(res.currents)
    idx = res.stats["node_indices"]
    for nm, col in idx.items():
        assert float(np.max(np.abs(res.voltage(nm) - ref_V[:, col]))) < 1e-7
    write_rawfile(res, tmp_path / "rc.raw", binary=True)
    back = read_rawfile(tmp_path / "rc.raw")
    np.testing.assert_array_equal(back["x"], res.times)
    for nm in idx:
        np.testing.assert_array_equal(back["voltages"][nm], res.voltage(nm))
    write_csv(res, tmp_path / "rc.csv")
    c = read_csv(tmp_path / "rc.csv")
    for nm in idx:
        np.testing.assert_allclose(c["voltages"][nm], res.voltage(nm), rtol=1e-9, atol=1e-300)
    # with a device-side save list only the saved column exists, under its name
    keep = sorted(idx)[:1]
    e.set_save_list(keep)
    res1 = e.run_transient_result()
    assert list(res1.voltages) == keep
    np.testing.assert_array_equal(res1.voltage(keep[0]), res.voltage(keep[0]))
    e.close()


@pytest.mark.parametrize("name", ["rc", "clipper", "ring", "mul3", "ringw"])
def test_dense_path_transient_against_reference(name, bundles):
    """A12: the reference's dense linear-solve path (solve(J + 1e-14 I, -f), solver_factories.py:728-800 - its default
    for circuits this small) through B200Engine(linear_solver="dense"): the reference's DENSE-path goldens, same
    accepted-step sequence.  The ring is the case where the two paths part: 67 accepted steps dense, 64 sparse."""
    from vajax_b200.analysis.engine import B200Engine

    g = load(name)
    t = tag(name, False, sparse=False)
    e = B200Engine(CIRCUITS[name](), bundles, use_device_graph=True, linear_solver="dense").prepare()
    times, V, st = e.run_transient()
    ref_t, ref_V, ref_st = g[t + "_tran_times"], g[t + "_tran_V"], tran_stats(g, t)
    n = st["n_steps"]
    assert st["error_flags"] == 0
    assert n == len(ref_t), (n, len(ref_t))
    assert st["rejected_steps"] == int(ref_st["rejected_steps"])
    assert st["total_nr_iterations"] == int(ref_st["total_nr_iterations"])
    assert scaled_err(times[:n], ref_t, rel_floor=1e-30, abs_floor=1e-30) < 1e-9
    assert float(np.max(np.abs(V[:n] - ref_V))) < 1e-7 * max(1.0, float(np.max(np.abs(ref_V))))
    if name == "ring":
        assert n == 67 and len(g[tag(name, False) + "_tran_times"]) == 64
    e.close()
