"""N4 output formats: raw / CSV files byte-identical to the ones the reference's own writers produced
(tests/golden/io/*, made by tests/golden/make_io_golden.py from vajax/io/rawfile_writer.py and
vajax/io/csv_writer.py), round trips, and the named-waveform view of a run (extract_results)."""
import filecmp
import importlib.util
import os

import numpy as np
import pytest

from vajax_b200.io import (TransientResult, extract_results, read_csv, read_rawfile, read_rawfile_header,
                           write_csv, write_rawfile)

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "io")
_spec = importlib.util.spec_from_file_location("make_io_golden", os.path.join(HERE, "golden", "make_io_golden.py"))
_gen = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_gen)


@pytest.mark.parametrize("kind", ["tran", "ac", "empty"])
@pytest.mark.parametrize("binary", [True, False])
def test_rawfile_bytes_equal_reference(tmp_path, kind, binary):
    out = tmp_path / "o.raw"
    write_rawfile(_gen.golden_result(kind), out, binary=binary)
    ref = os.path.join(GOLD, f"{kind}_{'binary' if binary else 'ascii'}.raw")
    assert filecmp.cmp(out, ref, shallow=False), f"{kind} binary={binary} differs from the reference writer's file"


@pytest.mark.parametrize("kind,precision,name", [("tran", 9, "tran.csv"), ("ac", 9, "ac.csv"), ("empty", 9, "empty.csv"),
                                                 ("tran", 4, "tran_p4.csv")])
def test_csv_bytes_equal_reference(tmp_path, kind, precision, name):
    out = tmp_path / "o.csv"
    write_csv(_gen.golden_result(kind), out, precision=precision)
    assert filecmp.cmp(out, os.path.join(GOLD, name), shallow=False)


@pytest.mark.parametrize("binary", [True, False])
def test_rawfile_round_trip(tmp_path, binary):
    r = _gen.golden_result("tran")
    p = tmp_path / "o.raw"
    write_rawfile(r, p, binary=binary)
    h = read_rawfile_header(p)
    assert h["title"] == "VAJAX Simulation" and h["plotname"] == "Transient Analysis"
    assert h["n_vars"] == 1 + len(r.voltages) and h["n_points"] == len(r.times) and h["binary"] is binary
    assert h["variables"][0] == ("time", "time")
    assert [v[0] for v in h["variables"][1:]] == list(r.voltages.keys())
    back = read_rawfile(p)
    assert back["x_name"] == "time"
    # binary carries the doubles themselves; %.15e holds 16 significant digits
    tol = 0.0 if binary else 5e-16
    np.testing.assert_allclose(back["x"], r.times, rtol=tol, atol=0.0)
    for nm, v in r.voltages.items():
        np.testing.assert_allclose(back["voltages"][nm], v, rtol=tol, atol=0.0)


def test_reads_the_reference_files():
    """Files written by the reference parse with this package's readers."""
    r = _gen.golden_result("tran")
    back = read_rawfile(os.path.join(GOLD, "tran_binary.raw"))
    np.testing.assert_array_equal(back["x"], r.times)
    for nm, v in r.voltages.items():
        np.testing.assert_array_equal(back["voltages"][nm], v)
    c = read_csv(os.path.join(GOLD, "tran.csv"))
    assert sorted(c["voltages"]) == sorted(r.voltages)
    np.testing.assert_allclose(c["times"], r.times, rtol=1e-9)
    np.testing.assert_allclose(c["voltages"]["n,odd"], r.voltages["n,odd"], rtol=1e-9)
    a = read_csv(os.path.join(GOLD, "ac.csv"))
    assert "frequencies" in a and "times" not in a


def test_extract_results_slices_and_names():
    """full_mna.py:143-175: first n_steps rows, one array per name, currents only if there are sources."""
    rng = np.random.default_rng(3)
    T, V, I = rng.random(10), rng.random((10, 4)), rng.random((10, 2))
    stats = {"n_steps": 6, "node_indices": {"a": 1, "B": 3}, "I_out": I, "current_indices": {"v1": 0, "vdd": 1}}
    t, v, i = extract_results(T, V, stats)
    assert t.shape == (6,) and np.array_equal(v["a"], V[:6, 1]) and np.array_equal(v["B"], V[:6, 3])
    assert np.array_equal(i["vdd"], I[:6, 1])
    res = TransientResult.from_run(T, V, stats)
    assert res.num_steps == 6 and np.array_equal(res.voltage("b"), V[:6, 3]) and np.array_equal(res.current("VDD"), I[:6, 1])
    with pytest.raises(KeyError):
        res.voltage("nope")
    _, _, none = extract_results(T, V, {"n_steps": 6, "node_indices": {"a": 1}, "I_out": I, "current_indices": {}})
    assert none == {}


def test_save_list_columns_write_through(tmp_path):
    """With a device-side save list V_out has one column per saved node and node_indices maps into them
    (engine.run_transient); the writers only see names."""
    V = np.arange(20.0).reshape(5, 4)
    stats = {"n_steps": 5, "node_indices": {"p0": 2, "p1": 0}}
    res = TransientResult.from_run(np.linspace(0, 1e-9, 5), V, stats)
    write_rawfile(res, tmp_path / "s.raw")
    back = read_rawfile(tmp_path / "s.raw")
    assert list(back["voltages"]) == ["p0", "p1"]
    np.testing.assert_array_equal(back["voltages"]["p0"], V[:, 2])
    write_csv(res, tmp_path / "s.csv")
    np.testing.assert_allclose(read_csv(tmp_path / "s.csv")["voltages"]["p1"], V[:, 0])
