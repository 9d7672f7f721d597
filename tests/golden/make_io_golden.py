#!/usr/bin/env python
"""Generate tests/golden/io/* with the REFERENCE's own writers (vajax/io/rawfile_writer.py,
vajax/io/csv_writer.py), loaded by file path so that ``vajax/__init__`` (which imports jax) never runs.

    python tests/golden/make_io_golden.py        # build container only (/root/reference)

The input is a fixed synthetic result (seeded), also returned by ``golden_result()`` for the tests."""
import importlib.util
import os
import sys
from types import SimpleNamespace

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("VAJAX_REFERENCE", "/root/reference")


def golden_result(kind: str = "tran"):
    rng = np.random.default_rng(20261018)
    n = 37
    t = np.cumsum(rng.uniform(1e-13, 4e-11, n))
    t[0] = 0.0
    names = ["vdd", "p0", "a1", "Out_7", "x1.n3", "n,odd"]        # unsorted on purpose; one name csv must quote
    volts = {nm: rng.normal(0.6, 0.5, n) * (10.0 ** rng.integers(-12, 3)) for nm in names}
    volts["p0"][3] = 0.0
    volts["a1"][5] = -1.2345678901234567e-300
    volts["vdd"][:] = 1.2
    if kind == "ac":
        return SimpleNamespace(times=t, frequencies=np.logspace(0, 9, n), voltages=volts)
    if kind == "empty":
        return SimpleNamespace(times=np.zeros(0), voltages={"a": np.zeros(0)})
    return SimpleNamespace(times=t, voltages=volts)


def _load(rel: str):
    spec = importlib.util.spec_from_file_location("ref_" + os.path.basename(rel)[:-3], os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    raw = _load("vajax/io/rawfile_writer.py")
    csvw = _load("vajax/io/csv_writer.py")
    out = os.path.join(HERE, "io")
    os.makedirs(out, exist_ok=True)
    for kind in ("tran", "ac", "empty"):
        r = golden_result(kind)
        raw.write_rawfile(r, os.path.join(out, f"{kind}_binary.raw"), binary=True)
        raw.write_rawfile(r, os.path.join(out, f"{kind}_ascii.raw"), binary=False)
        csvw.write_csv(r, os.path.join(out, f"{kind}.csv"))
    csvw.write_csv(golden_result("tran"), os.path.join(out, "tran_p4.csv"), precision=4)
    print("wrote", sorted(os.listdir(out)))


if __name__ == "__main__":
    sys.exit(main())
