#!/usr/bin/env python
"""Where does a dependency chain spend its time?  Runs each LU phase once with per-row completion
timestamps (vjx_lu_timeline) on the c6288-shaped circuit and prints, per block of levels, when the
last row of the block finished.  GPU box only:  python scripts/lu_timeline.py [nbits]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vajax_b200 import lib as vlib  # noqa: E402
from vajax_b200.analysis.engine import B200Engine  # noqa: E402
from vajax_b200.benchmarks import circuits  # noqa: E402
from vajax_b200.models.fixtures import get_bundle  # noqa: E402

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 16
bundles = {k: get_bundle(k) for k in ("resistor", "capacitor", "diode", "psp103", "psp103w")}
model = sys.argv[2] if len(sys.argv) > 2 else "psp103w"
circ = circuits.mul64_circuit(nb, mos_model=model) if nb == 64 else circuits.multiplier_circuit(nb, mos_model=model)
e = B200Engine(circ, bundles, use_device_graph=True).prepare()
e.run_transient(max_steps=6 if nb == 64 else 12)
st = e.lu_stats()
n = st["n"]
print("lu", st)
lib = vlib.load()
for phase, name in ((3, "factor+fwd"), (5, "usolve")):
    best = None
    for rep in range(3):
        ns = np.zeros(n)
        lv = np.zeros(n, dtype=np.int32)
        vlib.check(lib.vjx_lu_timeline(e.handle, phase, ns.ctypes.data_as(C.c_void_p), lv.ctypes.data_as(C.c_void_p), None))
        if best is None or ns.max() < best[0].max():
            best = (ns, lv)
    ns, lv = best
    ns = np.where(ns < 0, 0.0, ns)   # rows finished by an earlier kernel (dense trailing block) carry no stamp
    nl = lv.max() + 1
    print(f"== {name}: {nl} levels, last row done at {ns.max() / 1e3:.1f} us, first at {ns.min() / 1e3:.1f} us")
    tail = np.arange(n) >= (n - 185)
    for lo in range(0, nl, max(nl // 32, 1)):
        hi = min(lo + max(nl // 32, 1), nl)
        m = (lv >= lo) & (lv < hi)
        if m.any():
            print(f"   levels {lo:4d}-{hi - 1:4d}: rows {m.sum():6d}  done {ns[m].min() / 1e3:8.1f} .. {ns[m].max() / 1e3:8.1f} us"
                  f"   (tail rows {int((m & tail).sum())})")
    if os.environ.get("VJX_TL_DETAIL") and phase == 3:   # "lo:hi": every row of these levels (index, done us)
        lo, hi = (int(x) for x in os.environ["VJX_TL_DETAIL"].split(":"))
        for L in range(lo, hi):
            m = np.where(lv == L)[0]
            print(f"   level {L:4d}:", [(int(i), round(ns[i] / 1e3, 1)) for i in m[np.argsort(ns[m])][:12]])
    order = np.argsort(ns)
    print("   last 12 rows (index, level, us):", [(int(i), int(lv[i]), round(ns[i] / 1e3, 1)) for i in order[-12:]])
e.close()
