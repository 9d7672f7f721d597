"""I-V / C-V probe of a MOSFET fixture model through the oracle (test infrastructure; not on the product path).

    python scripts/psp103w_iv.py [psp103w|psp103]

One circuit with a grid of (vgs, vds) bias points, one transistor + two DC sources per point; a single DC
``nr_solve`` returns the whole surface as the source currents."""
import sys

import numpy as np

sys.path.insert(0, ".")
from oracle.engine import OracleEngine  # noqa: E402
from vajax_b200.benchmarks.circuits import _Net  # noqa: E402
from vajax_b200.models.fixtures import get_bundle  # noqa: E402


def surface(model, kind, w, l, vgs_list, vds_list):
    net = _Net(model)
    sg = 1.0 if kind == "n" else -1.0
    pts = [(a, b) for a in vgs_list for b in vds_list]
    for k, (vg, vd) in enumerate(pts):
        net.add(f"vd{k}", "vsource", [f"d{k}", "0"], dict(dc=sg * vd, type="dc"), openvaf=False)
        net.add(f"vg{k}", "vsource", [f"g{k}", "0"], dict(dc=sg * vg, type="dc"), openvaf=False)
        net.mos(f"m{k}", kind, f"d{k}", f"g{k}", "0", "0", w, l)
    c = net.circuit(dict(step=1e-12, stop=1e-9, icmode="op", tran_method="trap"))
    o = OracleEngine(c, {k: get_bundle(k) for k in ("resistor", "capacitor", "diode", "psp103", "psp103w")})
    o.setup()
    vs, isrc = o.eval_sources(0.0)
    X0 = np.zeros(o.n_total + o.n_vs)
    X, it, conv, max_f, Q, dQdt, Ivs, ls, mrc = o.nr_solve(X0, vs, isrc, np.zeros(o.n_unknowns), 0.0, max_iterations=100)
    Id = -Ivs[0::2].reshape(len(vgs_list), len(vds_list)) * sg
    Ig = -Ivs[1::2].reshape(len(vgs_list), len(vds_list)) * sg
    return Id, Ig, it, conv


if __name__ == "__main__":
    model = sys.argv[1] if len(sys.argv) > 1 else "psp103w"
    vg = np.linspace(0, 1.2, 7)
    vd = np.array([0.0, 0.05, 0.1, 0.2, 0.4, 0.6, 0.8, 1.0, 1.2])
    for kind, w in (("n", 0.5e-6), ("p", 1e-6)):
        Id, Ig, it, conv = surface(model, kind, w, 0.2e-6, vg, vd)
        print(f"{model} {kind}mos w={w} l=0.2u: DC NR iterations {it} converged {conv}")
        print("vds     " + " ".join(f"{v:9.2f}" for v in vd))
        for i, g in enumerate(vg):
            print(f"vgs {g:4.2f} " + " ".join(f"{x:9.2e}" for x in Id[i]))
        print("Ig (vds=1.2): " + " ".join(f"{x:9.2e}" for x in Ig[:, -1]))
