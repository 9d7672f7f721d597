#!/usr/bin/env python
"""GPU debug: drive a continuation case with the ORACLE, call the GPU nr_solve with the same inputs at every solve
(with and without a fresh LU analysis at that solve's Jacobian) and print the differences.  Test / tuning tool."""
import copy, ctypes as C, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from golden_util import CIRCUITS, GOLDEN
from oracle.engine import OracleEngine
from vajax_b200 import lib as vlib
from vajax_b200.analysis import dc_op
from vajax_b200.analysis.engine import B200Engine
from vajax_b200.models.fixtures import get_bundle

G = np.load(f"{GOLDEN}/reference_homotopy.npz")
bundles = {k: get_bundle(k) for k in ["resistor", "capacitor", "diode", "psp103", "opcov", "psp103w"]}
for case in sys.argv[1:] or ["ring_chain_default", "ring_gdev_rand_itl4", "mul3_gdev_itl6"]:
    meta = json.loads(str(G[f"{case}_result"]))
    circ = copy.deepcopy(CIRCUITS[meta["circuit"]]())
    circ["analysis"]["tran_itl"] = meta["tran_itl"]
    o = OracleEngine(circ, bundles); o.setup()
    e = B200Engine(circ, bundles).prepare()
    e2 = B200Engine(circ, bundles).prepare()
    print("==", case, meta)

    def both(X, vs, isrc, Qp, c0, gmin, gshunt, *rest):
        ro = o.nr_solve(X, vs, isrc, Qp, c0, gmin, gshunt, *rest)
        rg = e.nr_solve(X, vs, isrc, Qp, c0, gmin, gshunt, *rest)
        # fresh analysis at this solve's first Jacobian
        J = e2.build_system(X, vs, isrc, Qp, c0, gmin, gshunt, 0.0, 0.0, None, 0.0, None, None, 0)[0]
        Jd = e2._dev(J)
        vlib.check(vlib.load().vjx_lu_analyze(e2.handle, C.c_void_p(Jd.data_ptr()), e2._sarg()))
        rf = e2.nr_solve(X, vs, isrc, Qp, c0, gmin, gshunt, *rest)
        Jo = o.build_system(X, vs, isrc, Qp, c0, gmin, gshunt, 0.0, 0.0, np.zeros(o.n_unknowns), 0.0, np.zeros(o.n_unknowns), np.zeros(o.total_limit_states), 0)
        import scipy.sparse as sp
        A = sp.csr_matrix((np.asarray(Jo[0]), o.csr_indices, o.csr_indptr)).toarray()
        sv = np.linalg.svd(A, compute_uv=False)
        print(f"  gmin={gmin:.2e} gshunt={gshunt:.2e} vs0={vs[0] if len(vs) else 0:.3f} | oracle it={ro[1]} cv={bool(ro[2])} | gpu it={rg[1]} cv={rg[2]} dX={np.max(np.abs(rg[0]-ro[0])):.3e}"
              f" | gpu(fresh analysis) it={rf[1]} cv={rf[2]} dX={np.max(np.abs(rf[0]-ro[0])):.3e} | cond={sv[0]/sv[-1]:.3e} |X|max={np.max(np.abs(ro[0])):.3f}", flush=True)
        return ro

    cfg = dc_op.HomotopyConfig(gmin=1e-12, **meta["config"])
    vs, isv = (np.asarray(a, dtype=np.float64) for a in o.eval_sources(0.0))
    X0 = G[f"{case}_X0"]; Qz = np.zeros(o.n_unknowns)
    if meta["entry"] == "chain":
        dc_op.run_homotopy_chain(both, X0, vs, isv, Qz, cfg)
    elif meta["entry"] in ("gdev", "gshunt"):
        dc_op.gmin_stepping(both, X0, vs, isv, Qz, 1.0, cfg, meta["entry"])
    else:
        dc_op.source_stepping(both, X0, vs, isv, Qz, cfg)
    e.close(); e2.close()
