#!/usr/bin/env python
"""CPU-only look at the LU dependency structure of a benchmark circuit (no GPU needed): builds the
Jacobian with the oracle at a mid-rail state, runs the library's symbolic analysis and prints the
per-level row counts / lengths.  python scripts/lu_structure.py [nbits|64]  (test/tuning tool)"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.engine import OracleEngine  # noqa: E402
from vajax_b200 import lib as vlib  # noqa: E402
from vajax_b200.benchmarks.circuits import mul64_circuit, multiplier_circuit  # noqa: E402
from vajax_b200.models.fixtures import get_bundle  # noqa: E402

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 16
model = sys.argv[2] if len(sys.argv) > 2 else "psp103w"
keep_noi = len(sys.argv) > 3 and sys.argv[3] == "keep-noi"   # analyse the unfiltered pattern (what round 1 did)
bundles = {k: get_bundle(k) for k in ("resistor", "capacitor", "diode", "psp103", "psp103w")}
circ = mul64_circuit(nb, mos_model=model) if nb == 64 else multiplier_circuit(nb, mos_model=model)
t0 = time.time()
o = OracleEngine(circ, bundles)
o.setup()
print(f"oracle setup {time.time()-t0:.1f}s n_aug={o.n_aug} nse={o.nse}", flush=True)
X = np.zeros(o.n_total + o.n_vs)
X[:o.n_total] = o.init_mid_rail()
rng = np.random.default_rng(1)
X[1:o.n_total] += rng.normal(scale=0.05, size=o.n_total - 1)
vs, isv = o.eval_sources(1e-10)
nu = o.n_unknowns
c0 = 2.0 / 2e-12
out = o.build_system(X, vs, isv, np.zeros(nu), c0, 1e-12, 0.0, -c0, -1.0, np.zeros(nu), 0.0, np.zeros(nu),
                     np.zeros(o.total_limit_states), 1)
csr = np.array(out[0], dtype=np.float64)
ip, ix = o.csr_indptr.astype(np.int32), o.csr_indices.astype(np.int32)
rows = np.repeat(np.arange(o.n_aug), np.diff(ip))
noi = np.zeros(o.n_aug, bool)
if len(o.noi_indices):
    noi[np.asarray(o.noi_indices) - 1] = True
zero = (noi[rows] | ((rows < nu) & noi[ix])) & ~(noi[rows] & (rows == ix))
csr[zero] = 0.0
csr[noi[rows] & (rows == ix)] = 1.0
np.savez("/tmp/jac_%s_%d.npz" % (model, nb), indptr=ip, indices=ix, data=csr)
if not keep_noi:
    # the slots enforce_noi rewrites to zero are structural zeros for the factorisation (vjx_engine.cu ensure_lu)
    keep = ~zero
    ip = np.concatenate([[0], np.cumsum(np.bincount(rows[keep], minlength=o.n_aug))]).astype(np.int32)
    ix, csr = ix[keep].copy(), csr[keep].copy()
    print(f"pattern without the NOI-zero slots: {len(ix)} of {len(zero)} entries")
print(f"jacobian built {time.time()-t0:.1f}s", flush=True)
lib = vlib.load()
n = o.n_aug
lv_l, lv_u, l_len, u_len, waves = (np.zeros(n, np.int32) for _ in range(5))
p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
lib.vjx_symbolic_rows.argtypes = [C.c_int] + [C.c_void_p] * 3 + [C.c_double] + [C.c_void_p] * 5
t1 = time.time()
rc = lib.vjx_symbolic_rows(n, p(ip), p(ix), p(csr), 1e-3, p(lv_l), p(lv_u), p(l_len), p(u_len), p(waves))
print("symbolic rc", rc, f"{time.time()-t1:.1f}s")
nl = lv_l.max() + 1
print("levels", nl, "nnz L", l_len.sum(), "nnz U", u_len.sum(), "max l", l_len.max(), "max u", u_len.max())
cnt = np.bincount(lv_l, minlength=nl)
ll = np.bincount(lv_l, weights=l_len, minlength=nl)
ul = np.bincount(lv_l, weights=u_len, minlength=nl)
mx = np.zeros(nl)
np.maximum.at(mx, lv_l, l_len)
print("level: rows  sumL  sumU  maxL")
step = max(1, nl // 60)
for a in range(0, nl, step):
    b = min(nl, a + step)
    print(f"{a:5d}-{b-1:5d}: rows={int(cnt[a:b].sum()):7d} sumL={int(ll[a:b].sum()):9d} sumU={int(ul[a:b].sum()):9d} maxL={int(mx[a:b].max()):5d}")
single = np.where(cnt == 1)[0]
print("levels with one row:", len(single), " first such:", single[:5])
# heavy rows with the dense trailing block in place: sequential k-steps today vs dependency depths
bulk, nd = np.zeros(n, np.int32), C.c_int(0)
lib.vjx_symbolic_rows2.argtypes = [C.c_int] + [C.c_void_p] * 3 + [C.c_double, C.c_int] + [C.c_void_p] * 6 + [C.c_void_p]
rc = lib.vjx_symbolic_rows2(n, p(ip), p(ix), p(csr), 1e-3, 159, p(lv_l), p(lv_u), p(l_len), p(u_len), p(waves), p(bulk), C.byref(nd))
print(f"dense trailing block: {nd.value} rows")
np.savez(f"/tmp/lu_rows_{model}_{nb}.npz", level_l=lv_l, level_u=lv_u, l_len=l_len, u_len=u_len, waves=waves, bulk=bulk)
heavy = np.argsort(-bulk)[:12]
print("heaviest rows: index  k-steps of k_lu_rows  distinct dependency depths among them  (dense-block row?)")
for i in heavy:
    print(f"   {int(i):8d} {int(bulk[i]):6d} {int(waves[i]):6d}   {'dense' if i >= n - nd.value else ''}")
hl = bulk >= 32
if hl.any():
    print(f"rows with >= 32 k-steps: {int(hl.sum())}, k-steps {int(bulk[hl].sum())}, depths {int(waves[hl].sum())} "
          f"(ratio {bulk[hl].sum() / max(waves[hl].sum(), 1):.1f}x)")
