"""CPU-side tests of the product's host logic: the C-ABI library loads and exports every symbol
declared in include/vajax_b200.h, refuses to run without a GPU (no CPU fallback), the symbolic LU
analysis (C++) is correct on MNA-shaped matrices, the CUDA-C emitter produces compilable sources,
and the corner tables / sharding mirror the reference."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from vajax_b200 import lib as vlib

    vlib.build_library()
    return vlib.load()


def test_cabi_exports_every_declared_symbol(lib):
    from vajax_b200 import lib as vlib

    hdr = open(os.path.join(ROOT, "include", "vajax_b200.h")).read()
    declared = set(re.findall(r"\b(vjx_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"vjx_handle", "vjx_stream"}
    assert len(declared) >= 20
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    assert set(vlib.EXPORTS) <= declared | {"vjx_symbolic_probe"}, set(vlib.EXPORTS) - declared
    assert lib.vjx_version() >= 100


def test_no_cpu_fallback(lib, bundles):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from vajax_b200 import lib as vlib
    from vajax_b200.analysis.engine import B200Engine
    from vajax_b200.benchmarks.circuits import rc_circuit

    assert lib.vjx_device_count() == 0
    desc = vlib.CircuitDesc()
    h = C.c_void_p(None)
    rc = lib.vjx_create(C.byref(desc), 0, C.byref(h))
    assert rc != 0 and not h.value
    with pytest.raises(vlib.VjxError):
        B200Engine(rc_circuit(), bundles).prepare()


def _probe(lib, A, b, tol=1e-3):
    A = A.tocsr()
    A.sort_indices()
    n = A.shape[0]
    ip = np.ascontiguousarray(A.indptr, dtype=np.int32)
    ix = np.ascontiguousarray(A.indices, dtype=np.int32)
    dv = np.ascontiguousarray(A.data, dtype=np.float64)
    x, st = np.zeros(n), np.zeros(12)
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    rc = lib.vjx_symbolic_probe(n, p(ip), p(ix), p(dv), tol, p(np.ascontiguousarray(b)), p(x), p(st))
    keys = "nnz_lu flops lev_l lev_u max_row n resid scatter n_single n_core max_l_len".split()
    return rc, x, dict(zip(keys, st))


def _mna_like(n_nodes, n_vs, seed):
    """Random conductance matrix + voltage-source incidence rows/cols (zero diagonal block)."""
    import scipy.sparse as sp

    rng = np.random.default_rng(seed)
    rows, cols, vals = [], [], []
    for _ in range(3 * n_nodes):
        a, b = rng.integers(0, n_nodes, 2)
        g = rng.uniform(1e-4, 1e-2)
        for r, c, v in ((a, a, g), (b, b, g), (a, b, -g), (b, a, -g)):
            rows.append(r); cols.append(c); vals.append(v)
    for i in range(n_nodes):
        rows.append(i); cols.append(i); vals.append(1e-12)
    vs_nodes = rng.choice(n_nodes, size=n_vs, replace=False)  # two sources on one node would be singular
    for v in range(n_vs):
        node = int(vs_nodes[v])
        for r, c in ((node, n_nodes + v), (n_nodes + v, node)):
            rows.append(r); cols.append(c); vals.append(1.0)
    n = n_nodes + n_vs
    return sp.csr_matrix((vals, (rows, cols)), shape=(n, n))


@pytest.mark.parametrize("n_nodes,n_vs,seed", [(1, 1, 0), (3, 1, 1), (40, 3, 2), (400, 12, 3), (3000, 40, 4)])
def test_symbolic_analysis_against_superlu(lib, n_nodes, n_vs, seed):
    import scipy.sparse.linalg as spla

    A = _mna_like(n_nodes, n_vs, seed)
    b = np.random.default_rng(seed).normal(size=A.shape[0])
    rc, x, st = _probe(lib, A, b)
    assert rc == 0
    xr = spla.splu(A.tocsc()).solve(b)
    assert np.max(np.abs(x - xr)) <= 1e-9 * max(1.0, np.max(np.abs(xr)))
    assert st["n_single"] + st["n_core"] == A.shape[0]
    assert st["nnz_lu"] >= A.nnz and st["n"] == A.shape[0]


@pytest.mark.parametrize("n_nodes,n_vs,seed,hv_min,dense_cap", [
    (40, 3, 2, 1, 0), (40, 3, 2, 4, 159), (400, 12, 3, 1, 0), (400, 12, 3, 4, 0), (400, 12, 3, 4, 159),
    (400, 12, 3, 16, 48), (3000, 40, 4, 8, 159)])
def test_heavy_row_program_against_superlu(lib, monkeypatch, n_nodes, n_vs, seed, hv_min, dense_cap):
    """The heavy-row program (k-steps regrouped by dependency depth and target position, vjx_build_heavy) walked by
    the host emulation of the thread block that executes it on the GPU: same solution as SuperLU, with every row that
    has at least hv_min k-steps going through the program, with and without a dense trailing block."""
    import scipy.sparse.linalg as spla

    A = _mna_like(n_nodes, n_vs, seed)
    b = np.random.default_rng(seed).normal(size=A.shape[0])
    rc0, x0, st0 = _probe(lib, A, b)
    monkeypatch.setenv("VJX_PROBE_HV_MIN", str(hv_min))
    monkeypatch.setenv("VJX_PROBE_DENSE_CAP", str(dense_cap))
    rc, x, st = _probe(lib, A, b)
    assert rc == 0 and rc0 == 0
    xr = spla.splu(A.tocsc()).solve(b)
    assert np.max(np.abs(x - xr)) <= 1e-9 * max(1.0, np.max(np.abs(xr)))
    # same factorisation up to the order of the sums
    assert np.max(np.abs(x - x0)) <= 1e-10 * max(1.0, np.max(np.abs(x0)))


def test_symbolic_peels_supply_rails(lib):
    """A rail node pinned by a vsource and touched by every device must leave the core: the branch
    row is a row singleton, the branch-current column a column singleton."""
    import scipy.sparse as sp

    n_int = 200
    rows, cols, vals = [], [], []
    rail, br = n_int, n_int + 1
    for i in range(n_int):
        j = (i + 1) % n_int
        for r, c, v in ((i, i, 3e-3), (i, j, -1e-3), (j, i, -1e-3), (i, rail, -1e-3), (rail, i, -1e-3)):
            rows.append(r); cols.append(c); vals.append(v)
    rows += [rail, rail, br]; cols += [rail, br, rail]; vals += [0.2, 1.0, 1.0]
    A = sp.csr_matrix((vals, (rows, cols)), shape=(n_int + 2, n_int + 2))
    b = np.arange(1.0, n_int + 3.0)
    rc, x, st = _probe(lib, A, b)
    assert rc == 0 and st["n_single"] >= 2 and st["n_core"] <= n_int
    assert np.max(np.abs(A @ x - b)) < 1e-9 * np.max(np.abs(b))
    # the 200-long rail row never becomes an L row (the depth-aware ordering eliminates the 200-ring like a cyclic
    # reduction: ~log2(200) levels, L parts of up to ~2 log2(200) entries instead of a 200-level chain)
    assert st["max_l_len"] < 32


def test_symbolic_rejects_structurally_singular(lib):
    import scipy.sparse as sp

    A = sp.csr_matrix(np.array([[1.0, 2.0, 0.0], [0.0, 0.0, 0.0], [0.0, 3.0, 1.0]]))
    A = sp.csr_matrix(([1.0, 2.0, 3.0, 1.0], ([0, 0, 2, 2], [0, 1, 1, 2])), shape=(3, 3))
    rc, _, _ = _probe(lib, A, np.ones(3))
    assert rc != 0


@pytest.mark.parametrize("mt", ["resistor", "diode", "psp103"])
def test_emitter_sources_compile(mt, bundles):
    """MIR -> CUDA-C for init and eval; nvcc cross-compiles for sm_100a without a GPU."""
    from vajax_b200.codegen import cuda_emitter as em

    b = bundles[mt]
    src, meta = em.emit_init_source(b)
    assert "vjx_init_kernel" in src and meta["cache_size"] == len(b["init_mir_data"]["cache_mapping"])
    so = em.compile_plugin(src, f"{mt}_init")
    assert os.path.exists(so)
    kinds = b["param_kinds"]
    shared = [i for i, k in enumerate(kinds) if k not in ("voltage", "prev_state", "new_state", "enable_lim", "enable_integration")]
    varying = [i for i, k in enumerate(kinds) if k == "voltage"]
    lim = {i: (k, b["param_names"][i]) for i, k in enumerate(kinds) if k in ("prev_state", "new_state", "enable_lim", "enable_integration")}
    nc = len(b["init_mir_data"]["cache_mapping"])
    layout = em.compute_eval_layout(b, shared, varying, list(range(nc)), [], lim)
    for use_lim in (False, True):
        src, meta = em.emit_eval_source(b, layout, use_lim)
        assert "vjx_eval_kernel" in src
        assert meta["n_residuals"] == len(b["dae_data"]["residuals"]) and meta["n_jacobian"] == len(b["dae_data"]["jacobian"])
        assert meta["flops_static"]["insts"] > 0
    dll = C.CDLL(em.compile_plugin(src, f"{mt}_eval"))
    assert dll.vjx_model_abi_version() == 1


def test_emitter_rejects_loops(bundles):
    import copy

    from vajax_b200.mir.types import parse_function

    b = copy.deepcopy(bundles["diode"])
    mir = b["mir_data"]
    # turn the first jmp into a back edge to block0
    for ins in mir["instructions"]:
        if ins["opcode"] == "jmp":
            ins["destination"] = "block0"
            mir["blocks"]["block0"]["predecessors"].append(ins["block"])
            mir["blocks"][ins["block"]]["successors"] = ["block0"]
            break
    with pytest.raises(Exception):
        parse_function("eval", mir, b["str_constants"]).topo_order()


# ---- corners (reference: tests/test_corners.py:18-143; values corners.py:162-196) ------------
def test_corner_tables_and_sweep():
    from vajax_b200.analysis import corners as cz

    ff, ss = cz.PROCESS_CORNERS["FF"], cz.PROCESS_CORNERS["SS"]
    assert (ff.mobility_scale, ff.vth_shift, ff.tox_scale) == (1.15, -0.05, 0.95)
    assert (ss.mobility_scale, ss.vth_shift, ss.tox_scale) == (0.85, 0.05, 1.05)
    assert cz.TEMPERATURE_CORNERS == {"cold": 233.15, "room": 300.15, "hot": 398.15}
    pvt = cz.create_pvt_corners()
    assert len(pvt) == 27 and len({c.name for c in pvt}) == 27
    assert pvt[0].name == "FF_m40C_V90pct"
    sw = cz.sweep_corners(64)
    assert len(sw) == 64 and len({c.name for c in sw}) == 64
    mc = cz.monte_carlo_corners(8)
    assert len(mc) == 8 and mc[0].process.mobility_scale != mc[1].process.mobility_scale
    with pytest.raises(ValueError):
        cz.create_standard_corners(["XX"])


def test_apply_corner_is_pure_and_signed():
    from vajax_b200.analysis import corners as cz
    from vajax_b200.benchmarks.circuits import multiplier_circuit

    circ = multiplier_circuit(2)
    mos = next(d for d in circ["devices"] if d["model"] == "psp103")
    vdd = next(d for d in circ["devices"] if d["name"] == "vdd")
    kp0, vto0, dc0 = mos["params"]["kp"], mos["params"]["vto"], vdd["params"]["dc"]
    sw = cz.sweep_corners(64)
    assert sw[0].name == "TT_27C_V100pct"  # nominal first
    ff = next(c for c in sw if c.name == "FF_m40C_V90pct")
    c = cz.apply_corner(circ, ff)  # FF, 233.15 K, 0.9 vdd
    m2 = next(d for d in c["devices"] if d["name"] == mos["name"])
    v2 = next(d for d in c["devices"] if d["name"] == "vdd")
    assert m2["params"]["kp"] == kp0 * 1.15 and m2["params"]["vto"] == vto0 - 0.05
    assert v2["params"]["dc"] == dc0 * 0.9 and c["temperature_k"] == 233.15
    assert mos["params"]["kp"] == kp0 and vdd["params"]["dc"] == dc0  # input untouched
    assert cz.shard(10, 1, 4) == [1, 4] and cz.shard(10, 0, 4) == [0, 7] and cz.shard(10, 2, 4) == [2, 5, 8]
    parts = [cz.shard(64, r, 8) for r in range(8)]
    assert sorted(i for p in parts for i in p) == list(range(64)) and all(len(p) == 8 for p in parts)
    # Latin square: a rank sees every position-in-block (vdd x temperature parity) exactly once
    assert all(sorted(i % 8 for i in p) == list(range(8)) for p in parts)
    assert cz.shard(7, 0, 1) == list(range(7))


def test_ffi_module_needs_jax_and_says_so():
    """The JAX FFI registration (vajax_b200/ffi.py) is importable without JAX and fails loudly when used."""
    from vajax_b200 import ffi
    from vajax_b200 import lib as vlib

    try:
        import jax  # noqa: F401
        pytest.skip("JAX is installed here: the FFI path is exercised elsewhere")
    except ImportError:
        pass
    with pytest.raises(vlib.VjxError):
        ffi.register()
    src = open(os.path.join(ROOT, "vajax_b200", "csrc", "vjx_ffi.cc")).read()
    for name in ffi.TARGETS:   # the handler symbols ffi.register() asks for exist in the source
        assert f"XLA_FFI_DEFINE_HANDLER_SYMBOL({name}," in src


def test_handle_id_registry_rejects_unknown_ids(lib):
    lib.vjx_handle_from_id.restype = C.c_void_p
    lib.vjx_handle_from_id.argtypes = [C.c_longlong]
    lib.vjx_handle_id.restype = C.c_longlong
    lib.vjx_handle_id.argtypes = [C.c_void_p]
    assert lib.vjx_handle_from_id(12345) is None
    assert lib.vjx_handle_id(None) == 0


def test_ffi_handlers_type_check_against_their_bindings(tmp_path):
    """vajax_b200/csrc/vjx_ffi.cc (the XLA FFI custom-call handlers, north_star's primary call path) cannot be built for
    real here - jaxlib's headers are absent - but it is compiled against tests/stubs/xla/ffi/api/ffi.h, a stand-in of
    the binding DSL that static_asserts every handler is invocable with exactly the context / attribute / operand /
    result types of its Ffi::Bind() chain, and against include/vajax_b200.h (so the C-ABI calls inside type-check too).
    Catches signature rot; runs nothing."""
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = os.path.join(root, "vajax_b200", "csrc", "vjx_ffi.cc")
    obj = str(tmp_path / "vjx_ffi.o")
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Werror=return-type", "-I", os.path.join(root, "tests", "stubs"),
                        "-I", "/usr/local/cuda/include", src], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run(["g++", "-std=c++17", "-c", "-fPIC", "-I", os.path.join(root, "tests", "stubs"), "-I", "/usr/local/cuda/include",
                        src, "-o", obj], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    syms = subprocess.run(["nm", "-g", obj], capture_output=True, text=True).stdout
    for target, arity in (("vjx_ffi_solve_f64", 5), ("vjx_ffi_build_system_f64", 21), ("vjx_ffi_nr_solve_f64", 24)):
        assert f"{target}_arity" in syms, target
    for fn in ("vjx_lu_solve", "vjx_build_system", "vjx_nr_solve_async", "vjx_handle_from_id", "vjx_sizes"):
        assert f"U {fn}" in syms, fn      # resolved against libvajax_b200.so at link time


@pytest.mark.skipif(not os.path.isdir("/root/reference/vajax"), reason="needs the reference tree (build container only)")
def test_reference_adapter_reproduces_the_circuit_description():
    """vajax_b200.analysis.reference_adapter on the REFERENCE's own prepared objects (CircuitEngine.prepare(),
    FullMNAStrategy.ensure_setup(), _bcsr_indptr; executed on the NumPy JAX shim): the arrays that go into
    vjx_circuit_desc equal the reference's (voltage gather indices, stamp indices, CSR pattern: checked inside the
    adapter) and equal what this repo derives on its own from the flat circuit.  Host only, in a subprocess because
    the shim shadows `jax` in sys.modules."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PYTHONPATH=os.path.join(root, "tests"))
    r = subprocess.run([sys.executable, "check_reference_adapter.py", "rc", "clipper", "ring", "mul3"],
                       cwd=os.path.join(root, "tests", "golden"), env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [json.loads(ln) for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 8 and all(ln["arrays_equal"] >= 5 for ln in lines), lines
    assert {ln["circuit"] for ln in lines} == {"rc", "clipper", "ring", "mul3"}


def test_init_keys_shared_per_parameter_set():
    """B200Engine._init_keys (set-up of a sweep corner): one tuple per distinct instance parameter set, equal to the
    per-device formula (instance value, else model default, else 0)."""
    import random

    from vajax_b200.analysis.engine import B200Engine

    rnd = random.Random(3)
    names = [f"p{i}" for i in range(40)]
    defaults = {f"p{i}": float(i) for i in range(0, 40, 2)}
    sets = [{"p1": 1e-6, "p3": 2e-7}, {"p1": 1e-6, "p3": 2e-7, "p4": 5.0}, {"p2": 3, "type": -1.0}, {}]
    devs = [{"name": f"m{i}", "params": dict(rnd.choice(sets))} for i in range(500)]
    devs[7]["params"]["p5"] = [1.0]          # unhashable value: handled per device, float() raises as before
    devs[7]["params"].pop("p5")
    want = [tuple(float(d["params"][n]) if n in d["params"] else float(defaults.get(n, 0.0)) for n in names) for d in devs]
    got = B200Engine._init_keys(devs, names, defaults)
    assert got == want
    assert len({id(k) for k in got}) <= len(sets)      # shared objects, one per distinct set
    mixed = [{"params": {"a": 1.0, 3: 2.0}}]          # unorderable keys: falls back to the per-device path
    assert B200Engine._init_keys(mixed, ["a"], {}) == [(1.0,)]


def test_apply_corner_leaves_the_input_untouched_and_equals_a_deep_copy():
    import copy

    from vajax_b200.analysis import corners as cz
    from vajax_b200.benchmarks import circuits

    circ = circuits.multiplier_circuit(3, mos_model="psp103w")
    src = circuits.source_mix_circuit(stop=6e-8)          # pwl sources: list-valued parameters
    for c0 in (circ, src):
        before = copy.deepcopy(c0)
        for corner in cz.sweep_corners(64)[:5] + cz.monte_carlo_corners(2):
            got = cz.apply_corner(c0, corner)
            want = dict(c0)
            want["devices"] = copy.deepcopy(c0["devices"])
            cz.apply_process_corner(want["devices"], corner.process)
            cz.apply_voltage_corner(want["devices"], corner.voltage)
            assert got["devices"] == want["devices"]
            assert c0["devices"] == before["devices"]                      # input untouched
            for a, b in zip(got["devices"], c0["devices"]):               # nothing mutable is shared
                assert a is not b and a["params"] is not b["params"] and a["nodes"] is not b["nodes"]
                for k, v in a["params"].items():
                    if isinstance(v, (list, dict)):
                        assert v is not b["params"][k]


def test_emitted_source_cache_is_exact(bundles, monkeypatch):
    """engine._emit_eval_cached returns what emit_eval_source returns, per key: arguments, layout and the VJX_EVAL_*
    switches all take part in the key; the cached meta is handed out as a copy."""
    from vajax_b200.analysis import engine as eng
    from vajax_b200.codegen import cuda_emitter as em

    b = bundles["psp103"]
    kinds = b["param_kinds"]
    shared = [i for i, k in enumerate(kinds) if k not in ("voltage", "prev_state", "new_state", "enable_lim", "enable_integration")]
    varying = [i for i, k in enumerate(kinds) if k == "voltage"]
    nc = len(b["init_mir_data"]["cache_mapping"])
    lay1 = em.compute_eval_layout(b, shared, varying, list(range(nc)), [], {})
    lay2 = em.compute_eval_layout(b, shared[1:], varying + shared[:1], list(range(nc)), [], {})
    monkeypatch.delenv("VJX_EVAL_BLOCK", raising=False)
    for lay in (lay1, lay2):
        for ns, mb in ((1, 0), (3, 0), (1, 5)):
            want = em.emit_eval_source(b, lay, False, n_slices=ns, min_blocks=mb)
            for _ in range(2):
                got = eng._emit_eval_cached(b, lay, False, ns, mb)
                assert got[0] == want[0] and got[1] == want[1]
            got[1]["n_slices"] = -1                              # a caller's edit must not leak into the cache
            assert eng._emit_eval_cached(b, lay, False, ns, mb)[1] == want[1]
    a = eng._emit_eval_cached(b, lay1, False, 1, 0)[0]
    monkeypatch.setenv("VJX_EVAL_BLOCK", "64")
    c = eng._emit_eval_cached(b, lay1, False, 1, 0)[0]
    assert a != c and c == em.emit_eval_source(b, lay1, False, n_slices=1, min_blocks=0)[0]
