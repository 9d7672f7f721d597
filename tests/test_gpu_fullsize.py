"""Size-independent properties at BASELINE.json's c6288 scale (the oracle is too slow to rerun a
whole transient there): linear-solve residuals, run-to-run bit-identity of the stamped system,
host-polled vs device-graph loops, chunked runs, logic-level sanity of the multiplier outputs."""
import numpy as np
import pytest

from vajax_b200.benchmarks import circuits

pytestmark = pytest.mark.gpu


MODELS = ["psp103", "psp103w"]   # the light round-1 surrogate and the PSP103-weight model (models/psp103w.py)


@pytest.fixture(scope="module", params=MODELS)
def c6288(request, bundles):
    from vajax_b200.analysis.engine import B200Engine

    e = B200Engine(circuits.multiplier_circuit(16, mos_model=request.param), bundles, use_device_graph=True).prepare()
    e.mos_model = request.param
    yield e
    e.close()


def test_c6288_sizes(c6288):
    e = c6288
    n_mos = sum(1 for d in e.devices if d["model"] == e.mos_model)
    # psp103w stamps PSP103's full pattern (NOI row -> D G S B, NOI column -> G S D): 127,670 slots, of which the
    # factorisation sees 64,998 (the NOI-enforced slots are structural zeros); the light surrogate has 64,998
    assert e.nse == (127670 if e.mos_model == "psp103w" else 64998)
    assert n_mos == 10048 and e.n_vs == 34
    assert e.n_total == 5091 + 2 * n_mos  # two surviving internal unknowns per MOSFET
    assert e.n_total < 30000  # tests/test_vacask_benchmarks.py:340-347


def test_c6288_build_is_deterministic_and_lu_solves(c6288):
    import scipy.sparse as sp

    e = c6288
    rng = np.random.default_rng(1)
    X = np.zeros(e.n_total + e.n_vs)
    X[1:e.n_total] = rng.uniform(0, 1.2, e.n_unknowns)
    X[e.noi_indices] = 0.0
    vs, isv = e.eval_sources(5e-11)
    nu = e.n_unknowns
    a = e.build_system(X, vs, isv, np.zeros(nu), 1e12, 1e-12, 0.0, -1e12, -1.0, None, 0.0, None, None, 1)
    b = e.build_system(X, vs, isv, np.zeros(nu), 1e12, 1e-12, 0.0, -1e12, -1.0, None, 0.0, None, None, 1)
    for u, v in zip(a, b):
        assert np.array_equal(u, v)  # bit-identical run to run
    csr = a[0].copy()
    # NOI rows/cols as the NR loop enforces them (solver_factories.py:1268-1276)
    rows = np.repeat(np.arange(e.n_aug), np.diff(e.csr_indptr))
    noi = np.zeros(e.n_aug, bool)
    noi[e.noi_indices - 1] = True
    csr[noi[rows] | ((rows < nu) & noi[e.csr_indices])] = 0.0
    csr[noi[rows] & (rows == e.csr_indices)] = 1.0
    A = sp.csr_matrix((csr, e.csr_indices, e.csr_indptr), shape=(e.n_aug, e.n_aug))
    for k in range(2):
        rhs = rng.normal(size=e.n_aug)
        x = e.lu_solve(csr, rhs)
        r = A @ x - rhs
        assert np.max(np.abs(r)) <= 1e-9 * max(1.0, np.max(np.abs(rhs)), np.max(np.abs(A @ np.abs(x))))
        assert np.array_equal(x, e.lu_solve(csr, rhs))  # deterministic
    st = e.lu_stats()
    assert st["n"] == e.n_aug and st["n_singletons"] + st["n_core"] == e.n_aug
    assert st["n_singletons"] >= 2 * 10048  # NOI identity rows + branch flows + rails leave the core


def test_c6288_graph_equals_host_loop_and_chunks(c6288, bundles):
    from vajax_b200.analysis.engine import B200Engine

    e = c6288
    t1, V1, s1 = e.run_transient(max_steps=41)
    assert s1["error_flags"] == 0 and s1["n_steps"] == 41
    # chunked: same loop stopped and resumed every 7 attempts (checkpoint contract)
    t2, V2, s2 = e.run_transient(max_steps=41, chunk_attempts=7)
    assert s2["n_steps"] == 41 and np.array_equal(t1, t2) and np.array_equal(V1, V2)
    h = B200Engine(circuits.multiplier_circuit(16, mos_model=e.mos_model), bundles, use_device_graph=False).prepare()
    t3, V3, s3 = h.run_transient(max_steps=41)
    h.close()
    assert s3["total_nr_iterations"] == s1["total_nr_iterations"] and s3["rejected_steps"] == s1["rejected_steps"]
    assert np.array_equal(t1, t3) and np.array_equal(V1, V3)
    # physics sanity: rails hold, every recorded voltage is bounded by the rails (+ overshoot)
    vdd = e.node_names["vdd"]
    assert np.allclose(V1[1:41, vdd], 1.2, atol=1e-9)  # row 0 is the UIC state (all zeros)
    assert np.all(np.isfinite(V1[:41])) and V1[:41].min() > -0.6 and V1[:41].max() < 1.8


@pytest.mark.parametrize("graph", [False, True])
def test_save_list_filters_the_output_rows_on_the_device(bundles, graph):
    """N4: with a save list V_out holds one column per listed node, bit-identical to the same columns of the full
    record (reference: full record full_mna.py:1017-1023,2122; `save` lines runme.sim:296-308); None restores it."""
    from vajax_b200.analysis.engine import B200Engine

    circ = circuits.multiplier_circuit(4, stop=3e-10, step=2e-12, mos_model="psp103w")
    e = B200Engine(circ, bundles, use_device_graph=graph).prepare()
    t0, V0, s0 = e.run_transient()
    n = s0["n_steps"]
    assert V0.shape[1] == e.n_external
    names = [f"p{k}" for k in range(4, 7)] + ["vdd", "a0", "s2_1", "0"]
    e.set_save_list(names)
    t1, V1, s1 = e.run_transient()
    assert s1["n_steps"] == n and np.array_equal(t0, t1) and V1.shape[1] == len(names)
    for k, nm in enumerate(names):
        assert s1["node_indices"].get(nm, k) == k or nm == "0"
        assert np.array_equal(V1[:n, k], V0[:n, e.node_names[nm]]), nm
    assert e._tran["d2h_bytes"] < 0.2 * 8 * V0.shape[0] * V0.shape[1]
    e.set_save_list(None)
    t2, V2, s2 = e.run_transient()
    assert np.array_equal(V2, V0) and np.array_equal(t2, t0)
    e.close()


@pytest.mark.parametrize("model", MODELS)
def test_mul4_settles_to_the_product(bundles, model):
    """4x4 multiplier, operands 15 x 15 = 225: after the inputs have risen and the array has
    settled, the product bits read 0b11100001."""
    from vajax_b200.analysis.engine import B200Engine

    circ = circuits.multiplier_circuit(4, stop=4e-9, step=2e-12, mos_model=model)
    e = B200Engine(circ, bundles, use_device_graph=True).prepare()
    times, V, st = e.run_transient()
    n = st["n_steps"]
    assert st["error_flags"] == 0 and times[n - 1] >= 4e-9 * (1 - 1e-9)
    names = e.node_names
    bits = []
    bits.append(V[n - 1, names["pp0_0"]])
    bits.append(V[n - 1, names["s1_0"]])
    bits.append(V[n - 1, names["s2_0"]])
    bits.append(V[n - 1, names["s3_0"]])
    for k in range(3):
        bits.append(V[n - 1, names[f"p{4 + k}"]])
    bits.append(V[n - 1, names["rc2"]])
    got = sum((1 << i) for i, v in enumerate(bits) if v > 0.6)
    assert all(v < 0.2 or v > 1.0 for v in bits), bits
    assert got == 225, (got, bits)
    e.close()


def test_concurrent_corner_sweep_is_bit_identical(bundles):
    """Whole-circuit replicas on separate streams / host threads of ONE GPU (run_corners(concurrent=4)):
    same waveforms, bit for bit, as the sequential sweep - the LU's ticket / sentinel protocol does not
    depend on which other kernels share the SMs."""
    from vajax_b200.analysis import corners as cz

    circ = circuits.multiplier_circuit(4, stop=4e-10, step=2e-12)
    cs = cz.sweep_corners(64)[:6]
    a = cz.run_corners(circ, bundles, cs, use_device_graph=True)
    b = cz.run_corners(circ, bundles, cs, use_device_graph=True, concurrent=4)
    assert a.all_converged and b.all_converged
    for ra, rb in zip(a.results, b.results):
        assert ra.corner.name == rb.corner.name
        assert np.array_equal(ra.times, rb.times) and np.array_equal(ra.voltages, rb.voltages)


def test_live_engines_keep_their_own_model_constants(bundles):
    """Two circuit handles alive at the same time (corners with different temperature / process values)
    must not share the model plugin's __constant__ bank: the waveforms of a corner do not depend on
    which other engines exist (regression: constants were uploaded at create into one shared bank)."""
    from vajax_b200.analysis import corners as cz
    from vajax_b200.analysis.engine import B200Engine

    circ = circuits.multiplier_circuit(4, stop=2e-10, step=2e-12)
    cs = cz.sweep_corners(64)
    ca, cb = cz.apply_corner(circ, cs[5]), cz.apply_corner(circ, cs[40])
    alone = B200Engine(ca, bundles, use_device_graph=True)
    t0, V0, s0 = alone.run_transient()
    alone.close()
    a = B200Engine(ca, bundles, use_device_graph=True).prepare()
    b = B200Engine(cb, bundles, use_device_graph=True).prepare()   # created later: used to overwrite a's constants
    t1, V1, s1 = a.run_transient()
    t2, V2, s2 = b.run_transient()
    a.close()
    b.close()
    assert s1["n_steps"] == s0["n_steps"] and np.array_equal(t0, t1) and np.array_equal(V0, V1)
    assert not np.array_equal(V1, V2)   # the two corners really differ


def test_c6288_first_steps_match_the_oracle(c6288, bundles):
    """BASELINE.json's full c6288 size against the CPU oracle on the first accepted steps (the oracle
    needs ~0.15 s per step here): same accepted-step sequence (step, nr_iters, accept/reject), t and dt to
    the accuracy of the solves (1e-9 relative, north_star), waveforms within 1e-6 RMS relative."""
    from oracle.engine import OracleEngine

    K = 9 if c6288.mos_model == "psp103" else 6     # the oracle interprets 6.4 k MIR instructions per psp103w evaluation
    o = OracleEngine(circuits.multiplier_circuit(16, mos_model=c6288.mos_model), bundles)
    rec = []
    to, Vo, so = o.run_transient(max_steps=K, record=rec)
    tg, Vg, sg = c6288.run_transient(max_steps=K)
    assert sg["error_flags"] == 0 and sg["n_steps"] == so["n_steps"] == K
    assert sg["total_nr_iterations"] == so["total_nr_iterations"] and sg["rejected_steps"] == so["rejected_steps"]
    gr = c6288.step_records()
    assert len(gr) == len(rec)
    for a, b in zip(gr, rec):
        assert int(a[0]) == b["step"] and int(a[3]) == b["nr_iters"] and bool(a[4]) == b["accept"]
        assert abs(a[1] - b["t_next"]) <= 1e-9 * abs(b["t_next"]) and abs(a[2] - b["dt"]) <= 1e-9 * b["dt"]
    span = max(np.max(Vo[:K]) - np.min(Vo[:K]), 1e-12)
    rms = np.sqrt(np.mean((Vg[:K] - Vo[:K]) ** 2)) / span
    assert rms < 1e-6, rms
    assert np.max(np.abs(Vg[:K] - Vo[:K])) < 1e-7


@pytest.mark.timeout(900)
@pytest.mark.parametrize("model", MODELS)
def test_mul64_scale_properties(bundles, model):
    """BASELINE.json's largest configuration (mul64: 266,408 MOSFETs, 666,408 unknowns): the analysis sizes
    the reference documents, a few transient steps without device error flags, and the size-independent
    linear-solve properties (residual of A x = b on the current Jacobian, run-to-run bit identity)."""
    import scipy.sparse as sp

    from vajax_b200.analysis.engine import B200Engine

    e = B200Engine(circuits.mul64_circuit(64, mos_model=model), bundles, use_device_graph=True).prepare()
    assert e.n_aug == 666408 and sum(1 for d in e.devices if d["model"] == model) == 266408
    if model == "psp103w":   # the reference reports 3,259,918 CSR entries for mul64 (docs/performance_analysis.md:184)
        assert e.nse == 3260938 and abs(e.nse - 3259918) / 3259918 < 1e-3
    t, V, st = e.run_transient(max_steps=4, read_outputs=False)
    assert st["error_flags"] == 0 and st["n_steps"] == 4
    lu = e.lu_stats()
    assert lu["n"] == e.n_aug and lu["n_singletons"] + lu["n_core"] == e.n_aug
    sc = e.tran_scalars()
    X, Q, dQ = e.tran_state()
    vs, isv = e.eval_sources(sc["t"])
    nu = e.n_unknowns
    c0 = 2.0 / sc["dt"]
    csr = e.build_system(X, vs, isv, Q, c0, 1e-12, 0.0, -c0, -1.0, dQ, 0.0, None, None, 1)[0].copy()
    rows = np.repeat(np.arange(e.n_aug), np.diff(e.csr_indptr))
    noi = np.zeros(e.n_aug, bool)
    noi[e.noi_indices - 1] = True
    csr[noi[rows] | ((rows < nu) & noi[e.csr_indices])] = 0.0
    csr[noi[rows] & (rows == e.csr_indices)] = 1.0
    A = sp.csr_matrix((csr, e.csr_indices, e.csr_indptr), shape=(e.n_aug, e.n_aug))
    rhs = np.random.default_rng(3).normal(size=e.n_aug)
    x = e.lu_solve(csr, rhs)
    r = A @ x - rhs
    assert np.max(np.abs(r)) <= 1e-9 * max(1.0, np.max(np.abs(A @ np.abs(x))))
    assert np.array_equal(x, e.lu_solve(csr, rhs))
    e.close()


@pytest.mark.timeout(900)
@pytest.mark.parametrize("model", MODELS)
def test_mul64_first_steps_match_the_oracle_golden(bundles, model):
    """BASELINE.json's largest configuration against the CPU oracle: the oracle needs ~30 s per NR iteration
    at this size, so its first two accepted steps are a committed golden vector
    (tests/golden/mul64_first_steps.npz, generated by tests/golden/make_mul64_steps.py).  Same accepted-step
    sequence, t / dt to the accuracy of the solves, waveforms within 1e-6 RMS relative on the stored columns,
    and the row checksums over ALL 133,463 external nodes."""
    import os

    from vajax_b200.analysis.engine import B200Engine

    fn = "mul64_first_steps.npz" if model == "psp103" else f"mul64_first_steps_{model}.npz"
    path = os.path.join(os.path.dirname(__file__), "golden", fn)
    if not os.path.exists(path):
        pytest.skip(f"{fn} not generated (tests/golden/make_mul64_steps.py {model})")
    g = np.load(path)
    n = int(g["n_steps"])
    e = B200Engine(circuits.mul64_circuit(64, mos_model=model), bundles, use_device_graph=True).prepare()
    t, V, st = e.run_transient(max_steps=n)
    assert st["error_flags"] == 0 and st["n_steps"] == n and V.shape[1] == int(g["n_external"])
    assert st["total_nr_iterations"] == int(g["total_nr"]) and st["rejected_steps"] == int(g["rejected"])
    rec = e.step_records()
    assert len(rec) == len(g["records"])
    for a, b in zip(rec, g["records"]):
        assert int(a[0]) == int(b[0]) and int(a[3]) == int(b[3]) and bool(a[4]) == bool(b[4])
        assert abs(a[1] - b[1]) <= 1e-9 * abs(b[1]) and abs(a[2] - b[2]) <= 1e-9 * b[2]
    assert np.max(np.abs(t[:n] - g["times"])) <= 1e-9 * g["times"][-1]
    Vc = V[:n][:, g["cols"]]
    span = max(np.max(g["V_cols"]) - np.min(g["V_cols"]), 1e-12)
    assert np.sqrt(np.mean((Vc - g["V_cols"]) ** 2)) / span < 1e-6
    assert np.max(np.abs(Vc - g["V_cols"])) < 1e-7
    assert np.allclose(np.max(np.abs(V[:n]), axis=1), g["row_absmax"], rtol=0, atol=1e-7)
    assert np.allclose(np.sum(V[:n], axis=1), g["row_sum"], rtol=1e-9, atol=1e-4)
    e.close()


@pytest.mark.timeout(900)
def test_c6288_long_window_matches_the_accelerated_oracle(c6288, bundles):
    """The window bench.py times (150 accepted steps at BASELINE.json's c6288 size, rejected attempts included)
    against the CPU oracle: identical attempt records (step, NR iterations, accept / reject), t to 1e-8 and dt to
    1e-7 relative, waveforms to 1e-6 RMS.  The oracle runs with its two accelerations for timing legs - model functions compiled
    for the host (oracle/host_compiled.py, held to the interpreter by tests/test_host_compiled.py) and the cached
    ordering of the sparse solve - because the interpreter needs ~1.2 s per step here; the interpreter itself
    anchors the first steps (test above) and the small circuits."""
    from oracle import host_compiled
    from oracle.engine import OracleEngine

    K = 150
    o = OracleEngine(circuits.multiplier_circuit(16, mos_model=c6288.mos_model), bundles)
    o.reuse_ordering = True
    rec = []
    with host_compiled.installed():
        to, Vo, so = o.run_transient(max_steps=K, record=rec)
    tg, Vg, sg = c6288.run_transient(max_steps=K)
    gr = c6288.step_records()
    first_bad = next((i for i, (a, b) in enumerate(zip(gr, rec))
                      if not (int(a[0]) == b["step"] and int(a[3]) == b["nr_iters"] and bool(a[4]) == b["accept"])), None)
    assert first_bad is None, (first_bad, list(gr[first_bad]), rec[first_bad])
    assert sg["error_flags"] == 0 and sg["n_steps"] == so["n_steps"] == K and len(gr) == len(rec)
    assert sg["total_nr_iterations"] == so["total_nr_iterations"] and sg["rejected_steps"] == so["rejected_steps"]
    # dt is dt_prev * (lte_ratio / norm)^(1/(order+1)) with norm built from V_new - V_pred, a difference of nearly equal
    # numbers: the 1e-9 agreement of two converged NR solves is amplified by |V| / |V_new - V_pred| and carried from step
    # to step, so over 150 steps the bar is 1e-7 on dt and 1e-8 on t (the first-steps test above holds 1e-9)
    e_t = max(abs(a[1] - b["t_next"]) / abs(b["t_next"]) for a, b in zip(gr, rec))
    e_dt = max(abs(a[2] - b["dt"]) / b["dt"] for a, b in zip(gr, rec))
    span = max(np.max(Vo[:K]) - np.min(Vo[:K]), 1e-12)
    rms = np.sqrt(np.mean((Vg[:K] - Vo[:K]) ** 2)) / span
    dv = np.max(np.abs(Vg[:K] - Vo[:K]))
    print(f"c6288/{c6288.mos_model}: {K} steps, {len(rec)} attempts, {so['rejected_steps']} rejected, "
          f"{so['total_nr_iterations']} NR iterations; max rel dt {e_dt:.2e}, t {e_t:.2e}; rms {rms:.2e}, max |dV| {dv:.2e}")
    assert e_t < 1e-8 and e_dt < 1e-7, (e_t, e_dt)     # measured on the B200: 5.5e-10 / 6.3e-9
    assert rms < 1e-6, rms                             # north_star bar; measured 6.3e-11
    assert dv < 1e-7, dv                               # measured 2.6e-9


@pytest.mark.timeout(900)
def test_mul64_15_step_window_matches_the_oracle_golden(bundles):
    """The window the reference times on mul64 (15 steps, README.md:80) with the PSP103-weight model against the
    CPU oracle: tests/golden/mul64_steps15_psp103w.npz (make_mul64_steps.py psp103w 16; the oracle with compiled host
    model functions and cached ordering, 5.5 min in the build container).  Identical attempt records including the
    rejected third attempt, t / dt / waveform bars as in the c6288 window test, row checksums over all 133,463
    external nodes."""
    import os

    from vajax_b200.analysis.engine import B200Engine

    path = os.path.join(os.path.dirname(__file__), "golden", "mul64_steps15_psp103w.npz")
    g = np.load(path)
    n = int(g["n_steps"])
    e = B200Engine(circuits.mul64_circuit(64, mos_model="psp103w"), bundles, use_device_graph=True).prepare()
    t, V, st = e.run_transient(max_steps=n)
    assert st["error_flags"] == 0 and st["n_steps"] == n and V.shape[1] == int(g["n_external"])
    rec = e.step_records()
    assert [(int(a[0]), int(a[3]), bool(a[4])) for a in rec] == [(int(b[0]), int(b[3]), bool(b[4])) for b in g["records"]]
    assert st["total_nr_iterations"] == int(g["total_nr"]) and st["rejected_steps"] == int(g["rejected"]) == 1
    e_t = max(abs(a[1] - b[1]) / abs(b[1]) for a, b in zip(rec, g["records"]))
    e_dt = max(abs(a[2] - b[2]) / b[2] for a, b in zip(rec, g["records"]))
    Vc = V[:n][:, g["cols"]]
    span = max(np.max(g["V_cols"]) - np.min(g["V_cols"]), 1e-12)
    rms = np.sqrt(np.mean((Vc - g["V_cols"]) ** 2)) / span
    dv = np.max(np.abs(Vc - g["V_cols"]))
    print(f"mul64/psp103w: {n - 1} steps, {len(rec)} attempts; max rel dt {e_dt:.2e}, t {e_t:.2e}; rms {rms:.2e}, max |dV| {dv:.2e}")
    assert e_t < 1e-8 and e_dt < 1e-7, (e_t, e_dt)
    assert rms < 1e-6 and dv < 1e-7, (rms, dv)
    assert np.allclose(np.max(np.abs(V[:n]), axis=1), g["row_absmax"], rtol=0, atol=1e-7)
    assert np.allclose(np.sum(V[:n], axis=1), g["row_sum"], rtol=1e-9, atol=1e-4)
    e.close()


@pytest.mark.timeout(900)
def test_c6288_corner_sweep_matches_the_oracle_per_corner(bundles):
    """BASELINE.json configs[4] (PVT sweep of whole-circuit replicas) at c6288 size: four corners that differ in process,
    temperature and supply, run 4-in-flight on one GPU (run_corners(concurrent=4)), against the CPU oracle run
    corner by corner through the SAME driver (run_corners with the oracle as engine factory): identical attempt
    counts and NR iteration counts per corner, waveforms to 1e-6 RMS - the parity target of the sweep is 'a fresh
    engine per corner' (SURVEY 8e caveat D10)."""
    from oracle import host_compiled
    from oracle.engine import OracleEngine
    from vajax_b200.analysis import corners as cz

    circ = circuits.multiplier_circuit(16, mos_model="psp103w")
    allc = cz.sweep_corners(64)
    cs = [allc[0], allc[21], allc[42], allc[63]]     # TT/27C/1.0, FF/-40C/0.9, SS/0C/1.05, FS/125C/1.1
    assert len({c.name for c in cs}) == 4
    K = 40
    g = cz.run_corners(circ, bundles, cs, use_device_graph=True, concurrent=4, max_steps=K)

    def oracle_factory(c, b, device=0, **kw):
        o = OracleEngine(c, b)
        o.reuse_ordering = True
        return o

    with host_compiled.installed():
        o = cz.run_corners(circ, bundles, cs, engine_factory=oracle_factory, max_steps=K)
    assert g.all_converged and len(g.results) == len(o.results) == 4
    seen = set()
    for rg, ro in zip(g.results, o.results):
        assert rg.corner.name == ro.corner.name
        n = int(ro.stats["n_steps"])
        assert int(rg.stats["n_steps"]) == n == K
        assert int(rg.stats["total_nr_iterations"]) == int(ro.stats["total_nr_iterations"]), rg.corner.name
        assert int(rg.stats["rejected_steps"]) == int(ro.stats["rejected_steps"]), rg.corner.name
        assert np.max(np.abs(rg.times - ro.times)) <= 1e-8 * ro.times[-1]
        span = max(np.max(ro.voltages) - np.min(ro.voltages), 1e-12)
        rms = np.sqrt(np.mean((rg.voltages - ro.voltages) ** 2)) / span
        dv = np.max(np.abs(rg.voltages - ro.voltages))
        print(f"{rg.corner.name}: {n} steps, {ro.stats['total_nr_iterations']} NR iterations, {ro.stats['rejected_steps']} rejected; "
              f"rms {rms:.2e}, max |dV| {dv:.2e}, vmax {np.max(ro.voltages):.3f}")
        assert rms < 1e-6 and dv < 1e-7, (rg.corner.name, rms, dv)
        seen.add((int(ro.stats["total_nr_iterations"]), round(float(np.max(ro.voltages)), 3)))
    assert len(seen) > 1      # the corners really are different circuits


def _slot_abs_sums(o, parts, c0):
    """Per CSR slot: sum of |contribution| and number of contributions (the quantities the rounding error of a
    summation order is bounded by)."""
    s, cnt = np.zeros(o.nse), np.zeros(o.nse)
    for mt, m in o.models.items():
        v = np.abs(parts[mt]["jac_resist"] + c0 * parts[mt]["jac_react"])
        pos = m["csr_pos"]
        ok = (pos >= 0) & ~np.isnan(v)
        np.add.at(s, pos[ok], v[ok])
        np.add.at(cnt, pos[ok], 1.0)
    return s, cnt


def _slot_exact(o, parts, c0, k):
    """Correctly rounded sum (math.fsum) of the model contributions into slot k."""
    import math

    terms = []
    for mt, m in o.models.items():
        v = parts[mt]["jac_resist"] + c0 * parts[mt]["jac_react"]
        sel = (m["csr_pos"] == k) & ~np.isnan(v)
        terms.extend(v[sel].tolist())
    if k in set(np.asarray(o.diag_pos).tolist()):
        terms.append(o.options.gmin + 1e-9)        # diagonal regularisation + the gshunt these calls pass
    return math.fsum(terms), len(terms)


def _strict_build_system_parity(e, o, seed, label, steps=6):
    """Stamped system of the CUDA path against the interpreter oracle, per-entry metric (golden_util).

    (1) At an operating point the transient visits (the state after ``steps`` accepted steps), three coefficient sets.
    (2) At a random, unphysical bias point (every node uniform in [0, 1.2] V).

    Bars: f, Q, Ivs, mrc 1e-10 per entry.  J per entry: |got - want| <= 1e-12 max(|want|, 1e-12) + n_k u S_k, where n_k
    contributions of absolute sum S_k meet in slot k and u = 2^-53: the second term is the a-priori bound of recursive
    summation (Higham, Accuracy and Stability of Numerical Algorithms, eq. 4.4), i.e. what the oracle's / the reference's
    own left-to-right sum may be off by.  It is 1e-15 S_k for an ordinary slot (n_k <= 10) and matters for two kinds of
    slot only: sums that cancel, and the supply rails' diagonals at mul64 size, where 266 k near-identical conductances
    add up - identical addends round the same way every time, so the left-to-right sum drifts by ~2e-12 relative while
    the CUDA path's chunked tree reduction stays next to the exact sum (printed below: both against math.fsum).
    At (2) the PSP103-weight model also amplifies last-bit differences between elementary-function libraries to
    ~1e-12: the oracle itself moves by that much when its exp / log / pow come from glibc (oracle/host_compiled.py)
    instead of NumPy, so that self-difference is measured on the spot and the relative part of the bar is
    max(1e-12, 10 x it).  The share of Jacobian entries inside plain 1e-12 is printed."""
    from golden_util import ABS_FLOOR, BAR, BS_NAMES, entry_err
    from oracle import host_compiled

    nu = e.n_unknowns
    rng = np.random.default_rng(seed)
    e.run_transient(max_steps=steps, read_outputs=False)
    Xp, Qp, dq = e.tran_state()
    vs, isv = e.eval_sources(5e-11)
    q2 = rng.normal(0, 1e-15, nu)
    Xr = np.zeros(e.n_total + e.n_vs)
    Xr[1:e.n_total] = rng.uniform(0, 1.2, nu)
    Xr[e.noi_indices] = 0.0
    Xr[e.n_total:] = rng.normal(0, 1e-4, e.n_vs)
    eps = np.finfo(np.float64).eps
    for where, X, cases in (("operating point", Xp, ((0.0, 0.0, 0.0, 0.0, 0), (2e12, -2e12, -1.0, 0.0, 1), (1.5e11, -2e11, 0.0, 0.5e11, 3))),
                            ("random bias point", Xr, ((2e12, -2e12, -1.0, 0.0, 1),))):
        rep = {}
        for (c0, c1, d1, c2, it) in cases:
            got = e.build_system(X, vs, isv, Qp, c0, 1e-12, 1e-9, c1, d1, dq, c2, q2, None, it)
            want = o.build_system(X, vs, isv, Qp, c0, 1e-12, 1e-9, c1, d1, dq, c2, q2, None, it, return_parts=True)
            with host_compiled.installed():
                alt = o.build_system(X, vs, isv, Qp, c0, 1e-12, 1e-9, c1, d1, dq, c2, q2, None, it)
            S, N = _slot_abs_sums(o, want[6], c0)
            for nm, a, b, c in zip(BS_NAMES, got, want, alt):
                a, b, c = np.asarray(a), np.asarray(b), np.asarray(c)
                if b.size == 0:
                    continue
                assert np.array_equal(np.isnan(a), np.isnan(b)), (label, where, nm)
                m = ~np.isnan(b)
                noise = entry_err(c[m], b[m], ABS_FLOOR[nm])
                rel = max(BAR.get(nm, 1e-10), 10 * noise)
                scale = np.maximum(np.abs(b[m]), ABS_FLOOR[nm])
                d = np.abs(a[m] - b[m])
                slack = N[m] * (eps / 2) * S[m] if nm == "J" else 0.0
                bad = d > rel * scale + slack
                k = int(np.argmax(d / scale))
                r = rep.setdefault(nm, {"err": 0.0, "noise": 0.0})
                r["err"], r["noise"] = max(r["err"], float(np.max(d / scale))), max(r["noise"], noise)
                if nm == "J":
                    r["within_1e-12"] = min(r.get("within_1e-12", 100.0), 100.0 * float(np.mean(d <= 1e-12 * scale)))
                    if float(d[k] / scale[k]) >= r.get("_worst", -1.0):
                        r["_worst"] = float(d[k] / scale[k])
                        kk = int(np.flatnonzero(m)[k])
                        exact, n_terms = _slot_exact(o, want[6], c0, kk)
                        r["worst_slot"] = (f"slot {kk}: {n_terms} contributions, |got - fsum| / |fsum| = {abs(a[kk] - exact) / max(abs(exact), 1e-300):.1e}, "
                                           f"|oracle - fsum| / |fsum| = {abs(b[kk] - exact) / max(abs(exact), 1e-300):.1e} (model contributions + diagonal regularisation)")
                assert not bad.any(), (label, where, nm, c0, it, f"entry {int(np.argmax(bad))}: got {a[m][np.argmax(bad)]!r} want "
                                       f"{b[m][np.argmax(bad)]!r}", f"noise {noise:.1e}")
        print(label, where, {k: {kk: (vv if isinstance(vv, str) else f"{vv:.2e}" if kk != "within_1e-12" else f"{vv:.5f} %")
                                 for kk, vv in v.items() if not kk.startswith("_")} for k, v in rep.items()})


def test_c6288_build_system_strict_parity_with_the_interpreter_oracle(c6288, bundles):
    """north_star's 'stamped Jacobians within 1e-12 relative' at the metric's own size: 127,670 CSR slots of 10,048
    MOSFETs against the NumPy MIR interpreter (the checker proper, no acceleration), three coefficient sets."""
    from oracle.engine import OracleEngine

    o = OracleEngine(circuits.multiplier_circuit(16, mos_model=c6288.mos_model), bundles)
    o.setup()
    assert np.array_equal(o.csr_indptr, c6288.csr_indptr) and np.array_equal(o.csr_indices, c6288.csr_indices)
    _strict_build_system_parity(c6288, o, 7, f"c6288/{c6288.mos_model}")


@pytest.mark.timeout(1200)
def test_mul64_build_system_strict_parity_with_the_interpreter_oracle(bundles):
    """The same at BASELINE.json's largest configuration: 3,260,938 CSR slots, 266,408 PSP103-weight MOSFETs."""
    from oracle.engine import OracleEngine
    from vajax_b200.analysis.engine import B200Engine

    circ = circuits.mul64_circuit(64, mos_model="psp103w")
    e = B200Engine(circ, bundles, use_device_graph=False).prepare()
    o = OracleEngine(circ, bundles)
    o.setup()
    assert e.nse == len(o.csr_indices) == 3260938
    assert np.array_equal(o.csr_indptr, e.csr_indptr) and np.array_equal(o.csr_indices, e.csr_indices)
    _strict_build_system_parity(e, o, 11, "mul64/psp103w")
    e.close()
