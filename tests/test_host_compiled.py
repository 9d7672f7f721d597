"""The compiled host evaluation used by bench.py's CPU legs (oracle/host_compiled.py) against the NumPy MIR
interpreter (the parity checker): same build_system outputs on every fixture model, with and without limiting,
and the cached-ordering linear solve against the plain one.  Also a CPU-side check of the statement-level
MIR -> C translation that the CUDA emitter shares (cuda_emitter._Body)."""
import shutil

import numpy as np
import pytest

from golden_util import CIRCUITS
from oracle import host_compiled as hc
from oracle.engine import OracleEngine

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")

# relative bar: libm (compiled) and NumPy's vector loops (interpreter) differ in the last bits of exp / log / pow
BAR = {"rc": 0.0, "clipper": 1e-13, "opcov": 1e-12, "ring": 1e-12, "mul3": 1e-12, "ringw": 1e-11}


def _inputs(o, seed=5):
    rng = np.random.default_rng(seed)
    X = rng.normal(0.4, 0.5, o.n_total + o.n_vs)
    X[0] = 0.0
    ls = rng.normal(0.5, 0.2, o.total_limit_states) if o.total_limit_states else None
    return (X, rng.normal(1, 0.2, o.n_vs), rng.normal(0, 1e-4, len(o.isources)), rng.normal(0, 1e-13, o.n_unknowns)), ls


@pytest.mark.parametrize("name", list(BAR))
@pytest.mark.parametrize("limiting", [False, True])
def test_compiled_eval_equals_interpreter(name, limiting, bundles):
    o = OracleEngine(CIRCUITS[name](), bundles, use_device_limiting=limiting)
    o.setup()
    args, ls = _inputs(o)
    for c0, it in ((0.0, 0), (2e11, 1)):
        ref = o.build_system(*args, c0, limit_state_in=ls, nr_iteration=it)
        with hc.installed(threads=2):
            got = o.build_system(*args, c0, limit_state_in=ls, nr_iteration=it)
        for nm, a, b in zip(["J", "f", "Q", "Ivs", "lim", "mrc"], ref, got):
            a, b = np.asarray(a), np.asarray(b)
            assert np.array_equal(np.isnan(a), np.isnan(b)), nm
            ok = np.isfinite(a)
            err = np.max(np.abs(a[ok] - b[ok]) / np.maximum(np.abs(a[ok]), 1e-300), initial=0.0)
            assert err <= BAR[name], (name, limiting, nm, err)


def test_thread_count_does_not_change_results(bundles):
    o = OracleEngine(CIRCUITS["mul3"](), bundles)
    o.setup()
    args, _ = _inputs(o)
    outs = []
    for th in (1, 3, 8):
        with hc.installed(threads=th):
            outs.append(o.build_system(*args, 1e10))
    for other in outs[1:]:
        for a, b in zip(outs[0], other):
            assert np.array_equal(a, b)


def test_patch_is_scoped(bundles):
    from oracle import mir_interp

    orig = mir_interp.eval_model
    with hc.installed():
        assert mir_interp.eval_model is not orig
    assert mir_interp.eval_model is orig


def test_cached_ordering_solve(bundles):
    """reuse_ordering (timing legs only) solves the same systems: same NR iteration counts, waveforms to 1e-9."""
    circ = CIRCUITS["mul3"]()
    a = OracleEngine(circ, bundles)
    b = OracleEngine(circ, bundles)
    b.reuse_ordering = True
    a.setup()
    assert a.n_aug > 400   # above the oracle's dense-solve threshold: the sparse path is what is compared
    _, Va, sa = a.run_transient(max_steps=12)
    with hc.installed(threads=2):
        _, Vb, sb = b.run_transient(max_steps=12)
    n = sa["n_steps"]
    assert sb["n_steps"] == n and sb["total_nr_iterations"] == sa["total_nr_iterations"]
    assert float(np.max(np.abs(Va[:n] - Vb[:n]))) < 1e-9
