"""World-size-2 test of the corner-sweep sharding and gather on CPU: two gloo
processes, the oracle engine standing in for the GPU engine (tests may use the oracle), results
compared with a single-process sweep."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _OracleAsEngine:
    """Adapter: OracleEngine behind the B200Engine constructor / run_transient signature."""

    def __init__(self, circuit, bundles, device=0, **_kw):
        from oracle.engine import OracleEngine

        self.o = OracleEngine(circuit, bundles)

    def run_transient(self, max_steps=None):
        t, V, st = self.o.run_transient(max_steps=max_steps)
        st = dict(st)
        st.setdefault("error_flags", 0)
        return t, V, st


def _corners():
    from vajax_b200.analysis import corners as cz

    return cz.sweep_corners(64)[:5]  # 5 corners over 2 ranks: uneven shards 3 + 2


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist

    from vajax_b200.analysis import corners as cz
    from vajax_b200.benchmarks.circuits import diode_clipper_circuit
    from vajax_b200.models.fixtures import get_bundle

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        bundles = {k: get_bundle(k) for k in ("resistor", "capacitor", "diode", "psp103")}
        res = cz.run_corners(diode_clipper_circuit(stop=1e-4), bundles, _corners(), rank=rank, world_size=world,
                             engine_factory=_OracleAsEngine, gather=cz.make_dist_gather(0))
        if rank == 0:
            q.put([(r.corner.name, r.times, r.voltages, r.stats["n_steps"]) for r in res.results])
        else:
            assert res.results == []
            q.put(None)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_corner_sweep_matches_single_process(bundles):
    import torch.multiprocessing as mp

    from vajax_b200.analysis import corners as cz
    from vajax_b200.benchmarks.circuits import diode_clipper_circuit

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sharded = next(g for g in got if g is not None)
    single = cz.run_corners(diode_clipper_circuit(stop=1e-4), bundles, _corners(), engine_factory=_OracleAsEngine)
    assert [s[0] for s in sharded] == [r.corner.name for r in single.results]
    assert single.all_converged and single.num_corners == 5
    for (name, t, V, n), r in zip(sharded, single.results):
        assert n == r.stats["n_steps"] and np.array_equal(t, r.times) and np.array_equal(V, r.voltages), name
    # the corners really differ (temperature / vdd change the waveforms)
    assert not np.array_equal(single.results[0].voltages, single.results[4].voltages)


def test_concurrent_corners_match_sequential(bundles):
    """run_corners(concurrent=k) only changes how a rank schedules its corners (threads + streams)."""
    from vajax_b200.analysis import corners as cz
    from vajax_b200.benchmarks.circuits import diode_clipper_circuit

    circ = diode_clipper_circuit(stop=1e-4)
    a = cz.run_corners(circ, bundles, _corners(), engine_factory=_OracleAsEngine)
    b = cz.run_corners(circ, bundles, _corners(), engine_factory=_OracleAsEngine, concurrent=3)
    assert [r.corner.name for r in a.results] == [r.corner.name for r in b.results]
    for ra, rb in zip(a.results, b.results):
        assert np.array_equal(ra.times, rb.times) and np.array_equal(ra.voltages, rb.voltages)
