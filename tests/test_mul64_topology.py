"""mul64 topology parity: the flattened device list of ``mul64_circuit()`` (names, terminals, node
numbering, order) must equal what the reference's own parser + flattening produce from
vajax/benchmarks/data/mul64/runme.sim - compared through the digest recorded by
tests/golden/make_mul64_topology.py (which ran the reference's front end in the build container)."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))


def _kind(d):
    if d["model"] == "psp103":
        return "nmos" if d["params"]["type"] > 0 else "pmos"
    return d["model"]


def test_mul64_matches_reference_flattening():
    from make_mul64_topology import topology_digest

    from vajax_b200.benchmarks.circuits import mul64_circuit

    gold = json.load(open(os.path.join(HERE, "golden", "mul64_topology.json")))
    c = mul64_circuit()
    got = topology_digest(((d["name"], _kind(d), d["nodes"]) for d in c["devices"]), c["node_names"])
    for k in ("n_devices", "n_nodes", "devices_sha256", "nodes_sha256"):
        assert got[k] == gold[k], k
    # the counts the reference documents for mul64 (tests/test_vacask_benchmarks.py:678-713)
    n_mos = sum(1 for d in c["devices"] if d["model"] == "psp103")
    assert n_mos == 266408 and c["num_nodes"] == 133463
    assert sum(1 for d in c["devices"] if d["model"] == "vsource") == 130
    # instance geometry of the nmos/pmos wrappers (models.inc:3-11): ad = w*ld, pd = 2*(w+ld), ld = 0.5u
    m = next(d for d in c["devices"] if d["name"].endswith("pp_0_0.mn1.m"))
    assert m["params"]["w"] == 0.5e-6 and m["params"]["l"] == 0.2e-6
    assert m["params"]["ad"] == 0.5e-6 * 0.5e-6 and m["params"]["pd"] == 2 * (0.5e-6 + 0.5e-6)


def test_mul64_small_instance_counts():
    from vajax_b200.benchmarks.circuits import mul64_circuit

    for n in (4, 8):
        c = mul64_circuit(n)
        n_mos = sum(1 for d in c["devices"] if d["model"] == "psp103")
        # n^2 AND (6 T) + buffer (4 T) + (n-1) HA (28 T) + (n-2) rows of [HA + (n-2) FA (62 T)]
        assert n_mos == 6 * n * n + 4 + 28 * (n - 1) + (n - 2) * (28 + 62 * (n - 2))
