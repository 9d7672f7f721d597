"""Pin the mul64 topology: run the REFERENCE's own netlist front end on its own mul64 netlist and
record a digest of the flattened result (tests/golden/mul64_topology.json).

    python tests/golden/make_mul64_topology.py          # needs /root/reference (build container only)

The reference's parser (vajax/netlist/parser.py) and flattening (vajax/analysis/parsing.py) are pure
Python; `vajax/__init__.py` and `vajax/analysis/__init__.py` import jax, so both packages are
pre-seeded as empty namespaces, and `vajax.utils.safe_eval` (needs the absent `simpleeval`) is
replaced by a plain eval - parameter VALUES are not part of the digest, only names, models,
terminals and the node numbering.  The generate block of the netlist runs the in-tree va_builder.
"""
import hashlib
import json
import os
import sys
import types

REF = "/root/reference"


def topology_digest(entries, node_names):
    """entries: iterable of (device name, model kind, [node index...]) in netlist order."""
    h = hashlib.sha256()
    n = 0
    for name, kind, nodes in entries:
        h.update(f"{name} {kind} {' '.join(str(x) for x in nodes)}\n".encode())
        n += 1
    hn = hashlib.sha256()
    for nm, idx in sorted(node_names.items(), key=lambda kv: kv[1]):
        hn.update(f"{idx} {nm}\n".encode())
    return {"n_devices": n, "n_nodes": len(node_names), "devices_sha256": h.hexdigest(), "nodes_sha256": hn.hexdigest()}


def main():
    sys.path.insert(0, REF + "/va-builder")
    os.environ["PYTHONPATH"] = REF + "/va-builder" + os.pathsep + os.environ.get("PYTHONPATH", "")
    for name, sub in (("vajax", "vajax"), ("vajax.utils", "vajax/utils"), ("vajax.analysis", "vajax/analysis")):
        m = types.ModuleType(name)
        m.__path__ = [f"{REF}/{sub}"]
        sys.modules[name] = m
    se = types.ModuleType("vajax.utils.safe_eval")

    def safe_eval_expr(expr, env, default=0.0):
        try:
            return float(eval(expr, {"__builtins__": {}}, dict(env)))
        except Exception:
            return default

    se.safe_eval_expr = safe_eval_expr
    sys.modules["vajax.utils.safe_eval"] = se
    from vajax.analysis import parsing
    from vajax.netlist import parser as P

    path = REF + "/vajax/benchmarks/data/mul64/runme.sim"
    circ = P.VACASKParser().parse_file(path)
    flat = parsing.flatten_instances(circ, parsing.parse_elaborate_directive(open(path).read()), float)
    node_names, num_nodes = parsing.build_node_mapping(flat, circ.ground or "0")
    kind = {"psp103n": "nmos", "psp103p": "pmos", "v": "vsource", "r": "resistor"}
    out = topology_digest(((nm, kind[mdl], [node_names[t] for t in terms]) for nm, terms, mdl, _ in flat), node_names)
    out["source"] = "vajax/benchmarks/data/mul64/runme.sim via vajax.netlist.parser + vajax.analysis.parsing"
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "mul64_topology.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(out)


if __name__ == "__main__":
    main()
