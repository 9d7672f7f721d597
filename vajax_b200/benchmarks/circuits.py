"""Benchmark circuits as flat device lists (the data contract of the reference engine:
``CircuitEngine.devices`` = ``[{name, model, nodes, params, is_openvaf}]``, ``num_nodes``,
``node_names``, ``analysis_params``; reference: vajax/analysis/engine.py:421-618).

The netlist parser, subcircuit flattening and ``generate`` blocks are out of scope
(SURVEY.md §2.1); these builders produce the flattened result directly.

* ``rc``      - reference netlist text: docs/getting_started.md:78-95.
* ``ring``    - 9-stage PSP103 ring (docs/VACASK-ANALYSIS.md:5-96): 18 MOSFETs + isource + vsource.
* ``multiplier(n)`` - n x n carry-save array multiplier in the mul64 / c6288 gate library
  (AND = NAND2 + inverter = 6 T, NOR2 = 4 T; nmos w=0.5u, pmos w=1u, l=0.2u;
  reference: vajax/benchmarks/data/mul64/runme.sim generate block, docs/vacask_sim_format.md:94-106).
  n=16 gives 256 AND + 2,128 NOR = 10,048 MOSFETs, 32 input resistors and 34 vsources - the
  c6288 device counts of docs/parallelism_architecture.md:10-19 (the real ISCAS netlist is not
  in the reference tree).
"""

from __future__ import annotations

from typing import Dict, List, Optional

NMOS_CARD = dict(type=1.0, vto=0.35, kp=3.0e-4, nfac=1.3, theta=0.2, gamma=0.4, phi=0.8, cox=1.2e-2,
                 cgso=3e-10, cgdo=3e-10, cgbo=1e-11, js=1e-6, cj=1e-3, cjsw=1e-10, tcv=1e-3, tnom=27.0)
PMOS_CARD = dict(type=-1.0, vto=0.35, kp=1.2e-4, nfac=1.35, theta=0.15, gamma=0.4, phi=0.8, cox=1.2e-2,
                 cgso=3e-10, cgdo=3e-10, cgbo=1e-11, js=1e-6, cj=1e-3, cjsw=1e-10, tcv=1e-3, tnom=27.0)


# psp103w (models/psp103w.py): the two model cards of the reference's benchmarks differ in TYPE only
# (vajax/benchmarks/data/mul64/runme.sim model psp103n / psp103p); everything else is the model default
# the knobs of the process corners (analysis/corners.py) are spelled out so that a corner has something to scale
NMOS_CARD_W = dict(type=1.0, uoo=3.5e-2, vfbo=-0.70, toxo=1.5e-9)
PMOS_CARD_W = dict(type=-1.0, uoo=1.75e-2, vfbo=-0.70, toxo=1.5e-9)
MOS_MODELS = ("psp103", "psp103w")


class _Net:
    def __init__(self, mos_model: str = "psp103"):
        assert mos_model in MOS_MODELS, mos_model
        self.names: Dict[str, int] = {"0": 0}
        self.devices: List[dict] = []
        self.mos_model = mos_model

    def node(self, name: str) -> int:
        if name not in self.names:
            self.names[name] = len(self.names)
        return self.names[name]

    def add(self, name, model, nodes, params, openvaf=True):
        self.devices.append(dict(name=name, model=model, nodes=[self.node(n) if isinstance(n, str) else n for n in nodes],
                                 params=dict(params), is_openvaf=openvaf))

    def mos(self, name, kind, d, g, s, b, w, l, ld=0.5e-6):
        if self.mos_model == "psp103w":
            card = NMOS_CARD_W if kind == "n" else PMOS_CARD_W
        else:
            card = NMOS_CARD if kind == "n" else PMOS_CARD
        p = dict(card)
        p.update(w=w, l=l)
        p["as"] = w * ld
        p["ad"] = w * ld
        p["ps"] = 2 * (w + ld)
        p["pd"] = 2 * (w + ld)
        self.add(name, self.mos_model, [d, g, s, b], p)

    def renumber_sorted(self):
        """Node indices by sorted node name, ground = 0 (reference: build_node_mapping,
        vajax/analysis/parsing.py:146-168)."""
        order = ["0"] + sorted(n for n in self.names if n != "0")
        new = {n: i for i, n in enumerate(order)}
        remap = [0] * len(order)
        for n, old in self.names.items():
            remap[old] = new[n]
        for d in self.devices:
            d["nodes"] = [remap[x] for x in d["nodes"]]
        self.names = new

    def circuit(self, analysis: dict, temperature_k: float = 300.15) -> dict:
        return dict(num_nodes=len(self.names), node_names=dict(self.names), devices=self.devices,
                    analysis=dict(analysis), temperature_k=temperature_k)


def rc_circuit(stop: float = 10e-3, method: str = "trap") -> dict:
    n = _Net()
    n.add("vs", "vsource", ["1", "0"],
          dict(dc=0.0, type="pulse", val0=0.0, val1=1.0, rise=1e-6, fall=1e-6, width=1e-3, period=2e-3), openvaf=False)
    n.add("r1", "resistor", ["1", "2"], dict(r=1e3))
    n.add("c1", "capacitor", ["2", "0"], dict(c=1e-6))
    return n.circuit(dict(step=1e-6, stop=stop, maxstep=1e-6, tran_method=method, icmode="op"))


def diode_clipper_circuit(stop: float = 2e-3, rs: float = 10.0, method: str = "trap") -> dict:
    """Sine source -> R -> two anti-parallel diodes (with series resistance -> internal node)
    + load capacitor. Exercises branches, limiting, internal nodes and node collapse."""
    n = _Net()
    n.add("vs", "vsource", ["in", "0"], dict(dc=0.0, type="sine", sinedc=0.0, ampl=2.0, freq=1e3, phase=0.0), openvaf=False)
    n.add("r1", "resistor", ["in", "out"], dict(r=1e3))
    n.add("d1", "diode", ["out", "0"], {"is": 1e-14, "n": 1.0, "rs": rs, "cj0": 1e-12, "vj": 0.8, "tt": 1e-9, "area": 1.0})
    n.add("d2", "diode", ["0", "out"], {"is": 1e-14, "n": 1.0, "rs": 0.0, "cj0": 1e-12, "vj": 0.8, "tt": 1e-9, "area": 2.0})
    n.add("c1", "capacitor", ["out", "0"], dict(c=1e-9))
    return n.circuit(dict(step=1e-6, stop=stop, maxstep=2e-5, tran_method=method, icmode="op"))


def opcov_circuit(stop: float = 4e-4, method: str = "trap") -> dict:
    """Two instances of the opcode-coverage model (one with its internal node, one collapsed) driven by a sine
    source through a resistor, with a load capacitor: every MIR opcode / callback the reference translates is
    evaluated each NR iteration (models/fixtures.py::opcov_bundle)."""
    n = _Net()
    n.add("vs", "vsource", ["in", "0"], dict(dc=0.0, type="sine", sinedc=0.3, ampl=0.6, freq=5e3, phase=0.0), openvaf=False)
    n.add("r1", "resistor", ["in", "out"], dict(r=2e3))
    n.add("x1", "opcov", ["out", "0"], dict(g0=1e-3, pexp=1.7, nf=3.0, mode=2.0, rser=50.0, cq=1e-9, vto=0.4, isat=1e-12, scale=1.0))
    n.add("x2", "opcov", ["0", "out"], dict(g0=1e-3, pexp=2.2, nf=6.0, mode=2.0, rser=0.0, cq=1e-9, vto=0.4, isat=1e-12, scale=0.5))
    n.add("c1", "capacitor", ["out", "0"], dict(c=2e-9))
    return n.circuit(dict(step=1e-6, stop=stop, maxstep=1e-5, tran_method=method, icmode="op"))


def source_mix_circuit(stop: float = 6e-8, method: str = "trap") -> dict:
    """One RC low-pass per source kind the reference implements beyond dc / pulse / sine
    (vajax/analysis/sources.py:31-187): pwl, exp, am, fm on voltage sources, plus a pwl current source."""
    n = _Net()
    srcs = [
        ("pwl", dict(type="pwl", wave=[0, 0, 1e-9, 1, 5e-9, -1, 2e-8, 0.5])),
        ("exp", dict(type="exp", val0=0.0, val1=1.0, delay=1e-9, td2=2e-8, tau1=3e-9, tau2=5e-9)),
        ("am", dict(type="am", sinedc=0.2, ampl=0.8, freq=1e8, phase=30.0, delay=2e-9, modfreq=2e7, modphase=10.0, modindex=0.5)),
        ("fm", dict(type="fm", sinedc=0.1, ampl=0.7, freq=1e8, phase=20.0, delay=1e-9, modfreq=1.5e7, modphase=5.0, modindex=2.0)),
    ]
    for k, (nm, p) in enumerate(srcs):
        p = dict(p)
        p.setdefault("dc", 0.0)
        n.add(f"v{nm}", "vsource", [f"in_{nm}", "0"], p, openvaf=False)
        n.add(f"r{nm}", "resistor", [f"in_{nm}", f"out_{nm}"], dict(r=1e3 * (k + 1)))
        n.add(f"c{nm}", "capacitor", [f"out_{nm}", "0"], dict(c=1e-12))
    n.add("ipwl", "isource", ["0", "out_pwl"], dict(dc=0.0, type="pwl", wave=[0, 0, 1e-8, 1e-4, 3e-8, -1e-4, 5e-8, 0]), openvaf=False)
    return n.circuit(dict(step=1e-10, stop=stop, maxstep=5e-10, tran_method=method, icmode="op"))


def ring_circuit(stages: int = 9, stop: float = 20e-9, step: float = 0.05e-9, w: float = 10e-6,
                 l: float = 1e-6, pfact: float = 2.0, vdd: float = 1.2, method: str = "trap",
                 mos_model: str = "psp103") -> dict:
    n = _Net(mos_model)
    for k in range(1, stages + 1):
        n.node(str(k))
    n.node("vdd")
    n.add("vdd", "vsource", ["vdd", "0"], dict(dc=vdd, type="dc"), openvaf=False)
    n.add("i0", "isource", ["0", "1"],
          dict(dc=0.0, type="pulse", val0=0.0, val1=10e-6, delay=1e-9, rise=1e-9, fall=1e-9, width=1e-9, period=1.0),
          openvaf=False)
    for k in range(1, stages + 1):
        i, o = str(k), str(k % stages + 1)
        n.mos(f"u{k}:mp:m", "p", o, i, "vdd", "vdd", w * pfact, l)
        n.mos(f"u{k}:mn:m", "n", o, i, "0", "0", w, l)
    return n.circuit(dict(step=step, stop=stop, maxstep=step, tran_method=method, icmode="op"))


def multiplier_circuit(nbits: int = 16, stop: float = 2e-9, step: float = 2e-12, vdd: float = 1.2,
                       a_val: Optional[int] = None, b_val: Optional[int] = None,
                       wn: float = 0.5e-6, wp: float = 1e-6, l: float = 0.2e-6, mos_model: str = "psp103") -> dict:
    """n x n unsigned carry-save array multiplier, all-ones operands by default (mul64's stimulus),
    ``icmode=uic tran_lteratio=1.5 nr_convtol=1`` as in the c6288/mul64 control blocks
    (docs/transient_options.md:158-163)."""
    N = nbits
    a_val = (1 << N) - 1 if a_val is None else a_val
    b_val = (1 << N) - 1 if b_val is None else b_val
    net = _Net(mos_model)
    net.node("vdd")
    net.node("vss")
    net.add("vdd", "vsource", ["vdd", "0"], dict(dc=vdd, type="dc"), openvaf=False)
    net.add("vss", "vsource", ["vss", "0"], dict(dc=0.0, type="dc"), openvaf=False)
    cnt = [0]

    def uid(p):
        cnt[0] += 1
        return f"{p}{cnt[0]}"

    def inv(a, y):
        g = uid("inv")
        net.mos(g + ":mp", "p", y, a, "vdd", "vdd", wp, l)
        net.mos(g + ":mn", "n", y, a, "vss", "vss", wn, l)

    def nand2(a, b, y):
        g = uid("nand")
        mid = g + ":x"
        net.mos(g + ":mp1", "p", y, a, "vdd", "vdd", wp, l)
        net.mos(g + ":mp2", "p", y, b, "vdd", "vdd", wp, l)
        net.mos(g + ":mn1", "n", y, a, mid, "vss", wn, l)
        net.mos(g + ":mn2", "n", mid, b, "vss", "vss", wn, l)

    def and2(a, b, y):
        t = uid("andn")
        nand2(a, b, t)
        inv(t, y)

    def nor2(a, b, y):
        g = uid("nor")
        mid = g + ":x"
        net.mos(g + ":mp1", "p", mid, a, "vdd", "vdd", wp, l)
        net.mos(g + ":mp2", "p", y, b, mid, "vdd", wp, l)
        net.mos(g + ":mn1", "n", y, a, "vss", "vss", wn, l)
        net.mos(g + ":mn2", "n", y, b, "vss", "vss", wn, l)

    def full_adder(a, b, c, s, co):
        # 9-NOR full adder (dual of the 9-NAND adder; XOR3 and MAJ are self-dual)
        p = uid("fa")
        n1, n2, n3, n4, n5, n6, n7 = (f"{p}:n{k}" for k in range(1, 8))
        nor2(a, b, n1); nor2(a, n1, n2); nor2(b, n1, n3); nor2(n2, n3, n4)
        nor2(n4, c, n5); nor2(n4, n5, n6); nor2(c, n5, n7); nor2(n6, n7, s); nor2(n1, n5, co)

    def half_adder(a, b, s, co):
        # 7 NOR: xnor tree, sum = not(xnor), carry = nor(nor(a,b), sum); one buffer NOR
        p = uid("ha")
        n1, n2, n3, n4, n5 = (f"{p}:n{k}" for k in range(1, 6))
        nor2(a, b, n1); nor2(a, n1, n2); nor2(b, n1, n3); nor2(n2, n3, n4)
        nor2(n4, n4, s); nor2(n1, s, co); nor2(co, co, n5)

    # inputs: pulse sources behind 1-ohm driver resistors (mul64: 128 r with r=1.0)
    for nm, val in (("a", a_val), ("b", b_val)):
        for i in range(N):
            bit = (val >> i) & 1
            src = f"{nm}{i}_src"
            net.add(f"v{nm}{i}", "vsource", [src, "0"],
                    dict(dc=0.0, type="pulse", val0=0.0, val1=vdd * bit, delay=step * 10, rise=step * 25,
                         fall=step * 25, width=1.0, period=2.0), openvaf=False)
            net.add(f"r{nm}{i}", "resistor", [src, f"{nm}{i}"], dict(r=1.0))
    pp = [[f"pp{i}_{j}" for j in range(N)] for i in range(N)]
    for i in range(N):
        for j in range(N):
            and2(f"a{i}", f"b{j}", pp[i][j])
    # row 1: half adders
    s_prev = [None] * N
    c_prev = [None] * N
    for i in range(N - 1):
        s_prev[i], c_prev[i] = f"s1_{i}", f"c1_{i}"
        half_adder(pp[i + 1][0], pp[i][1], s_prev[i], c_prev[i])
    s_prev[N - 1] = pp[N - 1][1]
    for j in range(2, N):
        s_cur = [None] * N
        c_cur = [None] * N
        for i in range(N - 1):
            s_cur[i], c_cur[i] = f"s{j}_{i}", f"c{j}_{i}"
            full_adder(pp[i][j], s_prev[i + 1], c_prev[i], s_cur[i], c_cur[i])
        s_cur[N - 1] = pp[N - 1][j]
        s_prev, c_prev = s_cur, c_cur
    # vector-merging ripple adder
    carry = None
    for k in range(N - 1):
        so, co = f"p{N + k}", f"rc{k}"
        if carry is None:
            half_adder(s_prev[k + 1], c_prev[k], so, co)
        else:
            full_adder(s_prev[k + 1], c_prev[k], carry, so, co)
        carry = co
    return net.circuit(dict(step=step, stop=stop, icmode="uic", nr_convtol=1, tran_lteratio=1.5, tran_method="trap"))


def mul64_circuit(nbits: int = 64, stop: float = 2e-9, step: float = 2e-12, vdd: float = 1.2,
                  wn: float = 0.5e-6, wp: float = 1e-6, l: float = 0.2e-6, mos_model: str = "psp103") -> dict:
    """The reference's mul64 benchmark (vajax/benchmarks/data/mul64/runme.sim), flattened.

    Gate library of the netlist's generate block (:23-59): not = 2 T, nor = 4 T, and = 6 T
    (nand + inverter), xor = 3 not + 2 and + nor = 22 T, ha = xor + and = 28 T,
    fa = 2 xor + 2 and + nor + not = 62 T.  Array topology (:61-139): n^2 partial products, a two
    inverter buffer for p0, row 0 of n-1 half adders, rows 1..n-2 with a half adder in column 0
    and full adders in columns 1..n-2.  Test bench (:141-292): one ``drv`` (pulse source behind a
    1 ohm resistor, delay = rise = 0.1 ns) per operand bit, all bits driven high, rails vdd / vss
    behind DC sources, ``tran stop=2n step=2p icmode=uic``, ``nr_convtol=1 tran_lteratio=1.5``.
    n=64 gives 266,408 MOSFETs + 128 resistors + 130 vsources - the counts the reference reports
    for mul64.  Instances are emitted in the order of the reference's flattening
    (depth first, vajax/analysis/parsing.py:32-143) with its hierarchical names (``top.u1.fa_3_2.x1.a1.mn1.m``)
    and its node numbering (sorted names, parsing.py:146-168); tests/golden/mul64_topology.json holds
    the digest of the reference parser's own result for n=64."""
    N = nbits
    net = _Net(mos_model)
    net.node("vdd")
    net.node("vss")
    net.add("vdd", "vsource", ["vdd", "0"], dict(dc=vdd, type="dc"), openvaf=False)
    net.add("vss", "vsource", ["vss", "0"], dict(dc=0.0, type="dc"), openvaf=False)

    def inv(p, out, a):
        net.mos(p + ".mp.m", "p", out, a, "vdd", "vdd", wp, l)
        net.mos(p + ".mn.m", "n", out, a, "vss", "vss", wn, l)

    def nor(p, out, a, b):
        mid = p + ".int"
        net.mos(p + ".mp2.m", "p", mid, b, "vdd", "vdd", wp, l)
        net.mos(p + ".mp1.m", "p", out, a, mid, "vdd", wp, l)
        net.mos(p + ".mn1.m", "n", out, a, "vss", "vss", wn, l)
        net.mos(p + ".mn2.m", "n", out, b, "vss", "vss", wn, l)

    def and2(p, out, a, b):
        ox, mid = p + ".outx", p + ".int"
        net.mos(p + ".mp2.m", "p", ox, b, "vdd", "vdd", wp, l)
        net.mos(p + ".mp1.m", "p", ox, a, "vdd", "vdd", wp, l)
        net.mos(p + ".mn1.m", "n", ox, a, mid, "vss", wn, l)
        net.mos(p + ".mn2.m", "n", mid, b, "vss", "vss", wn, l)
        net.mos(p + ".mp3.m", "p", out, ox, "vdd", "vdd", wp, l)
        net.mos(p + ".mn3.m", "n", out, ox, "vss", "vss", wn, l)

    def xor(p, out, a, b):
        inv(p + ".inv_a", p + ".na", a)
        inv(p + ".inv_b", p + ".nb", b)
        and2(p + ".a1", p + ".t1", a, p + ".nb")
        and2(p + ".a2", p + ".t2", p + ".na", b)
        nor(p + ".n1", p + ".t3", p + ".t1", p + ".t2")
        inv(p + ".inv_out", out, p + ".t3")

    def ha(p, s, co, a, b):
        xor(p + ".x1", s, a, b)
        and2(p + ".a1", co, a, b)

    def fa(p, s, co, a, b, cin):
        xor(p + ".x1", p + ".ab_xor", a, b)
        xor(p + ".x2", s, p + ".ab_xor", cin)
        and2(p + ".a1", p + ".ab_and", a, b)
        and2(p + ".a2", p + ".xor_cin", p + ".ab_xor", cin)
        nor(p + ".n1", p + ".cout_n", p + ".ab_and", p + ".xor_cin")
        inv(p + ".inv_cout", co, p + ".cout_n")

    last_row = N - 2

    def sum_name(row, col):
        if col == 0:
            return f"top.p{row + 1}"
        if row == last_row:
            return f"top.p{N + col - 1}" if col < N - 1 else f"top.u1.s{row}_{col}"
        return f"top.u1.s{row}_{col}"

    def carry_name(row, col):
        if row == last_row and col == N - 2:
            return f"top.p{2 * N - 2}"
        return f"top.u1.c{row}_{col}"

    for i in range(N):
        for j in range(N):
            and2(f"top.u1.pp_{i}_{j}", f"top.u1.pp{i}_{j}", f"top.a{i}", f"top.b{j}")
    inv("top.u1.buf_p0_inv", "top.u1.p0_n", "top.u1.pp0_0")
    inv("top.u1.buf_p0", "top.p0", "top.u1.p0_n")
    for j in range(N - 1):
        ha(f"top.u1.ha_0_{j}", sum_name(0, j), carry_name(0, j), f"top.u1.pp0_{j + 1}", f"top.u1.pp1_{j}")
    for row in range(1, N - 1):
        for j in range(N - 1):
            sn, cn = sum_name(row, j), carry_name(row, j)
            b_in = f"top.u1.pp{row + 1}_{j}"
            a_in = sum_name(row - 1, j + 1) if j < N - 2 else carry_name(row - 1, N - 2)
            if j == 0:
                ha(f"top.u1.ha_{row}_{j}", sn, cn, a_in, b_in)
            else:
                fa(f"top.u1.fa_{row}_{j}", sn, cn, a_in, b_in, carry_name(row, j - 1))
    for nm in ("a", "b"):
        for i in range(N):
            src = f"top.d{nm}{i}.int"
            net.add(f"top.d{nm}{i}.vdrv", "vsource", [src, "0"],
                    dict(dc=0.0, type="pulse", val0=0.0, val1=vdd, delay=0.1e-9, rise=0.1e-9), openvaf=False)
            net.add(f"top.d{nm}{i}.rdrv", "resistor", [src, f"top.{nm}{i}"], dict(r=1.0))
    net.renumber_sorted()
    return net.circuit(dict(step=step, stop=stop, icmode="uic", nr_convtol=1, tran_lteratio=1.5, tran_method="trap"))


def split_psp_types(circuit: dict) -> dict:
    """The mul64 netlist declares two models (psp103n / psp103p) of the same module; the engine
    groups instances by model *type* (reference: engine.py:1113-1121).  Our fixtures use the module
    name ``psp103`` for both, with TYPE as a (varying) parameter, like the reference where TYPE is
    a model-card parameter."""
    out = dict(circuit)
    devs = []
    for d in circuit["devices"]:
        d = dict(d)
        if d["model"] in ("psp103n", "psp103p"):
            d["model"] = "psp103"
        devs.append(d)
    out["devices"] = devs
    return out
