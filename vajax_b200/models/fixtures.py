"""Model bundles in the OpenVAF MIR dict schema, built with ``vajax_b200.mir.builder``.

Why these exist: the reference compiles ``resistor.va`` / ``capacitor.va`` / ``diode.va`` /
``psp103.va`` with OpenVAF (Rust + LLVM) and hands the MIR dicts to its code generator
(reference: vajax/analysis/openvaf_models.py:338-545).  Neither OpenVAF nor
``psp103.va``/``diode.va`` exist in the build environment, so:

* ``resistor`` and ``capacitor`` restate the bundled Verilog-A sources
  (reference: vajax/devices/models/resistor.va, capacitor.va): ``I = V/r``, ``q = c*V``.
* ``diode`` is a junction diode with series resistance, an internal node ``CI`` that
  collapses when ``rs == 0``, ``$limit(..."pnjlim")``, ``$simparam("gmin")`` and a
  piecewise depletion charge - it exercises branches/PHI, callbacks, limiting, init cache
  and node collapse.
* ``mos_psp_shaped`` is a **surrogate** for PSP103: same terminal / internal-node structure
  (D,G,S,B + NOI,GP,SI,DI,BP,BI,BS,BD + one noise-branch flow = 13 unknowns; everything but
  NOI and the flow collapses for the benchmark model cards), TYPE-signed, smooth
  EKV-style channel current, charge-conserving gate/bulk/channel charges, junction diodes,
  overlap capacitances, NOI node with 1e40 S to ground (reference: CLAUDE.md:196-201).  It is
  NOT PSP103 physics; it exists so that ring / c6288-shaped / mul64 circuits run through the
  real data path.  A real PSP103 ``mir_data.pkl`` can replace it via ``load_bundle``.
"""

from __future__ import annotations

import pickle
from typing import Dict

from ..mir import builder as B
from ..mir.builder import Dual, IfElse, ModelBuilder, cmp, exp, ln, select, sqrt

KB_Q = 1.380649e-23 / 1.602176634e-19


def resistor_bundle() -> dict:
    m = ModelBuilder("resistor", ["A", "B"])
    r_i, r = m.model_param("r", 1.0, is_instance=True)
    _hn_i, _hn = m.model_param("has_noise", 1.0)
    _t_i, _t = m.temperature()
    mf = m.mfactor()
    v = m.voltage("A", "B")
    i = v / r
    # white_noise(4kT/r) contributes nothing to the large-signal residual
    m.eval.call("white_noise", [(Dual(m.eval, m.eval.fconst(4.0 * 1.380649e-23)) * _t / r).v, m.eval.sconst("thermal")])
    m.contrib("A", "B", resist=i * mf)
    return m.finalize()


def capacitor_bundle() -> dict:
    m = ModelBuilder("capacitor", ["A", "B"])
    c_i, c = m.model_param("c", 1e-12, is_instance=True)
    mf = m.mfactor()
    v = m.voltage("A", "B")
    q = c * v
    m.contrib("A", "B", react=q * mf)
    return m.finalize()


def _limexp(x: Dual) -> Dual:
    """exp(x) linearly extrapolated above x=80 (diamond if/else -> PHI)."""
    fb = x.fb
    c = cmp("fgt", x, 80.0)
    ie = IfElse(fb, c, "diamond")
    e80 = Dual(fb, fb.fconst(5.540622384393510e34))  # exp(80)
    a = e80 * (x - 79.0)
    ie.otherwise()
    b = exp(x)
    ie.end()
    return ie.phi(a, b)


def diode_bundle() -> dict:
    m = ModelBuilder("diode", ["A", "C"], ["CI"])
    is_i, is_e = m.model_param("is", 1e-14)
    n_i, n_e = m.model_param("n", 1.0)
    rs_i, rs_e = m.model_param("rs", 0.0)
    cj0_i, cj0_e = m.model_param("cj0", 0.0)
    vj_i, vj_e = m.model_param("vj", 1.0)
    mj_i, mj_e = m.model_param("m", 0.5)
    tt_i, tt_e = m.model_param("tt", 0.0)
    area_i, area_e = m.model_param("area", 1.0, is_instance=True)
    t_i, _t_e = m.temperature()
    mf = m.mfactor()
    fi = m.init

    # ---- init: thermal voltage, critical voltage, scaled currents -> cache
    vt_i = t_i * KB_Q * n_i
    isa_i = is_i * area_i
    vcrit_i = vt_i * ln(vt_i / (isa_i * 1.4142135623730951))
    cja_i = cj0_i * area_i
    rs_pos = cmp("fgt", rs_i, 0.0)
    ie = IfElse(fi, rs_pos, "triangle")
    g_then = area_i / rs_i
    ie.end()
    gs_i = ie.phi(g_then, 0.0)
    vt = m.cache(vt_i)
    isa = m.cache(isa_i)
    vcrit = m.cache(vcrit_i)
    cja = m.cache(cja_i)
    gs = m.cache(gs_i)
    # CI collapses into C when rs is not positive
    m.collapsible("CI", "C", decision_init=rs_pos, negate=True)

    # ---- eval
    fe = m.eval
    vraw = m.voltage("A", "CI")
    prev, _en = m.limit_params(0)
    # $limit(V(A,CI), "pnjlim", vt, vcrit): BuiltinLimit -> StoreLimit (the limited voltage is
    # treated as the unknown, so it inherits the raw branch derivatives)
    vlim_v = fe.call("BuiltinLimit { name: pnjlim, num_args: 4 }", [vraw.v, prev.v, vt.v, vcrit.v])
    vst_v = fe.call("StoreLimit(LimState(0))", [vlim_v])
    vd = Dual(fe, vst_v, dict(vraw.d))
    gmin = Dual(fe, fe.call("SimParam", [fe.sconst("gmin")]))
    idio = isa * (_limexp(vd / vt) - 1.0) + gmin * vd
    # depletion charge: piecewise at fc*vj (fc = 0.5)
    fcv = vj_e * 0.5
    c2 = cmp("flt", vd, fcv)
    ie3 = IfElse(fe, c2, "diamond")
    one_m = 1.0 - vd / vj_e
    qa = cja * vj_e * (1.0 - sqrt(one_m)) * 2.0  # m = 0.5 closed form
    ie3.otherwise()
    dv = vd - fcv
    qb = cja * vj_e * 0.5857864376269049 + cja * 1.4142135623730951 * (dv + dv * dv * (0.5 / vj_e))
    ie3.end()
    qdep = ie3.phi(qa, qb)
    _ = mj_e
    qd = tt_e * idio + qdep
    m.contrib("A", "CI", resist=idio * mf, react=qd * mf)
    vser = m.voltage("CI", "C")
    m.contrib("CI", "C", resist=vser * gs * mf)
    return m.finalize()


def opcov_bundle() -> dict:
    """Opcode-coverage model: every MIR opcode and callback class the reference translates
    (openvaf_jax/codegen/instruction.py:156-263 opcodes, :469-703 callbacks) is on the executed path of a small
    device, so that the reference code generator, the oracle interpreter and the emitted CUDA kernel can be
    compared on ``pow atan2 fmin fmax`` (not ``hypot``: the reference lists it among its unary functions first,
    instruction.py:84-96,222-225, and its generated ``jnp.hypot(x)`` raises - there is no reference result to match), the trigonometric / hyperbolic family, ``floor ceil clog2 log``,
    ``frem``, casts, integer / bitwise / boolean ops, every comparison, ``select copy optbarrier``, ``$simparam``,
    ``analysis()``, ``ddt``, ``ddx``, noise, ``abs/min/max`` calls, ``$limit`` with pnjlim AND fetlim, nested
    diamonds, triangles and a three-way PHI.  Topology: A --[nonlinear branch]-- M --[rser]-- B, M collapses
    into B when rser == 0.  Every branch current is a bounded, smooth-ish function of its voltage plus a linear
    conductance, so the device also survives a DC solve and a transient."""
    m = ModelBuilder("opcov", ["A", "B"], ["M"])
    P = {}
    for pname, dflt, inst in [("g0", 1e-3, False), ("pexp", 1.7, False), ("nf", 3.0, True), ("mode", 2.0, False),
                              ("rser", 50.0, False), ("cq", 1e-12, False), ("vto", 0.4, False), ("isat", 1e-12, False),
                              ("scale", 1.0, True)]:
        P[pname] = m.model_param(pname, dflt, is_instance=inst)
    t_i, t_e = m.temperature()
    mf = m.mfactor()
    fi, fe = m.init, m.eval
    I = lambda k: P[k][0]  # noqa: E731
    E = lambda k: P[k][1]  # noqa: E731

    # ---- init: integer arithmetic on the finger count, casts, cache values, collapse decision
    nf_i = fi.emit("ficast", [I("nf").v], ty="i")                        # int(floor(nf))
    nf2 = fi.emit("imul", [nf_i, fi.iconst(2)], ty="i")
    nf3 = fi.emit("iadd", [nf2, fi.iconst(1)], ty="i")
    nfh = fi.emit("idiv", [nf3, fi.iconst(2)], ty="i")                   # floor division
    nfr = fi.emit("irem", [nf3, fi.iconst(4)], ty="i")
    nfs = fi.emit("isub", [nfh, nfr], ty="i")
    nfn = fi.emit("ineg", [nfs], ty="i")
    nfx = fi.emit("ixor", [fi.emit("iand", [nf3, fi.iconst(6)], ty="i"), fi.emit("ior", [nfr, fi.iconst(8)], ty="i")], ty="i")
    nfsh = fi.emit("ishr", [fi.emit("ishl", [nfx, fi.iconst(3)], ty="i"), fi.iconst(1)], ty="i")
    nfnot = fi.emit("inot", [nfn], ty="i")
    geo_i = Dual(fi, fi.emit("ifcast", [fi.emit("iadd", [nfsh, nfnot], ty="i")]))
    odd = fi.emit("ine", [nfr, fi.iconst(0)], ty="b")
    big = fi.emit("ige", [nf_i, fi.iconst(3)], ty="b")
    both = fi.emit("beq", [odd, big], ty="b")
    fsel = Dual(fi, fi.emit("select", [both, fi.fconst(1.25), fi.fconst(0.75)]))
    vt_i = t_i * KB_Q
    geff_i = I("g0") * I("scale") * (1.0 + geo_i * 1e-3) * fsel
    lg2_i = B.clog2(I("nf") + 4.0)
    l10_i = B.log10(I("g0") * 1e6 + 10.0)
    rs_pos = cmp("fgt", I("rser"), 0.0)
    ie = IfElse(fi, rs_pos, "triangle")
    gs_then = 1.0 / I("rser")
    ie.end()
    gs_i = ie.phi(gs_then, 0.0)
    vt, geff, lg2, l10, gs = m.cache(vt_i), m.cache(geff_i), m.cache(lg2_i), m.cache(l10_i), m.cache(gs_i)
    vcrit = m.cache(vt_i * ln(vt_i / (I("isat") * 1.4142135623730951)))
    m.collapsible("M", "B", decision_init=rs_pos, negate=True)

    # ---- eval
    vraw = m.voltage("A", "M")
    prev0, _en = m.limit_params(0)
    prev1, _en = m.limit_params(1)
    # $limit(V(A,M), "pnjlim", vt, vcrit) feeds the junction term; $limit(V(A,M), "fetlim", vto) the FET-like term
    vl0 = fe.call("BuiltinLimit { name: pnjlim, num_args: 4 }", [vraw.v, prev0.v, vt.v, vcrit.v])
    vj = Dual(fe, fe.call("StoreLimit(LimState(0))", [vl0]), dict(vraw.d))
    vl1 = fe.call("BuiltinLimit { name: fetlim, num_args: 3 }", [vraw.v, prev1.v, E("vto").v])
    vf = Dual(fe, fe.call("StoreLimit(LimState(1))", [vl1]), dict(vraw.d))
    v = vraw
    gmin = Dual(fe, fe.call("SimParam", [fe.sconst("gmin")]))
    abst = Dual(fe, fe.call("SimParam", [fe.sconst("abstol")]))
    tnom = Dual(fe, fe.call("SimParamOpt", [fe.sconst("tnom"), fe.fconst(27.0)]))
    is_tran = fe.call("Analysis", [fe.sconst("tran")], ty="b")
    is_dc = fe.call("Analysis", [fe.sconst("dc")], ty="b")
    unit = geff * 1e-1  # currents of the exotic terms: ~1e-4 A/V scale

    # (1) powers / two-argument functions
    t1 = B.power(B.fabs(v) + 0.1, E("pexp")) * 0.3 + B.atan2(v, v * v + 0.5) \
        + B.fmin(v, 0.8) * 0.5 + B.fmax(v, -0.2) * 0.25
    # (2) trigonometric / hyperbolic family on bounded arguments
    vb = B.tanh(v)                      # in (-1, 1)
    t2 = B.sin(v) + B.cos(v * 2.0) + B.tan(vb * 0.5) + B.sinh(vb) + B.cosh(vb) + B.asin(vb * 0.5) + B.acos(vb * 0.5) \
        + B.atan(v) + B.asinh(v) + B.acosh(v * v + 1.5) + B.atanh(vb * 0.5)
    # (3) piecewise-constant functions of the voltage (zero derivative), float remainder, casts
    seg = B.floor(v * 3.0) + B.ceil(v * 2.0) * 0.5
    k_i = fe.emit("ficast", [(v * 4.0).v], ty="i")
    k_odd = fe.emit("ibcast", [fe.emit("iand", [k_i, fe.iconst(1)], ty="i")], ty="b")
    k_f = Dual(fe, fe.emit("bfcast", [k_odd]))
    k_b = Dual(fe, fe.emit("ifcast", [fe.emit("bicast", [fe.emit("fbcast", [seg.v], ty="b")], ty="i")]))
    rem = Dual(fe, fe.emit("frem", [(v * v + 0.3).v, fe.fconst(0.25)]))
    t3 = (seg * 0.01 + k_f * 0.02 + k_b * 0.03 + rem * 0.04) * v
    # (4) comparisons + select + boolean ops
    c_lt, c_le = cmp("flt", v, 0.1), cmp("fle", v, 0.3)
    c_gt, c_ge = cmp("fgt", v, 0.7), cmp("fge", v, 0.9)
    c_eq, c_ne = cmp("feq", v, 0.0), cmp("fne", v, 0.5)
    c_a = fe.emit("bne", [c_lt, c_ge], ty="b")
    c_b = fe.emit("bnot", [fe.emit("beq", [c_le, c_gt], ty="b")], ty="b")
    mode_i = fe.emit("ficast", [E("mode").v], ty="i")
    c_m = fe.emit("ieq", [mode_i, fe.iconst(2)], ty="b")
    c_m2 = fe.emit("ilt", [mode_i, fe.iconst(5)], ty="b")
    c_m3 = fe.emit("igt", [mode_i, fe.iconst(0)], ty="b")
    c_m4 = fe.emit("ile", [mode_i, fe.iconst(2)], ty="b")
    t4 = select(c_a, v * 0.2, v * 0.1) + select(c_b, v * v * 0.05, 0.0) + select(c_eq, v + 1.0, 0.0) * 1e-3 \
        + select(c_ne, v, 0.0) * 0.01 + select(c_m, v, 0.0) * 0.02 + select(c_m2, v, 0.0) * 0.03 \
        + select(c_m3, v, 0.0) * 0.04 + select(c_m4, v, 0.0) * 0.05
    # (5) logs, copy / optbarrier, call-style abs / min / max
    cp = Dual(fe, fe.emit("copy", [v.v]), dict(v.d))
    ob = Dual(fe, fe.emit("optbarrier", [(v * 0.5).v]), {k: fe.emit("fmul", [fe.fconst(0.5), d]) for k, d in v.d.items()})
    c_abs = Dual(fe, fe.call("abs", [v.v]), {k: fe.emit("select", [c_lt, d, d]) for k, d in v.d.items()})
    c_min = Dual(fe, fe.call("min", [v.v, fe.fconst(0.6)]))
    c_max = Dual(fe, fe.call("max", [v.v, fe.fconst(-0.1)]))
    t5 = B.log10(v * v + 1.0) + ln(v * v + 2.0) + sqrt(v * v + 0.01) + exp(vb) + cp * 0.1 + ob * 0.1 \
        + c_abs * 1e-3 + (c_min + c_max) * 1e-3 + lg2 * 1e-3 + l10 * 1e-3
    # (6) control flow: nested diamond inside a diamond, a triangle, and a three-way merge
    outer = IfElse(fe, c_gt, "diamond")
    inner = IfElse(fe, c_ge, "diamond")
    a_in = v * 0.9 + 0.07
    inner.otherwise()
    b_in = v * v
    inner.end()
    hi = inner.phi(a_in, b_in)
    outer.otherwise()
    half = v * 0.5       # (the fall-through value of a triangle has to be defined before the branch)
    tri = IfElse(fe, c_lt, "triangle")
    lo_then = B.sin(v) * 0.5
    tri.end()
    lo = tri.phi(lo_then, half)
    outer.end()
    t6 = outer.phi(hi, lo)
    # three-way PHI: block X branches to P or Q, Q branches to R or S; P, R, S all jump to the merge block
    b_p, b_q, b_r, b_s, b_m = (fe.new_block() for _ in range(5))
    fe.br(c_le, b_p, b_q)
    fe.cur = b_p
    w_p = v * 0.11
    fe.jmp(b_m)
    fe.cur = b_q
    fe.br(c_gt, b_r, b_s)
    fe.cur = b_r
    w_r = v * 0.07 + 0.028
    fe.jmp(b_m)
    fe.cur = b_s
    w_s = v * v * 0.1
    fe.jmp(b_m)
    fe.cur = b_m
    zero = fe.fconst(0.0)
    t7 = Dual(fe, fe.phi([(b_p, w_p.v), (b_r, w_r.v), (b_s, w_s.v)]),
              {k: fe.phi([(b_p, w_p.d.get(k, zero)), (b_r, w_r.d.get(k, zero)), (b_s, w_s.d.get(k, zero))]) for k in v.d})
    # (7) callbacks that evaluate to zero or pass through
    fe.call("white_noise", [(gmin * 4.0).v, fe.sconst("thermal")])
    fe.call("flicker_noise", [gmin.v, fe.fconst(1.0), fe.sconst("flicker")])
    ddx0 = Dual(fe, fe.call("ddx", [v.v]))
    fe.call("LimDiscontinuity", [])
    fe.call("Print { kind: Display, arg_tys: [] }", [fe.sconst("opcov\n")])
    fe.call("set_MinInclusive", [v.v])
    fe.call("SetRetFlag(Finish)", [])
    fe.call("CollapseHint(node2, Some(node1))", [])
    tran_f = Dual(fe, fe.emit("bfcast", [is_tran]))
    dc_f = Dual(fe, fe.emit("bfcast", [is_dc]))
    # junction (pnjlim-limited) and FET-like (fetlim-limited) terms
    ij = E("isat") * (_limexp(vj / vt) - 1.0)
    vov = vf - E("vto")
    ifet = geff * 0.5 * (vov + sqrt(vov * vov + 0.01)) * B.tanh(vf * 2.0)
    inl = geff * v + unit * (t1 + t2 + t3 + t4 + t5 + t6 + t7) * 1e-2 + ij + ifet + gmin * v + abst * 1e-3 * v \
        + (tnom - 27.0) * 1e-9 * v + ddx0 + tran_f * 1e-9 * v + dc_f * 2e-9 * v
    qch = Dual(fe, fe.call("ddt", [(E("cq") * (v + B.tanh(v) * 0.2)).v]), None)
    q = E("cq") * (v + B.tanh(v) * 0.2)
    qd = Dual(fe, qch.v, dict(q.d))
    _ = t_e
    m.contrib("A", "M", resist=inl * mf, react=qd * mf)
    m.contrib("M", "B", resist=m.voltage("M", "B") * gs * mf)
    return m.finalize()


def _softplus(x: Dual) -> Dual:
    """ln(1+exp(x)), passing x through above 40 (triangle-free diamond)."""
    fb = x.fb
    c = cmp("fgt", x, 40.0)
    ie = IfElse(fb, c, "diamond")
    a = x + 0.0
    ie.otherwise()
    b = ln(1.0 + exp(x))
    ie.end()
    return ie.phi(a, b)


def mos_psp_shaped_bundle(name: str = "psp103") -> dict:
    terms = ["D", "G", "S", "B"]
    internal = ["NOI", "GP", "SI", "DI", "BP", "BI", "BS", "BD", "noi_flow"]
    m = ModelBuilder(name, terms, internal, internal_kinds={"noi_flow": "BranchCurrent"})
    fi, fe = m.init, m.eval
    P: Dict[str, tuple] = {}
    for pname, dflt, inst in [
        ("type", 1.0, False), ("vto", 0.35, False), ("kp", 3e-4, False), ("nfac", 1.3, False),
        ("theta", 0.2, False), ("gamma", 0.4, False), ("phi", 0.8, False), ("cox", 1.2e-2, False),
        ("cgso", 3e-10, False), ("cgdo", 3e-10, False), ("cgbo", 1e-11, False),
        ("js", 1e-6, False), ("cj", 1e-3, False), ("cjsw", 1e-10, False),
        ("rg", 0.0, False), ("rsh", 0.0, False), ("rbulk", 0.0, False), ("tcv", 1e-3, False),
        ("tnom", 27.0, False),
        ("w", 1e-6, True), ("l", 1e-6, True), ("as", 0.0, True), ("ad", 0.0, True),
        ("ps", 0.0, True), ("pd", 0.0, True),
    ]:
        P[pname] = m.model_param(pname, dflt, is_instance=inst)
    t_i, _ = m.temperature()
    mf = m.mfactor()
    I = lambda k: P[k][0]  # noqa: E731

    # ---- init
    vt_i = t_i * KB_Q
    beta_i = I("kp") * I("w") / I("l")
    coxt_i = I("cox") * I("w") * I("l")
    vto_i = I("vto") - I("tcv") * (t_i - (I("tnom") + 273.15))
    is_s_i = I("js") * I("as") + 1e-18
    is_d_i = I("js") * I("ad") + 1e-18
    cjs_i = I("cj") * I("as") + I("cjsw") * I("ps")
    cjd_i = I("cj") * I("ad") + I("cjsw") * I("pd")
    cgs_i = I("cgso") * I("w")
    cgd_i = I("cgdo") * I("w")
    cgb_i = I("cgbo") * I("l")

    def cond_g(r: Dual) -> tuple:
        c = cmp("fgt", r, 0.0)
        ie = IfElse(fi, c, "triangle")
        g = 1.0 / r
        ie.end()
        return c, ie.phi(g, 0.0)

    c_rg, g_rg_i = cond_g(I("rg"))
    c_rs, g_rs_i = cond_g(I("rsh"))
    c_rb, g_rb_i = cond_g(I("rbulk"))
    two_n_vt2_i = 2.0 * I("nfac") * vt_i * vt_i
    inv2vt_i = 0.5 / vt_i
    vt, beta, coxt, vto = m.cache(vt_i), m.cache(beta_i), m.cache(coxt_i), m.cache(vto_i)
    is_s, is_d, cjs, cjd = m.cache(is_s_i), m.cache(is_d_i), m.cache(cjs_i), m.cache(cjd_i)
    cgs, cgd, cgb = m.cache(cgs_i), m.cache(cgd_i), m.cache(cgb_i)
    g_rg, g_rs, g_rb = m.cache(g_rg_i), m.cache(g_rs_i), m.cache(g_rb_i)
    two_n_vt2, inv2vt = m.cache(two_n_vt2_i), m.cache(inv2vt_i)
    sgn = m.cache(I("type") + 0.0)
    nfac = m.cache(I("nfac") + 0.0)
    theta = m.cache(I("theta") + 0.0)
    gamma = m.cache(I("gamma") + 0.0)
    phi = m.cache(I("phi") + 0.0)
    m.collapsible("GP", "G", c_rg, negate=True)
    m.collapsible("SI", "S", c_rs, negate=True)
    m.collapsible("DI", "D", c_rs, negate=True)
    m.collapsible("BP", "BI", c_rb, negate=True)
    m.collapsible("BS", "BI", c_rb, negate=True)
    m.collapsible("BD", "BI", c_rb, negate=True)
    m.collapsible("BI", "B", c_rb, negate=True)

    # ---- eval
    vgb = m.voltage("GP", "BP")
    vsb = m.voltage("SI", "BP")
    vdb = m.voltage("DI", "BP")
    vg, vs, vd = sgn * vgb, sgn * vsb, sgn * vdb
    vp = (vg - vto) / nfac
    sf = _softplus((vp - vs) * inv2vt)
    sr = _softplus((vp - vd) * inv2vt)
    i_f, i_r = sf * sf, sr * sr
    betaeff = beta / (1.0 + theta * sqrt(vp * vp + 0.01))
    gmin = Dual(fe, fe.call("SimParam", [fe.sconst("gmin")]))
    vds_raw = m.voltage("DI", "SI")
    ids = sgn * (two_n_vt2 * betaeff * (i_f - i_r)) + gmin * vds_raw
    m.contrib("DI", "SI", resist=ids * mf)
    # charges
    xf, xr = sqrt(i_f + 0.25), sqrt(i_r + 0.25)
    qi = (nfac * vt * -2.0) * ((xf * xf + xf * xr + xr * xr) / (xf + xr) * 0.6666666666666666 - 0.5) * coxt
    vpp = (vp + phi + sqrt((vp + phi) * (vp + phi) + 4e-4)) * 0.5
    qb = -(coxt * gamma * sqrt(vpp))
    qg = -(qi + qb)
    m.contrib("GP", None, react=sgn * qg * mf)
    m.contrib("BP", None, react=sgn * qb * mf)
    m.contrib("DI", None, react=sgn * qi * 0.5 * mf)
    m.contrib("SI", None, react=sgn * qi * 0.5 * mf)
    # overlap capacitances
    m.contrib("GP", "SI", react=cgs * m.voltage("GP", "SI") * mf)
    m.contrib("GP", "DI", react=cgd * m.voltage("GP", "DI") * mf)
    m.contrib("GP", "BP", react=cgb * vgb * mf)
    # junction diodes (bulk -> source/drain), TYPE-signed
    for bn, dn, isat, cj in (("BS", "SI", is_s, cjs), ("BD", "DI", is_d, cjd)):
        vj = m.voltage(bn, dn)
        ij = sgn * (isat * (_limexp(sgn * vj / vt) - 1.0)) + gmin * vj
        m.contrib(bn, dn, resist=ij * mf, react=cj * vj * mf)
    # series / bulk resistances (zero conductance when collapsed)
    m.contrib("G", "GP", resist=m.voltage("G", "GP") * g_rg * mf)
    m.contrib("S", "SI", resist=m.voltage("S", "SI") * g_rs * mf)
    m.contrib("D", "DI", resist=m.voltage("D", "DI") * g_rs * mf)
    m.contrib("BP", "BI", resist=m.voltage("BP", "BI") * g_rb * mf)
    m.contrib("BS", "BI", resist=m.voltage("BS", "BI") * g_rb * mf)
    m.contrib("BD", "BI", resist=m.voltage("BD", "BI") * g_rb * mf)
    m.contrib("B", "BI", resist=m.voltage("B", "BI") * g_rb * mf)
    # noise-correlation node: 1/mig = 1e40 S to ground plus the flow unknown
    vnoi = m.voltage("NOI")
    iflow = Dual(fe, fe.param("current", "I(noi_flow)"), {"noi_flow": fe.fconst(1.0)})
    m.contrib("NOI", None, resist=vnoi * 1e40 + iflow)
    m.contrib("noi_flow", None, resist=iflow + 0.0)
    return m.finalize()


def _psp103w_bundle() -> dict:
    from .psp103w import psp103w_bundle

    return psp103w_bundle()


_BUILDERS = {
    "psp103w": _psp103w_bundle,
    "resistor": resistor_bundle,
    "capacitor": capacitor_bundle,
    "diode": diode_bundle,
    "psp103": mos_psp_shaped_bundle,
    "opcov": opcov_bundle,
}
_CACHE: Dict[str, dict] = {}


def get_bundle(model_type: str) -> dict:
    """Return the (memoised) MIR bundle for a fixture model type."""
    if model_type not in _CACHE:
        _CACHE[model_type] = _BUILDERS[model_type]()
    return _CACHE[model_type]


def load_bundle(path: str) -> dict:
    """Load an externally produced bundle (the reference's ``mir_data.pkl``,
    reference: vajax/analysis/openvaf_models.py:440-458)."""
    with open(path, "rb") as f:
        return pickle.load(f)
