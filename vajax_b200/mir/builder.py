"""MIR fixture builder: a stand-in *producer* for the OpenVAF MIR dict schema.

The reference obtains each Verilog-A model's MIR from the Rust crate ``openvaf_py``
(reference: openvaf_jax/openvaf_py/src/lib.rs:378-557 ``get_mir_instructions``,
:560-765 ``get_init_mir_instructions``, :766-912 ``get_dae_system``).  That crate, the
OpenVAF submodule and ``psp103.va`` are absent from the build environment (SURVEY.md
§8c), so the dicts the CUDA-C emitter consumes have to come from somewhere else.  This
module emits dicts in **exactly that schema** from a small imperative DSL with
forward-mode automatic differentiation (OpenVAF does the same job in its
``sim_back`` crate: residual + Jacobian SSA values with ``mir_<N>`` references).

It produces the same top-level bundle the reference pickles as ``mir_data.pkl``
(reference: openvaf_jax/__init__.py:196-222 ``get_cache_data``), so a real PSP103 bundle
can be dropped in without touching the emitter.

This is setup-time tooling; nothing here runs inside the transient loop.
"""

from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

U32_MAX = 4294967295


class Val:
    """An SSA value id plus its static type ('f', 'i', 'b', 's')."""

    __slots__ = ("vid", "ty")

    def __init__(self, vid: str, ty: str = "f"):
        self.vid = vid
        self.ty = ty

    def __repr__(self):
        return f"Val({self.vid}:{self.ty})"


class FuncBuilder:
    """Builds one MIR function (``init`` or ``eval``) instruction by instruction.

    Pre-allocated constants follow OpenVAF (reference: openvaf_jax/mir/types.py:21-28):
    v0 gravestone, v1 false, v2 true, v3 0.0, v4 0, v5 1, v6 1.0, v7 -1.0.
    """

    def __init__(self, name: str):
        self.name = name
        self._next = 8
        self.constants: Dict[str, float] = {}
        self.bool_constants: Dict[str, bool] = {}
        self.int_constants: Dict[str, int] = {}
        self.str_constants_raw: Dict[str, int] = {}
        self.str_values: Dict[str, str] = {}
        self._fconst: Dict[float, str] = {0.0: "v3", 1.0: "v6", -1.0: "v7"}
        self._iconst: Dict[int, str] = {0: "v4", 1: "v5"}
        self._sconst: Dict[str, str] = {}
        self.instructions: List[dict] = []
        self.blocks: Dict[str, dict] = {"block0": {"predecessors": [], "successors": []}}
        self.cur = "block0"
        self._nblocks = 1
        self.params: List[Tuple[str, str, str]] = []  # (vid, kind, name)
        self.function_decls: Dict[str, dict] = {}
        self._decl_by_name: Dict[str, str] = {}

    # -- values ---------------------------------------------------------
    def _new(self, ty: str = "f") -> Val:
        v = Val(f"v{self._next}", ty)
        self._next += 1
        return v

    def fconst(self, x: float) -> Val:
        x = float(x)
        key = x
        if x != x:
            key = "nan"  # type: ignore[assignment]
        if key not in self._fconst:
            v = self._new("f")
            self._fconst[key] = v.vid
            self.constants[v.vid] = x
        return Val(self._fconst[key], "f")

    def iconst(self, x: int) -> Val:
        if x not in self._iconst:
            v = self._new("i")
            self._iconst[x] = v.vid
            self.int_constants[v.vid] = int(x)
        return Val(self._iconst[x], "i")

    def bconst(self, x: bool) -> Val:
        return Val("v2" if x else "v1", "b")

    def sconst(self, s: str) -> Val:
        if s not in self._sconst:
            v = self._new("s")
            self._sconst[s] = v.vid
            self.str_constants_raw[v.vid] = len(self._sconst)
            self.str_values[v.vid] = s
        return Val(self._sconst[s], "s")

    def param(self, kind: str, name: str, ty: str = "f") -> Val:
        v = self._new(ty)
        self.params.append((v.vid, kind, name))
        return v

    # -- instructions ---------------------------------------------------
    def emit(self, opcode: str, operands: Sequence[Val], ty: str = "f") -> Val:
        r = self._new(ty)
        self.instructions.append(
            {
                "block": self.cur,
                "result": r.vid,
                "opcode": opcode,
                "operands": [o.vid for o in operands],
            }
        )
        return r

    def call(self, func_name: str, operands: Sequence[Val], ty: str = "f", num_returns: int = 1) -> Val:
        if func_name not in self._decl_by_name:
            ref = f"inst{len(self.function_decls)}"
            self.function_decls[ref] = {
                "name": func_name,
                "num_args": len(operands),
                "num_returns": num_returns,
            }
            self._decl_by_name[func_name] = ref
        r = self._new(ty)
        self.instructions.append(
            {
                "block": self.cur,
                "result": r.vid,
                "opcode": "call",
                "operands": [o.vid for o in operands],
                "func_ref": self._decl_by_name[func_name],
            }
        )
        return r

    # -- control flow ---------------------------------------------------
    def new_block(self) -> str:
        name = f"block{self._nblocks}"
        self._nblocks += 1
        self.blocks[name] = {"predecessors": [], "successors": []}
        return name

    def _edge(self, a: str, b: str):
        self.blocks[a]["successors"].append(b)
        if a not in self.blocks[b]["predecessors"]:
            self.blocks[b]["predecessors"].append(a)

    def br(self, cond: Val, tb: str, fb: str):
        self.instructions.append(
            {
                "block": self.cur,
                "opcode": "br",
                "condition": cond.vid,
                "true_block": tb,
                "false_block": fb,
                "loop_entry": False,
            }
        )
        self._edge(self.cur, tb)
        self._edge(self.cur, fb)

    def jmp(self, dst: str):
        self.instructions.append({"block": self.cur, "opcode": "jmp", "destination": dst})
        self._edge(self.cur, dst)

    def exit(self):
        self.instructions.append({"block": self.cur, "opcode": "exit"})

    def phi(self, edges: Sequence[Tuple[str, Val]], ty: str = "f") -> Val:
        r = self._new(ty)
        self.instructions.append(
            {
                "block": self.cur,
                "result": r.vid,
                "opcode": "phi",
                "phi_operands": [{"value": v.vid, "block": b} for b, v in edges],
            }
        )
        return r

    def to_dict(self) -> dict:
        # lib.rs emits instructions grouped by block in layout order; keep that.
        order = {b: i for i, b in enumerate(self.blocks)}
        insts = sorted(
            enumerate(self.instructions), key=lambda t: (order[t[1]["block"]], t[0])
        )
        return {
            "constants": dict(self.constants),
            "bool_constants": dict(self.bool_constants),
            "int_constants": dict(self.int_constants),
            "str_constants": dict(self.str_constants_raw),
            "params": [p[0] for p in self.params],
            "instructions": [i for _, i in insts],
            "blocks": {k: {"predecessors": list(v["predecessors"]), "successors": list(v["successors"])} for k, v in self.blocks.items()},
            "function_decls": {k: dict(v) for k, v in self.function_decls.items()},
        }


class Dual:
    """A float SSA value with its partial derivatives w.r.t. node voltages.

    ``d`` maps a model node name (DAE unknown) to the SSA value of d(self)/dV(node);
    absent keys are structurally zero.
    """

    __slots__ = ("fb", "v", "d")

    def __init__(self, fb: FuncBuilder, v: Val, d: Optional[Dict[str, Val]] = None):
        self.fb = fb
        self.v = v
        self.d = d or {}

    # helpers
    def _lift(self, o) -> "Dual":
        if isinstance(o, Dual):
            return o
        return Dual(self.fb, self.fb.fconst(float(o)))

    def _lin(self, o: "Dual", sa: float, sb: float) -> Dict[str, Val]:
        fb = self.fb
        out: Dict[str, Val] = {}
        for k in list(self.d) + [k for k in o.d if k not in self.d]:
            a = self.d.get(k)
            b = o.d.get(k)
            if a is not None and b is not None:
                out[k] = fb.emit("fadd" if sb > 0 else "fsub", [a, b])
            elif a is not None:
                out[k] = a
            else:
                out[k] = b if sb > 0 else fb.emit("fneg", [b])
        return out

    def __add__(self, o):
        o = self._lift(o)
        return Dual(self.fb, self.fb.emit("fadd", [self.v, o.v]), self._lin(o, 1, 1))

    __radd__ = __add__

    def __sub__(self, o):
        o = self._lift(o)
        return Dual(self.fb, self.fb.emit("fsub", [self.v, o.v]), self._lin(o, 1, -1))

    def __rsub__(self, o):
        return self._lift(o).__sub__(self)

    def __neg__(self):
        fb = self.fb
        return Dual(fb, fb.emit("fneg", [self.v]), {k: fb.emit("fneg", [v]) for k, v in self.d.items()})

    def __mul__(self, o):
        o = self._lift(o)
        fb = self.fb
        r = fb.emit("fmul", [self.v, o.v])
        d: Dict[str, Val] = {}
        for k in list(self.d) + [k for k in o.d if k not in self.d]:
            terms = []
            if k in self.d:
                terms.append(fb.emit("fmul", [self.d[k], o.v]))
            if k in o.d:
                terms.append(fb.emit("fmul", [self.v, o.d[k]]))
            d[k] = terms[0] if len(terms) == 1 else fb.emit("fadd", terms)
        return Dual(fb, r, d)

    __rmul__ = __mul__

    def __truediv__(self, o):
        o = self._lift(o)
        fb = self.fb
        r = fb.emit("fdiv", [self.v, o.v])
        d: Dict[str, Val] = {}
        for k in list(self.d) + [k for k in o.d if k not in self.d]:
            # d(a/b) = a'/b - (a/b) * b'/b
            terms = None
            if k in self.d:
                terms = fb.emit("fdiv", [self.d[k], o.v])
            if k in o.d:
                t = fb.emit("fdiv", [fb.emit("fmul", [r, o.d[k]]), o.v])
                terms = fb.emit("fsub", [terms, t]) if terms is not None else fb.emit("fneg", [t])
            d[k] = terms
        return Dual(fb, r, d)

    def __rtruediv__(self, o):
        return self._lift(o).__truediv__(self)

    def _chain(self, r: Val, dr: Val) -> "Dual":
        fb = self.fb
        return Dual(fb, r, {k: fb.emit("fmul", [dr, v]) for k, v in self.d.items()})


def _unary(op: str):
    def f(x: Dual) -> Dual:
        fb = x.fb
        r = fb.emit(op, [x.v])
        if not x.d or op in ("floor", "ceil", "clog2"):
            return Dual(fb, r)
        one, e, xv = fb.fconst(1.0), fb.emit, x.v
        if op == "exp":
            dr = r
        elif op == "ln":
            dr = e("fdiv", [one, xv])
        elif op == "log":  # log10
            dr = e("fdiv", [fb.fconst(0.4342944819032518), xv])
        elif op == "sqrt":
            dr = e("fdiv", [fb.fconst(0.5), r])
        elif op == "tanh":
            dr = e("fsub", [one, e("fmul", [r, r])])
        elif op == "tan":
            dr = e("fadd", [one, e("fmul", [r, r])])
        elif op == "sin":
            dr = e("cos", [xv])
        elif op == "cos":
            dr = e("fneg", [e("sin", [xv])])
        elif op == "sinh":
            dr = e("cosh", [xv])
        elif op == "cosh":
            dr = e("sinh", [xv])
        elif op == "atan":
            dr = e("fdiv", [one, e("fadd", [one, e("fmul", [xv, xv])])])
        elif op == "asin":
            dr = e("fdiv", [one, e("sqrt", [e("fsub", [one, e("fmul", [xv, xv])])])])
        elif op == "acos":
            dr = e("fneg", [e("fdiv", [one, e("sqrt", [e("fsub", [one, e("fmul", [xv, xv])])])])])
        elif op == "asinh":
            dr = e("fdiv", [one, e("sqrt", [e("fadd", [e("fmul", [xv, xv]), one])])])
        elif op == "acosh":
            dr = e("fdiv", [one, e("sqrt", [e("fsub", [e("fmul", [xv, xv]), one])])])
        elif op == "atanh":
            dr = e("fdiv", [one, e("fsub", [one, e("fmul", [xv, xv])])])
        elif op in ("abs", "fabs"):
            dr = e("select", [e("flt", [xv, fb.fconst(0.0)], ty="b"), fb.fconst(-1.0), one])
        else:
            raise NotImplementedError(op)
        return x._chain(r, dr)

    return f


exp = _unary("exp")
ln = _unary("ln")
log10 = _unary("log")
sqrt = _unary("sqrt")
tanh = _unary("tanh")
tan = _unary("tan")
sin = _unary("sin")
cos = _unary("cos")
sinh = _unary("sinh")
cosh = _unary("cosh")
atan = _unary("atan")
asin = _unary("asin")
acos = _unary("acos")
asinh = _unary("asinh")
acosh = _unary("acosh")
atanh = _unary("atanh")
fabs = _unary("abs")
floor = _unary("floor")
ceil = _unary("ceil")
clog2 = _unary("clog2")


def _binary(op: str):
    def f(a, b) -> Dual:
        if not isinstance(a, Dual):
            a = b._lift(a)
        b = a._lift(b)
        fb = a.fb
        e = fb.emit
        r = e(op, [a.v, b.v])
        if not a.d and not b.d:
            return Dual(fb, r)
        if op == "pow":
            da = e("fmul", [b.v, e("pow", [a.v, e("fsub", [b.v, fb.fconst(1.0)])])]) if a.d else None
            db = e("fmul", [r, e("ln", [a.v])]) if b.d else None
        elif op == "atan2":  # atan2(y = a, x = b)
            den = e("fadd", [e("fmul", [a.v, a.v]), e("fmul", [b.v, b.v])])
            da = e("fdiv", [b.v, den]) if a.d else None
            db = e("fneg", [e("fdiv", [a.v, den])]) if b.d else None
        elif op == "hypot":
            da = e("fdiv", [a.v, r]) if a.d else None
            db = e("fdiv", [b.v, r]) if b.d else None
        elif op in ("fmin", "fmax"):
            c = e("flt" if op == "fmin" else "fgt", [a.v, b.v], ty="b")
            zero = fb.fconst(0.0)
            d = {}
            for k in list(a.d) + [k for k in b.d if k not in a.d]:
                d[k] = e("select", [c, a.d.get(k, zero), b.d.get(k, zero)])
            return Dual(fb, r, d)
        else:
            raise NotImplementedError(op)
        d = {}
        for k in list(a.d) + [k for k in b.d if k not in a.d]:
            terms = []
            if k in a.d:
                terms.append(e("fmul", [da, a.d[k]]))
            if k in b.d:
                terms.append(e("fmul", [db, b.d[k]]))
            d[k] = terms[0] if len(terms) == 1 else e("fadd", terms)
        return Dual(fb, r, d)

    return f


power = _binary("pow")
atan2 = _binary("atan2")
hypot = _binary("hypot")
fmin = _binary("fmin")
fmax = _binary("fmax")


def cmp(op: str, a: Dual, b) -> Val:
    b = a._lift(b)
    return a.fb.emit(op, [a.v, b.v], ty="b")


def select(c: Val, a: Dual, b) -> Dual:
    b = a._lift(b)
    fb = a.fb
    r = fb.emit("select", [c, a.v, b.v])
    d = {}
    zero = fb.fconst(0.0)
    for k in list(a.d) + [k for k in b.d if k not in a.d]:
        d[k] = fb.emit("select", [c, a.d.get(k, zero), b.d.get(k, zero)])
    return Dual(fb, r, d)


class IfElse:
    """Structured if/else producing real MIR blocks and PHI nodes.

    ``shape='diamond'``: br -> then / else -> merge (both arms are blocks).
    ``shape='triangle'``: br cond, then, merge (no else block; PHI takes the
    pre-branch value on the fall-through edge) - OpenVAF emits both shapes.
    """

    def __init__(self, fb: FuncBuilder, cond: Val, shape: str = "diamond"):
        self.fb = fb
        self.shape = shape
        self.head = fb.cur
        self.then_b = fb.new_block()
        self.else_b = fb.new_block() if shape == "diamond" else None
        self.merge_b = fb.new_block()
        fb.br(cond, self.then_b, self.else_b or self.merge_b)
        self.then_end = None
        self.else_end = None
        fb.cur = self.then_b

    def otherwise(self):
        fb = self.fb
        self.then_end = fb.cur
        fb.jmp(self.merge_b)
        assert self.else_b is not None
        fb.cur = self.else_b

    def end(self):
        fb = self.fb
        if self.shape == "diamond":
            self.else_end = fb.cur
            fb.jmp(self.merge_b)
        else:
            self.then_end = fb.cur
            fb.jmp(self.merge_b)
            self.else_end = self.head
        fb.cur = self.merge_b

    def phi(self, a: Dual, b) -> Dual:
        """Merge ``a`` (then arm) and ``b`` (else arm / fall-through) after end()."""
        fb = self.fb
        b = a._lift(b)
        r = fb.phi([(self.then_end, a.v), (self.else_end, b.v)])
        zero = fb.fconst(0.0)
        d = {}
        for k in list(a.d) + [k for k in b.d if k not in a.d]:
            d[k] = fb.phi([(self.then_end, a.d.get(k, zero)), (self.else_end, b.d.get(k, zero))])
        return Dual(fb, r, d)


class ModelBuilder:
    """Builds the whole per-model bundle (init MIR + eval MIR + DAE system + metadata)."""

    def __init__(self, name: str, terminals: Sequence[str], internal: Sequence[str] = (),
                 internal_kinds: Optional[Dict[str, str]] = None):
        self.name = name
        self.terminals = list(terminals)
        self.internal = list(internal)
        self.node_names = self.terminals + self.internal
        self.internal_kinds = internal_kinds or {}
        self.init = FuncBuilder("init")
        self.eval = FuncBuilder("eval")
        self.param_defaults: Dict[str, float] = {}
        self.osdi_params: List[dict] = []
        self._cache: List[Tuple[str, Val]] = []  # (init value id, eval param Val)
        self._resist: Dict[str, List[Tuple[float, Dual]]] = {n: [] for n in self.node_names}
        self._react: Dict[str, List[Tuple[float, Dual]]] = {n: [] for n in self.node_names}
        self.collapsible_pairs: List[Tuple[int, int]] = []
        self.collapse_decision_outputs: List[Tuple[int, str]] = []
        self._init_params: Dict[str, Val] = {}
        self._eval_named: Dict[Tuple[str, str], Val] = {}
        self.num_states = 0

    # -- parameters -----------------------------------------------------
    def model_param(self, name: str, default: float, is_instance: bool = False) -> Tuple[Dual, Dual]:
        """Declare a real parameter visible in both init and eval. Returns (init, eval) duals."""
        self.param_defaults[name.lower()] = float(default)
        self.osdi_params.append(
            {"name": name, "aliases": [], "units": "", "description": "", "flags": (4 if is_instance else 0),
             "is_instance": is_instance, "is_model_param": not is_instance}
        )
        iv = self.init.param("param", name)
        ev = self.eval.param("param", name)
        self._init_params[name] = iv
        self._eval_named[("param", name)] = ev
        return Dual(self.init, iv), Dual(self.eval, ev)

    def param_given(self, name: str) -> Tuple[Val, Val]:
        iv = self.init.param("param_given", name, ty="b")
        ev = self.eval.param("param_given", name, ty="b")
        return iv, ev

    def temperature(self) -> Tuple[Dual, Dual]:
        iv = self.init.param("temperature", "$temperature")
        ev = self.eval.param("temperature", "$temperature")
        return Dual(self.init, iv), Dual(self.eval, ev)

    def mfactor(self) -> Dual:
        return Dual(self.eval, self.eval.param("sysfun", "mfactor"))

    def voltage(self, hi: str, lo: Optional[str] = None) -> Dual:
        """Branch voltage V(hi,lo) as an eval parameter with unit derivatives."""
        name = f"V({hi},{lo})" if lo else f"V({hi})"
        key = ("voltage", name)
        if key not in self._eval_named:
            self._eval_named[key] = self.eval.param("voltage", name)
        v = self._eval_named[key]
        d = {hi: self.eval.fconst(1.0)}
        if lo:
            d[lo] = self.eval.fconst(-1.0)
        return Dual(self.eval, v, d)

    def limit_params(self, state: int) -> Tuple[Dual, Val]:
        """(prev_state_N value, enable_lim flag) eval params for a $limit call."""
        kp = ("prev_state", f"prev_state_{state}")
        if kp not in self._eval_named:
            self._eval_named[kp] = self.eval.param("prev_state", f"prev_state_{state}")
            self.eval.param("new_state", f"new_state_{state}")
            self.num_states = max(self.num_states, state + 1)
        ke = ("enable_lim", "enable_lim")
        if ke not in self._eval_named:
            self._eval_named[ke] = self.eval.param("enable_lim", "enable_lim", ty="b")
        return Dual(self.eval, self._eval_named[kp]), self._eval_named[ke]

    def cache(self, init_value: Dual) -> Dual:
        """Pass a value computed by init to eval through a cache slot."""
        assert init_value.fb is self.init
        ev = Val(f"v{self.eval._next}", "f")
        self.eval._next += 1
        self._cache.append((init_value.v.vid, ev))
        return Dual(self.eval, ev)

    # -- contributions --------------------------------------------------
    def contrib(self, hi: str, lo: Optional[str], resist: Optional[Dual] = None, react: Optional[Dual] = None):
        """I(hi,lo) <+ resist + ddt(react)."""
        for tgt, val in ((self._resist, resist), (self._react, react)):
            if val is None:
                continue
            tgt[hi].append((1.0, val))
            if lo is not None:
                tgt[lo].append((-1.0, val))

    def collapsible(self, n1: str, n2: Optional[str], decision_init: Optional[Val] = None, negate: bool = False):
        """Declare that n1 may collapse into n2 (None = ground) when decision_init is true."""
        i1 = self.node_names.index(n1)
        i2 = U32_MAX if n2 is None else self.node_names.index(n2)
        idx = len(self.collapsible_pairs)
        self.collapsible_pairs.append((i1, i2))
        if decision_init is not None:
            self.collapse_decision_outputs.append((idx, ("!" if negate else "") + decision_init.vid))

    # -- finalize -------------------------------------------------------
    def _sum(self, fb: FuncBuilder, terms: List[Tuple[float, Val]]) -> Optional[Val]:
        acc: Optional[Val] = None
        for s, v in terms:
            if acc is None:
                acc = v if s > 0 else fb.emit("fneg", [v])
            else:
                acc = fb.emit("fadd" if s > 0 else "fsub", [acc, v])
        return acc

    def finalize(self) -> dict:
        ev, it = self.eval, self.init
        # cache params come after the named params (their eval_param index >= n_named)
        n_named = len(ev.params)
        cache_mapping = []
        for k, (ivid, eparam) in enumerate(self._cache):
            ev.params.append((eparam.vid, "hidden_state", f"cache_{k}"))
            cache_mapping.append({"init_value": ivid, "eval_param": n_named + k})

        residuals = []
        for i, n in enumerate(self.node_names):
            r = self._sum(ev, [(s, d.v) for s, d in self._resist[n]])
            q = self._sum(ev, [(s, d.v) for s, d in self._react[n]])
            residuals.append(
                {
                    "equation_idx": i,
                    "node_idx": i,
                    "node_name": n,
                    "resist_var": "mir_" + (r.vid[1:] if r else "3"),
                    "react_var": "mir_" + (q.vid[1:] if q else "3"),
                    "resist_lim_rhs_var": "mir_3",
                    "react_lim_rhs_var": "mir_3",
                    "resist_small_signal_var": "mir_3",
                    "react_small_signal_var": "mir_3",
                }
            )
        jac = []
        for i, rn in enumerate(self.node_names):
            for j, cn in enumerate(self.node_names):
                rt = [(s, d.d[cn]) for s, d in self._resist[rn] if cn in d.d]
                qt = [(s, d.d[cn]) for s, d in self._react[rn] if cn in d.d]
                if not rt and not qt:
                    continue
                r = self._sum(ev, rt)
                q = self._sum(ev, qt)
                jac.append(
                    {
                        "entry_idx": len(jac),
                        "row_node_idx": i,
                        "col_node_idx": j,
                        "row_node_name": rn,
                        "col_node_name": cn,
                        "resist_var": "mir_" + (r.vid[1:] if r else "3"),
                        "react_var": "mir_" + (q.vid[1:] if q else "3"),
                        "has_resist": r is not None,
                        "has_react": q is not None,
                    }
                )
        ev.exit()
        it.exit()

        nodes = []
        for i, n in enumerate(self.node_names):
            nodes.append(
                {
                    "idx": i,
                    "name": n,
                    "kind": self.internal_kinds.get(n, "KirchoffLaw"),
                    "is_internal": i >= len(self.terminals),
                }
            )
        cp = []
        for i, (a, b) in enumerate(self.collapsible_pairs):
            e = {
                "pair_idx": i,
                "node1_idx": a,
                "node2_idx": b,
                "node1_name": self.node_names[a],
                "node2_name": "ground" if b == U32_MAX else self.node_names[b],
            }
            for idx, dv in self.collapse_decision_outputs:
                if idx == i:
                    e["decision_var"] = dv
            cp.append(e)
        dae = {
            "nodes": nodes,
            "residuals": residuals,
            "jacobian": jac,
            "terminals": list(self.terminals),
            "internal_nodes": list(self.internal),
            "num_terminals": len(self.terminals),
            "num_internal": len(self.internal),
            "collapsible_pairs": cp,
            "num_collapsible": len(cp),
            "small_signal_params": [],
            "num_small_signal_params": 0,
        }
        mir = ev.to_dict()
        init_mir = it.to_dict()
        init_mir["cache_mapping"] = cache_mapping
        str_constants = dict(ev.str_values)
        str_constants.update(it.str_values)
        # module.nodes is the Debug print of each SimUnknownKind: "node<i>" for KCL
        # unknowns, "br[...]" style for branch-current unknowns.
        mod_nodes = []
        for i, n in enumerate(self.node_names):
            k = self.internal_kinds.get(n, "KirchoffLaw")
            mod_nodes.append(f"node{i}" if k == "KirchoffLaw" else f"br[Branch(BranchId({i}))]")
        return {
            "mir_data": mir,
            "init_mir_data": init_mir,
            "dae_data": dae,
            "str_constants": str_constants,
            "param_names": [p[2] for p in ev.params[:n_named]],
            "param_kinds": [p[1] for p in ev.params[:n_named]],
            "nodes": mod_nodes,
            "collapse_decision_outputs": list(self.collapse_decision_outputs),
            "collapsible_pairs": list(self.collapsible_pairs),
            "num_collapsible": len(self.collapsible_pairs),
            "all_func_params": [(i, int(p[0][1:])) for i, p in enumerate(ev.params)],
            "init_param_names": [p[2] for p in it.params],
            "init_param_kinds": [p[1] for p in it.params],
            "param_defaults": dict(self.param_defaults),
            "osdi_descriptor": {
                "params": list(self.osdi_params),
                "nodes": [{"name": n, "units": "V", "residual_units": "A", "is_internal": i >= len(self.terminals)}
                          for i, n in enumerate(self.node_names)],
                "num_terminals": len(self.terminals),
                "num_states": self.num_states,
                "num_nodes": len(self.node_names),
            },
            "model_name": self.name,
        }
