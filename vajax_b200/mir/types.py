"""Parsed view of the OpenVAF MIR dict schema (the contract).

Producer schema: reference openvaf_jax/openvaf_py/src/lib.rs:378-557 (eval),
:560-765 (init, adds ``cache_mapping``), :766-912 (DAE system).  Consumer in the reference:
openvaf_jax/mir/types.py:175-297.  This module is an independent reader of the same dicts,
used by the CUDA-C emitter; it adds static type inference (the reference relies on Python's
dynamic typing) and a topological block order (the reference's CFGAnalyzer).
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

PREALLOC_F = {"v3": 0.0, "v6": 1.0, "v7": -1.0}
PREALLOC_I = {"v4": 0, "v5": 1}
PREALLOC_B = {"v1": False, "v2": True}


@dataclass
class Inst:
    opcode: str
    block: str
    result: Optional[str] = None
    operands: List[str] = field(default_factory=list)
    phi: List[Tuple[str, str]] = field(default_factory=list)  # (pred block, value)
    cond: Optional[str] = None
    true_block: Optional[str] = None
    false_block: Optional[str] = None
    target: Optional[str] = None
    func_name: str = ""


@dataclass
class Function:
    name: str
    params: List[str]
    fconst: Dict[str, float]
    iconst: Dict[str, int]
    bconst: Dict[str, bool]
    sconst: Dict[str, str]
    blocks: Dict[str, List[Inst]]
    preds: Dict[str, List[str]]
    succs: Dict[str, List[str]]
    entry: str = "block0"
    cache_mapping: List[dict] = field(default_factory=list)

    def topo_order(self) -> List[str]:
        """Reverse post-order from the entry block. Raises on loops: the reference casts
        loop-carried state to float32 (function_builder.py:301-304,442-453), a quirk we do
        not copy; PSP103/resistor/capacitor/diode eval have no loops."""
        seen: Dict[str, int] = {}
        order: List[str] = []
        stack = [(self.entry, iter(self._succ_of(self.entry)))]
        seen[self.entry] = 1
        while stack:
            node, it = stack[-1]
            adv = False
            for s in it:
                if s not in seen:
                    seen[s] = 1
                    stack.append((s, iter(self._succ_of(s))))
                    adv = True
                    break
                elif seen[s] == 1:
                    raise NotImplementedError(f"MIR loop through {s}: not supported by the CUDA emitter")
            if not adv:
                seen[node] = 2
                order.append(node)
                stack.pop()
        order.reverse()
        return order

    def _succ_of(self, b: str) -> List[str]:
        term = self.blocks[b][-1] if self.blocks[b] else None
        if term is None:
            return []
        if term.opcode == "br":
            return [term.true_block, term.false_block]
        if term.opcode == "jmp":
            return [term.target]
        return []


def parse_function(name: str, mir: dict, str_constants: Optional[Dict[str, str]] = None) -> Function:
    decls = dict(mir.get("function_decls", {}))
    fconst = dict(PREALLOC_F)
    fconst.update({k: float(v) for k, v in mir.get("constants", {}).items()})
    iconst = dict(PREALLOC_I)
    iconst.update({k: int(v) for k, v in mir.get("int_constants", {}).items()})
    bconst = dict(PREALLOC_B)
    bconst.update({k: bool(v) for k, v in mir.get("bool_constants", {}).items()})
    blocks: Dict[str, List[Inst]] = {b: [] for b in mir.get("blocks", {})}
    for d in mir.get("instructions", []):
        op = d.get("opcode", "").lower()
        ins = Inst(opcode=op, block=d.get("block", ""), result=d.get("result"), operands=list(d.get("operands", [])))
        if op == "phi":
            ins.phi = [(e["block"], e["value"]) for e in d.get("phi_operands", [])]
        elif op == "br":
            ins.cond = d.get("condition")
            ins.true_block = d.get("true_block")
            ins.false_block = d.get("false_block")
        elif op == "jmp":
            ins.target = d.get("destination") or d.get("target_block")
        elif op == "call":
            fn = d.get("func_name")
            if not fn:
                ref = d.get("func_ref")
                if ref in decls:
                    fn = decls[ref].get("name")
            ins.func_name = fn or ""
        blocks.setdefault(ins.block, []).append(ins)
    preds = {b: list(v.get("predecessors", [])) for b, v in mir.get("blocks", {}).items()}
    succs = {b: list(v.get("successors", [])) for b, v in mir.get("blocks", {}).items()}
    entry = "block0" if "block0" in blocks else next((b for b in blocks if not preds.get(b)), "block0")
    return Function(
        name=name,
        params=list(mir.get("params", [])),
        fconst=fconst,
        iconst=iconst,
        bconst=bconst,
        sconst=dict(str_constants or {}),
        blocks=blocks,
        preds=preds,
        succs=succs,
        entry=entry,
        cache_mapping=list(mir.get("cache_mapping", [])),
    )


ANALYSIS_TYPE_MAP = {"dc": 0, "static": 0, "ac": 1, "tran": 2, "transient": 2, "noise": 3, "nodeset": 4}


def classify_call(func_name: str) -> str:
    """Map a callback's Debug name to the behaviour class the reference implements
    (same substring tests, same order: openvaf_jax/codegen/instruction.py:469-703)."""
    fl = (func_name or "").lower()
    if not fl:
        return "zero"
    if "simparam" in fl:
        return "simparam"
    if "discontinuity" in fl:
        return "zero"
    if any(x in fl for x in ("display", "strobe", "write", "debug")):
        return "zero"
    if "analysis" in fl:
        return "analysis"
    if "ddt" in fl:
        return "passthrough"
    if "ddx" in fl:
        return "zero"
    if any(x in fl for x in ("white_noise", "flicker_noise", "noise_table")):
        return "zero"
    if fl == "abs":
        return "abs"
    if fl == "min":
        return "min"
    if fl == "max":
        return "max"
    if "storelimit" in fl:
        return "storelimit"
    if "loadlimit" in fl:
        return "loadlimit"
    if "$limit" in fl or "builtinlimit" in fl:
        return "builtinlimit"
    return "zero"
