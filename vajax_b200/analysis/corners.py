"""PVT corner / Monte-Carlo sweep definitions and the multi-GPU sweep driver.

Mirrors the reference's corner vocabulary (reference: vajax/analysis/corners.py:39-376 -
``ProcessCorner``, ``VoltageCorner``, ``CornerConfig``, ``PROCESS_CORNERS``,
``TEMPERATURE_CORNERS``, ``create_standard_corners``, ``create_pvt_corners``,
``apply_process_corner``, ``apply_voltage_corner``) and replaces the serial loop of
``CircuitEngine.run_corners`` (reference: vajax/analysis/engine.py:2235-2320) by whole-circuit
replicas: corner ``i`` runs on GPU ``i mod world_size``; nothing is exchanged during a run and the
recorded waveforms are gathered once at the end.

Parity target: "fresh engine per corner".  The reference re-uses a topology-keyed setup cache, so
same-temperature process/voltage edits are silently dropped there; here every
corner gets its own parameter / init-cache upload.
"""

from __future__ import annotations

import os

import copy
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence

import numpy as np

DEFAULT_TEMPERATURE_K = 300.15  # reference: vajax/config.py


@dataclass
class ProcessCorner:
    name: str
    mobility_scale: float = 1.0
    vth_shift: float = 0.0
    tox_scale: float = 1.0
    length_delta: float = 0.0
    model_params: Dict[str, Dict[str, float]] = field(default_factory=dict)


@dataclass
class VoltageCorner:
    name: str
    vdd_scale: float = 1.0
    source_values: Dict[str, float] = field(default_factory=dict)


@dataclass
class CornerConfig:
    name: str
    process: Optional[ProcessCorner] = None
    voltage: Optional[VoltageCorner] = None
    temperature: float = DEFAULT_TEMPERATURE_K


@dataclass
class CornerResult:
    corner: CornerConfig
    times: Any = None
    voltages: Any = None
    stats: Dict[str, Any] = field(default_factory=dict)
    converged: bool = True


@dataclass
class CornerSweepResult:
    corners: List[CornerConfig] = field(default_factory=list)
    results: List[CornerResult] = field(default_factory=list)

    @property
    def num_corners(self) -> int:
        return len(self.results)

    @property
    def num_converged(self) -> int:
        return sum(1 for r in self.results if r.converged)

    @property
    def all_converged(self) -> bool:
        return all(r.converged for r in self.results)

    def get_result(self, corner_name: str) -> Optional[CornerResult]:
        return next((r for r in self.results if r.corner.name == corner_name), None)

    def converged_results(self) -> List[CornerResult]:
        return [r for r in self.results if r.converged]


# values: reference corners.py:162-196
PROCESS_CORNERS: Dict[str, ProcessCorner] = {
    "FF": ProcessCorner("FF", mobility_scale=1.15, vth_shift=-0.05, tox_scale=0.95),
    "TT": ProcessCorner("TT"),
    "SS": ProcessCorner("SS", mobility_scale=0.85, vth_shift=0.05, tox_scale=1.05),
    "FS": ProcessCorner("FS", mobility_scale=1.15, vth_shift=0.05),
    "SF": ProcessCorner("SF", mobility_scale=0.85, vth_shift=-0.05),
}
TEMPERATURE_CORNERS: Dict[str, float] = {"cold": 233.15, "room": 300.15, "hot": 398.15}

# Parameter names touched per knob.  The first group of each tuple is the reference's list
# (corners.py:283-300); the names after the bar belong to the PSP103-shaped fixture model
# (vajax_b200/models/fixtures.py: kp / vto) and to psp103w (PSP's global names uoo / vfbo; toxo is in the reference's list).
MOBILITY_PARAMS = ("uo", "mu0", "u0", "betn", "betp", "mue") + ("kp", "uoo")
VTH_PARAMS = ("vth0", "vfb", "delvto", "dvt0") + ("vto", "vfbo")
TOX_PARAMS = ("tox", "toxe", "toxo", "toxp")
COX_PARAMS = ("cox",)  # cox ~ 1/tox


def _temp_tag(temp_k: float) -> str:
    c = temp_k - 273.15
    return {-40.0: "m40C", 27.0: "27C", 125.0: "125C"}.get(round(c, 6), f"{c:.0f}C")


def create_standard_corners(processes: Optional[Sequence[str]] = None, temperatures: Optional[Sequence[float]] = None,
                            vdd_scales: Optional[Sequence[float]] = None) -> List[CornerConfig]:
    """All process x temperature x vdd combinations, named ``<proc>_<temp>_V<pct>pct``."""
    out: List[CornerConfig] = []
    for pn in (processes or ["TT"]):
        if pn not in PROCESS_CORNERS:
            raise ValueError(f"Unknown process corner: {pn}")
        for t in (temperatures or [DEFAULT_TEMPERATURE_K]):
            for v in (vdd_scales or [1.0]):
                out.append(CornerConfig(name=f"{pn}_{_temp_tag(t)}_V{v * 100:.0f}pct", process=PROCESS_CORNERS[pn],
                                        voltage=VoltageCorner(name=f"VDD_{v}", vdd_scale=v), temperature=float(t)))
    return out


def create_pvt_corners(processes: Optional[Sequence[str]] = None, temperatures: Optional[Sequence[Any]] = None,
                       voltages: Optional[Sequence[float]] = None) -> List[CornerConfig]:
    """Default 3 x 3 x 3 = 27 corners (FF/TT/SS, cold/room/hot, 0.9/1.0/1.1)."""
    temps = [TEMPERATURE_CORNERS[t] if t in TEMPERATURE_CORNERS else float(t)
             for t in (temperatures or ["cold", "room", "hot"])]
    return create_standard_corners(list(processes or ["FF", "TT", "SS"]), temps, list(voltages or [0.9, 1.0, 1.1]))


def apply_process_corner(devices: List[Dict[str, Any]], corner: Optional[ProcessCorner]) -> None:
    """In-place edit of the OpenVAF devices' parameters."""
    if corner is None:
        return
    for dev in devices:
        if not dev.get("is_openvaf", False):
            continue
        p = dev.setdefault("params", {})
        for k in MOBILITY_PARAMS:
            if k in p:
                p[k] = float(p[k]) * corner.mobility_scale
        for k in VTH_PARAMS:
            if k in p:
                p[k] = float(p[k]) + corner.vth_shift
        for k in TOX_PARAMS:
            if k in p:
                p[k] = float(p[k]) * corner.tox_scale
        for k in COX_PARAMS:
            if k in p:
                p[k] = float(p[k]) / corner.tox_scale
        if corner.length_delta != 0 and "l" in p:
            p["l"] = float(p["l"]) + corner.length_delta
        for k, v in corner.model_params.get(dev.get("model", ""), {}).items():
            p[k] = v


def apply_voltage_corner(devices: List[Dict[str, Any]], corner: Optional[VoltageCorner]) -> None:
    """In-place edit of the voltage sources' DC values."""
    if corner is None:
        return
    for dev in devices:
        if dev.get("model") != "vsource":
            continue
        p = dev.setdefault("params", {})
        if dev.get("name", "") in corner.source_values:
            p["dc"] = corner.source_values[dev["name"]]
        elif "dc" in p:
            p["dc"] = float(p["dc"]) * corner.vdd_scale


def _copy_device(d: Dict[str, Any]) -> Dict[str, Any]:
    out = {k: (list(v) if isinstance(v, list) else dict(v) if isinstance(v, dict) else v) for k, v in d.items()}
    prm = out.get("params")
    if isinstance(prm, dict):
        out["params"] = {k: (copy.deepcopy(v) if isinstance(v, (list, dict)) else v) for k, v in prm.items()}
    return out


def apply_corner(circuit: dict, corner: CornerConfig) -> dict:
    """A fresh circuit dict with the corner applied (the input is left untouched)."""
    c = dict(circuit)
    # per device: a new dict, a new params dict (the only thing the corner edits) and new node / value lists; a generic
    # deepcopy of 10 k device dicts cost 0.45 s per corner, half of a c6288 corner's set-up after the first one
    c["devices"] = [_copy_device(d) for d in circuit["devices"]]
    apply_process_corner(c["devices"], corner.process)
    apply_voltage_corner(c["devices"], corner.voltage)
    c["temperature_k"] = float(corner.temperature)
    c["corner_name"] = corner.name
    return c


def sweep_corners(n: int = 64) -> List[CornerConfig]:
    """The 64-point sweep of BASELINE.json configs[4]: {FF,TT,SS,FS} x {233.15,273.15,300.15,398.15} K
    x vdd {0.9,1.0,1.05,1.1}, truncated / cycled to ``n``."""
    # nominal first (TT, 27 C, 100 % vdd), so that corner 0 is the circuit of the single-transient benchmark
    base = create_standard_corners(["TT", "FF", "SS", "FS"], [300.15, 233.15, 273.15, 398.15], [1.0, 0.9, 1.05, 1.1])
    return [base[i % len(base)] for i in range(n)]


def monte_carlo_corners(n: int = 64, seed: int = 64) -> List[CornerConfig]:
    """Monte-Carlo draws applied like a process corner: mobility ~ N(1, 0.05), vth ~ N(0, 0.02 V),
    tox ~ N(1, 0.02)."""
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        pc = ProcessCorner(f"MC{i}", mobility_scale=float(rng.normal(1.0, 0.05)), vth_shift=float(rng.normal(0.0, 0.02)),
                           tox_scale=float(rng.normal(1.0, 0.02)))
        out.append(CornerConfig(name=f"MC{i}", process=pc, voltage=VoltageCorner("VDD_1.0"), temperature=DEFAULT_TEMPERATURE_K))
    return out


def shard(n_items: int, rank: int, world_size: int) -> List[int]:
    """Static assignment of sweep items to ranks: item i -> rank (i + i // world_size) mod world_size.

    Every rank gets one item of each consecutive block of ``world_size`` (as plain block-cyclic
    ``i mod world_size`` would), but the position inside the block rotates from block to block.  Sweep
    tables are nested loops (process x temperature x vdd), so ``i mod world_size`` hands a rank the same
    vdd / temperature in every block and the ranks' loads differ systematically (a low-vdd corner needs
    more NR iterations per step); the rotation is a Latin square over (block, position) and evens that out
    without any run-time exchange.  Counts per rank are those of block-cyclic."""
    return [i for i in range(n_items) if (i + i // world_size) % world_size == rank]


def run_corners(circuit: dict, bundles: Dict[str, dict], corners: Sequence[CornerConfig], *, rank: int = 0,
                world_size: int = 1, device: int = 0, engine_factory=None, gather=None, max_steps: Optional[int] = None,
                concurrent: int = 1, **engine_kwargs) -> CornerSweepResult:
    """Run this rank's share of ``corners`` (one fresh engine per corner) and gather.

    ``concurrent`` > 1 runs that many corners at a time on this rank's GPU, each engine on its own CUDA
    stream from its own host thread (the C-ABI calls release the GIL): a c6288-sized circuit is bound by
    the latency of its dependency chains and leaves most of the GPU idle, so whole-circuit replicas
    overlap almost perfectly.  Results do not depend on ``concurrent``.

    ``engine_factory(circuit, bundles, device=..., **kw)`` defaults to ``B200Engine``; ``gather``
    is a callable ``(list_of_local_payloads) -> list_of_all_payloads`` (bench.py / the tests pass a
    ``torch.distributed`` gather; None = single process).  Results come back in corner order on
    every rank that ``gather`` returns data to.
    """
    if engine_factory is None:
        from .engine import B200Engine as engine_factory  # noqa: N813
    def run_one(i: int) -> dict:
        kw = dict(engine_kwargs)
        if concurrent > 1:
            kw.setdefault("stream", "own")
        if concurrent > 2 and engine_factory.__name__ == "B200Engine":
            # several corners in flight: one slice and a small register footprint per model-kernel launch, so that the
            # launches of different corners share the SMs instead of queueing behind each other (bit-identical results)
            kw.setdefault("eval_slices", 1)
            kw.setdefault("eval_min_blocks", 4)
            # ... and the persistent LU kernels of one corner take 2 / concurrent of the CTA slots instead of all of them
            # (8 corners in flight, c6288 / psp103w: 2.94 k -> 3.33 k steps/s; 16 in flight: 3.83 k;
            # gpurun_out/r2_t_sweep.txt).  VJX_SWEEP_LU_DIV overrides.
            kw.setdefault("lu_grid_div", int(os.environ.get("VJX_SWEEP_LU_DIV", max(1, min(8, concurrent // 2)))))
        eng = engine_factory(apply_corner(circuit, corners[i]), bundles, device=device, **kw)
        times, V, st = eng.run_transient(max_steps=max_steps)
        n = int(st["n_steps"])
        out = dict(index=i, times=np.asarray(times[:n]), voltages=np.asarray(V[:n]),
                   stats={k: v for k, v in st.items() if np.isscalar(v)})
        if hasattr(eng, "close"):
            eng.close()
        return out

    mine = shard(len(corners), rank, world_size)
    if concurrent > 1 and len(mine) > 1:
        from concurrent.futures import ThreadPoolExecutor

        with ThreadPoolExecutor(max_workers=int(concurrent)) as pool:
            local = list(pool.map(run_one, mine))
    else:
        local = [run_one(i) for i in mine]
    payloads = gather(local) if gather is not None else local
    res = CornerSweepResult(corners=list(corners))
    if payloads is None:
        return res
    for p in sorted(payloads, key=lambda d: d["index"]):
        st = p["stats"]
        res.results.append(CornerResult(corner=corners[p["index"]], times=p["times"], voltages=p["voltages"], stats=st,
                                        converged=int(st.get("error_flags", 0)) == 0))
    return res


def make_dist_gather(dst: int = 0):
    """Gather callable for ``run_corners`` over ``torch.distributed`` (NCCL on the GPU box, gloo in
    the CPU tests): waveforms travel as padded tensors through ``dist.gather`` (K13, SURVEY.md
    §2.3), the small per-corner metadata through ``gather_object``.  Returns the full payload list
    on rank ``dst`` and ``None`` elsewhere."""
    import torch
    import torch.distributed as dist

    def gather(local: List[dict]):
        rank, world = dist.get_rank(), dist.get_world_size()
        backend = dist.get_backend()
        dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
        meta = [dict(index=p["index"], n=int(len(p["times"])), n_ext=int(p["voltages"].shape[1]) if p["voltages"].ndim == 2 else 0,
                     stats=p["stats"]) for p in local]
        all_meta = [None] * world if rank == dst else None
        dist.gather_object(meta, all_meta, dst=dst)
        # common padded shape: (max corners per rank, max rows, 1 + n_ext)
        shape = torch.tensor([len(local), max([m["n"] for m in meta] + [0]), max([m["n_ext"] for m in meta] + [0])],
                             dtype=torch.int64, device=dev)
        dist.all_reduce(shape, op=dist.ReduceOp.MAX)
        nc, nr, ne = (int(v) for v in shape.tolist())
        buf = torch.zeros((max(nc, 1), max(nr, 1), 1 + max(ne, 1)), dtype=torch.float64, device=dev)
        for k, p in enumerate(local):
            n = len(p["times"])
            buf[k, :n, 0] = torch.from_numpy(np.ascontiguousarray(p["times"]))
            if n and p["voltages"].size:
                buf[k, :n, 1:1 + p["voltages"].shape[1]] = torch.from_numpy(np.ascontiguousarray(p["voltages"]))
        bufs = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
        dist.gather(buf, bufs, dst=dst)
        if rank != dst:
            return None
        out = []
        for r in range(world):
            arr = bufs[r].cpu().numpy()
            for k, m in enumerate(all_meta[r]):
                out.append(dict(index=m["index"], times=arr[k, :m["n"], 0].copy(),
                                voltages=arr[k, :m["n"], 1:1 + m["n_ext"]].copy(), stats=m["stats"]))
        return out

    return gather
