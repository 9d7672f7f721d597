#!/usr/bin/env python
"""bench.py - ms per accepted transient step of the Newton-Raphson hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c6288|mul|ring]

One "step" = one accepted timestep of the adaptive transient (all of its NR iterations, rejected
attempts included in the time - the reference's own definition, full_mna.py:1142-1148,1176).

* ``value``  device-resident loop, inputs already in HBM, CUDA events on the launch stream tight
  around each step's graph launch, a 256 MB L2 flush between steps (not timed).
* ``value_l2_resident``  the same K steps as ONE uninterrupted device-side loop (production mode: the
  working set of one circuit lives in L2; no host involvement between steps).
* ``e2e``  the same K steps through ``B200Engine.run_transient`` (the reference-facing API) with
  host buffers: initial state host->device, recorded waveforms device->host, wall clock.
* ``roofline``  the dominant kernel of a step, timed alone with CUDA events (vjx_time_phase).
* ``cpu_baseline``  the oracle port (compiled model functions on all host threads, cached-ordering SuperLU, NumPy) on a bounded sample.

Every rank runs ``--corners`` PVT corners of the circuit (default 8, so 8 GPUs run the 64-point sweep
of BASELINE.json configs[4]; corner i goes to GPU i mod N, corner 0 is the nominal circuit and is
reported alone as ``value_nominal``).  Weak scaling: fixed work per GPU, no collective inside the
step loop; for N > 1 the recorded waveforms are gathered to rank 0 with NCCL afterwards (in ``e2e``).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "ms/transient step (c6288) on B200"
UNIT = "ms/step"
HBM_FALLBACK_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            for k in ("hbm_gbs", "hbm_gb_s", "hbm"):
                if k in d:
                    return float(d[k]), "measured"
        except Exception:
            pass
    return HBM_FALLBACK_GBS, "fallback"


def load_extra():
    """Numbers measured on this pool's B200s by this repo's own probes (fp64 DFMA peak, ncu DRAM traffic)."""
    for fn in ("r2_measured_extra.json", "r1_measured_extra.json"):
        try:
            return json.load(open(os.path.join(ROOT, "profiles", fn)))
        except Exception:
            continue
    return {}


def eval_flops(eng):
    """Static fp64 operation count of the emitted model kernels x instances (cuda_emitter counts arithmetic,
    divisions and transcendental calls of the whole MIR; the executed path is shorter)."""
    tot = {"arith": 0, "div": 0, "transc": 0}
    for m in eng.models.values():
        f = m["meta"].get("flops_static", {})
        for k in tot:
            tot[k] += int(f.get(k, 0)) * m["n_dev"]
    return tot


MODEL_DESC = {"psp103w": "PSP103-weight MOSFETs (psp103w: R=13, J=56, 754 parameters, 6.4k MIR instructions per evaluation)",
              "psp103": "light PSP103-shaped surrogate MOSFETs (J=42, ~0.9k flops; round-1 stand-in)"}


def make_circuit(workload: str, nbits: int, corner: int = 0, model: str = "psp103w"):
    from vajax_b200.benchmarks import circuits

    if workload == "ring":
        return circuits.ring_circuit(mos_model=model), f"9-stage ring oscillator (47 unknowns), {MODEL_DESC[model]}"
    if workload == "mul":
        nb = nbits
    else:
        nb = 16
    circ = circuits.multiplier_circuit(nb, mos_model=model)
    if corner:  # corner 0 of the sweep is the nominal circuit itself
        from vajax_b200.analysis.corners import apply_corner, sweep_corners

        circ = apply_corner(circ, sweep_corners(64)[corner % 64])
    desc = (f"{nb}x{nb} carry-save array multiplier in the c6288/mul64 gate library, "
            f"{sum(1 for d in circ['devices'] if d['model'] == model)} {MODEL_DESC[model]}, trap, uic")
    return circ, desc


def config_dict(args, desc, world, n_unknowns, csr_nnz):
    """The same dict in both arms (the driver compares them)."""
    return {"workload": desc, "circuit": args.workload, "model": args.model, "n_unknowns": int(n_unknowns), "csr_nnz": int(csr_nnz),
            "parallelism": f"replicas x{world}" if world > 1 else "single circuit",
            "l2": "256 MB flush between timed steps (value); value_l2_resident is the uninterrupted device loop"}


def get_bundles():
    from vajax_b200.models.fixtures import get_bundle

    return {k: get_bundle(k) for k in ("resistor", "capacitor", "diode", "psp103", "psp103w")}


class ClockSampler:
    """nvidia-smi clocks during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def phase_bytes(eng, lu):
    """Algorithmic bytes per launch of each phase."""
    ev = 0
    jc = 0
    rc = 0
    for m in eng.models.values():
        D, R, J, L = m["n_dev"], m["R"], m["J"], m["num_limit_states"]
        nv, P, Cd = m["vnode1"].shape[1], m["dev_params"].shape[1], m["dev_cache"].shape[1]
        ev += D * 8 * (2 * nv + P + Cd + L) + D * 8 * (J + 4 * R + L)
        jc += D * J
        rc += D * R
    n, nse = eng.n_aug, eng.nse
    nnz = lu["nnz_lu"]
    return {
        "device_eval": ev,
        "jacobian_stamp": jc * (8 + 4) + nse * 8 + n * (8 + 4),
        "residual": 4 * rc * (8 + 4) + rc * (8 + 4) + 5 * n * 8,
        # refactorisation (read A, read + write L+U) with the forward solve folded in (b, y, 1/pivot)
        "lu_factor_fwd": 8 * (nse + 2 * nnz) + 3 * 8 * n,
        "lu_usolve": 8 * ((nnz - n) // 2 + n) + 3 * 8 * n,
        "nr_update": 14 * n * 8,
    }


# (name, vjx_time_phase id); id 4 was the separate forward solve, now part of id 3
PHASES = [("device_eval", 0), ("jacobian_stamp", 1), ("residual", 2), ("lu_factor_fwd", 3), ("lu_usolve", 5), ("nr_update", 6)]


def launches_per(eng, n_lu):
    nm = len(eng.models)
    isrc = 1 if eng.n_is > 0 else 0
    per_iter = isrc + nm + 1 + 1 + n_lu + 2   # models, stamp, residual, LU launches, k_delta_reduce + k_nr_update (no k_iter_begin since round 2)
    per_attempt = 2 + (isrc + nm + 1 + 1 + 1 + 1 + 1 + 1)
    return per_iter, per_attempt


def run_b200(args, rank, world, local_rank):
    import torch

    from vajax_b200.analysis.engine import B200Engine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - vajax_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    bundles = get_bundles()
    K, W = args.steps, max(args.warmup, 3)
    n_corners = args.corners
    engines = []
    desc = ""
    corner_ids = []
    from vajax_b200.analysis.corners import shard

    for gid in shard(n_corners * world, rank, world):  # rotated block-cyclic (corners.shard)
        corner_ids.append(gid)
        circ, desc = make_circuit(args.workload, args.nbits, corner=gid, model=args.model)
        # every engine on its own stream: the sweep leg below runs the corners of a rank concurrently
        engines.append(B200Engine(circ, bundles, device=local_rank, use_device_graph=not args.no_graph,
                                  stream="own").prepare())
    eng = engines[0]
    rows = 1 + K
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

    def barrier():
        if world > 1:
            import torch.distributed as dist

            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up: W steps of every engine (LU analysis, graph build, caches)
    for e in engines:
        e.tran_begin(max_steps=1 + W)
        e.tran_run()
    if world > 1:
        import torch.distributed as dist

        # ... and of the gather path: NCCL sets up its point-to-point channels on the first gather (about a
        # second, once per process group), which is start-up cost like the graph build, not part of a sweep
        probe = torch.zeros(1024, dtype=torch.float64, device=dev)
        dist.gather(probe, [torch.empty_like(probe) for _ in range(world)] if rank == 0 else None, dst=0)
        torch.cuda.synchronize()
    # ---- value: per-step graph launches with an L2 flush in between (events inside vjx_tran_run)
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    total_ms, steps_done, attempts, nr_total, rejected = 0.0, 0, 0, 0, 0
    per_corner = []
    t_wall0 = time.perf_counter()
    for e in engines:
        e.tran_begin(max_steps=rows)
        step = 1
        ms0 = total_ms
        while step < rows:
            flush.fill_(1.0)
            a = e.tran_run(step_limit=step + 1)
            if a == 0:
                break
            total_ms += e.tran_run_ms()[0]
            step += 1
        sc = e.tran_scalars()
        per_corner.append((total_ms - ms0) / max(sc["step_idx"] - 1, 1))
        steps_done += sc["step_idx"] - 1
        attempts += sc["attempts"]
        nr_total += sc["total_nr"]
        rejected += sc["rejected"]
        if sc["error_flags"]:
            raise SystemExit(f"bench.py: device error flags {sc['error_flags']}")
    barrier()
    wall_flushed = time.perf_counter() - t_wall0
    # ---- value_l2_resident: the same steps as one uninterrupted device loop per engine
    barrier()
    res_ms, res_steps = 0.0, 0
    for e in engines:
        e.tran_begin(max_steps=rows)
        e.tran_run(step_limit=rows)
        res_ms += e.tran_run_ms()[0]
        res_steps += e.tran_scalars()["step_idx"] - 1
    barrier()
    # ---- sweep: the corners of this rank run CONCURRENTLY (one stream + host thread each), no L2 flush:
    # a c6288-sized circuit is latency-bound and leaves most of the GPU idle
    sweep_ms, sweep_steps = 0.0, 0
    if n_corners > 1 and not args.no_graph:
        # engines configured for concurrency (run_corners(concurrent > 2) does the same): one output slice and 4 CTAs
        # per SM for the model kernels, so that the corners' launches overlap instead of queueing
        sweep_engines = []
        for gid in corner_ids:
            circ, _ = make_circuit(args.workload, args.nbits, corner=gid, model=args.model)
            sweep_engines.append(B200Engine(circ, bundles, device=local_rank, use_device_graph=True, stream="own",
                                            eval_slices=int(os.environ.get("VJX_SWEEP_SLICES", "1")),
                                            eval_min_blocks=int(os.environ.get("VJX_SWEEP_MINBLOCKS", "4")),
                                            lu_grid_div=int(os.environ.get("VJX_SWEEP_LU_DIV", max(1, min(8, len(corner_ids) // 2))))).prepare())
        for e in sweep_engines:   # warm-up: LU analysis, graph build
            e.tran_begin(max_steps=1 + W)
            e.tran_run()
        for e in sweep_engines:
            e.tran_begin(max_steps=rows)
        barrier()
        ev_s, ev_e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev_s.record()
        ths = [threading.Thread(target=e.tran_run, kwargs=dict(step_limit=rows)) for e in sweep_engines]
        for th in ths:
            th.start()
        for th in ths:
            th.join()
        torch.cuda.synchronize()
        ev_e.record()
        torch.cuda.synchronize()
        sweep_ms = ev_s.elapsed_time(ev_e)
        for e in sweep_engines:
            sc = e.tran_scalars()
            sweep_steps += sc["step_idx"] - 1
            if sc["error_flags"]:
                raise SystemExit(f"bench.py: device error flags {sc['error_flags']} in the concurrent sweep")
        for e in sweep_engines:
            e.close()
        barrier()
    clocks = sampler.stop()
    # ---- e2e through the public API with host buffers (H2D of the initial state, D2H of the waveforms)
    barrier()
    t0 = time.perf_counter()
    e2e_steps, h2d, d2h = 0, 0, 0
    outs = []
    for e in engines:
        times, V_out, st = e.run_transient(max_steps=rows)
        e2e_steps += st["n_steps"] - 1
        h2d += e._tran["h2d_bytes"]
        d2h += e._tran.get("d2h_bytes", 0)
        outs.append((times, V_out))
    torch.cuda.synchronize()
    e2e_local = time.perf_counter() - t0
    gather_s = 0.0
    if world > 1:
        import torch.distributed as dist

        # K13: gather the recorded waveforms of every corner on rank 0 (NCCL over NVLink)
        tg = time.perf_counter()
        mine = torch.from_numpy(np.stack([o[1] for o in outs])).to(dev)
        bufs = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
        dist.gather(mine, bufs, dst=0)
        torch.cuda.synchronize()
        gather_s = time.perf_counter() - tg
        e2e_local += gather_s
    # ---- reduce over ranks: MAX time, SUM steps
    vals = torch.tensor([total_ms, res_ms, e2e_local, wall_flushed, sweep_ms], dtype=torch.float64, device=dev)
    cnts = torch.tensor([steps_done, res_steps, e2e_steps, attempts, nr_total, rejected, h2d, d2h, sweep_steps],
                        dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist

        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnts, op=dist.ReduceOp.SUM)
    vals, cnts = vals.cpu().numpy(), cnts.cpu().numpy()
    if rank != 0:
        return
    g_steps, g_res_steps, g_e2e_steps = int(cnts[0]), int(cnts[1]), int(cnts[2])
    ms_per_step = vals[0] / max(g_steps, 1)
    # ---- roofline of the dominant kernel (rank 0, engine 0), timed alone
    lu = eng.lu_stats()
    pb = phase_bytes(eng, lu)
    peak, peak_kind = load_peaks()
    iters_per_step = cnts[4] / max(g_steps, 1)
    att_per_step = cnts[3] / max(g_steps, 1)
    phase = {}
    for nm, pid in PHASES:
        ms, nl = eng.time_phase(pid, 20)
        calls = {"device_eval": iters_per_step + att_per_step, "residual": iters_per_step + att_per_step}.get(nm, iters_per_step)
        phase[nm] = {"ms": ms, "launches": nl, "calls_per_step": calls, "ms_per_step": ms * calls,
                     "alg_bytes": pb[nm], "gbs": pb[nm] / (ms * 1e-3) / 1e9 if ms > 0 else 0.0}
    per_iter, per_attempt = launches_per(eng, phase["lu_factor_fwd"]["launches"] + phase["lu_usolve"]["launches"])
    gpu_launches = int(cnts[4] * per_iter + cnts[3] * per_attempt)
    extra = load_extra()
    ef = eval_flops(eng)
    ev_ms = phase["device_eval"]["ms"]
    if ev_ms > 0:
        # transcendental calls (exp/log/sqrt/pow) counted as 20 fp64 operations each, divisions as 8
        fl = ef["arith"] + 8 * ef["div"] + 20 * ef["transc"]
        phase["device_eval"].update({"flops_static": fl, "tflops": fl / (ev_ms * 1e-3) / 1e12,
                                     "fp64_peak_tflops": extra.get("fp64_tflops"),
                                     "frac_of_fp64_peak": (fl / (ev_ms * 1e-3) / 1e12) / extra["fp64_tflops"] if extra.get("fp64_tflops") else None})
    dom = max(phase, key=lambda k: phase[k]["ms_per_step"])
    pd = phase[dom]
    roof = {"kernel": dom, "bound": "hbm", "achieved": pd["gbs"], "peak": peak, "unit": "GB/s",
            "frac": pd["gbs"] / peak,
            "traffic": extra.get("traffic_bytes_per_launch", {}).get(args.workload, {}).get(dom), "peak_source": peak_kind,
            "launch_ms": pd["ms"], "share_of_step": pd["ms_per_step"] / max(ms_per_step, 1e-12),
            "kernels": "k_lu_load + k_lu_rows + k_lu_dense_tail" if dom == "lu_factor_fwd" else dom,
            "traffic_source": "ncu --set full capture kept under profiles/ (r2_measured_extra.json), not measured in this run",
            "note": "timed alone with CUDA events, L2-warm (working set < L2 at this size); bound by the dependency "
                    "chain of the elimination (latency of dependent fp64 operations and L2 round trips), not by bytes: "
                    "see profiles/r1_lu_notes.md, profiles/r2_dense_notes.md"}
    out = {
        "metric": METRIC, "value": ms_per_step, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_per_step, "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": config_dict(args, desc, world, eng.n_aug, eng.nse),
        "run_stats": {"steps_timed": g_steps, "attempts": int(cnts[3]), "nr_iterations": int(cnts[4]), "rejected": int(cnts[5]),
                      "corners_per_gpu": n_corners},
        "value_l2_resident": vals[1] / max(g_res_steps, 1),
        "value_nominal": per_corner[0] if corner_ids[0] == 0 else None,
        "rank0_corners": {"ids": corner_ids, "ms_per_step": per_corner},
        "e2e": {"value": vals[2] * 1e3 / max(g_e2e_steps, 1), "unit": UNIT,
                "h2d_bytes_per_step": cnts[6] / max(g_e2e_steps, 1), "d2h_bytes_per_step": cnts[7] / max(g_e2e_steps, 1),
                "gather_s": gather_s},
        "sweep_steps_per_s": 1e3 / ms_per_step if ms_per_step > 0 else None,
        "sweep_concurrent": {"steps_per_s": cnts[8] / (vals[4] * 1e-3) if vals[4] > 0 else None,
                             "ms_per_step": vals[4] / cnts[8] if cnts[8] > 0 else None, "corners_in_flight_per_gpu": n_corners,
                             "steps": int(cnts[8]),
                             "note": "same corners and steps as `value`, but the corners of a GPU run concurrently on "
                                     "separate streams (run_corners(concurrent=...)); CUDA events, max over ranks"},
        "gpu_launches": gpu_launches,
        "clocks": clocks,
        "roofline": roof,
        "phases": {k: {kk: (round(vv, 6) if isinstance(vv, float) else vv) for kk, vv in v.items()} for k, v in phase.items()},
        "lu": lu,
        "reported_reference": {"c6288_jax_gpu_T4_ms_per_step": 19.81, "c6288_jax_cpu_ms_per_step": 88.06,
                               "note": "reference README numbers on its real c6288 netlist + PSP103; not the same circuit/model"},
    }
    if world == 1 and args.mul64_steps > 0 and args.workload == "c6288":
        for e in engines:
            e.close()
        engines.clear()
        out["mul64"] = mul64_record(args, local_rank, peak)
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(args, sample_steps=args.cpu_steps)
        out["cpu_baseline"].pop("n_unknowns"), out["cpu_baseline"].pop("csr_nnz")
    _emit(json.dumps(out))


def mul64_record(args, local_rank, peak):
    """Second half of BASELINE.json's metric: ms per accepted step of the mul64 transient on one B200
    (266k MOSFETs, 666k unknowns; the reference times 15 steps, README.md:80)."""
    from vajax_b200.analysis.engine import B200Engine
    from vajax_b200.benchmarks.circuits import mul64_circuit

    t0 = time.perf_counter()
    e = B200Engine(mul64_circuit(64, mos_model=args.model), get_bundles(), device=local_rank, use_device_graph=True).prepare()
    prep = time.perf_counter() - t0
    W, K = 3, args.mul64_steps
    e.tran_begin(max_steps=1 + W)
    e.tran_run()
    e.tran_begin(max_steps=1 + W + K)
    e.tran_run(step_limit=1 + W)
    sc0 = e.tran_scalars()
    e.tran_run(step_limit=1 + W + K)
    ms = e.tran_run_ms()[0]
    sc = e.tran_scalars()
    steps = sc["step_idx"] - sc0["step_idx"]
    lu = e.lu_stats()
    pb = phase_bytes(e, lu)
    ph = {}
    for nm, pid in PHASES:
        t, nl = e.time_phase(pid, 5)
        ph[nm] = {"ms": round(t, 4), "launches": nl, "alg_bytes": pb[nm], "gbs": round(pb[nm] / (t * 1e-3) / 1e9, 2) if t > 0 else 0.0,
                  "frac_of_hbm_peak": round(pb[nm] / (t * 1e-3) / 1e9 / peak, 5) if t > 0 else 0.0}
    rec = {"metric": "ms/transient step (mul64) on B200", "value": ms / max(steps, 1), "unit": UNIT, "steps_timed": steps,
           "warmup_steps": W, "nr_iterations": sc["total_nr"] - sc0["total_nr"], "rejected": sc["rejected"] - sc0["rejected"],
           "n_unknowns": e.n_aug, "csr_nnz": e.nse, "prepare_s": round(prep, 2), "phases": ph, "lu": lu,
           "timing": "one uninterrupted device-side loop (CUDA graph with conditional nodes), CUDA events on the launch stream; "
                     "the working set (about 0.6 GB) is larger than L2",
           "error_flags": sc["error_flags"]}
    # e2e through the public API with host buffers and the netlist's own save list (N4: `save v(p0) ... v(p127)`,
    # vajax/benchmarks/data/mul64/runme.sim:296-308): UIC initial state host->device, K + W steps, the saved
    # waveforms device->host, wall clock
    import torch

    e.set_save_list([f"top.p{k}" for k in range(128) if f"top.p{k}" in e.node_names])
    e.run_transient(max_steps=1 + W)   # graph rebuild for the new output layout (set-up, not timed)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    _, V_out, st = e.run_transient(max_steps=1 + W + K)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    n = max(st["n_steps"] - 1, 1)
    rec["e2e"] = {"value": wall * 1e3 / n, "unit": UNIT, "steps": n, "h2d_bytes_per_step": e._tran["h2d_bytes"] / n,
                  "d2h_bytes_per_step": e._tran.get("d2h_bytes", 0) / n, "saved_nodes": int(V_out.shape[1]),
                  "wall_s": round(wall, 4), "setup_s": round(e.timings.get("tran_setup", 0.0), 4),
                  "device_loop_ms": round(st.get("device_time_ms", 0.0), 3),
                  "note": "B200Engine.run_transient, host buffers, save list = the 128 product bits"}
    e.close()
    return rec


def cpu_baseline(args, sample_steps: int, threads_note: str = "") -> dict:
    """The CPU restatement of the reference's algorithm (oracle/) on this box's host cores, bounded sample, shaped
    like the reference's own CPU path: model functions COMPILED and spread over all host threads (the reference
    jit-compiles ``vmap(eval_fn)`` with XLA-CPU; here the same straight-line fp64 code goes through ``g++ -O2`` +
    OpenMP, oracle/host_compiled.py), sparse LU with the fill-reducing ordering computed once per pattern (the
    reference's UMFPACK FFI caches its symbolic analysis; oracle/engine.py ``reuse_ordering``), assembly / NR /
    step control in NumPy.  ``interpreter_ms_per_step`` is the same transient through the NumPy MIR interpreter on one
    thread - the configuration the parity tests check against (and the only CPU leg of earlier rounds)."""
    from oracle import host_compiled
    from oracle.engine import OracleEngine

    circ, _ = make_circuit(args.workload, args.nbits, model=args.model)
    o = OracleEngine(circ, get_bundles())
    o.setup()
    threads = int(os.environ.get("VJX_HOST_THREADS", "0")) or (os.cpu_count() or 1)
    o.run_transient(max_steps=2)   # warm-up (imports, first factorisation)
    t0 = time.perf_counter()
    _, _, st = o.run_transient(max_steps=1 + min(sample_steps, 3))
    interp = (time.perf_counter() - t0) * 1e3 / max(st["n_steps"] - 1, 1)
    o.reuse_ordering = True
    runs = []
    with host_compiled.installed(threads=threads):
        o.run_transient(max_steps=3)   # warm-up: g++ on the first call (cached under oracle/_build), ordering, imports
        for _ in range(3):   # best of three
            t0 = time.perf_counter()
            _, _, st = o.run_transient(max_steps=1 + sample_steps)
            runs.append(time.perf_counter() - t0)
    dt = min(runs)
    n = max(st["n_steps"] - 1, 1)
    return {"value": dt * 1e3 / n, "unit": UNIT, "cores": threads, "kind": "port", "best_of": 3,
            "runs_ms_per_step": [round(r * 1e3 / n, 2) for r in runs],
            "interpreter_ms_per_step": round(interp, 2),
            "host_cores_available": os.cpu_count(), "n_unknowns": int(o.n_aug), "csr_nnz": int(len(o.csr_indices)),
            "sample": f"first {n} accepted steps of the same transient; the repo's CPU restatement of the reference's "
                      f"algorithm (NOT the reference's XLA-CPU + UMFPACK build, which cannot be installed here): model "
                      f"functions compiled with g++ -O2 and run on {threads} OpenMP threads, SciPy SuperLU with a cached "
                      f"ordering (1 thread), NumPy assembly and step control (1 thread){threads_note}"}


def run_reference(args, rank, world):
    """Reference arm: the reference's JAX implementation cannot be installed here (no jax, no
    openvaf_py), so the CPU restatement of its algorithm (oracle/) is timed."""
    if rank != 0:
        return
    K, W = args.steps, max(args.warmup, 3)
    sample = min(K, args.cpu_steps)
    cb = cpu_baseline(args, sample_steps=sample)
    _, desc = make_circuit(args.workload, args.nbits, model=args.model)
    out = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": world, "steps": K,
           "warmup": W, "ms_per_step": cb["value"], "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic",
           "config": config_dict(args, desc, world, cb.pop("n_unknowns"), cb.pop("csr_nnz")),
           "cpu_baseline": cb,
           "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(json.dumps(out))


def main():
    # the JSON line must be the only thing on stdout: libraries (NCCL's version banner) write there too
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    globals()["_emit"] = lambda line: (real_stdout.write(line + "\n"), real_stdout.flush())
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c6288", choices=["c6288", "mul", "ring"])
    ap.add_argument("--nbits", type=int, default=16)
    ap.add_argument("--model", default="psp103w", choices=["psp103w", "psp103"],
                    help="MOSFET model: psp103w = PSP103-weight (default, the headline), psp103 = round-1 light surrogate")
    ap.add_argument("--corners", type=int, default=8,
                    help="PVT corners (whole-circuit replicas) per GPU; 8 per GPU = the 64-point sweep at 8 GPUs; "
                         "corner 0 is the nominal circuit (value_nominal)")
    ap.add_argument("--cpu-steps", type=int, default=12)
    ap.add_argument("--mul64-steps", type=int, default=15, help="also time this many mul64 steps at N=1 (0 = skip)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="host-polled NR/step loops (for ncu launch lists)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_b200(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist

            dist.destroy_process_group()


if __name__ == "__main__":
    main()
