"""DC operating point with continuation: plain Newton first, then the gdev -> gshunt -> src chain.

Host-side control around seam S3.  Every solve is one ``nr_solve`` call with the reference's argument order
(``vjx_nr_solve``: the whole NR loop runs on the device); between solves only the three scalars
(iterations, converged, max|f|) come back to the host, the solution vector stays in device memory when the
solver is ``B200Engine.nr_solve`` with ``return_device=True`` (``device_solver``).

Behaviour follows the reference so that a circuit takes the same continuation steps and NR iterations:
  * ramp of an extra conductance, VACASK style: vajax/analysis/homotopy.py:90-276 (``gmin_stepping``)
  * source ramp with the conductance ramps as fall-back at factor 0: homotopy.py:279-467 (``source_stepping``)
  * chain driver: homotopy.py:470-567 (``run_homotopy_chain``)
  * entry point: vajax/analysis/dc_operating_point.py:128-242 (``compute_dc_operating_point``)
The reference's quirks are kept because they decide the step sequence: the "fast" threshold never widens the
conductance factor (``gmin_factor_max`` is unused there too), a failure after a success takes the fourth root
of the factor, and the source ramp halves its step on failure but scales it by ``source_scale`` on fast success.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, replace
from typing import Callable, Optional, Sequence, Tuple

import numpy as np

# fractions of ``max_iterations`` that classify one solve as fast / slow (homotopy.py:28-33)
FAST_FRACTION = 0.25
SLOW_FRACTION = 0.75


@dataclass
class HomotopyConfig:
    """Same fields and defaults as the reference's ``HomotopyConfig`` (homotopy.py:36-73)."""
    gmin: float = 1e-12
    gdev_start: float = 1e-3
    gdev_target: float = 1e-13
    gmin_factor: float = 10.0
    gmin_factor_min: float = 1.1
    gmin_factor_max: float = 100.0
    gmin_max: float = 1.0
    gmin_max_steps: int = 100
    source_step: float = 0.1
    source_step_min: float = 0.001
    source_scale: float = 2.0
    source_max_steps: int = 100
    chain: Tuple[str, ...] = ("gdev", "gshunt", "src")
    max_iterations: int = 100
    abstol: float = 1e-9
    debug: int = 0

    @classmethod
    def from_options(cls, o) -> "HomotopyConfig":
        """Field mapping of dc_operating_point.py:188-205 (SimulationOptions -> HomotopyConfig); options an
        object lacks fall back to the reference's SimulationOptions defaults (options.py:60-190)."""
        g = lambda k, d: getattr(o, k, d)  # noqa: E731
        return cls(gmin=g("gmin", 1e-12), gdev_start=g("homotopy_startgmin", 1e-3), gdev_target=g("homotopy_mingmin", 1e-15),
                   gmin_factor=g("homotopy_gminfactor", 10.0), gmin_factor_min=g("homotopy_mingminfactor", 1.00005),
                   gmin_factor_max=g("homotopy_maxgminfactor", 10.0), gmin_max=g("homotopy_maxgmin", 1e2),
                   gmin_max_steps=g("homotopy_gminsteps", 100), source_step=g("homotopy_srcstep", 0.01),
                   source_step_min=g("homotopy_minsrcstep", 1e-7), source_scale=g("homotopy_srcscale", 3.0),
                   source_max_steps=g("homotopy_srcsteps", 100), chain=tuple(g("op_homotopy", ("gdev", "gshunt", "src"))),
                   max_iterations=g("op_itlcont", 50), abstol=g("abstol", 1e-12), debug=0)


@dataclass
class HomotopyResult:
    """homotopy.py:76-86"""
    converged: bool
    V: object
    method: str = ""
    iterations: int = 0
    homotopy_steps: int = 0
    final_gmin: float = 0.0
    final_source_scale: float = 1.0


class _Solver:
    """One DC solve = ``nr_solve(X, vs, is, Q_prev, 0, gmin, gshunt, 0, 0, None, 0, None)``; counts what the
    continuation spent.  ``nr_solve`` takes the engine-style argument list (no ``device_arrays``: the per-instance
    data live in the handle); ``analysis.reference_adapter`` provides the reference-style wrapper."""

    def __init__(self, nr_solve: Callable, vs, isrc, Q_prev, trace: Optional[list] = None):
        self.fn, self.vs, self.isrc, self.Q_prev = nr_solve, vs, isrc, Q_prev
        self.iterations = 0
        self.solves = 0
        self.trace = trace

    def __call__(self, X, scale: float, gmin: float, gshunt: float):
        out = self.fn(X, self.vs * scale, self.isrc * scale, self.Q_prev, 0.0, gmin, gshunt, 0.0, 0.0, None, 0.0, None)
        X_new, iters, ok = out[0], int(out[1]), bool(out[2])
        self.iterations += iters
        self.solves += 1
        if self.trace is not None:
            self.trace.append((float(scale), float(gmin), float(gshunt), iters, ok))
        return X_new, iters, ok


def _conductance_ramp(solve: _Solver, X0, scale: float, cfg: HomotopyConfig, mode: str) -> HomotopyResult:
    """Ramp an extra conductance from ``gdev_start`` down to ``gdev_target`` (homotopy.py:90-276).  ``mode`` "gdev":
    added to the device gmin; "gshunt": a shunt from every node to ground."""
    it0, n0 = solve.iterations, solve.solves
    fast = int(cfg.max_iterations * FAST_FRACTION)
    slow = int(cfg.max_iterations * SLOW_FRACTION)
    g, factor = cfg.gdev_start, cfg.gmin_factor
    X, X_good, g_good, have_good = X0, X0, cfg.gdev_start, False
    method = f"{mode}_stepping"

    def done(ok, V, final_g):
        return HomotopyResult(bool(ok), V, method, solve.iterations - it0, solve.solves - n0, final_g)

    for _ in range(cfg.gmin_max_steps):
        gm, gs = (cfg.gmin + g, 0.0) if mode == "gdev" else (cfg.gmin, g)
        X_new, iters, ok = solve(X, scale, gm, gs)
        if ok:
            have_good, X_good, g_good = True, X_new, g
            if g <= cfg.gdev_target:
                break
            if iters > slow and iters > fast:
                factor = max(math.sqrt(factor), cfg.gmin_factor_min)
            if g / factor < cfg.gdev_target:
                factor, g = g / cfg.gdev_target, cfg.gdev_target
            else:
                g = g / factor
            X = X_new
        elif not have_good:
            g = g * factor
            if g > cfg.gmin_max:
                return done(False, X0, g)
        else:
            factor = factor ** 0.25
            if factor < cfg.gmin_factor_min:
                break
            X, g = X_good, g_good
    if not have_good:
        return done(False, X_good, g)
    X_fin, _, ok = solve(X_good, scale, cfg.gmin, 0.0)   # the solve at the circuit's own gmin
    return done(ok, X_fin if ok else X_good, cfg.gmin)


def gmin_stepping(nr_solve: Callable, V_init, vsource_vals, isource_vals, Q_prev, source_scale: float,
                  config: HomotopyConfig, mode: str = "gdev", trace: Optional[list] = None) -> HomotopyResult:
    """Reference-shaped entry (homotopy.py:90; without the ``device_arrays`` argument)."""
    return _conductance_ramp(_Solver(nr_solve, vsource_vals, isource_vals, Q_prev, trace), V_init, source_scale, config, mode)


def _source_ramp(solve: _Solver, X0, cfg: HomotopyConfig) -> HomotopyResult:
    """Ramp every independent source from 0 to 100 % (homotopy.py:279-467)."""
    it0, n0 = solve.iterations, solve.solves
    fast = int(cfg.max_iterations * FAST_FRACTION)
    slow = int(cfg.max_iterations * SLOW_FRACTION)

    def done(ok, V, scale):
        return HomotopyResult(bool(ok), V, "source_stepping", solve.iterations - it0, solve.solves - n0, 0.0, scale)

    X_new, _, ok = solve(X0, 0.0, cfg.gmin, 0.0)
    if ok:
        X_good = X_new
    else:
        # no solution with the sources off: conductance ramps at factor 0, device gmin first (homotopy.py:352-397)
        r = _conductance_ramp(solve, X0, 0.0, cfg, "gdev")
        if not r.converged:
            r = _conductance_ramp(solve, X0, 0.0, cfg, "gshunt")
        if not r.converged:
            return done(False, X0, 0.0)
        X_good = r.V
    at, step = 0.0, cfg.source_step
    for _ in range(cfg.source_max_steps):
        nxt = min(at + step, 1.0)
        X_new, iters, ok = solve(X_good, nxt, cfg.gmin, 0.0)
        if ok:
            X_good, at = X_new, nxt
            if at >= 1.0:
                return done(True, X_good, 1.0)
            if iters <= fast:
                step *= cfg.source_scale
            elif iters > slow:
                step = max(step / cfg.source_scale, cfg.source_step_min)
        else:
            step *= 0.5
            if step < cfg.source_step_min:
                break
    return done(at >= 1.0, X_good, at)


def source_stepping(nr_solve: Callable, V_init, vsource_vals, isource_vals, Q_prev, config: HomotopyConfig,
                    trace: Optional[list] = None) -> HomotopyResult:
    """Reference-shaped entry (homotopy.py:279)."""
    return _source_ramp(_Solver(nr_solve, vsource_vals, isource_vals, Q_prev, trace), V_init, config)


def run_homotopy_chain(nr_solve: Callable, V_init, vsource_vals, isource_vals, Q_prev, config: HomotopyConfig,
                       trace: Optional[list] = None) -> HomotopyResult:
    """Try the algorithms of ``config.chain`` in turn; a failed one hands its best vector to the next
    (homotopy.py:470-567)."""
    solve = _Solver(nr_solve, vsource_vals, isource_vals, Q_prev, trace)
    X = V_init
    for name in config.chain:
        if name in ("gdev", "gshunt"):
            r = _conductance_ramp(solve, X, 1.0, config, name)
        elif name == "src":
            r = _source_ramp(solve, X, config)
        else:
            continue
        if r.converged:
            return replace(r, method=f"chain_{r.method}", iterations=solve.iterations, homotopy_steps=solve.solves)
        X = r.V
    return HomotopyResult(False, X, "chain_failed", solve.iterations, solve.solves)


def device_solver(engine, reanalyze: bool = True) -> Callable:
    """``nr_solve`` for the continuation on a ``B200Engine``: the solution vector of one solve is handed to the
    next as a device tensor; only (iterations, converged, max|f|) cross to the host per solve.

    ``reanalyze``: the GPU LU keeps the pivot order of its analysis (fixed pattern, no pivoting in the numeric
    kernels), and a continuation moves gmin / gshunt over ten decades: a pivot that was 1e-3 S at the first solve is
    1e-12 S at the last.  With a stale order the ring oscillator's solves drift to |X| ~ 1e30 while still reporting
    convergence (the reference's UMFPACK / cuDSS pivot anew at every factorisation).  So every solve of the
    continuation starts with a new analysis at its own first Jacobian (``B200Engine.lu_reanalyze_at``: one build +
    the host analysis, milliseconds for the circuits that need a DC continuation)."""
    def nr_solve(X, vs, isrc, Q_prev, c0, gmin, gshunt, c1, d1, dQdt_prev, c2, Q_prev2):
        if reanalyze:
            engine.lu_reanalyze_at(X, vs, isrc, Q_prev, c0, gmin, gshunt)
        return engine.nr_solve(X, vs, isrc, Q_prev, c0, gmin, gshunt, c1, d1, dQdt_prev, c2, Q_prev2, None, None,
                               return_device=True)
    return nr_solve


def compute_dc_operating_point(engine, config: Optional[HomotopyConfig] = None, X_init=None, noi_indices: Optional[Sequence[int]] = None,
                               nr_solve: Optional[Callable] = None, trace: Optional[list] = None):
    """Plain Newton from the mid-rail start, the continuation chain if that fails, NOI nodes clamped to 0 V at the end
    (dc_operating_point.py:128-242).  ``engine`` provides ``init_mid_rail()``, ``eval_sources(0)``, ``n_unknowns``,
    ``n_total``, ``n_vs`` and (unless ``nr_solve`` is given) ``nr_solve``: a ``B200Engine`` or the oracle's engine.
    Returns ``(X, info)``: X = [V (n_total), branch currents (n_vs)] as a host array; info = dict(converged, method,
    iterations, homotopy_steps)."""
    cfg = config if config is not None else HomotopyConfig.from_options(engine.options)
    vs, isrc = (np.asarray(a, dtype=np.float64) for a in engine.eval_sources(0.0))
    nu = engine.n_unknowns
    if X_init is None:
        X_init = np.zeros(engine.n_total + engine.n_vs)
        X_init[:engine.n_total] = engine.init_mid_rail()
    if nr_solve is None:
        nr_solve = device_solver(engine) if hasattr(engine, "handle") else engine.nr_solve
    Qz = np.zeros(nu)
    # the plain solve runs at nr_solve's default gmin of 1e-12, not options.gmin (dc_operating_point.py:170-177)
    out = nr_solve(X_init, vs, isrc, Qz, 0.0, 1e-12, 0.0, 0.0, 0.0, None, 0.0, None)
    if trace is not None:
        trace.append((1.0, 1e-12, 0.0, int(out[1]), bool(out[2])))
    if bool(out[2]):
        X, info = out[0], dict(converged=True, method="direct", iterations=int(out[1]), homotopy_steps=0)
    else:
        r = run_homotopy_chain(nr_solve, X_init, vs, isrc, Qz, cfg, trace)
        # a chain that fails still returns its best vector: a metastable circuit starts the transient from it
        X, info = r.V, dict(converged=r.converged, method=r.method, iterations=int(out[1]) + r.iterations,
                            homotopy_steps=r.homotopy_steps, final_gmin=r.final_gmin, final_source_scale=r.final_source_scale)
    X = np.array(X.cpu().numpy() if hasattr(X, "cpu") else X, dtype=np.float64)
    noi = getattr(engine, "noi_indices", None) if noi_indices is None else noi_indices
    if noi is not None and len(noi):
        X[np.asarray(noi, dtype=np.int64)] = 0.0
    return X, info
