"""Named-waveform view of what ``B200Engine.run_transient`` returns.

Reference: ``TransientResult`` (vajax/analysis/engine.py:116-170) and ``extract_results``
(vajax/analysis/transient/full_mna.py:143-175): the first ``n_steps`` rows of the fixed-size output
buffers, one array per node / source name.  With a device-side save list (``set_save_list``, N4) the
columns of ``V_out`` are the saved nodes only and ``stats['node_indices']`` already maps names to them."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict, Tuple

import numpy as np


def extract_results(times, V_out, stats: Dict[str, Any]) -> Tuple[np.ndarray, Dict[str, np.ndarray], Dict[str, np.ndarray]]:
    n = int(stats["n_steps"])
    t = np.asarray(times)[:n]
    V = np.asarray(V_out)[:n]
    voltages = {name: V[:, col] for name, col in stats["node_indices"].items()}
    currents: Dict[str, np.ndarray] = {}
    if "I_out" in stats and stats.get("current_indices"):
        I = np.asarray(stats["I_out"])[:n]
        currents = {name: I[:, col] for name, col in stats["current_indices"].items()}
    return t, voltages, currents


@dataclass
class TransientResult:
    times: np.ndarray
    voltages: Dict[str, np.ndarray]
    currents: Dict[str, np.ndarray] = field(default_factory=dict)
    stats: Dict[str, Any] = field(default_factory=dict)

    @classmethod
    def from_run(cls, times, V_out, stats: Dict[str, Any]) -> "TransientResult":
        t, v, i = extract_results(times, V_out, stats)
        return cls(times=t, voltages=v, currents=i, stats=stats)

    @property
    def num_steps(self) -> int:
        return len(self.times)

    def _lookup(self, table: Dict[str, np.ndarray], name: str, what: str) -> np.ndarray:
        if name in table:
            return table[name]
        low = name.lower()
        for k, v in table.items():          # the reference falls back to a case-insensitive match
            if k.lower() == low:
                return v
        raise KeyError(f"{what} '{name}' not found. Available: {list(table.keys())}")

    def voltage(self, node: str) -> np.ndarray:
        return self._lookup(self.voltages, node, "Node")

    def current(self, source: str) -> np.ndarray:
        return self._lookup(self.currents, source, "Source")
