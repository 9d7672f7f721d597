"""ngspice-compatible raw files (reference: vajax/io/rawfile_writer.py:19-119 writer, :122-182 header reader).

Layout the reference writes, byte for byte:

    Title: VAJAX Simulation\\n Date: \\n Plotname: <Transient|AC> Analysis\\n Flags: real\\n
    No. Variables: 1+N\\n No. Points: P\\n Variables:\\n \\t<i>\\t<name>\\t<type>\\n ...
    binary : "Binary:\\n" then P records of (1+N) native float64, x first, nodes in dict order
    ascii  : "Values:\\n" then per point " <i>\\t<x %.15e>\\n" and one "\\t<v %.15e>\\n" per node

Written here as one interleaved (P, 1+N) array -> ``tobytes`` / one formatted block, so a mul64-sized
result (1,110 points x 128 saved nodes, or x 133 k nodes without a save list) is a single buffer write."""
from __future__ import annotations

from pathlib import Path
from typing import Any, Dict, List, Tuple, Union

import numpy as np


def _axis(result: Any) -> Tuple[str, str, np.ndarray]:
    if hasattr(result, "frequencies"):
        return "AC Analysis", "frequency", np.asarray(result.frequencies, dtype=np.float64)
    return "Transient Analysis", "time", np.asarray(result.times, dtype=np.float64)


def _table(result: Any) -> Tuple[str, List[Tuple[str, str]], np.ndarray]:
    plot, xname, x = _axis(result)
    names = list(result.voltages.keys())
    data = np.empty((len(x), 1 + len(names)), dtype=np.float64)
    data[:, 0] = x
    for j, nm in enumerate(names):
        data[:, 1 + j] = np.asarray(result.voltages[nm], dtype=np.float64)[:len(x)]
    variables = [(xname, xname)] + [(nm, "voltage") for nm in names]
    return plot, variables, data


def _header(plot: str, variables: List[Tuple[str, str]], n_points: int) -> str:
    lines = ["Title: VAJAX Simulation", "Date: ", f"Plotname: {plot}", "Flags: real",
             f"No. Variables: {len(variables)}", f"No. Points: {n_points}", "Variables:"]
    lines += [f"\t{i}\t{nm}\t{ty}" for i, (nm, ty) in enumerate(variables)]
    return "\n".join(lines) + "\n"


def write_rawfile(result: Any, output_path: Union[str, Path], binary: bool = True) -> None:
    plot, variables, data = _table(result)
    head = _header(plot, variables, data.shape[0])
    path = Path(output_path)
    if binary:
        with open(path, "wb") as f:
            f.write((head + "Binary:\n").encode("ascii"))
            f.write(np.ascontiguousarray(data).tobytes())
        return
    P, W = data.shape
    tab = np.char.add("\t", np.char.mod("%.15e", data))                     # (P, W) cells "\t<value>"
    first = np.char.add(np.char.add(" ", np.arange(P).astype(str)), tab[:, 0]) if P else tab[:, 0]
    body = "".join(first[i] + "\n" + "".join(v + "\n" for v in tab[i, 1:]) for i in range(P))
    with open(path, "w") as f:
        f.write(head + "Values:\n" + body)


def read_rawfile_header(path: Union[str, Path]) -> Dict[str, Any]:
    """title / plotname / n_vars / n_points / variables [(name, type)] / binary, plus ``data_offset``
    (byte position of the first value; an addition used by ``read_rawfile``)."""
    info: Dict[str, Any] = {"variables": [], "binary": False}
    in_vars = False
    with open(path, "rb") as f:
        while True:
            raw = f.readline()
            if not raw:
                break
            try:
                text = raw.decode("ascii").strip()
            except UnicodeDecodeError:
                break
            low = text.lower()
            if low.startswith("title:"):
                info["title"] = text[6:].strip()
            elif low.startswith("plotname:"):
                info["plotname"] = text[9:].strip()
            elif low.startswith("no. variables:"):
                info["n_vars"] = int(text[14:])
            elif low.startswith("no. points:"):
                info["n_points"] = int(text[11:])
            elif low == "variables:":
                in_vars = True
            elif low in ("binary:", "values:"):
                info["binary"] = low == "binary:"
                info["data_offset"] = f.tell()
                break
            elif in_vars and raw.startswith(b"\t"):
                parts = text.split()
                if len(parts) >= 3:
                    info["variables"].append((parts[1], parts[2]))
    return info


def read_rawfile(path: Union[str, Path]) -> Dict[str, Any]:
    """Round-trip reader (the reference has none): ``{'x_name', 'x', 'voltages': {name: array}}``."""
    info = read_rawfile_header(path)
    P, W = info["n_points"], info["n_vars"]
    with open(path, "rb") as f:
        f.seek(info["data_offset"])
        if info["binary"]:
            data = np.frombuffer(f.read(8 * P * W), dtype=np.float64).reshape(P, W)
        else:
            toks = f.read().split()
            rows = np.array(toks, dtype="S").reshape(P, W + 1)      # point index + W values
            data = rows[:, 1:].astype(np.float64)
    names = [nm for nm, _ in info["variables"]]
    return {"x_name": names[0], "x": data[:, 0].copy(),
            "voltages": {nm: data[:, 1 + j].copy() for j, nm in enumerate(names[1:])}}
