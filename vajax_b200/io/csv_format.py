"""CSV waveforms (reference: vajax/io/csv_writer.py:14-56 writer, :59-94 reader): header ``time,<sorted
node names>`` (``frequency`` for AC results), one row per point, every cell ``%.{precision}e``, rows
terminated ``\\r\\n`` (Python's csv default).  Rows are formatted as one block instead of cell by cell."""
from __future__ import annotations

import csv
import io
from pathlib import Path
from typing import Any, Union

import numpy as np


def write_csv(result: Any, output_path: Union[str, Path], precision: int = 9) -> None:
    if hasattr(result, "frequencies"):
        xname, x = "frequency", np.asarray(result.frequencies, dtype=np.float64)
    else:
        xname, x = "time", np.asarray(result.times, dtype=np.float64)
    names = sorted(result.voltages.keys())
    data = np.empty((len(x), 1 + len(names)), dtype=np.float64)
    data[:, 0] = x
    for j, nm in enumerate(names):
        data[:, 1 + j] = np.asarray(result.voltages[nm], dtype=np.float64)[:len(x)]
    head = io.StringIO()
    csv.writer(head).writerow([xname] + names)          # names with separators / quotes are escaped as csv does
    cells = np.char.mod(f"%.{int(precision)}e", data)
    body = "".join(",".join(r) + "\r\n" for r in cells)
    with open(Path(output_path), "w", newline="") as f:
        f.write(head.getvalue() + body)


def read_csv(input_path: Union[str, Path]) -> dict:
    with open(Path(input_path), "r", newline="") as f:
        rd = csv.reader(f)
        header = next(rd)
        rows = [r for r in rd if r]
    data = np.array(rows, dtype=np.float64).reshape(len(rows), len(header))
    out: dict = {"voltages": {nm: data[:, 1 + j].copy() for j, nm in enumerate(header[1:])}}
    out["frequencies" if header[0] == "frequency" else "times"] = data[:, 0].copy()
    return out
