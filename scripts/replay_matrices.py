#!/usr/bin/env python
"""Replay a captured NR matrix sequence (reference scripts/capture_benchmark_matrices.py layout, or a
jacobian.npz of scripts/extract_c6288_jacobian.py) through the GPU sparse LU.

    python scripts/replay_matrices.py <directory | jacobian.npz> [--reanalyze-every K]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vajax_b200.analysis import matrix_sequence as ms  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("path")
ap.add_argument("--reanalyze-every", type=int, default=0, help="redo the symbolic analysis / pivot order every K systems")
a = ap.parse_args()
seq = ms.load_jacobian_npz(a.path) if a.path.endswith(".npz") else ms.load_sequence(a.path)
print(f"{len(seq['data'])} systems, n = {seq['n']}, pattern nnz = {len(seq['indices'])}")
res = ms.replay(seq, reanalyze_every=a.reanalyze_every)
for r in res:
    print(f"  system {r['id']:4d}: scaled residual {r['residual']:.3e}, {r['wall_ms']:.3f} ms (host wall, incl. H2D/D2H)")
print(f"worst residual {max(r['residual'] for r in res):.3e}, median call {np.median([r['wall_ms'] for r in res]):.3f} ms")
