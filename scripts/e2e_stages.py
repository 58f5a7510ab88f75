"""Stage timing of the end-to-end path (score_create / score_solve / score_get_solution / score_destroy) for one
sub-batch, from pinned host arrays — where the host-side time of bench.py's e2e figure goes.

    python scripts/e2e_stages.py [n_instances] [repeats]
"""
import dataclasses
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
rep = int(sys.argv[2]) if len(sys.argv) > 2 else 3
prob = bench.make_batch(0, n, 20, 100)
import torch

from score_b200 import build

build.build()
from score_b200.solver import ScoreSolver

pinned = {}
for f in dataclasses.fields(prob):
    v = getattr(prob, f.name)
    if isinstance(v, np.ndarray):
        pinned[f.name] = torch.from_numpy(np.ascontiguousarray(v)).pin_memory()
prob = dataclasses.replace(prob, **{k: t.numpy() for k, t in pinned.items()})
with ScoreSolver(prob) as s:
    s.solve()
    outs = tuple(torch.empty(shp, dtype=torch.float64).pin_memory().numpy() for shp in s.solution_shapes())
for r in range(rep):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    s = ScoreSolver(prob)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    st = s.solve()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    s.solution(out=outs)
    t3 = time.perf_counter()
    s.close()
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    print(f"n={n} create {1e3*(t1-t0):.1f} ms  solve {1e3*(t2-t1):.1f} ms (device total {st.total_ms:.1f}: asm {st.assemble_ms:.1f} setup {st.setup_ms:.1f} "
          f"solve {st.solve_ms:.1f} extract {st.extract_ms:.1f})  read-back {1e3*(t3-t2):.1f} ms  destroy {1e3*(t4-t3):.1f} ms  "
          f"h2d {s.h2d_bytes/1e6:.0f} MB d2h {s.d2h_bytes/1e6:.0f} MB", flush=True)
