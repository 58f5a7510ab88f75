"""Per-job timeline of the streamed end-to-end path (ScoreSolverGroup.run_pipelined): when each sub-batch job's
score_create / score_solve / read-back / destroy started and ended.

    python scripts/e2e_trace.py [n_instances] [streams] [steps] [extra_threads]
"""
import dataclasses
import os
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
streams = int(sys.argv[2]) if len(sys.argv) > 2 else 2
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
extra = int(sys.argv[4]) if len(sys.argv) > 4 else 1
prob = bench.make_batch(0, n, 20, 100)
import torch

from score_b200 import build

build.build()
from score_b200.solver import ScoreSolver, ScoreSolverGroup

pinned = {}
for f in dataclasses.fields(prob):
    v = getattr(prob, f.name)
    if isinstance(v, np.ndarray):
        pinned[f.name] = torch.from_numpy(np.ascontiguousarray(v)).pin_memory()
prob = dataclasses.replace(prob, **{k: t.numpy() for k, t in pinned.items()})
g = ScoreSolverGroup(prob, n_streams=streams, create=False)
outs = tuple(torch.empty(shp, dtype=torch.float64).pin_memory().numpy() for shp in g._shapes())
views = g._views(outs)
if os.environ.get("E2E_PREWARM", "1") != "0":
    g.prewarm()
g.run_pipelined(out=outs, steps=2)  # warm (same queue depth: the memory pool grows here)
sem = threading.Semaphore(streams)
T0 = time.perf_counter()
now = lambda: 1e3 * (time.perf_counter() - T0)


creating = threading.Lock()


def one(job):
    j = job % len(g.parts)
    with creating:
        t = [now()]
        s = ScoreSolver(g.parts[j])
        t.append(now())
    with sem:
        t.append(now())
        st = s.solve()
        t.append(now())
    s.solution(out=views[j])
    t.append(now())
    s.close()
    t.append(now())
    return job, t, st.total_ms


with ThreadPoolExecutor(streams + extra) as pool:
    res = list(pool.map(one, range(steps * len(g.parts))))
torch.cuda.synchronize()
total = now()
for job, t, dev in res:
    print(f"job {job}: create {t[0]:7.1f}-{t[1]:7.1f}  wait {t[2]-t[1]:6.1f}  solve {t[2]:7.1f}-{t[3]:7.1f} (device {dev:6.1f})  "
          f"read-back {t[4]-t[3]:5.1f}  destroy {t[5]-t[4]:5.1f}")
print(f"total {total:.1f} ms for {steps} steps x {n} instances -> {steps * n / total * 1e3:.1f} solves/s")
