"""Probe of the pipelined step schedule: per-part solve times alone, together, and out of phase."""
import os, sys, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from score_b200 import build
build.build()
from score_b200.solver import ScoreSolverGroup
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
kw = {}
for a in sys.argv[2:]:
    k, v = a.split("=")
    kw[k] = float(v) if "." in v or "e" in v else int(v)
prob = bench.make_batch(0, n, 20, 100)
g = ScoreSolverGroup(prob, n_streams=2)
for s in g.solvers:
    s.solve(**kw)
for j, s in enumerate(g.solvers):
    t0 = time.perf_counter(); st = s.solve(**kw); dt = time.perf_counter() - t0
    print(f"part {j} alone: wall {1e3*dt:.1f} ms solve_ms {st.solve_ms:.1f} cycles {st.cycles}", flush=True)
t0 = time.perf_counter(); st = g.solve(**kw); dt = time.perf_counter() - t0
print(f"both together (in phase): wall {1e3*dt:.1f} ms", flush=True)
log = [[], []]
def run(j, steps, delay):
    time.sleep(delay)
    for k in range(steps):
        t0 = time.perf_counter(); st = g.solvers[j].solve(**kw); log[j].append((t0, time.perf_counter(), st.solve_ms))
for delay in (0.0, 0.07):
    log = [[], []]
    T0 = time.perf_counter()
    th = [threading.Thread(target=run, args=(j, 6, j * delay)) for j in range(2)]
    [t.start() for t in th]; [t.join() for t in th]
    tot = time.perf_counter() - T0
    print(f"pipelined 6 steps, offset {delay*1e3:.0f} ms: total {1e3*tot:.1f} ms = {1e3*tot/6:.1f} ms/step")
    for j in range(2):
        print("   part", j, " ".join(f"[{1e3*(a-T0):.0f}-{1e3*(b-T0):.0f}|{c:.0f}]" for a, b, c in log[j]))
g.close()

# ---- two handle sets (double-buffered steps), 2 parts each: 4 sub-batch solves in flight
for nsets, nparts in ((2, 2), (3, 2), (2, 4), (4, 2)):
    groups = [ScoreSolverGroup(prob, n_streams=nparts) for _ in range(nsets)]
    for gg in groups:
        gg.solve(**kw)
    steps = 12
    def runp(gg, j, k, delay):
        time.sleep(delay)
        for _ in range(k):
            gg.solvers[j].solve(**kw)
    T0 = time.perf_counter()
    th = []
    for si, gg in enumerate(groups):
        for j in range(nparts):
            th.append(threading.Thread(target=runp, args=(gg, j, steps // nsets, 0.03 * (si * nparts + j))))
    [t.start() for t in th]; [t.join() for t in th]
    tot = time.perf_counter() - T0
    print(f"{nsets} handle sets x {nparts} parts, {steps} steps: total {1e3*tot:.1f} ms = {1e3*tot/steps:.1f} ms/step", flush=True)
    for gg in groups:
        gg.close()
