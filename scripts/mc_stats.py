"""Solve N sweep instances on the GPU and dump the per-instance statistics table."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "mc_stats.npy")
kw = {}
for a in sys.argv[3:]:
    k, v = a.split("=")
    kw[k] = float(v) if "." in v or "e" in v else int(v)
prob = bench.make_batch(0, n, 20, 100)
from score_b200 import build
build.build()
from score_b200.solver import ScoreSolver
with ScoreSolver(prob) as s:
    st = s.solve(**kw)
    st = s.solve(**kw)
np.save(out, st.instances)
I = st.instances
print("solved", st.n_solved, "of", n, "ticks", st.ticks, "solve_ms", st.solve_ms, "GB/s", st.algorithmic_bytes / st.solve_ms / 1e6)
tot = I["cg_iters"] + I["newton_iters"]
print("ticks/instance percentiles", np.percentile(tot, [0, 10, 50, 90, 99, 100]))
print("newton percentiles", np.percentile(I["newton_iters"], [0, 10, 50, 90, 99, 100]))
print("unsolved:", np.nonzero(I["solved"] == 0)[0].tolist())
print("ls_failures>0:", int((I["ls_failures"] > 0).sum()))
bad = np.argsort(-tot)[:20]
for b in bad:
    print(b, I[b])
