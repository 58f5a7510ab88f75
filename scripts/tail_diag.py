"""Tail diagnosis of the Monte-Carlo sweep: which instances need many more ticks than the median, and why.

    python scripts/tail_diag.py [n_instances] [out.json] [key=value solver params ...]

Solves the sweep once with the per-Newton-step trace on (ScoreParams.verbose = 2) and dumps, for the slowest
instances, the trace (barrier parameter, step, PCG iterations, decrement, ladder shift per Newton step) together
with structural facts of the instance (ranges per robot, robots without landmark ranges, ...).
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "tail_diag.json")
kw = {}
for a in sys.argv[3:]:
    k, v = a.split("=")
    kw[k] = float(v) if "." in v or "e" in v else int(v)
prob = bench.make_batch(0, n, 20, 100)
from score_b200 import _lib, build

build.build()
from score_b200.solver import ScoreSolver

with ScoreSolver(prob) as s:
    st = s.solve(verbose=2, **kw)
    I = st.instances
    tot = I["cg_iters"] + I["newton_iters"]
    order = np.argsort(-tot)
    rep = {"n": n, "solved": int(st.n_solved), "cycles": int(st.cycles), "solve_ms": st.solve_ms,
           "tot_pct": np.percentile(tot, [0, 50, 90, 99, 100]).tolist(), "slow": [], "typical": []}

    def describe(i):
        tr = s.internal(_lib.SCORE_INT_TRACE, int(i)).reshape(-1, 8)
        nn = int(I[i]["newton_iters"]) + 1
        tr = tr[:nn]
        k0, k1 = prob.rng_off[i], prob.rng_off[i + 1]
        P = prob.pose_off[i + 1] - prob.pose_off[i]
        a, b = prob.rng_a[k0:k1], prob.rng_b[k0:k1]
        robot = lambda o: np.where(o < P, o // 100, -1)
        ra, rb = robot(a), robot(b)
        per_robot_lm = np.array([int(((ra == r) & (rb < 0)).sum() + ((rb == r) & (ra < 0)).sum()) for r in range(20)])
        per_robot_rr = np.array([int((((ra == r) | (rb == r)) & (ra >= 0) & (rb >= 0)).sum()) for r in range(20)])
        return {
            "inst": int(i), "newton": int(I[i]["newton_iters"]), "cg": int(I[i]["cg_iters"]),
            "ls_fail": int(I[i]["ls_failures"]), "kkt": float(I[i]["rel_kkt"]), "K": int(k1 - k0),
            "ranges_to_landmarks_per_robot": per_robot_lm.tolist(), "ranges_robot_robot_per_robot": per_robot_rr.tolist(),
            "zero_dist": int((prob.rng_dist[k0:k1] == 0).sum()),
            "trace_mu_step_cg_dec_shift": [[float(f"{r[0]:.3g}"), float(f"{r[1]:.3g}"), int(r[2]), float(f"{r[3]:.3g}"), int(r[6])]
                                           for r in tr],
        }

    for i in order[:12]:
        rep["slow"].append(describe(i))
    for i in order[n // 2 : n // 2 + 3]:
        rep["typical"].append(describe(i))
with open(out, "w") as f:
    json.dump(rep, f)
print("cycles", rep["cycles"], "solve_ms", rep["solve_ms"], "pct", rep["tot_pct"])
for r in rep["slow"] + rep["typical"]:
    tr = np.array(r["trace_mu_step_cg_dec_shift"])
    print(f"inst {r['inst']}: newton {r['newton']} cg {r['cg']} ls_fail {r['ls_fail']} K {r['K']} zero_dist {r['zero_dist']} "
          f"min lm-ranges/robot {min(r['ranges_to_landmarks_per_robot'])} min rr {min(r['ranges_robot_robot_per_robot'])}")
    print("   cg per newton:", tr[:, 2].astype(int).tolist())
    print("   mu:", [f"{v:.0e}" for v in tr[:, 0]])
    print("   step:", [f"{v:.2g}" for v in tr[:, 1]])
