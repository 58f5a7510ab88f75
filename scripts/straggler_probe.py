"""Find and trace the slowest instance of a sweep shard: FIRST=<index> python scripts/straggler_probe.py [n]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from score_b200 import _lib, build, generators
build.build()
from score_b200.lowering import lower_manhattan_arrays
from score_b200.solver import ScoreSolver
first = int(os.environ.get("FIRST", "0"))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
prob = bench.make_batch(first, n, 20, 100)
with ScoreSolver(prob) as s:
    st = s.solve()
I = st.instances
tot = I["cg_iters"] + I["newton_iters"]
worst = np.argsort(-tot)[:4]
print("cycles", st.cycles, "worst", [(first + int(w), int(I[w]["newton_iters"]), int(I[w]["cg_iters"])) for w in worst], flush=True)
w = int(worst[0])
p1 = lower_manhattan_arrays(generators.manhattan_2d_arrays(generators.MC_BASE_SEED + first + w, n_robots=20, n_steps=100), "QCQP", with_names=False)
with ScoreSolver(p1) as s:
    for spec in [""] + sys.argv[2:]:
        kw = {}
        for a in spec.split():
            k, v = a.split("=")
            kw[k] = float(v) if "." in v or "e" in v else int(v)
        st = s.solve(verbose=2, **kw)
        r = st.instances[0]
        tr = s.internal(_lib.SCORE_INT_TRACE, 0).reshape(-1, 8)[: int(r["newton_iters"]) + 1]
        print(f"[{spec}] inst {first + w}: newton {r['newton_iters']} cg {r['cg_iters']} kkt {r['rel_kkt']:.2e} lsfail {r['ls_failures']}")
        print("   mu  :", " ".join(f"{v:.0e}" for v in tr[:, 0]))
        print("   cg  :", " ".join(str(int(v)) for v in tr[:, 2]))
        print("   step:", " ".join(f"{v:.2g}" for v in tr[:, 1]))
        print("   dec :", " ".join(f"{v:.1e}" for v in tr[:, 3]))
