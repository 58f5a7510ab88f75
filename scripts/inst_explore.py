"""One sweep instance alone under solver-parameter variants: python scripts/inst_explore.py <sweep index> "k=v k=v" "k=v" ..."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from score_b200 import _lib, build, generators
build.build()
from score_b200.lowering import lower_manhattan_arrays
from score_b200.solver import ScoreSolver
g = int(sys.argv[1])
p1 = lower_manhattan_arrays(generators.manhattan_2d_arrays(generators.MC_BASE_SEED + g, n_robots=20, n_steps=100), "QCQP", with_names=False)
with ScoreSolver(p1) as s:
    for spec in [""] + sys.argv[2:]:
        kw = {}
        for a in spec.split():
            k, v = a.split("=")
            kw[k] = float(v) if "." in v or "e" in v else int(v)
        st = s.solve(**kw)
        r = st.instances[0]
        print(f"[{spec:40s}] solved {r['solved']} newton {r['newton_iters']:4d} cg {r['cg_iters']:5d} kkt {r['rel_kkt']:.2e} "
              f"r_stat {r['r_stat']:.2e} r_gap {r['r_gap']:.2e} f {r['objective']:.9f} lsfail {r['ls_failures']}", flush=True)
