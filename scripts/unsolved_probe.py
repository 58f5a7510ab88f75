"""Which sweep instances of shards [S0, S1) do not reach the tolerance, with their iteration records and a trace:
    python scripts/unsolved_probe.py S0 S1 [key=value solver params]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from score_b200 import _lib, build, generators
build.build()
from score_b200.lowering import lower_manhattan_arrays
from score_b200.solver import ScoreSolver
s0, s1 = int(sys.argv[1]), int(sys.argv[2])
kw = {}
for a in sys.argv[3:]:
    k, v = a.split("=")
    kw[k] = float(v) if "." in v or "e" in v else int(v)
bad = []
for sh in range(s0, s1):
    prob = bench.make_batch(sh * 1024, 1024, 20, 100)
    with ScoreSolver(prob) as s:
        st = s.solve(**kw)
    I = st.instances
    un = np.nonzero(I["rel_kkt"] > 1e-6)[0]
    print(f"shard {sh}: solved {st.n_solved}/1024 cycles {st.cycles} max kkt {I['rel_kkt'].max():.2e} unsolved {[sh * 1024 + int(u) for u in un]}", flush=True)
    bad += [sh * 1024 + int(u) for u in un]
for g in bad[:3]:
    p1 = lower_manhattan_arrays(generators.manhattan_2d_arrays(generators.MC_BASE_SEED + g, n_robots=20, n_steps=100), "QCQP", with_names=False)
    with ScoreSolver(p1) as s:
        st = s.solve(verbose=2, **kw)
        r = st.instances[0]
        tr = s.internal(_lib.SCORE_INT_TRACE, 0).reshape(-1, 8)[: int(r["newton_iters"]) + 1]
        print(f"inst {g} alone: solved {st.n_solved} newton {r['newton_iters']} cg {r['cg_iters']} kkt {r['rel_kkt']:.2e} lsfail {r['ls_failures']}")
        print("   mu  :", " ".join(f"{v:.0e}" for v in tr[:, 0]))
        print("   cg  :", " ".join(str(int(v)) for v in tr[:, 2]))
        print("   step:", " ".join(f"{v:.2g}" for v in tr[:, 1]))
        print("   dec :", " ".join(f"{v:.1e}" for v in tr[:, 3]))
        print("   cols 4-7:", [" ".join(f"{v:.2e}" for v in tr[-6:, c]) for c in range(4, 8)])
