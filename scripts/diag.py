"""Quick diagnostic: solve golden graphs and print per-instance stats."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from score_b200 import build
build.build()
from score_b200.graph_io import load_graph_npz
from score_b200.lowering import lower_factor_graph
from score_b200.solver import ScoreSolver

tol = float(os.environ.get("KKT", "1e-6"))
names = sys.argv[1:] or ["mc0_small", "man1", "goats", "man4", "mc0"]
for name in names:
    fg, extra = load_graph_npz(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    t0 = time.time()
    with ScoreSolver(lower_factor_graph(fg)) as s:
        t1 = time.time()
        st = s.solve(kkt_tol=tol)
        t2 = time.time()
        st2 = s.solve(kkt_tol=tol)
        poses, rounded, lms, dist = s.solution()
    rec = st.instances[0]
    print(f"{name}: solved={rec['solved']} f={rec['objective']:.9f} f*={float(extra['f_star']):.9f} kkt={rec['rel_kkt']:.3e} "
          f"newton={rec['newton_iters']} cg={rec['cg_iters']} lsfail={rec['ls_failures']} ticks={st.ticks} "
          f"asm={st.assemble_ms:.2f}ms setup={st.setup_ms:.2f}ms solve={st.solve_ms:.2f}ms (2nd {st2.solve_ms:.2f}ms) "
          f"create={t1-t0:.3f}s wall_solve={t2-t1:.3f}s GB/s={st2.algorithmic_bytes/st2.solve_ms/1e6:.1f}", flush=True)
