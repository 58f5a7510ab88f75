import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from score_b200 import build
build.build()
from score_b200.graph_io import load_graph_npz
from score_b200.lowering import lower_factor_graph
from score_b200.solver import ScoreSolver
for name in ["man1", "goats", "man4"]:
    fg, extra = load_graph_npz(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    with ScoreSolver(lower_factor_graph(fg)) as s:
        for spec in sys.argv[1:] or [""]:
            kw = {}
            for a in spec.split():
                k, v = a.split("=")
                kw[k] = float(v) if "." in v or "e" in v else int(v)
            st = s.solve(**kw)
            r = st.instances[0]
            print(f"{name} [{spec}] solved={r['solved']} newton={r['newton_iters']} cg={r['cg_iters']} kkt={r['rel_kkt']:.2e} lsfail={r['ls_failures']} f={r['objective']:.10f} ms={st.solve_ms:.1f}", flush=True)
