"""Matrix-free PCG operator (k_hessvec) against the assembled CSR pair: same iteration counts / objective to rounding,
and the timing of both on the sweep batch."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from score_b200 import build, generators
build.build()
from score_b200.graph_io import load_graph_npz
from score_b200.lowering import lower_factor_graph, lower_grid3d_arrays
from score_b200.solver import ScoreSolver

def one(prob, label):
    with ScoreSolver(prob) as s:
        res = {}
        for mode in (1, 0):
            st = s.solve(operator_mode=mode)
            st = s.solve(operator_mode=mode)
            res[mode] = (st, [a.copy() for a in s.solution()])
            r = st.instances[0]
            print(f"{label} mode={'csr' if mode else 'matrix-free'}: solved={st.n_solved} newton={r['newton_iters']} cg={r['cg_iters']} "
                  f"f={r['objective']:.12f} kkt={r['rel_kkt']:.2e} solve_ms={st.solve_ms:.3f}", flush=True)
        dp = np.abs(res[0][1][0] - res[1][1][0]).max()
        print(f"   max |pose diff| = {dp:.3e}  rel obj diff = "
              f"{abs(res[0][0].instances[0]['objective'] - res[1][0].instances[0]['objective']) / max(1, abs(res[1][0].instances[0]['objective'])):.2e}")

for name in ["man1", "goats", "man4"]:
    fg, _ = load_graph_npz(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    one(lower_factor_graph(fg), name)
one(lower_grid3d_arrays(generators.grid_3d_arrays(11, n_robots=3, n_steps=40, grid=8, n_landmarks=5, n_ranges=260)), "grid3d")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
prob = bench.make_batch(0, n, 20, 100)
with ScoreSolver(prob) as s:
    for mode in (1, 0):
        s.solve(operator_mode=mode)
        st = s.solve(operator_mode=mode)
        I = st.instances
        print(f"sweep {n} mode={'csr' if mode else 'matrix-free'}: solved={st.n_solved} solve_ms={st.solve_ms:.1f} cycles={st.cycles} "
              f"newton p50/max {np.percentile(I['newton_iters'],50):.0f}/{I['newton_iters'].max()} cg mean {I['cg_iters'].mean():.1f} "
              f"max {I['cg_iters'].max()} maxkkt {I['rel_kkt'].max():.2e}", flush=True)
