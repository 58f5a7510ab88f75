"""Fused PCG kernel against lockstep ticks: bit-identical per-instance results, and timings of both modes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from score_b200 import build, generators
build.build()
from score_b200.graph_io import load_graph_npz
from score_b200.lowering import concat, lower_factor_graph, lower_manhattan_arrays
from score_b200.solver import ScoreSolver

for name in ["man1", "goats", "man4", "mc0"]:
    fg, extra = load_graph_npz(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    with ScoreSolver(lower_factor_graph(fg)) as s:
        out = {}
        for mode, thr in (("lockstep", 0), ("fused", 1)):
            s.solve(tail_threshold=thr, operator_mode=1)
            st = s.solve(tail_threshold=thr, operator_mode=1)
            out[mode] = (st, [a.copy() for a in s.solution()])
            r = st.instances[0]
            print(f"{name} {mode}: solved={r['solved']} newton={r['newton_iters']} cg={r['cg_iters']} kkt={r['rel_kkt']:.2e} "
                  f"solve_ms={st.solve_ms:.3f} total_ms={st.total_ms:.3f} cycles={st.cycles}", flush=True)
        same = all(np.array_equal(a, b) for a, b in zip(out["lockstep"][1], out["fused"][1]))
        print(f"   bit-identical: {same}")
probs = [lower_manhattan_arrays(generators.manhattan_2d_arrays(generators.MC_BASE_SEED + i, n_robots=20, n_steps=100), "QCQP",
                                with_names=False) for i in (838, 201, 173, 276, 462, 963, 181, 889, 551)]
with ScoreSolver(concat(probs)) as s:
    for mode, thr in (("lockstep", 0), ("fused", 64), ("fused<=4", 4)):
        s.solve(tail_threshold=thr, operator_mode=1)
        st = s.solve(tail_threshold=thr, operator_mode=1)
        sol = [a.copy() for a in s.solution()]
        if mode == "lockstep":
            ref = sol
        print(f"9 slow/typical instances {mode}: solved={st.n_solved} solve_ms={st.solve_ms:.2f} cycles={st.cycles} "
              f"bit-identical={all(np.array_equal(a, b) for a, b in zip(ref, sol))}", flush=True)
