"""Relaxation + refinement on the shipped graphs (fixtures of GOATS-14 and the Manhattan pickle): cost and aligned
trajectory error before / after.   python scripts/refine_fixtures.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from score_b200 import build
build.build()
from score.solve_score import solve_and_refine
from score_b200.evaluate import evaluate_ate
from score_b200.graph_io import load_graph_npz
for name in ("goats", "man1", "man4"):
    fg, _ = load_graph_npz(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    relaxed, refined, rec = solve_and_refine([fg], "QCQP")
    relaxed, refined, rec = solve_and_refine([fg], "QCQP")  # second call: warm caches
    a0, a1 = evaluate_ate(relaxed[0], fg)["rmse"], evaluate_ate(refined[0], fg)["rmse"]
    print(f"{name:6s}: relaxed objective {relaxed[0].solver_cost:.4f}; non-convex cost {rec[0]['cost_initial']:.1f} -> {rec[0]['cost_final']:.3f} "
          f"in {rec[0]['outer_iterations']} LM iterations ({rec[0]['accepted_steps']} accepted); aligned ATE {a0:.3f} m -> {a1:.3f} m; "
          f"solve {relaxed[0].total_time * 1e3:.1f} ms, solve + refine {refined[0].total_time * 1e3:.1f} ms")
