#!/bin/bash
# compute-sanitizer memcheck + racecheck on a single small graph (man1) and a 7-instance batch.
# Usage (GPU box): bash scripts/sanitize.sh [out_dir]
out=${1:-gpurun_out}
mkdir -p "$out"
cat > /tmp/sanitize_target.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
from score_b200 import generators
from score_b200.graph_io import load_graph_npz
from score_b200.lowering import concat, lower_factor_graph, lower_manhattan_arrays
from score_b200.solver import ScoreSolver
fg, _ = load_graph_npz("tests/golden/man1.npz")
with ScoreSolver(lower_factor_graph(fg)) as s:
    st = s.solve()
    print("man1 solved", st.n_solved, "newton", st.instances[0]["newton_iters"], "cg", st.instances[0]["cg_iters"])
probs = [lower_manhattan_arrays(generators.manhattan_2d_arrays(generators.MC_BASE_SEED + i, n_robots=4, n_steps=30)) for i in range(7)]
with ScoreSolver(concat(probs)) as s:
    st = s.solve()
    s.solution()
    print("batch solved", st.n_solved, "of 7 (matrix-free operator)")
    st = s.solve(operator_mode=1)
    print("batch solved", st.n_solved, "of 7 (assembled CSR pair)")
    st = s.solve(operator_mode=1, tail_threshold=4)
    print("batch solved", st.n_solved, "of 7 (fused per-instance PCG kernel for the last 4)")
from score_b200.lowering import lower_grid3d_arrays
big = lower_grid3d_arrays(generators.grid_3d_arrays(5, n_robots=12, n_steps=12, grid=8, n_landmarks=20, n_ranges=1500))
with ScoreSolver(big) as s:
    st = s.solve()
    print("3D graph with a large coarse space (Schur complement + blocked sweeps) solved", st.n_solved, "newton", st.instances[0]["newton_iters"])
    rec, rs = s.refine(max_outer=6)
    print("3D refinement (chain preconditioner): cost", rec["cost_initial"][0], "->", rec["cost_final"][0])
with ScoreSolver(concat(probs)) as s:
    s.solve()
    rec, rs = s.refine(max_outer=6)
    rec2, _ = s.refine(max_outer=3, preconditioner=1, max_inner=20)
    print("batch refinement: accepted steps", rec["accepted_steps"].tolist(), "block-Jacobi", rec2["accepted_steps"].tolist())
PY
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python /tmp/sanitize_target.py > "$out/sanitizer_$tool.log" 2>&1
  echo "$tool exit $?" >> "$out/sanitizer_$tool.log"
  tail -5 "$out/sanitizer_$tool.log"
done
