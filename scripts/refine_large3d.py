"""Config 5 (one 3D graph, 100 k poses, 1 M ranges): relaxation, then local refinement; cost, time, aligned ATE."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from score_b200 import build, generators
build.build()
from score_b200.lowering import lower_grid3d_arrays
from score_b200.solver import ScoreSolver, trajectory_ate
R_, S_ = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (100, 1000)
L_, K_ = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (1000, 1000000)
arr = generators.grid_3d_arrays(generators.MC_BASE_SEED, n_robots=R_, n_steps=S_, grid=100, n_landmarks=L_, n_ranges=K_)
prob = lower_grid3d_arrays(arr)
gt = arr["pos"].reshape(-1, 3)
with ScoreSolver(prob) as s:
    st = s.solve()
    relaxed = s.solution()[0]
    t0 = time.time()
    rec, rs = s.refine(max_outer=int(os.environ.get("MAX_OUTER", "30")))
    wall = time.time() - t0
    poses, _ = s.refined()
a0 = trajectory_ate(relaxed[:, :, 3], gt)[0][0]
a1 = trajectory_ate(poses[:, :, 3], gt)[0][0]
print(f"P={prob.P} K={prob.K}: solve {st.solve_ms:.0f} ms solved {st.n_solved}; refine {rs['refine_ms']:.0f} ms (wall {wall:.2f} s), {rec['outer_iterations'][0]} LM iterations "
      f"({rec['accepted_steps'][0]} accepted), {rs['kernel_launches']} launches, stopped by tolerance: {rs['n_converged']}; cost {rec['cost_initial'][0]:.4g} -> {rec['cost_final'][0]:.6g}; "
      f"aligned ATE {a0:.3f} m -> {a1:.3f} m")
