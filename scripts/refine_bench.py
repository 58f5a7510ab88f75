"""Refinement after the solve on a batch of sweep instances: time, costs, ATE before / after.
    python scripts/refine_bench.py [n_instances]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from score_b200 import build, generators
build.build()
from score_b200.solver import ScoreSolver
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
prob = bench.make_batch(0, n, 20, 100)
gt = np.concatenate([generators.manhattan_2d_arrays(generators.MC_BASE_SEED + i, n_robots=20, n_steps=100)["pos"].reshape(-1, 2) for i in range(n)])
with ScoreSolver(prob) as s:
    st = s.solve()
    before = s.ate(gt)[0]
    rec, stats = s.refine()
    rec, stats = s.refine()
    poses, lms = s.refined()
from score_b200.solver import trajectory_ate
after = trajectory_ate(poses[:, :, 2], gt, traj_off=prob.pose_off)[0]
print(f"{n} instances: solve {st.solve_ms:.1f} ms; refine {stats['refine_ms']:.1f} ms, {stats['outer_iterations']} outer iterations, {stats['kernel_launches']} launches, converged {stats['n_converged']}")
print(f"cost: initial median {np.median(rec['cost_initial']):.1f} -> final median {np.median(rec['cost_final']):.2f}; outer per instance median {np.median(rec['outer_iterations'])} max {rec['outer_iterations'].max()}")
print(f"ATE (m, SE(2)-aligned per instance): relaxed median {np.median(before):.3f} max {before.max():.3f} -> refined median {np.median(after):.3f} max {after.max():.3f}")
