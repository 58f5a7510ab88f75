import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from score_b200 import _lib, build
build.build()
from score_b200.solver import ScoreSolver
from test_refine import _graph
for d in (2, 3):
    prob, _ = _graph(d)
    with ScoreSolver(prob) as s:
        s.solve()
        for kw in ({"preconditioner": 1}, {}, {"max_inner": 50}):
            for mo in (10, 30, 100):
                rec, st = s.refine(max_outer=mo, **kw)
                print(f"d={d} {str(kw):30s} max_outer {mo:3d}: cost -> {rec['cost_final'][0]:.6f} outer {rec['outer_iterations'][0]} accepted {rec['accepted_steps'][0]}  {st['refine_ms']:.1f} ms launches {st['kernel_launches']} conv {st['n_converged']}", flush=True)
