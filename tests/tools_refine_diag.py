"""Diagnostic (kept under tests/ because it executes the oracle): device gradient / Hessian-vector product of the
refinement against finite differences of the CPU restatement, and the cost after 1 .. 200 Levenberg-Marquardt iterations.
    python tests/tools_refine_diag.py   (GPU box)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from score_b200 import _lib, build
build.build()
from oracle import refine_oracle as ro
from score_b200.solver import ScoreSolver
from test_refine import _graph
for d in (2, 3):
    prob, _ = _graph(d)
    with ScoreSolver(prob) as s:
        s.solve()
        relaxed, rounded, lms0, _ = s.solution()
        poses0 = relaxed.copy(); poses0[:, :, :d] = rounded
        lam = 1e-4
        rec, st = s.refine(max_outer=1, max_inner=1, lambda0=lam)
        g = ro.from_device_slots(prob, s.internal(_lib.SCORE_INT_REF_GRAD))
        p = ro.from_device_slots(prob, s.internal(_lib.SCORE_INT_REF_DIR))
        q = ro.from_device_slots(prob, s.internal(_lib.SCORE_INT_REF_HDIR))
        dg = ro.from_device_slots(prob, s.internal(_lib.SCORE_INT_REF_DIAG))
        J = ro.tangent_jacobian(prob, poses0, lms0)
        r = ro.residuals(prob, poses0, lms0)
        g_ref, H = J.T @ r, J.T @ J
        print(f"d={d}: |g-g_ref|/|g_ref| = {np.linalg.norm(g - g_ref) / np.linalg.norm(g_ref):.2e}   |dg-diag H|/|diag| = {np.linalg.norm(dg - np.diag(H)) / np.linalg.norm(np.diag(H)):.2e}")
        q_ref = H @ p + lam * np.diag(H) * p
        print(f"      |q-q_ref|/|q_ref| = {np.linalg.norm(q - q_ref) / np.linalg.norm(q_ref):.2e}  |p| {np.linalg.norm(p):.3e}")
        bad = np.argsort(-np.abs(g - g_ref))[:6]
        print("      worst g entries", [(int(b), float(g[b]), float(g_ref[b])) for b in bad])
        badq = np.argsort(-np.abs(q - q_ref))[:6]
        print("      worst q entries", [(int(b), float(q[b]), float(q_ref[b])) for b in badq])
        for mo in (1, 2, 5, 10, 20, 50, 200):
            rec, st = s.refine(max_outer=mo)
            print(f"      max_outer {mo:3d}: cost {rec['cost_initial'][0]:.6f} -> {rec['cost_final'][0]:.6f} outer {rec['outer_iterations'][0]} accepted {rec['accepted_steps'][0]}")
        _, _, f_ref = ro.refine(prob, poses0, lms0)
        print("      scipy cost", f_ref)
