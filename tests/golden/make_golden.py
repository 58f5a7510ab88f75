"""Regenerate tests/golden/*.npz from the graphs shipped with the reference.

Run in the build container (needs /root/reference):
    python tests/golden/make_golden.py
Writes, per graph, the portable array form of the FactorGraphData (inputs) plus
the oracle's tight optimum (outputs): x*, f*, structural checksums of the
assembled least-squares matrix.  The GPU box has no /root/reference, so the
`-m gpu` parity tests read these files only.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import score_b200  # noqa: E402,F401
from oracle import score_oracle as so  # noqa: E402
from py_factor_graph.parsing.parse_pickle_file import parse_pickle_file  # noqa: E402
from score_b200 import generators  # noqa: E402
from score_b200.graph_io import robot_subgraph, save_graph_npz  # noqa: E402

REF = "/root/reference/examples"
HERE = os.path.dirname(os.path.abspath(__file__))


def sha16(a):
    return hashlib.sha256(np.asarray(a).astype("<i4").tobytes()).hexdigest()[:16]


def emit(name, fg):
    prob = so.assemble(fg, so.QCQP)
    sol = so.solve_qcqp_barrier(prob)
    x = so.polish_distances(prob, sol.x)
    kkt = so.kkt_qcqp(prob, x)
    B = prob.B
    save_graph_npz(
        fg,
        os.path.join(HERE, name + ".npz"),
        x_star=x,
        f_star=so.objective(prob, x),
        rel_kkt=kkt["rel_kkt"],
        newton_steps=sol.newton_steps,
        shape=np.asarray(B.shape),
        nnz=B.nnz,
        sha_indptr=sha16(B.indptr),
        sha_indices=sha16(B.indices),
        sum_values=B.data.sum(),
        sum_abs_values=np.abs(B.data).sum(),
        sum_w=prob.w.sum(),
    )
    print(name, B.shape, B.nnz, sha16(B.indptr), sha16(B.indices), so.objective(prob, x), kkt["rel_kkt"], sol.newton_steps)


if __name__ == "__main__":
    goats = parse_pickle_file(os.path.join(REF, "goats_14_data", "goats_14_6_2002_15_20.pkl"))
    man4 = parse_pickle_file(os.path.join(REF, "manhattan", "factor_graph.pickle"))
    emit("goats", goats)
    emit("man4", man4)
    emit("man1", robot_subgraph(man4, 0))
    emit("mc0_small", generators.manhattan_2d(generators.MC_BASE_SEED, n_robots=4, n_steps=40))
    emit("mc0", generators.monte_carlo_instance(0))
