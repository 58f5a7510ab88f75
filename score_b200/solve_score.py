"""``solve_score`` — drop-in for /root/reference/score/solve_score.py:54-86.

Same entry point, same ``FactorGraphData`` input and ``SolverResults`` output
layout as the reference; the Gurobi model build + barrier solve + SVD rounding
are replaced by one call chain into libscore_b200 (CUDA, sm_100a).  There is no
CPU fallback: without the built library or a CUDA device the call raises.
"""
from __future__ import annotations

import logging
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import pyfg_shim as _shim

_shim.install()

from py_factor_graph.factor_graph import FactorGraphData  # noqa: E402
from py_factor_graph.utils.solver_utils import SolverResults, VariableValues  # noqa: E402

from .lowering import (  # noqa: E402
    ACCEPTABLE_RELAXATIONS,
    QCQP_RELAXATION,
    SOCP_RELAXATION,
    LoweredProblem,
    check_valid_relaxation,
    concat,
    lower_factor_graph,
)
from .solver import ScoreSolver, SolveStats  # noqa: E402

logger = logging.getLogger(__name__)

#: solver tolerance: relative KKT.  The reference runs Gurobi with
#: BarQCPConvTol = 1e-1 (gurobi_utils.py:212); this solve is far tighter.
DEFAULT_KKT_TOL = 1e-6


def _check_factor_graph(data: FactorGraphData) -> None:
    """solve_score.py:28-32."""
    unconnected_variables = data.unconnected_variable_names
    assert len(unconnected_variables) == 0, f"Found {unconnected_variables} unconnected variables. "


def _split_args(args, relaxation_type):
    """Accept both ``solve_score(data, relaxation)`` (solve_score.py:54-57) and the example's
    ``solve_score(data, solver_params, relaxation)`` (examples/solve_goats_example_score.py:42-44)."""
    if len(args) == 0:
        return relaxation_type
    if len(args) == 1:
        if isinstance(args[0], str):
            return args[0]
        return relaxation_type  # a solver_params object: accepted and ignored
    if len(args) == 2:
        return args[1]
    raise TypeError("solve_score() takes at most 3 positional arguments")


def pack_results(
    prob: LoweredProblem,
    inst: int,
    poses: np.ndarray,
    rounded: np.ndarray,
    lms: np.ndarray,
    dist: np.ndarray,
    total_time: float,
    solved: bool,
    pose_chain_names,
    solver_cost: Optional[float] = None,
) -> SolverResults:
    """extract_solver_results / get_variable_values (gurobi_utils.py:114-136,190-203):
    homogeneous (d+1)x(d+1) poses with the rounded rotation and the untouched translation,
    raw landmarks, raw distance variables ((d,) for QCQP, (1,) for SOCP)."""
    d = prob.dim
    p0, l0, k0 = int(prob.pose_off[inst]), int(prob.lm_off[inst]), int(prob.rng_off[inst])
    pose_vals: Dict[str, np.ndarray] = {}
    for i, name in enumerate(prob.pose_names[inst]):
        T = np.eye(d + 1)
        T[:d, :d] = rounded[p0 + i]
        T[:d, d] = poses[p0 + i, :, d]
        pose_vals[name] = T
    lm_vals = {name: lms[l0 + i].copy() for i, name in enumerate(prob.landmark_names[inst])}
    dist_vals = {key: dist[k0 + i].copy() for i, key in enumerate(prob.range_keys[inst])}
    return SolverResults(
        variables=VariableValues(d, pose_vals, lm_vals, dist_vals),
        total_time=total_time,
        solved=bool(solved),
        pose_chain_names=pose_chain_names,
        solver_cost=solver_cost,
    )


def solve_score(data: FactorGraphData, *args, relaxation_type: str = QCQP_RELAXATION, device: int = 0,
                kkt_tol: float = DEFAULT_KKT_TOL, return_stats: bool = False, **solver_kw):
    """Solve the convex relaxation of range-aided SLAM for ``data`` on the GPU.

    args:
        data (FactorGraphData): the data describing the problem
        relaxation_type (str): "QCQP" (default) or "SOCP" (gurobi_utils.py:26-28)

    returns:
        SolverResults: rounded poses, landmarks, distance variables, solve time, solved flag
    """
    relaxation_type = _split_args(args, relaxation_type)
    _check_factor_graph(data)
    prob = lower_factor_graph(data, relaxation_type)
    with ScoreSolver(prob, device=device) as solver:
        stats = solver.solve(kkt_tol=kkt_tol, **solver_kw)
        poses, rounded, lms, dist = solver.solution()
    rec = stats.instances[0]
    if not rec["solved"]:
        logger.warning("solve_score: not converged (rel KKT %.3e after %d Newton steps)", rec["rel_kkt"],
                       rec["newton_iters"])
    res = pack_results(prob, 0, poses, rounded, lms, dist, stats.total_ms * 1e-3, rec["solved"],
                       data.get_pose_chain_names(), float(rec["objective"]))
    return (res, stats) if return_stats else res


def solve_score_batch(datas: Sequence[FactorGraphData], relaxation_type: str = QCQP_RELAXATION, device: int = 0,
                      kkt_tol: float = DEFAULT_KKT_TOL, return_stats: bool = False, **solver_kw):
    """Solve independent instances as one block-diagonal batch on one GPU."""
    check_valid_relaxation(relaxation_type)
    for data in datas:
        _check_factor_graph(data)
    prob = concat([lower_factor_graph(data, relaxation_type) for data in datas])
    with ScoreSolver(prob, device=device) as solver:
        stats = solver.solve(kkt_tol=kkt_tol, **solver_kw)
        poses, rounded, lms, dist = solver.solution()
    out = []
    for i, data in enumerate(datas):
        rec = stats.instances[i]
        out.append(pack_results(prob, i, poses, rounded, lms, dist, stats.total_ms * 1e-3 / len(datas),
                                rec["solved"], data.get_pose_chain_names(), float(rec["objective"])))
    return (out, stats) if return_stats else out


def solve_and_refine(datas: Sequence[FactorGraphData], relaxation_type: str = QCQP_RELAXATION, device: int = 0,
                     kkt_tol: float = DEFAULT_KKT_TOL, refine_kw: Optional[dict] = None, **solver_kw):
    """SCORE as the initialisation of a local search (/root/reference/README.md:63-67; the paper uses GTSAM for the
    second step): solve the relaxation of every instance as one batch, then refine the rounded estimates on the original
    non-convex cost on the same GPU (``score_refine``: batched Levenberg-Marquardt, csrc/refine.cuh).

    Returns (relaxed, refined, records): two lists of ``SolverResults`` in input order (the refined ones carry the
    refined poses / landmarks; their distance variables are the unit directions (QCQP) or lengths (SOCP) between the
    refined endpoints; ``solver_cost`` is the non-convex cost) and the per-instance refinement records."""
    check_valid_relaxation(relaxation_type)
    for data in datas:
        _check_factor_graph(data)
    prob = concat([lower_factor_graph(data, relaxation_type) for data in datas])
    d = prob.dim
    with ScoreSolver(prob, device=device) as solver:
        stats = solver.solve(kkt_tol=kkt_tol, **solver_kw)
        poses, rounded, lms, dist = solver.solution()
        rec, rstats = solver.refine(**(refine_kw or {}))
        rposes, rlms = solver.refined()
    # distance variables of the refined point: what the eliminated variable equals there
    owners = []
    for i in range(prob.n_instances):
        p0, p1, l0, l1 = prob.pose_off[i], prob.pose_off[i + 1], prob.lm_off[i], prob.lm_off[i + 1]
        own = np.concatenate([rposes[p0:p1, :, d], rlms[l0:l1]], axis=0)
        k0, k1 = prob.rng_off[i], prob.rng_off[i + 1]
        diff = own[prob.rng_a[k0:k1]] - own[prob.rng_b[k0:k1]]
        n = np.linalg.norm(diff, axis=1, keepdims=True)
        owners.append(diff / np.maximum(n, 1e-300) if relaxation_type == QCQP_RELAXATION else n)
    rdist = np.concatenate(owners, axis=0) if owners else dist
    relaxed, refined = [], []
    for i, data in enumerate(datas):
        r = stats.instances[i]
        per = stats.total_ms * 1e-3 / len(datas)
        relaxed.append(pack_results(prob, i, poses, rounded, lms, dist, per, r["solved"], data.get_pose_chain_names(),
                                    float(r["objective"])))
        refined.append(pack_results(prob, i, rposes, np.ascontiguousarray(rposes[:, :, :d]), rlms, rdist,
                                    per + rstats["refine_ms"] * 1e-3 / len(datas), r["solved"], data.get_pose_chain_names(),
                                    float(rec[i]["cost_final"])))
    return relaxed, refined, rec


def solve_problem_with_intermediate_iterates(data: FactorGraphData, relaxation_type: str = QCQP_RELAXATION,
                                             device: int = 0, max_iterates: int = 1000) -> List[SolverResults]:
    """solve_score.py:89-116: the reference re-solves with BarIterLimit = 0, 1, 2, ... and collects
    one SolverResults per cap.  Here the cap is on outer (Newton) iterations; the solver is
    deterministic, so the k-th entry is exactly the k-th iterate of the full solve."""
    logger.warning(
        "Solving the problem with intermediate iterates - this is for debugging or visualization "
        "only as it is much slower than a single solve. Use solve_score() for solving the problem"
    )
    check_valid_relaxation(relaxation_type)
    _check_factor_graph(data)
    prob = lower_factor_graph(data, relaxation_type)
    iterates: List[SolverResults] = []
    with ScoreSolver(prob, device=device) as solver:
        curr_iter = 0
        while curr_iter < max_iterates:
            stats = solver.solve(max_newton=curr_iter if curr_iter > 0 else -1)
            poses, rounded, lms, dist = solver.solution()
            rec = stats.instances[0]
            iterates.append(pack_results(prob, 0, poses, rounded, lms, dist, stats.total_ms * 1e-3, rec["solved"],
                                         data.get_pose_chain_names(), float(rec["objective"])))
            if rec["solved"]:
                break
            curr_iter += 1
    return iterates
