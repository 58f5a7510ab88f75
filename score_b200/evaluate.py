"""Evaluation step after ``solve_score``: SE(d)-aligned absolute trajectory error against
the ground truth the reference ships beside its inputs (``PoseVariable.true_position`` in
/root/reference/examples/manhattan/factor_graph.pickle, /root/reference/examples/goats_14_data/gt_traj_A.tum).

The reductions and the Kabsch rotation run on the GPU (``score_trajectory_ate`` / ``score_eval_ate``,
csrc/evaluate.cuh); this module only gathers names into arrays.  No CPU fallback.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np

from .solver import trajectory_ate


def ground_truth_positions(data) -> np.ndarray:
    """[P, d] true positions in ``data.pose_variables`` (chain-major) order."""
    d = int(data.dimension)
    return np.asarray([list(p.true_position) for chain in data.pose_variables for p in chain], np.float64).reshape(-1, d)


def chain_offsets(data) -> np.ndarray:
    """[n_chains + 1] offsets of the robot chains in chain-major pose order."""
    return np.concatenate([[0], np.cumsum([len(c) for c in data.pose_variables])]).astype(np.int32)


def estimated_positions(results, data) -> np.ndarray:
    """[P, d] estimated translations of a ``SolverResults`` in ``data.pose_variables`` order."""
    d = int(data.dimension)
    poses = results.variables.poses
    return np.asarray([poses[p.name][:d, d] for chain in data.pose_variables for p in chain], np.float64).reshape(-1, d)


def evaluate_ate(results, data, per_chain: bool = False, align: bool = True, device: int = 0) -> Dict[str, object]:
    """ATE of one solve: ``{"rmse": float | [per chain], "R": ..., "t": ...}`` with gt ~ R est + t."""
    est, gt = estimated_positions(results, data), ground_truth_positions(data)
    off = chain_offsets(data) if per_chain else None
    rmse, R, t = trajectory_ate(est, gt, off, align=align, device=device)
    if per_chain:
        return {"rmse": rmse, "R": R, "t": t}
    return {"rmse": float(rmse[0]), "R": R[0], "t": t[0]}


def evaluate_ate_batch(results_list: Sequence, datas: Sequence, align: bool = True, device: int = 0) -> np.ndarray:
    """One aligned RMSE per (results, data) pair, all trajectories in one launch."""
    ests = [estimated_positions(r, d) for r, d in zip(results_list, datas)]
    gts = [ground_truth_positions(d) for d in datas]
    if not ests:
        return np.empty(0)
    off = np.concatenate([[0], np.cumsum([len(e) for e in ests])]).astype(np.int32)
    rmse, _, _ = trajectory_ate(np.concatenate(ests), np.concatenate(gts), off, align=align, device=device)
    return rmse
