"""``VariableValues`` / ``SolverResults`` — the output containers of ``solve_score``.

Construction conventions follow the reference's call sites:
``VariableValues(dim, poses, landmarks, distances)`` positional
(/root/reference/score/utils/gurobi_utils.py:136) and
``SolverResults(variables=…, total_time=…, solved=…, pose_chain_names=…)``
by keyword (:196-203).
"""
from typing import Dict, List, Optional, Tuple

import numpy as np


class VariableValues:
    def __init__(
        self,
        dim: int,
        poses: Dict[str, np.ndarray],
        landmarks: Dict[str, np.ndarray],
        distances: Optional[Dict[Tuple[str, str], np.ndarray]] = None,
    ):
        if dim not in (2, 3):
            raise ValueError(f"Value {dim} is not 2 or 3")
        self.dim = dim
        self.poses = poses
        self.landmarks = landmarks
        self.distances = distances

    @property
    def rotations_theta(self) -> Dict[str, float]:
        assert self.dim == 2
        return {k: float(np.arctan2(v[1, 0], v[0, 0])) for k, v in self.poses.items()}

    @property
    def rotations_matrix(self) -> Dict[str, np.ndarray]:
        return {k: v[: self.dim, : self.dim] for k, v in self.poses.items()}

    @property
    def translations(self) -> Dict[str, np.ndarray]:
        out = {k: v[: self.dim, self.dim] for k, v in self.poses.items()}
        out.update(self.landmarks)
        return out


class SolverResults:
    def __init__(
        self,
        variables: VariableValues,
        total_time: float,
        solved: bool,
        pose_chain_names: Optional[List[List[str]]] = None,
        solver_cost: Optional[float] = None,
    ):
        self.variables = variables
        self.total_time = total_time
        self.solved = solved
        self.pose_chain_names = pose_chain_names
        self.solver_cost = solver_cost

    @property
    def dim(self) -> int:
        return self.variables.dim

    @property
    def poses(self):
        return self.variables.poses

    @property
    def translations(self):
        return self.variables.translations

    @property
    def rotations_theta(self):
        return self.variables.rotations_theta

    @property
    def rotations_matrix(self):
        return self.variables.rotations_matrix

    @property
    def landmarks(self):
        return self.variables.landmarks

    @property
    def distances(self):
        return self.variables.distances


def save_to_tum(solved_results: SolverResults, filepath: str, strip_extension: bool = False) -> List[str]:
    """Write one TUM trajectory file per pose chain (``idx x y z qx qy qz qw``).

    Same line layout as /root/reference/examples/goats_14_data/gt_traj_A.tum.
    """
    assert solved_results.pose_chain_names is not None
    base = filepath[: -len(".tum")] if filepath.endswith(".tum") else filepath
    written = []
    d = solved_results.dim
    for ci, chain in enumerate(solved_results.pose_chain_names):
        if not chain:
            continue
        path = f"{base}_{chain[0][0]}.tum"
        with open(path, "w") as f:
            for idx, name in enumerate(chain):
                T = solved_results.poses[name]
                t = T[:d, d]
                if d == 2:
                    th = np.arctan2(T[1, 0], T[0, 0])
                    q = (0.0, 0.0, np.sin(th / 2), np.cos(th / 2))
                    xyz = (t[0], t[1], 0.0)
                else:
                    q = _quat_from_rot(T[:3, :3])
                    xyz = (t[0], t[1], t[2])
                f.write(f"{idx} {xyz[0]} {xyz[1]} {xyz[2]} {q[0]} {q[1]} {q[2]} {q[3]}\n")
        written.append(path)
    return written


def _quat_from_rot(R: np.ndarray):
    tr = np.trace(R)
    if tr > 0:
        s = np.sqrt(tr + 1.0) * 2
        return ((R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s)
    i = int(np.argmax(np.diag(R)))
    j, k = (i + 1) % 3, (i + 2) % 3
    s = np.sqrt(1.0 + R[i, i] - R[j, j] - R[k, k]) * 2
    q = [0.0, 0.0, 0.0, 0.0]
    q[i] = 0.25 * s
    q[j] = (R[j, i] + R[i, j]) / s
    q[k] = (R[k, i] + R[i, k]) / s
    q[3] = (R[k, j] - R[j, k]) / s
    return tuple(q)
