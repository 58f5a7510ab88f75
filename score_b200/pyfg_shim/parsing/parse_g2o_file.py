"""g2o text files <-> ``FactorGraphData``.

The reference reaches other SLAM file types through PyFactorGraph ("a custom library ... to interface with a
broader range of SLAM file types (e.g. g2o)", /root/reference/README.md:53-56); PyFactorGraph itself is not under
/root/reference, so this is a reader / writer for the public g2o vocabulary, not a restatement of its parser:

    VERTEX_SE2 id x y theta                     EDGE_SE2 i j dx dy dtheta  I11 I12 I13 I22 I23 I33
    VERTEX_SE3:QUAT id x y z qx qy qz qw        EDGE_SE3:QUAT i j x y z qx qy qz qw  <21 upper-triangular information entries>
    VERTEX_XY id x y   /  VERTEX_TRACKXYZ id x y z          (landmarks)
    EDGE_RANGE a b dist information             (extension written / read by this module: range-aided SLAM has no
                                                 standard g2o tag; information = 1 / stddev^2)

Ids are plain integers (one robot chain "A", landmarks "L") or GTSAM symbol keys (character in the top byte, index
below: one chain per character, 'L' = landmarks).  Relative-pose information matrices are reduced to the two isotropic
precisions the SCORE cost uses (score/utils/gurobi_utils.py:504-526): the mean of the translation block's diagonal and
of the rotation block's diagonal.  An edge between consecutive poses of a chain is odometry, any other a loop closure.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np

from ..factor_graph import FactorGraphData
from ..measurements import FGRangeMeasurement, PoseMeasurement2D, PoseMeasurement3D
from ..variables import LandmarkVariable2D, LandmarkVariable3D, PoseVariable2D, PoseVariable3D

_SYMBOL_SHIFT = 56


def _quat_to_rot(q) -> np.ndarray:
    x, y, z, w = (float(v) for v in q)
    n = np.sqrt(x * x + y * y + z * z + w * w)
    x, y, z, w = x / n, y / n, z / n, w / n
    return np.array(
        [
            [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
            [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
            [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
        ]
    )


def _rot_to_quat(R) -> Tuple[float, float, float, float]:
    R = np.asarray(R, float)
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        w, x, y, z = 0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(1.0 + R[i, i] - R[j, j] - R[k, k]) * 2
        q = [0.0, 0.0, 0.0]
        q[i] = 0.25 * s
        q[j] = (R[j, i] + R[i, j]) / s
        q[k] = (R[k, i] + R[i, k]) / s
        w = (R[k, j] - R[j, k]) / s
        x, y, z = q
    return float(x), float(y), float(z), float(w)


def _split_id(raw: int) -> Tuple[str, int]:
    if raw >> _SYMBOL_SHIFT:
        return chr(raw >> _SYMBOL_SHIFT), raw & ((1 << _SYMBOL_SHIFT) - 1)
    return "", raw


def parse_g2o_file(filepath: str) -> FactorGraphData:
    poses: Dict[int, tuple] = {}
    lms: Dict[int, tuple] = {}
    edges: List[tuple] = []
    ranges: List[tuple] = []
    dim = None
    with open(filepath) as f:
        for line in f:
            tok = line.split()
            if not tok or tok[0].startswith("#"):
                continue
            tag = tok[0]
            if tag == "VERTEX_SE2":
                dim = dim or 2
                poses[int(tok[1])] = tuple(float(v) for v in tok[2:5])
            elif tag == "VERTEX_SE3:QUAT":
                dim = dim or 3
                poses[int(tok[1])] = tuple(float(v) for v in tok[2:9])
            elif tag in ("VERTEX_XY", "VERTEX_TRACKXYZ"):
                lms[int(tok[1])] = tuple(float(v) for v in tok[2:])
            elif tag == "EDGE_SE2":
                v = [float(x) for x in tok[3:]]
                edges.append((int(tok[1]), int(tok[2]), v[:3], 0.5 * (v[3] + v[6]), v[8]))
            elif tag == "EDGE_SE3:QUAT":
                v = [float(x) for x in tok[3:]]
                info = np.zeros((6, 6))
                info[np.triu_indices(6)] = v[7:28]
                d = np.diag(info)
                edges.append((int(tok[1]), int(tok[2]), v[:7], float(d[:3].mean()), float(d[3:].mean())))
            elif tag == "EDGE_RANGE":
                ranges.append((int(tok[1]), int(tok[2]), float(tok[3]), float(tok[4])))
            elif tag in ("FIX",):
                continue
            else:
                raise ValueError(f"{filepath}: unsupported g2o record {tag!r}")
    if dim is None:
        raise ValueError(f"{filepath}: no pose vertices")
    fg = FactorGraphData(dim)
    # chains: one per symbol character (plain ids: one chain "A"), poses in id order
    name_of: Dict[int, str] = {}
    chains: Dict[str, List[int]] = {}
    for raw in sorted(poses):
        ch, _ = _split_id(raw)
        chains.setdefault(ch or "A", []).append(raw)
    pos_in_chain: Dict[int, Tuple[int, int]] = {}
    for ci, ch in enumerate(sorted(chains)):
        for k, raw in enumerate(chains[ch]):
            _, idx = _split_id(raw)
            name_of[raw] = f"{ch}{idx}"
            pos_in_chain[raw] = (ci, k)
            v = poses[raw]
            if dim == 2:
                fg.add_pose_variable(PoseVariable2D(name_of[raw], (v[0], v[1]), v[2]), chain=ci)
            else:
                fg.add_pose_variable(PoseVariable3D(name_of[raw], v[:3], _quat_to_rot(v[3:7])), chain=ci)
    for raw in sorted(lms):
        _, idx = _split_id(raw)
        name_of[raw] = f"L{idx}"
        cls = LandmarkVariable2D if dim == 2 else LandmarkVariable3D
        fg.add_landmark_variable(cls(name_of[raw], tuple(lms[raw][:dim])))
    for i, j, v, kt, kr in edges:
        if dim == 2:
            m = PoseMeasurement2D(name_of[i], name_of[j], v[0], v[1], v[2], kt, kr)
        else:
            m = PoseMeasurement3D(name_of[i], name_of[j], np.asarray(v[:3]), _quat_to_rot(v[3:7]), kt, kr)
        ci, ki = pos_in_chain[i]
        cj, kj = pos_in_chain[j]
        if ci == cj and kj == ki + 1:
            fg.add_odom_measurement(ci, m)
        else:
            fg.add_loop_closure(m)
    for a, b, dist, info in ranges:
        fg.add_range_measurement(FGRangeMeasurement((name_of[a], name_of[b]), dist, 1.0 / np.sqrt(info)))
    return fg


def write_g2o_file(fg: FactorGraphData, filepath: str) -> None:
    """Write the measurements the SCORE cost reads (variables with their ground truth as initial values, relative-pose
    factors with isotropic information, ranges as EDGE_RANGE) with GTSAM symbol ids."""
    dim = int(fg.dimension)

    def key(name: str) -> int:
        return (ord(name[0]) << _SYMBOL_SHIFT) | int(name[1:])

    g = lambda x: format(float(x), ".17g")
    with open(filepath, "w") as f:
        for chain in fg.pose_variables:
            for p in chain:
                if dim == 2:
                    f.write(f"VERTEX_SE2 {key(p.name)} {g(p.true_position[0])} {g(p.true_position[1])} {g(p.true_theta)}\n")
                else:
                    q = _rot_to_quat(p.true_rotation)
                    f.write(f"VERTEX_SE3:QUAT {key(p.name)} " + " ".join(g(v) for v in (*p.true_position, *q)) + "\n")
        for l in fg.landmark_variables:
            tag = "VERTEX_XY" if dim == 2 else "VERTEX_TRACKXYZ"
            f.write(f"{tag} {key(l.name)} " + " ".join(g(v) for v in l.true_position) + "\n")
        rel = [m for c in fg.odom_measurements for m in c] + list(fg.loop_closure_measurements)
        for m in rel:
            kt, kr = m.translation_precision, m.rotation_precision
            if dim == 2:
                info = [kt, 0, 0, kt, 0, kr]
                f.write(f"EDGE_SE2 {key(m.base_pose)} {key(m.to_pose)} {g(m.x)} {g(m.y)} {g(m.theta)} " +
                        " ".join(g(v) for v in info) + "\n")
            else:
                info = np.diag([kt] * 3 + [kr] * 3)[np.triu_indices(6)]
                q = _rot_to_quat(m.rotation)
                f.write(f"EDGE_SE3:QUAT {key(m.base_pose)} {key(m.to_pose)} " +
                        " ".join(g(v) for v in (*m.translation, *q, *info)) + "\n")
        for m in fg.range_measurements:
            f.write(f"EDGE_RANGE {key(m.first_key)} {key(m.second_key)} {g(m.dist)} {g(1.0 / m.stddev ** 2)}\n")
