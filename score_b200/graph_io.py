"""Portable array (npz) form of a ``FactorGraphData``.

The reference only reads pickles of PyFactorGraph objects
(/root/reference/examples/solve_goats_example_score.py:18,40).  The npz form keeps
exactly the fields the hot path reads, in creation order, so a graph saved here
lowers to bit-identical arrays.  It is also what tests/golden/ stores.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np

from . import pyfg_shim as _shim

_shim.install()

from py_factor_graph.factor_graph import FactorGraphData  # noqa: E402
from py_factor_graph.measurements import FGRangeMeasurement, PoseMeasurement2D, PoseMeasurement3D  # noqa: E402
from py_factor_graph.priors import LandmarkPrior2D, LandmarkPrior3D  # noqa: E402
from py_factor_graph.variables import (  # noqa: E402
    LandmarkVariable2D,
    LandmarkVariable3D,
    PoseVariable2D,
    PoseVariable3D,
)


def graph_to_arrays(fg) -> Dict[str, np.ndarray]:
    d = int(fg.dimension)
    out: Dict[str, np.ndarray] = {"dim": np.asarray(d)}
    out["chain_len"] = np.asarray([len(c) for c in fg.pose_variables], np.int64)
    poses = [p for c in fg.pose_variables for p in c]
    out["pose_names"] = np.asarray([p.name for p in poses])
    out["pose_true_pos"] = np.asarray([list(p.true_position) for p in poses], np.float64).reshape(len(poses), d)
    if d == 2:
        out["pose_true_theta"] = np.asarray([p.true_theta for p in poses], np.float64)
    else:
        out["pose_true_rot"] = np.asarray([p.true_rotation for p in poses], np.float64).reshape(len(poses), 3, 3)
    out["lm_names"] = np.asarray([l.name for l in fg.landmark_variables])
    out["lm_true_pos"] = np.asarray([list(l.true_position) for l in fg.landmark_variables], np.float64).reshape(
        len(fg.landmark_variables), d
    )

    def pack_rel(prefix, meas_list):
        out[prefix + "_base"] = np.asarray([m.base_pose for m in meas_list])
        out[prefix + "_to"] = np.asarray([m.to_pose for m in meas_list])
        if d == 2:
            out[prefix + "_xyt"] = np.asarray([[m.x, m.y, m.theta] for m in meas_list], np.float64).reshape(-1, 3)
        else:
            out[prefix + "_t"] = np.asarray([m.translation for m in meas_list], np.float64).reshape(-1, 3)
            out[prefix + "_R"] = np.asarray([m.rotation for m in meas_list], np.float64).reshape(-1, 3, 3)
        out[prefix + "_prec"] = np.asarray(
            [[m.translation_precision, m.rotation_precision] for m in meas_list], np.float64
        ).reshape(-1, 2)

    out["odom_len"] = np.asarray([len(c) for c in fg.odom_measurements], np.int64)
    pack_rel("odom", [m for c in fg.odom_measurements for m in c])
    pack_rel("loop", list(fg.loop_closure_measurements))
    out["rng_first"] = np.asarray([m.association[0] for m in fg.range_measurements])
    out["rng_second"] = np.asarray([m.association[1] for m in fg.range_measurements])
    out["rng_dist_std"] = np.asarray([[m.dist, m.stddev] for m in fg.range_measurements], np.float64).reshape(-1, 2)
    out["prior_names"] = np.asarray([p.name for p in fg.landmark_priors])
    out["prior_pos"] = np.asarray([list(p.position) for p in fg.landmark_priors], np.float64).reshape(-1, d)
    out["prior_prec"] = np.asarray([p.translation_precision for p in fg.landmark_priors], np.float64)
    return out


def arrays_to_graph(arr) -> FactorGraphData:
    d = int(arr["dim"])
    fg = FactorGraphData(d)
    names = [str(n) for n in arr["pose_names"]]
    idx = 0
    for ci, n in enumerate(arr["chain_len"]):
        while len(fg.pose_variables) <= ci:
            fg.pose_variables.append([])
        for _ in range(int(n)):
            if d == 2:
                pv = PoseVariable2D(names[idx], tuple(arr["pose_true_pos"][idx]), float(arr["pose_true_theta"][idx]))
            else:
                pv = PoseVariable3D(names[idx], tuple(arr["pose_true_pos"][idx]), arr["pose_true_rot"][idx])
            fg.add_pose_variable(pv, chain=ci)
            idx += 1
    for i, n in enumerate(arr["lm_names"]):
        cls = LandmarkVariable2D if d == 2 else LandmarkVariable3D
        fg.add_landmark_variable(cls(str(n), tuple(arr["lm_true_pos"][i])))

    def unpack_rel(prefix, i):
        kt, kr = arr[prefix + "_prec"][i]
        b, t = str(arr[prefix + "_base"][i]), str(arr[prefix + "_to"][i])
        if d == 2:
            x, y, th = arr[prefix + "_xyt"][i]
            return PoseMeasurement2D(b, t, float(x), float(y), float(th), float(kt), float(kr))
        return PoseMeasurement3D(b, t, arr[prefix + "_t"][i], arr[prefix + "_R"][i], float(kt), float(kr))

    i = 0
    for ci, n in enumerate(arr["odom_len"]):
        while len(fg.odom_measurements) <= ci:
            fg.odom_measurements.append([])
        for _ in range(int(n)):
            fg.add_odom_measurement(ci, unpack_rel("odom", i))
            i += 1
    for i in range(len(arr["loop_base"])):
        fg.add_loop_closure(unpack_rel("loop", i))
    for i in range(len(arr["rng_first"])):
        dist, std = arr["rng_dist_std"][i]
        fg.add_range_measurement(
            FGRangeMeasurement((str(arr["rng_first"][i]), str(arr["rng_second"][i])), float(dist), float(std))
        )
    for i in range(len(arr["prior_names"])):
        cls = LandmarkPrior2D if d == 2 else LandmarkPrior3D
        fg.add_landmark_prior(cls(str(arr["prior_names"][i]), tuple(arr["prior_pos"][i]), float(arr["prior_prec"][i])))
    return fg


def save_graph_npz(fg, path: str, **extra) -> None:
    arrs = graph_to_arrays(fg)
    for k, v in extra.items():
        arrs["extra_" + k] = np.asarray(v)
    np.savez_compressed(path, **arrs)


def load_graph_npz(path: str) -> Tuple[FactorGraphData, Dict[str, np.ndarray]]:
    with np.load(path, allow_pickle=False) as z:
        arr = {k: z[k] for k in z.files}
    extra = {k[len("extra_"):]: v for k, v in arr.items() if k.startswith("extra_")}
    return arrays_to_graph(arr), extra


def robot_subgraph(fg, chain: int = 0) -> FactorGraphData:
    """Single-robot sub-graph: one chain, every landmark that keeps a
    range to it, the robot->landmark ranges; inter-robot ranges and other chains are dropped."""
    d = int(fg.dimension)
    out = FactorGraphData(d)
    keep = {p.name for p in fg.pose_variables[chain]}
    for p in fg.pose_variables[chain]:
        out.add_pose_variable(p, chain=0)
    lm_names = {l.name for l in fg.landmark_variables}
    used = set()
    rngs = []
    for m in fg.range_measurements:
        a, b = m.association
        if (a in keep and b in lm_names) or (b in keep and a in lm_names):
            rngs.append(m)
            used.add(b if b in lm_names else a)
    for l in fg.landmark_variables:
        if l.name in used:
            out.add_landmark_variable(l)
    for m in fg.odom_measurements[chain]:
        out.add_odom_measurement(0, m)
    for m in fg.loop_closure_measurements:
        if m.base_pose in keep and m.to_pose in keep:
            out.add_loop_closure(m)
    for m in rngs:
        out.add_range_measurement(m)
    for pr in fg.landmark_priors:
        if pr.name in used:
            out.add_landmark_prior(pr)
    return out
