"""Name -> index lowering of a ``FactorGraphData`` to structure-of-arrays.

This is the host half of what the reference does with dictionaries of Gurobi
variables (``VariableCollection``, /root/reference/score/utils/gurobi_utils.py:53-136)
and per-factor Python loops (:233-526): it fixes the variable order (poses
chain-major, then landmarks, then one distance variable per range, :233-310),
mirrors the reference's error behaviour, and hands plain arrays to the C ABI
(include/score_b200.h, ``ScoreProblemDesc``).  All arithmetic happens on the GPU.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Tuple

import numpy as np

SOCP_RELAXATION = "SOCP"
QCQP_RELAXATION = "QCQP"
ACCEPTABLE_RELAXATIONS = [SOCP_RELAXATION, QCQP_RELAXATION]


def check_valid_relaxation(relaxation: str) -> None:
    """gurobi_utils.py:139-144."""
    if relaxation not in ACCEPTABLE_RELAXATIONS:
        raise ValueError(
            f"Relaxation {relaxation} is not supported. "
            f"Acceptable relaxations are {ACCEPTABLE_RELAXATIONS}"
        )


def check_dimension(value) -> None:
    """is_dimension validator, gurobi_utils.py:37-50."""
    if not isinstance(value, (int, np.integer)) or isinstance(value, bool):
        raise ValueError(f"{value} is not an int")
    if value not in (2, 3):
        raise ValueError(f"Value {value} is not 2 or 3")


@dataclass
class LoweredProblem:
    """One instance (or, after ``concat``, a batch) in ScoreProblemDesc layout."""

    dim: int
    relaxation: str
    n_instances: int
    pose_off: np.ndarray
    lm_off: np.ndarray
    edge_off: np.ndarray
    rng_off: np.ndarray
    prior_off: np.ndarray
    seg_ptr: np.ndarray
    seg_inst: np.ndarray
    link_edge: np.ndarray
    edge_i: np.ndarray
    edge_j: np.ndarray
    edge_t: np.ndarray
    edge_R: np.ndarray
    edge_k: np.ndarray
    edge_tau: np.ndarray
    rng_a: np.ndarray
    rng_b: np.ndarray
    rng_dist: np.ndarray
    rng_w: np.ndarray
    prior_l: np.ndarray
    prior_t: np.ndarray
    prior_w: np.ndarray
    # names, kept for packing results (one list per instance)
    pose_names: List[List[str]] = field(default_factory=list)
    landmark_names: List[List[str]] = field(default_factory=list)
    range_keys: List[List[Tuple[str, str]]] = field(default_factory=list)

    @property
    def P(self) -> int:
        return int(self.pose_off[-1])

    @property
    def L(self) -> int:
        return int(self.lm_off[-1])

    @property
    def E(self) -> int:
        return int(self.edge_off[-1])

    @property
    def K(self) -> int:
        return int(self.rng_off[-1])

    @property
    def Lp(self) -> int:
        return int(self.prior_off[-1])

    @property
    def n_seg(self) -> int:
        return int(len(self.seg_inst))

    @property
    def dist_per(self) -> int:
        return self.dim if self.relaxation == QCQP_RELAXATION else 1


def _segments(link_edge: np.ndarray) -> np.ndarray:
    """Start index of every maximal run of poses linked by consecutive odometry."""
    starts = np.nonzero(link_edge < 0)[0]
    return np.concatenate([starts, [len(link_edge)]]).astype(np.int32)


def lower_factor_graph(fg, relaxation: str = QCQP_RELAXATION) -> LoweredProblem:
    """Lower one ``FactorGraphData``.

    Error behaviour mirrors the reference: ``ValueError`` for a bad relaxation
    (gurobi_utils.py:139-144), bad dimension (:37-50), duplicate variable names
    (:62-80) and unknown translation owners (:103-109); ``KeyError`` for a pose
    measurement that names an unknown pose (:99-100); ``IndexError`` when there
    is no first pose to pin (:181).
    """
    check_valid_relaxation(relaxation)
    d = fg.dimension
    check_dimension(d)
    d = int(d)

    pose_idx: Dict[str, int] = {}
    pose_names: List[str] = []
    chain_of: List[int] = []
    for ci, chain in enumerate(fg.pose_variables):
        for pose in chain:
            name = pose.name
            if not isinstance(name, str):
                raise ValueError(f"Variable name {name} is not a valid type: {type(name)}")
            if name in pose_idx:
                raise ValueError(f"Variable name {name} already exists in pose_vars")
            pose_idx[name] = len(pose_names)
            pose_names.append(name)
            chain_of.append(ci)
    P = len(pose_names)
    lm_idx: Dict[str, int] = {}
    landmark_names: List[str] = []
    for lm in fg.landmark_variables:
        name = lm.name
        if name in pose_idx:
            raise ValueError(f"Variable name {name} already exists in pose_vars")
        if name in lm_idx:
            raise ValueError(f"Variable name {name} already exists in landmark_vars")
        lm_idx[name] = len(landmark_names)
        landmark_names.append(name)
    L = len(landmark_names)

    # the pinned pose: fg.pose_variables[0][0] (gurobi_utils.py:181) -> IndexError if absent
    first_pose = fg.pose_variables[0][0]
    assert pose_idx[first_pose.name] == 0

    range_keys: List[Tuple[str, str]] = []
    seen = set()
    for meas in fg.range_measurements:
        key = (meas.first_key, meas.second_key)
        if relaxation == QCQP_RELAXATION and key in seen:
            # add_distance_variable -> _check_is_new_variable (gurobi_utils.py:62-67,296-306)
            raise ValueError(f"Variable name {key} already exists in distance_vars")
        seen.add(key)
        range_keys.append(key)
    K = len(range_keys)

    def owner(name: str) -> int:
        if name in pose_idx:
            return pose_idx[name]
        if name in lm_idx:
            return P + lm_idx[name]
        raise ValueError(f"Variable name {name} not found")

    edges = [m for chain in fg.odom_measurements for m in chain]
    n_odom = len(edges)
    edges += list(fg.loop_closure_measurements)
    E = len(edges)
    edge_i = np.zeros(E, np.int32)
    edge_j = np.zeros(E, np.int32)
    edge_t = np.zeros((E, d))
    edge_R = np.zeros((E, d, d))
    edge_k = np.zeros(E)
    edge_tau = np.zeros(E)
    for e, m in enumerate(edges):
        edge_i[e] = pose_idx[m.base_pose]  # KeyError like VariableCollection.get_pose_var
        edge_j[e] = pose_idx[m.to_pose]
        edge_t[e] = np.asarray(m.translation_vector, dtype=np.float64).reshape(d)
        edge_R[e] = np.asarray(m.rotation_matrix, dtype=np.float64).reshape(d, d)
        edge_k[e] = float(m.translation_precision)
        edge_tau[e] = float(m.rotation_precision)
    if E and np.any(edge_i == edge_j):
        raise ValueError("relative-pose measurement connects a pose to itself")

    # odometry links (p-1 -> p) inside a chain define the preconditioner's chain segments
    link_edge = -np.ones(P, np.int32)
    chain_arr = np.asarray(chain_of, np.int64)
    for e in range(n_odom):
        i, j = int(edge_i[e]), int(edge_j[e])
        if j == i + 1 and chain_arr[i] == chain_arr[j] and link_edge[j] < 0:
            link_edge[j] = e
    seg_ptr = _segments(link_edge)

    rng_a = np.zeros(K, np.int32)
    rng_b = np.zeros(K, np.int32)
    rng_dist = np.zeros(K)
    rng_w = np.zeros(K)
    for k, meas in enumerate(fg.range_measurements):
        rng_a[k] = owner(meas.first_key)
        rng_b[k] = owner(meas.second_key)
        rng_dist[k] = float(meas.dist)
        rng_w[k] = float(meas.precision)
    if K and np.any(rng_a == rng_b):
        raise ValueError("range measurement connects a variable to itself")

    priors = list(fg.landmark_priors)
    Lp = len(priors)
    prior_l = np.zeros(Lp, np.int32)
    prior_t = np.zeros((Lp, d))
    prior_w = np.zeros(Lp)
    for q, pr in enumerate(priors):
        if pr.name not in lm_idx and pr.name not in pose_idx:
            raise ValueError(f"Variable name {pr.name} not found")
        if pr.name not in lm_idx:
            raise ValueError(f"landmark prior on non-landmark variable {pr.name}")
        prior_l[q] = lm_idx[pr.name]
        prior_t[q] = np.asarray(pr.translation_vector, dtype=np.float64).reshape(d)
        prior_w[q] = float(pr.translation_precision)

    i32 = lambda *v: np.asarray(v, np.int32)
    return LoweredProblem(
        dim=d,
        relaxation=relaxation,
        n_instances=1,
        pose_off=i32(0, P),
        lm_off=i32(0, L),
        edge_off=i32(0, E),
        rng_off=i32(0, K),
        prior_off=i32(0, Lp),
        seg_ptr=seg_ptr,
        seg_inst=np.zeros(len(seg_ptr) - 1, np.int32),
        link_edge=link_edge,
        edge_i=edge_i,
        edge_j=edge_j,
        edge_t=edge_t,
        edge_R=edge_R,
        edge_k=edge_k,
        edge_tau=edge_tau,
        rng_a=rng_a,
        rng_b=rng_b,
        rng_dist=rng_dist,
        rng_w=rng_w,
        prior_l=prior_l,
        prior_t=prior_t,
        prior_w=prior_w,
        pose_names=[pose_names],
        landmark_names=[landmark_names],
        range_keys=[range_keys],
    )


def lower_manhattan_arrays(arr: dict, relaxation: str = QCQP_RELAXATION, with_names: bool = True) -> LoweredProblem:
    """Lower the array form produced by ``generators.manhattan_2d_arrays`` without
    materialising Python factor objects (identical result to lowering
    ``arrays_to_factor_graph(arr)``; used for the 1024-instance sweep)."""
    check_valid_relaxation(relaxation)
    d = 2
    R, S = arr["n_robots"], arr["n_steps"]
    P, L = R * S, len(arr["landmarks"])
    E = R * (S - 1)
    base = (np.arange(R)[:, None] * S + np.arange(S - 1)[None, :]).ravel().astype(np.int32)
    th = arr["odom_theta"].ravel()
    c, s = np.cos(th), np.sin(th)
    edge_R = np.stack([np.stack([c, -s], -1), np.stack([s, c], -1)], -2)
    link_edge = -np.ones(P, np.int32)
    link_edge[base + 1] = np.arange(E, dtype=np.int32)
    K = len(arr["rng_a"])
    i32 = lambda *v: np.asarray(v, np.int32)
    from .generators import _chain_prefix

    if with_names:
        pose_names = [f"{_chain_prefix(r)}{t}" for r in range(R) for t in range(S)]
        lm_names = [f"L{q}" for q in range(L)]
        all_names = pose_names + lm_names
        range_keys = [(all_names[a], all_names[b]) for a, b in zip(arr["rng_a"], arr["rng_b"])]
    else:  # throughput runs: results are consumed as arrays, no dict packing
        pose_names, lm_names, range_keys = [], [], []
    return LoweredProblem(
        dim=d,
        relaxation=relaxation,
        n_instances=1,
        pose_off=i32(0, P),
        lm_off=i32(0, L),
        edge_off=i32(0, E),
        rng_off=i32(0, K),
        prior_off=i32(0, 0),
        seg_ptr=(np.arange(R + 1) * S).astype(np.int32),
        seg_inst=np.zeros(R, np.int32),
        link_edge=link_edge,
        edge_i=base,
        edge_j=base + 1,
        edge_t=np.stack([arr["odom_x"].ravel(), arr["odom_y"].ravel()], -1),
        edge_R=edge_R,
        edge_k=np.full(E, arr["k_t"]),
        edge_tau=np.full(E, arr["k_r"]),
        rng_a=arr["rng_a"].astype(np.int32),
        rng_b=arr["rng_b"].astype(np.int32),
        rng_dist=np.asarray(arr["rng_dist"], np.float64),
        rng_w=np.full(K, 1.0 / arr["sigma_range"] ** 2),
        prior_l=np.zeros(0, np.int32),
        prior_t=np.zeros((0, d)),
        prior_w=np.zeros(0),
        pose_names=[pose_names],
        landmark_names=[lm_names],
        range_keys=[range_keys],
    )


def lower_grid3d_arrays(arr: dict, relaxation: str = QCQP_RELAXATION, with_names: bool = False) -> LoweredProblem:
    """Lower the array form produced by ``generators.grid_3d_arrays`` (config 5: 100k poses, 1M ranges)
    without materialising Python factor objects."""
    check_valid_relaxation(relaxation)
    d = 3
    R, S = arr["n_robots"], arr["n_steps"]
    P, L = R * S, len(arr["landmarks"])
    E = R * (S - 1)
    base = (np.arange(R)[:, None] * S + np.arange(S - 1)[None, :]).ravel().astype(np.int32)
    link_edge = -np.ones(P, np.int32)
    link_edge[base + 1] = np.arange(E, dtype=np.int32)
    K = len(arr["rng_a"])
    i32 = lambda *v: np.asarray(v, np.int32)
    from .generators import _chain_prefix

    if with_names:
        pose_names = [f"{_chain_prefix(r)}{t}" for r in range(R) for t in range(S)]
        lm_names = [f"L{q}" for q in range(L)]
        all_names = pose_names + lm_names
        range_keys = [(all_names[a], all_names[b]) for a, b in zip(arr["rng_a"], arr["rng_b"])]
    else:
        pose_names, lm_names, range_keys = [], [], []
    return LoweredProblem(
        dim=d, relaxation=relaxation, n_instances=1,
        pose_off=i32(0, P), lm_off=i32(0, L), edge_off=i32(0, E), rng_off=i32(0, K), prior_off=i32(0, 0),
        seg_ptr=(np.arange(R + 1) * S).astype(np.int32), seg_inst=np.zeros(R, np.int32), link_edge=link_edge,
        edge_i=base, edge_j=base + 1,
        edge_t=np.ascontiguousarray(arr["odom_t"].reshape(E, 3)), edge_R=np.ascontiguousarray(arr["odom_R"].reshape(E, 3, 3)),
        edge_k=np.full(E, arr["k_t"]), edge_tau=np.full(E, arr["k_r"]),
        rng_a=arr["rng_a"].astype(np.int32), rng_b=arr["rng_b"].astype(np.int32),
        rng_dist=np.asarray(arr["rng_dist"], np.float64), rng_w=np.full(K, 1.0 / arr["sigma_range"] ** 2),
        prior_l=np.zeros(0, np.int32), prior_t=np.zeros((0, d)), prior_w=np.zeros(0),
        pose_names=[pose_names], landmark_names=[lm_names], range_keys=[range_keys],
    )


def concat(problems: Sequence[LoweredProblem]) -> LoweredProblem:
    """Concatenate independent instances into one batch (block-diagonal problem)."""
    if not problems:
        raise ValueError("empty batch")
    d, relax = problems[0].dim, problems[0].relaxation
    for p in problems:
        if p.dim != d or p.relaxation != relax:
            raise ValueError("all instances of a batch must share dimension and relaxation")

    def cat_off(attr):
        out = [np.zeros(1, np.int64)]
        base = 0
        for p in problems:
            o = np.asarray(getattr(p, attr), np.int64)
            out.append(o[1:] + base)
            base += int(o[-1])
        res = np.concatenate(out)
        if res[-1] >= 2**31:
            raise ValueError("batch too large for 32-bit indexing")
        return res.astype(np.int32)

    pose_off = cat_off("pose_off")
    edge_off = cat_off("edge_off")
    # segment starts are global pose indices; link_edge holds global edge ids
    seg_ptr_parts, seg_inst_parts, link_parts = [], [], []
    pbase = ebase = ibase = 0
    for p in problems:
        seg_ptr_parts.append(np.asarray(p.seg_ptr[:-1], np.int64) + pbase)
        seg_inst_parts.append(np.asarray(p.seg_inst, np.int64) + ibase)
        le = np.asarray(p.link_edge, np.int64)
        link_parts.append(np.where(le >= 0, le + ebase, -1))
        pbase += p.P
        ebase += p.E
        ibase += p.n_instances
    seg_ptr = np.concatenate(seg_ptr_parts + [np.asarray([pbase])]).astype(np.int32)
    cat = lambda attr: np.concatenate([getattr(p, attr) for p in problems], axis=0)
    return LoweredProblem(
        dim=d,
        relaxation=relax,
        n_instances=ibase,
        pose_off=pose_off,
        lm_off=cat_off("lm_off"),
        edge_off=edge_off,
        rng_off=cat_off("rng_off"),
        prior_off=cat_off("prior_off"),
        seg_ptr=seg_ptr,
        seg_inst=np.concatenate(seg_inst_parts).astype(np.int32),
        link_edge=np.concatenate(link_parts).astype(np.int32),
        edge_i=cat("edge_i"),
        edge_j=cat("edge_j"),
        edge_t=cat("edge_t"),
        edge_R=cat("edge_R"),
        edge_k=cat("edge_k"),
        edge_tau=cat("edge_tau"),
        rng_a=cat("rng_a"),
        rng_b=cat("rng_b"),
        rng_dist=cat("rng_dist"),
        rng_w=cat("rng_w"),
        prior_l=cat("prior_l"),
        prior_t=cat("prior_t"),
        prior_w=cat("prior_w"),
        pose_names=[n for p in problems for n in p.pose_names],
        landmark_names=[n for p in problems for n in p.landmark_names],
        range_keys=[n for p in problems for n in p.range_keys],
    )


def slice_instances(prob: LoweredProblem, i0: int, i1: int) -> LoweredProblem:
    """Instances [i0, i1) of a batch as a batch of their own (inverse of ``concat``)."""
    if not (0 <= i0 < i1 <= prob.n_instances):
        raise ValueError("bad instance range")
    d = prob.dim
    p0, p1 = int(prob.pose_off[i0]), int(prob.pose_off[i1])
    e0, e1 = int(prob.edge_off[i0]), int(prob.edge_off[i1])
    k0, k1 = int(prob.rng_off[i0]), int(prob.rng_off[i1])
    q0, q1 = int(prob.prior_off[i0]), int(prob.prior_off[i1])
    seg_inst = np.asarray(prob.seg_inst)
    s0, s1 = int(np.searchsorted(seg_inst, i0, side="left")), int(np.searchsorted(seg_inst, i1, side="left"))
    le = np.asarray(prob.link_edge[p0:p1], np.int64)
    off = lambda a: (np.asarray(a[i0 : i1 + 1], np.int64) - int(a[i0])).astype(np.int32)
    names = lambda lst: lst[i0:i1] if len(lst) == prob.n_instances else []
    return LoweredProblem(
        dim=d, relaxation=prob.relaxation, n_instances=i1 - i0,
        pose_off=off(prob.pose_off), lm_off=off(prob.lm_off), edge_off=off(prob.edge_off), rng_off=off(prob.rng_off),
        prior_off=off(prob.prior_off),
        seg_ptr=(np.asarray(prob.seg_ptr[s0 : s1 + 1], np.int64) - p0).astype(np.int32),
        seg_inst=(seg_inst[s0:s1] - i0).astype(np.int32),
        link_edge=np.where(le >= 0, le - e0, -1).astype(np.int32),
        edge_i=prob.edge_i[e0:e1], edge_j=prob.edge_j[e0:e1], edge_t=prob.edge_t[e0:e1], edge_R=prob.edge_R[e0:e1],
        edge_k=prob.edge_k[e0:e1], edge_tau=prob.edge_tau[e0:e1],
        rng_a=prob.rng_a[k0:k1], rng_b=prob.rng_b[k0:k1], rng_dist=prob.rng_dist[k0:k1], rng_w=prob.rng_w[k0:k1],
        prior_l=prob.prior_l[q0:q1], prior_t=prob.prior_t[q0:q1], prior_w=prob.prior_w[q0:q1],
        pose_names=names(prob.pose_names), landmark_names=names(prob.landmark_names), range_keys=names(prob.range_keys),
    )
