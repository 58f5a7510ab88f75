"""Synthetic range-aided SLAM graphs for the sweep / large-graph workloads.

The reference ships no generator; parameters are inferred from its Manhattan
pickle: unit steps on an integer
grid, 90-degree turns with p=0.2, odometry noise sigma 0.01 m / 0.002 rad
(precisions 1e4 / 2.5e5), range noise sigma 1 m, range emission rate 0.1 per
(pose, landmark) and same-timestep (pose, pose) pair.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np

from . import pyfg_shim as _shim

_shim.install()

from py_factor_graph.factor_graph import FactorGraphData  # noqa: E402
from py_factor_graph.measurements import (  # noqa: E402
    FGRangeMeasurement,
    PoseMeasurement2D,
    PoseMeasurement3D,
)
from py_factor_graph.variables import (  # noqa: E402
    LandmarkVariable2D,
    LandmarkVariable3D,
    PoseVariable2D,
    PoseVariable3D,
)

MC_BASE_SEED = 20221003


def _chain_prefix(idx: int) -> str:
    chars = "ABCDEFGHIJKMNOPQRSTUVWXYZ"  # "L" is reserved for landmarks
    if idx < len(chars):
        return chars[idx]
    return chars[(idx // len(chars) - 1) % len(chars)] + chars[idx % len(chars)].lower() + "_"


def _rot2(th: float) -> np.ndarray:
    c, s = np.cos(th), np.sin(th)
    return np.array([[c, -s], [s, c]])


def manhattan_2d_arrays(
    seed: int,
    n_robots: int = 20,
    n_steps: int = 100,
    grid: int = 20,
    n_landmarks: int = 6,
    p_turn: float = 0.2,
    p_range: float = 0.1,
    sigma_xy: float = 0.01,
    sigma_th: float = 0.002,
    sigma_range: float = 1.0,
) -> dict:
    """Array form of one Manhattan-world instance.

    Returns poses (ground truth), odometry (x, y, theta per step), and ranges as
    (a, b, dist) with global translation-owner ids (pose p -> p chain-major,
    landmark q -> P + q), sorted by timestep then keys.  Every robot and every
    landmark is guaranteed at least one range (regenerated otherwise) so that
    ``unconnected_variable_names`` is empty.
    """
    rng = np.random.default_rng(seed)
    R, S, Lm = n_robots, n_steps, n_landmarks
    P = R * S
    while True:
        pos = np.zeros((R, S, 2))
        head = np.zeros((R, S), np.int64)
        start = rng.integers(0, grid + 1, size=(R, 2))
        h0 = rng.integers(0, 4, size=R)
        turn_u = rng.random((R, S))
        side_u = rng.random((R, S))
        dxy = np.array([[1, 0], [0, 1], [-1, 0], [0, -1]])
        for r in range(R):
            p = start[r].copy()
            h = int(h0[r])
            for s in range(S):
                pos[r, s] = p
                head[r, s] = h
                if s == S - 1:
                    break
                hh = h
                if turn_u[r, s] < p_turn:
                    hh = (h + (1 if side_u[r, s] < 0.5 else 3)) % 4
                q = p + dxy[hh]
                if not (0 <= q[0] <= grid and 0 <= q[1] <= grid):
                    # forced turn at a wall: first side that stays inside, else U-turn
                    for cand in ((h + 1) % 4, (h + 3) % 4, (h + 2) % 4) if side_u[r, s] < 0.5 else (
                        (h + 3) % 4,
                        (h + 1) % 4,
                        (h + 2) % 4,
                    ):
                        q = p + dxy[cand]
                        if 0 <= q[0] <= grid and 0 <= q[1] <= grid:
                            hh = cand
                            break
                h, p = hh, q
        th = head * (np.pi / 2)
        lms = rng.integers(0, grid + 1, size=(Lm, 2)).astype(float)
        # odometry: true relative pose in the base frame + noise
        dpos = pos[:, 1:] - pos[:, :-1]
        c, s_ = np.cos(th[:, :-1]), np.sin(th[:, :-1])
        ox = c * dpos[..., 0] + s_ * dpos[..., 1] + rng.normal(0, sigma_xy, size=(R, S - 1))
        oy = -s_ * dpos[..., 0] + c * dpos[..., 1] + rng.normal(0, sigma_xy, size=(R, S - 1))
        oth = ((th[:, 1:] - th[:, :-1] + np.pi) % (2 * np.pi) - np.pi) + rng.normal(0, sigma_th, size=(R, S - 1))
        # ranges: per timestep, per robot: landmarks then higher-index robots
        u_l = rng.random((S, R, Lm)) < p_range
        u_r = rng.random((S, R, R)) < p_range
        u_r &= np.triu(np.ones((R, R), bool), 1)[None]
        noise_l = rng.normal(0, sigma_range, size=(S, R, Lm))
        noise_r = rng.normal(0, sigma_range, size=(S, R, R))
        post = np.transpose(pos, (1, 0, 2))  # (S,R,2)
        d_l = np.linalg.norm(post[:, :, None, :] - lms[None, None], axis=-1) + noise_l
        d_r = np.linalg.norm(post[:, :, None, :] - post[:, None, :, :], axis=-1) + noise_r
        # merged order key: (s, r, kind(0=landmark,1=robot), target)
        sl, rl, ql = np.nonzero(u_l)
        sr, rr, r2 = np.nonzero(u_r)
        a = np.concatenate([rl * S + sl, rr * S + sr])
        b = np.concatenate([P + ql, r2 * S + sr])
        dist = np.concatenate([d_l[sl, rl, ql], d_r[sr, rr, r2]])
        key = np.concatenate(
            [((sl * R + rl) * 2 + 0) * (R + Lm) + ql, ((sr * R + rr) * 2 + 1) * (R + Lm) + r2]
        )
        order = np.argsort(key, kind="stable")
        a, b, dist = a[order], b[order], np.maximum(dist[order], 0.0)
        touched_l = np.zeros(Lm, bool)
        touched_l[b[b >= P] - P] = True
        if touched_l.all():
            break
    return {
        "dim": 2,
        "n_robots": R,
        "n_steps": S,
        "pos": pos,
        "theta": th,
        "landmarks": lms,
        "odom_x": ox,
        "odom_y": oy,
        "odom_theta": oth,
        "k_t": 1.0 / sigma_xy**2,
        "k_r": 1.0 / sigma_th**2,
        "rng_a": a.astype(np.int64),
        "rng_b": b.astype(np.int64),
        "rng_dist": dist,
        "sigma_range": sigma_range,
    }


def manhattan_2d(seed: int, **kw) -> FactorGraphData:
    """One Monte-Carlo instance (config 4) as a ``FactorGraphData``."""
    return arrays_to_factor_graph(manhattan_2d_arrays(seed, **kw))


def arrays_to_factor_graph(arr: dict) -> FactorGraphData:
    assert arr["dim"] == 2
    R, S = arr["n_robots"], arr["n_steps"]
    P = R * S
    fg = FactorGraphData(2)
    names: List[str] = []
    for r in range(R):
        pre = _chain_prefix(r)
        for s in range(S):
            nm = f"{pre}{s}"
            names.append(nm)
            fg.add_pose_variable(
                PoseVariable2D(nm, (arr["pos"][r, s, 0], arr["pos"][r, s, 1]), _wrap(arr["theta"][r, s]), float(s)),
                chain=r,
            )
    for q, lm in enumerate(arr["landmarks"]):
        names.append(f"L{q}")
        fg.add_landmark_variable(LandmarkVariable2D(f"L{q}", (lm[0], lm[1])))
    for r in range(R):
        for s in range(S - 1):
            fg.add_odom_measurement(
                r,
                PoseMeasurement2D(
                    names[r * S + s],
                    names[r * S + s + 1],
                    arr["odom_x"][r, s],
                    arr["odom_y"][r, s],
                    arr["odom_theta"][r, s],
                    arr["k_t"],
                    arr["k_r"],
                    float(s),
                ),
            )
    for a, b, dist in zip(arr["rng_a"], arr["rng_b"], arr["rng_dist"]):
        fg.add_range_measurement(
            FGRangeMeasurement((names[a], names[b]), float(dist), arr["sigma_range"], float(a % S))
        )
    return fg


def _wrap(a: float) -> float:
    return float((a + np.pi) % (2 * np.pi) - np.pi)


def monte_carlo_instance(i: int, n_robots: int = 20, n_steps: int = 100) -> FactorGraphData:
    """Instance ``i`` of the sweep: seed ``20221003 + i``."""
    return manhattan_2d(MC_BASE_SEED + i, n_robots=n_robots, n_steps=n_steps)


# ------------------------------------------------------------------------------------------------
# Large 3D graph
# ------------------------------------------------------------------------------------------------
def _rot_axis(axis: int, sign: int) -> np.ndarray:
    """Exact +-90 degree rotation about a coordinate axis (integer matrix)."""
    R = np.zeros((3, 3))
    a, b = [(1, 2), (2, 0), (0, 1)][axis]
    R[axis, axis] = 1.0
    R[a, b] = -sign
    R[b, a] = sign
    return R


def _so3_exp(w: np.ndarray) -> np.ndarray:
    """Rodrigues formula, batched: w (..., 3) -> (..., 3, 3)."""
    th = np.linalg.norm(w, axis=-1)[..., None, None]
    K = np.zeros(w.shape[:-1] + (3, 3))
    K[..., 0, 1], K[..., 0, 2] = -w[..., 2], w[..., 1]
    K[..., 1, 0], K[..., 1, 2] = w[..., 2], -w[..., 0]
    K[..., 2, 0], K[..., 2, 1] = -w[..., 1], w[..., 0]
    with np.errstate(divide="ignore", invalid="ignore"):
        a = np.where(th > 1e-12, np.sin(th) / th, 1.0 - th * th / 6.0)
        b = np.where(th > 1e-12, (1.0 - np.cos(th)) / (th * th), 0.5 - th * th / 24.0)
    return np.eye(3) + a * K + b * (K @ K)


def grid_3d_arrays(
    seed: int = MC_BASE_SEED,
    n_robots: int = 100,
    n_steps: int = 1000,
    grid: int = 100,
    n_landmarks: int = 1000,
    n_ranges: int = 1_000_000,
    p_turn: float = 0.2,
    sigma_t: float = 0.01,
    sigma_rot: float = 0.002,
    sigma_range: float = 1.0,
) -> dict:
    """One large 3D multi-robot graph: random walks on a grid^3 lattice (unit steps along the body x axis,
    axis-aligned 90-degree turns with probability p_turn, forced turns at the walls), odometry = true relative
    pose with N(0, sigma_t^2) translation noise and exp(N(0, sigma_rot^2 I)) rotation noise, and exactly
    n_ranges distinct range measurements drawn uniformly from all (pose, landmark) and same-timestep
    (pose, pose) pairs, sorted by timestep then keys."""
    rng = np.random.default_rng(seed)
    R, S, Lm = n_robots, n_steps, n_landmarks
    P = R * S
    turns = np.stack([_rot_axis(ax, sg) for ax in (1, 2) for sg in (1, -1)])  # about body y / z
    pos = np.zeros((R, S, 3))
    rot = np.zeros((R, S, 3, 3))
    p = rng.integers(0, grid + 1, size=(R, 3)).astype(np.float64)
    # start orientation: a random element of the 24-element cube group
    Rc = np.tile(np.eye(3), (R, 1, 1))
    for _ in range(3):
        Rc = Rc @ turns[rng.integers(0, 4, size=R)]
    turn_u = rng.random((R, S))
    turn_k = rng.integers(0, 4, size=(R, S))
    for s in range(S):
        pos[:, s] = p
        rot[:, s] = Rc
        if s == S - 1:
            break
        Rn = np.where((turn_u[:, s] < p_turn)[:, None, None], Rc @ turns[turn_k[:, s]], Rc)
        q = p + Rn[:, :, 0]
        bad = ((q < 0) | (q > grid)).any(axis=1)
        # forced turn at a wall: first of the four turns (then the U-turn) whose step stays inside
        for r in np.nonzero(bad)[0]:
            cands = [Rc[r] @ turns[(turn_k[r, s] + j) % 4] for j in range(4)] + [Rc[r] @ turns[0] @ turns[0]]
            for cand in cands:
                qq = p[r] + cand[:, 0]
                if ((qq >= 0) & (qq <= grid)).all():
                    Rn[r], q[r] = cand, qq
                    break
        Rc, p = Rn, q
    lms = rng.integers(0, grid + 1, size=(Lm, 3)).astype(np.float64)
    # odometry in the base frame
    Rt = np.transpose(rot[:, :-1], (0, 1, 3, 2))
    o_t = np.einsum("rsij,rsj->rsi", Rt, pos[:, 1:] - pos[:, :-1]) + rng.normal(0, sigma_t, size=(R, S - 1, 3))
    o_R = (Rt @ rot[:, 1:]) @ _so3_exp(rng.normal(0, sigma_rot, size=(R, S - 1, 3)))
    # ranges: distinct pair codes.  code < P*Lm: (pose, landmark); else same-timestep robot pair
    n_pl = P * Lm
    pairs_r = R * (R - 1) // 2
    n_tot = n_pl + S * pairs_r
    if n_ranges > n_tot:
        raise ValueError("more ranges requested than distinct pairs exist")
    codes = np.zeros(0, np.int64)
    while len(codes) < n_ranges:
        extra = rng.integers(0, n_tot, size=int((n_ranges - len(codes)) * 1.05) + 16)
        codes = np.unique(np.concatenate([codes, extra]))
    if len(codes) > n_ranges:
        codes = np.sort(rng.choice(codes, size=n_ranges, replace=False))
    is_pl = codes < n_pl
    c_pl, c_rr = codes[is_pl], codes[~is_pl] - n_pl
    a_pl, b_pl = c_pl // Lm, P + c_pl % Lm  # pose id is chain-major: r * S + s
    iu, ju = np.triu_indices(R, 1)
    s_rr, pr = c_rr // pairs_r, c_rr % pairs_r
    a_rr, b_rr = iu[pr] * S + s_rr, ju[pr] * S + s_rr
    a = np.concatenate([a_pl, a_rr])
    b = np.concatenate([b_pl, b_rr])
    # order: timestep of the first key, then robot, then (landmark targets before robot targets), then target
    s_key = a % S
    r_key = a // S
    order = np.lexsort((b, r_key, s_key))  # pose targets (b < P) sort before landmark targets (b >= P)
    a, b = a[order], b[order]
    allpos = np.concatenate([pos.reshape(P, 3), lms])
    dist = np.maximum(np.linalg.norm(allpos[a] - allpos[b], axis=1) + rng.normal(0, sigma_range, size=len(a)), 0.0)
    touched = np.zeros(Lm, bool)
    touched[b[b >= P] - P] = True
    if not touched.all():
        raise ValueError("a landmark has no range measurement; increase n_ranges")
    return {
        "dim": 3, "n_robots": R, "n_steps": S, "pos": pos, "rot": rot, "landmarks": lms, "odom_t": o_t, "odom_R": o_R,
        "k_t": 1.0 / sigma_t**2, "k_r": 1.0 / sigma_rot**2, "rng_a": a.astype(np.int64), "rng_b": b.astype(np.int64),
        "rng_dist": dist, "sigma_range": sigma_range,
    }


def grid_3d_factor_graph(arr: dict) -> FactorGraphData:
    """Object form (for the oracle and small tests; the large graph is lowered straight from the arrays)."""
    assert arr["dim"] == 3
    R, S = arr["n_robots"], arr["n_steps"]
    fg = FactorGraphData(3)
    names: List[str] = []
    for r in range(R):
        pre = _chain_prefix(r)
        for s in range(S):
            names.append(f"{pre}{s}")
            fg.add_pose_variable(PoseVariable3D(names[-1], arr["pos"][r, s], arr["rot"][r, s], float(s)), chain=r)
    for q, lm in enumerate(arr["landmarks"]):
        names.append(f"L{q}")
        fg.add_landmark_variable(LandmarkVariable3D(f"L{q}", lm))
    for r in range(R):
        for s in range(S - 1):
            fg.add_odom_measurement(r, PoseMeasurement3D(names[r * S + s], names[r * S + s + 1], arr["odom_t"][r, s],
                                                         arr["odom_R"][r, s], arr["k_t"], arr["k_r"], float(s)))
    for a, b, dist in zip(arr["rng_a"], arr["rng_b"], arr["rng_dist"]):
        fg.add_range_measurement(FGRangeMeasurement((names[a], names[b]), float(dist), arr["sigma_range"], float(a % S)))
    return fg
