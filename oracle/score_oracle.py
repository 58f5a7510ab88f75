"""CPU float64 oracle for SCORE's ``solve_score`` path.  TEST INFRASTRUCTURE ONLY.

This module restates, in numpy/scipy, the convex problem that the reference
builds with gurobipy and solves with Gurobi's barrier, and solves it tightly
with a log-barrier Newton method.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may import it; the
product path (``score_b200``) never does and has no CPU fallback.

PARITY UNPINNED: the reference ships no tests, golden vectors or expected
objective values, and its own solver (gurobipy, closed source, licence) plus
its data model (py_factor_graph, un-vendored) cannot run offline.  The oracle is
therefore pinned only by (i) the structural checksums of the assembled
least-squares matrix and (ii) the anchor optimum objectives recorded in
SURVEY.md App. B.3 / C.1, both reproduced in tests/.

Every function cites the reference lines it follows (paths relative to
/root/reference).
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

QCQP = "QCQP"
SOCP = "SOCP"


@dataclass
class OracleProblem:
    """Weighted least squares  f(x) = sum_i w_i (B x - b)_i^2  over x in C."""

    dim: int
    relaxation: str
    B: sp.csr_matrix  # unweighted rows
    w: np.ndarray  # row weights
    b: np.ndarray  # right-hand side
    n_cols: int
    P: int
    L: int
    K: int
    pose_names: List[str]
    landmark_names: List[str]
    range_keys: List[Tuple[str, str]]
    pin_cols: np.ndarray
    pin_vals: np.ndarray
    # SOCP only: per range the translation owners (global ids) for the cone
    rng_a: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int64))
    rng_b: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int64))
    rng_dist: np.ndarray = field(default_factory=lambda: np.zeros(0))
    rng_w: np.ndarray = field(default_factory=lambda: np.zeros(0))

    @property
    def dist_col0(self) -> int:
        d = self.dim
        return self.P * d * (d + 1) + self.L * d

    def trans_cols(self, owner: int) -> np.ndarray:
        """Columns of the translation owned by pose p (id p) or landmark q (id P+q).

        Follows VariableCollection.get_translation_var,
        score/utils/gurobi_utils.py:103-109.
        """
        d = self.dim
        if owner < self.P:
            return owner * d * (d + 1) + np.arange(d) * (d + 1) + d
        return self.P * d * (d + 1) + (owner - self.P) * d + np.arange(d)


def _check_valid_relaxation(relaxation: str) -> None:
    # score/utils/gurobi_utils.py:139-144
    if relaxation not in (SOCP, QCQP):
        raise ValueError(
            f"Relaxation {relaxation} is not supported. Acceptable relaxations are {[SOCP, QCQP]}"
        )


def assemble(fg, relaxation: str = QCQP) -> OracleProblem:
    """Variables, pin and objective rows in the reference's creation order.

    Columns: add_pose_variables / add_landmark_variables / add_distance_variables
    (score/utils/gurobi_utils.py:233-310).  Rows: get_full_cost_objective order
    odometry chains, loop closures, ranges, landmark priors (:358-377) with the
    per-factor algebra of get_relative_pose_cost_expression (:504-526),
    get_single_range_cost (:475-501) and get_all_landmark_prior_costs (:433-446).
    Pin: pin_pose on fg.pose_variables[0][0] (:181-183, :316-333).
    """
    _check_valid_relaxation(relaxation)
    d = fg.dimension
    if not isinstance(d, int) or d not in (2, 3):
        raise ValueError(f"Value {d} is not 2 or 3")
    blk = d * (d + 1)

    pose_col: Dict[str, int] = {}
    pose_names: List[str] = []
    for chain in fg.pose_variables:
        for pose in chain:
            if pose.name in pose_col:
                raise ValueError(f"Variable name {pose.name} already exists in pose_vars")
            pose_col[pose.name] = len(pose_names) * blk
            pose_names.append(pose.name)
    P = len(pose_names)
    lm_col: Dict[str, int] = {}
    landmark_names: List[str] = []
    for lm in fg.landmark_variables:
        if lm.name in pose_col:
            raise ValueError(f"Variable name {lm.name} already exists in pose_vars")
        if lm.name in lm_col:
            raise ValueError(f"Variable name {lm.name} already exists in landmark_vars")
        lm_col[lm.name] = P * blk + len(landmark_names) * d
        landmark_names.append(lm.name)
    L = len(landmark_names)
    range_keys: List[Tuple[str, str]] = []
    dist_col: Dict[Tuple[str, str], int] = {}
    per = d if relaxation == QCQP else 1
    for meas in fg.range_measurements:
        key = (meas.first_key, meas.second_key)
        if key in dist_col:
            if relaxation == QCQP:
                raise ValueError(f"Variable name {key} already exists in distance_vars")
        dist_col[key] = P * blk + L * d + len(range_keys) * per
        range_keys.append(key)
    K = len(range_keys)
    n_cols = P * blk + L * d + K * per

    def R_col(name: str, r: int, c: int) -> int:
        return pose_col[name] + r * (d + 1) + c

    def t_cols(name: str) -> List[int]:
        if name in pose_col:
            return [pose_col[name] + r * (d + 1) + d for r in range(d)]
        if name in lm_col:
            return [lm_col[name] + r for r in range(d)]
        raise ValueError(f"Variable name {name} not found")

    def owner(name: str) -> int:
        if name in pose_col:
            return pose_col[name] // blk
        if name in lm_col:
            return P + (lm_col[name] - P * blk) // d
        raise ValueError(f"Variable name {name} not found")

    rows: List[int] = []
    cols: List[int] = []
    vals: List[float] = []
    w: List[float] = []
    b: List[float] = []

    def new_row(weight: float, rhs: float, entries: List[Tuple[int, float]]) -> None:
        r = len(w)
        for c, v in entries:
            rows.append(r)
            cols.append(c)
            vals.append(v)
        w.append(weight)
        b.append(rhs)

    def add_relative_pose(meas) -> None:
        i, j = meas.base_pose, meas.to_pose
        if i not in pose_col or j not in pose_col:
            raise KeyError(i if i not in pose_col else j)
        t_meas = np.asarray(meas.translation_vector, dtype=np.float64)
        R_meas = np.asarray(meas.rotation_matrix, dtype=np.float64)
        ti, tj = t_cols(i), t_cols(j)
        # k * || t_j - t_i - R_i t_meas ||^2
        for r in range(d):
            ent = [(tj[r], 1.0), (ti[r], -1.0)]
            ent += [(R_col(i, r, c), -float(t_meas[c])) for c in range(d)]
            new_row(float(meas.translation_precision), 0.0, ent)
        # tau * || R_j - R_i R_meas ||_F^2
        for r in range(d):
            for c in range(d):
                ent = [(R_col(j, r, c), 1.0)]
                ent += [(R_col(i, r, m), -float(R_meas[m, c])) for m in range(d)]
                new_row(float(meas.rotation_precision), 0.0, ent)

    for chain in fg.odom_measurements:
        for meas in chain:
            add_relative_pose(meas)
    for meas in fg.loop_closure_measurements:
        add_relative_pose(meas)
    rng_a = np.zeros(K, np.int64)
    rng_b = np.zeros(K, np.int64)
    rng_dist = np.zeros(K)
    rng_w = np.zeros(K)
    for k, meas in enumerate(fg.range_measurements):
        key = (meas.first_key, meas.second_key)
        ta, tb = t_cols(key[0]), t_cols(key[1])
        rng_a[k], rng_b[k] = owner(key[0]), owner(key[1])
        rng_dist[k], rng_w[k] = float(meas.dist), float(meas.precision)
        dc = dist_col[key]
        if relaxation == QCQP:
            # w * || t_a - t_b - dist * delta ||^2
            for r in range(d):
                new_row(rng_w[k], 0.0, [(ta[r], 1.0), (tb[r], -1.0), (dc + r, -rng_dist[k])])
        else:
            # w * (delta - dist)^2
            new_row(rng_w[k], rng_dist[k], [(dc, 1.0)])
    for prior in fg.landmark_priors:
        tc = t_cols(prior.name)
        tv = np.asarray(prior.translation_vector, dtype=np.float64)
        for r in range(d):
            new_row(float(prior.translation_precision), float(tv[r]), [(tc[r], 1.0)])

    n_rows = len(w)
    B = sp.csr_matrix(
        (np.asarray(vals, np.float64), (np.asarray(rows, np.int64), np.asarray(cols, np.int64))),
        shape=(n_rows, n_cols),
    )
    # canonical CSR: indices ascending within a row, explicit zeros kept
    B.sort_indices()

    first = fg.pose_variables[0][0]  # IndexError if empty, like the reference (:181)
    pin_cols = np.array(
        [pose_col[first.name] + r * (d + 1) + c for r in range(d) for c in range(d + 1)], np.int64
    )
    pin_vals = np.array([1.0 if r == c else 0.0 for r in range(d) for c in range(d + 1)])
    return OracleProblem(
        dim=d,
        relaxation=relaxation,
        B=B,
        w=np.asarray(w),
        b=np.asarray(b),
        n_cols=n_cols,
        P=P,
        L=L,
        K=K,
        pose_names=pose_names,
        landmark_names=landmark_names,
        range_keys=range_keys,
        pin_cols=pin_cols,
        pin_vals=pin_vals,
        rng_a=rng_a,
        rng_b=rng_b,
        rng_dist=rng_dist,
        rng_w=rng_w,
    )


def objective(prob: OracleProblem, x: np.ndarray) -> float:
    r = prob.B @ x - prob.b
    return float(np.sum(prob.w * r * r))


# ----------------------------------------------------------------------------------------
# Tight solve: log-barrier Newton.  Interior point, like Gurobi's barrier
# (score/solve_score.py:76), but run to a far tighter tolerance than BarQCPConvTol=1e-1.
# ----------------------------------------------------------------------------------------
@dataclass
class OracleSolution:
    x: np.ndarray
    objective: float
    newton_steps: int
    seconds: float
    mu_final: float


def _qcqp_as_reduced(prob: OracleProblem):
    """Eliminate pinned columns: returns A (sqrt(W) B on free cols), r0, free index."""
    n = prob.n_cols
    free = np.ones(n, bool)
    free[prob.pin_cols] = False
    free_idx = np.nonzero(free)[0]
    x_pin = np.zeros(n)
    x_pin[prob.pin_cols] = prob.pin_vals
    sw = np.sqrt(prob.w)
    A = sp.diags(sw) @ prob.B[:, free_idx]
    r0 = sw * (prob.B @ x_pin - prob.b)
    return A.tocsc(), r0, free_idx, x_pin


def solve_qcqp_barrier(
    prob: OracleProblem,
    mu_final: float = 1e-11,
    newton_tol: float = 1e-9,
    max_newton: int = 400,
    verbose: bool = False,
    stop_kkt: Optional[float] = None,
) -> OracleSolution:
    """min f(x) s.t. ||delta_k|| <= 1, pin — add_distance_constraints QCQP branch
    (score/utils/gurobi_utils.py:341-344) + objective (:358-377)."""
    assert prob.relaxation == QCQP
    t0 = time.perf_counter()
    d = prob.dim
    A, r0, free_idx, x_pin = _qcqp_as_reduced(prob)
    nf = A.shape[1]
    H0 = (2.0 * (A.T @ A)).tocsc()
    g0 = 2.0 * (A.T @ r0)
    # position of delta blocks inside the reduced vector
    col_map = -np.ones(prob.n_cols, np.int64)
    col_map[free_idx] = np.arange(nf)
    K = prob.K
    dcols = col_map[prob.dist_col0 + np.arange(K * d)].reshape(K, d)
    assert (dcols >= 0).all()

    x = np.zeros(nf)  # delta = 0 strictly feasible

    def f_val(xv):
        r = A @ xv + r0
        return float(r @ r)

    # Schur complement on the (pose, landmark) block: the delta block of the Newton matrix is
    # block diagonal (one d x d block per range), so it is eliminated exactly and only the
    # smaller SPD matrix S = Hzz - Hzd Dk^-1 Hdz is factorised.
    is_d = np.zeros(nf, bool)
    is_d[dcols.ravel()] = True
    zi = np.nonzero(~is_d)[0]
    di = dcols.ravel()  # delta unknowns in range order
    H0r = H0.tocsr()
    Hzz = H0r[zi][:, zi].tocsc()
    Hzd = H0r[zi][:, di].tocsr()
    Hdz = Hzd.T.tocsr()
    Hdd0 = np.zeros((K, d, d))
    Hdd_full = H0r[di][:, di].tocoo()
    blk_of = Hdd_full.row // d
    assert np.array_equal(blk_of, Hdd_full.col // d), "delta block of the Hessian must be block diagonal"
    np.add.at(Hdd0, (blk_of, Hdd_full.row % d, Hdd_full.col % d), Hdd_full.data)
    kd = K * d
    bd_indptr = np.arange(0, kd * d + 1, d)
    bd_indices = (np.repeat(np.arange(K) * d, d * d).reshape(K, d, d) + np.arange(d)[None, None, :]).ravel()
    eye = np.eye(d)[None]

    steps = 0
    mu = 1.0
    while True:
        for _ in range(max_newton):
            dl = x[dcols]  # K x d
            s = 1.0 - np.sum(dl * dl, axis=1)
            grad = H0 @ x + g0
            gb = (2.0 * mu / s)[:, None] * dl
            np.add.at(grad, dcols.ravel(), gb.ravel())
            # barrier blocks: mu * (2/s I + 4 dl dl^T / s^2)
            blocks = Hdd0 + (2.0 * mu / s)[:, None, None] * eye + (4.0 * mu / (s * s))[:, None, None] * (
                dl[:, :, None] * dl[:, None, :]
            )
            # ranges with dist == 0 have an all-zero column: only the barrier term keeps them SPD
            Dinv = sp.csr_matrix((np.linalg.inv(blocks).ravel(), bd_indices, bd_indptr), shape=(kd, kd))
            S = (Hzz - (Hzd @ Dinv @ Hdz)).tocsc()
            gz, gd_ = grad[zi], grad[di]
            # S is SPD: symmetric ordering + no pivoting (~9x faster than the unsymmetric default)
            lu = spla.splu(S, permc_spec="MMD_AT_PLUS_A", diag_pivot_thresh=0.0, options=dict(SymmetricMode=True))
            dzv = lu.solve(-(gz - Hzd @ (Dinv @ gd_)))
            ddv = -(Dinv @ (gd_ + Hdz @ dzv))
            dx = np.zeros(nf)
            dx[zi] = dzv
            dx[di] = ddv
            steps += 1
            dec = float(-grad @ dx)
            # backtracking keeping strict feasibility
            phi0 = f_val(x) - mu * np.sum(np.log(s))
            t = 1.0
            ddl = dx[dcols]
            while True:
                xn = x + t * dx
                dn = dl + t * ddl
                sn = 1.0 - np.sum(dn * dn, axis=1)
                if (sn > 0).all():
                    phin = f_val(xn) - mu * np.sum(np.log(sn))
                    if phin <= phi0 - 1e-4 * t * dec or t < 1e-12:
                        break
                t *= 0.5
            x = xn
            if verbose:
                print(f"mu={mu:.1e} step={steps} t={t:.2e} dec={dec:.3e} f={f_val(x):.9f}")
            if dec * t < newton_tol * max(1.0, abs(phi0)) or dec < 1e-14:
                break
        if mu <= mu_final:
            break
        if stop_kkt is not None and mu <= 1e-6:
            # certificate-driven stop (bench.py's CPU arm): end the path as soon as the polished point certifies
            xc = x_pin.copy()
            xc[free_idx] = x
            if kkt_qcqp(prob, polish_distances(prob, xc))["rel_kkt"] <= stop_kkt:
                break
        mu *= 0.1
    xf = x_pin.copy()
    xf[free_idx] = x
    return OracleSolution(
        x=xf, objective=objective(prob, xf), newton_steps=steps, seconds=time.perf_counter() - t0, mu_final=mu
    )


def reduced_from_qcqp(prob: OracleProblem, x: np.ndarray):
    """Split x into pose blocks (P,d,d+1), landmarks (L,d), delta (K,d)."""
    d = prob.dim
    blk = d * (d + 1)
    poses = x[: prob.P * blk].reshape(prob.P, d, d + 1)
    lms = x[prob.P * blk : prob.P * blk + prob.L * d].reshape(prob.L, d)
    per = d if prob.relaxation == QCQP else 1
    dist = x[prob.dist_col0 :].reshape(prob.K, per)
    return poses, lms, dist


def polish_distances(prob: OracleProblem, x: np.ndarray) -> np.ndarray:
    """Replace the auxiliary variables by their exact minimisers given (R,t,l)
    (SURVEY.md App. A.4): QCQP delta = proj_ball((t_a-t_b)/dist), SOCP delta = max(dist, ||t_a-t_b||)."""
    x = x.copy()
    d = prob.dim
    for k in range(prob.K):
        v = x[prob.trans_cols(int(prob.rng_a[k]))] - x[prob.trans_cols(int(prob.rng_b[k]))]
        nv = np.linalg.norm(v)
        r = prob.rng_dist[k]
        if prob.relaxation == QCQP:
            c0 = prob.dist_col0 + k * d
            if r > 0:
                u = v / r
                nu = np.linalg.norm(u)
                x[c0 : c0 + d] = u if nu <= 1 else u / nu
            # r == 0: column of zeros, delta irrelevant; leave as is
        else:
            x[prob.dist_col0 + k] = max(r, nv)
    return x


def socp_from_qcqp_solution(prob_q: OracleProblem, prob_s: OracleProblem, xq: np.ndarray) -> np.ndarray:
    """Both relaxations share the optimal (R,t,l)."""
    d = prob_q.dim
    nz = prob_q.P * d * (d + 1) + prob_q.L * d
    xs = np.zeros(prob_s.n_cols)
    xs[:nz] = xq[:nz]
    return polish_distances(prob_s, xs)


# ----------------------------------------------------------------------------------------
# Optimality certificate, evaluated on the unscaled problem
# ----------------------------------------------------------------------------------------
def project_C(prob: OracleProblem, x: np.ndarray) -> np.ndarray:
    """Pin clamp + per-range ball (QCQP).  SOCP cone handled in kkt_socp."""
    x = x.copy()
    x[prob.pin_cols] = prob.pin_vals
    if prob.relaxation == QCQP:
        d = prob.dim
        dl = x[prob.dist_col0 :].reshape(prob.K, d)
        nrm = np.linalg.norm(dl, axis=1)
        scale = np.where(nrm > 1.0, 1.0 / np.maximum(nrm, 1e-300), 1.0)
        x[prob.dist_col0 :] = (dl * scale[:, None]).ravel()
    return x


def kkt_qcqp(prob: OracleProblem, x: np.ndarray, y: Optional[np.ndarray] = None) -> Dict[str, float]:
    assert prob.relaxation == QCQP
    res = prob.B @ x - prob.b
    y_link = 2.0 * prob.w * res
    if y is None:
        y = y_link
    g = prob.B.T @ y
    r_link = np.linalg.norm(y - y_link) / (1.0 + np.linalg.norm(y))
    r_stat = np.linalg.norm(x - project_C(prob, x - g)) / (1.0 + np.linalg.norm(x))
    p = float(np.sum(prob.w * res * res))
    d = prob.dim
    gd = g[prob.dist_col0 :].reshape(prob.K, d)
    with np.errstate(divide="ignore", invalid="ignore"):
        quad = np.where(prob.w > 0, y * y / prob.w, 0.0)
    D = -0.25 * float(np.sum(quad)) - float(prob.b @ y) + float(g[prob.pin_cols] @ prob.pin_vals)
    D -= float(np.sum(np.linalg.norm(gd, axis=1)))
    r_gap = abs(p - D) / (1.0 + abs(p) + abs(D))
    free = np.ones(prob.n_cols, bool)
    free[prob.pin_cols] = False
    free[prob.dist_col0 :] = False
    return {
        "r_link": float(r_link),
        "r_stat": float(r_stat),
        "r_gap": float(r_gap),
        "rel_kkt": float(max(r_link, r_stat, r_gap)),
        "primal": p,
        "dual": D,
        "dual_infeas": float(np.linalg.norm(g[free])),
    }


def round_rotations(poses: np.ndarray) -> np.ndarray:
    """round_to_special_orthogonal per pose (score/utils/matrix_utils.py:59-79):
    U diag(1,..,det(U Vh)) Vh."""
    P, d, _ = poses.shape
    out = np.zeros((P, d, d))
    for p in range(P):
        U, _, Vh = np.linalg.svd(poses[p, :, :d])
        R = U @ Vh
        if np.linalg.det(R) < 0:
            R = U @ np.diag([1.0] * (d - 1) + [-1.0]) @ Vh
        out[p] = R
    return out


def align_trajectory(est: np.ndarray, gt: np.ndarray, align: bool = True):
    """Checker for the evaluation step: the rigid alignment
    min_{R in SO(d), t} sum ||gt_i - (R est_i + t)||^2 by the Kabsch SVD — the same U diag(1,..,det) Vh rule as
    round_to_special_orthogonal (score/utils/matrix_utils.py:59-79) applied to the cross-covariance — and the
    absolute trajectory error after it.  Returns (rmse, R, t); an empty trajectory gives (nan, I, 0)."""
    est, gt = np.asarray(est, float), np.asarray(gt, float)
    n, d = est.shape
    if n == 0:
        return float("nan"), np.eye(d), np.zeros(d)
    if not align:
        return float(np.sqrt(((gt - est) ** 2).sum() / n)), np.eye(d), np.zeros(d)
    me, mg = est.mean(0), gt.mean(0)
    H = (gt - mg).T @ (est - me)
    U, _, Vh = np.linalg.svd(H)
    R = U @ Vh
    if np.linalg.det(R) < 0:
        R = U @ np.diag([1.0] * (d - 1) + [-1.0]) @ Vh
    t = mg - R @ me
    res = gt - (est @ R.T + t)
    return float(np.sqrt((res ** 2).sum() / n)), R, t


def extract(prob: OracleProblem, x: np.ndarray):
    """VariableCollection.get_variable_values (score/utils/gurobi_utils.py:114-136):
    homogeneous (d+1)x(d+1) poses with rounded rotation and untouched translation."""
    d = prob.dim
    poses, lms, dist = reduced_from_qcqp(prob, x)
    Rr = round_rotations(poses)
    out_p = {}
    for p, name in enumerate(prob.pose_names):
        T = np.eye(d + 1)
        T[:d, :d] = Rr[p]
        T[:d, d] = poses[p, :, d]
        out_p[name] = T
    out_l = {name: lms[q].copy() for q, name in enumerate(prob.landmark_names)}
    out_d = {key: dist[k].copy() for k, key in enumerate(prob.range_keys)}
    return out_p, out_l, out_d


def solve(fg, relaxation: str = QCQP, **kw):
    """End-to-end oracle: returns (OracleProblem, x*, OracleSolution)."""
    _check_valid_relaxation(relaxation)
    assert len(fg.unconnected_variable_names) == 0  # score/solve_score.py:28-32
    prob_q = assemble(fg, QCQP)
    sol = solve_qcqp_barrier(prob_q, **kw)
    xq = polish_distances(prob_q, sol.x)
    if relaxation == QCQP:
        return prob_q, xq, sol
    prob_s = assemble(fg, SOCP)
    xs = socp_from_qcqp_solution(prob_q, prob_s, xq)
    return prob_s, xs, sol
