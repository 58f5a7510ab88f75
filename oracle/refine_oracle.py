"""TEST INFRASTRUCTURE ONLY — CPU restatement of the local refinement's cost (csrc/refine.cuh), with a SciPy
least-squares solve to compare against.  Only tests/ may import this.

The cost is the reference's objective (/root/reference/score/utils/gurobi_utils.py:358-526) with R_p in SO(d) and the
auxiliary distance variables eliminated — the original non-convex range-aided SLAM cost the paper's local search
(GTSAM, /root/reference/README.md:63-67) minimises from SCORE's estimate:
    f = sum_edges k ||t_j - t_i - R_i t~||^2 + tau ||R_j - R_i R~||_F^2 + sum_ranges w (||p_a - p_b|| - r~)^2
      + sum_priors w ||l - prior||^2 ,   first pose fixed.
GTSAM itself is not under /root/reference: parity of the refinement is unpinned by the reference; what is checked is
first-order optimality and agreement with scipy.optimize.least_squares from the same start.
Single instances only (a LoweredProblem with n_instances == 1).
"""
import numpy as np


def _hat(w):
    return np.array([[0.0, -w[2], w[1]], [w[2], 0.0, -w[0]], [-w[1], w[0], 0.0]])


def exp_so(w, d):
    """Exp of so(d): d = 2: w is the angle; d = 3: rotation vector (Rodrigues)."""
    if d == 2:
        c, s = np.cos(w[0]), np.sin(w[0])
        return np.array([[c, -s], [s, c]])
    th = np.linalg.norm(w)
    K = _hat(w)
    if th < 1e-12:
        return np.eye(3) + K
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * (K @ K)


def residuals(prob, poses, lms):
    """Weighted residual vector r with f = r.r; poses [P, d, d+1] as [R|t], lms [L, d]."""
    assert prob.n_instances == 1
    d = prob.dim
    R, t = poses[:, :, :d], poses[:, :, d]
    i, j = prob.edge_i, prob.edge_j
    et = np.asarray(prob.edge_t).reshape(-1, d)
    eR = np.asarray(prob.edge_R).reshape(-1, d, d)
    rt = t[j] - t[i] - np.einsum("erc,ec->er", R[i], et)
    rR = R[j] - np.einsum("erm,emc->erc", R[i], eR)
    out = [np.sqrt(prob.edge_k)[:, None] * rt, (np.sqrt(prob.edge_tau)[:, None, None] * rR).reshape(len(i), -1)]
    own = np.concatenate([t, lms], axis=0) if len(lms) else t
    if len(prob.rng_a):
        n = np.linalg.norm(own[prob.rng_a] - own[prob.rng_b], axis=1)
        out.append((np.sqrt(prob.rng_w) * (n - prob.rng_dist))[:, None])
    if len(prob.prior_l):
        pt = np.asarray(prob.prior_t).reshape(-1, d)
        out.append(np.sqrt(prob.prior_w)[:, None] * (lms[prob.prior_l] - pt))
    return np.concatenate([o.ravel() for o in out])


def cost(prob, poses, lms) -> float:
    r = residuals(prob, poses, lms)
    return float(r @ r)


def _retract(prob, poses0, lms0, xi):
    """Tangent coordinates xi = [per free pose (dt, omega)..., per landmark dl...] applied at (poses0, lms0)."""
    d = prob.dim
    nr = 1 if d == 2 else 3
    dof = d + nr
    P = poses0.shape[0]
    poses = poses0.copy()
    for p in range(1, P):  # pose 0 is pinned
        v = xi[(p - 1) * dof : p * dof]
        poses[p, :, d] = poses0[p, :, d] + v[:d]
        poses[p, :, :d] = poses0[p, :, :d] @ exp_so(v[d:], d)
    lms = lms0 + xi[(P - 1) * dof :].reshape(-1, d) if len(lms0) else lms0
    return poses, lms


def tangent_gradient(prob, poses, lms):
    """J^T r in the tangent coordinates of csrc/refine.cuh at (poses, lms), by central differences."""
    d = prob.dim
    dof = d + (1 if d == 2 else 3)
    n = (poses.shape[0] - 1) * dof + lms.size
    g = np.zeros(n)
    h = 1e-6
    for k in range(n):
        e = np.zeros(n)
        e[k] = h
        g[k] = (cost(prob, *_retract(prob, poses, lms, e)) - cost(prob, *_retract(prob, poses, lms, -e))) / (2 * h)
    return 0.5 * g  # d(r.r)/dxi = 2 J^T r


def refine(prob, poses0, lms0, max_nfev=200):
    """scipy.optimize.least_squares (trust-region reflective) from (poses0, lms0); returns (poses, lms, cost)."""
    from scipy.optimize import least_squares

    d = prob.dim
    dof = d + (1 if d == 2 else 3)
    n = (poses0.shape[0] - 1) * dof + lms0.size
    fun = lambda xi: residuals(prob, *_retract(prob, poses0, lms0, xi))
    res = least_squares(fun, np.zeros(n), method="trf", xtol=1e-14, ftol=1e-14, gtol=1e-12, max_nfev=max_nfev)
    poses, lms = _retract(prob, poses0, lms0, res.x)
    return poses, lms, cost(prob, poses, lms)


def tangent_jacobian(prob, poses, lms, h=1e-6):
    """d residuals / d xi at (poses, lms) by central differences: [n_residuals, n_tangent]."""
    d = prob.dim
    dof = d + (1 if d == 2 else 3)
    n = (poses.shape[0] - 1) * dof + lms.size
    cols = []
    for k in range(n):
        e = np.zeros(n)
        e[k] = h
        cols.append((residuals(prob, *_retract(prob, poses, lms, e)) - residuals(prob, *_retract(prob, poses, lms, -e))) / (2 * h))
    return np.stack(cols, axis=1)


def from_device_slots(prob, v):
    """Tangent vector in the device's column-space slots (pose p: first dof entries of its d(d+1) block, then the
    landmarks) -> the xi ordering used here (pose 0, pinned, dropped)."""
    d = prob.dim
    dof, blk = d + (1 if d == 2 else 3), d * (d + 1)
    P = prob.P
    pose = np.asarray(v[: P * blk]).reshape(P, blk)[1:, :dof].ravel()
    return np.concatenate([pose, np.asarray(v[P * blk :]).ravel()])
