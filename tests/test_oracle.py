"""CPU tests of the oracle (oracle/score_oracle.py) against everything that pins it.

The reference ships no tests or golden vectors ("parity unpinned"), so the
pins are: the closed-form layout of SURVEY App. A, the structural checksums of App. B.3, the anchor
optimum objectives of App. C.1, the committed fixtures under tests/golden/ (inputs = the graphs the
reference ships, outputs = the oracle's tight optimum), and oracle-free optimality certificates.
"""
import hashlib
import os

import numpy as np
import pytest

from oracle import score_oracle as so


def _sha16(a):
    return hashlib.sha256(np.asarray(a).astype("<i4").tobytes()).hexdigest()[:16]


# SURVEY.md App. B.3 (QCQP least-squares matrix of the two graphs the reference ships)
CHECKSUMS = {
    "goats": dict(shape=(7184, 7198), nnz=22908, indptr="d04b9dba9a517146", indices="3ab4c6c6a9462b34",
                  sum_values=-862201.098072, sum_abs=877453.724426, sum_w=342395539.5556, first_lm_col=4074,
                  first_dist_col=4082),
    "man4": dict(shape=(11896, 11932), nnz=38880, indptr="377637ae04a8b978", indices="ffe195c1c10c2c61",
                 sum_values=-26142.744512, sum_abs=51941.171485, sum_w=1627922320.0, first_lm_col=9600,
                 first_dist_col=9612),
}
# SURVEY.md App. C.1
ANCHORS = {"goats": 330.48687, "man4": 33.665861}


@pytest.mark.parametrize("name", ["goats", "man4"])
def test_structural_checksums(golden, name):
    fg, _ = golden(name)
    prob = so.assemble(fg, so.QCQP)
    c = CHECKSUMS[name]
    B = prob.B
    assert B.shape == c["shape"] and B.nnz == c["nnz"]
    assert _sha16(B.indptr) == c["indptr"]
    assert _sha16(B.indices) == c["indices"]
    assert abs(B.data.sum() - c["sum_values"]) < 1e-5
    assert abs(np.abs(B.data).sum() - c["sum_abs"]) < 1e-5
    assert abs(prob.w.sum() - c["sum_w"]) < 1e-3
    d = prob.dim
    assert prob.P * d * (d + 1) == c["first_lm_col"]
    assert prob.dist_col0 == c["first_dist_col"]
    # pinned columns 0..5 = (1,0,0,0,1,0): pin_pose, gurobi_utils.py:316-333
    assert np.array_equal(prob.pin_cols, np.arange(6))
    assert np.array_equal(prob.pin_vals, [1, 0, 0, 0, 1, 0])
    nnz_row = np.diff(B.indptr)
    assert nnz_row.min() >= 3 and nnz_row.max() <= 4


def test_goats_first_rows(golden):
    """SURVEY App. A.3 self-check rows (get_relative_pose_cost_expression :504-526, get_single_range_cost :475-501)."""
    fg, _ = golden("goats")
    prob = so.assemble(fg, so.QCQP)
    B = prob.B

    def row(i):
        sl = slice(B.indptr[i], B.indptr[i + 1])
        return B.indices[sl].tolist(), B.data[sl]

    c, v = row(0)
    assert c == [0, 1, 2, 8]
    assert np.allclose(v, [0.042836, 0.056850, -1, 1], atol=1e-6) and prob.w[0] == 2500
    c, v = row(2)
    assert c == [0, 1, 6]
    assert np.allclose(v, [-1.0, 0.000928, 1], atol=1e-6) and prob.w[2] == 125000
    c, v = row(7182)
    assert c == [4070, 4080, 7196]
    assert np.allclose(v, [1, -1, -146.62072], atol=1e-5) and abs(prob.w[7182] - 1.7778) < 1e-4


@pytest.mark.parametrize("name", ["goats", "man4", "man1", "mc0", "mc0_small"])
def test_golden_optimum_is_certified(golden, name):
    """x* of every fixture passes the oracle-free certificate (App. A.7) and reproduces f*."""
    fg, extra = golden(name)
    prob = so.assemble(fg, so.QCQP)
    x = extra["x_star"]
    assert x.shape == (prob.n_cols,)
    f = so.objective(prob, x)
    assert abs(f - float(extra["f_star"])) <= 1e-12 * max(1.0, abs(f))
    kkt = so.kkt_qcqp(prob, x)
    # mc0_small is degenerate (f* = 0: every range is slack at the optimum), so the absolute duality gap
    # of the barrier point (~1e-5, all of it the -sum ||g_delta|| term) is not scaled down by |f|
    assert kkt["rel_kkt"] <= (1e-5 if name == "mc0_small" else 1e-6), kkt
    assert abs(kkt["rel_kkt"] - float(extra["rel_kkt"])) <= 1e-9
    # feasibility: pins and unit balls (add_distance_constraints :341-344)
    assert np.array_equal(x[prob.pin_cols], prob.pin_vals)
    dn = np.linalg.norm(x[prob.dist_col0:].reshape(-1, prob.dim), axis=1)
    assert dn.max() <= 1 + 1e-9
    if name in ANCHORS:
        assert abs(f - ANCHORS[name]) <= 1e-6 * ANCHORS[name]
    assert _sha16(prob.B.indptr) == str(extra["sha_indptr"])
    assert _sha16(prob.B.indices) == str(extra["sha_indices"])


@pytest.mark.parametrize("name", ["mc0_small", "man1"])
def test_barrier_solve_reproduces_fixture(golden, name):
    """The oracle's own solve (seconds at these sizes) lands on the committed optimum."""
    fg, extra = golden(name)
    prob = so.assemble(fg, so.QCQP)
    sol = so.solve_qcqp_barrier(prob)
    x = so.polish_distances(prob, sol.x)
    f_star = float(extra["f_star"])
    assert abs(so.objective(prob, x) - f_star) <= 1e-7 * max(1.0, f_star)
    assert so.kkt_qcqp(prob, x)["rel_kkt"] <= (2e-5 if name == "mc0_small" else 1e-6)  # degenerate f* = 0, see above
    nz = prob.dist_col0
    # pose/landmark part agrees where the optimum is unique (translations of the pinned chain)
    d = prob.dim
    P0 = len(fg.pose_variables[0])
    t_cols = (np.arange(P0)[:, None] * d * (d + 1) + np.arange(d)[None, :] * (d + 1) + d).ravel()
    assert np.abs(x[:nz][t_cols] - extra["x_star"][:nz][t_cols]).max() < 1e-3


@pytest.mark.parametrize("name", ["mc0_small", "man1", "goats"])
def test_qcqp_and_socp_coincide(golden, name):
    """after eliminating the auxiliary variable both relaxations have the same
    objective at the same (R, t, l) (gurobi_utils.py:475-501 with :336-352)."""
    fg, extra = golden(name)
    pq = so.assemble(fg, so.QCQP)
    ps = so.assemble(fg, so.SOCP)
    xq = extra["x_star"]
    xs = so.socp_from_qcqp_solution(pq, ps, xq)
    assert ps.n_cols == pq.dist_col0 + pq.K
    assert abs(so.objective(ps, xs) - so.objective(pq, xq)) <= 1e-9 * max(1.0, so.objective(pq, xq))
    # cone feasibility: ||t_a - t_b|| <= delta_k, delta_k >= 0  (:345-352)
    for k in range(0, pq.K, max(1, pq.K // 50)):
        ta = xs[ps.trans_cols(int(ps.rng_a[k]))]
        tb = xs[ps.trans_cols(int(ps.rng_b[k]))]
        dk = xs[ps.dist_col0 + k]
        assert dk >= 0 and np.linalg.norm(ta - tb) <= dk + 1e-9


def test_socp_layout(golden):
    """SOCP: one scalar distance column per range and one row `delta_k` with rhs = measured distance."""
    fg, _ = golden("mc0_small")
    ps = so.assemble(fg, so.SOCP)
    pq = so.assemble(fg, so.QCQP)
    d = ps.dim
    E = (pq.B.shape[0] - pq.K * d) // (d + d * d)
    assert ps.B.shape == (E * (d + d * d) + ps.K, pq.dist_col0 + ps.K)
    r0 = E * (d + d * d)
    for k in (0, ps.K - 1):
        sl = slice(ps.B.indptr[r0 + k], ps.B.indptr[r0 + k + 1])
        assert ps.B.indices[sl].tolist() == [ps.dist_col0 + k] and ps.B.data[sl].tolist() == [1.0]
        assert ps.b[r0 + k] == fg.range_measurements[k].dist


def test_bad_relaxation_raises(golden):
    fg, _ = golden("mc0_small")
    with pytest.raises(ValueError):
        so.assemble(fg, "LP")


@pytest.mark.parametrize("d", [2, 3])
def test_round_rotations_is_svd_rule(d):
    """round_to_special_orthogonal (matrix_utils.py:59-79): U diag(1,..,det(U Vh)) Vh; result in SO(d)."""
    rng = np.random.default_rng(3)
    M = rng.standard_normal((500, d, d))
    M[::5, :, 0] *= -1
    poses = np.concatenate([M, rng.standard_normal((500, d, 1))], axis=2)
    R = so.round_rotations(poses)
    assert np.abs(R @ R.transpose(0, 2, 1) - np.eye(d)).max() < 1e-12
    assert np.abs(np.linalg.det(R) - 1).max() < 1e-12
    for i in range(0, 500, 25):
        U, _, Vh = np.linalg.svd(M[i])
        Rr = U @ Vh
        if np.linalg.det(Rr) < 0:
            Rr = U @ np.diag([1.0] * (d - 1) + [-1.0]) @ Vh
        assert np.abs(R[i] - Rr).max() < 1e-10
    # nearest rotation: maximises tr(R^T M) over SO(d) -> beats random rotations
    for i in range(20):
        Q, _ = np.linalg.qr(rng.standard_normal((d, d)))
        if np.linalg.det(Q) < 0:
            Q[:, 0] *= -1
        assert np.trace(R[i].T @ M[i]) >= np.trace(Q.T @ M[i]) - 1e-12


REF_EXAMPLES = "/root/reference/examples"


@pytest.mark.skipif(not os.path.isdir(REF_EXAMPLES), reason="reference checkout not present (GPU box)")
@pytest.mark.parametrize("name,path", [("goats", "goats_14_data/goats_14_6_2002_15_20.pkl"),
                                       ("man4", "manhattan/factor_graph.pickle")])
def test_fixture_inputs_match_reference_pickles(golden, name, path):
    """The committed fixture inputs are the reference's shipped pickles, array for array."""
    import score_b200  # noqa: F401  (installs the py_factor_graph shim)
    from py_factor_graph.parsing.parse_pickle_file import parse_pickle_file
    from score_b200.graph_io import graph_to_arrays

    fg_ref = parse_pickle_file(os.path.join(REF_EXAMPLES, path))
    fg_fix, _ = golden(name)
    a, b = graph_to_arrays(fg_ref), graph_to_arrays(fg_fix)
    assert a.keys() == b.keys()
    for k in a:
        assert np.array_equal(np.asarray(a[k]), np.asarray(b[k])), k
    assert len(fg_ref.unconnected_variable_names) == 0
