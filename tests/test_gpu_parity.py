"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the golden fixtures."""
import hashlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _sha16(a):
    return hashlib.sha256(np.asarray(a).astype("<i4").tobytes()).hexdigest()[:16]


def _solver(fg, relax="QCQP"):
    from score_b200.lowering import lower_factor_graph
    from score_b200.solver import ScoreSolver

    return ScoreSolver(lower_factor_graph(fg, relax))


def _full_x(prob, poses, lms, dist):
    d = prob.dim
    blk = d * (d + 1)
    x = np.zeros(prob.n_cols)
    x[: prob.P * blk] = poses.ravel()
    x[prob.P * blk : prob.P * blk + prob.L * d] = lms.ravel()
    x[prob.dist_col0 :] = dist.ravel()
    return x


@pytest.mark.parametrize("name", ["goats", "man4", "man1", "mc0_small"])
@pytest.mark.parametrize("relax", ["QCQP", "SOCP"])
def test_assembly_bit_exact(built_lib, golden, name, relax):
    """Sparsity pattern / row + column indexing bit-exact against the oracle's restatement of the
    reference variable order; values, weights and rhs to the last bit as well."""
    from oracle import score_oracle as so
    from score_b200 import _lib

    fg, extra = golden(name)
    prob = so.assemble(fg, relax)
    with _solver(fg, relax) as s:
        indptr, indices, values, w, b, shape = s.csr(_lib.SCORE_CSR_FULL, 0)
    B = prob.B
    assert shape == B.shape
    assert np.array_equal(indptr, B.indptr)
    assert np.array_equal(indices, B.indices)
    assert np.array_equal(values, B.data)
    assert np.array_equal(w, prob.w)
    assert np.array_equal(b, prob.b)
    if relax == "QCQP":
        assert _sha16(indptr) == str(extra["sha_indptr"])
        assert _sha16(indices) == str(extra["sha_indices"])


@pytest.mark.parametrize("name", ["goats", "man4"])
def test_reduced_operator_and_transpose(built_lib, golden, name):
    import scipy.sparse as sp
    from oracle import score_oracle as so
    from score_b200 import _lib

    fg, _ = golden(name)
    prob = so.assemble(fg, "QCQP")
    nz = prob.dist_col0
    with _solver(fg) as s:
        s.solve(max_newton=-1)
        ip, ix, v, w, b, shape = s.csr(_lib.SCORE_CSR_REDUCED, 0)
        tip, tix, tv, _, _, tshape = s.csr(_lib.SCORE_CSR_REDUCED_T, 0)
    Bz = prob.B[:, :nz].tocsr()
    Bz.sort_indices()
    assert shape == Bz.shape and tshape == Bz.shape[::-1]
    assert np.array_equal(ip, Bz.indptr) and np.array_equal(ix, Bz.indices) and np.array_equal(v, Bz.data)
    BT = sp.csr_matrix((v, ix, ip), shape=shape).T.tocsr()
    BT.sort_indices()
    assert np.array_equal(tip, BT.indptr) and np.array_equal(tix, BT.indices) and np.array_equal(tv, BT.data)


@pytest.mark.parametrize("d", [2, 3])
def test_round_so_matches_svd(built_lib, d):
    """round_to_special_orthogonal: U diag(1,..,det) V^T  (reference matrix_utils.py:59-79)."""
    from oracle import score_oracle as so
    from score_b200.solver import round_to_special_orthogonal_batch

    rng = np.random.default_rng(7)
    mats = rng.standard_normal((4000, d, d))
    mats[::3] *= 0.01
    mats[1::7] = so.round_rotations(np.concatenate([mats[1::7], np.zeros((len(mats[1::7]), d, 1))], axis=2)) * 0.6
    mats[5::11, :, 0] *= -1  # negative determinants
    got = round_to_special_orthogonal_batch(mats)
    ref = so.round_rotations(np.concatenate([mats, np.zeros((len(mats), d, 1))], axis=2))
    # skip numerically ambiguous inputs (two smallest singular values nearly equal with det<0)
    sv = np.linalg.svd(mats, compute_uv=False)
    ok = (np.linalg.det(mats) > 0) | (sv[:, -2] - sv[:, -1] > 1e-6 * sv[:, 0])
    assert ok.mean() > 0.95
    assert np.abs(got - ref)[ok].max() < 1e-9
    assert np.abs(got @ np.transpose(got, (0, 2, 1)) - np.eye(d)).max() < 1e-12
    assert np.abs(np.linalg.det(got) - 1).max() < 1e-12


@pytest.mark.parametrize("name,relax", [("goats", "QCQP"), ("goats", "SOCP"), ("man4", "QCQP"), ("man1", "QCQP"),
                                        ("mc0", "QCQP"), ("mc0_small", "QCQP")])
def test_solution_parity(built_lib, golden, name, relax):
    """Tolerances from BASELINE.json north_star: relative objective gap <= 1e-4, scaled KKT residuals
    <= 1e-6 (certified independently by the oracle's evaluator on the returned point), translations
    within 1e-3 m of the oracle optimum where the optimum is unique (pinned single chain)."""
    from oracle import score_oracle as so

    fg, extra = golden(name)
    f_star = float(extra["f_star"])
    with _solver(fg, relax) as s:
        st = s.solve()
        poses, rounded, lms, dist = s.solution()
    rec = st.instances[0]
    assert rec["solved"] == 1
    assert abs(rec["objective"] - f_star) <= 1e-4 * max(1.0, abs(f_star))
    prob = so.assemble(fg, relax)
    x = _full_x(prob, poses, lms, dist)
    assert abs(so.objective(prob, x) - rec["objective"]) <= 1e-9 * max(1.0, abs(f_star))
    if relax == "QCQP":
        kkt = so.kkt_qcqp(prob, x)
        assert kkt["rel_kkt"] <= 1e-6, kkt
        assert abs(kkt["rel_kkt"] - rec["rel_kkt"]) <= 1e-2 * rec["rel_kkt"] + 1e-12
    else:
        # SOCP cone feasibility: ||t_a - t_b|| <= delta, delta >= 0
        for k in range(0, prob.K, 37):
            v = x[prob.trans_cols(int(prob.rng_a[k]))] - x[prob.trans_cols(int(prob.rng_b[k]))]
            assert np.linalg.norm(v) <= x[prob.dist_col0 + k] * (1 + 1e-12)
    d = prob.dim
    if name in ("goats", "man1"):
        # ATE (RMS translation error, no alignment needed: pose 0 is pinned in both) within 1e-3 m
        xs = extra["x_star"][: prob.P * d * (d + 1)].reshape(prob.P, d, d + 1)
        err = np.linalg.norm(poses[:, :, d] - xs[:, :, d], axis=1)
        assert np.sqrt(np.mean(err**2)) <= 1e-3
        assert err.max() <= 5e-3
    # rounded rotations: same rule as the oracle's SVD rounding
    ref_round = so.round_rotations(poses)
    assert np.abs(rounded - ref_round).max() < 1e-8


@pytest.mark.parametrize("name", ["goats", "man1"])
def test_tight_solve_trajectory(built_lib, golden, name):
    """Following the central path a little further (rel KKT 1e-8) costs a few more Newton steps and pins
    every translation of the (unique) single-chain optimum to the oracle within 1e-4 m."""
    from oracle import score_oracle as so

    fg, extra = golden(name)
    with _solver(fg) as s:
        st = s.solve(kkt_tol=1e-8)
        poses, _, _, _ = s.solution()
    rec = st.instances[0]
    assert rec["solved"] == 1 and rec["rel_kkt"] <= 1e-8
    d = fg.dimension
    xs = extra["x_star"][: poses.size].reshape(poses.shape)
    assert np.abs(poses[:, :, d] - xs[:, :, d]).max() <= 1e-4
    assert abs(rec["objective"] - float(extra["f_star"])) <= 1e-7 * float(extra["f_star"])


def test_solve_score_entry_point(built_lib, golden):
    """Drop-in call exactly as the reference's users make it (solve_score.py:54-57)."""
    from score.solve_score import solve_score
    from score.utils.gurobi_utils import QCQP_RELAXATION

    fg, extra = golden("man1")
    res = solve_score(fg, QCQP_RELAXATION)
    assert res.solved and res.total_time > 0
    d = fg.dimension
    names = [p.name for c in fg.pose_variables for p in c]
    assert list(res.variables.poses.keys()) == names
    T = res.variables.poses[names[0]]
    assert T.shape == (d + 1, d + 1) and np.allclose(T, np.eye(d + 1), atol=1e-12)  # pinned pose
    for T in list(res.variables.poses.values())[::50]:
        assert np.allclose(T[:d, :d] @ T[:d, :d].T, np.eye(d), atol=1e-9) and abs(np.linalg.det(T[:d, :d]) - 1) < 1e-9
        assert np.allclose(T[d], [0] * d + [1])
    key = (fg.range_measurements[0].first_key, fg.range_measurements[0].second_key)
    assert res.variables.distances[key].shape == (d,)
    assert res.pose_chain_names == fg.get_pose_chain_names()
    # the example's 3-positional-argument form (examples/solve_goats_example_score.py:42-44)
    res2 = solve_score(fg, object(), QCQP_RELAXATION)
    assert res2.solved
    with pytest.raises(ValueError):
        solve_score(fg, "LP")


def test_batch_matches_single_bitwise(built_lib, golden):
    """An instance solved inside a batch gives bit-identical results to solving it alone."""
    from score_b200.lowering import concat, lower_factor_graph
    from score_b200.solver import ScoreSolver

    a = lower_factor_graph(golden("man1")[0])
    b = lower_factor_graph(golden("mc0_small")[0])
    c = lower_factor_graph(golden("goats")[0])
    with ScoreSolver(concat([a, b, c, a])) as s:
        st = s.solve()
        poses, rounded, lms, dist = s.solution()
    assert st.n_solved == 4
    with ScoreSolver(a) as s1:
        st1 = s1.solve()
        p1, r1, l1, d1 = s1.solution()
    assert st1.instances[0]["cg_iters"] == st.instances[0]["cg_iters"] == st.instances[3]["cg_iters"]
    assert np.array_equal(p1, poses[: a.P]) and np.array_equal(p1, poses[-a.P :])
    assert np.array_equal(d1, dist[: a.K]) and np.array_equal(l1, lms[: a.L])
    with ScoreSolver(c) as s3:
        s3.solve()
        p3, _, _, _ = s3.solution()
    assert np.array_equal(p3, poses[a.P + b.P : a.P + b.P + c.P])


@pytest.mark.parametrize("name", ["man4", "mc0", "mc0_small", "grid3d", "grid3d_big", "man21"])
def test_coarse_inverse_matches_dense_reference(built_lib, golden, name):
    """The on-chip coarse level (sorted-list accumulation + register-tiled Gauss-Jordan, coarse.cuh) against a
    dense numpy build of A_c = Z^T H_range Z from the same per-range curvature blocks and frames."""
    from score_b200 import _lib, generators
    from score_b200.lowering import lower_factor_graph

    if name == "grid3d":
        fg = generators.grid_3d_factor_graph(
            generators.grid_3d_arrays(7, n_robots=4, n_steps=30, grid=8, n_landmarks=5, n_ranges=300))
    elif name == "grid3d_big":  # nc = 11 * 12 + 20 * 3 = 192 > kCoarseMax: global-memory build + blocked sweeps (dense.cuh)
        fg = generators.grid_3d_factor_graph(
            generators.grid_3d_arrays(5, n_robots=12, n_steps=20, grid=8, n_landmarks=20, n_ranges=2500))
    elif name == "man21":  # 21 robots: nc = 20 * 6 + 6 * 2 = 132, just past the on-chip limit of 128
        fg = generators.manhattan_2d(generators.MC_BASE_SEED + 5, n_robots=21, n_steps=25)
    else:
        fg, _ = golden(name)
    p = lower_factor_graph(fg)
    d, d1 = p.dim, p.dim + 1
    blk, nm = d * d1, d * (d + 1) // 2
    from score_b200.solver import ScoreSolver

    with ScoreSolver(p) as s:
        st = s.solve()
        Ainv = s.internal(_lib.SCORE_INT_COARSE_INV)
        mk = s.internal(_lib.SCORE_INT_RANGE_CURV).reshape(p.K, nm)
        G = s.internal(_lib.SCORE_INT_FRAMES).reshape(p.P, d, d1)
    assert st.n_solved == 1
    nsegfree, nb = p.n_seg - 1, (p.n_seg - 1) * blk
    nc = nb + p.L * d
    # large coarse spaces without landmark-landmark ranges eliminate the landmark block (Schur complement, dense.cuh):
    # what is stored then is T^-1 = the segment-base block of A_c^-1
    schur = Ainv.size == nb * nb and nb != nc
    assert Ainv.size == (nb * nb if schur else nc * nc)
    assert schur == (name in ("grid3d_big", "man21"))
    Ainv = Ainv.reshape(nb, nb) if schur else Ainv.reshape(nc, nc)
    seg_of_pose = np.searchsorted(p.seg_ptr, np.arange(p.P), side="right") - 1

    def jac(owner):  # d x nc map from coarse coordinates to the translation of an endpoint
        J = np.zeros((d, nc))
        if owner >= p.P:
            q = owner - p.P
            J[:, nb + q * d : nb + (q + 1) * d] = np.eye(d)
        else:
            slot = seg_of_pose[owner] - 1
            if slot >= 0:
                h = np.append(G[owner][:, d], 1.0)
                for r in range(d):
                    J[r, slot * blk + r * d1 : slot * blk + (r + 1) * d1] = h
        return J

    A = np.zeros((nc, nc))
    iu = np.triu_indices(d)
    for k in range(p.K):
        M = np.zeros((d, d))
        M[iu] = mk[k]
        M = M + M.T - np.diag(np.diag(M))
        Ja, Jb = jac(int(p.rng_a[k])), jac(int(p.rng_b[k]))
        Z = Ja - Jb
        A += Z.T @ M @ Z
        # regularisation 1e-6 * 2 w on the diagonal blocks (one term per incidence; a - b when both share a slot)
        w2r = 1e-6 * 2.0 * p.rng_w[k]
        sa = seg_of_pose[p.rng_a[k]] - 1 if p.rng_a[k] < p.P else nsegfree + p.rng_a[k] - p.P
        sb = seg_of_pose[p.rng_b[k]] - 1 if p.rng_b[k] < p.P else nsegfree + p.rng_b[k] - p.P
        if sa == sb:
            A += w2r * Z.T @ Z
        else:
            A += w2r * (Ja.T @ Ja + Jb.T @ Jb)
    for q, wq in zip(p.prior_l, p.prior_w):
        A[nb + q * d + np.arange(d), nb + q * d + np.arange(d)] += 2.0 * wq
    for i in range(nc):
        if not A[i, i] > 0:
            A[i, i] = 1.0
    ref = np.linalg.inv(A)
    assert np.array_equal(Ainv, Ainv.T)
    cond = np.linalg.cond(A)
    if schur:
        ref = ref[:nb, :nb]
        err = np.abs(Ainv - ref).max() / np.abs(ref).max()
        assert err <= max(1e-8, 1e-13 * cond), (err, cond)
        return
    err = np.abs(Ainv - ref).max() / np.abs(ref).max()
    # on-chip path (register-tiled sweeps): 1e-8; the global-memory blocked sweeps (no pivoting either) are checked
    # against the forward-error bound of an inverse, cond * eps with a modest constant
    big = nc > 128
    assert err <= (max(1e-8, 1e-13 * cond) if big else 1e-8 * max(1.0, cond * 1e-8)), (err, cond)
    # residual: an inverse with forward error delta leaves |Ainv A - I| <= delta * cond in the worst case
    assert np.abs(Ainv @ A - np.eye(nc)).max() <= (max(1e-6, 1e-10 * cond) if big else 1e-6), (np.abs(Ainv @ A - np.eye(nc)).max(), cond)


@pytest.mark.parametrize("relax", ["QCQP", "SOCP"])
def test_solution_parity_3d(built_lib, relax):
    """d = 3 (SURVEY 8(d) config 5 at test size): assembly bit-exact, solution certified by the oracle, objective
    equal to the oracle's own barrier solve, SO(3) rounding equal to the SVD rule."""
    from oracle import score_oracle as so
    from score_b200 import _lib, generators

    fg = generators.grid_3d_factor_graph(
        generators.grid_3d_arrays(11, n_robots=3, n_steps=40, grid=8, n_landmarks=5, n_ranges=260))
    prob = so.assemble(fg, relax)
    with _solver(fg, relax) as s:
        indptr, indices, values, w, b, shape = s.csr(_lib.SCORE_CSR_FULL, 0)
        st = s.solve()
        poses, rounded, lms, dist = s.solution()
    assert shape == prob.B.shape
    assert np.array_equal(indptr, prob.B.indptr) and np.array_equal(indices, prob.B.indices)
    assert np.array_equal(values, prob.B.data) and np.array_equal(w, prob.w) and np.array_equal(b, prob.b)
    rec = st.instances[0]
    assert rec["solved"] == 1
    pq, xq, _ = so.solve(fg, so.QCQP)
    f_star = so.objective(pq, xq)
    assert abs(rec["objective"] - f_star) <= 1e-4 * max(1.0, abs(f_star))
    x = _full_x(prob, poses, lms, dist)
    assert abs(so.objective(prob, x) - rec["objective"]) <= 1e-9 * max(1.0, abs(f_star))
    if relax == "QCQP":
        assert so.kkt_qcqp(prob, x)["rel_kkt"] <= 1e-6
    assert np.abs(rounded - so.round_rotations(poses)).max() < 1e-8
    assert np.abs(np.linalg.det(rounded) - 1).max() < 1e-12


def test_large_coarse_space_single_graph_matches_oracle(built_lib):
    """A single 3-D graph whose coarse space (nc = 11 * 12 + 30 * 3 = 222) takes the global-memory path — sorted-list
    accumulation into a dense matrix, blocked symmetric sweeps (dense.cuh), warp-per-row application — against the
    oracle: objective, independent KKT certificate, translations.  (The row-partition test's graph.)"""
    from oracle import score_oracle as so
    from score_b200 import generators

    fg = generators.grid_3d_factor_graph(
        generators.grid_3d_arrays(3, n_robots=12, n_steps=60, grid=12, n_landmarks=30, n_ranges=6000))
    prob = so.assemble(fg, so.QCQP)
    with _solver(fg) as s:
        st = s.solve()
        poses, rounded, lms, dist = s.solution()
        assert s.internal(0).size == 132 * 132  # SCORE_INT_COARSE_INV: coarse level on, landmark block eliminated (nb = 11 * 12)
    rec = st.instances[0]
    assert rec["solved"] == 1
    pq, xq, _ = so.solve(fg, so.QCQP)
    f_star = so.objective(pq, xq)
    assert abs(rec["objective"] - f_star) <= 1e-6 * max(1.0, abs(f_star))
    x = _full_x(prob, poses, lms, dist)
    assert abs(so.objective(prob, x) - rec["objective"]) <= 1e-9 * max(1.0, abs(f_star))
    kkt = so.kkt_qcqp(prob, x)
    assert kkt["rel_kkt"] <= 1e-6, kkt
    d = 3
    xs = xq[: pq.P * d * (d + 1)].reshape(pq.P, d, d + 1)
    # robots are tied to each other only through ranges (the relaxed optimum is flat in their relative placement):
    # compare translations on the pinned robot's chain, where the optimum is well determined
    err = np.linalg.norm(poses[:60, :, d] - xs[:60, :, d], axis=1)
    assert err.max() <= 1e-3
    assert np.abs(rounded - so.round_rotations(poses)).max() < 1e-8


def test_batch_with_large_coarse_instances_matches_oracle(built_lib):
    """A batch that mixes instances below and above the on-chip coarse limit (21 robots: nc = 132 > 128): every
    instance gets a coarse level (no silent fall-off), is certified by the oracle, and is bit-identical to solving
    it alone."""
    from oracle import score_oracle as so
    from score_b200 import generators
    from score_b200.lowering import concat, lower_factor_graph
    from score_b200.solver import ScoreSolver

    fgs = [generators.manhattan_2d(generators.MC_BASE_SEED + 5, n_robots=21, n_steps=25),
           generators.manhattan_2d(generators.MC_BASE_SEED + 1, n_robots=4, n_steps=30),
           generators.manhattan_2d(generators.MC_BASE_SEED + 6, n_robots=24, n_steps=20),
           generators.manhattan_2d(generators.MC_BASE_SEED + 2, n_robots=20, n_steps=25)]
    probs = [lower_factor_graph(fg) for fg in fgs]
    with ScoreSolver(concat(probs)) as s:
        st = s.solve()
        poses, rounded, lms, dist = s.solution()
        ncs = [int(round(np.sqrt(s.internal(0, i).size))) for i in range(4)]
    # (the two large instances store the inverse of their Schur complement on the segment bases)
    assert ncs == [20 * 6, 3 * 6 + 12, 23 * 6, 19 * 6 + 12]
    assert st.n_solved == 4
    po = np.cumsum([0] + [p.P for p in probs])
    lo = np.cumsum([0] + [p.L for p in probs])
    ko = np.cumsum([0] + [p.K for p in probs])
    for i, fg in enumerate(fgs):
        prob = so.assemble(fg, so.QCQP)
        x = _full_x(prob, poses[po[i]:po[i + 1]], lms[lo[i]:lo[i + 1]], dist[ko[i]:ko[i + 1]])
        rec = st.instances[i]
        assert abs(so.objective(prob, x) - rec["objective"]) <= 1e-9 * max(1.0, abs(rec["objective"]))
        assert so.kkt_qcqp(prob, x)["rel_kkt"] <= 1e-6
        if i in (0, 2):
            pq, xq, _ = so.solve(fg, so.QCQP)
            f_star = so.objective(pq, xq)
            assert abs(rec["objective"] - f_star) <= 1e-6 * max(1.0, abs(f_star))
    with ScoreSolver(probs[0]) as s1:
        st1 = s1.solve()
        p1 = s1.solution()[0]
    assert st1.instances[0]["cg_iters"] == st.instances[0]["cg_iters"]
    assert np.array_equal(p1, poses[: probs[0].P])


def test_stream_group_matches_single_handle_bitwise(built_lib, golden):
    """Sub-batches solved concurrently on their own streams (ScoreSolverGroup, incl. the pipelined
    create -> solve -> read-back path) give bit-identical per-instance results to one handle."""
    from score_b200 import generators
    from score_b200.lowering import concat, lower_factor_graph, lower_manhattan_arrays
    from score_b200.solver import ScoreSolver, ScoreSolverGroup

    probs = [lower_manhattan_arrays(generators.manhattan_2d_arrays(generators.MC_BASE_SEED + i, n_robots=4, n_steps=30))
             for i in range(7)]
    batch = concat(probs)
    with ScoreSolver(batch) as s:
        st = s.solve()
        ref = s.solution()
    assert st.n_solved == 7
    with ScoreSolverGroup(batch, n_streams=3) as g:
        stg = g.solve()
        got = g.solution()
        stp, piped, h2d, d2h = g.run_pipelined()
    assert stg.n_solved == 7 and stp.n_solved == 7 and h2d > 0 and d2h > 0
    assert np.array_equal(stg.instances["cg_iters"], st.instances["cg_iters"])
    for a, b, c in zip(ref, got, piped):
        assert np.array_equal(a, b) and np.array_equal(a, c)
    # more sub-batches than host threads (staggered queue): same bits again
    with ScoreSolverGroup(batch, n_streams=2, n_parts=5) as g:
        assert len(g.parts) == 5 and g.n_streams == 2
        stq = g.solve()
        gotq = g.solution()
        _, pipedq, h2d1, d2h1 = g.run_pipelined()
        _, piped2, h2d2, d2h2 = g.run_pipelined(steps=3)  # three steps streamed as one job queue
        assert (h2d2, d2h2) == (h2d1, d2h1) and d2h1 == d2h  # bytes are reported per step
        for a, c in zip(ref, piped2):
            assert np.array_equal(a, c)
    assert stq.n_solved == 7 and np.array_equal(stq.instances["cg_iters"], st.instances["cg_iters"])
    for a, b, c in zip(ref, gotq, pipedq):
        assert np.array_equal(a, b) and np.array_equal(a, c)


def test_intermediate_iterates(built_lib, golden):
    """solve_problem_with_intermediate_iterates (solve_score.py:89-116): one SolverResults per outer-iteration cap;
    the objective decreases towards the optimum and the last entry is the converged solve."""
    from score.solve_score import solve_problem_with_intermediate_iterates

    fg, extra = golden("mc0_small")
    its = solve_problem_with_intermediate_iterates(fg, "QCQP", max_iterates=60)
    assert 3 <= len(its) <= 60
    assert its[-1].solved and not its[0].solved
    costs = [r.solver_cost for r in its]
    assert costs[-1] <= costs[0] + 1e-12
    names = [p.name for c in fg.pose_variables for p in c]
    assert all(list(r.variables.poses.keys()) == names for r in its)


def _edge_case_graph(kind):
    """Small 2D graphs exercising the rarely-hit branches of the reference model build."""
    import score_b200  # noqa: F401
    from py_factor_graph.factor_graph import FactorGraphData
    from py_factor_graph.measurements import FGRangeMeasurement, PoseMeasurement2D
    from py_factor_graph.priors import LandmarkPrior2D
    from py_factor_graph.variables import LandmarkVariable2D, PoseVariable2D

    rng = np.random.default_rng(42)
    fg = FactorGraphData(2)
    n = 30
    th = np.cumsum(rng.normal(0, 0.3, n))
    pos = np.cumsum(np.stack([np.cos(th), np.sin(th)], 1), 0)
    for i in range(n):
        fg.add_pose_variable(PoseVariable2D(f"A{i}", tuple(pos[i]), float(th[i]), float(i)), chain=0)
    for q, p in enumerate([(3.0, 8.0), (12.0, -4.0), (20.0, 6.0)]):
        fg.add_landmark_variable(LandmarkVariable2D(f"L{q}", p))

    def rel(i, j, noise=0.02):
        c, s = np.cos(th[i]), np.sin(th[i])
        d = pos[j] - pos[i]
        return PoseMeasurement2D(f"A{i}", f"A{j}", c * d[0] + s * d[1] + rng.normal(0, noise),
                                 -s * d[0] + c * d[1] + rng.normal(0, noise), th[j] - th[i] + rng.normal(0, 0.01),
                                 1.0 / noise**2, 1e4, float(i))

    skip = {14} if kind == "broken_chain" else set()
    for i in range(n - 1):
        if i not in skip:
            fg.add_odom_measurement(0, rel(i, i + 1))
    if kind in ("loop_closures", "broken_chain"):
        for i, j in [(2, 20), (5, 27), (10, 16)]:
            fg.add_loop_closure(rel(i, j, 0.05))
    if kind != "no_ranges":
        lms = np.array([l.true_position for l in fg.landmark_variables])
        for i in range(0, n, 2):
            for q in range(3):
                dist = np.linalg.norm(pos[i] - lms[q]) + rng.normal(0, 0.5)
                fg.add_range_measurement(FGRangeMeasurement((f"A{i}", f"L{q}"), max(dist, 0.0), 0.5, float(i)))
        fg.add_range_measurement(FGRangeMeasurement(("L0", "L1"), 0.0, 1.0, 0.0))  # dist == 0: all-zero delta column
    else:
        fg.landmark_variables.clear()
        fg.existing_landmark_variables.clear() if hasattr(fg, "existing_landmark_variables") else None
    if kind == "priors":
        fg.add_landmark_prior(LandmarkPrior2D("L0", (3.5, 7.5), 4.0))
        fg.add_landmark_prior(LandmarkPrior2D("L2", (19.0, 6.5), 0.25))
    assert len(fg.unconnected_variable_names) == 0
    return fg


@pytest.mark.parametrize("kind", ["priors", "loop_closures", "broken_chain", "no_ranges"])
@pytest.mark.parametrize("relax", ["QCQP", "SOCP"])
def test_edge_case_graphs_match_oracle(built_lib, kind, relax):
    """Landmark priors (get_all_landmark_prior_costs :433-446), loop closures (:407-430), an odometry chain with a
    missing link, a range with dist == 0, and a graph without ranges (add_distance_variables early return :281-283):
    assembly bit-exact, optimum equal to the oracle's."""
    from oracle import score_oracle as so
    from score_b200 import _lib

    fg = _edge_case_graph(kind)
    prob = so.assemble(fg, relax)
    with _solver(fg, relax) as s:
        indptr, indices, values, w, b, shape = s.csr(_lib.SCORE_CSR_FULL, 0)
        st = s.solve()
        poses, rounded, lms, dist = s.solution()
    assert shape == prob.B.shape
    assert np.array_equal(indptr, prob.B.indptr) and np.array_equal(indices, prob.B.indices)
    assert np.array_equal(values, prob.B.data) and np.array_equal(w, prob.w) and np.array_equal(b, prob.b)
    rec = st.instances[0]
    assert rec["solved"] == 1
    pq, xq, _ = so.solve(fg, so.QCQP)
    f_star = so.objective(pq, xq)
    assert abs(rec["objective"] - f_star) <= 1e-6 * max(1.0, abs(f_star))
    x = _full_x(prob, poses, lms, dist)
    assert abs(so.objective(prob, x) - rec["objective"]) <= 1e-9 * max(1.0, abs(f_star))
    if relax == "QCQP":
        assert so.kkt_qcqp(prob, x)["rel_kkt"] <= 1e-6
    nz = pq.P * 6
    assert np.abs(poses.ravel() - xq[:nz]).max() <= 1e-3  # single pinned chain: unique optimum


@pytest.mark.parametrize("name", ["goats", "man4", "grid3d", "loops"])
def test_matrix_free_operator_matches_assembled_pair(built_lib, golden, name):
    """PCG iterations with the factor-wise Hessian-vector product (k_hessvec, the default) against the same solve on the
    assembled CSR pair (row pass + column pass): same optimum to rounding, both certified by the oracle's evaluator.
    `loops` has relative-pose factors that are not odometry links (loop closures across a broken chain)."""
    from oracle import score_oracle as so
    from score_b200 import generators

    if name == "grid3d":
        fg = generators.grid_3d_factor_graph(
            generators.grid_3d_arrays(11, n_robots=3, n_steps=40, grid=8, n_landmarks=5, n_ranges=260))
    elif name == "loops":
        fg = _edge_case_graph("broken_chain")
    else:
        fg, _ = golden(name)
    prob = so.assemble(fg, so.QCQP)
    out = {}
    with _solver(fg) as s:
        for mode in (0, 1):
            st = s.solve(operator_mode=mode, kkt_tol=1e-8)
            out[mode] = (st.instances[0].copy(), [a.copy() for a in s.solution()])
    for mode in (0, 1):
        rec, (poses, rounded, lms, dist) = out[mode]
        assert rec["solved"] == 1
        assert so.kkt_qcqp(prob, _full_x(prob, poses, lms, dist))["rel_kkt"] <= 1e-8
    f0, f1 = out[0][0]["objective"], out[1][0]["objective"]
    assert abs(f0 - f1) <= 1e-7 * max(1.0, abs(f1))
    d = prob.dim
    if name in ("goats", "loops"):  # single pinned chain: unique optimum
        assert np.abs(out[0][1][0][:, :, d] - out[1][1][0][:, :, d]).max() <= 1e-4


def test_fused_pcg_kernel_is_bit_identical_to_lockstep_ticks(built_lib, golden):
    """The opt-in fused per-instance PCG kernel (fused.cuh) calls the same block-level bodies with the same block
    decomposition as the lockstep tick kernels: identical bits (both on the assembled CSR pair)."""
    from score_b200.lowering import concat, lower_factor_graph
    from score_b200.solver import ScoreSolver

    batch = concat([lower_factor_graph(golden(n)[0]) for n in ("man1", "mc0_small", "goats")])
    with ScoreSolver(batch) as s:
        st0 = s.solve(operator_mode=1)
        ref = [a.copy() for a in s.solution()]
        st1 = s.solve(operator_mode=1, tail_threshold=8)
        got = s.solution()
    assert st0.n_solved == 3 and st1.n_solved == 3
    assert np.array_equal(st0.instances["cg_iters"], st1.instances["cg_iters"])
    for a, b in zip(ref, got):
        assert np.array_equal(a, b)


def test_pipelined_steps_match_plain_solve_bitwise(built_lib):
    """ScoreSolverGroup.solve_steps (steps double-buffered over two sets of handles, no barrier between steps): every
    step's per-instance records and the final solution equal a plain single-handle solve bit for bit."""
    from score_b200 import generators
    from score_b200.lowering import concat, lower_manhattan_arrays
    from score_b200.solver import ScoreSolver, ScoreSolverGroup

    batch = concat([lower_manhattan_arrays(generators.manhattan_2d_arrays(generators.MC_BASE_SEED + i, n_robots=4, n_steps=30))
                    for i in range(6)])
    with ScoreSolver(batch) as s:
        st = s.solve()
        ref = s.solution()
    with ScoreSolverGroup(batch, n_streams=2) as g:
        steps = g.solve_steps(5, n_sets=2)
        got = g.solution()
    assert len(steps) == 5
    for k in steps:
        assert k.n_solved == 6
        assert np.array_equal(k.instances["cg_iters"], st.instances["cg_iters"])
        assert np.array_equal(k.instances["objective"], st.instances["objective"])
    for a, b in zip(ref, got):
        assert np.array_equal(a, b)


def test_dynamic_sweep_queue_matches_plain_solve_bitwise(built_lib):
    """sharding.SweepQueue over pooled handles (device-resident) and over create -> solve -> read-back jobs (host buffers):
    every job's per-instance records and read-back arrays equal a plain single-handle solve of the same instances."""
    from score_b200 import generators
    from score_b200.lowering import concat, lower_manhattan_arrays, slice_instances
    from score_b200.sharding import SweepQueue
    from score_b200.solver import HandlePool, ScoreSolver, StreamedJobs

    batch = concat([lower_manhattan_arrays(generators.manhattan_2d_arrays(generators.MC_BASE_SEED + i, n_robots=3 + i % 2,
                                                                           n_steps=24 + 2 * i)) for i in range(7)])
    cuts = [0, 2, 5, 7]
    parts = [slice_instances(batch, a, b) for a, b in zip(cuts[:-1], cuts[1:])]
    with ScoreSolver(batch) as s:
        st = s.solve()
        ref = s.solution()
    assert st.n_solved == 7
    with HandlePool(parts, copies=2) as pool:
        costs = [float(w.cycles) for w in pool.warm()]
        with SweepQueue(len(parts), lambda step, part, w: pool.solve(part), inflight=4) as q:
            res = q.run(4, costs)
    assert sorted((s_, p) for s_, p, _ in res) == [(s_, p) for s_ in range(4) for p in range(3)]
    for _, p, r in res:
        a, b = cuts[p], cuts[p + 1]
        assert r.n_solved == b - a
        assert np.array_equal(r.instances["cg_iters"], st.instances["cg_iters"][a:b])
        assert np.array_equal(r.instances["objective"], st.instances["objective"][a:b])
    # host-buffer jobs: every worker reads back into its own slot; the slot of the last job of a worker holds that part
    d = batch.dim
    big = lambda: (np.empty((batch.P, d, d + 1)), np.empty((batch.P, d, d)), np.empty((batch.L, d)),
                   np.empty((batch.K, batch.dist_per)))
    outs = [big() for _ in range(3)]
    jobs = StreamedJobs(parts, outs, inflight=2)
    last = {}

    def job(step, part, w):
        r = jobs(step, part, w)
        last[w] = (part, [v.copy() for v in jobs.views(w, parts[part])])
        return r

    with SweepQueue(len(parts), job, inflight=3) as q:
        res = q.run(2, costs)
    assert len(res) == 6 and jobs.h2d_bytes > 0 and jobs.d2h_bytes > 0
    for w, (p, arrs) in last.items():
        a, b = cuts[p], cuts[p + 1]
        assert np.array_equal(arrs[0], ref[0][batch.pose_off[a]:batch.pose_off[b]])
        assert np.array_equal(arrs[1], ref[1][batch.pose_off[a]:batch.pose_off[b]])
        assert np.array_equal(arrs[2], ref[2][batch.lm_off[a]:batch.lm_off[b]])
        assert np.array_equal(arrs[3], ref[3][batch.rng_off[a]:batch.rng_off[b]])


@pytest.mark.parametrize("index", [462, 4136])
def test_weakly_active_sweep_instances_certify(built_lib, index):
    """Two Monte-Carlo sweep instances whose optimum has weakly active range terms (the Newton decrement stalls on their
    kinks: 4136 needs a tighter forcing term at a held barrier parameter, 462 a smaller barrier parameter): both must
    reach the 1e-6 certificate, and the oracle's evaluator must agree on the returned point."""
    from oracle import score_oracle as so
    from score_b200 import generators
    from score_b200.lowering import lower_manhattan_arrays
    from score_b200.solver import ScoreSolver

    seed = generators.MC_BASE_SEED + index
    prob = lower_manhattan_arrays(generators.manhattan_2d_arrays(seed, n_robots=20, n_steps=100), "QCQP", with_names=False)
    with ScoreSolver(prob) as s:
        st = s.solve()
        poses, _, lms, dist = s.solution()
    assert st.n_solved == 1 and st.instances[0]["rel_kkt"] <= 1e-6
    op = so.assemble(generators.manhattan_2d(seed, n_robots=20, n_steps=100), so.QCQP)
    x = _full_x(op, poses, lms, dist)
    assert so.kkt_qcqp(op, x)["rel_kkt"] <= 1.5e-6  # tolerance of the test: the device's 1e-6 plus evaluator rounding
