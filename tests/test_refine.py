"""Local refinement after the relaxation: score_refine against the CPU restatement of its cost
(oracle/refine_oracle.py) and scipy.optimize.least_squares from the same start."""
import numpy as np
import pytest


def _graph(d, seed=3):
    from score_b200 import generators
    from score_b200.lowering import lower_grid3d_arrays, lower_manhattan_arrays

    if d == 2:
        arr = generators.manhattan_2d_arrays(generators.MC_BASE_SEED + seed, n_robots=3, n_steps=14)
        return lower_manhattan_arrays(arr, "QCQP", with_names=False), arr
    arr = generators.grid_3d_arrays(seed, n_robots=3, n_steps=10, grid=6, n_landmarks=4, n_ranges=90)
    return lower_grid3d_arrays(arr), arr


@pytest.mark.parametrize("d", [2, 3])
def test_refine_oracle_cost_is_the_reference_objective_on_so_d(d):
    """CPU only: the refinement cost at (poses, landmarks) IS the reference's objective (oracle restatement of
    gurobi_utils.py:358-526) evaluated with every distance variable on the unit sphere along its range (delta = v / |v|,
    which the relaxation only bounds by |delta| <= 1); plus a finite-difference check of the tangent parametrisation."""
    from oracle import refine_oracle as ro
    from oracle import score_oracle as so

    prob, _ = _graph(d)
    rng = np.random.default_rng(0)
    P, L = prob.P, prob.L
    poses = np.zeros((P, d, d + 1))
    for p in range(P):
        poses[p, :, :d] = ro.exp_so(rng.normal(size=1 if d == 2 else 3), d)
        poses[p, :, d] = rng.normal(size=d) * 3
    lms = rng.normal(size=(L, d)) * 3
    f = ro.cost(prob, poses, lms)
    r = ro.residuals(prob, poses, lms)
    assert r.shape[0] == prob.E * (d + d * d) + prob.K + prob.Lp * d and np.isclose(f, r @ r)
    # the reference's (restated) QCQP objective at x = (poses, landmarks, delta_k = unit vector from b to a): the same number
    from score_b200 import generators

    fg = (generators.manhattan_2d(generators.MC_BASE_SEED + 3, n_robots=3, n_steps=14) if d == 2
          else generators.grid_3d_factor_graph(_graph(3)[1]))
    op = so.assemble(fg, so.QCQP)
    own = np.concatenate([poses[:, :, d], lms], axis=0)
    v = own[prob.rng_a] - own[prob.rng_b]
    x = np.concatenate([poses.ravel(), lms.ravel(), (v / np.linalg.norm(v, axis=1, keepdims=True)).ravel()])
    assert x.shape[0] == op.n_cols and np.isclose(so.objective(op, x), f, rtol=1e-12)
    # gradient check of the tangent parametrisation: directional derivative by finite differences
    g = ro.tangent_gradient(prob, poses, lms)
    xi = rng.normal(size=g.shape) * 1e-5
    f1 = ro.cost(prob, *ro._retract(prob, poses, lms, xi))
    assert abs((f1 - f) - 2 * g @ xi) <= 1e-3 * abs(2 * g @ xi) + 1e-9 * (1 + f)


@pytest.mark.gpu
@pytest.mark.parametrize("d", [2, 3])
def test_refine_reaches_a_stationary_point_and_matches_least_squares(built_lib, d):
    from oracle import refine_oracle as ro
    from score_b200.solver import ScoreSolver

    prob, _ = _graph(d)
    with ScoreSolver(prob) as s:
        st = s.solve()
        assert st.n_solved == 1
        relaxed, rounded, lms0, _ = s.solution()
        rec, stats = s.refine()
        poses, lms = s.refined()
    poses0 = relaxed.copy()
    poses0[:, :, :d] = rounded  # the start the device used: rounded rotations, relaxed translations / landmarks
    f0 = ro.cost(prob, poses0, lms0)
    f1 = ro.cost(prob, poses, lms)
    assert stats["n_converged"] == 1 and rec["accepted_steps"][0] >= 1
    assert np.isclose(rec["cost_initial"][0], f0, rtol=1e-9) and np.isclose(rec["cost_final"][0], f1, rtol=1e-9)
    assert f1 < f0
    # rotations stay in SO(d); the pinned pose does not move
    R = poses[:, :, :d]
    assert np.abs(R @ np.transpose(R, (0, 2, 1)) - np.eye(d)).max() < 1e-10 and np.all(np.linalg.det(R) > 0.999)
    assert np.array_equal(poses[0], poses0[0])
    # first-order optimality in the tangent coordinates
    g = ro.tangent_gradient(prob, poses, lms)
    assert np.linalg.norm(g) <= 1e-4 * (1.0 + f1)
    # same basin as SciPy's trust-region solve from the same start (tolerance of the test: 1e-6 relative cost)
    _, _, f_ref = ro.refine(prob, poses0, lms0)
    assert f1 <= f_ref * (1 + 1e-6) + 1e-9


@pytest.mark.gpu
def test_refine_batch_matches_single_and_custom_start(built_lib):
    """A batch refines every instance as if alone (bit-identical costs); a caller-supplied start is honoured."""
    from score_b200 import generators
    from score_b200.lowering import concat, lower_manhattan_arrays
    from score_b200.solver import ScoreSolver

    probs = [lower_manhattan_arrays(generators.manhattan_2d_arrays(generators.MC_BASE_SEED + 10 + i, n_robots=3 + i % 2,
                                                                   n_steps=16 + 2 * i), "QCQP", with_names=False) for i in range(4)]
    singles = []
    for p in probs:
        with ScoreSolver(p) as s:
            s.solve()
            rec, st1 = s.refine()
            singles.append((rec[0], s.refined(), st1["n_converged"]))
    with ScoreSolver(concat(probs)) as s:
        s.solve()
        rec, stats = s.refine()
        poses, lms = s.refined()
        assert stats["n_converged"] == sum(c for _, _, c in singles)
        batch = concat(probs)
        for i, (r1, (p1, l1), _) in enumerate(singles):
            assert rec[i]["cost_final"] == r1["cost_final"] and rec[i]["outer_iterations"] == r1["outer_iterations"]
            assert np.array_equal(poses[batch.pose_off[i]:batch.pose_off[i + 1]], p1)
            assert np.array_equal(lms[batch.lm_off[i]:batch.lm_off[i + 1]], l1)
        # restart from the refined point: already stationary, nothing to gain
        rec2, _ = s.refine(init=(poses, lms))
        assert np.all(rec2["cost_final"] <= rec["cost_final"] * (1 + 1e-12))
        assert np.allclose(rec2["cost_initial"], rec["cost_final"], rtol=1e-12)
    with pytest.raises(ValueError):
        with ScoreSolver(probs[0]) as s:
            s.solve()
            s.refine(init=(np.zeros((1, 2, 3)), np.zeros((1, 2))))


@pytest.mark.gpu
def test_refine_before_solve_is_an_error(built_lib):
    from score_b200.solver import ScoreSolver

    prob, _ = _graph(2)
    with ScoreSolver(prob) as s:
        with pytest.raises(RuntimeError):
            s.refine()
        with pytest.raises(RuntimeError):
            s.refined()


@pytest.mark.gpu
def test_refinement_improves_the_trajectory_error(built_lib):
    """The relaxation contracts the trajectory (~58 m RMSE on GOATS before any local search); from its
    rounded estimate the refinement must land much closer to the ground truth (SE(2)-aligned ATE, whole instance)."""
    from score_b200 import generators
    from score_b200.evaluate import trajectory_ate
    from score_b200.lowering import lower_manhattan_arrays
    from score_b200.solver import ScoreSolver

    arr = generators.manhattan_2d_arrays(generators.MC_BASE_SEED + 1, n_robots=4, n_steps=40)
    prob = lower_manhattan_arrays(arr, "QCQP", with_names=False)
    gt = arr["pos"].reshape(-1, 2)
    with ScoreSolver(prob) as s:
        s.solve()
        relaxed = s.solution()[0]
        rec, stats = s.refine()
        poses, _ = s.refined()
    before = float(trajectory_ate(relaxed[:, :, 2], gt)[0][0])
    after = float(trajectory_ate(poses[:, :, 2], gt)[0][0])
    assert rec["cost_final"][0] < rec["cost_initial"][0]
    assert after < 0.5 * before and after < 1.0  # metres; range noise is 1 m, odometry 1 cm per step


@pytest.mark.gpu
def test_solve_and_refine_entry_point(built_lib):
    """score.solve_score.solve_and_refine on FactorGraphData objects: relaxed and refined SolverResults in input order."""
    from score.solve_score import solve_and_refine
    from score_b200 import generators

    graphs = [generators.manhattan_2d(generators.MC_BASE_SEED + 20 + i, n_robots=3, n_steps=12 + 3 * i) for i in range(2)]
    relaxed, refined, rec = solve_and_refine(graphs, "QCQP")
    assert len(relaxed) == len(refined) == 2 and len(rec) == 2
    for fg, a, b, r in zip(graphs, relaxed, refined, rec):
        assert a.solved and list(a.poses) == list(b.poses) and a.pose_chain_names == b.pose_chain_names
        assert r["cost_final"] < r["cost_initial"] and np.isclose(b.solver_cost, r["cost_final"])
        first = fg.pose_variables[0][0].name
        assert np.array_equal(b.poses[first], np.eye(3))  # pin_pose: the first pose stays [I | 0]
        for name, T in b.poses.items():
            assert np.allclose(T[:2, :2] @ T[:2, :2].T, np.eye(2), atol=1e-10) and np.allclose(T[2], [0, 0, 1])
        for key, dv in b.variables.distances.items():
            assert dv.shape == (2,) and abs(np.linalg.norm(dv) - 1.0) < 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("d", [2, 3])
def test_chain_and_block_jacobi_preconditioners_agree(built_lib, d):
    """The odometry-chain preconditioner (block LDL^T per segment, the default) only changes how fast the PCG solves
    converge: same minimum as with block-Jacobi, in far fewer kernel launches."""
    from score_b200.solver import ScoreSolver

    prob, _ = _graph(d)
    with ScoreSolver(prob) as s:
        s.solve()
        rec_c, st_c = s.refine()
        pc, lc = s.refined()
        rec_j, st_j = s.refine(preconditioner=1)
        pj, lj = s.refined()
    assert st_c["n_converged"] == 1 and st_j["n_converged"] == 1
    assert abs(rec_c["cost_final"][0] - rec_j["cost_final"][0]) <= 1e-6 * rec_j["cost_final"][0]
    assert np.abs(pc - pj).max() < 1e-2 and np.abs(lc - lj).max() < 1e-2  # flat directions of the minimum (metres)
    assert st_c["kernel_launches"] < st_j["kernel_launches"] / 3


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["priors", "loop_closures", "broken_chain", "no_ranges"])
def test_refine_on_edge_case_graphs(built_lib, kind):
    """Landmark priors, loop closures (relative-pose factors that are not odometry links), a chain broken into two
    segments, no ranges at all, a zero-distance landmark-landmark range: the device gradient of the linearisation
    matches finite differences of the CPU restatement, and the refinement ends at a stationary point with a lower cost."""
    from oracle import refine_oracle as ro
    from score_b200 import _lib
    from score_b200.lowering import lower_factor_graph
    from score_b200.solver import ScoreSolver
    from test_gpu_parity import _edge_case_graph

    prob = lower_factor_graph(_edge_case_graph(kind), "QCQP")
    d = 2
    with ScoreSolver(prob) as s:
        assert s.solve().n_solved == 1
        relaxed, rounded, lms0, _ = s.solution()
        poses0 = relaxed.copy()
        poses0[:, :, :d] = rounded
        s.refine(max_outer=1, max_inner=1)  # one linearisation at the start point
        g = ro.from_device_slots(prob, s.internal(_lib.SCORE_INT_REF_GRAD))
        J = ro.tangent_jacobian(prob, poses0, lms0)
        g_ref = J.T @ ro.residuals(prob, poses0, lms0)
        assert np.linalg.norm(g - g_ref) <= 1e-6 * (1.0 + np.linalg.norm(g_ref))
        dg = ro.from_device_slots(prob, s.internal(_lib.SCORE_INT_REF_DIAG))
        assert np.allclose(dg, np.diag(J.T @ J), rtol=1e-5, atol=1e-6 * np.abs(np.diag(J.T @ J)).max())
        rec, stats = s.refine()
        poses, lms = s.refined()
    f0, f1 = ro.cost(prob, poses0, lms0), ro.cost(prob, poses, lms)
    assert np.isclose(rec["cost_initial"][0], f0, rtol=1e-9) and np.isclose(rec["cost_final"][0], f1, rtol=1e-9)
    assert f1 <= f0 and stats["n_converged"] == 1
    assert np.linalg.norm(ro.tangent_gradient(prob, poses, lms)) <= 1e-4 * (1.0 + f1)
