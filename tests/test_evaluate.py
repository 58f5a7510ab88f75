"""Evaluation step after the path: SE(d)-aligned absolute trajectory error.
CPU tests pin the oracle's alignment on closed-form cases; GPU tests compare the CUDA path (through the C ABI)
with it on ragged batches, degenerate trajectories and the solved golden graphs."""
import numpy as np
import pytest


def _rand_rot(rng, d):
    Q, _ = np.linalg.qr(rng.normal(size=(d, d)))
    if np.linalg.det(Q) < 0:
        Q[:, -1] *= -1
    return Q


def _batch(seed, d, sizes, noise):
    rng = np.random.default_rng(seed)
    ests, gts, Rs, ts = [], [], [], []
    for n in sizes:
        e = rng.normal(size=(n, d)) * 20.0
        R, t = _rand_rot(rng, d), rng.normal(size=d) * 50.0
        g = e @ R.T + t + noise * rng.normal(size=(n, d))
        ests.append(e)
        gts.append(g)
        Rs.append(R)
        ts.append(t)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    return np.concatenate(ests), np.concatenate(gts), off, Rs, ts


# ---------------------------------------------------------------- CPU: the oracle itself
@pytest.mark.parametrize("d", [2, 3])
def test_oracle_alignment_recovers_known_transform(d):
    from oracle import score_oracle as so

    est, gt, off, Rs, ts = _batch(1, d, [50], 0.0)
    rmse, R, t = so.align_trajectory(est, gt)
    assert rmse < 1e-12
    assert np.allclose(R, Rs[0], atol=1e-12) and np.allclose(t, ts[0], atol=1e-10)
    assert abs(np.linalg.det(R) - 1.0) < 1e-12


def test_oracle_alignment_is_a_minimiser_and_never_reflects():
    from oracle import score_oracle as so

    rng = np.random.default_rng(3)
    est = rng.normal(size=(40, 2))
    gt = est * np.array([1.0, -1.0]) + 0.01 * rng.normal(size=(40, 2))  # a mirror image: best SO(2) fit is poor
    rmse, R, t = so.align_trajectory(est, gt)
    assert abs(np.linalg.det(R) - 1.0) < 1e-12 and rmse > 0.3
    for _ in range(50):  # no nearby rigid transform does better
        th = rng.normal() * 0.05
        dR = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
        r2 = np.sqrt(((gt - (est @ (dR @ R).T + t + 0.01 * rng.normal(size=2))) ** 2).sum() / 40)
        assert r2 >= rmse - 1e-12
    r0, R0, t0 = so.align_trajectory(est, gt, align=False)
    assert np.array_equal(R0, np.eye(2)) and np.array_equal(t0, np.zeros(2))
    assert np.isclose(r0, np.sqrt(((gt - est) ** 2).sum() / 40))
    assert np.isnan(so.align_trajectory(np.zeros((0, 3)), np.zeros((0, 3)))[0])


def test_host_gathering_follows_pose_variable_order(golden):
    from score_b200.evaluate import chain_offsets, ground_truth_positions

    fg, _ = golden("man4")
    gt = ground_truth_positions(fg)
    off = chain_offsets(fg)
    assert gt.shape == (1600, 2) and off.tolist() == [0, 400, 800, 1200, 1600]
    assert tuple(gt[400]) == tuple(fg.pose_variables[1][0].true_position)


# ---------------------------------------------------------------- GPU: CUDA path vs oracle
@pytest.mark.gpu
@pytest.mark.parametrize("d", [2, 3])
@pytest.mark.parametrize("align", [True, False])
def test_trajectory_ate_matches_oracle_on_ragged_batch(built_lib, d, align):
    from oracle import score_oracle as so
    from score_b200.solver import trajectory_ate

    sizes = [1, 2, 3, 0, 31, 32, 33, 255, 256, 257, 1000, 2000, 5000, 0, 7]
    est, gt, off, _, _ = _batch(10 + d, d, sizes, 0.3)
    rmse, R, t = trajectory_ate(est, gt, off, align=align)
    for i, n in enumerate(sizes):
        r_o, R_o, t_o = so.align_trajectory(est[off[i]:off[i + 1]], gt[off[i]:off[i + 1]], align=align)
        if n == 0:
            assert np.isnan(rmse[i]) and np.array_equal(R[i], np.eye(d)) and np.array_equal(t[i], np.zeros(d))
            continue
        assert abs(rmse[i] - r_o) <= 1e-10 * max(1.0, r_o), (i, n, rmse[i], r_o)
        if n > d:  # fewer points: the minimiser is not unique, only the error is
            assert np.allclose(R[i], R_o, atol=1e-9) and np.allclose(t[i], t_o, atol=1e-7)
        assert abs(np.linalg.det(R[i]) - 1.0) < 1e-12


@pytest.mark.gpu
def test_trajectory_ate_exact_and_degenerate_cases(built_lib):
    from oracle import score_oracle as so
    from score_b200.solver import trajectory_ate

    # exact rigid motion: zero error, transform recovered
    for d in (2, 3):
        est, gt, off, Rs, ts = _batch(5, d, [300, 40], 0.0)
        rmse, R, t = trajectory_ate(est, gt, off)
        assert rmse.max() < 1e-10
        for i in range(2):
            assert np.allclose(R[i], Rs[i], atol=1e-10) and np.allclose(t[i], ts[i], atol=1e-8)
    # collinear 3D trajectory (rank-1 covariance): rotation not unique, error is
    rng = np.random.default_rng(0)
    s = rng.normal(size=(100, 1))
    est = s * np.array([[1.0, 2.0, -1.0]])
    gt = s * np.array([[0.0, 0.0, np.sqrt(6.0)]]) + 3.0 + 0.05 * rng.normal(size=(100, 3))
    rmse, R, _ = trajectory_ate(est, gt)
    assert abs(rmse[0] - so.align_trajectory(est, gt)[0]) < 1e-9
    assert np.allclose(R[0] @ R[0].T, np.eye(3), atol=1e-12) and abs(np.linalg.det(R[0]) - 1) < 1e-12
    # all points identical (zero covariance): identity rotation, pure translation
    est = np.ones((10, 2))
    gt = np.full((10, 2), 4.0)
    rmse, R, t = trajectory_ate(est, gt)
    assert rmse[0] == 0.0 and np.array_equal(R[0], np.eye(2)) and np.allclose(t[0], 3.0)
    # mirror image: never a reflection
    est = rng.normal(size=(64, 3))
    gt = est * np.array([1.0, 1.0, -1.0])
    rmse, R, _ = trajectory_ate(est, gt)
    assert abs(np.linalg.det(R[0]) - 1.0) < 1e-12
    assert abs(rmse[0] - so.align_trajectory(est, gt)[0]) < 1e-9
    # argument errors follow the library's conventions
    with pytest.raises(ValueError):
        trajectory_ate(np.zeros((4, 4)), np.zeros((4, 4)))
    with pytest.raises(ValueError):
        trajectory_ate(np.zeros((4, 2)), np.zeros((4, 2)), [0, 3, 2])


@pytest.mark.gpu
def test_trajectory_ate_batch_is_bitwise_equal_to_single(built_lib):
    from score_b200.solver import trajectory_ate

    sizes = [700, 13, 2048, 300]
    est, gt, off, _, _ = _batch(77, 3, sizes, 0.1)
    rmse, R, t = trajectory_ate(est, gt, off)
    for i in range(len(sizes)):
        r1, R1, t1 = trajectory_ate(est[off[i]:off[i + 1]], gt[off[i]:off[i + 1]])
        assert r1[0] == rmse[i] and np.array_equal(R1[0], R[i]) and np.array_equal(t1[0], t[i])


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["man4", "goats"])
def test_handle_ate_matches_oracle_and_host_path(built_lib, golden, name):
    """score_eval_ate on the solved handle == oracle alignment of the returned translations, per instance and per
    robot chain; evaluate_ate on the packed SolverResults gives the same numbers."""
    from oracle import score_oracle as so
    from score_b200.evaluate import chain_offsets, evaluate_ate, ground_truth_positions
    from score_b200.lowering import lower_factor_graph
    from score_b200.solve_score import solve_score
    from score_b200.solver import ScoreSolver

    fg, _ = golden(name)
    gt = ground_truth_positions(fg)
    off = chain_offsets(fg)
    d = fg.dimension
    with ScoreSolver(lower_factor_graph(fg, "QCQP")) as s:
        with pytest.raises(RuntimeError):
            s.ate(gt)  # before score_solve: state error
        s.solve()
        poses = s.solution()[0]
        rmse, R, t = s.ate(gt)
        rmse_c, R_c, t_c = s.ate(gt, off)
        with pytest.raises(ValueError):
            s.ate(gt[:-1])
    est = poses[:, :, d]
    r_o, R_o, t_o = so.align_trajectory(est, gt)
    assert abs(rmse[0] - r_o) <= 1e-9 * max(1.0, r_o)
    assert np.allclose(R[0], R_o, atol=1e-9) and np.allclose(t[0], t_o, atol=1e-6)
    for c in range(len(off) - 1):
        r_oc = so.align_trajectory(est[off[c]:off[c + 1]], gt[off[c]:off[c + 1]])[0]
        assert abs(rmse_c[c] - r_oc) <= 1e-9 * max(1.0, r_oc)
    res = solve_score(fg, "QCQP")
    ev = evaluate_ate(res, fg)
    assert abs(ev["rmse"] - rmse[0]) <= 1e-9 * max(1.0, rmse[0])
    evc = evaluate_ate(res, fg, per_chain=True)
    assert np.allclose(evc["rmse"], rmse_c, rtol=1e-9)


@pytest.mark.gpu
def test_handle_ate_on_monte_carlo_batch(built_lib):
    """One trajectory per instance of a batched sweep; unaligned error of the estimate against itself is 0."""
    import bench
    from oracle import score_oracle as so
    from score_b200.solver import ScoreSolver

    prob = bench.make_batch(0, 6, 4, 30)
    rng = np.random.default_rng(0)
    with ScoreSolver(prob) as s:
        s.solve()
        poses = s.solution()[0]
        est = poses[:, :, 2]
        gt = est + 0.05 * rng.normal(size=est.shape)
        rmse, R, t = s.ate(gt)
        zero = s.ate(est, align=False)[0]
    assert np.array_equal(zero, np.zeros(6))
    for i in range(6):
        a, b = prob.pose_off[i], prob.pose_off[i + 1]
        r_o = so.align_trajectory(est[a:b], gt[a:b])[0]
        assert abs(rmse[i] - r_o) <= 1e-10


# ---------------------------------------------------------------- golden fixture: ATE of the oracle optimum x*
def _golden_ate():
    import json
    import os

    from conftest import GOLDEN

    with open(os.path.join(GOLDEN, "ate.json")) as f:
        return json.load(f)


def _xstar_translations(fg, extra):
    from score_b200.evaluate import ground_truth_positions

    d = int(fg.dimension)
    gt = ground_truth_positions(fg)
    est = np.asarray(extra["x_star"])[: len(gt) * d * (d + 1)].reshape(len(gt), d, d + 1)[:, :, d]
    return est, gt


@pytest.mark.parametrize("name", ["goats", "man4", "man1", "mc0_small"])
def test_oracle_reproduces_golden_ate(golden, name):
    """tests/golden/ate.json (make_ate_golden.py) from the stored optimum; GOATS also against the figure SURVEY.md
    8(c) records for the reference's graph: the relaxed estimate sits ~58 m RMSE from ground truth after alignment."""
    from oracle import score_oracle as so
    from score_b200.evaluate import chain_offsets

    fg, extra = golden(name)
    est, gt = _xstar_translations(fg, extra)
    ref = _golden_ate()[name]
    rmse, R, t = so.align_trajectory(est, gt)
    assert abs(rmse - ref["rmse"]) <= 1e-9 * max(1.0, ref["rmse"])
    assert np.allclose(R, ref["R"], atol=1e-10) and np.allclose(t, ref["t"], atol=1e-7)
    off = chain_offsets(fg)
    per = [so.align_trajectory(est[a:b], gt[a:b])[0] for a, b in zip(off[:-1], off[1:])]
    assert np.allclose(per, ref["rmse_per_chain"], rtol=1e-9)
    if name == "goats":
        assert 57.0 < rmse < 59.0


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["goats", "man4", "man1", "mc0_small"])
def test_gpu_ate_matches_golden(built_lib, golden, name):
    """CUDA path on the stored optimum == golden numbers; the GPU's own solve lands on the same error."""
    from score_b200.evaluate import chain_offsets, evaluate_ate
    from score_b200.solve_score import solve_score
    from score_b200.solver import trajectory_ate

    fg, extra = golden(name)
    est, gt = _xstar_translations(fg, extra)
    ref = _golden_ate()[name]
    rmse, R, t = trajectory_ate(est, gt)
    assert abs(rmse[0] - ref["rmse"]) <= 1e-9 * max(1.0, ref["rmse"])
    assert np.allclose(R[0], ref["R"], atol=1e-9) and np.allclose(t[0], ref["t"], atol=1e-6)
    assert abs(trajectory_ate(est, gt, align=False)[0][0] - ref["rmse_unaligned"]) <= 1e-9 * ref["rmse_unaligned"]
    per = trajectory_ate(est, gt, chain_offsets(fg))[0]
    assert np.allclose(per, ref["rmse_per_chain"], rtol=1e-9)
    if name in ("goats", "man1"):  # single chains: the optimum is unique, the GPU solve reproduces its error
        ev = evaluate_ate(solve_score(fg, "QCQP"), fg)
        assert abs(ev["rmse"] - ref["rmse"]) <= 1e-3


def test_trajectory_ate_argument_errors_need_no_device(built_lib):
    """The C entry point validates its arguments before touching the device (error behaviour of the library:
    ValueError for invalid input), and zero trajectories are a no-op."""
    from score_b200.solver import trajectory_ate

    with pytest.raises(ValueError):
        trajectory_ate(np.zeros((4, 4)), np.zeros((4, 4)))  # dimension 4
    with pytest.raises(ValueError):
        trajectory_ate(np.zeros((4, 2)), np.zeros((4, 2)), [0, 3, 2])  # offsets not monotone
    with pytest.raises(ValueError):
        trajectory_ate(np.zeros((4, 2)), np.zeros((5, 2)))  # shapes differ
    with pytest.raises(ValueError):
        trajectory_ate(np.zeros((4, 2)), np.zeros((4, 2)), [0, 9])  # offsets beyond the arrays
    rmse, R, t = trajectory_ate(np.zeros((0, 2)), np.zeros((0, 2)), [0])
    assert rmse.shape == (0,) and R.shape == (0, 2, 2) and t.shape == (0, 2)
