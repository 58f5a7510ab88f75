"""CPU tests of the host side: py_factor_graph shim, name->index lowering, batching, entry-point
argument/error behaviour (mirrors /root/reference/score/solve_score.py and utils/gurobi_utils.py)."""
import os
import pickle

import numpy as np
import pytest
import scipy.sparse as sp

import score_b200  # noqa: F401  (installs the shim when py_factor_graph is absent)
from oracle import score_oracle as so
from py_factor_graph.factor_graph import FactorGraphData
from py_factor_graph.measurements import FGRangeMeasurement, PoseMeasurement2D, PoseMeasurement3D
from py_factor_graph.priors import LandmarkPrior2D
from py_factor_graph.utils.matrix_utils import round_to_special_orthogonal
from py_factor_graph.utils.solver_utils import SolverResults, VariableValues, save_to_tum
from py_factor_graph.variables import LandmarkVariable2D, PoseVariable2D
from score_b200 import generators
from score_b200.graph_io import arrays_to_graph, graph_to_arrays, load_graph_npz, robot_subgraph, save_graph_npz
from score_b200.lowering import concat, lower_factor_graph, lower_manhattan_arrays


def csr_from_lowered(p, inst=0):
    """Reference-ordered QCQP least-squares matrix rebuilt from the ScoreProblemDesc arrays alone
    (include/score_b200.h "Column / row order"), to check the lowering against the oracle."""
    d = p.dim
    blk = d * (d + 1)
    P0, P1 = p.pose_off[inst], p.pose_off[inst + 1]
    L0, L1 = p.lm_off[inst], p.lm_off[inst + 1]
    Pn, Ln = P1 - P0, L1 - L0
    rows, cols, vals, w, b = [], [], [], [], []
    r = 0

    def tcol(owner, c):
        return owner * blk + c * (d + 1) + d if owner < Pn else Pn * blk + (owner - Pn) * d + c

    for e in range(p.edge_off[inst], p.edge_off[inst + 1]):
        i, j = int(p.edge_i[e]), int(p.edge_j[e])
        for a in range(d):  # translation rows: t_j - t_i - R_i t~
            ent = {j * blk + a * (d + 1) + d: 1.0, i * blk + a * (d + 1) + d: -1.0}
            for c in range(d):
                ent[i * blk + a * (d + 1) + c] = -p.edge_t[e][c]
            for cc, v in ent.items():
                rows.append(r), cols.append(cc), vals.append(v)
            w.append(p.edge_k[e]), b.append(0.0)
            r += 1
        for a in range(d):  # rotation rows: R_j - R_i R~
            for c in range(d):
                ent = {j * blk + a * (d + 1) + c: 1.0}
                for m in range(d):
                    ent[i * blk + a * (d + 1) + m] = -p.edge_R[e][m, c]
                for cc, v in ent.items():
                    rows.append(r), cols.append(cc), vals.append(v)
                w.append(p.edge_tau[e]), b.append(0.0)
                r += 1
    k0 = p.rng_off[inst]
    for k in range(k0, p.rng_off[inst + 1]):
        for c in range(d):
            rows += [r, r, r]
            cols += [tcol(int(p.rng_a[k]), c), tcol(int(p.rng_b[k]), c), Pn * blk + Ln * d + (k - k0) * d + c]
            vals += [1.0, -1.0, -p.rng_dist[k]]
            w.append(p.rng_w[k]), b.append(0.0)
            r += 1
    for q in range(p.prior_off[inst], p.prior_off[inst + 1]):
        for c in range(d):
            rows.append(r), cols.append(Pn * blk + int(p.prior_l[q]) * d + c), vals.append(1.0)
            w.append(p.prior_w[q]), b.append(p.prior_t[q][c])
            r += 1
    n_cols = Pn * blk + Ln * d + (p.rng_off[inst + 1] - k0) * d
    B = sp.csr_matrix((vals, (rows, cols)), shape=(r, n_cols))
    B.sort_indices()
    return B, np.asarray(w), np.asarray(b)


@pytest.mark.parametrize("name", ["goats", "man4", "man1", "mc0_small"])
def test_lowering_matches_oracle_layout(golden, name):
    fg, _ = golden(name)
    p = lower_factor_graph(fg, "QCQP")
    prob = so.assemble(fg, so.QCQP)
    B, w, b = csr_from_lowered(p)
    assert B.shape == prob.B.shape
    assert np.array_equal(B.indptr, prob.B.indptr) and np.array_equal(B.indices, prob.B.indices)
    assert np.array_equal(B.data, prob.B.data)
    assert np.array_equal(w, prob.w) and np.array_equal(b, prob.b)
    assert p.pose_names[0] == prob.pose_names and p.landmark_names[0] == prob.landmark_names
    assert p.range_keys[0] == prob.range_keys
    # chain segments tile the poses and follow pose_variables (one per robot here)
    assert p.seg_ptr[0] == 0 and p.seg_ptr[-1] == p.P
    assert p.n_seg == len(fg.pose_variables)
    assert (p.link_edge[p.seg_ptr[:-1]] == -1).all() and (np.delete(p.link_edge, p.seg_ptr[:-1]) >= 0).all()


def test_goats_pose_names_are_non_contiguous(golden):
    """names A0, A1, A10, ... — indices come from list order, never from the name."""
    fg, _ = golden("goats")
    p = lower_factor_graph(fg)
    assert p.P == 679 and p.L == 4 and p.E == 678 and p.K == 1558
    nums = [int(n[1:]) for n in p.pose_names[0]]
    assert nums != list(range(len(nums)))


def test_concat_offsets_and_blocks(golden):
    a = lower_factor_graph(golden("mc0_small")[0])
    b = lower_factor_graph(golden("man1")[0])
    c = concat([a, b, a])
    assert c.n_instances == 3
    assert c.pose_off.tolist() == [0, a.P, a.P + b.P, 2 * a.P + b.P]
    assert c.rng_off.tolist() == [0, a.K, a.K + b.K, 2 * a.K + b.K]
    assert c.seg_inst.tolist() == [0] * a.n_seg + [1] * b.n_seg + [2] * a.n_seg
    assert c.seg_ptr[-1] == c.P and np.all(np.diff(c.seg_ptr) > 0)
    # edge / range indices stay instance-local, link_edge is global
    assert np.array_equal(c.edge_i[a.E : a.E + b.E], b.edge_i)
    le = c.link_edge[a.P : a.P + b.P]
    assert np.array_equal(le[le >= 0], b.link_edge[b.link_edge >= 0] + a.E)
    for inst, src in enumerate([a, b, a]):
        B1, w1, _ = csr_from_lowered(c, inst)
        B0, w0, _ = csr_from_lowered(src, 0)
        assert (B1 != B0).nnz == 0 and np.array_equal(w0, w1)
    with pytest.raises(ValueError):
        concat([])
    with pytest.raises(ValueError):
        concat([a, lower_factor_graph(golden("mc0_small")[0], "SOCP")])


def test_array_lowering_equals_object_lowering():
    arr = generators.manhattan_2d_arrays(generators.MC_BASE_SEED + 5, n_robots=3, n_steps=12)
    fg = generators.arrays_to_factor_graph(arr)
    a = lower_manhattan_arrays(arr, "QCQP")
    b = lower_factor_graph(fg, "QCQP")
    for f in ("pose_off", "lm_off", "edge_off", "rng_off", "seg_ptr", "link_edge", "edge_i", "edge_j", "rng_a", "rng_b"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
    for f in ("edge_t", "edge_R", "edge_k", "edge_tau", "rng_dist", "rng_w"):
        assert np.allclose(getattr(a, f), getattr(b, f), rtol=0, atol=1e-15), f
    assert a.pose_names == b.pose_names and a.range_keys == b.range_keys


def test_generator_is_deterministic_and_connected():
    a = generators.manhattan_2d_arrays(123, n_robots=4, n_steps=20)
    b = generators.manhattan_2d_arrays(123, n_robots=4, n_steps=20)
    for k in a:
        assert np.array_equal(np.asarray(a[k]), np.asarray(b[k])), k
    fg = generators.monte_carlo_instance(3, n_robots=4, n_steps=20)
    assert len(fg.unconnected_variable_names) == 0
    assert fg.dimension == 2 and len(fg.pose_variables) == 4 and len(fg.landmark_variables) == 6
    assert all(m.dist >= 0 for m in fg.range_measurements)


def test_npz_roundtrip(tmp_path, golden):
    fg, _ = golden("mc0_small")
    path = os.path.join(tmp_path, "g.npz")
    save_graph_npz(fg, path, answer=np.arange(3))
    fg2, extra = load_graph_npz(path)
    a, b = graph_to_arrays(fg), graph_to_arrays(fg2)
    for k in a:
        assert np.array_equal(np.asarray(a[k]), np.asarray(b[k])), k
    assert extra["answer"].tolist() == [0, 1, 2]
    fg3 = arrays_to_graph(a)
    assert fg3.get_pose_chain_names() == fg.get_pose_chain_names()


def test_robot_subgraph_is_config2(golden):
    """robot-A sub-graph of the Manhattan pickle (400 poses, 399 odometry, 212 ranges)."""
    man4, _ = golden("man4")
    sub = robot_subgraph(man4, 0)
    assert len(sub.pose_variables) == 1 and len(sub.pose_variables[0]) == 400
    assert len(sub.odom_measurements[0]) == 399
    assert len(sub.range_measurements) == 212
    assert len(sub.unconnected_variable_names) == 0
    man1, _ = golden("man1")
    a, b = graph_to_arrays(sub), graph_to_arrays(man1)
    for k in a:
        assert np.array_equal(np.asarray(a[k]), np.asarray(b[k])), k


# ---- error behaviour of the entry point (no GPU needed: every check happens before the C ABI) ---------
def _tiny_graph(d=2):
    fg = FactorGraphData(d)
    fg.add_pose_variable(PoseVariable2D("A0", (0.0, 0.0), 0.0))
    fg.add_pose_variable(PoseVariable2D("A1", (1.0, 0.0), 0.0))
    fg.add_landmark_variable(LandmarkVariable2D("L0", (1.0, 1.0)))
    fg.add_odom_measurement(0, PoseMeasurement2D("A0", "A1", 1.0, 0.0, 0.0, 100.0, 1000.0))
    fg.add_range_measurement(FGRangeMeasurement(("A1", "L0"), 1.0, 0.5))
    return fg


def test_entry_point_errors():
    from score.solve_score import solve_score
    from score.utils.gurobi_utils import QCQP_RELAXATION, SOCP_RELAXATION

    assert (QCQP_RELAXATION, SOCP_RELAXATION) == ("QCQP", "SOCP")  # gurobi_utils.py:26-28
    fg = _tiny_graph()
    with pytest.raises(ValueError):  # _check_valid_relaxation, gurobi_utils.py:139-144
        solve_score(fg, "LP")
    with pytest.raises(ValueError):  # the example's 3-positional form: (data, solver_params, relaxation)
        solve_score(fg, object(), "LP")
    fg.add_landmark_variable(LandmarkVariable2D("L1", (5.0, 5.0)))
    with pytest.raises(AssertionError):  # _check_factor_graph, solve_score.py:28-32
        solve_score(fg)


def test_lowering_errors():
    fg = _tiny_graph()
    fg.add_range_measurement(FGRangeMeasurement(("A1", "L0"), 2.0, 0.5))
    with pytest.raises(ValueError):  # duplicate distance variable name (QCQP only), gurobi_utils.py:62-67,296-306
        lower_factor_graph(fg, "QCQP")
    assert lower_factor_graph(fg, "SOCP").K == 2
    fg = _tiny_graph()
    fg.add_range_measurement(FGRangeMeasurement(("A0", "nope"), 2.0, 0.5))
    with pytest.raises(ValueError):  # get_translation_var, gurobi_utils.py:103-109
        lower_factor_graph(fg)
    fg = _tiny_graph()
    fg.dimension = 4
    with pytest.raises(ValueError):  # is_dimension, gurobi_utils.py:37-50
        lower_factor_graph(fg)
    empty = FactorGraphData(2)
    with pytest.raises(IndexError):  # fg.pose_variables[0][0], gurobi_utils.py:181
        lower_factor_graph(empty)
    fg = _tiny_graph()
    fg.add_landmark_prior(LandmarkPrior2D("L0", (1.0, 2.0), 4.0))
    p = lower_factor_graph(fg)
    assert p.Lp == 1 and p.prior_t.tolist() == [[1.0, 2.0]] and p.prior_w.tolist() == [4.0]


def test_product_path_has_no_cpu_fallback(built_lib):
    """Without a CUDA device the solve raises (SCORE_ERR_CUDA); nothing routes through the oracle."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from score.solve_score import solve_score

    with pytest.raises(RuntimeError):
        solve_score(_tiny_graph())
    import score_b200.solve_score as mod
    import score_b200.solver as smod

    for m in (mod, smod):
        src = open(m.__file__).read()
        assert "import oracle" not in src and "from oracle" not in src


# ---- shim semantics --------------------------------------------------------------------
def test_shim_measurement_semantics():
    m = FGRangeMeasurement(("A3", "L1"), 7.5, 0.75)
    assert (m.first_key, m.second_key) == ("A3", "L1")
    assert abs(m.precision - 1 / 0.75**2) < 1e-15  # GOATS min_measure_weight 1.7778
    pm = PoseMeasurement2D("A0", "A1", 1.0, 2.0, 0.3, 10.0, 20.0)
    assert pm.translation_vector.tolist() == [1.0, 2.0]
    assert np.allclose(pm.rotation_matrix, [[np.cos(0.3), -np.sin(0.3)], [np.sin(0.3), np.cos(0.3)]])
    R = np.array([[0.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]])
    pm3 = PoseMeasurement3D("A0", "A1", np.array([1.0, 2.0, 3.0]), R, 10.0, 20.0)
    assert pm3.translation_vector.tolist() == [1.0, 2.0, 3.0] and np.array_equal(pm3.rotation_matrix, R)


def test_shim_results_layout_and_tum(tmp_path):
    T = np.eye(3)
    T[:2, 2] = [1.0, 2.0]
    vv = VariableValues(2, {"A0": np.eye(3), "A1": T}, {"L0": np.array([3.0, 4.0])}, {("A1", "L0"): np.array([0.6, 0.8])})
    res = SolverResults(variables=vv, total_time=0.5, solved=True, pose_chain_names=[["A0", "A1"]])
    assert res.poses["A1"][1, 2] == 2.0 and res.translations["A1"].tolist() == [1.0, 2.0]
    assert res.landmarks["L0"].tolist() == [3.0, 4.0] and res.distances[("A1", "L0")].tolist() == [0.6, 0.8]
    assert abs(res.rotations_theta["A1"]) < 1e-15
    files = save_to_tum(res, os.path.join(tmp_path, "traj.tum"))
    lines = open(files[0]).read().strip().splitlines()
    assert len(lines) == 2
    cols = lines[1].split()
    assert len(cols) == 8 and float(cols[1]) == 1.0 and float(cols[2]) == 2.0 and float(cols[7]) == 1.0


def test_shim_rounding_matches_reference_rule():
    rng = np.random.default_rng(0)
    for d in (2, 3):
        for _ in range(20):
            M = rng.standard_normal((d, d))
            R = round_to_special_orthogonal(M)
            assert np.allclose(R @ R.T, np.eye(d), atol=1e-12) and abs(np.linalg.det(R) - 1) < 1e-12
    with pytest.raises(AssertionError):
        round_to_special_orthogonal(np.zeros((2, 3)))


def test_shim_pickle_roundtrip(golden):
    """FactorGraphData pickles carry py_factor_graph.* class paths."""
    fg, _ = golden("mc0_small")
    blob = pickle.dumps(fg)
    assert b"py_factor_graph" in blob
    fg2 = pickle.loads(blob)
    a, b = graph_to_arrays(fg), graph_to_arrays(fg2)
    for k in a:
        assert np.array_equal(np.asarray(a[k]), np.asarray(b[k])), k


def test_grid3d_generator_and_lowering():
    """SURVEY 8(d) config 5 at test size: unit lattice steps, exact range count, distinct keys, and the direct
    array lowering equals lowering the FactorGraphData objects."""
    from score_b200.lowering import lower_grid3d_arrays

    kw = dict(n_robots=3, n_steps=25, grid=6, n_landmarks=4, n_ranges=120)
    arr = generators.grid_3d_arrays(5, **kw)
    arr2 = generators.grid_3d_arrays(5, **kw)
    for k in arr:
        assert np.array_equal(np.asarray(arr[k]), np.asarray(arr2[k])), k
    assert len(arr["rng_a"]) == 120
    assert len({(a, b) for a, b in zip(arr["rng_a"], arr["rng_b"])}) == 120
    steps = np.linalg.norm(np.diff(arr["pos"], axis=1), axis=-1)
    assert np.allclose(steps, 1.0) and arr["pos"].min() >= 0 and arr["pos"].max() <= 6
    Rm = arr["rot"]
    assert np.abs(Rm @ np.transpose(Rm, (0, 1, 3, 2)) - np.eye(3)).max() < 1e-12 and np.allclose(np.linalg.det(Rm), 1)
    oR = arr["odom_R"]
    assert np.abs(oR @ np.transpose(oR, (0, 1, 3, 2)) - np.eye(3)).max() < 1e-12
    # inter-robot ranges connect poses of the same timestep
    rr = arr["rng_b"] < 75
    assert np.array_equal(arr["rng_a"][rr] % 25, arr["rng_b"][rr] % 25)
    fg = generators.grid_3d_factor_graph(arr)
    assert fg.dimension == 3 and len(fg.unconnected_variable_names) == 0
    a, b = lower_grid3d_arrays(arr, with_names=True), lower_factor_graph(fg)
    for f in ("pose_off", "lm_off", "edge_off", "rng_off", "seg_ptr", "link_edge", "edge_i", "edge_j", "rng_a", "rng_b",
              "edge_t", "edge_R", "edge_k", "edge_tau", "rng_dist", "rng_w"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
    assert a.range_keys == b.range_keys
    B, w, _ = csr_from_lowered(a)
    prob = so.assemble(fg, so.QCQP)
    assert np.array_equal(B.indptr, prob.B.indptr) and np.array_equal(B.indices, prob.B.indices)
    assert np.array_equal(B.data, prob.B.data) and np.array_equal(w, prob.w)
    with pytest.raises(ValueError):
        generators.grid_3d_arrays(5, n_robots=2, n_steps=3, grid=4, n_landmarks=1, n_ranges=10**6)


def test_slice_instances_inverts_concat():
    import dataclasses

    from score_b200.lowering import slice_instances

    ps = [lower_manhattan_arrays(generators.manhattan_2d_arrays(100 + i, n_robots=2 + i % 3, n_steps=8 + i), "QCQP")
          for i in range(5)]
    c = concat(ps)
    for i0, i1 in [(0, 5), (1, 3), (4, 5), (0, 2)]:
        a, b = slice_instances(c, i0, i1), concat(ps[i0:i1])
        for f in dataclasses.fields(a):
            x, y = getattr(a, f.name), getattr(b, f.name)
            if isinstance(x, np.ndarray):
                assert np.array_equal(x, y) and x.dtype == y.dtype, f.name
            else:
                assert x == y, f.name
    with pytest.raises(ValueError):
        slice_instances(c, 3, 3)


def test_solver_group_partitions_a_batch_into_a_queue_of_sub_batches():
    """ScoreSolverGroup(n_streams, n_parts): contiguous sub-batches that tile the batch; the host-thread count is
    clamped to the number of parts, the part count to the number of instances (no device needed: create=False)."""
    from score_b200 import generators
    from score_b200.lowering import concat, lower_manhattan_arrays
    from score_b200.solver import ScoreSolverGroup

    probs = [lower_manhattan_arrays(generators.manhattan_2d_arrays(generators.MC_BASE_SEED + i, n_robots=2, n_steps=6))
             for i in range(7)]
    batch = concat(probs)
    g = ScoreSolverGroup(batch, n_streams=2, n_parts=4, create=False)
    assert g.n_streams == 2 and len(g.parts) == 4 and g.cuts[0] == 0 and g.cuts[-1] == 7
    assert sum(p.n_instances for p in g.parts) == 7 and all(p.n_instances >= 1 for p in g.parts)
    assert sum(p.P for p in g.parts) == batch.P and sum(p.K for p in g.parts) == batch.K
    for j, p in enumerate(g.parts):  # part j holds instances cuts[j] .. cuts[j+1] unchanged
        a = g.cuts[j]
        assert np.array_equal(p.rng_dist, batch.rng_dist[batch.rng_off[a]:batch.rng_off[g.cuts[j + 1]]])
    g.close()
    g = ScoreSolverGroup(batch, n_streams=3, create=False)  # default: one part per stream
    assert g.n_streams == 3 and len(g.parts) == 3
    g.close()
    g = ScoreSolverGroup(batch, n_streams=16, n_parts=64, create=False)  # clamped to the instance count
    assert len(g.parts) == 7 and g.n_streams == 7
    g.close()


def test_streamed_pipeline_queue_semantics(monkeypatch):
    """ScoreSolverGroup.run_pipelined(steps=K) without a device (ScoreSolver replaced by a recorder): K x parts jobs,
    one score_create at a time, at most n_streams solves at a time, one more job may already be created while they
    run, byte counts reported per step, every job read back into its own part's view."""
    import threading
    import time

    from score_b200 import generators, solver as solver_mod
    from score_b200.lowering import concat, lower_manhattan_arrays

    probs = [lower_manhattan_arrays(generators.manhattan_2d_arrays(generators.MC_BASE_SEED + i, n_robots=2, n_steps=6))
             for i in range(4)]
    batch = concat(probs)
    lock = threading.Lock()
    state = {"creating": 0, "max_creating": 0, "solving": 0, "max_solving": 0, "alive": 0, "max_alive": 0, "jobs": 0,
             "readbacks": []}

    class FakeSolver:
        def __init__(self, prob, device=0):
            with lock:
                state["creating"] += 1
                state["max_creating"] = max(state["max_creating"], state["creating"])
                state["alive"] += 1
                state["max_alive"] = max(state["max_alive"], state["alive"])
            time.sleep(0.01)
            self.prob, self.h2d_bytes, self.d2h_bytes = prob, 100 * prob.n_instances, 0
            with lock:
                state["creating"] -= 1

        def __enter__(self):
            return self

        def __exit__(self, *exc):
            self.close()

        def close(self):
            with lock:
                state["alive"] -= 1

        def solve(self, **kw):
            with lock:
                state["solving"] += 1
                state["max_solving"] = max(state["max_solving"], state["solving"])
                state["jobs"] += 1
            time.sleep(0.03)
            with lock:
                state["solving"] -= 1
            n = self.prob.n_instances
            inst = np.zeros(n, dtype=solver_mod._INST_DTYPE)
            inst["solved"] = 1
            z = np.zeros(len(solver_mod.KERNEL_NAMES))
            return solver_mod.SolveStats(n, n, 5, 7, 0.0, 0.0, 1.0, 0.0, 1.0, 0, 0, 0, 0.0, inst, kernel_ms=z,
                                         kernel_bytes=z, kernel_count=z, kernel_bytes_total=z, cycles=1)

        def solution(self, out=None):
            out[0][...] = float(self.prob.n_instances)  # mark the view of this part
            self.d2h_bytes = 10 * self.prob.n_instances
            with lock:
                state["readbacks"].append(out[0].shape[0])
            return out

    monkeypatch.setattr(solver_mod, "ScoreSolver", FakeSolver)
    g = solver_mod.ScoreSolverGroup(batch, n_streams=2, n_parts=2, create=False)
    stats, out, h2d, d2h = g.run_pipelined(steps=3)
    g.close()
    assert state["jobs"] == 6 and stats.n_instances == 12 and stats.n_solved == 12  # 3 steps x 2 parts x 2 instances
    assert state["max_creating"] == 1
    assert state["max_solving"] <= 2
    assert state["max_alive"] <= 3  # two solving + one created ahead
    assert (h2d, d2h) == (100 * 4, 10 * 4)  # per step
    assert np.all(out[0] == 2.0)  # every pose row was written by a 2-instance part
    assert sorted(state["readbacks"]) == sorted([g.parts[0].P, g.parts[1].P] * 3)


def test_pipelined_steps_schedule(monkeypatch):
    """ScoreSolverGroup.solve_steps without a device: K steps double-buffered over n_sets handle sets — every (set, part)
    handle is re-solved by its own thread, sets x parts solves in flight at most, one merged record per step, and a
    handle is never solved by two threads at once."""
    import threading
    import time

    from score_b200 import generators, solver as solver_mod
    from score_b200.lowering import concat, lower_manhattan_arrays

    probs = [lower_manhattan_arrays(generators.manhattan_2d_arrays(generators.MC_BASE_SEED + i, n_robots=2, n_steps=6))
             for i in range(4)]
    batch = concat(probs)
    lock = threading.Lock()
    state = {"solving": 0, "max_solving": 0, "handles": 0, "overlap": False}

    class FakeSolver:
        def __init__(self, prob, device=0):
            self.prob, self.busy, self.count = prob, False, 0
            with lock:
                state["handles"] += 1

        def close(self):
            pass

        def solve(self, **kw):
            with lock:
                state["overlap"] |= self.busy
                self.busy = True
                state["solving"] += 1
                state["max_solving"] = max(state["max_solving"], state["solving"])
            time.sleep(0.02)
            with lock:
                self.busy = False
                state["solving"] -= 1
                self.count += 1
            n = self.prob.n_instances
            inst = np.zeros(n, dtype=solver_mod._INST_DTYPE)
            inst["solved"] = 1
            z = np.zeros(len(solver_mod.KERNEL_NAMES))
            return solver_mod.SolveStats(n, n, 5, 7, 0.0, 0.0, 1.0, 0.0, 1.0, 0, 0, 0, 0.0, inst, kernel_ms=z,
                                         kernel_bytes=z, kernel_count=z, kernel_bytes_total=z, cycles=1)

    monkeypatch.setattr(solver_mod, "ScoreSolver", FakeSolver)
    g = solver_mod.ScoreSolverGroup(batch, n_streams=2)
    steps = g.solve_steps(5, n_sets=2)
    assert len(steps) == 5 and all(s.n_instances == 4 and s.n_solved == 4 for s in steps)
    assert state["handles"] == 4 and not state["overlap"]  # 2 sets x 2 parts, each handle used by one thread at a time
    assert 2 <= state["max_solving"] <= 4
    counts = sorted(s.count for st in g._sets for s in st)
    assert counts == [2, 2, 3, 3]  # set 0 serves steps 0, 2, 4; set 1 serves steps 1, 3
    g.close()


def test_g2o_roundtrip(tmp_path):
    """g2o text (README.md:53-56 names it as the interchange format PyFactorGraph reads) -> FactorGraphData: what the
    SCORE cost reads survives a write / parse round trip bit for bit (2D multi-robot graph with ranges, 3D graph)."""
    from py_factor_graph.parsing.parse_g2o_file import parse_g2o_file, write_g2o_file

    fg2 = generators.manhattan_2d(generators.MC_BASE_SEED + 3, n_robots=3, n_steps=12)
    fg3 = generators.grid_3d_factor_graph(generators.grid_3d_arrays(5, n_robots=2, n_steps=10, grid=6, n_landmarks=3, n_ranges=40))
    for k, fg in enumerate((fg2, fg3)):
        path = os.path.join(tmp_path, f"g{k}.g2o")
        write_g2o_file(fg, path)
        back = parse_g2o_file(path)
        a, b = lower_factor_graph(fg), lower_factor_graph(back)
        assert back.get_pose_chain_names() == fg.get_pose_chain_names()
        for name in ("edge_i", "edge_j", "rng_a", "rng_b", "link_edge", "seg_ptr"):
            assert np.array_equal(getattr(a, name), getattr(b, name)), name
        for name in ("edge_k", "edge_tau", "rng_dist"):
            assert np.array_equal(getattr(a, name), getattr(b, name)), name
        assert np.allclose(a.edge_t, b.edge_t, rtol=0, atol=0) and np.allclose(a.rng_w, b.rng_w, rtol=1e-15)
        assert np.allclose(a.edge_R, b.edge_R, rtol=0, atol=1e-15)  # 3D: through a unit quaternion
    # plain integer ids and a standard landmark-free 2D file
    path = os.path.join(tmp_path, "plain.g2o")
    with open(path, "w") as f:
        f.write("VERTEX_SE2 0 0 0 0\nVERTEX_SE2 1 1 0 0\nVERTEX_SE2 2 2 0 0.1\n")
        f.write("EDGE_SE2 0 1 1 0 0 100 0 0 100 0 400\nEDGE_SE2 1 2 1 0 0.1 100 0 0 100 0 400\nEDGE_SE2 0 2 2 0 0.1 50 0 0 50 0 200\n")
    fg = parse_g2o_file(path)
    assert [p.name for p in fg.pose_variables[0]] == ["A0", "A1", "A2"]
    assert len(fg.odom_measurements[0]) == 2 and len(fg.loop_closure_measurements) == 1
    assert fg.loop_closure_measurements[0].translation_precision == 50 and fg.loop_closure_measurements[0].rotation_precision == 200
    with open(path, "a") as f:
        f.write("EDGE_BEARING 0 1 0.3 10\n")
    with pytest.raises(ValueError):
        parse_g2o_file(path)


def test_dynamic_queue_jobs_without_a_device(monkeypatch):
    """sharding.SweepQueue over solver.HandlePool / solver.StreamedJobs with ScoreSolver replaced by a recorder: a pooled
    handle is never used by two jobs at once (a third borrower of a sub-batch waits), streamed jobs create one handle at
    a time and cap the solves in flight, every (step, part) job runs once, byte counts add up."""
    import threading
    import time

    from score_b200 import generators, solver as solver_mod
    from score_b200.lowering import concat, lower_manhattan_arrays, slice_instances
    from score_b200.sharding import SweepQueue

    probs = [lower_manhattan_arrays(generators.manhattan_2d_arrays(generators.MC_BASE_SEED + i, n_robots=2, n_steps=6))
             for i in range(6)]
    batch = concat(probs)
    parts = [slice_instances(batch, a, b) for a, b in ((0, 2), (2, 3), (3, 6))]
    lock = threading.Lock()
    st = {"busy": {}, "overlap": 0, "creating": 0, "max_creating": 0, "solving": 0, "max_solving": 0, "made": 0}

    class FakeSolver:
        def __init__(self, prob, device=0):
            with lock:
                st["creating"] += 1
                st["max_creating"] = max(st["max_creating"], st["creating"])
                st["made"] += 1
            time.sleep(0.005)
            self.prob, self.h2d_bytes, self.d2h_bytes = prob, 100 * prob.n_instances, 0
            with lock:
                st["creating"] -= 1

        def __enter__(self):
            return self

        def __exit__(self, *exc):
            self.close()

        def close(self):
            pass

        def solve(self, **kw):
            with lock:
                st["overlap"] += 1 if st["busy"].get(id(self)) else 0
                st["busy"][id(self)] = True
                st["solving"] += 1
                st["max_solving"] = max(st["max_solving"], st["solving"])
            time.sleep(0.02)
            with lock:
                st["busy"][id(self)] = False
                st["solving"] -= 1
            n = self.prob.n_instances
            inst = np.zeros(n, dtype=solver_mod._INST_DTYPE)
            z = np.zeros(len(solver_mod.KERNEL_NAMES))
            return solver_mod.SolveStats(n, n, 5, 7, 0.0, 0.0, 1.0, 0.0, 1.0, 0, 0, 0, 0.0, inst, kernel_ms=z,
                                         kernel_bytes=z, kernel_count=z, kernel_bytes_total=z, cycles=3 * n)

        def solution(self, out=None):
            for a, shp in zip(out, self.solution_shapes()):
                assert a.shape == shp
            self.d2h_bytes = 10 * self.prob.n_instances
            return out

        def solution_shapes(self):
            p, d = self.prob, self.prob.dim
            return (p.P, d, d + 1), (p.P, d, d), (p.L, d), (p.K, p.dist_per)

    monkeypatch.setattr(solver_mod, "ScoreSolver", FakeSolver)
    # device-resident: 2 handles per sub-batch, 5 workers -> more borrowers than handles for the hot sub-batch
    with solver_mod.HandlePool(parts, copies=2) as pool:
        assert st["made"] == 6
        costs = [float(w.cycles) for w in pool.warm()]
        assert costs == [6.0, 3.0, 9.0]
        with SweepQueue(3, lambda step, part, w: pool.solve(part), inflight=5) as q:
            res = q.run(4, costs)
    assert sorted((s, p) for s, p, _ in res) == [(s, p) for s in range(4) for p in range(3)]
    assert all(r.n_instances == parts[p].n_instances for _, p, r in res)
    assert st["overlap"] == 0  # no handle ran two solves at once
    # host buffers: one create at a time, at most 2 solves in flight, per-worker output slots
    d = batch.dim
    slot = lambda: (np.empty((batch.P, d, d + 1)), np.empty((batch.P, d, d)), np.empty((batch.L, d)),
                    np.empty((batch.K, batch.dist_per)))
    st.update(max_creating=0, max_solving=0, made=0)
    jobs = solver_mod.StreamedJobs(parts, [slot() for _ in range(3)], inflight=2)
    with SweepQueue(3, jobs, inflight=3) as q:
        res = q.run(2, costs)
    assert len(res) == 6 and st["made"] == 6 and st["max_creating"] == 1 and st["max_solving"] <= 2
    assert jobs.h2d_bytes == 2 * 100 * 6 and jobs.d2h_bytes == 2 * 10 * 6
