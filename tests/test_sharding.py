"""Multi-rank host logic on CPU: instance sharding with a world_size-2 gloo process group.

The GPU solve is replaced by the CPU oracle through the `solve_fn` hook of solve_score_sharded (tests may
execute the oracle); what is under test is the partition, the absence of any data-path collective,
and the ordered gather of the per-instance results."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_is_balanced_and_deterministic():
    from score_b200.sharding import contiguous_shard, partition_instances

    rng = np.random.default_rng(0)
    costs = rng.integers(50_000, 70_000, size=1024).astype(float)
    for world in (1, 2, 4, 8):
        parts = partition_instances(costs, world)
        allids = np.concatenate(parts)
        assert sorted(allids.tolist()) == list(range(1024))
        loads = np.array([costs[p].sum() for p in parts])
        assert loads.max() / loads.mean() < 1.001
        assert all(len(p) == 1024 // world for p in parts)
        again = partition_instances(costs, world)
        assert all(np.array_equal(a, b) for a, b in zip(parts, again))
    assert partition_instances([], 2)[0].size == 0
    assert [len(p) for p in partition_instances([5, 1, 1], 4)] == [1, 1, 1, 0]
    assert list(contiguous_shard(4, 1, 2)) == [4, 5, 6, 7]
    with pytest.raises(ValueError):
        partition_instances([1.0], 0)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_solve(graphs):
    from oracle import score_oracle as so

    out = []
    for fg in graphs:
        prob, x, sol = so.solve(fg, so.QCQP, mu_final=1e-9)
        out.append({"f": so.objective(prob, x), "first_pose": fg.pose_variables[0][0].name, "n": prob.n_cols,
                    "pid": os.getpid()})
    return out


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from score_b200 import generators
    from score_b200.sharding import solve_score_sharded

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        # ragged sweep: instance sizes differ, so the cost-balanced partition is not the round-robin one
        graphs = [generators.manhattan_2d(generators.MC_BASE_SEED + i, n_robots=2 + (i % 3), n_steps=8 + 2 * i)
                  for i in range(5)]
        res = solve_score_sharded(graphs, "QCQP", solve_fn=_oracle_solve, dst=0)
        everywhere = solve_score_sharded(graphs, "QCQP", solve_fn=_oracle_solve, dst=None)
        q.put((rank, res, everywhere))
    finally:
        dist.destroy_process_group()


def test_sharded_solve_gloo_world2():
    import torch.multiprocessing as mp

    from score_b200 import generators

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(2):
        rank, res, everywhere = q.get(timeout=240)
        got[rank] = (res, everywhere)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res0, all0 = got[0]
    res1, all1 = got[1]
    assert res1 is None and res0 is not None  # gathered on rank 0 only
    graphs = [generators.manhattan_2d(generators.MC_BASE_SEED + i, n_robots=2 + (i % 3), n_steps=8 + 2 * i)
              for i in range(5)]
    single = _oracle_solve(graphs)
    assert len(res0) == 5
    for a, b in zip(res0, single):  # same instance order, same numbers as a single-process run
        assert a["n"] == b["n"] and a["first_pose"] == b["first_pose"]
        assert abs(a["f"] - b["f"]) <= 1e-9 * max(1.0, abs(b["f"]))
    assert len({r["pid"] for r in res0}) == 2  # both ranks did work
    for a, b, c in zip(all0, all1, res0):  # dst=None: every rank holds the full ordered list
        assert a["f"] == b["f"] == c["f"]


def test_gather_detects_double_assignment():
    from score_b200.sharding import gather_results

    assert gather_results([1, 0], ["b", "a"], 2) == ["a", "b"]
    with pytest.raises(RuntimeError):
        gather_results([0, 0], ["a", "b"], 2)
    with pytest.raises(RuntimeError):
        gather_results([0], ["a"], 2)


def test_sweep_queue_single_process_order_and_coverage():
    """One process: every (step, part) job runs exactly once, longest sub-batch first inside a step."""
    import threading

    from score_b200.sharding import SweepQueue

    claimed, lock = [], threading.Lock()

    def job(step, part, worker):
        with lock:
            claimed.append((step, part))
        return part * 10 + step

    costs = [3.0, 9.0, 1.0, 9.0]
    assert SweepQueue.job_order(costs, 4) == [1, 3, 0, 2]
    assert SweepQueue.job_order(None, 3) == [0, 1, 2]
    with SweepQueue(4, job, inflight=1) as q:  # one worker: the claim order is the queue order
        res = q.run(3, costs)
        assert claimed == [(s, p) for s in range(3) for p in (1, 3, 0, 2)]
        assert [(s, p) for s, p, _ in res] == claimed and all(r == p * 10 + s for s, p, r in res)
        claimed.clear()
        res = q.run(2)  # a second run starts a fresh counter
        assert sorted(claimed) == [(s, p) for s in range(2) for p in range(4)]
    with SweepQueue(5, job, inflight=3) as q:
        claimed.clear()
        res = q.run(4, [1, 2, 3, 4, 5])
        assert sorted((s, p) for s, p, _ in res) == [(s, p) for s in range(4) for p in range(5)]
    with pytest.raises(ValueError):
        SweepQueue(0, job)
    with pytest.raises(ValueError):
        SweepQueue.job_order([1.0], 2)


def _queue_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import time

    import torch.distributed as dist

    from score_b200.sharding import JobCounter, SweepQueue

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        store = dist.distributed_c10d._get_default_store()

        def job(step, part, worker):  # rank 1 is the slow GPU: every job takes 4x longer there
            time.sleep(0.02 if rank == 0 else 0.08)
            return (rank, step, part)

        runs = []
        with SweepQueue(6, job, JobCounter(store, key="test/q"), inflight=2) as q:
            for _ in range(2):
                dist.barrier()
                runs.append(q.run(5, [1, 6, 2, 5, 3, 4]))
            dist.barrier()  # the store lives in rank 0: it must not go away while the slow rank still claims jobs
        out.put((rank, runs))
    finally:
        dist.destroy_process_group()


def test_sweep_queue_gloo_world2():
    """Two ranks drain one queue through the process group's store: every job exactly once per run, the slow rank
    claims fewer."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_queue_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(out.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for run in range(2):
        jobs = [(s, p) for r in (0, 1) for s, p, _ in got[r][run]]
        assert sorted(jobs) == [(s, p) for s in range(5) for p in range(6)]  # each job exactly once, none lost
        assert all(res == (r, s, p) for r in (0, 1) for s, p, res in got[r][run])
        n0, n1 = len(got[0][run]), len(got[1][run])
        assert n0 > n1 >= 1  # the fast rank took more of the queue
