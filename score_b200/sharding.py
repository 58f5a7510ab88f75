"""Multi-GPU sharding of independent instances (Monte-Carlo sweep).

One process per GPU (``torch.distributed``; NCCL on the B200 box, gloo in the CPU tests).  The
instances of a sweep are independent convex programs, so they are split across ranks with NO data-path
collective: every rank lowers, uploads and solves its own shard through the C ABI, and only the
per-instance result records travel (one gather at the end).  The reference has no counterpart — it is
single-process — this is the data-parallel axis of ``solve_score``.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import numpy as np


def instance_cost(prob) -> float:
    """Work estimate of one lowered instance: non-zeros of its reduced operator
    (edges: d(d+2) + d^2(d+1) each, ranges: 2d each, priors: d each)."""
    d = prob.dim
    return float(prob.E * (d * (d + 2) + d * d * (d + 1)) + prob.K * 2 * d + prob.Lp * d)


def partition_instances(costs: Sequence[float], world_size: int) -> List[np.ndarray]:
    """Deterministic longest-processing-time partition: instance ids per rank, ascending inside a rank.

    Every rank computes the same partition from the same costs (no communication).  Ties are broken by
    instance id so the result does not depend on the sort implementation.
    """
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    costs = np.asarray(costs, dtype=np.float64)
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0.0] * world_size
    count = [0] * world_size
    shards: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda j: (load[j], count[j], j))
        shards[r].append(i)
        load[r] += float(costs[i])
        count[r] += 1
    return [np.asarray(sorted(s), dtype=np.int64) for s in shards]


def contiguous_shard(n: int, rank: int, world_size: int) -> range:
    """Block partition used by bench.py's weak-scaling sweep (rank r owns instances [r*n, (r+1)*n))."""
    return range(rank * n, (rank + 1) * n)


def gather_results(local_ids: Sequence[int], local_results: Sequence[object], n_total: int, group=None,
                   dst: Optional[int] = 0) -> Optional[List[object]]:
    """Collect per-instance results from every rank, in instance order.

    ``dst=None`` returns the full list on every rank, otherwise only on rank ``dst`` (None elsewhere).
    Works without an initialised process group (single process).
    """
    import torch.distributed as dist

    payload = (list(map(int, local_ids)), list(local_results))
    if not (dist.is_available() and dist.is_initialized()):
        parts = [payload]
    else:
        world = dist.get_world_size(group)
        if dst is None:
            parts = [None] * world
            dist.all_gather_object(parts, payload, group=group)
        else:
            parts = [None] * world if dist.get_rank(group) == dst else None
            dist.gather_object(payload, parts, dst=dst, group=group)
            if parts is None:
                return None
    out: List[object] = [None] * n_total
    seen = 0
    for ids, res in parts:
        for i, r in zip(ids, res):
            if out[i] is not None:
                raise RuntimeError(f"instance {i} was solved by two ranks")
            out[i] = r
            seen += 1
    if seen != n_total:
        raise RuntimeError(f"gathered {seen} results for {n_total} instances")
    return out


def solve_score_sharded(datas, relaxation_type: str = "QCQP", device: Optional[int] = None, group=None,
                        dst: Optional[int] = 0, solve_fn: Optional[Callable] = None, **solver_kw):
    """Solve a list of independent ``FactorGraphData`` instances across the ranks of a process group.

    Every rank is handed the same ``datas``; it solves the instances ``partition_instances`` assigns
    to it as one block-diagonal batch on its GPU (``solve_score_batch``) and the ``SolverResults`` are
    gathered in input order.  ``solve_fn(list_of_graphs) -> list_of_results`` replaces the GPU batch
    solve in the CPU (gloo) tests.
    """
    import torch.distributed as dist

    from .lowering import check_valid_relaxation, lower_factor_graph

    check_valid_relaxation(relaxation_type)
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    costs = [instance_cost(lower_factor_graph(fg, relaxation_type)) for fg in datas]
    mine = partition_instances(costs, world)[rank]
    local: List[object] = []
    if len(mine):
        shard = [datas[int(i)] for i in mine]
        if solve_fn is None:
            from .solve_score import solve_score_batch

            if device is None:
                import torch

                device = rank % max(1, torch.cuda.device_count())
            local = solve_score_batch(shard, relaxation_type, device=device, **solver_kw)
        else:
            local = list(solve_fn(shard))
        if len(local) != len(mine):
            raise RuntimeError("solve_fn returned a different number of results than instances")
    return gather_results(mine, local, len(datas), group=group, dst=dst)


class JobCounter:
    """Fetch-and-add counter shared by all ranks: the c10d key-value store's atomic ``add`` (TCPStore / the default
    store of the process group), or a lock-protected integer inside one process (``store=None``)."""

    def __init__(self, store=None, key: str = "score_b200/sweep_queue"):
        import threading

        self.store, self.key = store, key
        self._lock, self._local = threading.Lock(), {}

    def next(self, epoch: int) -> int:
        """The next unclaimed job number of run ``epoch`` (0, 1, 2, ... across all ranks and threads)."""
        if self.store is not None:
            return int(self.store.add(f"{self.key}/{epoch}", 1)) - 1
        with self._lock:
            v = self._local.get(epoch, 0)
            self._local[epoch] = v + 1
            return v


class SweepQueue:
    """Dynamic schedule of a Monte-Carlo sweep over the ranks ("the only loss is tail imbalance
    from per-instance iteration counts => dynamic chunking").

    The sweep is cut into ``n_parts`` sub-batches that every rank can solve (the lowered inputs are small: ~0.2 MB per
    instance, so each rank keeps all of them; on the device-resident path as ready handles).  ``steps`` passes over the
    sweep are ONE queue of ``steps x n_parts`` jobs; ``inflight`` host threads per rank claim the next job from a
    counter shared by all ranks (``JobCounter``) and run it on their GPU, so that a rank that drew an ill-conditioned
    sub-batch (its slowest instance needs 2-3x the cycles of a typical one) simply claims fewer jobs instead of holding
    every step back.  Inside a pass the sub-batches are handed out longest first (``costs``, e.g. the solve times of a
    warm-up pass), which keeps the last jobs of the queue short.  There is still no data-path collective: the counter is
    a host-side store operation per job (16 per 8192-instance pass).

    ``job_fn(step, part, worker) -> result`` runs one job (bench.py / ``solve_parts``: a ``score_solve`` on a pooled
    handle, or create -> solve -> read back -> destroy from host buffers).  Results are bit-identical to a static
    schedule: an instance's arithmetic does not depend on where or next to what it is solved.
    """

    def __init__(self, n_parts: int, job_fn: Callable, counter: Optional[JobCounter] = None, inflight: int = 4):
        if n_parts < 1 or inflight < 1:
            raise ValueError("n_parts and inflight must be >= 1")
        self.n_parts, self.job_fn, self.inflight = int(n_parts), job_fn, int(inflight)
        self.counter = counter if counter is not None else JobCounter()
        self.epoch = 0
        self._pool = None

    def close(self) -> None:
        if self._pool is not None:
            self._pool.shutdown(wait=True)
            self._pool = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @staticmethod
    def job_order(costs: Optional[Sequence[float]], n_parts: int) -> List[int]:
        """Sub-batch order inside a pass: descending cost, ties by index (the same on every rank)."""
        if costs is None:
            return list(range(n_parts))
        if len(costs) != n_parts:
            raise ValueError("one cost per sub-batch")
        return sorted(range(n_parts), key=lambda i: (-float(costs[i]), i))

    def run(self, steps: int, costs: Optional[Sequence[float]] = None) -> List[tuple]:
        """Work through this run's queue together with the other ranks (every rank must call ``run`` the same number
        of times with the same ``steps`` / ``costs``).  Returns the jobs THIS rank ran as (step, part, result), in
        completion order; returns when the queue is empty and this rank's last job has finished.  A rank that returns
        early must keep the counter's store alive until the others are done: put a barrier between the last ``run``
        and ``destroy_process_group`` (the store is served by rank 0)."""
        from concurrent.futures import ThreadPoolExecutor

        order = self.job_order(costs, self.n_parts)
        total = int(steps) * self.n_parts
        epoch = self.epoch
        self.epoch += 1
        if self._pool is None:
            self._pool = ThreadPoolExecutor(self.inflight)

        def worker(w):
            done = []
            while True:
                j = self.counter.next(epoch)
                if j >= total:
                    return done
                step, part = j // self.n_parts, order[j % self.n_parts]
                done.append((step, part, self.job_fn(step, part, w)))

        out: List[tuple] = []
        for d in self._pool.map(worker, range(self.inflight)):
            out.extend(d)
        return out


def row_block_range(n_blocks: int, rank: int, world_size: int) -> range:
    """Row blocks (of 768 measurement rows) rank ``rank`` owns in a row-partitioned solve — the same split
    ``score_comm_init`` computes on the device side (api.cu)."""
    return range(rank * n_blocks // world_size, (rank + 1) * n_blocks // world_size)


def solve_row_partitioned(prob, device: Optional[int] = None, group=None, **solver_kw):
    """Solve ONE large lowered instance with its measurement rows partitioned over the ranks of a
    ``torch.distributed`` group (BASELINE configs[4]).

    Every rank passes the same ``prob``; the column-space vectors are replicated, each rank applies B and B^T
    for its own rows and one NCCL all-reduce per iteration sums the B^T u partials together with the scalar
    partial sums.  Returns ``(stats, (poses, rounded, landmarks, dist))`` on every rank (identical bits).
    """
    import torch
    import torch.distributed as dist

    from .solver import ScoreSolver

    if prob.n_instances != 1:
        raise ValueError("row partitioning applies to a single instance; shard batches with solve_score_sharded")
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    if device is None:
        device = rank % max(1, torch.cuda.device_count())
    solver = ScoreSolver(prob, device=device)
    try:
        if world > 1:
            box = [ScoreSolver.nccl_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(box, src=0, group=group)
            solver.comm_init(world, rank, box[0])
        stats = solver.solve(**solver_kw)
        return stats, solver.solution()
    finally:
        solver.close()
