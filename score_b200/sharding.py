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
