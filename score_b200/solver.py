"""Thin object wrapper over the C ABI handle (include/score_b200.h)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np

from . import _lib
from .lowering import QCQP_RELAXATION, LoweredProblem, slice_instances


@dataclass
class SolveStats:
    n_instances: int
    n_solved: int
    ticks: int
    kernel_launches: int
    assemble_ms: float
    setup_ms: float
    solve_ms: float
    extract_ms: float
    total_ms: float
    nnz_reduced: int
    rows: int
    cols: int
    algorithmic_bytes: float
    instances: np.ndarray  # structured array, one record per instance
    kernel_ms: Optional[np.ndarray] = None  # profile mode: per-kernel event time summed over the profiled cycles
    profiled_cycles: int = 0
    kernel_bytes: Optional[np.ndarray] = None  # algorithmic bytes of one launch of each tick kernel, all instances active
    kernel_count: Optional[np.ndarray] = None  # profile mode: launches of each tick kernel in the profiled cycles
    kernel_bytes_total: Optional[np.ndarray] = None  # algorithmic bytes of each tick kernel over the whole solve
    cycles: int = 0
    kernel_ms_full: Optional[np.ndarray] = None  # profile mode: launches whose work list is the whole batch
    kernel_count_full: Optional[np.ndarray] = None


KERNEL_NAMES = ["k_rowpass", "k_linesearch", "k_ctrl_a", "k_rowupdate", "k_coarse_build", "k_colpass", "k_precond_rev",
                "k_coarse_apply", "k_precond_fwd", "k_ctrl_b", "k_pupdate", "k_pcg_fused", "k_hessvec"]


_INST_DTYPE = np.dtype(
    [
        ("solved", np.int32),
        ("newton_iters", np.int32),
        ("cg_iters", np.int32),
        ("ls_failures", np.int32),
        ("objective", np.float64),
        ("rel_kkt", np.float64),
        ("r_stat", np.float64),
        ("r_gap", np.float64),
    ]
)

_REFINE_DTYPE = np.dtype([("cost_initial", np.float64), ("cost_final", np.float64), ("outer_iterations", np.int32),
                          ("accepted_steps", np.int32)])


def _check(rc: int) -> None:
    if rc == 0:
        return
    msg = _lib.last_error()
    if rc == -1:
        raise ValueError(msg)
    raise RuntimeError(f"libscore_b200 error {rc}: {msg}")


def _i32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


class ScoreSolver:
    """Owns one device-resident problem (a single graph or a batch of instances)."""

    def __init__(self, prob: LoweredProblem, device: int = 0):
        self._lib = _lib.load()
        self.prob = prob
        self.device = device
        self._h = C.c_void_p()
        # keep the contiguous arrays alive for the duration of score_create
        keep = {}
        desc = _lib.ScoreProblemDesc()
        desc.dim = prob.dim
        desc.relaxation = _lib.SCORE_RELAX_QCQP if prob.relaxation == QCQP_RELAXATION else _lib.SCORE_RELAX_SOCP
        desc.n_instances = prob.n_instances
        desc.P, desc.L, desc.E, desc.K, desc.Lp = prob.P, prob.L, prob.E, prob.K, prob.Lp
        desc.n_seg = prob.n_seg
        for name in ("pose_off", "lm_off", "edge_off", "rng_off", "prior_off", "seg_ptr", "seg_inst", "link_edge",
                     "edge_i", "edge_j", "rng_a", "rng_b", "prior_l"):
            keep[name] = _i32(getattr(prob, name))
            setattr(desc, name, keep[name].ctypes.data_as(C.POINTER(C.c_int32)))
        for name in ("edge_t", "edge_R", "edge_k", "edge_tau", "rng_dist", "rng_w", "prior_t", "prior_w"):
            keep[name] = _f64(getattr(prob, name))
            setattr(desc, name, keep[name].ctypes.data_as(C.POINTER(C.c_double)))
        _check(self._lib.score_create(C.byref(desc), device, C.byref(self._h)))
        self.h2d_bytes = int(sum(a.nbytes for a in keep.values()))
        self.last_stats: Optional[SolveStats] = None

    # -- lifecycle -------------------------------------------------------------------------
    def close(self) -> None:
        if self._h:
            self._lib.score_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- row-partitioned multi-GPU solve of one large instance -------------------------------
    @staticmethod
    def nccl_unique_id() -> bytes:
        """128-byte NCCL id (call on rank 0, ship to the other ranks)."""
        buf = C.create_string_buffer(128)
        _check(_lib.load().score_nccl_unique_id(buf))
        return buf.raw

    def comm_init(self, n_ranks: int, rank: int, unique_id: bytes) -> None:
        """Join the communicator: from now on ``solve()`` splits the measurement rows over the ranks and
        all-reduces the B^T u partials once per iteration (score_comm_init)."""
        if len(unique_id) != 128:
            raise ValueError("NCCL unique id must be 128 bytes")
        _check(self._lib.score_comm_init(self._h, n_ranks, rank, unique_id))

    # -- solve -----------------------------------------------------------------------------
    def solve(
        self,
        kkt_tol: float = 1e-6,
        max_newton: int = 0,
        max_cg: int = 0,
        max_ticks: int = 0,
        cg_forcing: float = 0.0,
        cg_per_cycle: int = 0,
        stream: int = 0,
        profile_cycles: int = 0,
        profile_skip: int = 0,
        cg_grow_after: int = 0,
        cg_grow_every: int = 0,
        coarse_every: int = 0,
        mu0: float = 0.0,
        mu_factor: float = 0.0,
        center_tol: float = 0.0,
        mu_min: float = 0.0,
        verbose: int = 0,
        tail_threshold: int = 0,
        operator_mode: int = 0,
        hi_prio_threshold: int = 0,
        center_tol_late: float = 0.0,
    ) -> SolveStats:
        prm = _lib.ScoreParams()
        prm.device = self.device
        prm.max_newton, prm.max_cg, prm.max_ticks = max_newton, max_cg, max_ticks
        prm.kkt_tol, prm.cg_forcing = kkt_tol, cg_forcing
        prm.cg_per_cycle = cg_per_cycle
        prm.verbose = verbose
        prm.tail_threshold = tail_threshold
        prm.operator_mode = operator_mode
        prm.hi_prio_threshold = hi_prio_threshold
        prm.center_tol_late = center_tol_late
        prm.cg_grow_after, prm.cg_grow_every = cg_grow_after, cg_grow_every
        prm.coarse_every = coarse_every
        prm.stream = C.c_void_p(stream) if stream else None
        prm.profile_cycles, prm.profile_skip = profile_cycles, profile_skip
        prm.mu0, prm.mu_factor, prm.center_tol, prm.mu_min = mu0, mu_factor, center_tol, mu_min
        st = _lib.ScoreStats()
        inst = np.zeros(self.prob.n_instances, dtype=_INST_DTYPE)
        _check(self._lib.score_solve(self._h, C.byref(prm), C.byref(st),
                                     inst.ctypes.data_as(C.POINTER(_lib.ScoreInstanceStats))))
        self.last_stats = SolveStats(
            st.n_instances, st.n_solved, st.ticks, st.kernel_launches, st.assemble_ms, st.setup_ms, st.solve_ms,
            st.extract_ms, st.total_ms, st.nnz_reduced, st.rows, st.cols, st.algorithmic_bytes, inst,
            kernel_ms=np.array(st.kernel_ms[:len(KERNEL_NAMES)]), profiled_cycles=int(st.profiled_cycles),
            kernel_bytes=np.array(st.kernel_bytes[:len(KERNEL_NAMES)]),
            kernel_count=np.array(st.kernel_count[:len(KERNEL_NAMES)]),
            kernel_bytes_total=np.array(st.kernel_bytes_total[:len(KERNEL_NAMES)]), cycles=int(st.cycles),
            kernel_ms_full=np.array(st.kernel_ms_full[:len(KERNEL_NAMES)]),
            kernel_count_full=np.array(st.kernel_count_full[:len(KERNEL_NAMES)]),
        )
        return self.last_stats

    def solution_shapes(self):
        p = self.prob
        d = p.dim
        return (p.P, d, d + 1), (p.P, d, d), (p.L, d), (p.K, p.dist_per)

    def solution(self, out=None) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
        """(pose_blocks [P,d,d+1], pose_rounded [P,d,d], landmarks [L,d], dist [K,d] or [K,1]).

        ``out``: optional tuple of four C-contiguous float64 arrays of ``solution_shapes()`` to fill (e.g. views of
        pinned host memory, which makes the device-to-host copy run at full PCIe speed)."""
        p = self.prob
        if out is None:
            poses, rounded, lms, dist = (np.empty(s) for s in self.solution_shapes())
        else:
            poses, rounded, lms, dist = out
            for a, s in zip(out, self.solution_shapes()):
                if a.shape != s or a.dtype != np.float64 or not a.flags.c_contiguous:
                    raise ValueError("solution(out=...): arrays must be C-contiguous float64 of solution_shapes()")
        _check(self._lib.score_get_solution(self._h, poses.ctypes.data, rounded.ctypes.data,
                                            lms.ctypes.data if p.L else None, dist.ctypes.data if p.K else None))
        self.d2h_bytes = poses.nbytes + rounded.nbytes + lms.nbytes + dist.nbytes
        return poses, rounded, lms, dist

    def ate(self, gt_positions, traj_off=None, align: bool = True):
        """Absolute trajectory error of the solved translations against ``gt_positions`` [P, d] on the device
        (score_eval_ate): one trajectory per instance, or per ``traj_off`` range of the batch's global pose
        numbering (e.g. ``lowering.chain_offsets(prob)`` for one trajectory per robot).  Returns
        (rmse [n_traj], R [n_traj, d, d], t [n_traj, d]) with gt ~ R est + t."""
        p = self.prob
        d = p.dim
        gt = _f64(gt_positions)
        if gt.shape != (p.P, d):
            raise ValueError(f"gt_positions must have shape {(p.P, d)}, got {gt.shape}")
        off = None if traj_off is None else _i32(traj_off)
        n = p.n_instances if off is None else len(off) - 1
        rmse, R, t = np.empty(n), np.empty((n, d, d)), np.empty((n, d))
        _check(self._lib.score_eval_ate(self._h, n, None if off is None else off.ctypes.data, gt.ctypes.data,
                                        1 if align else 0, rmse.ctypes.data, R.ctypes.data, t.ctypes.data))
        return rmse, R, t

    def refine(self, max_outer: int = 0, max_inner: int = 0, rel_tol: float = 0.0, lambda0: float = 0.0, cg_tol: float = 0.0,
               init=None, preconditioner: int = 0):
        """Local refinement of the last solve's rounded estimate on the original non-convex cost (score_refine: batched
        Levenberg-Marquardt, /root/reference/README.md:63-67 — the step the paper hands to GTSAM).  ``init``: optional
        (poses [P, d, d+1] as [R|t] with R in SO(d), landmarks [L, d]) to start from instead.  Returns
        (per-instance record array: cost_initial, cost_final, outer_iterations, accepted_steps; stats dict); the refined
        estimate is read with ``refined()``."""
        p = self.prob
        prm = _lib.ScoreRefineParams()
        prm.max_outer, prm.max_inner = max_outer, max_inner
        prm.rel_tol, prm.lambda0, prm.cg_tol = rel_tol, lambda0, cg_tol
        prm.preconditioner = preconditioner
        st = _lib.ScoreRefineStats()
        inst = np.zeros(p.n_instances, dtype=_REFINE_DTYPE)
        ip = il = None
        if init is not None:
            ip, il = _f64(init[0]), _f64(init[1])
            if ip.shape != (p.P, p.dim, p.dim + 1) or il.shape != (p.L, p.dim):
                raise ValueError(f"init must be (poses {(p.P, p.dim, p.dim + 1)}, landmarks {(p.L, p.dim)})")
        _check(self._lib.score_refine(self._h, C.byref(prm), None if ip is None else ip.ctypes.data,
                                      None if il is None else il.ctypes.data, C.byref(st), inst.ctypes.data))
        return inst, {"n_converged": int(st.n_converged), "outer_iterations": int(st.outer_iterations),
                      "kernel_launches": int(st.kernel_launches), "refine_ms": float(st.refine_ms)}

    def refined(self) -> Tuple[np.ndarray, np.ndarray]:
        """(poses [P, d, d+1] as [R|t] with R in SO(d), landmarks [L, d]) after ``refine``."""
        p = self.prob
        poses, lms = np.empty((p.P, p.dim, p.dim + 1)), np.empty((p.L, p.dim))
        _check(self._lib.score_get_refined(self._h, poses.ctypes.data, lms.ctypes.data if p.L else None))
        return poses, lms

    def internal(self, which: int, inst: int = 0) -> np.ndarray:
        """Solver internals of one instance (score_get_internal): flat float64 array."""
        n = C.c_int64()
        _check(self._lib.score_get_internal(self._h, which, inst, None, 0, C.byref(n)))
        out = np.empty(n.value)
        _check(self._lib.score_get_internal(self._h, which, inst, out.ctypes.data if n.value else None, n.value,
                                            C.byref(n)))
        return out

    def sizes(self) -> Tuple[int, int, int]:
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        _check(self._lib.score_get_sizes(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def csr(self, which: int = _lib.SCORE_CSR_FULL, inst: int = 0):
        """Assembled matrix of one instance: (indptr, indices, values, weights, rhs, shape)."""
        nr, nc, nnz = C.c_int64(), C.c_int64(), C.c_int64()
        _check(self._lib.score_get_csr(self._h, which, inst, C.byref(nr), C.byref(nc), C.byref(nnz),
                                       None, None, None, None, None))
        indptr = np.empty(nr.value + 1, np.int32)
        indices = np.empty(nnz.value, np.int32)
        values = np.empty(nnz.value)
        weights = np.zeros(nr.value)
        rhs = np.zeros(nr.value)
        _check(self._lib.score_get_csr(self._h, which, inst, C.byref(nr), C.byref(nc), C.byref(nnz),
                                       indptr.ctypes.data, indices.ctypes.data, values.ctypes.data,
                                       weights.ctypes.data, rhs.ctypes.data))
        return indptr, indices, values, weights, rhs, (nr.value, nc.value)


class ScoreSolverGroup:
    """A batch split into sub-batches, each with its own handle and CUDA stream, solved by ``n_streams`` host threads.

    Independent instances never interact, so the sub-batches are solved concurrently: kernels of different
    streams fill each other's ramp-up / tail gaps on the GPU, and in the create -> solve -> read-back pipeline
    the host-side upload of one sub-batch overlaps the device-side solve of another (``run_pipelined``).
    Per-instance results are bit-identical to a single-handle solve (every instance is reduced in fixed order).
    """

    def __init__(self, prob: LoweredProblem, n_streams: int = 2, device: int = 0, create: bool = True,
                 n_parts: int = 0):
        """``n_parts`` (default: ``n_streams``) sub-batches are worked through by ``n_streams`` host threads, each
        sub-batch on its own CUDA stream.  With more parts than streams the sub-batches start staggered, so the
        sparsely occupied last cycles of one (a few slow instances) run under the dense first cycles of the next."""
        from concurrent.futures import ThreadPoolExecutor

        n = max(1, min(int(n_parts) if n_parts > 0 else int(n_streams), prob.n_instances))
        n_streams = max(1, min(int(n_streams), n))
        cuts = [round(j * prob.n_instances / n) for j in range(n + 1)]
        self.parts = [slice_instances(prob, cuts[j], cuts[j + 1]) for j in range(n)]
        self.prob, self.device, self.cuts = prob, device, cuts
        self.n_streams = n_streams
        self.pool = ThreadPoolExecutor(n_streams)
        self.pipe_pool = None
        self.solvers: List[Optional[ScoreSolver]] = [None] * n
        if create:
            self.solvers = list(self.pool.map(lambda part: ScoreSolver(part, device=device), self.parts))

    def close(self) -> None:
        for st in getattr(self, "_sets", [self.solvers])[1:]:
            for s in st:
                s.close()
        self._sets = [self.solvers]
        for s in self.solvers:
            if s is not None:
                s.close()
        self.solvers = [None] * len(self.parts)
        self.pool.shutdown(wait=True)
        if getattr(self, "_steps_pool", None) is not None:
            self._steps_pool.shutdown(wait=True)
            self._steps_pool = None
        if self.pipe_pool is not None:
            self.pipe_pool.shutdown(wait=True)
            self.pipe_pool = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @staticmethod
    def merge_stats(stats: List[SolveStats]) -> SolveStats:
        f = stats[0]
        summed = lambda name: sum(getattr(s, name) for s in stats)
        longest = lambda name: max(getattr(s, name) for s in stats)
        return SolveStats(
            summed("n_instances"), summed("n_solved"), longest("ticks"), summed("kernel_launches"), longest("assemble_ms"),
            longest("setup_ms"), longest("solve_ms"), longest("extract_ms"), longest("total_ms"), summed("nnz_reduced"),
            summed("rows"), summed("cols"), summed("algorithmic_bytes"), np.concatenate([s.instances for s in stats]),
            kernel_ms=sum(s.kernel_ms for s in stats), profiled_cycles=f.profiled_cycles,
            kernel_bytes=sum(s.kernel_bytes for s in stats), kernel_count=sum(s.kernel_count for s in stats),
            kernel_bytes_total=sum(s.kernel_bytes_total for s in stats), cycles=longest("cycles"),
            kernel_ms_full=sum(s.kernel_ms_full for s in stats) if f.kernel_ms_full is not None else None,
            kernel_count_full=sum(s.kernel_count_full for s in stats) if f.kernel_count_full is not None else None,
        )

    def solve(self, **kw) -> SolveStats:
        kw.pop("stream", None)  # every handle runs on its own (library-owned, non-blocking) stream
        return self.merge_stats(list(self.pool.map(lambda s: s.solve(**kw), self.solvers)))

    def solve_steps(self, steps: int, n_sets: int = 1, stagger: bool = True, **kw) -> List[SolveStats]:
        """Solve the whole batch ``steps`` times as a pipeline: every sub-batch handle is re-solved by its own host thread
        with NO barrier between the steps, the sub-batches starting a fraction of a solve apart.  With ``n_sets`` > 1 the
        steps are double-buffered over that many sets of handles (step k runs on set k mod n_sets), so that
        ``n_sets x n_parts`` sub-batch solves are in flight.

        Why: a sweep's last cycles are sparse (a handful of ill-conditioned instances keep a sub-batch alive long after
        the others have finished) and leave the GPU mostly idle; with several sub-batch solves in flight and out of phase,
        the tails of some run under the dense first cycles of others (profiles/pipeline_probe_r2.txt).  Returns one merged
        SolveStats per step.  Per-instance results are bit-identical to ``solve`` (instances never interact)."""
        import time
        from concurrent.futures import ThreadPoolExecutor

        kw.pop("stream", None)
        n = len(self.solvers)
        n_sets = max(1, min(int(n_sets), steps))
        if not hasattr(self, "_sets"):
            self._sets = [self.solvers]
        while len(self._sets) < n_sets:  # further handle sets over the same sub-batches
            self._sets.append(list(self.pool.map(lambda part: ScoreSolver(part, device=self.device), self.parts)))
        if getattr(self, "_steps_pool", None) is None or self._steps_pool._max_workers < n * n_sets:
            if getattr(self, "_steps_pool", None) is not None:
                self._steps_pool.shutdown(wait=True)
            self._steps_pool = ThreadPoolExecutor(n * n_sets)
        delay = getattr(self, "_last_part_s", 0.0) / (n * n_sets) if stagger else 0.0
        counts = [len(range(si, steps, n_sets)) for si in range(n_sets)]  # steps served by every set

        def run(job):
            si, j = divmod(job, n)
            if delay > 0 and job > 0:
                time.sleep(job * delay)
            out = []
            for _ in range(counts[si]):
                t0 = time.perf_counter()
                out.append(self._sets[si][j].solve(**kw))
                self._last_part_s = time.perf_counter() - t0
            return out

        per_job = list(self._steps_pool.map(run, range(n * n_sets)))
        return [self.merge_stats([per_job[(k % n_sets) * n + j][k // n_sets] for j in range(n)]) for k in range(steps)]

    def solution(self, out=None):
        shapes = self._shapes()
        if out is None:
            out = tuple(np.empty(s) for s in shapes)
        views = self._views(out)
        list(self.pool.map(lambda a: a[0].solution(out=a[1]), zip(self.solvers, views)))
        return out

    def _shapes(self):
        p = self.prob
        d = p.dim
        return (p.P, d, d + 1), (p.P, d, d), (p.L, d), (p.K, p.dist_per)

    def _views(self, out):
        p, c = self.prob, self.cuts
        views = []
        for j in range(len(self.parts)):
            a, b = c[j], c[j + 1]
            views.append((out[0][p.pose_off[a]:p.pose_off[b]], out[1][p.pose_off[a]:p.pose_off[b]],
                          out[2][p.lm_off[a]:p.lm_off[b]], out[3][p.rng_off[a]:p.rng_off[b]]))
        return views

    def prewarm(self, extra: int = 1) -> None:
        """Create and destroy ``n_streams + 1 + extra`` handles at once, so that the library's chunk / stream / event
        caches hold enough for the streamed pipeline plus a spare set (a handle that has to go to the driver for
        memory while others solve costs hundreds of milliseconds)."""
        k = self.n_streams + 1 + max(0, int(extra))
        hs = [ScoreSolver(self.parts[i % len(self.parts)], device=self.device) for i in range(k)]
        for h in hs:
            h.close()

    def run_pipelined(self, out=None, steps: int = 1, inflight: int = 0, **kw):
        """create -> solve -> read-back -> destroy of every sub-batch in its own thread (the end-to-end path).

        ``steps`` > 1 repeats the whole batch that many times as ONE queue of sub-batch jobs (every job uploads its
        inputs again and reads its results back): ``n_streams`` jobs solve at a time while one more thread already
        runs the host-side ``score_create`` (table build + upload) of the next job, so the GPU does not idle during
        the uploads — the way a long Monte-Carlo sweep is streamed through a GPU.
        Returns (stats, arrays, h2d_bytes, d2h_bytes), the byte counts per step."""
        import threading
        from concurrent.futures import ThreadPoolExecutor

        kw.pop("stream", None)
        if out is None:
            out = tuple(np.empty(s) for s in self._shapes())
        views = self._views(out)
        n = len(self.parts)
        inflight = int(inflight) if inflight > 0 else self.n_streams  # sub-batch solves running at the same time
        solving = threading.Semaphore(inflight)
        creating = threading.Lock()  # one score_create at a time: concurrent creates only contend (host cores, driver)
        out_locks = [threading.Lock() for _ in range(n)]

        def one(job):
            j = job % n
            with creating:
                s = ScoreSolver(self.parts[j], device=self.device)
            with s:
                with solving:
                    st = s.solve(**kw)
                with out_locks[j]:
                    s.solution(out=views[j])
                return st, s.h2d_bytes, s.d2h_bytes

        steps = max(1, int(steps))
        if steps == 1:
            res = list(self.pool.map(one, range(n)))
        else:
            if self.pipe_pool is None or self.pipe_pool._max_workers < inflight + 1:
                if self.pipe_pool is not None:
                    self.pipe_pool.shutdown(wait=True)
                self.pipe_pool = ThreadPoolExecutor(inflight + 1)  # kept for the life of the group
            res = list(self.pipe_pool.map(one, range(steps * n)))
        return (self.merge_stats([r[0] for r in res]), out, sum(r[1] for r in res) // steps,
                sum(r[2] for r in res) // steps)


class HandlePool:
    """Ready handles for a list of sub-batches, ``copies`` per sub-batch (device-resident inputs): the job function of
    a dynamically scheduled sweep (``sharding.SweepQueue``) borrows one for the duration of a ``score_solve``.  Two
    copies let two passes over the same sub-batch be in flight on one GPU; a third borrower waits."""

    def __init__(self, parts: List[LoweredProblem], device: int = 0, copies: int = 2, threads: int = 4):
        import queue
        from concurrent.futures import ThreadPoolExecutor

        self.parts, self.device = list(parts), device
        self._free = [queue.Queue() for _ in self.parts]
        self._all: List[ScoreSolver] = []
        with ThreadPoolExecutor(max(1, threads)) as ex:
            made = list(ex.map(lambda g: ScoreSolver(self.parts[g], device=device),
                               [g for g in range(len(self.parts)) for _ in range(max(1, copies))]))
        for k, h in enumerate(made):
            self._all.append(h)
            self._free[k // max(1, copies)].put(h)

    def solve(self, part: int, **kw) -> SolveStats:
        kw.pop("stream", None)  # every handle runs on its own library-owned stream
        h = self._free[part].get()
        try:
            return h.solve(**kw)
        finally:
            self._free[part].put(h)

    def warm(self, threads: int = 4, **kw) -> List[SolveStats]:
        """Solve on EVERY handle once (first-use work: graph capture, workspace) — one SolveStats per sub-batch (of its
        last copy; iteration counts are deterministic, so every rank sees the same)."""
        from concurrent.futures import ThreadPoolExecutor

        kw.pop("stream", None)
        with ThreadPoolExecutor(max(1, threads)) as ex:
            sts = list(ex.map(lambda h: h.solve(**kw), self._all))
        copies = len(self._all) // max(1, len(self.parts))
        return [sts[g * copies + copies - 1] for g in range(len(self.parts))]

    def close(self) -> None:
        for h in self._all:
            h.close()
        self._all = []

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class StreamedJobs:
    """create -> solve -> read-back -> destroy of one sub-batch per call, from HOST arrays (the end-to-end job of a
    dynamically scheduled sweep): at most ``inflight`` solves at a time on this GPU while one more thread may already be
    inside ``score_create`` of the next job; results land in the calling worker's own (pinned) output slot."""

    def __init__(self, parts: List[LoweredProblem], outs: List[tuple], device: int = 0, inflight: int = 4):
        import threading

        self.parts, self.outs, self.device = list(parts), outs, device
        self._solving = threading.Semaphore(max(1, inflight))
        self._creating = threading.Lock()
        self.h2d_bytes = self.d2h_bytes = 0
        self._acc = threading.Lock()

    def views(self, worker: int, prob: LoweredProblem):
        o = self.outs[worker]
        return (o[0][: prob.P], o[1][: prob.P], o[2][: prob.L], o[3][: prob.K])

    def __call__(self, step: int, part: int, worker: int, **kw) -> SolveStats:
        prob = self.parts[part]
        with self._creating:
            s = ScoreSolver(prob, device=self.device)
        with s:
            with self._solving:
                st = s.solve(**kw)
            s.solution(out=self.views(worker, prob))
            with self._acc:
                self.h2d_bytes += s.h2d_bytes
                self.d2h_bytes += s.d2h_bytes
        return st


def round_to_special_orthogonal_batch(mats: np.ndarray, device: int = 0) -> np.ndarray:
    """SO(d) rounding of a stack of d x d matrices on the GPU (score_round_so)."""
    mats = _f64(mats)
    if mats.ndim != 3 or mats.shape[1] != mats.shape[2]:
        raise AssertionError("matrix must be square")
    d = mats.shape[1]
    out = np.empty_like(mats)
    _check(_lib.load().score_round_so(d, mats.shape[0], mats.ctypes.data, out.ctypes.data, device))
    return out


def trajectory_ate(est, gt, traj_off=None, align: bool = True, device: int = 0):
    """SE(d)-aligned absolute trajectory error of a batch of trajectories on the GPU (score_trajectory_ate).

    est, gt: [n, d] positions (d = 2 or 3); traj_off: [n_traj + 1] offsets (default: one trajectory).
    Returns (rmse [n_traj], R [n_traj, d, d], t [n_traj, d]) with gt ~ R est + t; empty trajectories give NaN."""
    est, gt = _f64(est), _f64(gt)
    if est.ndim != 2 or est.shape != gt.shape:
        raise ValueError("est and gt must both be [n, d]")
    d = est.shape[1]
    off = _i32([0, est.shape[0]] if traj_off is None else traj_off)
    if off.ndim != 1 or len(off) < 1 or (len(off) > 1 and int(off[-1]) > est.shape[0]):
        raise ValueError("traj_off must be a 1-D offset array within the point arrays")
    n = len(off) - 1
    rmse, R, t = np.empty(n), np.empty((n, d, d)), np.empty((n, d))
    _check(_lib.load().score_trajectory_ate(d, n, off.ctypes.data, est.ctypes.data, gt.ctypes.data, 1 if align else 0,
                                            rmse.ctypes.data, R.ctypes.data, t.ctypes.data, device))
    return rmse, R, t
