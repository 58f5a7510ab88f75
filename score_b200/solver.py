"""Thin object wrapper over the C ABI handle (include/score_b200.h)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np

from . import _lib
from .lowering import QCQP_RELAXATION, LoweredProblem


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


KERNEL_NAMES = ["k_rowpass", "k_linesearch", "k_ctrl_a", "k_rowupdate", "k_coarse_build", "k_colpass", "k_precond_rev",
                "k_coarse_apply", "k_precond_fwd", "k_ctrl_b", "k_pupdate"]


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
    ) -> SolveStats:
        prm = _lib.ScoreParams()
        prm.device = self.device
        prm.max_newton, prm.max_cg, prm.max_ticks = max_newton, max_cg, max_ticks
        prm.kkt_tol, prm.cg_forcing = kkt_tol, cg_forcing
        prm.cg_per_cycle = cg_per_cycle
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
        )
        return self.last_stats

    def solution(self) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
        """(pose_blocks [P,d,d+1], pose_rounded [P,d,d], landmarks [L,d], dist [K,d] or [K,1])."""
        p = self.prob
        d = p.dim
        poses = np.empty((p.P, d, d + 1))
        rounded = np.empty((p.P, d, d))
        lms = np.empty((p.L, d))
        dist = np.empty((p.K, p.dist_per))
        _check(self._lib.score_get_solution(self._h, poses.ctypes.data, rounded.ctypes.data,
                                            lms.ctypes.data if p.L else None, dist.ctypes.data if p.K else None))
        self.d2h_bytes = poses.nbytes + rounded.nbytes + lms.nbytes + dist.nbytes
        return poses, rounded, lms, dist

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


def round_to_special_orthogonal_batch(mats: np.ndarray, device: int = 0) -> np.ndarray:
    """SO(d) rounding of a stack of d x d matrices on the GPU (score_round_so)."""
    mats = _f64(mats)
    if mats.ndim != 3 or mats.shape[1] != mats.shape[2]:
        raise AssertionError("matrix must be square")
    d = mats.shape[1]
    out = np.empty_like(mats)
    _check(_lib.load().score_round_so(d, mats.shape[0], mats.ctypes.data, out.ctypes.data, device))
    return out
