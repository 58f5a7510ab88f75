"""ctypes binding of libscore_b200.so (C ABI in include/score_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` (or ``python -m
score_b200.build``).  There is no CPU fallback: if the shared object is missing
or no CUDA device is usable, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SCORE_B200_LIB: an alternative build of the same library (A/B runs of compile-time variants, scripts/runs/)
LIB_PATH = os.environ.get("SCORE_B200_LIB") or os.path.join(_HERE, "libscore_b200.so")

SCORE_RELAX_QCQP, SCORE_RELAX_SOCP = 0, 1
SCORE_CSR_FULL, SCORE_CSR_REDUCED, SCORE_CSR_REDUCED_T = 0, 1, 2
SCORE_INT_COARSE_INV, SCORE_INT_RANGE_CURV, SCORE_INT_FRAMES, SCORE_INT_TRACE = 0, 1, 2, 3
SCORE_INT_REF_GRAD, SCORE_INT_REF_DIR, SCORE_INT_REF_HDIR, SCORE_INT_REF_DIAG = 4, 5, 6, 7
SCORE_OK, SCORE_ERR_INVALID, SCORE_ERR_CUDA, SCORE_ERR_STATE, SCORE_ERR_ALLOC = 0, -1, -2, -3, -4

_i32p = C.POINTER(C.c_int32)
_f64p = C.POINTER(C.c_double)


class ScoreProblemDesc(C.Structure):
    _fields_ = [
        ("dim", C.c_int32),
        ("relaxation", C.c_int32),
        ("n_instances", C.c_int32),
        ("reserved0", C.c_int32),
        ("P", C.c_int64),
        ("L", C.c_int64),
        ("E", C.c_int64),
        ("K", C.c_int64),
        ("Lp", C.c_int64),
        ("n_seg", C.c_int64),
        ("pose_off", _i32p),
        ("lm_off", _i32p),
        ("edge_off", _i32p),
        ("rng_off", _i32p),
        ("prior_off", _i32p),
        ("seg_ptr", _i32p),
        ("seg_inst", _i32p),
        ("link_edge", _i32p),
        ("edge_i", _i32p),
        ("edge_j", _i32p),
        ("edge_t", _f64p),
        ("edge_R", _f64p),
        ("edge_k", _f64p),
        ("edge_tau", _f64p),
        ("rng_a", _i32p),
        ("rng_b", _i32p),
        ("rng_dist", _f64p),
        ("rng_w", _f64p),
        ("prior_l", _i32p),
        ("prior_t", _f64p),
        ("prior_w", _f64p),
    ]


class ScoreParams(C.Structure):
    _fields_ = [
        ("device", C.c_int32),
        ("max_newton", C.c_int32),
        ("max_cg", C.c_int32),
        ("max_ticks", C.c_int32),
        ("kkt_tol", C.c_double),
        ("cg_forcing", C.c_double),
        ("cg_per_cycle", C.c_int32),
        ("verbose", C.c_int32),
        ("stream", C.c_void_p),
        ("profile_cycles", C.c_int32),
        ("profile_skip", C.c_int32),
        ("mu0", C.c_double),
        ("mu_factor", C.c_double),
        ("center_tol", C.c_double),
        ("mu_min", C.c_double),
        ("cg_grow_after", C.c_int32),
        ("cg_grow_every", C.c_int32),
        ("coarse_every", C.c_int32),
        ("tail_threshold", C.c_int32),
        ("operator_mode", C.c_int32),
        ("hi_prio_threshold", C.c_int32),
        ("reserved3", C.c_int32),
        ("center_tol_late", C.c_double),
    ]


class ScoreInstanceStats(C.Structure):
    _fields_ = [
        ("solved", C.c_int32),
        ("newton_iters", C.c_int32),
        ("cg_iters", C.c_int32),
        ("ls_failures", C.c_int32),
        ("objective", C.c_double),
        ("rel_kkt", C.c_double),
        ("r_stat", C.c_double),
        ("r_gap", C.c_double),
    ]


class ScoreStats(C.Structure):
    _fields_ = [
        ("n_instances", C.c_int32),
        ("n_solved", C.c_int32),
        ("ticks", C.c_int64),
        ("cycles", C.c_int64),
        ("kernel_launches", C.c_int64),
        ("assemble_ms", C.c_double),
        ("setup_ms", C.c_double),
        ("solve_ms", C.c_double),
        ("extract_ms", C.c_double),
        ("total_ms", C.c_double),
        ("nnz_reduced", C.c_int64),
        ("rows", C.c_int64),
        ("cols", C.c_int64),
        ("algorithmic_bytes", C.c_double),
        ("kernel_ms", C.c_double * 16),
        ("kernel_count", C.c_int64 * 16),
        ("profiled_cycles", C.c_int64),
        ("kernel_bytes", C.c_double * 16),
        ("kernel_bytes_total", C.c_double * 16),
        ("kernel_ms_full", C.c_double * 16),
        ("kernel_count_full", C.c_int64 * 16),
    ]


class ScoreRefineParams(C.Structure):
    _fields_ = [
        ("max_outer", C.c_int32),
        ("max_inner", C.c_int32),
        ("rel_tol", C.c_double),
        ("lambda0", C.c_double),
        ("cg_tol", C.c_double),
        ("stream", C.c_void_p),
        ("preconditioner", C.c_int32),
        ("reserved", C.c_int32),
    ]


class ScoreRefineStats(C.Structure):
    _fields_ = [
        ("n_instances", C.c_int32),
        ("n_converged", C.c_int32),
        ("outer_iterations", C.c_int32),
        ("kernel_launches", C.c_int32),
        ("refine_ms", C.c_double),
    ]


class ScoreRefineInstanceStats(C.Structure):
    _fields_ = [
        ("cost_initial", C.c_double),
        ("cost_final", C.c_double),
        ("outer_iterations", C.c_int32),
        ("accepted_steps", C.c_int32),
    ]


EXPORTED_SYMBOLS = [
    "score_create",
    "score_solve",
    "score_get_sizes",
    "score_get_solution",
    "score_get_csr",
    "score_nccl_unique_id",
    "score_comm_init",
    "score_get_internal",
    "score_round_so",
    "score_trajectory_ate",
    "score_eval_ate",
    "score_refine",
    "score_get_refined",
    "score_destroy",
    "score_release_cached",
    "score_last_error",
    "score_version",
]

_lib = None


class ScoreLibraryError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the shared library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ScoreLibraryError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(score_b200 has no CPU fallback)"
        )
    lib = C.CDLL(LIB_PATH)
    lib.score_create.argtypes = [C.POINTER(ScoreProblemDesc), C.c_int32, C.POINTER(C.c_void_p)]
    lib.score_create.restype = C.c_int
    lib.score_solve.argtypes = [C.c_void_p, C.POINTER(ScoreParams), C.POINTER(ScoreStats), C.POINTER(ScoreInstanceStats)]
    lib.score_solve.restype = C.c_int
    lib.score_get_sizes.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.score_get_sizes.restype = C.c_int
    lib.score_get_solution.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.score_get_solution.restype = C.c_int
    lib.score_get_csr.argtypes = [
        C.c_void_p,
        C.c_int32,
        C.c_int32,
        C.POINTER(C.c_int64),
        C.POINTER(C.c_int64),
        C.POINTER(C.c_int64),
        C.c_void_p,
        C.c_void_p,
        C.c_void_p,
        C.c_void_p,
        C.c_void_p,
    ]
    lib.score_get_csr.restype = C.c_int
    lib.score_nccl_unique_id.argtypes = [C.c_char_p]
    lib.score_nccl_unique_id.restype = C.c_int
    lib.score_comm_init.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_char_p]
    lib.score_comm_init.restype = C.c_int
    lib.score_get_internal.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
    lib.score_get_internal.restype = C.c_int
    lib.score_round_so.argtypes = [C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32]
    lib.score_round_so.restype = C.c_int
    lib.score_trajectory_ate.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
    lib.score_trajectory_ate.restype = C.c_int
    lib.score_eval_ate.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                   C.c_void_p]
    lib.score_eval_ate.restype = C.c_int
    lib.score_refine.argtypes = [C.c_void_p, C.POINTER(ScoreRefineParams), C.c_void_p, C.c_void_p,
                                 C.POINTER(ScoreRefineStats), C.c_void_p]
    lib.score_refine.restype = C.c_int
    lib.score_get_refined.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.score_get_refined.restype = C.c_int
    lib.score_destroy.argtypes = [C.c_void_p]
    lib.score_destroy.restype = None
    lib.score_release_cached.argtypes = []
    lib.score_release_cached.restype = None
    lib.score_last_error.restype = C.c_char_p
    lib.score_version.restype = C.c_char_p
    _lib = lib
    return lib


def last_error() -> str:
    return load().score_last_error().decode("utf-8", "replace")
