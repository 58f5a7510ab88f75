/*
 * score_b200.h — C ABI of the B200-native SCORE solver (libscore_b200.so).
 *
 * Drop-in boundary for the hot path of MarineRoboticsGroup/score:
 *   score/solve_score.py:54-86   solve_score(data, relaxation_type)
 * i.e. everything the reference does between "I have a FactorGraphData" and
 * "I have relaxed + rounded variable values":
 *   model build        score/utils/gurobi_utils.py:173-187  (initialize_model)
 *   barrier solve      score/solve_score.py:76               (model.optimize())
 *   value extraction   score/utils/gurobi_utils.py:114-136   (get_variable_values)
 *
 * The Python wrapper (score_b200/solve_score.py) does the name -> index lowering
 * and the dict packing; everything numeric happens behind these entry points on
 * the GPU.  There is no CPU fallback: every call fails with SCORE_ERR_CUDA when no
 * sm_100 device is usable.
 *
 * Conventions
 *   - plain C types only; all arrays are caller-owned; pointers in ScoreProblemDesc
 *     may be host OR device pointers (copied with cudaMemcpyDefault at create time);
 *     output pointers of the getters are HOST pointers unless stated otherwise.
 *   - indices are int32, zero based, instance-local; sizes are int64.
 *   - return value 0 on success, negative ScoreStatus on error; the message is
 *     available from score_last_error() (thread-local).
 *   - one handle is used from one host thread at a time.
 *
 * Column / row order (bit-exact with the reference's variable creation order,
 * gurobi_utils.py:233-310 and objective accumulation order :358-377):
 *   columns: pose p -> p*d*(d+1) + r*(d+1) + c   (c<d: R[r,c], c==d: t[r])
 *            landmark q -> P*d*(d+1) + q*d + r
 *            QCQP delta_k[r] -> P*d*(d+1) + L*d + k*d + r ; SOCP delta_k -> ... + k
 *   rows:    edges (odometry chains in order, then loop closures): d translation
 *            rows then d*d rotation rows (row-major) each; then ranges; then
 *            landmark priors.
 */
#ifndef SCORE_B200_H_
#define SCORE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  SCORE_OK = 0,
  SCORE_ERR_INVALID = -1,   /* bad argument / inconsistent description          */
  SCORE_ERR_CUDA = -2,      /* CUDA runtime failure or no usable device         */
  SCORE_ERR_STATE = -3,     /* call out of order (e.g. get_solution before solve) */
  SCORE_ERR_ALLOC = -4
} ScoreStatus;

enum { SCORE_RELAX_QCQP = 0, SCORE_RELAX_SOCP = 1 };

/* Which matrix score_get_csr returns. */
enum {
  SCORE_CSR_FULL = 0,     /* the reference's least-squares matrix incl. delta columns (assembly parity) */
  SCORE_CSR_REDUCED = 1,  /* operator the solver iterates with: pose+landmark columns only              */
  SCORE_CSR_REDUCED_T = 2 /* its transpose (CSR of B^T)                                                 */
};

/*
 * Lowered factor graph, structure-of-arrays.  A batch of independent instances is
 * the concatenation of their arrays plus the *_off offset tables (n_instances+1
 * entries each; may be NULL when n_instances == 1).
 *
 * Replaces the dict-of-MVar bookkeeping of VariableCollection
 * (gurobi_utils.py:53-136) and the per-factor Python loops (:233-526).
 */
typedef struct {
  int32_t dim;          /* 2 or 3 (is_dimension, gurobi_utils.py:37-50)            */
  int32_t relaxation;   /* SCORE_RELAX_QCQP / SCORE_RELAX_SOCP (:26-28)             */
  int32_t n_instances;
  int32_t reserved0;
  int64_t P, L, E, K, Lp;      /* totals over the batch                             */
  int64_t n_seg;               /* odometry chain segments (paths of consecutive poses) */
  const int32_t *pose_off, *lm_off, *edge_off, *rng_off, *prior_off; /* [n_instances+1] */
  const int32_t *seg_ptr;      /* [n_seg+1] global pose index where each segment starts */
  const int32_t *seg_inst;     /* [n_seg] owning instance                              */
  const int32_t *link_edge;    /* [P] global edge id of the odometry edge (p-1 -> p), -1 at segment starts */
  /* relative-pose factors (odometry then loop closures), get_relative_pose_cost_expression :504-526 */
  const int32_t *edge_i, *edge_j;   /* [E] instance-local pose indices (base, to)   */
  const double *edge_t;             /* [E*d]   measured translation                 */
  const double *edge_R;             /* [E*d*d] measured rotation, row-major         */
  const double *edge_k, *edge_tau;  /* [E] translation / rotation precision         */
  /* range factors, get_single_range_cost :475-501 */
  const int32_t *rng_a, *rng_b;     /* [K] translation owner: pose p -> p, landmark q -> P_inst + q */
  const double *rng_dist, *rng_w;   /* [K] measured distance, precision             */
  /* landmark priors, get_all_landmark_prior_costs :433-446 */
  const int32_t *prior_l;           /* [Lp] instance-local landmark index           */
  const double *prior_t;            /* [Lp*d]                                       */
  const double *prior_w;            /* [Lp]                                         */
} ScoreProblemDesc;

typedef struct {
  int32_t device;          /* CUDA device ordinal                                   */
  int32_t max_newton;      /* outer (Newton) iteration cap, <=0: default; -1: evaluate the start point only */
  int32_t max_cg;          /* inner PCG iteration cap per Newton step, <=0: default  */
  int32_t max_ticks;       /* global cap on solver ticks, <=0: default               */
  double kkt_tol;          /* relative KKT tolerance, <=0: 1e-6    */
  double cg_forcing;       /* inexact-Newton forcing term eta, <=0: 0.2 (x0.3 on the last barrier stages) */
  int32_t cg_per_cycle;    /* PCG ticks between two line-search ticks of the batch, <=0: default 4 */
  int32_t verbose;
  void *stream;            /* cudaStream_t to run on, NULL: library-owned stream     */
  int32_t profile_cycles;  /* >0: time every kernel of this many cycles with CUDA events (un-graphed) */
  int32_t profile_skip;    /* cycles to run before the profiled ones                 */
  /* interior-point path following (0: defaults) */
  double mu0;              /* initial barrier parameter, default 0.1; < 0: no barrier (plain semismooth Newton) */
  double mu_factor;        /* barrier reduction per centred stage, default 0.1       */
  double center_tol;       /* stage ends when Newton decrement^2 / mu <= this, default 16 (early stages, mu > 1e-5) */
  double mu_min;           /* smallest barrier parameter, default 1e-16              */
  int32_t cg_grow_after;   /* from this cycle on the PCG ticks per cycle double every cg_grow_every cycles, <=0: never */
  int32_t cg_grow_every;   /* <=0: 8                                                 */
  int32_t coarse_every;    /* rebuild the coarse inverse every this many Newton steps of a barrier stage, <=0: 1 */
  int32_t tail_threshold;  /* > 0: run the fused per-instance PCG kernel (one thread-block cluster per instance) once at
                              most this many instances are unfinished; <= 0: lockstep ticks only (default) */
  int32_t operator_mode;   /* PCG operator: 0 matrix-free, factor by factor (default); 1 assembled CSR pair (row pass +
                              column pass) */
  int32_t hi_prio_threshold; /* > 0: the cycles after at most this many instances are left unfinished run on a
                              high-priority stream (library-owned stream only); <= 0: never (default; measured: no gain) */
  int32_t reserved3;
  double center_tol_late;  /* the same threshold on the last barrier stages (mu <= 1e-5, the ones whose iterates are
                              certified), default 1 */
} ScoreParams;

/* Per-instance result record. */
typedef struct {
  int32_t solved;        /* 1: rel KKT <= tol (maps to SolverResults.solved)          */
  int32_t newton_iters;
  int32_t cg_iters;      /* total PCG iterations = operator applications            */
  int32_t ls_failures;   /* Newton steps whose line search found no decreasing candidate */
  double objective;      /* f(x) = sum w (Bx-b)^2                                    */
  double rel_kkt, r_stat, r_gap;
} ScoreInstanceStats;

typedef struct {
  int32_t n_instances, n_solved;
  int64_t ticks;           /* solver ticks executed by the batch (line-search + PCG ticks)           */
  int64_t cycles;          /* lockstep cycles (one line-search tick + evaluation tick + PCG ticks)   */
  int64_t kernel_launches; /* kernels launched by this call (incl. inside graphs)    */
  double assemble_ms, setup_ms, solve_ms, extract_ms, total_ms; /* CUDA-event times on the solver stream */
  int64_t nnz_reduced, rows, cols;  /* operator size over the batch                  */
  double algorithmic_bytes;         /* bytes the instances had to move over the whole solve */
  /* profile mode: summed CUDA-event time / launch count of each tick kernel over the profiled cycles; order
   * rowpass, linesearch, ctrl_a, rowupdate, coarse_build, colpass, precond_rev, coarse_apply, precond_fwd,
   * ctrl_b, pupdate, pcg_fused, hessvec */
  double kernel_ms[16];
  int64_t kernel_count[16];
  int64_t profiled_cycles;
  /* algorithmic bytes of ONE launch of each tick kernel with every instance active */
  double kernel_bytes[16];
  /* algorithmic bytes of each tick kernel summed over the whole solve (per-instance iteration counts) */
  double kernel_bytes_total[16];
  /* profile mode: the subset of kernel_ms / kernel_count whose work list is the whole batch while every instance is
   * still running — line-search-only kernels in the line-search tick, all others in the first PCG tick of a cycle
   * (meaningful when the profiled cycles are early ones) */
  double kernel_ms_full[16];
  int64_t kernel_count_full[16];
} ScoreStats;

typedef struct ScoreHandle_ *ScoreHandle;

/* Upload a lowered problem.  (VariableCollection + data plumbing.) */
int score_create(const ScoreProblemDesc *desc, int32_t device, ScoreHandle *out);

/* Assemble, precondition, solve and round.  Replaces initialize_model + model.optimize()
 * (solve_score.py:72-85).  inst_stats may be NULL or point to n_instances records. */
int score_solve(ScoreHandle h, const ScoreParams *params, ScoreStats *stats, ScoreInstanceStats *inst_stats);

/* Sizes of the full variable vector, for buffer allocation. */
int score_get_sizes(ScoreHandle h, int64_t *n_cols_full, int64_t *n_rows_full, int64_t *nnz_full);

/* Solution in the reference's layouts (get_variable_values, gurobi_utils.py:114-136):
 *   pose_blocks   [P*d*(d+1)]  relaxed [R|t] row-major (Var.X of each pose MVar)
 *   pose_rounded  [P*d*d]      nearest SO(d) rotation (round_to_special_orthogonal)
 *   landmarks     [L*d]
 *   dist          [K*d] (QCQP) or [K] (SOCP)
 * any pointer may be NULL. */
int score_get_solution(ScoreHandle h, double *pose_blocks, double *pose_rounded, double *landmarks, double *dist);

/* Assembled matrix (which = SCORE_CSR_*), instance `inst` of the batch, instance-local
 * row/column numbering.  Pass NULL arrays to query sizes only.  weights/rhs have n_rows
 * entries (only filled for SCORE_CSR_FULL / _REDUCED). */
int score_get_csr(ScoreHandle h, int32_t which, int32_t inst, int64_t *n_rows, int64_t *n_cols, int64_t *nnz,
                  int32_t *indptr, int32_t *indices, double *values, double *weights, double *rhs);

/* Row-partitioned multi-GPU solve of ONE large instance (BASELINE configs[4]): every rank creates the
 * same handle on its own GPU, then joins a communicator; score_solve then splits the measurement rows over the
 * ranks and sums the B^T u partials (plus the scalar partial sums, in the same buffer) with one NCCL all-reduce per
 * iteration.  Rank 0 makes the 128-byte id and ships it to the others by any means (the Python wrapper uses
 * torch.distributed).  The reference has no counterpart (single process). */
int score_nccl_unique_id(char *out128);
int score_comm_init(ScoreHandle h, int32_t n_ranks, int32_t rank, const char *id128);

/* Test / diagnostic access to solver internals of instance `inst` after score_solve (host buffer of
 * `capacity` doubles; *count receives the number of doubles the array has; pass out = NULL to query):
 *   SCORE_INT_COARSE_INV  nc x nc inverse coarse matrix of the last Newton step (0 doubles: coarse level off; large
 *                         coarse spaces whose landmark block is eliminated: its nb x nb segment-base block)
 *   SCORE_INT_RANGE_CURV  K_inst x d(d+1)/2 curvature blocks 2 w H_k of the range terms (upper, row-major)
 *   SCORE_INT_FRAMES      P_inst x d x (d+1) dead-reckoned frames of the odometry-chain preconditioner
 *   SCORE_INT_TRACE       (only after a solve with ScoreParams.verbose >= 2) 256 x 8 doubles, one record per Newton
 *                         step: barrier parameter, step length, PCG iterations, decrement^2, F_mu, r.s, ladder shift, eta  *   SCORE_INT_REF_GRAD / _DIR / _HDIR / _DIAG  (after score_refine) tangent-space vectors of the refinement in the
 *                         column-space slots of the instance (pose p: first dof entries of its block, then landmarks):
 *                         gradient J^T W r of the last linearisation, last PCG direction p, (J^T W J + lambda D) p, D
 */
enum { SCORE_INT_COARSE_INV = 0, SCORE_INT_RANGE_CURV = 1, SCORE_INT_FRAMES = 2, SCORE_INT_TRACE = 3, SCORE_INT_REF_GRAD = 4,
       SCORE_INT_REF_DIR = 5, SCORE_INT_REF_HDIR = 6, SCORE_INT_REF_DIAG = 7 };
int score_get_internal(ScoreHandle h, int32_t which, int32_t inst, double *out, int64_t capacity, int64_t *count);

/* Stand-alone SO(d) rounding of n d x d matrices (host pointers) on the device:
 * round_to_special_orthogonal, score/utils/matrix_utils.py:59-79. */
int score_round_so(int32_t dim, int64_t n, const double *mats, double *out, int32_t device);

/* Evaluation step after the path (the reference ships ground truth beside its inputs —
 * PoseVariable.true_position in examples/manhattan/factor_graph.pickle, examples/goats_14_data/gt_traj_A.tum — and
 * leaves the trajectory-error computation to downstream tooling): absolute trajectory error of n_traj
 * trajectories after the best rigid alignment  min_{R in SO(d), t} sum ||gt_i - (R est_i + t)||^2  (Kabsch; the
 * rotation is the same U diag(1,..,det) V^T rule as round_to_special_orthogonal, matrix_utils.py:59-79).
 *   traj_off [n_traj+1]  point offsets of the trajectories (non-decreasing)
 *   est, gt  [n*dim]     estimated / true positions, row-major (host pointers)
 *   align                1: SE(d) alignment, 0: compare as is (R = I, t = 0)
 *   rmse [n_traj], R [n_traj*dim*dim], t [n_traj*dim]   outputs, any may be NULL; an empty trajectory gives NaN
 * One CTA per trajectory, fixed-order reductions (bit-reproducible). */
int score_trajectory_ate(int32_t dim, int32_t n_traj, const int32_t *traj_off, const double *est, const double *gt,
                         int32_t align, double *rmse, double *R, double *t, int32_t device);

/* Same, on the translations of the solution held by a solved handle (no device-to-host copy of the estimate).
 * traj_off: [n_traj+1] offsets into the batch's global pose numbering (e.g. one trajectory per robot chain);
 * NULL = one trajectory per instance (n_traj is then ignored).  gt_pos: [P*dim] host. */
int score_eval_ate(ScoreHandle h, int32_t n_traj, const int32_t *traj_off, const double *gt_pos, int32_t align,
                   double *rmse, double *R, double *t);

/* Local refinement after the relaxation (/root/reference/README.md:63-67: SCORE's estimate is
 * the initialisation of a local search — GTSAM in the paper).  Batched Levenberg-Marquardt on the original non-convex
 * cost (the reference's objective, gurobi_utils.py:358-526, with R_p in SO(d) and the distance variables eliminated:
 * ranges enter as w (||p_a - p_b|| - r)^2), first pose of every instance fixed; Gauss-Newton systems by block-Jacobi
 * PCG, matrix-free over the factor incidence lists.  Starts from the rounded solution of the last score_solve (rotations
 * rounded, translations and landmarks as relaxed), or from init_poses [P*d*(d+1)] ([R|t] row-major, R in SO(d)) and
 * init_landmarks [L*d] when both are given (host or device pointers).  Call after score_solve (the handle's factor
 * lists are built there).  No reference counterpart in /root/reference itself. */
typedef struct ScoreRefineParams {
  int32_t max_outer;  /* Levenberg-Marquardt iterations per instance, <=0: 100 */
  int32_t max_inner;  /* PCG iterations per Gauss-Newton system, <=0: 200 */
  double rel_tol;     /* stop when an accepted step lowers the cost by less than rel_tol (1 + cost), <=0: 1e-10 */
  double lambda0;     /* initial damping lambda of (J'WJ + lambda I), <=0: 1e-3 */
  double cg_tol;      /* PCG relative residual (preconditioned norm), <=0: 1e-4 (looser solves were seen to end in
                         other, worse local minima than an exact Levenberg-Marquardt iteration) */
  void *stream;       /* cudaStream_t, NULL: the handle's own stream */
  int32_t preconditioner; /* 0: block LDL^T along every odometry chain segment (default); 1: block-Jacobi (A/B, tests) */
  int32_t reserved;
} ScoreRefineParams;
typedef struct ScoreRefineStats {
  int32_t n_instances;
  int32_t n_converged;      /* instances that stopped by rel_tol (or at a stationary point), not by max_outer */
  int32_t outer_iterations; /* Levenberg-Marquardt iterations the batch ran (the slowest instance's) */
  int32_t kernel_launches;
  double refine_ms;
} ScoreRefineStats;
typedef struct ScoreRefineInstanceStats {
  double cost_initial, cost_final;
  int32_t outer_iterations, accepted_steps;
} ScoreRefineInstanceStats;
int score_refine(ScoreHandle h, const ScoreRefineParams *params, const double *init_poses, const double *init_landmarks,
                 ScoreRefineStats *stats, ScoreRefineInstanceStats *per_instance);
/* Refined estimate: poses [P*d*(d+1)] as [R|t] row-major with R in SO(d), landmarks [L*d] (host buffers). */
int score_get_refined(ScoreHandle h, double *poses, double *landmarks);

void score_destroy(ScoreHandle h);
/* Handles take their device memory, stream and events from process-wide caches that score_destroy refills (so
 * that sweeps of similar problems create and destroy handles without any driver allocation call); this returns
 * everything cached to the driver.  No reference counterpart. */
void score_release_cached(void);
const char *score_last_error(void);
const char *score_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SCORE_B200_H_ */
