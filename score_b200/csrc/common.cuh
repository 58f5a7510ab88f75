// Shared device-side structures and helpers for libscore_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>
#include <type_traits>

#include "../../include/score_b200.h"

namespace score {

constexpr int kRowsPerBlock = 768;  // divisible by 2 and 3 so a range's d rows never straddle blocks
constexpr int kColsPerBlock = 768;
constexpr int kThreads = 256;
constexpr int kSegThreads = 128;  // chain-scan CTA width
constexpr int kNumCand = 12;      // line-search candidates 2^(1 - c/2); slot kNumCand is a = 0
constexpr int kCoarseMax = 128;   // largest per-instance coarse space handled by the dense on-chip solve (4 x 4 register tiles)
constexpr int kCoarseThreads = 1024;
constexpr int kCoarseBigMax = 8192;  // largest coarse space of the dense global-memory path (single-instance handles)
constexpr int kLsSums = kNumCand + 1 + 3;

// Per-instance phase.  PH_WAIT: this Newton system is solved to its forcing tolerance; idle until the
// batch-wide line-search tick of the cycle.
enum Phase : int { PH_CG = 0, PH_LS = 1, PH_DONE = 2, PH_WAIT = 3 };
// Tick kind the host schedules (the same for every instance of the batch: the batch advances in lockstep).
enum TickMode : int { TM_LS = 0, TM_EVAL = 1, TM_CG = 2, TM_CG_LAST = 3, TM_FILE = 4 };
enum ColKind : int { CB_POSE = 0, CB_LANDMARK = 1 };

// Work descriptor of one CTA of a row-pass / column-pass kernel.  A CTA never
// spans two instances, so per-instance partial sums are reduced in a fixed order.
struct BlockDesc {
  int inst, i0, i1, kind;
};

// Per-instance solver state, owned by the controller kernels.
struct InstState {
  int phase, skip_ls, end_cg, solved;
  int newton_it, cg_it, total_cg, ls_fail;
  int eval_now, want_eval, n_eval, stall;  // true-KKT evaluation ticks
  double alpha, beta, rs, rs0, eta, step;
  int c_age, ls_shift;                     // Newton steps the current coarse inverse has served; line-search ladder shift
  int fused_cg, holds;                     // PCG iterations done inside the fused kernel (fused.cuh); stage ends at which
                                           // the current barrier parameter was held (ctrl_a)
  double mu_c;                             // barrier parameter the coarse inverse was built at
  double mu, mu_ls, dec;                   // barrier parameter (current / used by this tick's line search), Newton decrement
  double mu_out;                           // barrier parameter of the auxiliary variables the last certificate was taken with
  double dec_prev;                         // Newton decrement^2 of the previous step of this barrier stage (0: first step)
  double F, Fmu, kkt, r_stat, r_gap, gnorm, xnorm;
  // the mu-driven parts of the last certificate (sum lambda s in the gap, min(lambda, s) in the stationarity residual) and
  // the factor the forcing term has been tightened by while the barrier parameter was held (ctrl_a); 0 reads as 1
  double gap_mu, stat_mu, eta_scale;
};

struct DevProblem {
  int d, blk, rpe, npe;  // dim, d*(d+1), rows per edge, nnz per edge
  int relax, n_inst;
  int P, L, E, K, Lp, n_seg;
  int nz, m, nnz;  // reduced operator: columns, rows, non-zeros over the batch
  // instance offset tables [n_inst+1]
  int *pose_off, *lm_off, *edge_off, *rng_off, *prior_off;
  int *zoff, *roff, *nnzoff;
  int *seg_ptr, *seg_inst, *link_edge;
  int *seg_begin;  // [n_inst+1] first segment of each instance
  int4 *seg_tab;   // [n_inst x maxseg] {first pose, end pose, segment, 0} of segment j of an instance (first pose -1: none):
                   // what a chain-scan CTA needs about its work item in one 16-byte load (WorkLists.maxseg is the stride)
  // factors
  int *edge_i, *edge_j;
  double *edge_t, *edge_R, *edge_k, *edge_tau;
  int *rng_a, *rng_b;
  double *rng_dist, *rng_w;
  int *prior_l;
  double *prior_t, *prior_w;
  // reduced operator B (CSR) and its transpose
  int *indptr, *cols;
  double *vals;
  int *t_indptr, *t_rows;
  double *t_vals;
  double *w, *b;
  // odometry-chain preconditioner: G = dead-reckoned frame [Rg|tg] per pose (d x (d+1)),
  // M = G^-T D^-1 G^-1 per pose ((d+1) x (d+1)), landmark diagonal inverse
  double *G, *M, *lm_inv;
  // coarse level (free segment bases + landmarks of an instance): slot of every range endpoint,
  // per-instance offsets into the coarse vectors / matrices, size and on/off flag
  // static sorted lists the coarse matrix is summed from (coarse.cuh)
  int c_ninc, c_npair;
  int *c_inc_off, *c_inc_code;       // [n_inst+1] ; [c_ninc] (local range << 2) | endpoint (0: a, 1: b, 2: a - b)
  double *c_inc_h, *c_inc_w2;        // [c_ninc x (d+1)] frame of the endpoint ; [c_ninc] 2 w
  int *c_drun_off, *c_drun_slot, *c_drun_begin;  // diagonal runs: [n_inst+1] ; [n_drun] ; [n_drun+1]
  int *c_pr_off, *c_pr_code;         // [n_inst+1] ; [c_npair] (local range << 1) | (1: endpoint b is the lower slot)
  double *c_pr_h;                    // [c_npair x 2 (d+1)] frames of the lower / higher slot endpoint
  int *c_orun_lo, *c_orun_hi, *c_orun_begin;     // off-diagonal runs: [n_orun] x2 ; [n_orun+1]
  int *c_orun_off;                   // [n_inst+1] first off-diagonal run of every instance
  int *c_owarp;                      // [n_inst x 33] first off-diagonal run of every warp of the build CTA
  int *c_off, *c_moff, *c_n, *c_nb;  // [n_inst(+1)]
  // factor incidence lists of the matrix-free Hessian-vector product (hessvec.cuh): owner = global pose index or
  // P + global landmark index
  int n_inc;
  int *inc_ptr;                  // [P + L + 1]
  uint4 *inc_rec;                // [n_inc] incidence records (hessvec.cuh)
  int n_nonlink;                 // relative-pose factors that are not odometry links (loop closures)
  double *c_Ainv;                 // per instance nc x nc inverse coarse Hessian
  double *c_rhs, *c_sol;          // coarse right-hand side / solution
};

struct SolverVecs {
  // column space
  double *z, *dz, *r, *s, *p, *t, *ytmp;
  // row space
  double *res, *u, *bdz;
  double *mk;  // [K x d(d+1)/2] curvature block 2 w (tan I + (rad - tan) v v^T / n^2) of every range term (upper, row-major)
  // partial sums
  double *part_row;  // [n_row_blocks] pHp
  double *part_ls;   // [n_row_blocks * kLsSums]
  double *part_upd;  // [n_row_blocks * 4]  F ; evaluation ticks also |delta|^2, sum lambda s, sum min(lambda, s)^2
  double *hloc;      // [nz] B^T u of this rank's rows (row-partitioned solve: SpMV and update are separate passes)
  const double *hglob;  // [nz] its sum over the ranks
  double *part_col;  // [n_col_blocks * 4]  |g|^2, g.z, |z|^2
  double *part_seg;  // [n_seg] r.s over chain segments
  double *part_lm;   // [n_inst] r.s over landmarks
  // matrix-free PCG operator (hessvec.cuh): h = B^T H_r B p and the partial p'Hp of every pose / landmark block
  double *h;         // [nz]
  double *part_hv;   // [n_pose_blocks]
  double *part_gr;   // [n_pose_blocks * 4] certificate sums |g|^2, g.z, |z|^2 of the matrix-free gradient kernel
  int mf, pad1;      // 1: PCG iterations apply the operator factor by factor; 0: assembled CSR pair (row + column pass)
  // diagnostic trace (ScoreParams.verbose >= 2): per instance kTraceRec doubles per Newton step, trace_cap steps
  double *trace;
  int trace_cap;
};
constexpr int kTraceRec = 8;

// Work descriptor of one CTA of the matrix-free Hessian-vector kernel: poses / landmarks [i0, i1) of instance `inst`
// (instance-local), plus what the kernel would otherwise fetch through two more dependent loads.
struct HvBlock {
  int inst, i0, i1, kind;  // kind < 0: no such block (padding of the dense item table)
  int z0, pg0, Pi, lg0;    // column base, global index of the instance's first pose, its pose count, first global landmark
  int bid, jb, je, pad;    // index in the compact table (partial sums); pose blocks: the block's run of incidence records
};

struct BlockTables {
  BlockDesc *rb, *cb;  // row blocks, column blocks
  HvBlock *pb;         // pose / landmark blocks of the Hessian-vector kernel
  HvBlock *pbd;        // the same as a dense item table [n_inst x WorkLists.maxpb]: one 48-byte load per work item
  int n_rb, n_cb, n_pb;
  int *rb_begin, *cb_begin, *pb_begin;  // [n_inst+1]
};

// Compacted lists of the instances a tick has work for, rebuilt on the device by k_ctrl_b at the end of every
// tick (double-buffered by parity).  Worker kernels are launched with a fixed, SM-sized grid and loop over
// (listed instance) x (block of that instance), so finished / idle instances cost nothing.
//   run:  instances inside a Newton solve or about to take a line-search tick (PH_CG, PH_LS)
//   ls:   the PH_LS subset (line-search kernels)
//   wait: PH_WAIT instances (only the controller visits them, to wake them at the next line-search tick)
//   ev:   instances that asked for a certificate evaluation in this cycle's evaluation tick
enum ListKind : int { WL_RUN = 0, WL_LS = 1, WL_WAIT = 2, WL_EVAL = 3 };
struct WorkLists {
  int *par;     // [1] parity of the lists the workers read
  int *ticket;  // [1] completion ticket of the k_ctrl_b CTAs (the last one flips the parity)
  int *cnt;     // [2][3] entries of run / ls / wait per parity
  int *cnt_ev;  // [1]
  int *lists;   // [2][3][n_inst] run / ls / wait per parity, then [n_inst] ev   (flat: no dynamic indexing of
  int *ev;      //  kernel-parameter arrays, which would force a local-memory copy of the parameters)
  int n_inst;
  int maxvc;                         // most kVecChunk-column chunks any instance has (element-wise vector kernels)
  int maxrb, maxcb, maxseg, maxpb;  // most row blocks / column blocks / chain segments / pose blocks any instance has
  int rb_lo, rb_hi;          // row blocks this rank owns (row-partitioned multi-GPU solve; [0, n_rb) otherwise)
  __device__ __forceinline__ int *list(int parity, int kind) const { return lists + (size_t)(parity * 3 + kind) * n_inst; }
};
__device__ __forceinline__ void wl_get(const WorkLists &W, int kind, const int *&list, int &n) {
  if (kind == WL_EVAL) {
    list = W.ev;
    n = *W.cnt_ev;
    return;
  }
  const int p = *W.par;
  n = W.cnt[p * 3 + kind];
  list = W.list(p, kind);
}

struct SolverCfg {
  int max_newton, max_cg;
  double kkt_tol, forcing;
  double mu0, mu_factor, mu_min, mu_eval, center_tol, center_tol_late, coarse_reg;
  int coarse_every, pad;
};

__host__ __device__ inline int find_inst(const int *off, int n_inst, int idx) {
  // largest i with off[i] <= idx  (off has n_inst+1 entries, off[n_inst] > idx)
  int lo = 0, hi = n_inst;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (off[mid] <= idx)
      lo = mid;
    else
      hi = mid;
  }
  return lo;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Deterministic block sum (fixed tree); result valid in thread 0.
template <int NT>
__device__ __forceinline__ double block_sum(double v, double *smem /* >= NT/32 */) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) smem[wid] = v;
  __syncthreads();
  double out = 0.0;
  if (wid == 0) {
    out = (lane < NT / 32) ? smem[lane] : 0.0;
    out = warp_sum(out);
  }
  return out;
}

// Sixteen warp sums at once: every butterfly step halves the number of values a lane carries (the two halves
// of the warp keep different values), 16 shuffles instead of 80.  Lane l returns the total of v[(l >> 1) & 15].
// Fixed order, so bit-reproducible.
__device__ __forceinline__ double warp_sum16(double (&v)[16]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int h = 8, o = 16; h >= 1; h >>= 1, o >>= 1) {
    const bool up = lane & o;
#pragma unroll
    for (int i = 0; i < h; ++i) {
      const double send = up ? v[i] : v[i + h], keep = up ? v[i + h] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// Sixteen deterministic block sums with two barriers: total of v[i] is returned in thread i (i < 16), 0 elsewhere.
template <int NT>
__device__ __forceinline__ double block_sum16(double (&v)[16], double *smem /* >= 16 * NT/32 */) {
  constexpr int NW = NT / 32;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const double mine = warp_sum16(v);
  __syncthreads();
  if (!(lane & 1)) smem[(lane >> 1) * NW + wid] = mine;
  __syncthreads();
  double out = 0.0;
  if (threadIdx.x < 16) {
#pragma unroll
    for (int w = 0; w < NW; ++w) out += smem[threadIdx.x * NW + w];
  }
  return out;
}

}  // namespace score

extern thread_local std::string g_score_last_error;

#define SCORE_CUDA_CHECK(expr)                                                                   \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess) {                                                                     \
      g_score_last_error = std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + \
                           ":" + std::to_string(__LINE__) + ")";                                 \
      return SCORE_ERR_CUDA;                                                                     \
    }                                                                                            \
  } while (0)
