// Fused PCG solve: one thread-block cluster per instance runs whole Newton-system solves (every PCG iteration
// until the forcing tolerance is met) inside ONE kernel, with cluster barriers where the tick kernels of solver.cuh /
// precond.cuh have kernel boundaries.
//
// Why: a PCG tick of the lockstep scheduler is seven dependent kernels.  When only a few instances are still
// running — the last cycles of a sweep (a handful of ill-conditioned instances need 5-10x the median number of PCG
// iterations), or a handle that holds a single graph — every kernel is nearly empty and the tick costs seven
// launch + dependency latencies (~58 us measured) while the arithmetic takes a few microseconds.  Here the seven
// phases of an iteration are separated by hardware cluster barriers (barrier.cluster, release / acquire at cluster
// scope) and an instance's data stays in L2 between them.
//
// The phases call the SAME body functions as the stand-alone kernels, with the same block decomposition (row blocks,
// column blocks, chain segments) and therefore the same fixed-order partial sums: an instance solved here is
// bit-identical to the same instance advanced by lockstep ticks.  The bodies are instantiated with coherent loads
// (the vectors change between phases; ld.global.nc would be allowed to return stale lines) and the chain scans run
// on the first kSegThreads threads of the CTA behind a named barrier.
#pragma once
#include <cooperative_groups.h>

#include "precond.cuh"
#include "solver.cuh"

namespace score {

constexpr int kClusterSize = 8;  // CTAs per instance (portable maximum)
constexpr int kGroupSize = 32;   // CTAs per instance of the software-barrier variant

// Barrier of the `nb` CTAs that work on one instance: arrival counter + generation in global memory.  Thread 0 of every
// CTA releases its CTA's writes (__threadfence), arrives, spins on the generation, acquires.  All CTAs of a group must
// be resident at the same time (the launch keeps the grid below the device's capacity).
__device__ __forceinline__ void group_barrier(int *cnt, volatile int *gen, int nb) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const int g = *gen;
    if (atomicAdd(cnt, 1) == nb - 1) {
      *cnt = 0;
      __threadfence();
      *gen = g + 1;
    } else {
      while (*gen == g) {
      }
    }
    __threadfence();
  }
  __syncthreads();
}

// BAR = 0: one thread-block cluster per instance, hardware cluster barriers.  BAR = 1: kGroupSize CTAs per instance,
// software barriers through global memory (more CTAs per instance: every phase is a single round of blocks).
template <int D, int BAR>
__global__ void __launch_bounds__(kThreads) k_pcg_fused(DevProblem P, SolverVecs V, BlockTables T, InstState *st, SolverCfg cfg,
                                                       int *n_done, WorkLists W, int max_iters, int *bar_mem) {
  namespace cg = cooperative_groups;
  constexpr int NB = BAR ? kGroupSize : kClusterSize;
  const int crank = BAR ? (int)(blockIdx.x % NB) : (int)cg::this_cluster().block_rank();
  const int cid = blockIdx.x / NB, ncl = gridDim.x / NB;
  auto sync_all = [&]() {
    if (BAR)
      group_barrier(bar_mem + 2 * cid, bar_mem + 2 * cid + 1, NB);
    else
      cg::this_cluster().sync();
  };
  const int *act;
  int n_act;
  wl_get(W, WL_RUN, act, n_act);
  if (blockIdx.x == 0 && threadIdx.x == 0) {  // the lists k_ctrl_b(TM_FILE) will fill after this kernel
    const int q = (*W.par) ^ 1;
    W.cnt[q * 3 + 0] = W.cnt[q * 3 + 1] = W.cnt[q * 3 + 2] = 0;
  }
  const bool warp0 = threadIdx.x < 32;
  for (int ai = cid; ai < n_act; ai += ncl) {
    const int inst = act[ai];
    const int rb0 = T.rb_begin[inst], rb1 = T.rb_begin[inst + 1];
    const int cb0 = T.cb_begin[inst], cb1 = T.cb_begin[inst + 1];
    const int sg0 = P.seg_begin[inst], nseg = P.seg_begin[inst + 1] - sg0;
    for (int it = 0; it < max_iters; ++it) {
      // phase written by CTA 0 before the last barrier of the previous iteration: the same value in every CTA
      if (st[inst].phase != PH_CG || st[inst].eval_now) break;
      // ---- q = B p, u = H_r q, partial p'Hp
      for (int bid = rb0 + crank; bid < rb1; bid += NB) {
        rowpass_body<D, false>(P, V, T, st, bid);
        __syncthreads();
      }
      sync_all();
      // ---- step length
      if (crank == 0 && warp0) ctrl_a_body(V, T, st, cfg, inst);
      sync_all();
      // ---- h = B^T u, dz += alpha p, r -= alpha h
      for (int bid = cb0 + crank; bid < cb1; bid += NB) {
        colpass_body<false>(P, V, T, st, TM_CG, CS_FUSED, bid);
        __syncthreads();
      }
      sync_all();
      // ---- s = P r: reverse scans (+ coarse right-hand side), then coarse solve + forward scans
      if (threadIdx.x < kSegThreads)
        for (int j = crank; j <= nseg; j += NB) {
          if (j < nseg)
            precond_rev_body<D, true>(P, V, st, sg0 + j, inst, P.seg_ptr[sg0 + j], P.seg_ptr[sg0 + j + 1]);
          else
            precond_rev_body<D, true>(P, V, st, P.n_seg + inst, inst, 0, 0);
          seg_bar<true>();
        }
      sync_all();
      if (threadIdx.x < kSegThreads)
        for (int j = crank; j < nseg; j += NB) {
          precond_fwd_body<D, true>(P, V, st, sg0 + j, inst, P.seg_ptr[sg0 + j], P.seg_ptr[sg0 + j + 1], true);
          seg_bar<true>();
        }
      sync_all();
      // ---- r.s, beta, convergence
      if (crank == 0 && warp0) {
        ctrl_b_body(P, V, T, st, cfg, n_done, TM_CG, inst);
        if (threadIdx.x == 0) st[inst].fused_cg += 1;
      }
      sync_all();
      // ---- p = s + beta p
      for (int bid = cb0 + crank; bid < cb1; bid += NB) pupdate_body(V, T, st, bid);
      sync_all();
    }
  }
}

}  // namespace score
