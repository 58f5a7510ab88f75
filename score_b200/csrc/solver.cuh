// Solver tick kernels.
//
// The convex program the reference hands to Gurobi's barrier (score/solve_score.py:76; constraints
// score/utils/gurobi_utils.py:316-352, objective :358-526) is solved here by a primal interior-point
// method in its cone-eliminated form.  For a barrier parameter mu the auxiliary distance variable of
// every range is minimised out exactly (SURVEY.md App. A.4 is the mu = 0 case):
//     phi_mu(n) = min_{0<=rho<1}  w (n - r rho)^2 - mu log(1 - rho^2),      n = ||t_a - t_b||,
//     F_mu(z)   = sum_edges w (B z)^2 + sum_ranges phi_mu(||t_a - t_b||) + priors,   z = (R, t, l),
// which is smooth and convex; its minimisers trace the central path of the log-barrier formulation of
// the QCQP (ball constraint ||delta|| <= 1) — the same path an interior-point solver follows — and
// F_0 is the reference objective with delta (QCQP) / the cone variable (SOCP) at their optimal values.
// Each F_mu is minimised by Newton steps; every Newton system is solved by conjugate gradients with
// the two-level preconditioner of precond.cuh, and a one-pass multi-candidate line search picks the
// step.  mu shrinks whenever the Newton decrement says the iterate is centred.
//
// The batch advances in lockstep cycles scheduled by the host (api.cu): one line-search tick (TM_LS: step
// along the Newton direction, gradient + curvature + coarse matrix at the new point), one evaluation tick
// (TM_EVAL: certificate of the un-smoothed problem, only for instances that asked for it) and n_cg PCG ticks
// (TM_CG).  An instance whose PCG reaches its forcing tolerance idles (PH_WAIT) until the next line-search
// tick; one whose PCG needs longer simply keeps iterating through it (the line-search tick's kernel sequence
// contains the PCG one) and takes a later opportunity.  The expensive line-search kernels therefore run only
// once per cycle, at high batch occupancy, and PCG ticks launch nothing else.  All scalars live on the device; the host only replays CUDA graphs of whole cycles.
#pragma once
#include "common.cuh"

#ifndef SCORE_LS_MINB
#define SCORE_LS_MINB 3
#endif
#ifndef SCORE_RU_MINB
#define SCORE_RU_MINB 4
#endif

namespace score {

__device__ __forceinline__ double ls_candidate(int c) {
  // 2^(1 - c/2), c = 0..kNumCand-1 (compile-time constants in the unrolled loops)
  return ldexp((c & 1) ? 1.4142135623730951 : 1.0, 1 - (c + 1) / 2);
}
// After a line search without any decrease the same direction is tried again on a ladder of 64x shorter steps
// (shift 1), then 4096x (shift 2), ...
__device__ __forceinline__ double ls_scale(int shift) { return ldexp(1.0, -6 * shift); }

// 1 / x from the hardware seed (rcp.approx.ftz.f64, ~23 bits) and one Newton step (~46 bits): a handful of instructions
// where the IEEE quotient is 20-30.  Used where the quotient feeds an iteration that corrects itself.
__device__ __forceinline__ double rcp_fast(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  return fma(fma(-x, r, 1.0), r, r);
}

// eps = 1 - rho*, the root in (0,1] of (q - 1 + e) e (2 - e) = kap (1 - e)   (stationarity of phi_mu in rho
// with q = n / r, kap = mu / (w r^2)); parametrised by eps so that rho -> 1 keeps full relative accuracy.
__device__ __forceinline__ double barrier_eps(double q, double kap) {
  const double qm = q - 1.0;
  const double sq = sqrt(qm * qm + 2.0 * kap);
  double e = (qm >= 0.0) ? kap / (qm + sq) : 0.5 * (sq - qm);
  e = fmin(e, 1.0);
  for (int it = 0; it < 6; ++it) {  // quadratic convergence; most ranges need 1-3 steps
    const double F = (qm + e) * e * (2.0 - e) - kap * (1.0 - e);
    const double dF = e * (2.0 - e) + (qm + e) * (2.0 - 2.0 * e) + kap;
    double ne = fma(-F, rcp_fast(dF), e);  // (the iteration's fixed point is the root whatever the quotient's last bits)
    if (!(ne > 0.0)) ne = 0.5 * e;
    ne = fmin(ne, 1.0);
    const bool done = fabs(ne - e) <= 2e-16 * e;
    e = ne;
    if (done) break;
  }
  return e;
}

// The same root for G step sizes of one range at once (line search): the G Newton iterations are independent, so
// running them in lockstep gives the scheduler G divisions to interleave instead of one dependent chain.  `tol` is
// the relative step at which the iteration stops: phi_mu is the MINIMUM over rho, so its value is second-order
// insensitive to the root (envelope) and, Newton converging quadratically, a last step below 1e-7 leaves the value
// exact to rounding — the confirming iteration barrier_eps needs for the curvature terms is not needed here.
template <int G>
__device__ __forceinline__ void barrier_eps_group(const double (&qm)[G], double kap, double tol, double (&e)[G]) {
#pragma unroll
  for (int g = 0; g < G; ++g) {
    // starting point from the two asymptotes; single-precision square root and fast reciprocal are plenty for a guess
    const double sq = (double)sqrtf((float)(qm[g] * qm[g] + 2.0 * kap));
    double e0 = (qm[g] >= 0.0) ? kap * rcp_fast(qm[g] + sq) : 0.5 * (sq - qm[g]);
    if (!(e0 > 0.0)) e0 = 0.5 * kap;  // (the rounded square root can cancel qm exactly)
    e[g] = fmin(e0, 1.0);
  }
  for (int it = 0; it < 6; ++it) {
    bool all = true;
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const double eg = e[g];
      const double F = (qm[g] + eg) * eg * (2.0 - eg) - kap * (1.0 - eg);
      const double dF = eg * (2.0 - eg) + (qm[g] + eg) * (2.0 - 2.0 * eg) + kap;
      // Newton step with the fast reciprocal: the kernel is bound by FP64 instruction issue, the quotient was most of
      // an iteration, and a step that is exact to 1e-13 relative converges exactly like the exact one
      double ne = fma(-F, rcp_fast(dF), eg);
      if (!(ne > 0.0)) ne = 0.5 * eg;
      ne = fmin(ne, 1.0);
      all = all && (fabs(ne - eg) <= tol * eg);
      e[g] = ne;
    }
    if (all) break;
  }
}

// One range term at distance n: value phi_mu(n), tan = phi'/(2 w n), rad = phi''/(2 w).
struct RangeTerm {
  double val, tan, rad;
  double rho, gapr;  // norm of the auxiliary variable delta = rho v / n and the residual n - r rho it leaves
};
__device__ __forceinline__ RangeTerm range_term(double n, double r, double w, double mu, bool need_val) {
  RangeTerm t;
  if (!(r > 0.0)) {  // dist == 0: plain quadratic w n^2 (the delta column is all zeros)
    t.val = w * n * n;
    t.tan = 1.0;
    t.rad = 1.0;
    t.rho = 0.0;
    t.gapr = n;
    return t;
  }
  if (mu == 0.0) {  // exact elimination: w max(0, n - r)^2
    const double e = n - r;
    const bool act = e > 0.0;
    t.val = act ? w * e * e : 0.0;
    t.tan = act ? e / n : 0.0;
    t.rad = act ? 1.0 : 0.0;
    t.rho = act ? 1.0 : n / r;
    t.gapr = act ? e : 0.0;
    return t;
  }
  const double q = n / r, kap = mu / (w * r * r);
  const double e = barrier_eps(q, kap);
  const double rho = 1.0 - e, om = e * (2.0 - e);
  const double c = kap * (1.0 + rho * rho) / (om * om);
  t.rad = c / (1.0 + c);
  t.tan = (q > 0.0) ? (q - 1.0 + e) / q : t.rad;
  t.rho = rho;
  t.gapr = r * (q - 1.0 + e);
  t.val = 0.0;
  if (need_val) t.val = w * t.gapr * t.gapr - mu * log(om);
  return t;
}

// Value only (line search): phi_mu(n).  (Warm-starting the root search from a neighbouring step size was tried
// and rejected: the stationarity cubic has other real roots and Newton then occasionally lands on one of them.)
__device__ __forceinline__ double range_value(double n, double r, double w, double mu) {
  if (!(r > 0.0)) return w * n * n;
  if (mu == 0.0) {
    const double e = fmax(n - r, 0.0);
    return w * e * e;
  }
  const double q = n / r, kap = mu / (w * r * r);
  const double e = barrier_eps(q, kap);
  const double gap = r * (q - 1.0 + e);
  return w * gap * gap - mu * log(e * (2.0 - e));
}

// res = B z - b   (initialisation only)
__global__ void __launch_bounds__(kThreads) k_residual(DevProblem P, const double *__restrict__ z, double *res) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= P.m) return;
  double acc = -P.b[row];
  for (int k = P.indptr[row]; k < P.indptr[row + 1]; ++k) acc += P.vals[k] * z[P.cols[k]];
  res[row] = acc;
}

// ---- K1: q = B x.  PH_CG: x = p, u = 2 W H_r q (H_r = per-range curvature block tan I + (rad - tan) vv^T/n^2 of
// F_mu at the current residual, stored as M_k = 2 w H_r by k_rowupdate), partial p'Hp.  PH_LS: x = dz, bdz = q.
// NC: gathers of the solver vectors may use the non-coherent path (stand-alone kernels: the vectors are constant for
// the kernel's lifetime).  The fused PCG kernel (fused.cuh) rewrites them between its phases and needs coherent loads.
template <bool NC>
__device__ __forceinline__ double ldv(const double *p) {
  return NC ? __ldg(p) : *p;
}

template <int D, bool NC = true>
__device__ __forceinline__ void rowpass_body(DevProblem P, SolverVecs V, BlockTables T, const InstState *st, const int bid) {
  __shared__ double sq[kRowsPerBlock];
  __shared__ double red[kThreads / 32];
  const BlockDesc bd = T.rb[bid];
  const int phase = st[bd.inst].phase;
  if ((phase != PH_CG && phase != PH_LS) || st[bd.inst].eval_now) return;
  if (phase == PH_CG && V.mf) return;                 // PCG iterations apply the operator matrix-free (k_hessvec)
  if (phase == PH_LS && st[bd.inst].skip_ls) return;  // start point: no direction yet (bdz stays 0)
  const double *x = (phase == PH_CG) ? V.p : V.dz;
  const int nrows = bd.i1 - bd.i0;
  {
    // rows li = tid + t * kThreads, t < 3 (kRowsPerBlock = 3 kThreads): the index loads of all three rows are
    // issued before the value / gather loads so that each thread keeps several requests in flight
    constexpr int RPT = kRowsPerBlock / kThreads;
    int k0[RPT], k1[RPT];
#pragma unroll
    for (int t = 0; t < RPT; ++t) {
      const int li = threadIdx.x + t * kThreads;
      const bool ok = li < nrows;
      k0[t] = ok ? P.indptr[bd.i0 + li] : 0;
      k1[t] = ok ? P.indptr[bd.i0 + li + 1] : 0;
    }
    double acc[RPT];
#pragma unroll
    for (int t = 0; t < RPT; ++t) acc[t] = 0.0;
#pragma unroll
    for (int j = 0; j < 6; ++j) {  // rows of the reduced operator hold at most d + 2 <= 5 entries
#pragma unroll
      for (int t = 0; t < RPT; ++t) {
        const int k = k0[t] + j;
        if (k < k1[t]) acc[t] += P.vals[k] * ldv<NC>(x + P.cols[k]);
      }
    }
#pragma unroll
    for (int t = 0; t < RPT; ++t) {
      for (int k = k0[t] + 6; k < k1[t]; ++k) acc[t] += P.vals[k] * ldv<NC>(x + P.cols[k]);  // (never taken today)
      const int li = threadIdx.x + t * kThreads;
      if (li < nrows) sq[li] = acc[t];
    }
  }
  __syncthreads();
  if (phase == PH_LS) {
    for (int li = threadIdx.x; li < nrows; li += kThreads) V.bdz[bd.i0 + li] = sq[li];
    return;
  }
  const int inst = bd.inst;
  const int rr0 = P.roff[inst] + (P.edge_off[inst + 1] - P.edge_off[inst]) * P.rpe;
  const int rr1 = rr0 + (P.rng_off[inst + 1] - P.rng_off[inst]) * D;
  double acc = 0.0;
  for (int li = threadIdx.x; li < nrows; li += kThreads) {
    const int row = bd.i0 + li;
    const double q = sq[li];
    double u;
    if (row >= rr0 && row < rr1) {
      // range rows: u = M_k q  (M_k already carries 2 w)
      const int rel = row - rr0, comp = rel % D, base = row - comp - bd.i0;
      const double *mk = V.mk + (size_t)(P.rng_off[inst] + rel / D) * (D * (D + 1) / 2);
      u = 0.0;
#pragma unroll
      for (int c = 0; c < D; ++c) {
        const int a = comp < c ? comp : c, b = comp < c ? c : comp;
        u += mk[a * D - a * (a - 1) / 2 + (b - a)] * sq[base + c];
      }
    } else {
      u = 2.0 * P.w[row] * q;
    }
    V.u[row] = u;
    acc += q * u;
  }
  const double tot = block_sum<kThreads>(acc, red);
  if (threadIdx.x == 0) V.part_row[bid] = tot;
}

template <int D>
__global__ void __launch_bounds__(kThreads) k_rowpass(DevProblem P, SolverVecs V, BlockTables T, const InstState *st, WorkLists W) {
  const int *act;
  int n_act;
  wl_get(W, WL_RUN, act, n_act);
  for (long long item = blockIdx.x; item < (long long)n_act * W.maxrb; item += gridDim.x) {
    const int inst = act[item / W.maxrb], bid = T.rb_begin[inst] + (int)(item % W.maxrb);
    if (bid < T.rb_begin[inst + 1] && bid >= W.rb_lo && bid < W.rb_hi) rowpass_body<D>(P, V, T, st, bid);
  }
}

// ---- K_ls: F_mu(res + a bdz) at kNumCand step sizes plus a = 0, in one pass.
// Plain rows are quadratic in a (three sums); range rows are evaluated per candidate.
template <int D>
__device__ __forceinline__ void linesearch_body(DevProblem P, SolverVecs V, BlockTables T, const InstState *st, const int bid) {
  __shared__ double red[16 * (kThreads / 32)];
  const BlockDesc bd = T.rb[bid];
  const int inst = bd.inst;
  if (st[inst].phase != PH_LS || st[inst].skip_ls || st[inst].eval_now) return;
  const double mu = st[inst].mu;
  const double lsc = ls_scale(st[inst].ls_shift);
  const int rr0 = P.roff[inst] + (P.edge_off[inst + 1] - P.edge_off[inst]) * P.rpe;
  const int rr1 = rr0 + (P.rng_off[inst + 1] - P.rng_off[inst]) * D;
  double sums[3] = {0.0, 0.0, 0.0};
  // plain (quadratic) rows of this block: relative-pose rows before the ranges, prior rows after them
  auto plain = [&](int r0, int r1) {
    for (int row = r0 + threadIdx.x; row < r1; row += kThreads) {
      const double v = V.res[row], q = V.bdz[row], wr = P.w[row];
      sums[0] += wr * v * v;
      sums[1] += wr * v * q;
      sums[2] += wr * q * q;
    }
  };
  plain(bd.i0, min(bd.i1, rr0));
  plain(max(bd.i0, rr1), bd.i1);
  // range terms: the kNumCand + 1 evaluation points of a range (slot kNumCand: a = 0) are split into two groups of
  // GS, handled by an even / odd pair of threads — 2 x ranges work items per block, an equal number per thread
  // (a block never splits a range: kRowsPerBlock is a multiple of D).  The odd group's last slot is idle padding.
  constexpr int GS = 7;
  static_assert(kNumCand == 12 && kLsSums == 16, "group layout: even threads c = 0..6, odd threads c = 7..11 and a = 0");
  const int grp = threadIdx.x & 1;
  double acc[GS];
#pragma unroll
  for (int g = 0; g < GS; ++g) acc[g] = 0.0;
  const int ra = max(bd.i0, rr0), rb = min(bd.i1, rr1);
  if (rb > ra) {
    const int nrng = (rb - ra) / D, kfirst = P.rng_off[inst] + (ra - rr0) / D;
    for (int it = threadIdx.x; it < 2 * nrng; it += kThreads) {  // it & 1 == grp (kThreads is even)
      const int q = it >> 1, row = ra + q * D, k = kfirst + q;
      const double rr = P.rng_dist[k], wk = P.rng_w[k];
      double A = 0.0, Bq = 0.0, C = 0.0;
#pragma unroll
      for (int c = 0; c < D; ++c) {
        const double v = V.res[row + c], qq = V.bdz[row + c];
        A += v * v;
        Bq += v * qq;
        C += qq * qq;
      }
      Bq *= lsc;
      C *= lsc * lsc;
      double nn[GS];  // distance at this group's step sizes (compile-time constants; the ladder scale is in Bq, C)
#pragma unroll
      for (int g = 0; g < GS; ++g) {
        const double a0 = ls_candidate(g), a1 = (GS + g < kNumCand) ? ls_candidate(GS + g) : 0.0;
        const double a = grp ? a1 : a0;
        nn[g] = sqrt(fmax(0.0, A + a * (2.0 * Bq + a * C)));
      }
      if (!(rr > 0.0) || mu == 0.0) {  // dist == 0 / no barrier: closed forms, no root to find
#pragma unroll
        for (int g = 0; g < GS; ++g) acc[g] += range_value(nn[g], rr, wk, mu);
        continue;
      }
      const double rinv = 1.0 / rr, wr2 = wk * rr * rr, kap = mu / wr2;
      double qm[GS], e[GS];
#pragma unroll
      for (int g = 0; g < GS; ++g) qm[g] = fma(nn[g], rinv, -1.0);
      barrier_eps_group<GS>(qm, kap, 1e-7, e);
#pragma unroll
      for (int g = 0; g < GS; ++g) {
        const double gap = qm[g] + e[g];
        acc[g] += wr2 * gap * gap - mu * log(e[g] * (2.0 - e[g]));
      }
    }
  }
  double v[kLsSums];
#pragma unroll
  for (int i = 0; i < 3; ++i) v[i] = sums[i];
#pragma unroll
  for (int g = 0; g < GS; ++g) v[3 + g] = grp ? 0.0 : acc[g];
#pragma unroll
  for (int g = 0; g < kLsSums - 3 - GS; ++g) v[3 + GS + g] = grp ? acc[g] : 0.0;
  const double tot = block_sum16<kThreads>(v, red);
  if (threadIdx.x < kLsSums) V.part_ls[(size_t)bid * kLsSums + threadIdx.x] = tot;
}

template <int D>
__global__ void __launch_bounds__(kThreads, SCORE_LS_MINB) k_linesearch(DevProblem P, SolverVecs V, BlockTables T, const InstState *st, WorkLists W) {
  const int *act;
  int n_act;
  wl_get(W, WL_LS, act, n_act);
  for (long long item = blockIdx.x; item < (long long)n_act * W.maxrb; item += gridDim.x) {
    const int inst = act[item / W.maxrb], bid = T.rb_begin[inst] + (int)(item % W.maxrb);
    if (bid < T.rb_begin[inst + 1] && bid >= W.rb_lo && bid < W.rb_hi) linesearch_body<D>(P, V, T, st, bid);
  }
}

// Fixed-order sum of part[i0..i1) with stride `stride`, by one warp (the controllers run one warp per
// instance); valid in every lane.
__device__ __forceinline__ double ctrl_sum(const double *part, int i0, int i1, int stride, int off) {
  double acc = 0.0;
  for (int i = i0 + (threadIdx.x & 31); i < i1; i += 32) acc += part[(size_t)i * stride + off];
  return warp_sum(acc);
}

// PCG step length rs / p'Hp (0: breakdown, the solve ends)
__device__ __forceinline__ double cg_step_length(double pHp, double rs) {
  return (pHp > 0.0 && pHp > 1e-30 * fabs(rs) && isfinite(pHp)) ? rs / pHp : 0.0;
}

// ---- ctrl_a: PCG step length / line-search decision / barrier update.  One warp per listed instance.
__device__ __forceinline__ void ctrl_a_body(SolverVecs V, BlockTables T, InstState *st, SolverCfg cfg, const int inst) {
  const int lane = threadIdx.x & 31;
  InstState &S = st[inst];
  const int phase = S.phase;
  if ((phase != PH_CG && phase != PH_LS) || S.eval_now) return;
  const int b0 = T.rb_begin[inst], b1 = T.rb_begin[inst + 1];
  if (phase == PH_CG) {
    const double pHp = V.mf ? ctrl_sum(V.part_hv, T.pb_begin[inst], T.pb_begin[inst + 1], 1, 0) : ctrl_sum(V.part_row, b0, b1, 1, 0);
    if (lane == 0) {
      const double alpha = cg_step_length(pHp, S.rs);
      S.alpha = alpha;
      if (alpha != 0.0)
        S.dec += alpha * S.rs;  // -g.dz accumulates: Newton decrement^2 of the current solve
      else
        S.end_cg = 1;
    }
    return;
  }
  // PH_LS
  if (S.skip_ls) {
    if (lane == 0) {
      S.step = 0.0;
      S.mu_ls = S.mu;
    }
    return;
  }
  double tot[kLsSums];
  for (int i = 0; i < kLsSums; ++i) tot[i] = ctrl_sum(V.part_ls, b0, b1, kLsSums, i);
  if (lane == 0) {
    double best = tot[0] + tot[3 + kNumCand], step = 0.0;
    for (int c = 0; c < kNumCand; ++c) {
      const double a = ls_candidate(c) * ls_scale(S.ls_shift);
      const double Fc = tot[0] + a * (2.0 * tot[1] + a * tot[2]) + tot[3 + c];
      if (Fc < best) {
        best = Fc;
        step = a;
      }
    }
    // Below the resolution of the line search: when the predicted decrease (half the Newton decrement^2) is lost in the
    // rounding of F_mu (~1e-13 relative: a sum of ~1e4 terms), no candidate can be told from a = 0.  The Newton step
    // itself is still good to many digits there, so take it in full instead of reporting a failed search.
    if (step == 0.0 && S.ls_shift == 0 && S.dec >= 0.0 && S.dec <= 1e-11 * (1.0 + fabs(best))) step = 1.0;
    S.step = step;
    S.mu_ls = S.mu;
    // path following: once the Newton decrement (in units of mu) says the iterate is centred — or no
    // candidate improved — the barrier parameter shrinks; the gradient of this tick already uses it
    const double lam2 = (S.mu > 0.0) ? S.dec / S.mu : 0.0;
    S.want_eval = 0;
    // no decrease although the iterate is not centred: the Newton step is too long for the ladder -> solve the same
    // system again (tighter forcing term, ctrl_b) and search 64x shorter steps; give up on the stage after 3 shifts
    // (the early stages only have to stay near the path; the last ones feed the certificate and are centred tighter)
    const double ctol = (S.mu <= cfg.mu_eval) ? cfg.center_tol_late : cfg.center_tol;
    const bool retry = step == 0.0 && lam2 > ctol && S.ls_shift < 3;
    S.ls_shift = retry ? S.ls_shift + 1 : 0;
    // a decrement that has stopped falling near the resolution of F_mu cannot be driven lower: the stage is as centred
    // as it gets.  (Judged by stagnation, not by a fixed level: the certificate's gap term g.z is bounded by
    // sqrt(decrement * z'Hz), so every further decade the Newton iteration still delivers is needed.)
    const bool at_floor = S.mu <= cfg.mu_eval && S.dec <= 1e-9 * (1.0 + fabs(best)) && S.dec_prev > 0.0 && S.dec > 0.5 * S.dec_prev;
    const double mu_before = S.mu;
    if (S.mu > 0.0 && !retry && (lam2 <= ctol || at_floor || step == 0.0)) {
      if (S.mu <= cfg.mu_eval) S.want_eval = 1;
      // The certificate's mu-driven parts (sum lambda s in the gap, min(lambda, s) in the stationarity residual) scale
      // with mu (sqrt(mu)).  Once their estimate at the CURRENT mu — from the last evaluation, taken at mu_out — is a
      // quarter of the tolerance, what the certificate still lacks is Newton accuracy (its gap term g.z is bounded by
      // sqrt(decrement * z'Hz)), not a smaller mu: a smaller mu only sharpens the kinks of weakly active range terms and
      // the inexact Newton iteration stalls on them (sweep instance 4136: decrement stuck at 1e-10 from mu = 1e-10 down
      // to 1e-16).  So hold mu and tighten the forcing term instead — for up to three stage ends per level: where the
      // decrement stalls on the kinks themselves (instance 462 at mu = 1e-9) only a smaller mu helps, and the iteration
      // moves on with the tightened forcing term.
      const double ratio = (S.mu_out > 0.0) ? S.mu / S.mu_out : 1.0;
      const bool hold = S.n_eval > 0 && S.mu <= cfg.mu_eval && ratio <= 1.0 && S.gap_mu * ratio <= 0.25 * cfg.kkt_tol &&
                        S.stat_mu * sqrt(ratio) <= 0.25 * cfg.kkt_tol && S.holds < 3;
      if (hold) {
        S.holds += 1;
        S.eta_scale = fmax(0.3 * (S.eta_scale > 0.0 ? S.eta_scale : 1.0), 1e-2);
      } else {
        if (S.mu <= cfg.mu_min) S.stall += 1;
        S.mu = fmax(S.mu * cfg.mu_factor, cfg.mu_min);
        S.holds = 0;
      }
    }
    S.dec_prev = (S.mu != mu_before) ? 0.0 : S.dec;
    if (S.mu == 0.0) S.want_eval = 1;
    // long late stages: look at the true certificate every 4th Newton step as well
    if (S.mu <= cfg.mu_eval && (S.newton_it & 3) == 3) S.want_eval = 1;
    if (step == 0.0) S.ls_fail += 1;
  }
}

__global__ void __launch_bounds__(kSegThreads) k_ctrl_a(SolverVecs V, BlockTables T, InstState *st, SolverCfg cfg, WorkLists W,
                                                       int mode) {
  const int *act;
  int n_act;
  wl_get(W, WL_RUN, act, n_act);
  const int nw = kSegThreads / 32;
  for (int ai = blockIdx.x * nw + (threadIdx.x >> 5); ai < n_act; ai += gridDim.x * nw) ctrl_a_body(V, T, st, cfg, act[ai]);
  if (blockIdx.x == 0 && threadIdx.x == 0) {  // the lists k_ctrl_b of this tick will fill
    const int q = (*W.par) ^ 1;
    W.cnt[q * 3 + 0] = W.cnt[q * 3 + 1] = W.cnt[q * 3 + 2] = 0;
    if (mode == TM_LS) *W.cnt_ev = 0;
  }
}

// ---- K_upd (PH_LS): res += step * bdz ; per-range curvature factors and u = dF_mu/d(res) ; partial F_mu.
// Evaluation ticks (eval_now): u and sums of the un-smoothed problem (mu = 0) at the current point.
template <int D>
__device__ __forceinline__ void rowupdate_body(DevProblem P, SolverVecs V, BlockTables T, const InstState *st,
                                                        int mode, const int bid) {
  __shared__ double red[kThreads / 32];
  const BlockDesc bd = T.rb[bid];
  const int inst = bd.inst;
  const bool eval = mode == TM_EVAL;
  if (st[inst].phase == PH_DONE) return;
  if (eval ? !st[inst].eval_now : (st[inst].phase != PH_LS || st[inst].eval_now)) return;
  const double step = eval ? 0.0 : st[inst].step;
  // Evaluation ticks certify the point x = (z, delta) with delta on the central path of the barrier parameter the
  // iterate was last centred for (mu_ls): any delta in the unit ball is feasible, and with this one the stationarity
  // residual in z is the gradient the Newton iteration drives to zero, while the complementarity terms are bounded by
  // mu.  (With the exact minimiser delta = proj(v / r), mu = 0, weakly active ranges keep the residual of order
  // sqrt(w mu) until mu is tiny: a handful of instances then needed mu = 1e-14 and thousands of PCG iterations.)
  const double mu = eval ? st[inst].mu_ls : st[inst].mu;
  const int rr0 = P.roff[inst] + (P.edge_off[inst + 1] - P.edge_off[inst]) * P.rpe;
  const int rr1 = rr0 + (P.rng_off[inst + 1] - P.rng_off[inst]) * D;
  double Facc = 0.0, dacc = 0.0, gacc = 0.0, sacc = 0.0;  // objective, |delta|^2, sum lambda s, sum min(lambda, s)^2
  auto plain = [&](int r0, int r1) {
    for (int row = r0 + threadIdx.x; row < r1; row += kThreads) {
      const double wr = P.w[row];
      const double v = V.res[row] + step * V.bdz[row];
      if (!eval) V.res[row] = v;
      Facc += wr * v * v;
      V.u[row] = 2.0 * wr * v;
    }
  };
  plain(bd.i0, min(bd.i1, rr0));
  plain(max(bd.i0, rr1), bd.i1);
  // range terms: one thread per range
  const int ra = max(bd.i0, rr0), rb = min(bd.i1, rr1);
  if (rb > ra) {
    const int nrng = (rb - ra) / D, kfirst = P.rng_off[inst] + (ra - rr0) / D;
    for (int q = threadIdx.x; q < nrng; q += kThreads) {
      const int row = ra + q * D, k = kfirst + q;
      const double rr = P.rng_dist[k], wr = P.rng_w[k];
      double v[D], n2 = 0.0;
#pragma unroll
      for (int c = 0; c < D; ++c) {
        v[c] = V.res[row + c] + step * V.bdz[row + c];
        n2 += v[c] * v[c];
      }
      const double nv = sqrt(n2);
      const RangeTerm t = range_term(nv, rr, wr, mu, !eval);
      Facc += eval ? wr * t.gapr * t.gapr : t.val;  // (certificate: the objective itself, without the barrier term)
#pragma unroll
      for (int c = 0; c < D; ++c) {
        if (!eval) V.res[row + c] = v[c];
        V.u[row + c] = 2.0 * wr * t.tan * v[c];
      }
      if (!eval) {
        const double coef = (n2 > 0.0) ? (t.rad - t.tan) / n2 : 0.0;
        double *mk = V.mk + (size_t)k * (D * (D + 1) / 2);
        int m = 0;
#pragma unroll
        for (int a = 0; a < D; ++a)
#pragma unroll
          for (int b = a; b < D; ++b) mk[m++] = 2.0 * wr * (((a == b) ? t.tan : 0.0) + coef * v[a] * v[b]);
      } else if (rr > 0.0) {
        const double lam = 2.0 * wr * rr * t.gapr, sl = 1.0 - t.rho;  // multiplier ||g_delta|| and slack of the ball
        dacc += t.rho * t.rho;
        gacc += lam * sl;
        sacc += fmin(lam, sl) * fmin(lam, sl);
      }
    }
  }
  const double Ftot = block_sum<kThreads>(Facc, red);
  if (!eval) {
    if (threadIdx.x == 0) V.part_upd[(size_t)bid * 4 + 0] = Ftot;
    return;
  }
  const double dtot = block_sum<kThreads>(dacc, red);
  const double gtot = block_sum<kThreads>(gacc, red);
  const double stot = block_sum<kThreads>(sacc, red);
  if (threadIdx.x == 0) {
    V.part_upd[(size_t)bid * 4 + 0] = Ftot;
    V.part_upd[(size_t)bid * 4 + 1] = dtot;
    V.part_upd[(size_t)bid * 4 + 2] = gtot;
    V.part_upd[(size_t)bid * 4 + 3] = stot;
  }
}

template <int D>
__global__ void __launch_bounds__(kThreads, SCORE_RU_MINB) k_rowupdate(DevProblem P, SolverVecs V, BlockTables T, const InstState *st,
                                                        int mode, WorkLists W) {
  const int *act;
  int n_act;
  wl_get(W, mode == TM_EVAL ? WL_EVAL : WL_LS, act, n_act);
  for (long long item = blockIdx.x; item < (long long)n_act * W.maxrb; item += gridDim.x) {
    const int inst = act[item / W.maxrb], bid = T.rb_begin[inst] + (int)(item % W.maxrb);
    if (bid < T.rb_begin[inst + 1] && bid >= W.rb_lo && bid < W.rb_hi) rowupdate_body<D>(P, V, T, st, mode, bid);
  }
}

// ---- K2: h = B^T u.  PH_CG: dz += alpha p, r -= alpha h.  PH_LS: z += step dz, dz = 0, r = -h (h is the
// gradient of F_mu at the new point).  Evaluation ticks: partial |g|^2, g.z, |z|^2 of the true gradient only.
// `split` (row-partitioned multi-GPU solve): 0 fused; 1 SpMV only (h of this rank's rows -> V.hloc); 2 update only
// (h read from V.hglob, the sum over the ranks).
enum ColSplit : int { CS_FUSED = 0, CS_SPMV = 1, CS_APPLY = 2 };
template <bool NC = true>
__device__ __forceinline__ void colpass_body(DevProblem P, SolverVecs V, BlockTables T, const InstState *st,
                                                      int mode, int split, const int bid) {
  __shared__ double red[kThreads / 32];
  const BlockDesc bd = T.cb[bid];
  const int inst = bd.inst;
  const int phase = st[inst].phase;
  const bool eval = mode == TM_EVAL;
  if (phase == PH_DONE || phase == PH_WAIT) return;
  if (eval ? !st[inst].eval_now : (st[inst].eval_now != 0)) return;
  const double alpha = st[inst].alpha, step = st[inst].step;
  const int pin_end = P.zoff[inst] + P.blk;
  const double *hsrc = V.hglob;
  if (V.mf && phase == PH_CG && !eval) {  // h = B^T H_r B p was formed factor by factor (k_hessvec)
    split = CS_APPLY;
    hsrc = V.h;
  }
  double gg = 0.0, gz = 0.0, zz = 0.0;
  auto apply = [&](int col, double h) {
    if (split == CS_SPMV) {
      V.hloc[col] = h;
      return;
    }
    if (col < pin_end) h = 0.0;
    if (eval) {
      const double zc = V.z[col];
      gg += h * h;
      gz += h * zc;
      zz += zc * zc;
    } else if (phase == PH_CG) {  // (explicit fma: k_cg_update performs the same two operations, bit for bit)
      V.dz[col] = fma(alpha, V.p[col], V.dz[col]);
      V.r[col] = fma(-alpha, h, V.r[col]);
    } else {
      V.z[col] += step * V.dz[col];
      V.dz[col] = 0.0;
      V.r[col] = -h;
    }
  };
  const int ncols = bd.i1 - bd.i0;
  if (bd.kind == CB_POSE) {
    for (int lc = threadIdx.x; lc < ncols; lc += kThreads) {
      const int col = bd.i0 + lc;
      double h = 0.0;
      if (split == CS_APPLY) {
        h = hsrc[col];
      } else {
        for (int k = P.t_indptr[col]; k < P.t_indptr[col + 1]; ++k) h += P.t_vals[k] * ldv<NC>(V.u + P.t_rows[k]);
      }
      apply(col, h);
    }
  } else {
    // heavy (landmark) columns: one warp per column, lanes stride over the entries
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int lc = wid; lc < ncols; lc += kThreads / 32) {
      const int col = bd.i0 + lc;
      double h = 0.0;
      if (split == CS_APPLY) {
        h = hsrc[col];
      } else {
        for (int k = P.t_indptr[col] + lane; k < P.t_indptr[col + 1]; k += 32)
          h += P.t_vals[k] * ldv<NC>(V.u + P.t_rows[k]);
        h = warp_sum(h);
      }
      if (lane == 0) apply(col, h);
    }
  }
  if (eval && split != CS_SPMV) {
    const double a = block_sum<kThreads>(gg, red);
    const double b = block_sum<kThreads>(gz, red);
    const double c = block_sum<kThreads>(zz, red);
    if (threadIdx.x == 0) {
      V.part_col[(size_t)bid * 4 + 0] = a;
      V.part_col[(size_t)bid * 4 + 1] = b;
      V.part_col[(size_t)bid * 4 + 2] = c;
    }
  }
}

__global__ void __launch_bounds__(kThreads) k_colpass(DevProblem P, SolverVecs V, BlockTables T, const InstState *st,
                                                      int mode, int split, WorkLists W) {
  const int *act;
  int n_act;
  wl_get(W, mode == TM_EVAL ? WL_EVAL : WL_RUN, act, n_act);
  for (long long item = blockIdx.x; item < (long long)n_act * W.maxcb; item += gridDim.x) {
    const int inst = act[item / W.maxcb], bid = T.cb_begin[inst] + (int)(item % W.maxcb);
    if (bid < T.cb_begin[inst + 1]) colpass_body(P, V, T, st, mode, split, bid);
  }
}

// ---- ctrl_b: after the preconditioner.  PCG bookkeeping / Newton bookkeeping / termination.
__device__ __forceinline__ void ctrl_b_body(DevProblem P, SolverVecs V, BlockTables T, InstState *st, SolverCfg cfg,
                                            int *n_done, int mode, const int inst) {
  const int lane = threadIdx.x & 31;
  InstState &S = st[inst];
  const int phase = S.phase;
  if (phase == PH_DONE) return;
  const int rb0 = T.rb_begin[inst], rb1 = T.rb_begin[inst + 1];
  if (mode == TM_EVAL) {
    if (!S.eval_now) return;
    // certificate of the un-smoothed problem with the auxiliary variables at their
    // exact minimisers: r_link = 0, r_stat = |g_free| / (1 + |x|), p - D = g_free . z
    const int cb0 = T.cb_begin[inst], cb1 = T.cb_begin[inst + 1];
    const double F = ctrl_sum(V.part_upd, rb0, rb1, 4, 0);
    const double dn2 = ctrl_sum(V.part_upd, rb0, rb1, 4, 1);
    const double gsum = ctrl_sum(V.part_upd, rb0, rb1, 4, 2);
    const double ssum = ctrl_sum(V.part_upd, rb0, rb1, 4, 3);
    // (matrix-free solve: the gradient kernel's sums per pose / landmark block)
    const double *pc = V.mf ? V.part_gr : V.part_col;
    const int c0 = V.mf ? T.pb_begin[inst] : cb0, c1 = V.mf ? T.pb_begin[inst + 1] : cb1;
    const double gg = ctrl_sum(pc, c0, c1, 4, 0);
    const double gz = ctrl_sum(pc, c0, c1, 4, 1);
    const double zz = ctrl_sum(pc, c0, c1, 4, 2);
    if (lane != 0) return;
    // SURVEY.md App. A.7 at x = (z, delta): r_stat = ||x - proj_C(x - g)|| / (1 + ||x||) with min(lambda, s) per range in
    // the delta block; p - D = g_free . z + sum lambda s
    const double pd = gz + gsum;
    S.F = F;
    S.gnorm = sqrt(gg + ssum);
    S.xnorm = sqrt(zz + dn2);
    S.r_stat = S.gnorm / (1.0 + S.xnorm);
    S.r_gap = fabs(pd) / (1.0 + fabs(F) + fabs(F - pd));
    S.gap_mu = fabs(gsum) / (1.0 + fabs(F) + fabs(F - pd));
    S.stat_mu = sqrt(ssum) / (1.0 + S.xnorm);
    S.mu_out = S.mu_ls;
    S.kkt = fmax(S.r_stat, S.r_gap);
    S.n_eval += 1;
    S.eval_now = 0;
    const bool ok = S.kkt <= cfg.kkt_tol;
    if (ok || S.newton_it >= cfg.max_newton || S.stall > 6 || !isfinite(F)) {
      S.phase = PH_DONE;
      S.solved = (ok && isfinite(F)) ? 1 : 0;
      atomicAdd(n_done, 1);
    }
    return;
  }
  if (phase == PH_CG) {
    // one more PCG iteration done (PCG ticks, and line-search ticks for instances still inside a Newton solve)
    double rs_new = ctrl_sum(V.part_seg, P.seg_begin[inst], P.seg_begin[inst + 1], 1, 0);
    if (lane == 0) {
      rs_new += V.part_lm[inst];
      S.cg_it += 1;
      S.total_cg += 1;
      if (S.end_cg || !(rs_new > S.eta * S.eta * S.rs0) || S.cg_it >= cfg.max_cg) {
        S.phase = PH_WAIT;  // Newton system solved to its forcing tolerance: idle until the next line-search tick
      } else {
        S.beta = rs_new / S.rs;
        S.rs = rs_new;
      }
    }
  }
  if (mode == TM_CG_LAST) {  // the next tick of the batch is a line-search tick: waiting instances take it
    if (lane == 0 && S.phase == PH_WAIT) {
      S.phase = PH_LS;
      S.skip_ls = 0;
      S.end_cg = 0;
    }
    return;
  }
  if (mode != TM_LS || phase != PH_LS) return;
  double rs_new = ctrl_sum(V.part_seg, P.seg_begin[inst], P.seg_begin[inst + 1], 1, 0);
  // PH_LS: a new point (or the initial point) has just been evaluated with barrier parameter S.mu
  const double Fmu = ctrl_sum(V.part_upd, rb0, rb1, 4, 0);
  if (lane != 0) return;
  rs_new += V.part_lm[inst];
  S.Fmu = Fmu;
  if (V.trace && S.newton_it < V.trace_cap) {  // one record per Newton step: what the step that just ended looked like
    double *t = V.trace + ((size_t)inst * V.trace_cap + S.newton_it) * kTraceRec;
    t[0] = S.mu_ls;
    t[1] = S.step;
    t[2] = (double)S.cg_it;
    t[3] = S.dec;
    t[4] = Fmu;
    t[5] = rs_new;
    t[6] = (double)S.ls_shift;
    t[7] = S.eta;
  }
  if (!S.skip_ls) S.newton_it += 1;
  S.skip_ls = 0;
  // start the next Newton solve
  S.phase = PH_CG;
  S.rs0 = S.rs = rs_new;
  S.beta = 0.0;
  S.cg_it = 0;
  S.end_cg = 0;
  S.dec = 0.0;
  // inexact-Newton forcing term: tighter on the last barrier stages (the certificate needs the accuracy) and right
  // after a line search that found no decrease (the direction was too inexact to be a descent direction)
  S.eta = cfg.forcing * ((S.mu <= cfg.mu_eval) ? 0.3 : 1.0) * ((S.step == 0.0 && S.newton_it > 0) ? 0.1 : 1.0) *
          (S.eta_scale > 0.0 ? S.eta_scale : 1.0);
  if (S.want_eval || !(rs_new > 0.0) || S.newton_it >= cfg.max_newton || !isfinite(Fmu)) S.eval_now = 1;
  S.want_eval = 0;
}

// One warp per listed instance (run list, plus the waiting instances, or the evaluation list); afterwards the
// instance is filed into the lists of the next tick, and the last CTA to finish publishes them.
__global__ void __launch_bounds__(kSegThreads) k_ctrl_b(DevProblem P, SolverVecs V, BlockTables T, InstState *st,
                                                       SolverCfg cfg, int *n_done, int mode, WorkLists W) {
  const int nw = kSegThreads / 32, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * nw + (threadIdx.x >> 5), stride = gridDim.x * nw;
  if (mode == TM_EVAL) {
    const int n_ev = *W.cnt_ev;
    for (int ai = gw; ai < n_ev; ai += stride) ctrl_b_body(P, V, T, st, cfg, n_done, mode, W.ev[ai]);
    return;
  }
  const int p = *W.par, q = p ^ 1;
  const int n_run = W.cnt[p * 3 + WL_RUN], n_wait = W.cnt[p * 3 + WL_WAIT];
  for (int ai = gw; ai < n_run + n_wait; ai += stride) {
    const int inst = (ai < n_run) ? W.list(p, WL_RUN)[ai] : W.list(p, WL_WAIT)[ai - n_run];
    if (mode == TM_FILE) {
      // after the fused PCG kernel: no bookkeeping left to do, waiting instances take the next line-search tick
      if (lane == 0 && st[inst].phase == PH_WAIT) {
        st[inst].phase = PH_LS;
        st[inst].skip_ls = 0;
        st[inst].end_cg = 0;
      }
    } else if (ai < n_run) {
      ctrl_b_body(P, V, T, st, cfg, n_done, mode, inst);
    } else if (mode == TM_CG_LAST && lane == 0 && st[inst].phase == PH_WAIT) {
      st[inst].phase = PH_LS;
      st[inst].skip_ls = 0;
      st[inst].end_cg = 0;
    }
    if (lane == 0) {
      const int ph = st[inst].phase;
      if (ph == PH_CG || ph == PH_LS) W.list(q, WL_RUN)[atomicAdd(&W.cnt[q * 3 + WL_RUN], 1)] = inst;
      if (ph == PH_LS) W.list(q, WL_LS)[atomicAdd(&W.cnt[q * 3 + WL_LS], 1)] = inst;
      if (ph == PH_WAIT) W.list(q, WL_WAIT)[atomicAdd(&W.cnt[q * 3 + WL_WAIT], 1)] = inst;
      if (mode == TM_LS && st[inst].eval_now) W.ev[atomicAdd(W.cnt_ev, 1)] = inst;
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0 && atomicAdd(W.ticket, 1) == (int)gridDim.x - 1) {
    *W.ticket = 0;
    *W.par = q;
  }
}

// ---- K4 (PH_CG): p = s + beta p   (beta = 0 right after a Newton step; idempotent across an evaluation tick)
__device__ __forceinline__ void pupdate_body(SolverVecs V, BlockTables T, const InstState *st, const int bid) {
  const BlockDesc bd = T.cb[bid];
  if (st[bd.inst].phase != PH_CG) return;
  const double beta = st[bd.inst].beta;
  for (int col = bd.i0 + threadIdx.x; col < bd.i1; col += kThreads) V.p[col] = fma(beta, V.p[col], V.s[col]);
}

__global__ void __launch_bounds__(kThreads) k_pupdate(SolverVecs V, BlockTables T, const InstState *st, WorkLists W) {
  const int *act;
  int n_act;
  wl_get(W, WL_RUN, act, n_act);
  for (long long item = blockIdx.x; item < (long long)n_act * W.maxcb; item += gridDim.x) {
    const int inst = act[item / W.maxcb], bid = T.cb_begin[inst] + (int)(item % W.maxcb);
    if (bid < T.cb_begin[inst + 1]) pupdate_body(V, T, st, bid);
  }
}

// ---- Element-wise column-space updates of the PCG iteration, 128-bit accesses.  Work item = (listed instance) x
// (chunk of kVecChunk consecutive columns of that instance): one CTA streams 16 KB per vector.
constexpr int kVecChunk = 2048;  // doubles per chunk: kThreads threads x 4 double2

// PH_CG (matrix-free operator): dz += alpha p ; r -= alpha h, h = 0 on the pinned pose's columns.
// OWN_ALPHA (PCG ticks): there is no controller kernel between k_hessvec and this one — every CTA sums the instance's
// p'Hp partials itself (warp 0, the controller's fixed order: the same bits in every CTA) and the CTA of chunk 0 records
// the step length / Newton decrement / breakdown flag for the kernels that follow.  One launch less per PCG tick.
template <bool OWN_ALPHA>
__global__ void __launch_bounds__(kThreads) k_cg_update(DevProblem P, SolverVecs V, BlockTables T, InstState *st, WorkLists W) {
  __shared__ double s_alpha;
  const int *act;
  int n_act;
  wl_get(W, WL_RUN, act, n_act);
  if (OWN_ALPHA && blockIdx.x == 0 && threadIdx.x == 0) {  // (what k_ctrl_a does at this point of a tick)
    const int q = (*W.par) ^ 1;
    W.cnt[q * 3 + 0] = W.cnt[q * 3 + 1] = W.cnt[q * 3 + 2] = 0;
  }
  for (long long item = blockIdx.x; item < (long long)n_act * W.maxvc; item += gridDim.x) {
    const int inst = act[item / W.maxvc], chunk = (int)(item % W.maxvc);
    const int z0 = P.zoff[inst], n = P.zoff[inst + 1] - z0, c0 = chunk * kVecChunk;
    if (c0 >= n || st[inst].phase != PH_CG || st[inst].eval_now) continue;
    double alpha;
    if (OWN_ALPHA) {
      __syncthreads();
      if (threadIdx.x < 32) {
        const double pHp = ctrl_sum(V.part_hv, T.pb_begin[inst], T.pb_begin[inst + 1], 1, 0);
        if (threadIdx.x == 0) {
          const double a = cg_step_length(pHp, st[inst].rs);
          s_alpha = a;
          if (chunk == 0) {
            st[inst].alpha = a;
            if (a != 0.0)
              st[inst].dec += a * st[inst].rs;
            else
              st[inst].end_cg = 1;
          }
        }
      }
      __syncthreads();
      alpha = s_alpha;
    } else {
      alpha = st[inst].alpha;
    }
    const int c1 = min(n, c0 + kVecChunk), pin = P.blk;
    double *dz = V.dz + z0, *r = V.r + z0;
    const double *p = V.p + z0, *h = V.h + z0;
    if ((z0 & 1) == 0) {
      const int v0 = c0 / 2, v1 = c1 / 2;  // whole double2 elements
      for (int v = v0 + threadIdx.x; v < v1; v += kThreads) {
        const double2 pp = reinterpret_cast<const double2 *>(p)[v];
        double2 hh = reinterpret_cast<const double2 *>(h)[v];
        double2 dd = reinterpret_cast<double2 *>(dz)[v], rr = reinterpret_cast<double2 *>(r)[v];
        if (2 * v < pin) hh.x = 0.0;
        if (2 * v + 1 < pin) hh.y = 0.0;
        dd.x = fma(alpha, pp.x, dd.x);
        dd.y = fma(alpha, pp.y, dd.y);
        rr.x = fma(-alpha, hh.x, rr.x);
        rr.y = fma(-alpha, hh.y, rr.y);
        reinterpret_cast<double2 *>(dz)[v] = dd;
        reinterpret_cast<double2 *>(r)[v] = rr;
      }
      if ((c1 & 1) && c1 == n && threadIdx.x == 0) {  // odd tail column of the instance
        const int c = c1 - 1;
        const double hv = c < pin ? 0.0 : h[c];
        dz[c] = fma(alpha, p[c], dz[c]);
        r[c] = fma(-alpha, hv, r[c]);
      }
    } else {
      for (int c = c0 + threadIdx.x; c < c1; c += kThreads) {
        const double hv = c < pin ? 0.0 : h[c];
        dz[c] = fma(alpha, p[c], dz[c]);
        r[c] = fma(-alpha, hv, r[c]);
      }
    }
  }
}

// PH_CG: p = s + beta p  (the vectorised twin of k_pupdate)
__global__ void __launch_bounds__(kThreads) k_pupdate_vec(DevProblem P, SolverVecs V, const InstState *st, WorkLists W) {
  const int *act;
  int n_act;
  wl_get(W, WL_RUN, act, n_act);
  for (long long item = blockIdx.x; item < (long long)n_act * W.maxvc; item += gridDim.x) {
    const int inst = act[item / W.maxvc], chunk = (int)(item % W.maxvc);
    const int z0 = P.zoff[inst], n = P.zoff[inst + 1] - z0, c0 = chunk * kVecChunk;
    if (c0 >= n || st[inst].phase != PH_CG) continue;
    const double beta = st[inst].beta;
    const int c1 = min(n, c0 + kVecChunk);
    double *p = V.p + z0;
    const double *s = V.s + z0;
    if ((z0 & 1) == 0) {
      const int v0 = c0 / 2, v1 = c1 / 2;
      for (int v = v0 + threadIdx.x; v < v1; v += kThreads) {
        const double2 ss = reinterpret_cast<const double2 *>(s)[v];
        double2 pp = reinterpret_cast<double2 *>(p)[v];
        pp.x = fma(beta, pp.x, ss.x);
        pp.y = fma(beta, pp.y, ss.y);
        reinterpret_cast<double2 *>(p)[v] = pp;
      }
      if ((c1 & 1) && c1 == n && threadIdx.x == 0) p[c1 - 1] = fma(beta, p[c1 - 1], s[c1 - 1]);
    } else {
      for (int c = c0 + threadIdx.x; c < c1; c += kThreads) p[c] = fma(beta, p[c], s[c]);
    }
  }
}

}  // namespace score
