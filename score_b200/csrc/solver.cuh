// Solver tick kernels.
//
// The convex program the reference hands to Gurobi (score/solve_score.py:76; constraints
// score/utils/gurobi_utils.py:316-352, objective :358-526) is solved here in its cone-eliminated form:
// minimising over the auxiliary distance variable of each range (QCQP: delta in the unit ball,
// SOCP: delta >= ||t_a - t_b||) leaves the smooth convex function
//     F(z) = sum_edges w (B z)^2 + sum_ranges w max(0, ||t_a - t_b|| - r~)^2 + priors ,  z = (R, t, l),
// with the pinned pose held fixed.  F is minimised by a semismooth Newton
// method; every Newton system is solved by conjugate gradients preconditioned with the odometry
// chain (precond.cuh), and a one-pass multi-candidate line search picks the step.
//
// One "tick" = the fixed kernel sequence
//     rowpass -> linesearch -> ctrl_a -> rowupdate -> colpass -> precond -> ctrl_b -> pupdate
// Every instance of a batch advances by one operation per tick according to its own phase
// (PH_CG: one PCG iteration; PH_LS: line search + gradient at the new point), so instances never
// wait for each other.  All scalars live on the device; the host only replays a CUDA graph.
#pragma once
#include "common.cuh"

namespace score {

__device__ __forceinline__ double ls_candidate(int c) {
  // 2^(1 - c/2), c = 0..kNumCand-1
  return ldexp((c & 1) ? 1.4142135623730951 : 1.0, 1 - (c + 1) / 2);
}

// res = B z - b   (initialisation only)
__global__ void __launch_bounds__(kThreads) k_residual(DevProblem P, const double *__restrict__ z, double *res) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= P.m) return;
  double acc = -P.b[row];
  for (int k = P.indptr[row]; k < P.indptr[row + 1]; ++k) acc += P.vals[k] * z[P.cols[k]];
  res[row] = acc;
}

// ---- K1: q = B x.  PH_CG: x = p, u = 2 W J q (J = generalised Jacobian of the per-range shrink at the
// current residual), partial p'Hp.  PH_LS: x = dz, bdz = q.
template <int D>
__global__ void __launch_bounds__(kThreads) k_rowpass(DevProblem P, SolverVecs V, BlockTables T, const InstState *st) {
  __shared__ double sq[kRowsPerBlock];
  __shared__ double red[kThreads / 32];
  const BlockDesc bd = T.rb[blockIdx.x];
  const int phase = st[bd.inst].phase;
  if (phase == PH_DONE) return;
  const double *__restrict__ x = (phase == PH_CG) ? V.p : V.dz;
  const int nrows = bd.i1 - bd.i0;
  for (int li = threadIdx.x; li < nrows; li += kThreads) {
    const int row = bd.i0 + li;
    const int k0 = P.indptr[row], k1 = P.indptr[row + 1];
    double acc = 0.0;
    for (int k = k0; k < k1; ++k) acc += P.vals[k] * __ldg(x + P.cols[k]);
    sq[li] = acc;
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
    double out = q;
    if (row >= rr0 && row < rr1) {
      const int rel = row - rr0, comp = rel % D, base = row - comp;
      const double rr = P.rng_dist[P.rng_off[inst] + rel / D];
      double v[D], n2 = 0.0, dot = 0.0;
#pragma unroll
      for (int c = 0; c < D; ++c) {
        v[c] = V.res[base + c];
        n2 += v[c] * v[c];
        dot += v[c] * sq[base - bd.i0 + c];
      }
      const double nv = sqrt(n2);
      if (nv > rr) {
        const double inv = 1.0 / nv;
        out = (1.0 - rr * inv) * q + (rr * inv) * (dot * inv) * (v[comp] * inv);
      } else {
        out = 0.0;
      }
    }
    const double u = 2.0 * P.w[row] * out;
    V.u[row] = u;
    acc += q * u;
  }
  const double tot = block_sum<kThreads>(acc, red);
  if (threadIdx.x == 0) V.part_row[blockIdx.x] = tot;
}

// ---- K_ls: phi(a) = F(res + a bdz) at kNumCand step sizes plus a = 0, in one pass.
// Plain rows are quadratic in a (three sums); range rows are evaluated per candidate.
template <int D>
__global__ void __launch_bounds__(kThreads) k_linesearch(DevProblem P, SolverVecs V, BlockTables T, const InstState *st) {
  __shared__ double red[kThreads / 32];
  const BlockDesc bd = T.rb[blockIdx.x];
  const int inst = bd.inst;
  if (st[inst].phase != PH_LS || st[inst].skip_ls) return;
  const int rr0 = P.roff[inst] + (P.edge_off[inst + 1] - P.edge_off[inst]) * P.rpe;
  const int rr1 = rr0 + (P.rng_off[inst + 1] - P.rng_off[inst]) * D;
  double sums[kLsSums];
#pragma unroll
  for (int i = 0; i < kLsSums; ++i) sums[i] = 0.0;
  const int nrows = bd.i1 - bd.i0;
  for (int li = threadIdx.x; li < nrows; li += kThreads) {
    const int row = bd.i0 + li;
    if (row >= rr0 && row < rr1) {
      const int rel = row - rr0;
      if (rel % D != 0) continue;
      const int k = P.rng_off[inst] + rel / D;
      const double rr = P.rng_dist[k], wk = P.rng_w[k];
      double A = 0.0, Bq = 0.0, C = 0.0;
#pragma unroll
      for (int c = 0; c < D; ++c) {
        const double v = V.res[row + c], q = V.bdz[row + c];
        A += v * v;
        Bq += v * q;
        C += q * q;
      }
#pragma unroll
      for (int c = 0; c < kNumCand; ++c) {
        const double a = ls_candidate(c);
        const double e = fmax(0.0, sqrt(fmax(0.0, A + a * (2.0 * Bq + a * C))) - rr);
        sums[3 + c] += wk * e * e;
      }
      const double e0 = fmax(0.0, sqrt(A) - rr);
      sums[3 + kNumCand] += wk * e0 * e0;
    } else {
      const double v = V.res[row], q = V.bdz[row], wr = P.w[row];
      sums[0] += wr * v * v;
      sums[1] += wr * v * q;
      sums[2] += wr * q * q;
    }
  }
#pragma unroll
  for (int i = 0; i < kLsSums; ++i) {
    const double tot = block_sum<kThreads>(sums[i], red);
    if (threadIdx.x == 0) V.part_ls[(size_t)blockIdx.x * kLsSums + i] = tot;
  }
}

// Fixed-order sum of part[i0..i1) with stride `stride`, by a 128-thread CTA; valid in thread 0.
__device__ __forceinline__ double ctrl_sum(const double *part, int i0, int i1, int stride, int off, double *red) {
  double acc = 0.0;
  for (int i = i0 + threadIdx.x; i < i1; i += kSegThreads) acc += part[(size_t)i * stride + off];
  return block_sum<kSegThreads>(acc, red);
}

// ---- ctrl_a: PCG step length / line-search decision.  One CTA per instance.
__global__ void __launch_bounds__(kSegThreads) k_ctrl_a(SolverVecs V, BlockTables T, InstState *st) {
  __shared__ double red[kSegThreads / 32];
  const int inst = blockIdx.x;
  InstState &S = st[inst];
  const int phase = S.phase;
  if (phase == PH_DONE) return;
  const int b0 = T.rb_begin[inst], b1 = T.rb_begin[inst + 1];
  if (phase == PH_CG) {
    double pHp = ctrl_sum(V.part_row, b0, b1, 1, 0, red);
    double pt = 0.0;
    if (S.lam > 0.0) pt = ctrl_sum(V.part_col, T.cb_begin[inst], T.cb_begin[inst + 1], 4, 3, red);
    if (threadIdx.x == 0) {
      pHp += S.lam * pt;
      if (pHp > 0.0 && pHp > 1e-30 * fabs(S.rs) && isfinite(pHp)) {
        S.alpha = S.rs / pHp;
      } else {
        S.alpha = 0.0;
        S.end_cg = 1;
      }
    }
    return;
  }
  // PH_LS
  if (S.skip_ls) {
    if (threadIdx.x == 0) S.step = 0.0;
    return;
  }
  double tot[kLsSums];
  for (int i = 0; i < kLsSums; ++i) tot[i] = ctrl_sum(V.part_ls, b0, b1, kLsSums, i, red);
  if (threadIdx.x == 0) {
    double best = tot[0] + tot[3 + kNumCand], step = 0.0;
    for (int c = 0; c < kNumCand; ++c) {
      const double a = ls_candidate(c);
      const double Fc = tot[0] + a * (2.0 * tot[1] + a * tot[2]) + tot[3 + c];
      if (Fc < best) {
        best = Fc;
        step = a;
      }
    }
    S.step = step;
  }
}

// ---- K_upd (PH_LS): res += step * bdz ; u = y = 2 W shrink(res) ; partial F and |delta|^2.
template <int D>
__global__ void __launch_bounds__(kThreads) k_rowupdate(DevProblem P, SolverVecs V, BlockTables T, const InstState *st) {
  __shared__ double red[kThreads / 32];
  const BlockDesc bd = T.rb[blockIdx.x];
  const int inst = bd.inst;
  if (st[inst].phase != PH_LS) return;
  const double step = st[inst].step;
  const int rr0 = P.roff[inst] + (P.edge_off[inst + 1] - P.edge_off[inst]) * P.rpe;
  const int rr1 = rr0 + (P.rng_off[inst + 1] - P.rng_off[inst]) * D;
  const int nrows = bd.i1 - bd.i0;
  double Facc = 0.0, dacc = 0.0;
  // two sweeps so that every thread of a range reads the un-updated residual of its siblings
  double newv[(kRowsPerBlock + kThreads - 1) / kThreads];
  int j = 0;
  for (int li = threadIdx.x; li < nrows; li += kThreads, ++j) {
    const int row = bd.i0 + li;
    double out;
    if (row >= rr0 && row < rr1) {
      const int rel = row - rr0, comp = rel % D, base = row - comp;
      const int k = P.rng_off[inst] + rel / D;
      const double rr = P.rng_dist[k];
      double v[D], n2 = 0.0;
#pragma unroll
      for (int c = 0; c < D; ++c) {
        v[c] = V.res[base + c] + step * V.bdz[base + c];
        n2 += v[c] * v[c];
      }
      newv[j] = v[comp];
      const double nv = sqrt(n2);
      const double fac = (nv > rr) ? 1.0 - rr / nv : 0.0;
      out = fac * v[comp];
      if (comp == 0 && rr > 0.0) {
        const double dn = fmin(1.0, nv / rr);
        dacc += dn * dn;
      }
    } else {
      newv[j] = V.res[row] + step * V.bdz[row];
      out = newv[j];
    }
    const double wr = P.w[row];
    V.u[row] = 2.0 * wr * out;
    Facc += wr * out * out;
  }
  __syncthreads();
  j = 0;
  for (int li = threadIdx.x; li < nrows; li += kThreads, ++j) V.res[bd.i0 + li] = newv[j];
  const double Ftot = block_sum<kThreads>(Facc, red);
  const double dtot = block_sum<kThreads>(dacc, red);
  if (threadIdx.x == 0) {
    V.part_upd[(size_t)blockIdx.x * 2 + 0] = Ftot;
    V.part_upd[(size_t)blockIdx.x * 2 + 1] = dtot;
  }
}

// ---- K2: h = B^T u (+ lam t).  PH_CG: dz += alpha p, r -= alpha h.  PH_LS: z += step dz, dz = 0,
// r = -h (h is the gradient at the new point), partial |g|^2, g.z, |z|^2.
__global__ void __launch_bounds__(kThreads) k_colpass(DevProblem P, SolverVecs V, BlockTables T, const InstState *st) {
  __shared__ double red[kThreads / 32];
  const BlockDesc bd = T.cb[blockIdx.x];
  const int inst = bd.inst;
  const int phase = st[inst].phase;
  if (phase == PH_DONE) return;
  const double alpha = st[inst].alpha, lam = st[inst].lam, step = st[inst].step;
  const int pin_end = P.zoff[inst] + P.blk;
  double gg = 0.0, gz = 0.0, zz = 0.0;
  auto apply = [&](int col, double h) {
    if (col < pin_end) h = 0.0;
    if (phase == PH_CG) {
      if (lam > 0.0) h += lam * V.t[col];
      V.dz[col] += alpha * V.p[col];
      V.r[col] -= alpha * h;
    } else {
      const double zn = V.z[col] + step * V.dz[col];
      V.z[col] = zn;
      V.dz[col] = 0.0;
      V.r[col] = -h;
      gg += h * h;
      gz += h * zn;
      zz += zn * zn;
    }
  };
  const int ncols = bd.i1 - bd.i0;
  if (bd.kind == CB_POSE) {
    for (int lc = threadIdx.x; lc < ncols; lc += kThreads) {
      const int col = bd.i0 + lc;
      double h = 0.0;
      for (int k = P.t_indptr[col]; k < P.t_indptr[col + 1]; ++k) h += P.t_vals[k] * __ldg(V.u + P.t_rows[k]);
      apply(col, h);
    }
  } else {
    // heavy (landmark) columns: one warp per column, lanes stride over the entries
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int lc = wid; lc < ncols; lc += kThreads / 32) {
      const int col = bd.i0 + lc;
      double h = 0.0;
      for (int k = P.t_indptr[col] + lane; k < P.t_indptr[col + 1]; k += 32)
        h += P.t_vals[k] * __ldg(V.u + P.t_rows[k]);
      h = warp_sum(h);
      if (lane == 0) apply(col, h);
    }
  }
  if (phase == PH_LS) {
    const double a = block_sum<kThreads>(gg, red);
    const double b = block_sum<kThreads>(gz, red);
    const double c = block_sum<kThreads>(zz, red);
    if (threadIdx.x == 0) {
      V.part_col[(size_t)blockIdx.x * 4 + 0] = a;
      V.part_col[(size_t)blockIdx.x * 4 + 1] = b;
      V.part_col[(size_t)blockIdx.x * 4 + 2] = c;
    }
  }
}

// ---- ctrl_b: after the preconditioner.  PCG bookkeeping / Newton bookkeeping / termination.
__global__ void __launch_bounds__(kSegThreads) k_ctrl_b(DevProblem P, SolverVecs V, BlockTables T, InstState *st,
                                                       SolverCfg cfg, int *n_done) {
  __shared__ double red[kSegThreads / 32];
  const int inst = blockIdx.x;
  InstState &S = st[inst];
  const int phase = S.phase;
  if (phase == PH_DONE) return;
  double rs_new = ctrl_sum(V.part_seg, P.seg_begin[inst], P.seg_begin[inst + 1], 1, 0, red);
  if (phase == PH_CG) {
    if (threadIdx.x == 0) {
      rs_new += V.part_lm[inst];
      S.cg_it += 1;
      S.total_cg += 1;
      if (S.end_cg || !(rs_new > S.eta * S.eta * S.rs0) || S.cg_it >= cfg.max_cg) {
        S.phase = PH_LS;
        S.skip_ls = 0;
        S.end_cg = 0;
      } else {
        S.beta = rs_new / S.rs;
        S.rs = rs_new;
      }
    }
    return;
  }
  // PH_LS: a new point (or the initial point) has just been evaluated
  const int rb0 = T.rb_begin[inst], rb1 = T.rb_begin[inst + 1];
  const int cb0 = T.cb_begin[inst], cb1 = T.cb_begin[inst + 1];
  const double F = ctrl_sum(V.part_upd, rb0, rb1, 2, 0, red);
  const double dn2 = ctrl_sum(V.part_upd, rb0, rb1, 2, 1, red);
  const double gg = ctrl_sum(V.part_col, cb0, cb1, 4, 0, red);
  const double gz = ctrl_sum(V.part_col, cb0, cb1, 4, 1, red);
  const double zz = ctrl_sum(V.part_col, cb0, cb1, 4, 2, red);
  if (threadIdx.x != 0) return;
  rs_new += V.part_lm[inst];
  // relative KKT of SURVEY.md App. A.7 with the auxiliary variables at their exact minimisers:
  // r_link = 0, r_stat = |g_free| / (1 + |x|), p - D = g_free . z  =>  r_gap = |g.z| / (1 + |p| + |D|)
  S.F = F;
  S.gnorm = sqrt(gg);
  S.xnorm = sqrt(zz + dn2);
  S.r_stat = S.gnorm / (1.0 + S.xnorm);
  S.r_gap = fabs(gz) / (1.0 + fabs(F) + fabs(F - gz));
  S.kkt = fmax(S.r_stat, S.r_gap);
  if (!S.skip_ls) {
    S.newton_it += 1;
    if (S.step == 0.0) {  // no candidate decreased F: damp the next Newton system
      S.ls_fail += 1;
      S.lam = fmax(S.lam, 1e-6) * 10.0;
    } else if (S.step >= 0.99) {
      S.lam = (S.lam < 1e-9) ? 0.0 : S.lam / 3.0;
    } else if (S.step < 0.3) {
      S.lam *= 3.0;
    }
  }
  S.skip_ls = 0;
  const bool ok = S.kkt <= cfg.kkt_tol || !(rs_new > 0.0);
  if (ok || S.newton_it >= cfg.max_newton || S.ls_fail > 40 || !isfinite(F)) {
    S.phase = PH_DONE;
    S.solved = (ok && isfinite(F)) ? 1 : 0;
    atomicAdd(n_done, 1);
    return;
  }
  S.phase = PH_CG;
  S.rs0 = S.rs = rs_new;
  S.beta = 0.0;
  S.cg_it = 0;
  S.end_cg = 0;
  S.eta = cfg.forcing;
}

// ---- K4 (PH_CG): p = s + beta p ; t = r + beta t  (t = P^{-1} p, used by the damping term) ; partial p.t
__global__ void __launch_bounds__(kThreads) k_pupdate(SolverVecs V, BlockTables T, const InstState *st) {
  __shared__ double red[kThreads / 32];
  const BlockDesc bd = T.cb[blockIdx.x];
  if (st[bd.inst].phase != PH_CG) return;
  const double beta = st[bd.inst].beta;
  const bool need_pt = st[bd.inst].lam > 0.0;
  double acc = 0.0;
  for (int col = bd.i0 + threadIdx.x; col < bd.i1; col += kThreads) {
    const double pn = V.s[col] + beta * V.p[col];
    const double tn = V.r[col] + beta * V.t[col];
    V.p[col] = pn;
    V.t[col] = tn;
    acc += pn * tn;
  }
  if (need_pt) {
    const double tot = block_sum<kThreads>(acc, red);
    if (threadIdx.x == 0) V.part_col[(size_t)blockIdx.x * 4 + 3] = tot;
  }
}

}  // namespace score
