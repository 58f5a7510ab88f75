// Local refinement after the relaxation (/root/reference/README.md:63-67: "SCORE ... as an
// initialization for local-search": the paper hands SCORE's estimate to GTSAM).  Batched Levenberg-Marquardt on the
// ORIGINAL non-convex range-aided SLAM cost — the reference's objective (score/utils/gurobi_utils.py:358-526) with
// R_p in SO(d) and the auxiliary distance variables eliminated:
//     f = sum_edges k ||t_j - t_i - R_i t~||^2 + tau ||R_j - R_i R~||_F^2
//       + sum_ranges w (||p_a - p_b|| - r~)^2 + sum_priors w ||l - prior||^2 ,        first pose of every instance fixed,
// from the rounded solution of score_solve (or caller-supplied initial values).
//
// Unknowns per pose: (dt in R^d, omega in R^{d(d-1)/2}) with t <- t + dt, R <- R Exp(omega^); per landmark dl in R^d.
// Every damped Gauss-Newton system (J^T W J + lambda I) delta = -J^T W r (Levenberg's damping: Marquardt's diag(J^T W J)
// scaling crawls here — the stiff odometry terms put 1e4 on every diagonal entry, so the low-curvature motions of
// whole chains that the ranges ask for are damped 1e4 times harder than they curve; measured on the CPU restatement:
// cost 527 after 60 iterations against 40.66 after 15) is solved by block-Jacobi-preconditioned
// CG, matrix-free and in GATHER form over the same incidence lists as k_hessvec (hessvec.cuh): one thread per pose sums
// the contributions of the factors incident on it in list order, one CTA-wide fixed-order reduction per landmark — no
// atomics, bit-reproducible.  The state x lives in the layout of the solver's column space (pose blocks [R|t] of
// d(d+1) doubles, then the landmarks); the tangent-space vectors use the same slots (first dof entries of a pose
// block), so the solver's offsets, block tables and incidence records address both.
// All instances of a batch advance in lockstep with per-instance scalars (step lengths, damping, accept / reject).
#pragma once
#include "hessvec.cuh"

namespace score {

template <int D>
struct RefDims {
  static constexpr int NR = (D == 2) ? 1 : 3;  // rotational degrees of freedom
  static constexpr int DOF = D + NR;
  static constexpr int D1 = D + 1, BLK = D * (D + 1);
  static constexpr int NRES = D + D * D;       // residual entries of a relative-pose factor
};

struct RefState {
  double cost, cost_trial, lambda, cost0;
  double rs, rs0, alpha, beta, eta;
  double nu;  // growth factor of the damping after a rejected step (Nielsen's rule)
  int cg_it, cg_done, outer, done, accepted, n_accept, converged, pad1;  // converged: stopped by the tolerance, not by max_outer
};

struct RefVecs {
  double *x, *xt;              // state / trial state (column-space layout)
  double *g, *dg;              // gradient J^T W r and diagonal of J^T W J (tangent layout)
  double *dl, *r, *s, *p, *q;  // CG vectors (tangent layout)
  double *Db, *Mi;             // per pose DOF x DOF (then per landmark d x d): diagonal blocks of J^T W J / inverse of the damped blocks
  double *part_cost, *part_dot;  // per pose / landmark block partial sums
  double *part_aux;              // [2 per block] g.delta and |delta|^2 of the step on trial (predicted decrease)
  // odometry-chain preconditioner in the tangent space: block-tridiagonal (DOF x DOF blocks) along every chain segment
  double *Ob;        // [P x DOF x DOF] coupling block J_{p-1}^T W J_p of the odometry link into pose p (rows: pose p-1)
  double *Sinv, *Lb; // [P x DOF x DOF] inverse pivots S_p^-1 and multipliers L_p = O_p^T S_{p-1}^-1 of the block LDL^T
  double *part_seg;  // [n_seg] r.s over a chain segment
  RefState *st;
};

// M = R G_k for the k-th generator of so(d) (2D: G = [[0,-1],[1,0]]; 3D: G_k = [e_k]_x); R = rotation part of block X
template <int D>
__device__ __forceinline__ void rot_gen(const double (&X)[D * (D + 1)], const int k, double (&M)[D * D]) {
  constexpr int D1 = D + 1;
  if (D == 2) {
#pragma unroll
    for (int r = 0; r < D; ++r) {
      M[r * D + 0] = X[r * D1 + 1];
      M[r * D + 1] = -X[r * D1 + 0];
    }
  } else {
    const int a = (k + 1) % 3, b = (k + 2) % 3;  // column a of M = +R[:, b], column b = -R[:, a], column k = 0
#pragma unroll
    for (int r = 0; r < D; ++r) {
      M[r * D + k] = 0.0;
      M[r * D + a] = X[r * D1 + b];
      M[r * D + b] = -X[r * D1 + a];
    }
  }
}

// Linearisation of one relative-pose factor (i -> j) at (Xi, Xj): residual, and the generator images
//   A_k = R_i G_k t~ (d),  B_k = R_i G_k R~ (d x d),  C_k = R_j G_k (d x d)
// so that  d rt = dt_j - dt_i - sum_k w_i[k] A_k ,   d rR = sum_k w_j[k] C_k - sum_k w_i[k] B_k.
template <int D>
struct EdgeLin {
  static constexpr int NR = RefDims<D>::NR;
  double rt[D], rR[D * D];
  double A[NR][D], B[NR][D * D], C[NR][D * D];
};

template <int D>
__device__ __forceinline__ void edge_residual(const double (&Xi)[D * (D + 1)], const double (&Xj)[D * (D + 1)],
                                              const double (&tm)[D], const double (&Rm)[D * D], double (&rt)[D],
                                              double (&rR)[D * D]) {
  constexpr int D1 = D + 1;
#pragma unroll
  for (int r = 0; r < D; ++r) {
    double a = Xj[r * D1 + D] - Xi[r * D1 + D];
#pragma unroll
    for (int c = 0; c < D; ++c) a -= Xi[r * D1 + c] * tm[c];
    rt[r] = a;
#pragma unroll
    for (int c = 0; c < D; ++c) {
      double b = Xj[r * D1 + c];
#pragma unroll
      for (int m = 0; m < D; ++m) b -= Xi[r * D1 + m] * Rm[m * D + c];
      rR[r * D + c] = b;
    }
  }
}

template <int D>
__device__ __forceinline__ void edge_linearize(const double (&Xi)[D * (D + 1)], const double (&Xj)[D * (D + 1)],
                                               const double (&tm)[D], const double (&Rm)[D * D], EdgeLin<D> &L) {
  constexpr int NR = RefDims<D>::NR;
  edge_residual<D>(Xi, Xj, tm, Rm, L.rt, L.rR);
#pragma unroll
  for (int k = 0; k < NR; ++k) {
    double Mi[D * D];
    rot_gen<D>(Xi, k, Mi);
    rot_gen<D>(Xj, k, L.C[k]);
#pragma unroll
    for (int r = 0; r < D; ++r) {
      double a = 0.0;
#pragma unroll
      for (int c = 0; c < D; ++c) a += Mi[r * D + c] * tm[c];
      L.A[k][r] = a;
#pragma unroll
      for (int c = 0; c < D; ++c) {
        double b = 0.0;
#pragma unroll
        for (int m = 0; m < D; ++m) b += Mi[r * D + m] * Rm[m * D + c];
        L.B[k][r * D + c] = b;
      }
    }
  }
}

// out += J_side^T W u for a residual-space vector u = (ut, uR) of the factor; AT_J: the `to` pose's side
template <int D, bool AT_J>
__device__ __forceinline__ void edge_JT(const EdgeLin<D> &L, const double kw, const double tw, const double (&ut)[D],
                                        const double (&uR)[D * D], double (&out)[RefDims<D>::DOF]) {
  constexpr int NR = RefDims<D>::NR;
  if (AT_J) {
#pragma unroll
    for (int r = 0; r < D; ++r) out[r] += kw * ut[r];
#pragma unroll
    for (int k = 0; k < NR; ++k) {
      double a = 0.0;
#pragma unroll
      for (int i = 0; i < D * D; ++i) a += L.C[k][i] * uR[i];
      out[D + k] += tw * a;
    }
  } else {
#pragma unroll
    for (int r = 0; r < D; ++r) out[r] -= kw * ut[r];
#pragma unroll
    for (int k = 0; k < NR; ++k) {
      double a = 0.0, b = 0.0;
#pragma unroll
      for (int r = 0; r < D; ++r) a += L.A[k][r] * ut[r];
#pragma unroll
      for (int i = 0; i < D * D; ++i) b += L.B[k][i] * uR[i];
      out[D + k] -= kw * a + tw * b;
    }
  }
}

// u = J_i vi + J_j vj
template <int D>
__device__ __forceinline__ void edge_Jv(const EdgeLin<D> &L, const double (&vi)[RefDims<D>::DOF],
                                        const double (&vj)[RefDims<D>::DOF], double (&ut)[D], double (&uR)[D * D]) {
  constexpr int NR = RefDims<D>::NR;
#pragma unroll
  for (int r = 0; r < D; ++r) {
    double a = vj[r] - vi[r];
#pragma unroll
    for (int k = 0; k < NR; ++k) a -= vi[D + k] * L.A[k][r];
    ut[r] = a;
  }
#pragma unroll
  for (int i = 0; i < D * D; ++i) {
    double b = 0.0;
#pragma unroll
    for (int k = 0; k < NR; ++k) b += vj[D + k] * L.C[k][i] - vi[D + k] * L.B[k][i];
    uR[i] = b;
  }
}

// diagonal block J_side^T W J_side (DOF x DOF, row-major, accumulated)
template <int D, bool AT_J>
__device__ __forceinline__ void edge_diag(const EdgeLin<D> &L, const double kw, const double tw,
                                          double (&Dg)[RefDims<D>::DOF * RefDims<D>::DOF]) {
  constexpr int NR = RefDims<D>::NR, DOF = RefDims<D>::DOF;
#pragma unroll
  for (int r = 0; r < D; ++r) Dg[r * DOF + r] += kw;
#pragma unroll
  for (int k = 0; k < NR; ++k)
#pragma unroll
    for (int l = 0; l < NR; ++l) {
      double a = 0.0;
      if (AT_J) {
#pragma unroll
        for (int i = 0; i < D * D; ++i) a += L.C[k][i] * L.C[l][i];
        a *= tw;
      } else {
        double b = 0.0;
#pragma unroll
        for (int r = 0; r < D; ++r) b += L.A[k][r] * L.A[l][r];
#pragma unroll
        for (int i = 0; i < D * D; ++i) a += L.B[k][i] * L.B[l][i];
        a = tw * a + kw * b;
      }
      Dg[(D + k) * DOF + D + l] += a;
    }
  if (!AT_J) {
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
      for (int k = 0; k < NR; ++k) {
        Dg[r * DOF + D + k] += kw * L.A[k][r];
        Dg[(D + k) * DOF + r] += kw * L.A[k][r];
      }
  }
}

// tangent-space slot (first coordinate) of the owner a column record names: pose p -> p * BLK, landmark: its own column
template <int D>
__device__ __forceinline__ int rec_slot(const unsigned rec) {
  constexpr int BLK = D * (D + 1);
  const int col = (int)(rec & 0x7fffffffu);
  return (rec >> 31) ? (col / BLK) * BLK : col;
}

enum RefMode : int { RM_LIN = 0, RM_HV = 1, RM_COST = 2 };

template <int D>
__device__ __forceinline__ void load_tan(const double *v, const int slot, double (&o)[RefDims<D>::DOF]) {
#pragma unroll
  for (int i = 0; i < RefDims<D>::DOF; ++i) o[i] = v[slot + i];
}

// One pose / landmark block.  MODE = RM_LIN: g, dg, Db, partial cost at xs.  RM_HV: q = (J^T W J + lambda I) p, partial
// p.q.  RM_COST: partial cost at xs.
template <int D, int MODE>
__device__ __forceinline__ void ref_visit(const DevProblem &P, const RefVecs &R, const HvBlock bd, const double *xs_all) {
  using RD = RefDims<D>;
  constexpr int DOF = RD::DOF, BLK = RD::BLK, D1 = RD::D1;
  __shared__ double red[16 * (kThreads / 32)];
  const int inst = bd.inst, bid = bd.bid, tid = threadIdx.x;
  if (R.st[inst].done || (MODE == RM_HV && R.st[inst].cg_done)) return;
  const int z0 = bd.z0, Pi = bd.Pi;
  const double *xs = xs_all + z0;
  const double *pv = R.p + z0;
  const double lambda = R.st[inst].lambda;
  const IncRec *__restrict__ recs = reinterpret_cast<const IncRec *>(P.inc_rec);
  double acc_s = 0.0;  // cost (LIN, COST) or p.q (HV)
  if (bd.kind == CB_POSE) {
    const int p = bd.i0 + tid, pg = bd.pg0 + p;
    if (p < bd.i1) {
      double Xo[BLK], vo[DOF], out[DOF], Dg[DOF * DOF];
      load_pose<D>(xs, 0, p, Xo);
#pragma unroll
      for (int i = 0; i < DOF; ++i) out[i] = 0.0;
      if (MODE == RM_HV) load_tan<D>(pv, p * BLK, vo);
      if (MODE == RM_LIN) {
#pragma unroll
        for (int i = 0; i < DOF * DOF; ++i) Dg[i] = 0.0;
      }
      auto edge = [&](const int e, const int other, const bool at_j, const bool link_in) {
        double Xn[BLK], tm[D], Rm[D * D], k2, tau2;
        load_pose<D>(xs, 0, other, Xn);
        load_edge<D>(P, e, tm, Rm, k2, tau2);
        const double kw = 0.5 * k2, tw = 0.5 * tau2;  // (load_edge returns the doubled precisions)
        if (MODE == RM_COST) {
          if (at_j) {
            double rt[D], rR[D * D];
            edge_residual<D>(Xn, Xo, tm, Rm, rt, rR);
#pragma unroll
            for (int r = 0; r < D; ++r) acc_s += kw * rt[r] * rt[r];
#pragma unroll
            for (int i = 0; i < D * D; ++i) acc_s += tw * rR[i] * rR[i];
          }
          return;
        }
        EdgeLin<D> L;
        if (at_j)
          edge_linearize<D>(Xn, Xo, tm, Rm, L);
        else
          edge_linearize<D>(Xo, Xn, tm, Rm, L);
        if (MODE == RM_LIN) {
          if (at_j) {
            edge_JT<D, true>(L, kw, tw, L.rt, L.rR, out);
            edge_diag<D, true>(L, kw, tw, Dg);
            if (link_in) {  // coupling block of the odometry link (rows: pose p-1, columns: this pose)
              double *Ob = R.Ob + (size_t)pg * DOF * DOF;
#pragma unroll
              for (int a = 0; a < DOF; ++a)
#pragma unroll
                for (int b = 0; b < DOF; ++b) {
                  double v = 0.0;
                  if (a < D && b < D) {
                    v = (a == b) ? -kw : 0.0;
                  } else if (a >= D && b < D) {
                    v = -kw * L.A[a - D][b];
                  } else if (a >= D && b >= D) {
#pragma unroll
                    for (int i = 0; i < D * D; ++i) v -= tw * L.B[a - D][i] * L.C[b - D][i];
                  }
                  Ob[a * DOF + b] = v;
                }
            }
#pragma unroll
            for (int r = 0; r < D; ++r) acc_s += kw * L.rt[r] * L.rt[r];
#pragma unroll
            for (int i = 0; i < D * D; ++i) acc_s += tw * L.rR[i] * L.rR[i];
          } else {
            edge_JT<D, false>(L, kw, tw, L.rt, L.rR, out);
            edge_diag<D, false>(L, kw, tw, Dg);
          }
        } else {  // RM_HV
          double vn[DOF], ut[D], uR[D * D];
          load_tan<D>(pv, other * BLK, vn);
          if (at_j) {
            edge_Jv<D>(L, vn, vo, ut, uR);
            edge_JT<D, true>(L, kw, tw, ut, uR, out);
          } else {
            edge_Jv<D>(L, vo, vn, ut, uR);
            edge_JT<D, false>(L, kw, tw, ut, uR, out);
          }
        }
      };
      const int e_in = P.link_edge[pg], e_out = (p + 1 < Pi) ? P.link_edge[pg + 1] : -1;
      if (e_in >= 0) edge(e_in, p - 1, true, true);
      if (e_out >= 0) edge(e_out, p + 1, false, false);
      for (int j = P.inc_ptr[pg]; j < P.inc_ptr[pg + 1]; ++j) {
        const IncRec rec = recs[j];
        const int kind = (int)(rec.x >> kIncShift), id = (int)(rec.x & kIncMask);
        if (kind == INC_EJ) {
          edge(id, (int)rec.z, true, false);
        } else if (kind == INC_EI) {
          edge(id, (int)rec.z, false, false);
        } else if (kind == INC_RA || kind == INC_RB) {
          double tp[D], u[D], n2 = 0.0;
          load_trans_rec<D>(xs, rec.z, tp);
#pragma unroll
          for (int r = 0; r < D; ++r) {
            u[r] = Xo[r * D1 + D] - tp[r];
            n2 += u[r] * u[r];
          }
          const double n = sqrt(n2), w = P.rng_w[id], res = n - P.rng_dist[id];
          if (MODE != RM_HV && kind == INC_RA) acc_s += w * res * res;
          if (MODE == RM_COST || !(n > 0.0)) continue;
          const double inv = 1.0 / n;
#pragma unroll
          for (int r = 0; r < D; ++r) u[r] *= inv;
          if (MODE == RM_LIN) {
#pragma unroll
            for (int r = 0; r < D; ++r) {
              out[r] += w * res * u[r];
#pragma unroll
              for (int c = 0; c < D; ++c) Dg[r * DOF + c] += w * u[r] * u[c];
            }
          } else {
            const int ps = rec_slot<D>(rec.z);
            double dv = 0.0;
#pragma unroll
            for (int r = 0; r < D; ++r) dv += u[r] * (vo[r] - pv[ps + r]);
#pragma unroll
            for (int r = 0; r < D; ++r) out[r] += w * dv * u[r];
          }
        }
      }
      const bool pinned = p == 0;  // pin_pose, gurobi_utils.py:316-333: the first pose of the first chain does not move
      const int slot = z0 + p * BLK;
      if (MODE == RM_LIN) {
        double *Db = R.Db + (size_t)pg * DOF * DOF;
#pragma unroll
        for (int i = 0; i < DOF * DOF; ++i) Db[i] = Dg[i];
#pragma unroll
        for (int i = 0; i < DOF; ++i) {
          R.g[slot + i] = pinned ? 0.0 : out[i];
          R.dg[slot + i] = Dg[i * DOF + i];
        }
      } else if (MODE == RM_HV) {
#pragma unroll
        for (int i = 0; i < DOF; ++i) {
          const double qi = pinned ? 0.0 : out[i] + lambda * vo[i];
          R.q[slot + i] = qi;
          acc_s += qi * vo[i];
        }
      }
    }
    const double tot = block_sum<kThreads>(acc_s, red);
    if (tid == 0) (MODE == RM_HV ? R.part_dot : R.part_cost)[bid] = tot;
    return;
  }
  // landmark block: every landmark reduced by the whole CTA in fixed order
  double tot_s = 0.0;  // thread 0
  for (int q = bd.i0; q < bd.i1; ++q) {
    double v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = 0.0;  // [0, D): vector; [D, D + D(D+1)/2): diagonal block (LIN); [15]: scalar
    const int lslot = Pi * BLK + q * D;
    double lo[D], vo[D];
#pragma unroll
    for (int r = 0; r < D; ++r) {
      lo[r] = xs[lslot + r];
      vo[r] = (MODE == RM_HV) ? pv[lslot + r] : 0.0;
    }
    const int j0 = P.inc_ptr[P.P + bd.lg0 + q], j1 = P.inc_ptr[P.P + bd.lg0 + q + 1];
    for (int j = j0 + tid; j < j1; j += kThreads) {
      const IncRec rec = recs[j];
      const int kind = (int)(rec.x >> kIncShift), id = (int)(rec.x & kIncMask);
      if (kind == INC_PR) {
        const double w = P.prior_w[id];
#pragma unroll
        for (int r = 0; r < D; ++r) {
          const double res = lo[r] - P.prior_t[(size_t)id * D + r];
          if (MODE != RM_HV) v[15] += w * res * res;
          if (MODE == RM_LIN) {
            v[r] += w * res;
            v[D + r * D - r * (r - 1) / 2] += w;  // diagonal entry (r, r) of the packed upper triangle
          }
          if (MODE == RM_HV) v[r] += w * vo[r];
        }
        continue;
      }
      double tp[D], u[D], n2 = 0.0;
      load_trans_rec<D>(xs, rec.z, tp);
#pragma unroll
      for (int r = 0; r < D; ++r) {
        u[r] = lo[r] - tp[r];
        n2 += u[r] * u[r];
      }
      const double n = sqrt(n2), w = P.rng_w[id], res = n - P.rng_dist[id];
      if (MODE != RM_HV && kind == INC_RA) v[15] += w * res * res;
      if (MODE == RM_COST || !(n > 0.0)) continue;
      const double inv = 1.0 / n;
#pragma unroll
      for (int r = 0; r < D; ++r) u[r] *= inv;
      if (MODE == RM_LIN) {
        int m = D;
#pragma unroll
        for (int r = 0; r < D; ++r) {
          v[r] += w * res * u[r];
#pragma unroll
          for (int c = r; c < D; ++c) v[m++] += w * u[r] * u[c];
        }
      } else {
        const int ps = rec_slot<D>(rec.z);
        double dv = 0.0;
#pragma unroll
        for (int r = 0; r < D; ++r) dv += u[r] * (vo[r] - pv[ps + r]);
#pragma unroll
        for (int r = 0; r < D; ++r) v[r] += w * dv * u[r];
      }
    }
    const double tot = block_sum16<kThreads>(v, red);  // thread i < 16 holds the total of v[i]
    __shared__ double sh[16];
    if (tid < 16) sh[tid] = tot;
    __syncthreads();
    if (tid == 0) {
      const int slot = z0 + lslot;
      if (MODE == RM_LIN) {
        double *Db = R.Db + (size_t)P.P * DOF * DOF + (size_t)(bd.lg0 + q) * D * D;
        int m = D;
        for (int r = 0; r < D; ++r)
          for (int c = r; c < D; ++c) {
            Db[r * D + c] = Db[c * D + r] = sh[m];
            if (r == c) R.dg[slot + r] = sh[m];
            ++m;
          }
        for (int r = 0; r < D; ++r) R.g[slot + r] = sh[r];
        tot_s += sh[15];
      } else if (MODE == RM_HV) {
        for (int r = 0; r < D; ++r) {
          const double qi = sh[r] + lambda * vo[r];
          R.q[slot + r] = qi;
          tot_s += qi * vo[r];
        }
      } else {
        tot_s += sh[15];
      }
    }
    __syncthreads();
  }
  if (tid == 0) (MODE == RM_HV ? R.part_dot : R.part_cost)[bid] = tot_s;
}

template <int D, int MODE>
__global__ void __launch_bounds__(kThreads) k_ref_visit(DevProblem P, RefVecs R, BlockTables T, int use_trial) {
  ref_visit<D, MODE>(P, R, T.pb[blockIdx.x], use_trial ? R.xt : R.x);
}

// In-place inverse of a small SPD matrix (n <= 6) by Gauss-Jordan without pivoting
template <int N>
__device__ __forceinline__ void spd_inverse_n(double (&A)[N * N]) {
  double inv[N * N];
#pragma unroll
  for (int i = 0; i < N * N; ++i) inv[i] = ((i / N) == (i % N)) ? 1.0 : 0.0;
#pragma unroll
  for (int c = 0; c < N; ++c) {
    const double piv = 1.0 / A[c * N + c];
#pragma unroll
    for (int j = 0; j < N; ++j) {
      A[c * N + j] *= piv;
      inv[c * N + j] *= piv;
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
      if (i == c) continue;
      const double f = A[i * N + c];
#pragma unroll
      for (int j = 0; j < N; ++j) {
        A[i * N + j] -= f * A[c * N + j];
        inv[i * N + j] -= f * inv[c * N + j];
      }
    }
  }
#pragma unroll
  for (int i = 0; i < N * N; ++i) A[i] = inv[i];
}

// s = M^-1 r for one pose / landmark from the stored inverse block
template <int N>
__device__ __forceinline__ void blk_apply(const double *Mi, const double (&r)[N], double (&s)[N]) {
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double a = 0.0;
#pragma unroll
    for (int j = 0; j < N; ++j) a += Mi[i * N + j] * r[j];
    s[i] = a;
  }
}

enum RefVecMode : int { RV_START = 0, RV_UPDATE = 1, RV_PUPDATE = 2, RV_TRIAL = 3 };

// Element-wise passes, one thread per pose / landmark of a block:
//   RV_START : Mi = (Db + lambda I)^-1;  dl = 0, r = -g, s = Mi r, p = s;  partial r.s
//   RV_UPDATE: dl += alpha p, r -= alpha q, s = Mi r;  partial r.s
//   RV_PUPDATE: p = s + beta p
//   RV_TRIAL : xt = retract(x, dl); partial g.dl and |dl|^2 (predicted decrease)
// (with the chain preconditioner the poses' s and r.s come from k_ref_chain_apply, and p = s from a PUPDATE with beta = 0)
template <int D, int MODE>
__global__ void __launch_bounds__(kThreads) k_ref_vec(DevProblem P, RefVecs R, BlockTables T, const int chain) {
  using RD = RefDims<D>;
  constexpr int DOF = RD::DOF, BLK = RD::BLK, D1 = RD::D1;
  __shared__ double red[kThreads / 32];
  const HvBlock bd = T.pb[blockIdx.x];
  const int inst = bd.inst, tid = threadIdx.x;
  const RefState S = R.st[inst];
  if (S.done) return;
  if ((MODE == RV_UPDATE || MODE == RV_PUPDATE) && S.cg_done) return;
  const int z0 = bd.z0, Pi = bd.Pi;
  double acc = 0.0, acc2 = 0.0;
  const bool pose = bd.kind == CB_POSE;
  const int i = bd.i0 + tid;
  if (i < bd.i1) {
    const int slot = z0 + (pose ? i * BLK : Pi * BLK + i * D);
    const int n = pose ? DOF : D;
    const double *Db = pose ? R.Db + (size_t)(bd.pg0 + i) * DOF * DOF : R.Db + (size_t)P.P * DOF * DOF + (size_t)(bd.lg0 + i) * D * D;
    double *Mi = R.Mi + (Db - R.Db);
    const bool pinned = pose && i == 0;
    if (MODE == RV_START) {
      double rr[DOF], ss[DOF];
      if (pose) {
        double A[DOF * DOF];
#pragma unroll
        for (int k = 0; k < DOF * DOF; ++k) A[k] = Db[k];
#pragma unroll
        for (int k = 0; k < DOF; ++k) {
          const double dk = A[k * DOF + k];
          A[k * DOF + k] = dk + S.lambda + ((dk > 0.0) ? 0.0 : 1.0);  // a coordinate without curvature: identity
        }
        spd_inverse_n<DOF>(A);
#pragma unroll
        for (int k = 0; k < DOF * DOF; ++k) Mi[k] = A[k];
#pragma unroll
        for (int k = 0; k < DOF; ++k) rr[k] = pinned ? 0.0 : -R.g[slot + k];
        blk_apply<DOF>(Mi, rr, ss);
      } else {
        double A[D * D], r2[D], s2[D];
#pragma unroll
        for (int k = 0; k < D * D; ++k) A[k] = Db[k];
#pragma unroll
        for (int k = 0; k < D; ++k) {
          const double dk = A[k * D + k];
          A[k * D + k] = dk + S.lambda + ((dk > 0.0) ? 0.0 : 1.0);
        }
        spd_inverse_n<D>(A);
#pragma unroll
        for (int k = 0; k < D * D; ++k) Mi[k] = A[k];
#pragma unroll
        for (int k = 0; k < D; ++k) r2[k] = -R.g[slot + k];
        blk_apply<D>(Mi, r2, s2);
#pragma unroll
        for (int k = 0; k < D; ++k) {
          rr[k] = r2[k];
          ss[k] = s2[k];
        }
      }
      const bool by_chain = chain && pose;  // the chain kernel computes s (and r.s) of the poses; p = s follows it
      for (int k = 0; k < n; ++k) {
        R.dl[slot + k] = 0.0;
        R.r[slot + k] = rr[k];
        R.s[slot + k] = ss[k];
        R.p[slot + k] = chain ? 0.0 : ss[k];
        if (!by_chain) acc += rr[k] * ss[k];
      }
    } else if (MODE == RV_UPDATE) {
      double rr[DOF], ss[DOF];
      for (int k = 0; k < n; ++k) {
        R.dl[slot + k] += S.alpha * R.p[slot + k];
        rr[k] = pinned ? 0.0 : R.r[slot + k] - S.alpha * R.q[slot + k];
        R.r[slot + k] = rr[k];
      }
      if (pose) {
        blk_apply<DOF>(Mi, rr, ss);
      } else {
        double r2[D], s2[D];
#pragma unroll
        for (int k = 0; k < D; ++k) r2[k] = rr[k];
        blk_apply<D>(Mi, r2, s2);
#pragma unroll
        for (int k = 0; k < D; ++k) ss[k] = s2[k];
      }
      for (int k = 0; k < n; ++k) {
        R.s[slot + k] = ss[k];
        if (!(chain && pose)) acc += rr[k] * ss[k];
      }
    } else if (MODE == RV_PUPDATE) {
      for (int k = 0; k < n; ++k) R.p[slot + k] = R.s[slot + k] + S.beta * R.p[slot + k];
    } else if (MODE == RV_TRIAL) {
      for (int k = 0; k < n; ++k) {
        const double dk = R.dl[slot + k];
        acc += R.g[slot + k] * dk;
        acc2 += dk * dk;
      }
      if (pose) {
        double X[BLK], E[D * D];
        load_pose<D>(R.x + z0, 0, i, X);
        const double *dl = R.dl + slot;
        if (D == 2) {
          const double c = cos(dl[D]), s = sin(dl[D]);
          E[0] = c, E[1] = -s, E[2] = s, E[3] = c;
        } else {  // Rodrigues: Exp(w) = I + a K + b K^2, a = sin(th)/th, b = (1 - cos th)/th^2
          const double w0 = dl[D], w1 = dl[D + 1], w2 = dl[D + 2], th2 = w0 * w0 + w1 * w1 + w2 * w2, th = sqrt(th2);
          const double a = (th > 1e-8) ? sin(th) / th : 1.0 - th2 / 6.0;
          const double b = (th > 1e-8) ? (1.0 - cos(th)) / th2 : 0.5 - th2 / 24.0;
          const double K[9] = {0.0, -w2, w1, w2, 0.0, -w0, -w1, w0, 0.0};
          for (int r = 0; r < 3; ++r)
            for (int cc = 0; cc < 3; ++cc) {
              double k2 = 0.0;
              for (int m = 0; m < 3; ++m) k2 += K[r * 3 + m] * K[m * 3 + cc];
              E[r * 3 + cc] = ((r == cc) ? 1.0 : 0.0) + a * K[r * 3 + cc] + b * k2;
            }
        }
        double *xt = R.xt + z0 + i * BLK;
#pragma unroll
        for (int r = 0; r < D; ++r) {
#pragma unroll
          for (int cc = 0; cc < D; ++cc) {
            double v = 0.0;
#pragma unroll
            for (int m = 0; m < D; ++m) v += X[r * D1 + m] * E[m * D + cc];
            xt[r * D1 + cc] = v;
          }
          xt[r * D1 + D] = X[r * D1 + D] + dl[r];
        }
      } else {
        for (int k = 0; k < D; ++k) R.xt[slot + k] = R.x[slot + k] + R.dl[slot + k];
      }
    }
  }
  if (MODE == RV_START || MODE == RV_UPDATE) {
    const double tot = block_sum<kThreads>(acc, red);
    if (tid == 0) R.part_dot[bd.bid] = tot;
  }
  if (MODE == RV_TRIAL) {
    const double t1 = block_sum<kThreads>(acc, red);
    const double t2 = block_sum<kThreads>(acc2, red);
    if (tid == 0) {
      R.part_aux[2 * bd.bid] = t1;
      R.part_aux[2 * bd.bid + 1] = t2;
    }
  }
}

template <int N>
__device__ __forceinline__ void ld_blk(const double *src, double (&o)[N * N]) {
#pragma unroll
  for (int i = 0; i < N * N; ++i) o[i] = src[i];
}

// Block LDL^T of the tangent-space chain matrix of one segment: diagonal blocks Db + lambda I, couplings Ob.  One thread
// per segment (sequential along the chain; once per outer iteration).
template <int D>
__global__ void k_ref_chain_factor(DevProblem P, RefVecs R) {
  constexpr int DOF = RefDims<D>::DOF, NN = DOF * DOF;
  const int sg = blockIdx.x * blockDim.x + threadIdx.x;
  if (sg >= P.n_seg) return;
  const int inst = P.seg_inst[sg];
  if (R.st[inst].done) return;
  const double lam = R.st[inst].lambda;
  const int p0 = P.seg_ptr[sg], p1 = P.seg_ptr[sg + 1], pin = P.pose_off[inst];
  double Sp[NN];  // S_{p-1}^-1
  bool prev_free = false;
  for (int pg = p0; pg < p1; ++pg) {
    double S[NN], L[NN];
    ld_blk<DOF>(R.Db + (size_t)pg * NN, S);
#pragma unroll
    for (int k = 0; k < DOF; ++k) {
      const double dk = S[k * DOF + k];
      S[k * DOF + k] = dk + lam + ((dk > 0.0) ? 0.0 : 1.0);
    }
    const bool pinned = pg == pin;
    if (pg > p0 && prev_free && !pinned) {
      double O[NN];
      ld_blk<DOF>(R.Ob + (size_t)pg * NN, O);
      // L = O^T Sp ;  S -= L O
#pragma unroll
      for (int a = 0; a < DOF; ++a)
#pragma unroll
        for (int b = 0; b < DOF; ++b) {
          double v = 0.0;
#pragma unroll
          for (int m = 0; m < DOF; ++m) v += O[m * DOF + a] * Sp[m * DOF + b];
          L[a * DOF + b] = v;
        }
#pragma unroll
      for (int a = 0; a < DOF; ++a)
#pragma unroll
        for (int b = 0; b < DOF; ++b) {
          double v = 0.0;
#pragma unroll
          for (int m = 0; m < DOF; ++m) v += L[a * DOF + m] * O[m * DOF + b];
          S[a * DOF + b] -= v;
        }
    } else {
#pragma unroll
      for (int i = 0; i < NN; ++i) L[i] = 0.0;
    }
    if (pinned) {
#pragma unroll
      for (int i = 0; i < NN; ++i) S[i] = ((i / DOF) == (i % DOF)) ? 1.0 : 0.0;
    }
    spd_inverse_n<DOF>(S);
    double *So = R.Sinv + (size_t)pg * NN, *Lo = R.Lb + (size_t)pg * NN;
#pragma unroll
    for (int i = 0; i < NN; ++i) {
      So[i] = S[i];
      Lo[i] = L[i];
      Sp[i] = S[i];
    }
    prev_free = !pinned;
  }
}

// s = M_chain^-1 r for the poses of one segment (forward / backward substitution) and the segment's r.s.
template <int D>
__global__ void k_ref_chain_apply(DevProblem P, RefVecs R, const int start) {
  constexpr int DOF = RefDims<D>::DOF, NN = DOF * DOF, BLK = RefDims<D>::BLK;
  const int sg = blockIdx.x * blockDim.x + threadIdx.x;
  if (sg >= P.n_seg) return;
  const int inst = P.seg_inst[sg];
  if (R.st[inst].done || (!start && R.st[inst].cg_done)) return;  // (start: cg_done still holds the previous solve's flag)
  const int p0 = P.seg_ptr[sg], p1 = P.seg_ptr[sg + 1], pin = P.pose_off[inst];
  const long base = (long)P.zoff[inst] - (long)pin * BLK;
  double y[DOF];
#pragma unroll
  for (int k = 0; k < DOF; ++k) y[k] = 0.0;
#pragma unroll 4  // (the loads of the next poses do not depend on the recurrence: let them be issued ahead)
  for (int pg = p0; pg < p1; ++pg) {  // forward: y_p = r_p - L_p y_{p-1}, kept in s
    const double *L = R.Lb + (size_t)pg * NN;
    const double *rp = R.r + base + (long)pg * BLK;
    double yn[DOF];
#pragma unroll
    for (int a = 0; a < DOF; ++a) {
      double v = (pg == pin) ? 0.0 : rp[a];
#pragma unroll
      for (int m = 0; m < DOF; ++m) v -= L[a * DOF + m] * y[m];
      yn[a] = v;
    }
    double *sp = R.s + base + (long)pg * BLK;
#pragma unroll
    for (int a = 0; a < DOF; ++a) sp[a] = y[a] = yn[a];
  }
  double x[DOF], dot = 0.0;
#pragma unroll
  for (int k = 0; k < DOF; ++k) x[k] = 0.0;
#pragma unroll 4
  for (int pg = p1 - 1; pg >= p0; --pg) {  // backward: x_p = S_p^-1 (y_p - O_{p+1} x_{p+1})
    double *sp = R.s + base + (long)pg * BLK;
    double t[DOF];
#pragma unroll
    for (int a = 0; a < DOF; ++a) t[a] = sp[a];
    if (pg + 1 < p1 && pg != pin) {
      const double *O = R.Ob + (size_t)(pg + 1) * NN;
#pragma unroll
      for (int a = 0; a < DOF; ++a)
#pragma unroll
        for (int m = 0; m < DOF; ++m) t[a] -= O[a * DOF + m] * x[m];
    }
    const double *Si = R.Sinv + (size_t)pg * NN;
    const double *rp = R.r + base + (long)pg * BLK;
#pragma unroll
    for (int a = 0; a < DOF; ++a) {
      double v = 0.0;
#pragma unroll
      for (int m = 0; m < DOF; ++m) v += Si[a * DOF + m] * t[m];
      x[a] = (pg == pin) ? 0.0 : v;
    }
#pragma unroll
    for (int a = 0; a < DOF; ++a) {
      sp[a] = x[a];
      dot += x[a] * rp[a];
    }
  }
  R.part_seg[sg] = dot;
}

enum RefCtrl : int { RC_COST0 = 0, RC_START = 1, RC_ALPHA = 2, RC_BETA = 3, RC_ACCEPT = 4 };

struct RefCfg {
  int max_outer, max_inner;
  double rel_tol, lambda0, eta;
  int chain, pad;  // chain: odometry-chain preconditioner (block LDL^T per segment) instead of block-Jacobi for the poses
};

// Per-instance scalars; one warp per instance, fixed-order sums of the block partials.
template <int MODE>
__global__ void k_ref_ctrl(BlockTables T, RefVecs R, RefCfg cfg, int n_inst, int *n_done, const int *seg_begin) {
  const int lane = threadIdx.x & 31;
  const int inst = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (inst >= n_inst) return;
  RefState &S = R.st[inst];
  if (S.done) return;
  const int b0 = T.pb_begin[inst], b1 = T.pb_begin[inst + 1];
  const double *part = (MODE == RC_COST0 || MODE == RC_ACCEPT) ? R.part_cost : R.part_dot;
  if ((MODE == RC_ALPHA || MODE == RC_BETA) && S.cg_done) return;
  double acc = 0.0, gd = 0.0, dDd = 0.0;
  for (int b = b0 + lane; b < b1; b += 32) acc += part[b];
  acc = warp_sum(acc);
  if ((MODE == RC_START || MODE == RC_BETA) && seg_begin) {  // chain preconditioner: r.s of the poses comes per segment
    double a2 = 0.0;
    for (int sg = seg_begin[inst] + lane; sg < seg_begin[inst + 1]; sg += 32) a2 += R.part_seg[sg];
    acc += warp_sum(a2);
  }
  if (MODE == RC_ACCEPT) {
    for (int b = b0 + lane; b < b1; b += 32) {
      gd += R.part_aux[2 * b];
      dDd += R.part_aux[2 * b + 1];
    }
    gd = warp_sum(gd);
    dDd = warp_sum(dDd);
  }
  if (lane != 0) return;
  if (MODE == RC_COST0) {
    S.cost = S.cost0 = acc;
  } else if (MODE == RC_START) {
    S.rs = S.rs0 = acc;
    S.beta = 0.0;
    S.cg_it = 0;
    S.cg_done = !(acc > 0.0);
    S.accepted = 0;
    if (S.cg_done) atomicAdd(n_done + 1, 1);  // [1]: instances whose PCG solve of this outer iteration has ended
  } else if (MODE == RC_ALPHA) {
    S.alpha = (acc > 0.0 && isfinite(acc)) ? S.rs / acc : 0.0;
    if (S.alpha == 0.0) {
      S.cg_done = 1;
      atomicAdd(n_done + 1, 1);
    }
  } else if (MODE == RC_BETA) {
    S.cg_it += 1;
    if (!(acc > cfg.eta * cfg.eta * S.rs0) || S.cg_it >= cfg.max_inner) {
      S.cg_done = 1;
      atomicAdd(n_done + 1, 1);
    } else {
      S.beta = acc / S.rs;
      S.rs = acc;
    }
  } else {  // RC_ACCEPT: trial cost against the current one; damping by the gain ratio (Nielsen's rule)
    S.cost_trial = acc;
    S.outer += 1;
    // Predicted decrease of f = r'r for a CG iterate delta (started at 0) of (H + lambda I) delta = -g:
    // delta'(H + lambda I) delta = -g'delta, hence f - model = -2 g'delta - delta'H delta = -g'delta + lambda |delta|^2.
    const double decrease = S.cost - acc;
    const double predicted = -gd + S.lambda * dDd;
    const double rho = (predicted > 0.0) ? decrease / predicted : -1.0;
    const bool ok = isfinite(acc) && decrease > 0.0 && rho > 0.0;
    S.accepted = ok ? 1 : 0;
    bool finished = S.outer >= cfg.max_outer;
    if (ok) {
      S.n_accept += 1;
      if (decrease <= cfg.rel_tol * (1.0 + S.cost)) finished = true, S.converged = 1;
      S.cost = acc;
      const double t = 2.0 * rho - 1.0;
      S.lambda = fmax(S.lambda * fmax(1.0 / 3.0, 1.0 - t * t * t), 1e-12);
      S.nu = 2.0;
    } else {
      // no decrease: more damping — or, when even the CG start residual is zero / damping is exhausted, a stationary point
      if (!(S.rs0 > 0.0) || S.lambda >= 1e10) finished = true, S.converged = 1;
      S.lambda = fmin(S.lambda * S.nu, 1e12);
      S.nu = fmin(2.0 * S.nu, 1e6);
    }
    if (finished) {
      S.done = 1;  // (an accepted last step is still committed: k_ref_vec<RV_COMMIT> runs before `done` is honoured)
      atomicAdd(n_done, 1);
    }
  }
}

// x from the last solve (rounded rotations, relaxed translations, landmarks) or from caller-supplied values
__global__ void k_ref_init(DevProblem P, RefVecs R, const double *poses_rt, const double *rounded, const double *lms,
                           double lambda0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < P.n_inst) {
    RefState S;
    memset(&S, 0, sizeof(S));
    S.lambda = lambda0;
    S.nu = 2.0;
    R.st[i] = S;
  }
  if (i >= P.nz) return;
  const int d = P.d, d1 = d + 1;
  const int inst = find_inst(P.zoff, P.n_inst, i);
  const int loc = i - P.zoff[inst];
  const int Pi = P.pose_off[inst + 1] - P.pose_off[inst];
  double v;
  if (loc < Pi * P.blk) {
    const size_t pg = (size_t)P.pose_off[inst] + loc / P.blk;
    const int e = loc % P.blk, r = e / d1, c = e % d1;
    v = (c < d && rounded) ? rounded[pg * d * d + r * d + c] : poses_rt[pg * P.blk + e];
  } else {
    v = lms[(size_t)P.lm_off[inst] * d + (loc - Pi * P.blk)];
  }
  R.x[i] = v;
  R.xt[i] = v;
  R.g[i] = R.dg[i] = R.dl[i] = R.r[i] = R.s[i] = R.p[i] = R.q[i] = 0.0;
}

// commit must also reach instances that finish with an accepted step: run it before `done` is looked at
template <int D>
__global__ void __launch_bounds__(kThreads) k_ref_commit(DevProblem P, RefVecs R, BlockTables T) {
  constexpr int BLK = D * (D + 1);
  const HvBlock bd = T.pb[blockIdx.x];
  const RefState S = R.st[bd.inst];
  if (!S.accepted) return;
  const bool pose = bd.kind == CB_POSE;
  const int i = bd.i0 + threadIdx.x;
  if (i >= bd.i1) return;
  const int slot = bd.z0 + (pose ? i * BLK : bd.Pi * BLK + i * D), len = pose ? BLK : D;
  for (int k = 0; k < len; ++k) R.x[slot + k] = R.xt[slot + k];
}

// clear the accept flag of finished instances so that later commits leave them alone
__global__ void k_ref_clear(RefVecs R, int n_inst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_inst && R.st[i].done) R.st[i].accepted = 0;
}

}  // namespace score
