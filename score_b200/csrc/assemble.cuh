// On-device lowering of the factor graph to the least-squares operator
//   f(x) = sum_i w_i (B x - b)_i^2
// in the reference's variable / row order.  One thread per factor writes that
// factor's fixed-size row block; indptr is closed-form so no scan is needed.
//
// Reference algebra restated here (paths relative to /root/reference):
//   relative pose  score/utils/gurobi_utils.py:504-526   k||t_j - t_i - R_i t~||^2 + tau||R_j - R_i R~||_F^2
//   range (QCQP)   score/utils/gurobi_utils.py:475-501   w||t_a - t_b - r~ delta||^2
//   range (SOCP)   same lines                            w(delta - r~)^2
//   landmark prior score/utils/gurobi_utils.py:433-446   w||l - prior||^2
//   column order   score/utils/gurobi_utils.py:233-310 ; row order :358-377
#pragma once
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace score {

enum AsmMode : int { ASM_REDUCED = 0, ASM_FULL_QCQP = 1, ASM_FULL_SOCP = 2 };

struct AsmOut {
  int *indptr, *cols;
  double *vals, *w, *b;
  int *nnz_row;  // optional: owning row of every stored entry (for the transpose)
};

__global__ void k_assemble(DevProblem P, int mode, int inst0, int inst1, AsmOut out) {
  const int d = P.d, blk = P.blk, rpe = P.rpe, npe = P.npe;
  const int e0 = P.edge_off[inst0], k0 = P.rng_off[inst0], l0 = P.prior_off[inst0];
  const int nE = P.edge_off[inst1] - e0, nK = P.rng_off[inst1] - k0, nLp = P.prior_off[inst1] - l0;
  const int rpr = (mode == ASM_FULL_SOCP) ? 1 : d;
  const int npr = (mode == ASM_REDUCED) ? 2 * d : (mode == ASM_FULL_QCQP ? 3 * d : 1);
  const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= (long)nE + nK + nLp) return;

  int inst;
  if (tid < nE)
    inst = find_inst(P.edge_off, P.n_inst, e0 + (int)tid);
  else if (tid < nE + nK)
    inst = find_inst(P.rng_off, P.n_inst, k0 + (int)(tid - nE));
  else
    inst = find_inst(P.prior_off, P.n_inst, l0 + (int)(tid - nE - nK));

  const int Ei = P.edge_off[inst + 1] - P.edge_off[inst];
  const int Ki = P.rng_off[inst + 1] - P.rng_off[inst];
  const int Pi = P.pose_off[inst + 1] - P.pose_off[inst];
  const int Li = P.lm_off[inst + 1] - P.lm_off[inst];
  const int rowbase = (P.edge_off[inst] - e0) * rpe + (P.rng_off[inst] - k0) * rpr + (P.prior_off[inst] - l0) * d;
  const int nnzbase = (P.edge_off[inst] - e0) * npe + (P.rng_off[inst] - k0) * npr + (P.prior_off[inst] - l0) * d;
  const int colbase = (mode == ASM_REDUCED) ? P.zoff[inst] : 0;

  auto put = [&](int nz, int row, int col, double v) {
    out.cols[nz] = col;
    out.vals[nz] = v;
    if (out.nnz_row) out.nnz_row[nz] = row;
  };

  if (tid < nE) {
    const int e = e0 + (int)tid, el = e - P.edge_off[inst];
    const int i = P.edge_i[e], j = P.edge_j[e];
    const int ci = colbase + i * blk, cj = colbase + j * blk;
    const double *tm = P.edge_t + (size_t)e * d;
    const double *Rm = P.edge_R + (size_t)e * d * d;
    const double kk = P.edge_k[e], tau = P.edge_tau[e];
    int row = rowbase + el * rpe, nz = nnzbase + el * npe;
    const bool j_after = cj > ci;
    for (int r = 0; r < d; ++r) {  // translation rows: t_j[r] - t_i[r] - sum_c t~[c] R_i[r,c]
      out.indptr[row] = nz;
      out.w[row] = kk;
      out.b[row] = 0.0;
      if (!j_after) put(nz++, row, cj + r * (d + 1) + d, 1.0);
      for (int c = 0; c < d; ++c) put(nz++, row, ci + r * (d + 1) + c, -tm[c]);
      put(nz++, row, ci + r * (d + 1) + d, -1.0);
      if (j_after) put(nz++, row, cj + r * (d + 1) + d, 1.0);
      ++row;
    }
    for (int r = 0; r < d; ++r)
      for (int c = 0; c < d; ++c) {  // rotation rows: R_j[r,c] - sum_m R~[m,c] R_i[r,m]
        out.indptr[row] = nz;
        out.w[row] = tau;
        out.b[row] = 0.0;
        if (!j_after) put(nz++, row, cj + r * (d + 1) + c, 1.0);
        for (int mm = 0; mm < d; ++mm) put(nz++, row, ci + r * (d + 1) + mm, -Rm[mm * d + c]);
        if (j_after) put(nz++, row, cj + r * (d + 1) + c, 1.0);
        ++row;
      }
  } else if (tid < nE + nK) {
    const int k = k0 + (int)(tid - nE), kl = k - P.rng_off[inst];
    const int a = P.rng_a[k], bb = P.rng_b[k];
    const double dist = P.rng_dist[k], wk = P.rng_w[k];
    const int dcol0 = colbase + Pi * blk + Li * d;
    if (mode == ASM_FULL_SOCP) {
      const int row = rowbase + Ei * rpe + kl, nz = nnzbase + Ei * npe + kl;
      out.indptr[row] = nz;
      out.w[row] = wk;
      out.b[row] = dist;
      put(nz, row, dcol0 + kl, 1.0);
    } else {
      for (int r = 0; r < d; ++r) {
        const int row = rowbase + Ei * rpe + kl * d + r;
        int nz = nnzbase + Ei * npe + kl * npr + r * (npr / d);
        const int ca = (a < Pi) ? colbase + a * blk + r * (d + 1) + d : colbase + Pi * blk + (a - Pi) * d + r;
        const int cb = (bb < Pi) ? colbase + bb * blk + r * (d + 1) + d : colbase + Pi * blk + (bb - Pi) * d + r;
        out.indptr[row] = nz;
        out.w[row] = wk;
        out.b[row] = 0.0;
        if (ca < cb) {
          put(nz++, row, ca, 1.0);
          put(nz++, row, cb, -1.0);
        } else {
          put(nz++, row, cb, -1.0);
          put(nz++, row, ca, 1.0);
        }
        if (mode == ASM_FULL_QCQP) put(nz++, row, dcol0 + kl * d + r, -dist);
      }
    }
  } else {
    const int pl = l0 + (int)(tid - nE - nK), pll = pl - P.prior_off[inst];
    const int lq = P.prior_l[pl];
    for (int r = 0; r < d; ++r) {
      const int row = rowbase + Ei * rpe + Ki * rpr + pll * d + r;
      const int nz = nnzbase + Ei * npe + Ki * npr + pll * d + r;
      out.indptr[row] = nz;
      out.w[row] = P.prior_w[pl];
      out.b[row] = P.prior_t[(size_t)pl * d + r];
      put(nz, row, colbase + Pi * blk + lq * d + r, 1.0);
    }
  }
  if (tid == 0) {
    const int rows = nE * rpe + nK * rpr + nLp * d;
    const int nnz = nE * npe + nK * npr + nLp * d;
    out.indptr[rows] = nnz;
  }
}

__global__ void k_iota(int *a, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = i;
}

// After a stable sort of the entries by column: gather rows/values and derive indptr.
__global__ void k_transpose_fill(int nnz, int ncols, const int *__restrict__ sorted_cols, const int *__restrict__ perm,
                                 const int *__restrict__ nnz_row, const double *__restrict__ vals, int *t_indptr,
                                 int *t_rows, double *t_vals) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nnz) return;
  const int src = perm[k];
  t_rows[k] = nnz_row[src];
  t_vals[k] = vals[src];
  const int c = sorted_cols[k];
  const int cprev = (k == 0) ? -1 : sorted_cols[k - 1];
  for (int cc = cprev + 1; cc <= c; ++cc) t_indptr[cc] = k;
  if (k == nnz - 1)
    for (int cc = c + 1; cc <= ncols; ++cc) t_indptr[cc] = nnz;
}

// The CSR of B^T is obtained by a stable LSD radix sort of the entries by column (cub, api.cu) —
// stability keeps the row-major order, so every column lists its rows ascending — followed by
// k_transpose_fill.

}  // namespace score
