// Value extraction: SO(d) rounding of the relaxed rotations and the exact auxiliary distances.
//
// Reference: VariableCollection.get_variable_values (score/utils/gurobi_utils.py:114-136) which calls
// round_to_special_orthogonal (score/utils/matrix_utils.py:59-79):  R = U diag(1,..,1,det(U V^T)) V^T,
// i.e. the maximiser of tr(R^T M) over SO(d).  One thread per pose:
//   d = 2: closed form  (c, s) = (M00 + M11, M10 - M01) / hypot(.)
//   d = 3: Horn's unit-quaternion form — top eigenvector of the 4x4 symmetric matrix N(M), found by
//          cyclic Jacobi sweeps; no singular-value division, so rank-deficient M is handled too.
#pragma once
#include "common.cuh"
#include "solver.cuh"

namespace score {

__device__ inline void round_so2(const double *M, int ldm, double *R) {
  const double a = M[0] + M[ldm + 1], b = M[ldm] - M[1];
  const double h = hypot(a, b);
  double c = 1.0, s = 0.0;
  if (h > 0.0) {
    c = a / h;
    s = b / h;
  }
  R[0] = c;
  R[1] = -s;
  R[2] = s;
  R[3] = c;
}

__device__ inline void round_so3(const double *M, int ldm, double *R) {
  // S = M^T  (S_xy = M_yx)
  const double Sxx = M[0], Sxy = M[ldm], Sxz = M[2 * ldm];
  const double Syx = M[1], Syy = M[ldm + 1], Syz = M[2 * ldm + 1];
  const double Szx = M[2], Szy = M[ldm + 2], Szz = M[2 * ldm + 2];
  double A[4][4] = {{Sxx + Syy + Szz, Syz - Szy, Szx - Sxz, Sxy - Syx},
                    {Syz - Szy, Sxx - Syy - Szz, Sxy + Syx, Szx + Sxz},
                    {Szx - Sxz, Sxy + Syx, -Sxx + Syy - Szz, Syz + Szy},
                    {Sxy - Syx, Szx + Sxz, Syz + Szy, -Sxx - Syy + Szz}};
  double V[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
  double scale = 0.0;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) scale = fmax(scale, fabs(A[i][j]));
  if (scale == 0.0) {
    for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
    return;
  }
  for (int sweep = 0; sweep < 16; ++sweep) {
    double off = 0.0;
    for (int p = 0; p < 3; ++p)
      for (int q = p + 1; q < 4; ++q) off += A[p][q] * A[p][q];
    if (off <= 1e-34 * scale * scale) break;
    for (int p = 0; p < 3; ++p)
      for (int q = p + 1; q < 4; ++q) {
        const double apq = A[p][q];
        if (fabs(apq) <= 1e-300) continue;
        const double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
        const double t = ((theta >= 0.0) ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 4; ++k) {
          const double akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - s * akq;
          A[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 4; ++k) {
          const double apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - s * aqk;
          A[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 4; ++k) {
          const double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq;
          V[k][q] = s * vkp + c * vkq;
        }
      }
  }
  int best = 0;
  for (int i = 1; i < 4; ++i)
    if (A[i][i] > A[best][best]) best = i;
  double w = V[0][best], x = V[1][best], y = V[2][best], z = V[3][best];
  const double n = 1.0 / sqrt(w * w + x * x + y * y + z * z);
  w *= n;
  x *= n;
  y *= n;
  z *= n;
  R[0] = 1 - 2 * (y * y + z * z);
  R[1] = 2 * (x * y - w * z);
  R[2] = 2 * (x * z + w * y);
  R[3] = 2 * (x * y + w * z);
  R[4] = 1 - 2 * (x * x + z * z);
  R[5] = 2 * (y * z - w * x);
  R[6] = 2 * (x * z - w * y);
  R[7] = 2 * (y * z + w * x);
  R[8] = 1 - 2 * (x * x + y * y);
}

// mats: n matrices, element (i,j) of matrix p at mats[p*stride + i*ldm + j]
__global__ void k_round_so(int d, long n, const double *__restrict__ mats, long stride, int ldm, double *out) {
  const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  double R[9];
  if (d == 2)
    round_so2(mats + p * stride, ldm, R);
  else
    round_so3(mats + p * stride, ldm, R);
  for (int i = 0; i < d * d; ++i) out[p * d * d + i] = R[i];
}

// Auxiliary variables given (t, l): the point of the central path the solve was certified at (barrier parameter
// InstState::mu_out; the reference's barrier solver returns such interior values too), or, for mu_out = 0, their exact
// minimisers:
//   QCQP: delta_k = rho v / n, rho = 1 - eps(n / r~, mu / (w r~^2))      (mu = 0: proj_ball(v / r~)),  v = t_a - t_b
//   SOCP: delta_k = n + r~ (1 - rho)                                      (mu = 0: max(r~, n))  — the same cost, cone-feasible
__global__ void k_distances(DevProblem P, const double *__restrict__ z, const InstState *st, double *dist) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= P.K) return;
  const int d = P.d;
  const int inst = find_inst(P.rng_off, P.n_inst, k);
  const int Pi = P.pose_off[inst + 1] - P.pose_off[inst];
  const int a = P.rng_a[k], b = P.rng_b[k];
  double v[3], n2 = 0.0;
  for (int r = 0; r < d; ++r) {
    const int ca = (a < Pi) ? P.zoff[inst] + a * P.blk + r * (d + 1) + d : P.zoff[inst] + Pi * P.blk + (a - Pi) * d + r;
    const int cb = (b < Pi) ? P.zoff[inst] + b * P.blk + r * (d + 1) + d : P.zoff[inst] + Pi * P.blk + (b - Pi) * d + r;
    v[r] = z[ca] - z[cb];
    n2 += v[r] * v[r];
  }
  const double nv = sqrt(n2), rr = P.rng_dist[k], mu = st[inst].mu_out;
  double rho = 0.0;  // dist == 0: the delta column is all zeros, any value does
  if (rr > 0.0) rho = (mu > 0.0) ? 1.0 - barrier_eps(nv / rr, mu / (P.rng_w[k] * rr * rr)) : fmin(1.0, nv / rr);
  if (P.relax == SCORE_RELAX_SOCP) {
    dist[k] = (rr > 0.0) ? nv + rr * (1.0 - rho) : nv;
  } else {
    const double sc = (nv > 0.0) ? rho / nv : 0.0;
    for (int r = 0; r < d; ++r) dist[(size_t)k * d + r] = v[r] * sc;
  }
}

// Split z into the contiguous pose / landmark output arrays of the batch.
__global__ void k_split_z(DevProblem P, const double *__restrict__ z, double *poses, double *lms) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.nz) return;
  const int inst = find_inst(P.zoff, P.n_inst, i);
  const int loc = i - P.zoff[inst];
  const int Pi = P.pose_off[inst + 1] - P.pose_off[inst];
  if (loc < Pi * P.blk)
    poses[(size_t)P.pose_off[inst] * P.blk + loc] = z[i];
  else
    lms[(size_t)P.lm_off[inst] * P.d + (loc - Pi * P.blk)] = z[i];
}

}  // namespace score
