// Dense SPD inverse in global memory: blocked symmetric sweeps.
//
// Coarse spaces that do not fit the on-chip register-tiled solve of coarse.cuh (kCoarseMax < nc <= kCoarseBigMax)
// are inverted here.  The sweep operator on a pivot block K (W = A_KK^-1)
//     A_KK <- -W,   A_KJ <- W A_KJ,   A_IK <- A_IK W,   A_IJ <- A_IJ - A_IK W A_KJ
// is the composition of the scalar symmetric sweeps of the block's pivots, so after all blocks the matrix holds
// -A^-1 (the same operator coarse_build_body applies in registers).  Per pivot block of kDB = 32 pivots:
//   k_dense_pivot   one CTA inverts the 32 x 32 pivot block in shared memory                         -> W
//   k_dense_panel   raw row panel A_K,: (with A_KK - I in the pivot columns) and scaled panel W A_K,:  -> raw, scl
//   k_dense_update  rank-32 update of the whole matrix with the two panels (64 x 64 tiles, 4 x 4 per thread); with
//                   A_KK - I published in place of A_KK the generic update also produces the new pivot rows and
//                   columns, and only the pivot block itself is patched to -W.
// k_dense_update is a general C -= X^T Y kernel (k-major panels); the Schur complement of the landmark block
// (coarse.cuh) reuses it with kdim = landmark coordinates.
// Every kernel is gated on the device by the owning instance's phase (the launches are static graph nodes).
#pragma once
#include "common.cuh"

namespace score {

constexpr int kDB = 32;  // pivots per block
constexpr int kDT = 64;  // update tile

// gate < 0: always on (stand-alone use); otherwise instance `gate` must be taking a line-search tick
__device__ __forceinline__ bool dense_gate(const InstState *st, int gate) {
  return gate < 0 || (st[gate].phase == PH_LS && !st[gate].eval_now);
}

__global__ void __launch_bounds__(kDB * kDB) k_dense_pivot(const double *__restrict__ A, int n, int lda, int kt,
                                                          double *__restrict__ W, const InstState *st, int gate) {
  if (!dense_gate(st, gate)) return;
  __shared__ double a[kDB][kDB + 1];
  const int i = threadIdx.y, j = threadIdx.x, gi = kt * kDB + i, gj = kt * kDB + j;
  double v = (gi < n && gj < n) ? A[(size_t)gi * lda + gj] : (i == j ? 1.0 : 0.0);
  a[i][j] = v;
  __syncthreads();
  for (int k = 0; k < kDB; ++k) {
    const double piv = a[k][k], aik = a[i][k], akj = a[k][j];
    __syncthreads();
    const double inv = 1.0 / piv;
    if (i == k && j == k)
      v = -inv;
    else if (i == k || j == k)
      v *= inv;
    else
      v -= aik * akj * inv;
    a[i][j] = v;
    __syncthreads();
  }
  W[i * kDB + j] = -v;
}

__global__ void __launch_bounds__(256) k_dense_panel(const double *__restrict__ A, int n, int lda, int kt,
                                                    const double *__restrict__ W, double *__restrict__ raw,
                                                    double *__restrict__ scl, int ldp, const InstState *st, int gate) {
  if (!dense_gate(st, gate)) return;
  __shared__ double w[kDB][kDB + 1];
  for (int t = threadIdx.x; t < kDB * kDB; t += 256) w[t / kDB][t % kDB] = W[t];
  __syncthreads();
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j >= n) return;
  const int k0 = kt * kDB;
  double a[kDB];
#pragma unroll
  for (int k = 0; k < kDB; ++k) {
    const int gi = k0 + k;
    a[k] = (gi < n) ? A[(size_t)gi * lda + j] : 0.0;
    if (gi == j) a[k] -= 1.0;
  }
#pragma unroll 4
  for (int r = 0; r < kDB; ++r) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < kDB; ++k) s += w[r][k] * a[k];
    scl[(size_t)r * ldp + j] = s;
  }
#pragma unroll
  for (int r = 0; r < kDB; ++r) raw[(size_t)r * ldp + j] = a[r];
}

// C[i][j] -= sum_{k < kdim} X[k][i] Y[k][j]   (i, j < n); entries of pivot block kt are set to -W when W != nullptr.
__global__ void __launch_bounds__(256) k_dense_update(double *__restrict__ C, int n, int ldc, const double *__restrict__ X,
                                                     const double *__restrict__ Y, int ldp, int kdim, int kt,
                                                     const double *__restrict__ W, const InstState *st, int gate) {
  if (!dense_gate(st, gate)) return;
  __shared__ __align__(16) double sx[kDB][kDT], sy[kDB][kDT];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int i0 = blockIdx.y * kDT, j0 = blockIdx.x * kDT;
  double acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.0;
  for (int kc = 0; kc < kdim; kc += kDB) {
    for (int t = threadIdx.x; t < kDB * kDT; t += 256) {
      const int k = t / kDT, q = t % kDT;
      const bool kok = kc + k < kdim;
      sx[k][q] = (kok && i0 + q < n) ? X[(size_t)(kc + k) * ldp + i0 + q] : 0.0;
      sy[k][q] = (kok && j0 + q < n) ? Y[(size_t)(kc + k) * ldp + j0 + q] : 0.0;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < kDB; ++k) {
      const double2 xa = *reinterpret_cast<const double2 *>(&sx[k][ty * 4]);
      const double2 xb = *reinterpret_cast<const double2 *>(&sx[k][ty * 4 + 2]);
      const double2 ya = *reinterpret_cast<const double2 *>(&sy[k][tx * 4]);
      const double2 yb = *reinterpret_cast<const double2 *>(&sy[k][tx * 4 + 2]);
      const double xr[4] = {xa.x, xa.y, xb.x, xb.y}, yc[4] = {ya.x, ya.y, yb.x, yb.y};
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] += xr[r] * yc[c];
    }
    __syncthreads();
  }
  const int p0 = kt * kDB;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = i0 + ty * 4 + r;
    if (i >= n) continue;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int j = j0 + tx * 4 + c;
      if (j >= n) continue;
      double v = C[(size_t)i * ldc + j] - acc[r][c];
      if (W && i >= p0 && i < p0 + kDB && j >= p0 && j < p0 + kDB) v = -W[(i - p0) * kDB + (j - p0)];
      C[(size_t)i * ldc + j] = v;
    }
  }
}

// out = -(A + A^T) / 2  (the swept matrix holds -A^-1; exactly symmetric result)
__global__ void __launch_bounds__(256) k_dense_finish(const double *__restrict__ A, int n, int lda, double *__restrict__ out,
                                                     int ldo, const InstState *st, int gate) {
  if (!dense_gate(st, gate)) return;
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  if (t >= (long long)n * n) return;
  const int i = (int)(t / n), j = (int)(t % n);
  out[(size_t)i * ldo + j] = -0.5 * (A[(size_t)i * lda + j] + A[(size_t)j * lda + i]);
}

// ---- Schur complement of a block-diagonal landmark block -----------------------------------------------------
// Coarse matrix A = [[S11, S12], [S12^T, Dl]] with S11 the free segment bases (nb coordinates), Dl the landmarks
// (nl coordinates).  When no range term joins two landmarks Dl is block diagonal (one d x d block per landmark) and
//     A^-1 c :  v = Dl^-1 c_l ;  y_s = T^-1 (c_s - S12 v) ;  y_l = v - Dl^-1 S12^T y_s ,      T = S11 - S12 Dl^-1 S12^T ,
// so only the nb x nb matrix T is swept (BASELINE configs[4]: nb = 1 188 of nc = 4 188: 44x fewer flops than the full
// sweep) and an application reads T^-1, S21 = S12^T and Y = Dl^-1 S21 (68 MB instead of 140 MB).

// S21[l][i] = A[i][nb + l],  Y = Dl^-1 S21 (both k-major panels, row stride ldp),  Dinv[q] = (d x d block q of Dl)^-1.
template <int D>
__global__ void __launch_bounds__(256) k_schur_prep(const double *__restrict__ A, int n, int nb, double *__restrict__ S21,
                                                   double *__restrict__ Y, double *__restrict__ Dinv, int ldp,
                                                   const InstState *st, int gate) {
  if (!dense_gate(st, gate)) return;
  const int nq = (n - nb) / D;
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  if (t >= (long long)nq * nb) return;
  const int q = (int)(t / nb), i = (int)(t % nb);
  double Dq[D * D], inv[D * D];
#pragma unroll
  for (int r = 0; r < D; ++r)
#pragma unroll
    for (int c = 0; c < D; ++c) {
      Dq[r * D + c] = A[(size_t)(nb + q * D + r) * n + nb + q * D + c];
      inv[r * D + c] = (r == c) ? 1.0 : 0.0;
    }
#pragma unroll
  for (int c = 0; c < D; ++c) {  // Gauss-Jordan on the SPD block
    const double piv = 1.0 / Dq[c * D + c];
#pragma unroll
    for (int j = 0; j < D; ++j) {
      Dq[c * D + j] *= piv;
      inv[c * D + j] *= piv;
    }
#pragma unroll
    for (int r = 0; r < D; ++r) {
      if (r == c) continue;
      const double f = Dq[r * D + c];
#pragma unroll
      for (int j = 0; j < D; ++j) {
        Dq[r * D + j] -= f * Dq[c * D + j];
        inv[r * D + j] -= f * inv[c * D + j];
      }
    }
  }
  double sv[D];
#pragma unroll
  for (int m = 0; m < D; ++m) sv[m] = A[(size_t)i * n + nb + q * D + m];
#pragma unroll
  for (int r = 0; r < D; ++r) {
    double acc = 0.0;
#pragma unroll
    for (int m = 0; m < D; ++m) acc += inv[r * D + m] * sv[m];
    S21[(size_t)(q * D + r) * ldp + i] = sv[r];
    Y[(size_t)(q * D + r) * ldp + i] = acc;
  }
  if (i == 0) {
#pragma unroll
    for (int e = 0; e < D * D; ++e) Dinv[(size_t)q * D * D + e] = inv[e];
  }
}

constexpr int kSchurChunks = 64;  // landmark-coordinate chunks of the S12 v product (fixed-order two-stage sum)

// v = Dl^-1 c_l for landmark coordinate l
template <int D>
__device__ __forceinline__ double schur_v(const double *__restrict__ Dinv, const double *cl, int l) {
  const int q = l / D, r = l % D;
  double acc = 0.0;
#pragma unroll
  for (int m = 0; m < D; ++m) acc += Dinv[(size_t)q * D * D + r * D + m] * cl[q * D + m];
  return acc;
}

// part[ch][i] = sum over the landmark coordinates l of chunk ch of S21[l][i] v[l]
template <int D>
__global__ void __launch_bounds__(256) k_schur_apply1(const double *__restrict__ S21, const double *__restrict__ Dinv,
                                                     const double *crhs, int nb, int nl, int ldp, double *__restrict__ part,
                                                     const InstState *st, int inst) {
  if (st[inst].phase == PH_DONE || st[inst].phase == PH_WAIT || st[inst].eval_now) return;
  __shared__ double vs[512];
  const int per = (nl + kSchurChunks - 1) / kSchurChunks;
  const int l0 = blockIdx.y * per, l1 = min(nl, l0 + per);
  const int i = blockIdx.x * 256 + threadIdx.x;
  double acc = 0.0;
  for (int lb = l0; lb < l1; lb += 512) {
    const int cnt = min(512, l1 - lb);
    __syncthreads();
    for (int t = threadIdx.x; t < cnt; t += 256) vs[t] = schur_v<D>(Dinv, crhs + nb, lb + t);
    __syncthreads();
    if (i < nb)
      for (int t = 0; t < cnt; ++t) acc += S21[(size_t)(lb + t) * ldp + i] * vs[t];
  }
  if (i < nb) part[(size_t)blockIdx.y * nb + i] = acc;
}

// w = c_s - sum over the chunks (fixed order)
__global__ void __launch_bounds__(256) k_schur_apply2(const double *crhs, const double *__restrict__ part, int nb,
                                                     double *__restrict__ w, const InstState *st, int inst) {
  if (st[inst].phase == PH_DONE || st[inst].phase == PH_WAIT || st[inst].eval_now) return;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= nb) return;
  double acc = 0.0;
#pragma unroll
  for (int ch = 0; ch < kSchurChunks; ++ch) acc += part[(size_t)ch * nb + i];
  w[i] = crhs[i] - acc;
}

// out[row] = sum_j M[row][j] x[j]  (one warp per row); SUBV: out[row] = v[row] - that sum with v = Dl^-1 c_l
template <int D, bool SUBV>
__global__ void __launch_bounds__(256) k_schur_matvec(const double *__restrict__ M, int rows, int cols, int ldm, const double *x,
                                                     double *__restrict__ out, const double *__restrict__ Dinv, const double *cl,
                                                     const InstState *st, int inst) {
  if (st[inst].phase == PH_DONE || st[inst].phase == PH_WAIT || st[inst].eval_now) return;
  const int lane = threadIdx.x & 31, row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const double *__restrict__ a = M + (size_t)row * ldm;
  double acc0 = 0.0, acc1 = 0.0;
  int j = lane;
  for (; j + 32 < cols; j += 64) {
    acc0 += __ldg(a + j) * x[j];
    acc1 += __ldg(a + j + 32) * x[j + 32];
  }
  if (j < cols) acc0 += __ldg(a + j) * x[j];
  const double tot = warp_sum(acc0 + acc1);
  if (lane == 0) out[row] = SUBV ? schur_v<D>(Dinv, cl, row) - tot : tot;
}

// Host side: in-place sweep of A (n x n, row stride lda) -> -A^-1.  W: kDB*kDB doubles, raw/scl: kDB x ldp panels.
// Returns the number of kernels launched.
inline int launch_dense_sweep(double *A, int n, int lda, double *W, double *raw, double *scl, int ldp,
                              const InstState *st, int gate, cudaStream_t s) {
  const int nblk = (n + kDB - 1) / kDB, nt = (n + kDT - 1) / kDT;
  for (int kt = 0; kt < nblk; ++kt) {
    k_dense_pivot<<<1, dim3(kDB, kDB), 0, s>>>(A, n, lda, kt, W, st, gate);
    k_dense_panel<<<(n + 255) / 256, 256, 0, s>>>(A, n, lda, kt, W, raw, scl, ldp, st, gate);
    k_dense_update<<<dim3(nt, nt), 256, 0, s>>>(A, n, lda, raw, scl, ldp, kDB, kt, W, st, gate);
  }
  return 3 * nblk;
}

}  // namespace score
