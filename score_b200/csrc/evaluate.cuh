// Evaluation step after the path: absolute trajectory error of the estimated
// translations against ground truth after the best rigid SE(d) alignment, batched over trajectories.
//
// The reference ships ground truth beside its inputs (PoseVariable.true_position in the pickles,
// examples/goats_14_data/gt_traj_A.tum) and reports aligned trajectory error in its paper; the alignment itself
// is the Kabsch / Umeyama (scale 1) closed form:
//     R = argmax_{R in SO(d)} tr(R^T H),  H = sum_i (g_i - mean g)(e_i - mean e)^T,   t = mean g - R mean e,
// whose maximiser is the same U diag(1,..,det) V^T rule as round_to_special_orthogonal
// (score/utils/matrix_utils.py:59-79), so the rounding device functions of extract.cuh are reused.
//
// One CTA per trajectory, three fixed-order passes (means, cross-covariance, residual) — no atomics, so a
// trajectory evaluated inside a batch is bit-identical to evaluating it alone.  The passes re-read at most a few
// thousand points that stay in L1/L2; the kernel is a reduction over 16 d bytes per pose.
#pragma once
#include "common.cuh"
#include "extract.cuh"

namespace score {

constexpr int kAteThreads = 256;

// Point i of the estimate lives at est[i * es + c * ecs + ec0] (c < D): the relaxed pose blocks [R|t] of a handle
// (es = d(d+1), ecs = d+1, ec0 = d) or a plain [n, d] array (es = d, ecs = 1, ec0 = 0).  gt is [n, d].
// off: [n_traj + 1] point offsets.  Outputs per trajectory: rmse, R (d x d row-major), t (d), any may be null.
// An empty trajectory gives rmse = NaN, R = I, t = 0.
template <int D>
__global__ void __launch_bounds__(kAteThreads) k_ate(int n_traj, const int *__restrict__ off, const double *__restrict__ est,
                                                    long es, int ecs, int ec0, const double *__restrict__ gt, int align,
                                                    double *out_rmse, double *out_R, double *out_t) {
  __shared__ double red[kAteThreads / 32];
  __shared__ double sh[2 * D + D * D + D];  // mean e, mean g, R, t
  for (int tr = blockIdx.x; tr < n_traj; tr += gridDim.x) {
    const int i0 = off[tr], i1 = off[tr + 1], n = i1 - i0;
    double *me = sh, *mg = sh + D, *R = sh + 2 * D, *t = sh + 2 * D + D * D;
    // ---- pass 1: means
    double acc[2 * D];
#pragma unroll
    for (int c = 0; c < 2 * D; ++c) acc[c] = 0.0;
    if (align) {
      for (int i = i0 + threadIdx.x; i < i1; i += kAteThreads) {
#pragma unroll
        for (int c = 0; c < D; ++c) {
          acc[c] += est[(long)i * es + c * ecs + ec0];
          acc[D + c] += gt[(long)i * D + c];
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 2 * D; ++c) {
      const double tot = block_sum<kAteThreads>(acc[c], red);
      if (threadIdx.x == 0) sh[c] = (n > 0) ? tot / n : 0.0;
    }
    __syncthreads();
    // ---- pass 2: cross-covariance H[a][b] = sum (g_a - mg_a)(e_b - me_b), then the rotation
    double h[D * D];
#pragma unroll
    for (int c = 0; c < D * D; ++c) h[c] = 0.0;
    if (align) {
      for (int i = i0 + threadIdx.x; i < i1; i += kAteThreads) {
        double e[D], g[D];
#pragma unroll
        for (int c = 0; c < D; ++c) {
          e[c] = est[(long)i * es + c * ecs + ec0] - me[c];
          g[c] = gt[(long)i * D + c] - mg[c];
        }
#pragma unroll
        for (int a = 0; a < D; ++a)
#pragma unroll
          for (int b = 0; b < D; ++b) h[a * D + b] += g[a] * e[b];
      }
    }
    double H[D * D];
#pragma unroll
    for (int c = 0; c < D * D; ++c) H[c] = block_sum<kAteThreads>(h[c], red);
    if (threadIdx.x == 0) {
      double Rr[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
      if (D == 2) {
        Rr[1] = 0.0;
        Rr[2] = 0.0;
        Rr[3] = 1.0;
      }
      if (align && n > 0) {
        if (D == 2)
          round_so2(H, D, Rr);
        else
          round_so3(H, D, Rr);
      }
#pragma unroll
      for (int c = 0; c < D * D; ++c) R[c] = Rr[c];
#pragma unroll
      for (int a = 0; a < D; ++a) {
        double v = mg[a];
#pragma unroll
        for (int b = 0; b < D; ++b) v -= Rr[a * D + b] * me[b];
        t[a] = align ? v : 0.0;
      }
    }
    __syncthreads();
    // ---- pass 3: residual
    double sse = 0.0;
    for (int i = i0 + threadIdx.x; i < i1; i += kAteThreads) {
      double e[D];
#pragma unroll
      for (int c = 0; c < D; ++c) e[c] = est[(long)i * es + c * ecs + ec0];
#pragma unroll
      for (int a = 0; a < D; ++a) {
        double v = gt[(long)i * D + a] - t[a];
#pragma unroll
        for (int b = 0; b < D; ++b) v -= R[a * D + b] * e[b];
        sse += v * v;
      }
    }
    const double tot = block_sum<kAteThreads>(sse, red);
    if (threadIdx.x == 0) {
      if (out_rmse) out_rmse[tr] = (n > 0) ? sqrt(tot / n) : nan("");
      if (out_R)
        for (int c = 0; c < D * D; ++c) out_R[(long)tr * D * D + c] = R[c];
      if (out_t)
        for (int c = 0; c < D; ++c) out_t[(long)tr * D + c] = t[c];
    }
    __syncthreads();  // sh / red are reused by the next trajectory
  }
}

}  // namespace score
