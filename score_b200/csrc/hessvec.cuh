// Matrix-free Hessian-vector product of the PCG iteration:  h = B^T H_r B x  and  x'Hx,  straight from the factors.
//
// The assembled pair (B, B^T) costs 12 bytes per stored entry and orientation, and the row-space vector u = H_r B x
// is written by the row pass only to be read back by the column pass.  The factors themselves are far smaller —
// a relative-pose factor (score/utils/gurobi_utils.py:504-526: k||t_j - t_i - R_i t~||^2 + tau||R_j - R_i R~||_F^2) is
// its measurement (d + d^2 doubles), two precisions and two pose indices for d + d^2 rows of d+2 / d+1 entries; a range
// term (:475-501) is two owner indices and its d(d+1)/2 curvature entries for d rows — so the PCG iteration applies
// the operator factor by factor instead (~5x fewer matrix bytes).
//
// Gather form, no atomics: every unknown block (pose: d(d+1) coordinates, landmark: d) owns the list of the factors
// incident on it — built once per solve by a stable device radix sort of (owner, factor) pairs, so the order inside a
// list is the factor order of the reference's objective (:358-377) — and sums their contributions in that fixed
// order.  One thread per pose (its block of x is d(d+1) contiguous doubles, read with 128-bit loads; the chain
// neighbours' blocks are the adjacent threads' own blocks), one CTA-wide fixed-order reduction per landmark.
// A factor's x'Hx term is counted once (relative-pose factors at their `to` pose, ranges at their first owner).
// Bit-reproducible; the assembled CSR pair stays in use for the line-search ticks and for score_get_csr.
#pragma once
#include "common.cuh"

namespace score {

// incidence codes: kind in the top 3 bits, factor id (global over the batch) below
enum IncKind : int { INC_EJ = 0, INC_EI = 1, INC_RA = 2, INC_RB = 3, INC_PR = 4 };
constexpr int kIncShift = 28;
constexpr int kIncMask = (1 << kIncShift) - 1;
constexpr int kPosesPerBlock = 256;  // pose block of the Hessian-vector kernel: one thread per pose
constexpr int kLmPerBlock = 8;       // landmarks per landmark block (handled one after the other by the whole CTA)

// (owner, code) pairs in factor order; owner = global pose index, or P + global landmark index
__global__ void k_inc_fill(DevProblem P, int *__restrict__ keys, int *__restrict__ vals) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < P.E) {
    const int e = (int)t, inst = find_inst(P.edge_off, P.n_inst, e), p0 = P.pose_off[inst];
    keys[2 * e] = p0 + P.edge_i[e];
    vals[2 * e] = (INC_EI << kIncShift) | e;
    keys[2 * e + 1] = p0 + P.edge_j[e];
    vals[2 * e + 1] = (INC_EJ << kIncShift) | e;
  } else if (t < (long)P.E + P.K) {
    const int k = (int)(t - P.E), inst = find_inst(P.rng_off, P.n_inst, k);
    const int p0 = P.pose_off[inst], Pi = P.pose_off[inst + 1] - p0, l0 = P.lm_off[inst];
    const int a = P.rng_a[k], b = P.rng_b[k];
    const int j = 2 * P.E + 2 * k;
    keys[j] = (a < Pi) ? p0 + a : P.P + l0 + (a - Pi);
    vals[j] = (INC_RA << kIncShift) | k;
    keys[j + 1] = (b < Pi) ? p0 + b : P.P + l0 + (b - Pi);
    vals[j + 1] = (INC_RB << kIncShift) | k;
  } else if (t < (long)P.E + P.K + P.Lp) {
    const int q = (int)(t - P.E - P.K), inst = find_inst(P.prior_off, P.n_inst, q);
    const int j = 2 * P.E + 2 * P.K + q;
    keys[j] = P.P + P.lm_off[inst] + P.prior_l[q];
    vals[j] = (INC_PR << kIncShift) | q;
  }
}

// inc_ptr from the sorted owners (same construction as k_transpose_fill)
__global__ void k_inc_ptr(int n, int n_owner, const int *__restrict__ sorted_owner, int *__restrict__ ptr) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int c = sorted_owner[k], cprev = (k == 0) ? -1 : sorted_owner[k - 1];
  for (int cc = cprev + 1; cc <= c; ++cc) ptr[cc] = k;
  if (k == n - 1)
    for (int cc = c + 1; cc <= n_owner; ++cc) ptr[cc] = n;
}

template <int D>
struct PoseBlock {
  double v[D * (D + 1)];
};

// x block of pose `p` (instance-local) of instance with column base `z0`
template <int D>
__device__ __forceinline__ void load_pose(const double *x, int z0, int p, double (&out)[D * (D + 1)]) {
  constexpr int BLK = D * (D + 1);
  const double *src = x + z0 + p * BLK;
  if (D == 2) {  // 48-byte blocks on a 16-byte aligned base (zoff is even in 2D): three 128-bit loads
    const double2 *s2 = reinterpret_cast<const double2 *>(src);
#pragma unroll
    for (int i = 0; i < BLK / 2; ++i) {
      const double2 t = s2[i];
      out[2 * i] = t.x;
      out[2 * i + 1] = t.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < BLK; ++i) out[i] = src[i];
  }
}

// translation of owner `o` (instance-local owner numbering: pose p -> p, landmark q -> Pi + q)
template <int D>
__device__ __forceinline__ void load_trans(const double *x, int z0, int Pi, int o, double (&t)[D]) {
  constexpr int BLK = D * (D + 1);
  if (o < Pi) {
#pragma unroll
    for (int r = 0; r < D; ++r) t[r] = x[z0 + o * BLK + r * (D + 1) + D];
  } else {
#pragma unroll
    for (int r = 0; r < D; ++r) t[r] = x[z0 + Pi * BLK + (o - Pi) * D + r];
  }
}

// u = M_k q for a range (M_k: symmetric d x d, upper row-major, already carrying 2 w)
template <int D>
__device__ __forceinline__ void range_apply(const double *__restrict__ mk, const double (&q)[D], double (&u)[D]) {
#pragma unroll
  for (int a = 0; a < D; ++a) {
    double acc = 0.0;
#pragma unroll
    for (int b = 0; b < D; ++b) {
      const int lo = a < b ? a : b, hi = a < b ? b : a;
      acc += mk[lo * D - lo * (lo - 1) / 2 + (hi - lo)] * q[b];
    }
    u[a] = acc;
  }
}

template <int D>
__device__ __forceinline__ void hessvec_body(DevProblem P, SolverVecs V, BlockTables T, const InstState *st, const int bid) {
  constexpr int D1 = D + 1, BLK = D * D1, NM = D * (D + 1) / 2;
  __shared__ double red[16 * (kThreads / 32)];
  __shared__ double qsh;
  const BlockDesc bd = T.pb[bid];
  const int inst = bd.inst;
  if (st[inst].phase != PH_CG || st[inst].eval_now) return;
  const double *x = V.p;
  const int z0 = P.zoff[inst], p0 = P.pose_off[inst], Pi = P.pose_off[inst + 1] - p0;
  double quad = 0.0;
  if (bd.kind == CB_POSE) {
    const int p = bd.i0 + threadIdx.x;  // instance-local pose
    if (p < bd.i1) {
      double xo[BLK], h[BLK];
      load_pose<D>(x, z0, p, xo);
#pragma unroll
      for (int i = 0; i < BLK; ++i) h[i] = 0.0;
      const int j0 = P.inc_ptr[p0 + p], j1 = P.inc_ptr[p0 + p + 1];
      for (int j = j0; j < j1; ++j) {
        const int code = P.inc_code[j], kind = code >> kIncShift, id = code & kIncMask;
        if (kind == INC_EJ || kind == INC_EI) {
          const bool is_j = kind == INC_EJ;
          double xn[BLK];  // the other pose of the factor
          load_pose<D>(x, z0, is_j ? P.edge_i[id] : P.edge_j[id], xn);
          const double *tm = P.edge_t + (size_t)id * D, *Rm = P.edge_R + (size_t)id * D * D;
          const double k2 = 2.0 * P.edge_k[id], tau2 = 2.0 * P.edge_tau[id];
          double xi[BLK], xj[BLK];  // base pose i, `to` pose j
#pragma unroll
          for (int i = 0; i < BLK; ++i) {
            xi[i] = is_j ? xn[i] : xo[i];
            xj[i] = is_j ? xo[i] : xn[i];
          }
          double ut[D], uR[D * D], qq = 0.0;
#pragma unroll
          for (int r = 0; r < D; ++r) {
            double qt = xj[r * D1 + D] - xi[r * D1 + D];
#pragma unroll
            for (int c = 0; c < D; ++c) qt -= xi[r * D1 + c] * tm[c];
            ut[r] = k2 * qt;
            qq += qt * ut[r];
#pragma unroll
            for (int c = 0; c < D; ++c) {
              double qr = xj[r * D1 + c];
#pragma unroll
              for (int m = 0; m < D; ++m) qr -= xi[r * D1 + m] * Rm[m * D + c];
              uR[r * D + c] = tau2 * qr;
              qq += qr * uR[r * D + c];
            }
          }
          if (is_j) {
            quad += qq;  // the factor's x'Hx term is counted at its `to` pose
#pragma unroll
            for (int r = 0; r < D; ++r) {
              h[r * D1 + D] += ut[r];
#pragma unroll
              for (int c = 0; c < D; ++c) h[r * D1 + c] += uR[r * D + c];
            }
          } else {
#pragma unroll
            for (int r = 0; r < D; ++r) {
              h[r * D1 + D] -= ut[r];
#pragma unroll
              for (int m = 0; m < D; ++m) {
                double acc = ut[r] * tm[m];
#pragma unroll
                for (int c = 0; c < D; ++c) acc += uR[r * D + c] * Rm[m * D + c];
                h[r * D1 + m] -= acc;
              }
            }
          }
        } else if (kind == INC_RA || kind == INC_RB) {
          const bool is_a = kind == INC_RA;
          double tn[D], q[D], u[D];
          load_trans<D>(x, z0, Pi, is_a ? P.rng_b[id] : P.rng_a[id], tn);
#pragma unroll
          for (int r = 0; r < D; ++r) q[r] = is_a ? xo[r * D1 + D] - tn[r] : tn[r] - xo[r * D1 + D];
          range_apply<D>(V.mk + (size_t)id * NM, q, u);
#pragma unroll
          for (int r = 0; r < D; ++r) {
            h[r * D1 + D] += is_a ? u[r] : -u[r];
            if (is_a) quad += q[r] * u[r];
          }
        }
      }
      double *dst = V.h + z0 + p * BLK;
      if (D == 2) {
        double2 *d2 = reinterpret_cast<double2 *>(dst);
#pragma unroll
        for (int i = 0; i < BLK / 2; ++i) d2[i] = make_double2(h[2 * i], h[2 * i + 1]);
      } else {
#pragma unroll
        for (int i = 0; i < BLK; ++i) dst[i] = h[i];
      }
    }
    const double tot = block_sum<kThreads>(quad, red);
    if (threadIdx.x == 0) V.part_hv[bid] = tot;
    return;
  }
  // landmark block: landmarks bd.i0 .. bd.i1 (instance-local), each reduced by the whole CTA in fixed order
  const int l0 = P.lm_off[inst];
  double qtot = 0.0;  // thread 0 only
  for (int q = bd.i0; q < bd.i1; ++q) {
    const int o = Pi + q;
    double xo[D];
    load_trans<D>(x, z0, Pi, o, xo);
    double v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = 0.0;
    const int j0 = P.inc_ptr[P.P + l0 + q], j1 = P.inc_ptr[P.P + l0 + q + 1];
    for (int j = j0 + threadIdx.x; j < j1; j += kThreads) {
      const int code = P.inc_code[j], kind = code >> kIncShift, id = code & kIncMask;
      if (kind == INC_PR) {  // w ||l - prior||^2: curvature 2 w
        const double w2 = 2.0 * P.prior_w[id];
#pragma unroll
        for (int r = 0; r < D; ++r) {
          v[r] += w2 * xo[r];
          v[D] += w2 * xo[r] * xo[r];
        }
        continue;
      }
      const bool is_a = kind == INC_RA;
      double tn[D], qv[D], u[D];
      load_trans<D>(x, z0, Pi, is_a ? P.rng_b[id] : P.rng_a[id], tn);
#pragma unroll
      for (int r = 0; r < D; ++r) qv[r] = is_a ? xo[r] - tn[r] : tn[r] - xo[r];
      range_apply<D>(V.mk + (size_t)id * NM, qv, u);
#pragma unroll
      for (int r = 0; r < D; ++r) {
        v[r] += is_a ? u[r] : -u[r];
        if (is_a) v[D] += qv[r] * u[r];
      }
    }
    const double tot = block_sum16<kThreads>(v, red);  // thread i < 16 holds the total of v[i]
    if (threadIdx.x < D) V.h[z0 + Pi * BLK + q * D + threadIdx.x] = tot;
    if (threadIdx.x == D) qsh = tot;
    __syncthreads();
    if (threadIdx.x == 0) qtot += qsh;
  }
  if (threadIdx.x == 0) V.part_hv[bid] = qtot;
}

template <int D>
__global__ void __launch_bounds__(kThreads) k_hessvec(DevProblem P, SolverVecs V, BlockTables T, const InstState *st, WorkLists W) {
  const int *act;
  int n_act;
  wl_get(W, WL_RUN, act, n_act);
  for (long long item = blockIdx.x; item < (long long)n_act * W.maxpb; item += gridDim.x) {
    const int inst = act[item / W.maxpb], bid = T.pb_begin[inst] + (int)(item % W.maxpb);
    if (bid < T.pb_begin[inst + 1]) hessvec_body<D>(P, V, T, st, bid);
    __syncthreads();
  }
}

}  // namespace score
