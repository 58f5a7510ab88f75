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

// incidence records (16 bytes, one 128-bit load): x = kind (top 3 bits) | factor id (global over the batch); y / z =
// where the owner's / the partner's translation lives: the z-relative column of its first coordinate with bit 31 set
// for a pose (coordinate stride d+1; landmark: stride 1); for a relative-pose factor z is the instance-local index of
// the other pose.  Odometry links (the factor (p-1 -> p) that link_edge names) are NOT in the
// lists: the kernel takes them straight from link_edge, with the neighbours' blocks at p-1 / p+1.
enum IncKind : int { INC_EJ = 0, INC_EI = 1, INC_RA = 2, INC_RB = 3, INC_PR = 4 };
constexpr int kIncShift = 28;
constexpr int kIncMask = (1 << kIncShift) - 1;
constexpr int kPosesPerBlock = 256;  // pose block of the Hessian-vector kernel: one thread per pose
constexpr int kLmPerBlock = 8;       // landmarks per landmark block (handled one after the other by the whole CTA)
constexpr int kHvTile = 1024;         // range records whose contributions a pose block stages in shared memory at a time
typedef uint4 IncRec;                 // x: kind | id, y: owner column record, z: partner column record / other pose

__device__ __forceinline__ IncRec inc_make(int kind, int id, unsigned own, unsigned other) {
  return make_uint4((unsigned)((kind << kIncShift) | id), own, other, 0u);
}

// (owner, record) pairs in factor order; owner = global pose index, or P + global landmark index; odometry links get
// the sentinel owner P + L (sorted past every real list)
__global__ void k_inc_fill(DevProblem P, int *__restrict__ keys, IncRec *__restrict__ vals) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int blk = P.blk, d = P.d;
  if (t < P.E) {
    const int e = (int)t, inst = find_inst(P.edge_off, P.n_inst, e), p0 = P.pose_off[inst];
    const int i = P.edge_i[e], j = P.edge_j[e];
    const bool link = P.link_edge[p0 + j] == e;
    keys[2 * e] = link ? P.P + P.L : p0 + i;
    vals[2 * e] = inc_make(INC_EI, e, 0u, (unsigned)j);
    keys[2 * e + 1] = link ? P.P + P.L : p0 + j;
    vals[2 * e + 1] = inc_make(INC_EJ, e, 0u, (unsigned)i);
  } else if (t < (long)P.E + P.K) {
    const int k = (int)(t - P.E), inst = find_inst(P.rng_off, P.n_inst, k);
    const int p0 = P.pose_off[inst], Pi = P.pose_off[inst + 1] - p0, l0 = P.lm_off[inst];
    const int a = P.rng_a[k], b = P.rng_b[k];
    auto tcol = [&](int o) -> unsigned {
      return (o < Pi) ? (0x80000000u | (unsigned)(o * blk + d)) : (unsigned)(Pi * blk + (o - Pi) * d);
    };
    const int j = 2 * P.E + 2 * k;
    keys[j] = (a < Pi) ? p0 + a : P.P + l0 + (a - Pi);
    vals[j] = inc_make(INC_RA, k, tcol(a), tcol(b));
    keys[j + 1] = (b < Pi) ? p0 + b : P.P + l0 + (b - Pi);
    vals[j + 1] = inc_make(INC_RB, k, tcol(b), tcol(a));
  } else if (t < (long)P.E + P.K + P.Lp) {
    const int q = (int)(t - P.E - P.K), inst = find_inst(P.prior_off, P.n_inst, q);
    const int j = 2 * P.E + 2 * P.K + q;
    keys[j] = P.P + P.lm_off[inst] + P.prior_l[q];
    vals[j] = inc_make(INC_PR, q, 0u, 0u);
  }
}

// inc_ptr from the sorted owners (same construction as k_transpose_fill); owners > n_owner - 1 (the sentinel) end
// the last real list
__global__ void k_inc_ptr(int n, int n_owner, const int *__restrict__ sorted_owner, int *__restrict__ ptr) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int c = min(sorted_owner[k], n_owner), cprev = (k == 0) ? -1 : min(sorted_owner[k - 1], n_owner);
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
__device__ __forceinline__ void range_apply(const double *mk, const double (&q)[D], double (&u)[D]) {
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

// translation at z-relative column record `other` (bit 31: pose, stride d+1)
template <int D>
__device__ __forceinline__ void load_trans_rec(const double *xz, unsigned other, double (&t)[D]) {
  const int col = (int)(other & 0x7fffffffu), stride = (other >> 31) ? D + 1 : 1;
#pragma unroll
  for (int r = 0; r < D; ++r) t[r] = xz[col + r * stride];
}

// Contribution of one relative-pose factor (measurement tm, Rm; doubled precisions k2, tau2) to the h block of one of its
// poses.  xi: base pose block, xj: `to` pose block.  AT_J: accumulate the `to` pose's part (and the factor's x'Hx
// term), else the base pose's part.
template <int D, bool AT_J>
__device__ __forceinline__ void edge_contrib(const double (&xi)[D * (D + 1)], const double (&xj)[D * (D + 1)],
                                             const double (&tm)[D], const double (&Rm)[D * D], double k2, double tau2,
                                             double (&h)[D * (D + 1)], double &quad) {
  constexpr int D1 = D + 1;
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
  if (AT_J) {
    quad += qq;
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
}

template <int D>
__device__ __forceinline__ void load_edge(const DevProblem &P, int e, double (&tm)[D], double (&Rm)[D * D], double &k2,
                                          double &tau2) {
  const double *t = P.edge_t + (size_t)e * D, *R = P.edge_R + (size_t)e * D * D;
  if (D == 2) {
    const double2 a = *reinterpret_cast<const double2 *>(t);
    const double2 b = reinterpret_cast<const double2 *>(R)[0], c = reinterpret_cast<const double2 *>(R)[1];
    tm[0] = a.x, tm[1] = a.y;
    Rm[0] = b.x, Rm[1] = b.y, Rm[2] = c.x, Rm[3] = c.y;
  } else {
#pragma unroll
    for (int i = 0; i < D; ++i) tm[i] = t[i];
#pragma unroll
    for (int i = 0; i < D * D; ++i) Rm[i] = R[i];
  }
  k2 = 2.0 * P.edge_k[e];
  tau2 = 2.0 * P.edge_tau[e];
}

// Contribution of a range term to its OWNER's translation block: c = M_k (t_owner - t_partner)  (for the first owner
// this is +M_k (t_a - t_b), for the second -M_k (t_a - t_b): the same expression).  Returns d.c, the term's x'Hx.
template <int D>
__device__ __forceinline__ double range_contrib(const double *xz, const double *mk_all, const IncRec rec, double (&c)[D]) {
  constexpr int NM = D * (D + 1) / 2;
  const int id = (int)(rec.x & kIncMask);
  double to[D], tp[D], m[NM], dv[D];
  load_trans_rec<D>(xz, rec.y, to);
  load_trans_rec<D>(xz, rec.z, tp);
#pragma unroll
  for (int i = 0; i < NM; ++i) m[i] = mk_all[(size_t)id * NM + i];
#pragma unroll
  for (int r = 0; r < D; ++r) dv[r] = to[r] - tp[r];
  range_apply<D>(m, dv, c);
  double dc = 0.0;
#pragma unroll
  for (int r = 0; r < D; ++r) dc += dv[r] * c[r];
  return dc;
}

// The descriptor carries everything the CTA needs to address its loads, so the instance state, the links, the list
// bounds, the own x block and the first incidence record of every thread go out in ONE round (all addresses are valid
// whatever the state says); the state is looked at after they have been issued.
template <int D>
__device__ __forceinline__ void hessvec_body(DevProblem P, SolverVecs V, const InstState *st, const HvBlock bd) {
  constexpr int D1 = D + 1, BLK = D * D1;
  __shared__ double red[16 * (kThreads / 32)];
  __shared__ double qsh;
  __shared__ double contrib[kHvTile * D];
  const int inst = bd.inst, bid = bd.bid;
  const int phase = st[inst].phase, evn = st[inst].eval_now;
  const int z0 = bd.z0, Pi = bd.Pi;
  const double *xz = V.p + z0;  // this instance's block of the direction vector
  const IncRec *__restrict__ recs = reinterpret_cast<const IncRec *>(P.inc_rec);
  const int tid = threadIdx.x;
  double quad = 0.0;
  if (bd.kind == CB_POSE) {
    const int np = bd.i1 - bd.i0, pgf = bd.pg0 + bd.i0;  // poses of this block, global index of the first one
    const int p = bd.i0 + tid;                           // instance-local pose of this thread
    const bool active = tid < np;
    const int jb = bd.jb, je = bd.je;
    int j0 = 0, j1 = 0, e_in = -1, e_out = -1;
    double h[BLK], xo[BLK];
    IncRec pre = make_uint4(0u, 0u, 0u, 0u);  // this thread's first record of the block's run
    if (jb + tid < je) pre = recs[jb + tid];
#pragma unroll
    for (int i = 0; i < BLK; ++i) h[i] = 0.0;
    if (active) {
      // odometry links straight from link_edge: the neighbours' blocks are at p - 1 / p + 1
      e_in = P.link_edge[pgf + tid];
      e_out = (p + 1 < Pi) ? P.link_edge[pgf + tid + 1] : -1;
      j0 = P.inc_ptr[pgf + tid];
      j1 = P.inc_ptr[pgf + tid + 1];
      load_pose<D>(xz, 0, p, xo);
    }
    if (phase != PH_CG || evn) return;
    if (active) {
      if (e_in >= 0) {  // (p-1 -> p): this pose is the `to` pose
        double xn[BLK], tm[D], Rm[D * D], k2, tau2;
        load_pose<D>(xz, 0, p - 1, xn);
        load_edge<D>(P, e_in, tm, Rm, k2, tau2);
        edge_contrib<D, true>(xn, xo, tm, Rm, k2, tau2, h, quad);
      }
      if (e_out >= 0) {  // (p -> p+1): this pose is the base pose
        double xn[BLK], tm[D], Rm[D * D], k2, tau2;
        load_pose<D>(xz, 0, p + 1, xn);
        load_edge<D>(P, e_out, tm, Rm, k2, tau2);
        edge_contrib<D, false>(xo, xn, tm, Rm, k2, tau2, h, quad);
      }
    }
    // range terms of the block's poses: their records are one contiguous run [jb, je).  Phase 1: all threads stride
    // over the run (coalesced 128-bit record loads, independent gathers) and stage each term's contribution in shared
    // memory; phase 2: every pose adds the contributions of its own list in list order.
    for (int t0 = jb; t0 < je; t0 += kHvTile) {
      const int t1 = min(je, t0 + kHvTile);
#pragma unroll 2
      for (int j = t0 + tid; j < t1; j += kThreads) {
        const IncRec rec = (j == jb + tid) ? pre : recs[j];
        const int kind = (int)(rec.x >> kIncShift);
        if (kind == INC_RA || kind == INC_RB) {
          double c[D];
          const double dc = range_contrib<D>(xz, V.mk, rec, c);
          if (kind == INC_RA) quad += dc;  // the term's x'Hx is counted at its first owner
#pragma unroll
          for (int r = 0; r < D; ++r) contrib[(j - t0) * D + r] = c[r];
        }
      }
      __syncthreads();
      if (active) {
        for (int j = max(j0, t0); j < min(j1, t1); ++j) {
          if (P.n_nonlink > 0) {  // relative-pose factors that are not odometry links (loop closures): rare, done in place
            const IncRec rec = recs[j];
            const int kind = (int)(rec.x >> kIncShift), id = (int)(rec.x & kIncMask);
            if (kind == INC_EJ || kind == INC_EI) {
              double xn[BLK], tm[D], Rm[D * D], k2, tau2;
              load_pose<D>(xz, 0, (int)rec.z, xn);
              load_edge<D>(P, id, tm, Rm, k2, tau2);
              if (kind == INC_EJ)
                edge_contrib<D, true>(xn, xo, tm, Rm, k2, tau2, h, quad);
              else
                edge_contrib<D, false>(xo, xn, tm, Rm, k2, tau2, h, quad);
              continue;
            }
          }
#pragma unroll
          for (int r = 0; r < D; ++r) h[r * D1 + D] += contrib[(j - t0) * D + r];
        }
      }
      __syncthreads();
    }
    if (active) {
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
    if (tid == 0) V.part_hv[bid] = tot;
    return;
  }
  // landmark block: landmarks bd.i0 .. bd.i1 (instance-local), each reduced by the whole CTA in fixed order
  if (phase != PH_CG || evn) return;
  double qtot = 0.0;  // thread 0 only
  for (int q = bd.i0; q < bd.i1; ++q) {
    double v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = 0.0;
    const int j0 = P.inc_ptr[P.P + bd.lg0 + q], j1 = P.inc_ptr[P.P + bd.lg0 + q + 1];
#pragma unroll 2
    for (int j = j0 + tid; j < j1; j += kThreads) {
      const IncRec rec = recs[j];
      const int kind = (int)(rec.x >> kIncShift), id = (int)(rec.x & kIncMask);
      if (kind == INC_PR) {  // w ||l - prior||^2: curvature 2 w
        const double w2 = 2.0 * P.prior_w[id];
#pragma unroll
        for (int r = 0; r < D; ++r) {
          const double xr = xz[Pi * BLK + q * D + r];
          v[r] += w2 * xr;
          v[D] += w2 * xr * xr;
        }
        continue;
      }
      double c[D];
      const double dc = range_contrib<D>(xz, V.mk, rec, c);
#pragma unroll
      for (int r = 0; r < D; ++r) v[r] += c[r];
      if (kind == INC_RA) v[D] += dc;
    }
    const double tot = block_sum16<kThreads>(v, red);  // thread i < 16 holds the total of v[i]
    if (tid < D) V.h[z0 + Pi * BLK + q * D + tid] = tot;
    if (tid == D) qsh = tot;
    __syncthreads();
    if (tid == 0) qtot += qsh;
  }
  if (tid == 0) V.part_hv[bid] = qtot;
}

// ---- The line-search ticks' two operator applications, factor by factor as well -----------------------------------
// (1) rows: out = B x (- b) for the rows of one row block — bdz = B dz before the line search, res = B z - b at the
//     start — one thread per factor, the factor's rows in the reference's order (assemble.cuh).
// (2) gradient: h = B^T u gathered per pose / landmark from the row-space vector u = dF/d(res) that k_rowupdate wrote,
//     fused with what the column pass did with it (new point z += step dz, dz = 0, r = -h; or the certificate's sums).
// With these the assembled CSR pair is not touched by a matrix-free solve at all (it is still built for
// score_get_csr, for operator_mode = 1 and for the row-partitioned multi-GPU solve).
enum RowsMode : int { RM_BDZ = 0, RM_RES = 1 };

template <int D>
__device__ __forceinline__ void rows_mf_body(DevProblem P, SolverVecs V, BlockTables T, const InstState *st, const int bid,
                                             const int mode) {
  constexpr int D1 = D + 1, BLK = D * D1, RPE = D + D * D;
  const BlockDesc bd = T.rb[bid];
  const int inst = bd.inst;
  if (mode == RM_BDZ && (st[inst].phase != PH_LS || st[inst].skip_ls || st[inst].eval_now)) return;
  const double *xz = (mode == RM_BDZ ? V.dz : V.z) + P.zoff[inst];
  double *out = mode == RM_BDZ ? V.bdz : V.res;
  const int r0 = P.roff[inst], e0 = P.edge_off[inst], Ei = P.edge_off[inst + 1] - e0;
  const int k0 = P.rng_off[inst], Ki = P.rng_off[inst + 1] - k0, q0 = P.prior_off[inst];
  const int Pi = P.pose_off[inst + 1] - P.pose_off[inst];
  const int rr0 = r0 + Ei * RPE, rp0 = rr0 + Ki * D;
  // relative-pose factors whose rows lie in this block (kRowsPerBlock is a multiple of the rows per factor)
  {
    const int a = max(bd.i0, r0), b = min(bd.i1, rr0);
    for (int el = (a - r0) / RPE + threadIdx.x; el < (b - r0) / RPE; el += kThreads) {
      const int e = e0 + el;
      double xi[BLK], xj[BLK], tm[D], Rm[D * D], k2, tau2;
      load_pose<D>(xz, 0, P.edge_i[e], xi);
      load_pose<D>(xz, 0, P.edge_j[e], xj);
      load_edge<D>(P, e, tm, Rm, k2, tau2);
      double *o = out + r0 + (size_t)el * RPE;
#pragma unroll
      for (int r = 0; r < D; ++r) {
        double qt = xj[r * D1 + D] - xi[r * D1 + D];
#pragma unroll
        for (int c = 0; c < D; ++c) qt -= xi[r * D1 + c] * tm[c];
        o[r] = qt;
#pragma unroll
        for (int c = 0; c < D; ++c) {
          double qr = xj[r * D1 + c];
#pragma unroll
          for (int m = 0; m < D; ++m) qr -= xi[r * D1 + m] * Rm[m * D + c];
          o[D + r * D + c] = qr;
        }
      }
    }
  }
  {  // range terms: rows t_a - t_b
    const int a = max(bd.i0, rr0), b = min(bd.i1, rp0);
    for (int kl = (a - rr0) / D + threadIdx.x; kl < (b - rr0) / D; kl += kThreads) {
      const int k = k0 + kl;
      double ta[D], tb[D];
      load_trans<D>(xz, 0, Pi, P.rng_a[k], ta);
      load_trans<D>(xz, 0, Pi, P.rng_b[k], tb);
#pragma unroll
      for (int r = 0; r < D; ++r) out[rr0 + (size_t)kl * D + r] = ta[r] - tb[r];
    }
  }
  {  // landmark priors: rows l - prior
    const int a = max(bd.i0, rp0), b = bd.i1;
    for (int ql = (a - rp0) / D + threadIdx.x; ql < (b - rp0) / D; ql += kThreads) {
      const int q = q0 + ql, lq = P.prior_l[q];
#pragma unroll
      for (int r = 0; r < D; ++r)
        out[rp0 + (size_t)ql * D + r] = xz[Pi * BLK + lq * D + r] - (mode == RM_RES ? P.prior_t[(size_t)q * D + r] : 0.0);
    }
  }
}

template <int D>
__global__ void __launch_bounds__(kThreads) k_rows_mf(DevProblem P, SolverVecs V, BlockTables T, const InstState *st, WorkLists W,
                                                     int mode) {
  const int *act;
  int n_act;
  wl_get(W, WL_RUN, act, n_act);
  for (long long item = blockIdx.x; item < (long long)n_act * W.maxrb; item += gridDim.x) {
    const int inst = act[item / W.maxrb], bid = T.rb_begin[inst] + (int)(item % W.maxrb);
    if (bid < T.rb_begin[inst + 1]) rows_mf_body<D>(P, V, T, st, bid, mode);
  }
}

// the same over every row block of the batch (initial residual: no work lists yet)
template <int D>
__global__ void __launch_bounds__(kThreads) k_rows_mf_all(DevProblem P, SolverVecs V, BlockTables T, const InstState *st, int mode) {
  for (int bid = blockIdx.x; bid < T.n_rb; bid += gridDim.x) rows_mf_body<D>(P, V, T, st, bid, mode);
}

// weights and right-hand sides of the rows (what the assembly writes beside the matrix)
__global__ void k_row_weights(DevProblem P, double *__restrict__ w, double *__restrict__ b) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int d = P.d, rpe = P.rpe;
  if (t < P.E) {
    const int e = (int)t, inst = find_inst(P.edge_off, P.n_inst, e);
    const int row = P.roff[inst] + (e - P.edge_off[inst]) * rpe;
    for (int r = 0; r < rpe; ++r) {
      w[row + r] = r < d ? P.edge_k[e] : P.edge_tau[e];
      b[row + r] = 0.0;
    }
  } else if (t < (long)P.E + P.K) {
    const int k = (int)(t - P.E), inst = find_inst(P.rng_off, P.n_inst, k);
    const int row = P.roff[inst] + (P.edge_off[inst + 1] - P.edge_off[inst]) * rpe + (k - P.rng_off[inst]) * d;
    for (int r = 0; r < d; ++r) {
      w[row + r] = P.rng_w[k];
      b[row + r] = 0.0;
    }
  } else if (t < (long)P.E + P.K + P.Lp) {
    const int q = (int)(t - P.E - P.K), inst = find_inst(P.prior_off, P.n_inst, q);
    const int row = P.roff[inst] + (P.edge_off[inst + 1] - P.edge_off[inst]) * rpe +
                    (P.rng_off[inst + 1] - P.rng_off[inst]) * d + (q - P.prior_off[inst]) * d;
    for (int r = 0; r < d; ++r) {
      w[row + r] = P.prior_w[q];
      b[row + r] = P.prior_t[(size_t)q * d + r];
    }
  }
}

// range weight incident on every pose translation / inverse Hessian diagonal of every landmark coordinate from the
// incidence lists (the matrix-free twin of k_diag_setup; k_lm_finish inverts lm_inv afterwards)
__global__ void k_diag_setup_mf(DevProblem P, double *wsum) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= P.P + P.L) return;
  double acc = 0.0;
  for (int j = P.inc_ptr[o]; j < P.inc_ptr[o + 1]; ++j) {
    const IncRec rec = P.inc_rec[j];
    const int kind = (int)(rec.x >> kIncShift), id = (int)(rec.x & kIncMask);
    if (kind == INC_RA || kind == INC_RB)
      acc += P.rng_w[id];
    else if (kind == INC_PR)
      acc += P.prior_w[id];
  }
  if (o < P.P) {
    wsum[o] = acc;
  } else {
    for (int r = 0; r < P.d; ++r) P.lm_inv[(size_t)(o - P.P) * P.d + r] = acc;
  }
}

// h = B^T u per pose / landmark, then the column-space update of the tick.  mode TM_LS: instances in PH_LS take the
// step (z += step dz, dz = 0, r = -h).  mode TM_EVAL: partial |g|^2, g.z, |z|^2 of the instances being certified.
template <int D>
__device__ __forceinline__ void grad_mf_body(DevProblem P, SolverVecs V, BlockTables T, const InstState *st, const int bid,
                                             const int mode) {
  constexpr int D1 = D + 1, BLK = D * D1, RPE = D + D * D;
  __shared__ double red[16 * (kThreads / 32)];
  const HvBlock bd = T.pb[bid];
  const int inst = bd.inst;
  const bool eval = mode == TM_EVAL;
  const int phase = st[inst].phase;
  if (phase == PH_DONE) return;
  if (eval ? !st[inst].eval_now : (phase != PH_LS || st[inst].eval_now)) return;
  const int z0 = bd.z0, Pi = bd.Pi;
  const int r0 = P.roff[inst], e0 = P.edge_off[inst], Ei = P.edge_off[inst + 1] - e0;
  const int k0 = P.rng_off[inst], Ki = P.rng_off[inst + 1] - k0, q0 = P.prior_off[inst];
  const int rr0 = r0 + Ei * RPE, rp0 = rr0 + Ki * D;
  const double *u = V.u;
  const IncRec *__restrict__ recs = reinterpret_cast<const IncRec *>(P.inc_rec);
  const double step = st[inst].step;
  const int tid = threadIdx.x;
  double gg = 0.0, gz = 0.0, zz = 0.0;
  auto finish = [&](int col, double h) {  // one column: certificate sums, or the step to the new point
    if (col < P.blk) h = 0.0;  // pinned pose
    double *zc = V.z + z0 + col;
    if (eval) {
      const double zv = *zc;
      gg += h * h;
      gz += h * zv;
      zz += zv * zv;
    } else {
      double *dzc = V.dz + z0 + col;
      *zc = fma(step, *dzc, *zc);
      *dzc = 0.0;
      V.r[z0 + col] = -h;
    }
  };
  if (bd.kind == CB_POSE) {
    const int p = bd.i0 + tid;
    if (p < bd.i1) {
      const int pg = bd.pg0 + p;
      double h[BLK];
#pragma unroll
      for (int i = 0; i < BLK; ++i) h[i] = 0.0;
      // rows of a relative-pose factor: d translation rows, then d x d rotation rows
      auto edge_rows = [&](int e, bool at_j) {
        const double *ue = u + r0 + (size_t)(e - e0) * RPE;
        if (at_j) {
#pragma unroll
          for (int r = 0; r < D; ++r) {
            h[r * D1 + D] += ue[r];
#pragma unroll
            for (int c = 0; c < D; ++c) h[r * D1 + c] += ue[D + r * D + c];
          }
        } else {
          double tm[D], Rm[D * D], k2, tau2;
          load_edge<D>(P, e, tm, Rm, k2, tau2);
#pragma unroll
          for (int r = 0; r < D; ++r) {
            h[r * D1 + D] -= ue[r];
#pragma unroll
            for (int m = 0; m < D; ++m) {
              double acc = ue[r] * tm[m];
#pragma unroll
              for (int c = 0; c < D; ++c) acc += ue[D + r * D + c] * Rm[m * D + c];
              h[r * D1 + m] -= acc;
            }
          }
        }
      };
      const int e_in = P.link_edge[pg], e_out = (p + 1 < Pi) ? P.link_edge[pg + 1] : -1;
      if (e_in >= 0) edge_rows(e_in, true);
      if (e_out >= 0) edge_rows(e_out, false);
      for (int j = P.inc_ptr[pg]; j < P.inc_ptr[pg + 1]; ++j) {
        const IncRec rec = recs[j];
        const int kind = (int)(rec.x >> kIncShift), id = (int)(rec.x & kIncMask);
        if (kind == INC_RA || kind == INC_RB) {
          const double *uk = u + rr0 + (size_t)(id - k0) * D;
#pragma unroll
          for (int r = 0; r < D; ++r) h[r * D1 + D] += (kind == INC_RA) ? uk[r] : -uk[r];
        } else if (kind == INC_EJ) {
          edge_rows(id, true);
        } else if (kind == INC_EI) {
          edge_rows(id, false);
        }
      }
#pragma unroll
      for (int i = 0; i < BLK; ++i) finish(p * BLK + i, h[i]);
    }
  } else {
    for (int q = bd.i0; q < bd.i1; ++q) {
      double v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = 0.0;
      const int j0 = P.inc_ptr[P.P + bd.lg0 + q], j1 = P.inc_ptr[P.P + bd.lg0 + q + 1];
      for (int j = j0 + tid; j < j1; j += kThreads) {
        const IncRec rec = recs[j];
        const int kind = (int)(rec.x >> kIncShift), id = (int)(rec.x & kIncMask);
        const double *uk = (kind == INC_PR) ? u + rp0 + (size_t)(id - q0) * D : u + rr0 + (size_t)(id - k0) * D;
#pragma unroll
        for (int r = 0; r < D; ++r) v[r] += (kind == INC_RB) ? -uk[r] : uk[r];
      }
      const double tot = block_sum16<kThreads>(v, red);  // thread i < 16 holds the total of v[i]
      if (tid < D) finish(Pi * BLK + q * D + tid, tot);
    }
  }
  if (eval) {
    const double a = block_sum<kThreads>(gg, red);
    const double b = block_sum<kThreads>(gz, red);
    const double c = block_sum<kThreads>(zz, red);
    if (tid == 0) {
      V.part_gr[(size_t)bid * 4 + 0] = a;
      V.part_gr[(size_t)bid * 4 + 1] = b;
      V.part_gr[(size_t)bid * 4 + 2] = c;
    }
  }
}

template <int D>
__global__ void __launch_bounds__(kThreads) k_grad_mf(DevProblem P, SolverVecs V, BlockTables T, const InstState *st, int mode,
                                                     WorkLists W) {
  const int *act;
  int n_act;
  wl_get(W, mode == TM_EVAL ? WL_EVAL : WL_RUN, act, n_act);
  for (long long item = blockIdx.x; item < (long long)n_act * W.maxpb; item += gridDim.x) {
    const int inst = act[item / W.maxpb], bid = T.pb_begin[inst] + (int)(item % W.maxpb);
    if (bid < T.pb_begin[inst + 1]) grad_mf_body<D>(P, V, T, st, bid, mode);
    __syncthreads();
  }
}

#ifndef SCORE_HV_MINB
#define SCORE_HV_MINB 4
#endif
template <int D>
__global__ void __launch_bounds__(kThreads, SCORE_HV_MINB) k_hessvec(DevProblem P, SolverVecs V, BlockTables T, const InstState *st, WorkLists W) {
  const int *act;
  int n_act;
  wl_get(W, WL_RUN, act, n_act);
  for (long long item = blockIdx.x; item < (long long)n_act * W.maxpb; item += gridDim.x) {
    const int inst = act[item / W.maxpb];
    const HvBlock bd = T.pbd[(size_t)inst * W.maxpb + (int)(item % W.maxpb)];
    if (bd.kind >= 0) hessvec_body<D>(P, V, st, bd);
    __syncthreads();
  }
}

// jb / je of every pose block (compact and dense tables) once the incidence lists exist
__global__ void k_hv_fill(DevProblem P, BlockTables T, int n_dense) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T.n_pb + n_dense) return;
  HvBlock *b = i < T.n_pb ? T.pb + i : T.pbd + (i - T.n_pb);
  if (b->kind != CB_POSE) return;
  b->jb = P.inc_ptr[b->pg0 + b->i0];
  b->je = P.inc_ptr[b->pg0 + b->i1];
}

}  // namespace score
