// Odometry-chain preconditioner.
//
// Along a chain segment the odometry residuals C_p = X_{p+1} - X_p T~_p  (X_p = [R_p|t_p], T~_p the
// measured SE(d) step; score/utils/gurobi_utils.py:504-526 written in matrix form) are an invertible
// linear change of variables.  With G_p the dead-reckoned frame (G_{p+1} = G_p T~_p) one has
//     X_p = ( sum_{q<=p} U_q G_q^{-1} ) G_p ,   U_first = X_first, U_q = C_{q-1},
// so the odometry Hessian is diagonal in U and its inverse is two prefix sums wrapped in per-pose
// (d+1)x(d+1) frame changes:  P = G . cumsum . M . cumsum^T . G^T  with  M_q = G_q^{-T} D_q^{-1} G_q^{-1},
// D = diag(2 tau, ..., 2 tau, 2 k).  Segment bases (no odometry curvature) use the inverse of the
// range Hessian block  sum_p 2 w_p h_p h_p^T,  h_p = (tg_p, 1);  the pinned pose gets M = 0.
#pragma once
#include "common.cuh"

// CTAs of kSegThreads threads per SM the chain-scan kernels are compiled for.  d = 2: 16 = every thread slot of the SM
// (32 registers per thread, a few spilled words).  These kernels are bound by the latency of their dependent rounds of
// loads, and 16 resident segments per SM measured 614 us per PCG tick of the 1024-instance sweep against 649 us with
// the compiler's own choice (40 / 56 registers, 12 / 9 CTAs per SM) — profiles/occupancy_r2.txt.  d = 3 (12-entry
// pose blocks): 8 CTAs, 64 registers.
#ifndef SCORE_PC_MINB
#define SCORE_PC_MINB 16
#endif
#ifndef SCORE_PC_MINB3
#define SCORE_PC_MINB3 8
#endif

namespace score {

// One thread per segment: G_p = [Rg|tg], sequential composition (setup only).
__global__ void k_dead_reckon(DevProblem P) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= P.n_seg) return;
  const int d = P.d, blk = P.blk;
  const int p0 = P.seg_ptr[s], p1 = P.seg_ptr[s + 1];
  double Rg[9], tg[3], Rn[9], tn[3];
  for (int i = 0; i < d; ++i) {
    tg[i] = 0.0;
    for (int j = 0; j < d; ++j) Rg[i * d + j] = (i == j) ? 1.0 : 0.0;
  }
  for (int p = p0; p < p1; ++p) {
    if (p > p0) {
      const int e = P.link_edge[p];
      const double *Rm = P.edge_R + (size_t)e * d * d;
      const double *tm = P.edge_t + (size_t)e * d;
      for (int i = 0; i < d; ++i) {
        double acc = tg[i];
        for (int m = 0; m < d; ++m) acc += Rg[i * d + m] * tm[m];
        tn[i] = acc;
        for (int j = 0; j < d; ++j) {
          double a = 0.0;
          for (int m = 0; m < d; ++m) a += Rg[i * d + m] * Rm[m * d + j];
          Rn[i * d + j] = a;
        }
      }
      for (int i = 0; i < d; ++i) {
        tg[i] = tn[i];
        for (int j = 0; j < d; ++j) Rg[i * d + j] = Rn[i * d + j];
      }
    }
    double *Gp = P.G + (size_t)p * blk;
    for (int i = 0; i < d; ++i) {
      for (int j = 0; j < d; ++j) Gp[i * (d + 1) + j] = Rg[i * d + j];
      Gp[i * (d + 1) + d] = tg[i];
    }
  }
}

// Range weight incident on each pose translation (deterministic: walks the t-column of B^T in row
// order) and the inverse Hessian diagonal of every landmark column.
__global__ void k_diag_setup(DevProblem P, double *wsum) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int d = P.d, blk = P.blk;
  if (i < P.P) {
    const int inst = find_inst(P.pose_off, P.n_inst, i);
    const int col = P.zoff[inst] + (i - P.pose_off[inst]) * blk + d;  // t[0] column
    const int Ei = P.edge_off[inst + 1] - P.edge_off[inst];
    const int Ki = P.rng_off[inst + 1] - P.rng_off[inst];
    const int rr0 = P.roff[inst] + Ei * P.rpe, rr1 = rr0 + Ki * d;
    double acc = 0.0;
    for (int k = P.t_indptr[col]; k < P.t_indptr[col + 1]; ++k) {
      const int row = P.t_rows[k];
      if (row >= rr0 && row < rr1) acc += P.w[row] * P.t_vals[k] * P.t_vals[k];
    }
    wsum[i] = acc;
  } else if (i < P.P + P.L * d) {
    const int j = i - P.P;  // global landmark coordinate index
    const int lq = j / d, r = j % d;
    const int inst = find_inst(P.lm_off, P.n_inst, lq);
    const int Pi = P.pose_off[inst + 1] - P.pose_off[inst];
    const int col = P.zoff[inst] + Pi * blk + (lq - P.lm_off[inst]) * d + r;
    double acc = 0.0;
    for (int k = P.t_indptr[col]; k < P.t_indptr[col + 1]; ++k) acc += P.w[P.t_rows[k]] * P.t_vals[k] * P.t_vals[k];
    P.lm_inv[j] = acc;  // raw sum; k_lm_finish inverts it (after the sum over ranks in a row-partitioned solve)
  }
}

__global__ void k_lm_finish(DevProblem P) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= P.L * P.d) return;
  const double acc = P.lm_inv[j];
  P.lm_inv[j] = (acc > 0.0) ? 1.0 / (2.0 * acc) : 1.0;
}

// In-place inverse of a small SPD matrix (n <= 4) by Gauss-Jordan.
__device__ inline void spd_inverse(double *A, int n) {
  double inv[16];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) inv[i * n + j] = (i == j) ? 1.0 : 0.0;
  for (int c = 0; c < n; ++c) {
    const double piv = 1.0 / A[c * n + c];
    for (int j = 0; j < n; ++j) {
      A[c * n + j] *= piv;
      inv[c * n + j] *= piv;
    }
    for (int i = 0; i < n; ++i) {
      if (i == c) continue;
      const double f = A[i * n + c];
      for (int j = 0; j < n; ++j) {
        A[i * n + j] -= f * A[c * n + j];
        inv[i * n + j] -= f * inv[c * n + j];
      }
    }
  }
  for (int i = 0; i < n * n; ++i) A[i] = inv[i];
}

// One thread per pose: M_p for chain-interior poses; one extra sequential loop for segment bases.
__global__ void k_build_M(DevProblem P, const double *wsum) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P.P) return;
  const int d = P.d, d1 = d + 1, blk = P.blk;
  double Mp[16];                                          // full block, stored below as its upper triangle
  double *Mout = P.M + (size_t)p * (d1 * (d1 + 1) / 2);  // M_p is symmetric: (d+1)(d+2)/2 doubles, row-major upper
  auto store = [&]() {
    int k = 0;
    for (int i = 0; i < d1; ++i)
      for (int j = i; j < d1; ++j) Mout[k++] = Mp[i * d1 + j];
  };
  const int e = P.link_edge[p];
  if (e >= 0) {
    // A = G^{-1} = [[Rg^T, -Rg^T tg],[0,1]] ;  M = A^T D^{-1} A
    const double *Gp = P.G + (size_t)p * blk;
    double A[16];
    for (int i = 0; i < d; ++i) {
      double acc = 0.0;
      for (int j = 0; j < d; ++j) {
        A[i * d1 + j] = Gp[j * d1 + i];
        acc -= Gp[j * d1 + i] * Gp[j * d1 + d];
      }
      A[i * d1 + d] = acc;
    }
    for (int j = 0; j < d; ++j) A[d * d1 + j] = 0.0;
    A[d * d1 + d] = 1.0;
    const double ir = 1.0 / (2.0 * P.edge_tau[e]), it = 1.0 / (2.0 * P.edge_k[e]);
    for (int i = 0; i < d1; ++i)
      for (int j = 0; j < d1; ++j) {
        double acc = 0.0;
        for (int m = 0; m < d1; ++m) acc += A[m * d1 + i] * ((m < d) ? ir : it) * A[m * d1 + j];
        Mp[i * d1 + j] = acc;
      }
    store();
    return;
  }
  // segment base
  const int inst = find_inst(P.pose_off, P.n_inst, p);
  if (p == P.pose_off[inst]) {  // pinned pose: pin_pose, gurobi_utils.py:316-333
    for (int i = 0; i < d1 * d1; ++i) Mp[i] = 0.0;
    store();
    return;
  }
  // find the segment end: next pose with link_edge < 0 or instance end
  int q = p + 1;
  const int pend = P.pose_off[inst + 1];
  while (q < pend && P.link_edge[q] >= 0) ++q;
  double H[16];
  for (int i = 0; i < d1 * d1; ++i) H[i] = 0.0;
  for (int pp = p; pp < q; ++pp) {
    const double ww = 2.0 * wsum[pp];
    if (ww == 0.0) continue;
    const double *Gp = P.G + (size_t)pp * blk;
    double h[4];
    for (int i = 0; i < d; ++i) h[i] = Gp[i * d1 + d];
    h[d] = 1.0;
    for (int i = 0; i < d1; ++i)
      for (int j = 0; j < d1; ++j) H[i * d1 + j] += ww * h[i] * h[j];
  }
  double tr = 0.0;
  for (int i = 0; i < d1; ++i) tr += H[i * d1 + i];
  if (!(tr > 0.0)) {
    double sc = 1.0;
    if (q > p + 1) sc = 1.0 / (2.0 * P.edge_k[P.link_edge[p + 1]]);
    for (int i = 0; i < d1; ++i)
      for (int j = 0; j < d1; ++j) Mp[i * d1 + j] = (i == j) ? sc : 0.0;
    store();
    return;
  }
  for (int i = 0; i < d1; ++i) H[i * d1 + i] += 1e-9 * tr;
  spd_inverse(H, d1);
  for (int i = 0; i < d1 * d1; ++i) Mp[i] = H[i];
  store();
}

// z0: every segment dead-reckoned from an identity base; landmarks at the origin.
__global__ void k_init_z(DevProblem P, double *z) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.nz) return;
  const int inst = find_inst(P.zoff, P.n_inst, i);
  const int loc = i - P.zoff[inst];
  const int Pi = P.pose_off[inst + 1] - P.pose_off[inst];
  if (loc < Pi * P.blk)
    z[i] = P.G[(size_t)P.pose_off[inst] * P.blk + loc];
  else
    z[i] = 0.0;
}

// N doubles of one pose (N even) from / to a 16-byte aligned address: N/2 128-bit accesses.  A thread-per-pose access
// with a 48- or 96-byte stride touches the same lines either way; half the instructions are half the first-level-cache
// wavefronts, which is what bounds these kernels (profiles/occupancy_r2.txt).  `al` (CTA-uniform): the address is aligned.
template <int N>
__device__ __forceinline__ void ld_block(const double *src, double (&o)[N], const bool al) {
  if (al && N % 2 == 0) {
    const double2 *s2 = reinterpret_cast<const double2 *>(src);
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
      const double2 t = s2[i];
      o[2 * i] = t.x;
      o[2 * i + 1] = t.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) o[i] = src[i];
  }
}
template <int N>
__device__ __forceinline__ void st_block(double *dst, const double (&v)[N], const bool al) {
  if (al && N % 2 == 0) {
    double2 *d2 = reinterpret_cast<double2 *>(dst);
#pragma unroll
    for (int i = 0; i < N / 2; ++i) d2[i] = make_double2(v[2 * i], v[2 * i + 1]);
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) dst[i] = v[i];
  }
}

// Barrier of the kSegThreads threads that run a chain-scan body.  SUB = false: they are the whole CTA.  SUB = true:
// they are the first kSegThreads threads of a larger CTA (the fused PCG kernel, fused.cuh) and meet at named barrier 1.
template <bool SUB>
__device__ __forceinline__ void seg_bar() {
  if (SUB)
    asm volatile("bar.sync 1, %0;" ::"n"(kSegThreads) : "memory");
  else
    __syncthreads();
}

// block_sum over the kSegThreads scan threads (same tree as block_sum<kSegThreads>)
template <bool SUB>
__device__ __forceinline__ double seg_sum(double v, double *smem /* >= kSegThreads/32 */) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  seg_bar<SUB>();
  if (lane == 0) smem[wid] = v;
  seg_bar<SUB>();
  double out = 0.0;
  if (wid == 0) {
    out = (lane < kSegThreads / 32) ? smem[lane] : 0.0;
    out = warp_sum(out);
  }
  return out;
}

// Inclusive scan of NV doubles per thread across the CTA (plus running carry across tiles).
template <int NV, bool SUB>
__device__ __forceinline__ void cta_scan(double (&v)[NV], double (*wtot)[NV], double *carry) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  constexpr int NW = kSegThreads / 32;
#pragma unroll
  for (int c = 0; c < NV; ++c) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double n = __shfl_up_sync(0xffffffffu, v[c], o);
      if (lane >= o) v[c] += n;
    }
  }
  if (lane == 31) {
#pragma unroll
    for (int c = 0; c < NV; ++c) wtot[wid][c] = v[c];
  }
  seg_bar<SUB>();
#pragma unroll
  for (int c = 0; c < NV; ++c) {
    double add = carry[c];
    for (int w = 0; w < wid; ++w) add += wtot[w][c];
    v[c] += add;
  }
  double ncarry = 0.0;
  if (threadIdx.x < NV) {
    ncarry = carry[threadIdx.x];
    for (int w = 0; w < NW; ++w) ncarry += wtot[w][threadIdx.x];
  }
  seg_bar<SUB>();
  if (threadIdx.x < NV) carry[threadIdx.x] = ncarry;
}

// ---- s = P r, pass 1 (reverse): S_p = sum_{q>=p} r_q G_q^T ;  Y_p = S_p M_p -> ytmp.
// With the coarse level on, the base block S_first goes to the coarse right-hand side instead.
// CTA per chain segment; CTAs past n_seg handle the landmark block of one instance each.
// `s` = segment (s < n_seg; [p0, p1) its global pose range) or n_seg + inst (the landmark block of instance `inst`).
// The caller passes what it already knows (inst; the segment's bounds from the dense item table) so that the body's
// own loads — instance state, offsets — go out in ONE round before the data loads instead of four dependent ones.
template <int D, bool SUB = false>
__device__ __forceinline__ void precond_rev_body(DevProblem P, SolverVecs V, const InstState *st, const int s, const int inst,
                                                 const int p0, const int p1) {
  constexpr int D1 = D + 1, NV = D * D1;
  __shared__ double wtot[kSegThreads / 32][NV];
  __shared__ double carry[NV];
  __shared__ double red[kSegThreads / 32];
  const int tid = threadIdx.x;
  // one round: every address below is valid whatever the state says
  const int phase = st[inst].phase, evn = st[inst].eval_now;
  const int zo = P.zoff[inst], po = P.pose_off[inst], cn = P.c_n[inst], coff = P.c_off[inst];
  if (s >= P.n_seg) {
    const int po1 = P.pose_off[inst + 1], l0 = P.lm_off[inst], l1 = P.lm_off[inst + 1], cnb = P.c_nb[inst];
    if (phase == PH_DONE || phase == PH_WAIT || evn) return;
    const int Pi = po1 - po;
    const int n = (l1 - l0) * D;
    const int c0 = zo + Pi * P.blk, j0 = l0 * D;
    if (cn > 0) {
      double *c = P.c_rhs + coff + cnb;
      for (int j = tid; j < n; j += kSegThreads) c[j] = V.r[c0 + j];
      return;
    }
    double acc = 0.0;
    for (int j = tid; j < n; j += kSegThreads) {
      const double rv = V.r[c0 + j];
      const double sv = rv * P.lm_inv[j0 + j];
      V.s[c0 + j] = sv;
      acc += rv * sv;
    }
    const double tot = seg_sum<SUB>(acc, red);
    if (tid == 0) V.part_lm[inst] = tot;
    return;
  }
  const int sb = P.seg_begin[inst];
  if (phase == PH_DONE || phase == PH_WAIT || evn) return;
  const int len = p1 - p0;
  const long colbase = (long)zo - (long)po * NV;
  const bool al = (colbase & 1) == 0;  // pose blocks of the column-space vectors start on 16-byte boundaries
  const int slot = s - sb - 1;  // -1: pinned segment
  const bool to_coarse = cn > 0 && slot >= 0;

  if (tid < NV) carry[tid] = 0.0;
  seg_bar<SUB>();
  for (int t0 = 0; t0 < len; t0 += kSegThreads) {
    const int idx = t0 + tid;
    const bool valid = idx < len;
    const int pg = p1 - 1 - idx;
    double v[NV];
#pragma unroll
    for (int c = 0; c < NV; ++c) v[c] = 0.0;
    if (valid) {
      double g[NV], rv[NV];
      ld_block<NV>(P.G + (size_t)pg * NV, g, true);
      ld_block<NV>(V.r + colbase + (long)pg * NV, rv, al);
#pragma unroll
      for (int r = 0; r < D; ++r) {
#pragma unroll
        for (int c = 0; c < D; ++c) {
          double acc = rv[r * D1 + D] * g[c * D1 + D];
#pragma unroll
          for (int m = 0; m < D; ++m) acc += rv[r * D1 + m] * g[c * D1 + m];
          v[r * D1 + c] = acc;
        }
        v[r * D1 + D] = rv[r * D1 + D];
      }
    }
    cta_scan<NV, SUB>(v, wtot, carry);
    if (valid) {
      if (pg == p0 && to_coarse) {
        double *c = P.c_rhs + coff + slot * NV;
#pragma unroll
        for (int i = 0; i < NV; ++i) c[i] = v[i];
      } else {
        constexpr int NM = D1 * (D1 + 1) / 2;  // upper triangle of the symmetric M_p
        double mm[NM], y[NV];
        ld_block<NM>(P.M + (size_t)pg * NM, mm, true);
#pragma unroll
        for (int r = 0; r < D; ++r)
#pragma unroll
          for (int c = 0; c < D1; ++c) {
            double acc = 0.0;
#pragma unroll
            for (int m = 0; m < D1; ++m) {
              const int lo = m < c ? m : c, hi = m < c ? c : m;
              acc += v[r * D1 + m] * mm[lo * D1 - lo * (lo - 1) / 2 + (hi - lo)];
            }
            y[r * D1 + c] = acc;
          }
        st_block<NV>(V.ytmp + colbase + (long)pg * NV, y, al);
      }
    }
  }
}

// Listed instance x (its chain segments, then its landmark block: slot W.maxseg).
template <int D>
__global__ void __launch_bounds__(kSegThreads, D == 2 ? SCORE_PC_MINB : SCORE_PC_MINB3) k_precond_rev(DevProblem P, SolverVecs V, const InstState *st, WorkLists W) {
  const int *act;
  int n_act;
  wl_get(W, WL_RUN, act, n_act);
  const int per = W.maxseg + 1;
  for (long long item = blockIdx.x; item < (long long)n_act * per; item += gridDim.x) {
    const int inst = act[item / per], j = (int)(item % per);
    if (j == W.maxseg) {
      precond_rev_body<D>(P, V, st, P.n_seg + inst, inst, 0, 0);
    } else {
      const int4 sd = P.seg_tab[(size_t)inst * W.maxseg + j];  // {first pose, end pose, segment} or first pose < 0
      if (sd.x >= 0) precond_rev_body<D>(P, V, st, sd.z, inst, sd.x, sd.y);
    }
    __syncthreads();  // the scan's shared carry is reused by the next item
  }
}

// ---- s = P r, pass 2 (forward): Xh_p = sum_{q<=p} Y_q ;  s_p = Xh_p G_p ;  partial r.s
// With `fuse` the coarse application y = A_c^-1 c (coarse.cuh) happens here instead of in a kernel of its own:
// the CTA of free segment sl needs exactly the blk rows of y that start its prefix sum, the CTA of the pinned
// first segment takes the landmark rows (-> s, partial r.s).  Every row is one warp's lane-strided dot product,
// the same summation order as k_coarse_apply, so the two variants agree to the bit.
template <int D, bool SUB = false>
__device__ __forceinline__ void precond_fwd_body(DevProblem P, SolverVecs V, const InstState *st, const int s, const int inst,
                                                 const int p0, const int p1, const bool fuse) {
  constexpr int D1 = D + 1, NV = D * D1, NWS = kSegThreads / 32;
  __shared__ double wtot[kSegThreads / 32][NV];
  __shared__ double carry[NV];
  __shared__ double red[kSegThreads / 32];
  __shared__ double ybase[NV];
  __shared__ double ylm[kCoarseMax];
  const int tid = threadIdx.x;
  // one round of loads (every address is valid whatever the state says)
  const int phase = st[inst].phase, evn = st[inst].eval_now;
  const int zo = P.zoff[inst], po = P.pose_off[inst], po1 = P.pose_off[inst + 1], sb = P.seg_begin[inst];
  const int cn = P.c_n[inst], coff = P.c_off[inst], cmoff = P.c_moff[inst], cnb = P.c_nb[inst];
  if (phase == PH_DONE || phase == PH_WAIT || evn) return;
  const int len = p1 - p0;
  const long colbase = (long)zo - (long)po * NV;
  const bool al = (colbase & 1) == 0;
  const int sl = s - sb;
  const int nco = (fuse && cn <= kCoarseMax) ? cn : 0;  // larger coarse spaces: kernels of their own
  if (nco > 0) {
    const double *__restrict__ Ai = P.c_Ainv + cmoff;  // constant during PCG ticks (rebuilt in line-search ticks)
    const double *cv = P.c_rhs + coff;                  // written by the reverse pass: coherent loads
    const int nb = cnb, lane = tid & 31, wid = tid >> 5;
    const int row0 = (sl >= 1) ? (sl - 1) * NV : nb, nrow = (sl >= 1) ? NV : nco - nb;
    // three rows per warp step (rows wid, wid + 4, wid + 8): their loads are all in flight before the first reduction;
    // every row is still summed lane-strided in j, then by the warp tree
    constexpr int RW = 3;
    for (int rr0 = wid; rr0 < nrow; rr0 += RW * NWS) {
      double a[RW];
#pragma unroll
      for (int t = 0; t < RW; ++t) a[t] = 0.0;
      for (int j = lane; j < nco; j += 32) {
        const double cj = cv[j];
#pragma unroll
        for (int t = 0; t < RW; ++t)
          if (rr0 + t * NWS < nrow) a[t] += __ldg(Ai + (size_t)(row0 + rr0 + t * NWS) * nco + j) * cj;
      }
#pragma unroll
      for (int t = 0; t < RW; ++t) {
        const double tot = warp_sum(a[t]);
        if (lane == 0 && rr0 + t * NWS < nrow) (sl >= 1 ? ybase : ylm)[rr0 + t * NWS] = tot;
      }
    }
  }
  if (tid < NV) carry[tid] = 0.0;
  seg_bar<SUB>();
  if (nco > 0 && sl == 0) {  // landmark block: s = y, partial r.s
    const int Pi = po1 - po, c0 = zo + Pi * NV, nlm = nco - cnb;
    double acc = 0.0;
    for (int j = tid; j < nlm; j += kSegThreads) {
      const double sv = ylm[j];
      V.s[c0 + j] = sv;
      acc += sv * V.r[c0 + j];
    }
    const double tot = seg_sum<SUB>(acc, red);
    if (tid == 0) V.part_lm[inst] = tot;
  }
  const bool ystart = nco > 0 && sl >= 1;  // the segment's base block starts from the coarse solution
  double dot = 0.0;
  for (int t0 = 0; t0 < len; t0 += kSegThreads) {
    const int idx = t0 + tid;
    const bool valid = idx < len;
    const int pg = p0 + idx;
    double v[NV];
#pragma unroll
    for (int c = 0; c < NV; ++c) v[c] = 0.0;
    if (valid) {
      double *yp = V.ytmp + colbase + (long)pg * NV;
      if (ystart && idx == 0) {
#pragma unroll
        for (int c = 0; c < NV; ++c) yp[c] = v[c] = ybase[c];
      } else {
        ld_block<NV>(yp, v, al);
      }
    }
    cta_scan<NV, SUB>(v, wtot, carry);
    if (valid) {
      double g[NV], rv[NV], sv[NV];
      ld_block<NV>(P.G + (size_t)pg * NV, g, true);
      ld_block<NV>(V.r + colbase + (long)pg * NV, rv, al);
#pragma unroll
      for (int r = 0; r < D; ++r) {
        double tacc = v[r * D1 + D];
#pragma unroll
        for (int c = 0; c < D; ++c) {
          double acc = 0.0;
#pragma unroll
          for (int m = 0; m < D; ++m) acc += v[r * D1 + m] * g[m * D1 + c];
          sv[r * D1 + c] = acc;
          dot += acc * rv[r * D1 + c];
          tacc += v[r * D1 + c] * g[c * D1 + D];
        }
        sv[r * D1 + D] = tacc;
        dot += tacc * rv[r * D1 + D];
      }
      st_block<NV>(V.s + colbase + (long)pg * NV, sv, al);
    }
  }
  const double tot = seg_sum<SUB>(dot, red);
  if (tid == 0) V.part_seg[s] = tot;
}

template <int D>
__global__ void __launch_bounds__(kSegThreads, D == 2 ? SCORE_PC_MINB : SCORE_PC_MINB3) k_precond_fwd(DevProblem P, SolverVecs V, const InstState *st, WorkLists W,
                                                            const bool fuse) {
  const int *act;
  int n_act;
  wl_get(W, WL_RUN, act, n_act);
  for (long long item = blockIdx.x; item < (long long)n_act * W.maxseg; item += gridDim.x) {
    const int inst = act[item / W.maxseg];
    const int4 sd = P.seg_tab[(size_t)inst * W.maxseg + (int)(item % W.maxseg)];
    if (sd.x >= 0) precond_fwd_body<D>(P, V, st, sd.z, inst, sd.x, sd.y, fuse);
    __syncthreads();
  }
}

}  // namespace score
