// Coarse level of the preconditioner.
//
// The odometry Hessian is blind to two families of directions: rigid re-placement of a whole chain
// segment (its base block U_first) and the landmarks.  Their curvature comes only from the range terms
// and couples all segments of an instance.  Per instance these nc = (n_seg - 1) blk + L d coordinates
// form a small dense block A_c = Z^T H_range Z (Z: tree coordinates -> x), rebuilt at every Newton step
// from the current per-range curvature blocks M_k, inverted on chip, and applied inside the
// preconditioner between the two prefix sums:  P = T blockdiag(D^-1, A_c^-1) T^T.
//
// A "slot" is one coarse block: free segment s (blk coordinates, the base frame of the segment) or
// landmark q (d coordinates).  With h_p = (tg_p, 1) the dead-reckoned position of pose p in its segment
// frame (landmarks: h = (0, 1)), J = I_d (x) h^T maps slot coordinates to the translation of an endpoint
// and a range k between slots (a, b) contributes
//     block(a, a) += M_k (x) h_a h_a^T        block(a, b) -= M_k (x) h_a h_b^T   (and transposed).
// The contributions are summed from two static, host-sorted lists (api.cu: build_coarse_tables):
//   * incidences sorted by slot      -> diagonal blocks, one warp per slot run;
//   * ranges sorted by slot pair     -> upper off-diagonal blocks, contiguous runs per warp;
// so every block has exactly one owning warp, the summation order is fixed and no atomics are needed.
// The nc x nc matrix is then inverted by blocked symmetric sweeps (Gauss-Jordan in its symmetry-preserving form)
// with the matrix held in registers (TS x TS tile per thread; per block of TS pivots one raw and one scaled row
// panel broadcast through shared memory, one barrier).
#pragma once
#include "common.cuh"

namespace score {

template <int D>
struct CoarseDims {
  static constexpr int D1 = D + 1;
  static constexpr int BLK = D * D1;
  static constexpr int NM = D * (D + 1) / 2;      // unique entries of the symmetric d x d curvature block
  static constexpr int NH = D1 * (D1 + 1) / 2;    // unique entries of h h^T
  static constexpr int NPD = NM * NH;             // products per incidence (diagonal blocks)
  static constexpr int NPO = NM * D1 * D1;        // products per range (off-diagonal blocks)
  static constexpr int SB = (D == 2) ? 32 : 16;   // ranges staged per warp batch
  static constexpr int REC_D = NM + D1;           // staged doubles per incidence
  static constexpr int REC_O = NM + 2 * D1;       // staged doubles per pair entry
  static constexpr int STAGE = SB * REC_O;        // per-warp staging doubles
};

// (a, b) of the m-th unique entry of a symmetric D x D matrix stored row-major upper.
template <int D>
__host__ __device__ __forceinline__ void sym_pair(int m, int &a, int &b) {
  a = 0;
  int rowlen = D;
  while (m >= rowlen) {
    m -= rowlen;
    ++a;
    --rowlen;
  }
  b = a + m;
}
template <int D>
__host__ __device__ __forceinline__ int sym_index(int a, int b) {
  if (a > b) {
    const int t = a;
    a = b;
    b = t;
  }
  return a * D - a * (a - 1) / 2 + (b - a);
}

// Static per-entry frames of the two sorted lists (after the dead reckoning): h of every incidence and of
// both endpoints of every pair entry.
template <int D>
__global__ void k_coarse_static(DevProblem P) {
  constexpr int D1 = D + 1, BLK = D * D1;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  auto frame = [&](int inst, int owner, double *h) {
    const int Pi = P.pose_off[inst + 1] - P.pose_off[inst];
    if (owner < Pi) {
      const double *Gp = P.G + (size_t)(P.pose_off[inst] + owner) * BLK;
#pragma unroll
      for (int c = 0; c < D; ++c) h[c] = Gp[c * D1 + D];
    } else {
#pragma unroll
      for (int c = 0; c < D; ++c) h[c] = 0.0;
    }
    h[D] = 1.0;
  };
  if (j < P.c_ninc) {
    const int inst = find_inst(P.c_inc_off, P.n_inst, j);
    const int code = P.c_inc_code[j], k = P.rng_off[inst] + (code >> 2), e = code & 3;
    double h[D1];
    if (e == 2) {  // both endpoints in this slot: J_a - J_b
      double hb[D1];
      frame(inst, P.rng_a[k], h);
      frame(inst, P.rng_b[k], hb);
#pragma unroll
      for (int c = 0; c < D1; ++c) h[c] -= hb[c];
    } else {
      frame(inst, e == 0 ? P.rng_a[k] : P.rng_b[k], h);
    }
#pragma unroll
    for (int c = 0; c < D1; ++c) P.c_inc_h[(size_t)j * D1 + c] = h[c];
  }
  if (j < P.c_npair) {
    const int inst = find_inst(P.c_pr_off, P.n_inst, j);
    const int code = P.c_pr_code[j], k = P.rng_off[inst] + (code >> 1), flip = code & 1;
    double ha[D1], hb[D1];
    frame(inst, P.rng_a[k], ha);
    frame(inst, P.rng_b[k], hb);
#pragma unroll
    for (int c = 0; c < D1; ++c) {
      P.c_pr_h[(size_t)j * 2 * D1 + c] = flip ? hb[c] : ha[c];        // lower slot
      P.c_pr_h[(size_t)j * 2 * D1 + D1 + c] = flip ? ha[c] : hb[c];   // higher slot
    }
  }
}

// coarse index of coordinate (r, c) of a slot; -1 when the slot has no such coordinate
template <int D>
__device__ __forceinline__ int coarse_index(int slot, int nsegfree, int nb, int r, int c) {
  if (slot < nsegfree) return slot * (D * (D + 1)) + r * (D + 1) + c;
  return (c == D) ? nb + (slot - nsegfree) * D + r : -1;
}

// Sum the range contributions into A (row stride lda; zero-initialised by the caller): diagonal blocks from
// the slot-sorted incidence runs (run r -> warp r mod nw) and upper off-diagonal blocks from the pair-sorted
// runs.  `wid` / `nw`: this warp's index / the number of cooperating warps; `mystage`: CoarseDims::STAGE
// doubles of shared memory private to the warp.  Every block has exactly one owning warp.
template <int D>
__device__ __forceinline__ void coarse_accumulate(const DevProblem &P, const SolverVecs &V, const int inst, const double reg,
                                                  double *A, const int lda, const int wid, const int nw,
                                                  const bool use_owarp, double *mystage) {
  using CD = CoarseDims<D>;
  constexpr int D1 = CD::D1, BLK = CD::BLK, NM = CD::NM, NH = CD::NH, SB = CD::SB;
  const int lane = threadIdx.x & 31;
  const int nb = P.c_nb[inst], nsegfree = nb / BLK;
  const int k0 = P.rng_off[inst];
  // ---- diagonal blocks: runs of the slot-sorted incidence list, run r -> warp r % NW
  {
    constexpr int NACC = (CD::NPD + 31) / 32;
    int pm[NACC], pc[NACC], pcc[NACC];  // per accumulator: curvature entry, (c, c') of h h^T
#pragma unroll
    for (int a = 0; a < NACC; ++a) {
      const int p = lane + 32 * a;
      pm[a] = (p < CD::NPD) ? p / NH : -1;
      int c, cc;
      sym_pair<D1>((p < CD::NPD) ? p % NH : 0, c, cc);
      pc[a] = c;
      pcc[a] = cc;
    }
    const int r0 = P.c_drun_off[inst], r1 = P.c_drun_off[inst + 1];
    for (int run = r0 + wid; run < r1; run += nw) {
      const int slot = P.c_drun_slot[run];
      const int jb = P.c_drun_begin[run], je = P.c_drun_begin[run + 1];
      double acc[NACC];
#pragma unroll
      for (int a = 0; a < NACC; ++a) acc[a] = 0.0;
      for (int j0 = jb; j0 < je; j0 += SB) {
        const int cnt = min(SB, je - j0);
        __syncwarp();
        if (lane < cnt) {
          const int j = j0 + lane;
          const int k = k0 + (P.c_inc_code[j] >> 2);
          const double w2r = reg * P.c_inc_w2[j];
          double *rec = mystage + lane * CD::REC_D;
#pragma unroll
          for (int m = 0; m < NM; ++m) {
            int a, b;
            sym_pair<D>(m, a, b);
            rec[m] = V.mk[(size_t)k * NM + m] + ((a == b) ? w2r : 0.0);
          }
#pragma unroll
          for (int c = 0; c < D1; ++c) rec[NM + c] = P.c_inc_h[(size_t)j * D1 + c];
        }
        __syncwarp();
        for (int q = 0; q < cnt; ++q) {
          const double *rec = mystage + q * CD::REC_D;
#pragma unroll
          for (int a = 0; a < NACC; ++a)
            if (pm[a] >= 0) acc[a] += rec[pm[a]] * (rec[NM + pc[a]] * rec[NM + pcc[a]]);
        }
      }
      // flush: entry ((r, c), (r', c')) = M[r, r'] h[c] h[c'] and its symmetric images
#pragma unroll
      for (int a = 0; a < NACC; ++a) {
        if (pm[a] < 0) continue;
        int r, rr;
        sym_pair<D>(pm[a], r, rr);
        const int c = pc[a], cc = pcc[a];
        const int i0 = coarse_index<D>(slot, nsegfree, nb, r, c), i1 = coarse_index<D>(slot, nsegfree, nb, rr, cc);
        const int i2 = coarse_index<D>(slot, nsegfree, nb, r, cc), i3 = coarse_index<D>(slot, nsegfree, nb, rr, c);
        const double v = acc[a];
        if (i0 >= 0 && i1 >= 0) {
          A[i0 * lda + i1] += v;
          if (i1 != i0) A[i1 * lda + i0] += v;
        }
        if (c != cc && r != rr && i2 >= 0 && i3 >= 0) {  // (r, c') x (r', c): distinct image only when both differ
          A[i2 * lda + i3] += v;
          A[i3 * lda + i2] += v;
        }
      }
    }
  }
  // ---- upper off-diagonal blocks: contiguous runs of the pair-sorted range list per warp
  {
    constexpr int NACC = (CD::NPO + 31) / 32;
    int pm[NACC], pc[NACC], pcc[NACC];
#pragma unroll
    for (int a = 0; a < NACC; ++a) {
      const int p = lane + 32 * a;
      pm[a] = (p < CD::NPO) ? p / (D1 * D1) : -1;
      pc[a] = (p / D1) % D1;
      pcc[a] = p % D1;
    }
    // small matrices: the host-balanced contiguous chunk of this warp; big ones: round robin over all warps
    const int *wsplit = P.c_owarp + (size_t)inst * (kCoarseThreads / 32 + 1);
    const int run_b = use_owarp ? wsplit[wid] : P.c_orun_off[inst] + wid;
    const int run_e = use_owarp ? wsplit[wid + 1] : P.c_orun_off[inst + 1];
    const int run_step = use_owarp ? 1 : nw;
    for (int run = run_b; run < run_e; run += run_step) {
      const int lo = P.c_orun_lo[run], hi = P.c_orun_hi[run];
      const int jb = P.c_orun_begin[run], je = P.c_orun_begin[run + 1];
      double acc[NACC];
#pragma unroll
      for (int a = 0; a < NACC; ++a) acc[a] = 0.0;
      for (int j0 = jb; j0 < je; j0 += SB) {
        const int cnt = min(SB, je - j0);
        __syncwarp();
        if (lane < cnt) {
          const int j = j0 + lane;
          const int k = k0 + (P.c_pr_code[j] >> 1);
          double *rec = mystage + lane * CD::REC_O;
#pragma unroll
          for (int m = 0; m < NM; ++m) rec[m] = V.mk[(size_t)k * NM + m];
#pragma unroll
          for (int c = 0; c < 2 * D1; ++c) rec[NM + c] = P.c_pr_h[(size_t)j * 2 * D1 + c];
        }
        __syncwarp();
        for (int q = 0; q < cnt; ++q) {
          const double *rec = mystage + q * CD::REC_O;
#pragma unroll
          for (int a = 0; a < NACC; ++a)
            if (pm[a] >= 0) acc[a] += rec[pm[a]] * (rec[NM + pc[a]] * rec[NM + D1 + pcc[a]]);
        }
      }
      // block(lo, hi)[(r, c), (r', c')] = -M[r, r'] h_lo[c] h_hi[c'];  M symmetric -> also (r', c), (r, c')
#pragma unroll
      for (int a = 0; a < NACC; ++a) {
        if (pm[a] < 0) continue;
        int r, rr;
        sym_pair<D>(pm[a], r, rr);
        const int c = pc[a], cc = pcc[a];
        const double v = -acc[a];
        const int i0 = coarse_index<D>(lo, nsegfree, nb, r, c), i1 = coarse_index<D>(hi, nsegfree, nb, rr, cc);
        if (i0 >= 0 && i1 >= 0) A[i0 * lda + i1] += v;
        if (r != rr) {
          const int i2 = coarse_index<D>(lo, nsegfree, nb, rr, c), i3 = coarse_index<D>(hi, nsegfree, nb, r, cc);
          if (i2 >= 0 && i3 >= 0) A[i2 * lda + i3] += v;
        }
      }
    }
  }
}

// Build + invert.  One CTA (1024 threads) per instance that is in a line-search tick.
template <int D, int TS>
__device__ __forceinline__ void coarse_build_body(DevProblem P, SolverVecs V, InstState *st, double reg, int every,
                                                  const int inst) {
  using CD = CoarseDims<D>;
  constexpr int NP = 32 * TS;  // padded matrix dimension
  constexpr int NS = NP + 1;   // shared-memory row stride (odd: transposed reads are bank-conflict free)
  constexpr int NW = kCoarseThreads / 32;
  extern __shared__ double sm[];
  const int n = P.c_n[inst];
  if (n <= 0 || n > NP || st[inst].phase != PH_LS || st[inst].eval_now) return;
  // lagged coarse level: within one barrier stage the inverse is reused for `every` Newton steps
  if (every > 1 && st[inst].mu == st[inst].mu_c && st[inst].c_age < every) {
    __syncthreads();
    if (threadIdx.x == 0) st[inst].c_age += 1;
    return;
  }
  double *A = sm;                       // NP x NS
  double *stage = A + NP * NS;          // NW x STAGE (NP * NS is even: 16-byte aligned)
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (int i = tid; i < NP * NS; i += kCoarseThreads) A[i] = 0.0;
  __syncthreads();
  const int nb = P.c_nb[inst];
  double *mystage = stage + wid * CD::STAGE;

  coarse_accumulate<D>(P, V, inst, reg, A, NS, wid, NW, true, mystage);
  __syncthreads();
  // landmark priors (w ||l - prior||^2) add 2 w on the diagonal
  if (tid == 0) {
    for (int pl = P.prior_off[inst]; pl < P.prior_off[inst + 1]; ++pl) {
      const int q = P.prior_l[pl];
      for (int r = 0; r < D; ++r) A[(nb + q * D + r) * NS + nb + q * D + r] += 2.0 * P.prior_w[pl];
    }
  }
  __syncthreads();
  // ---- load the register tile (mirroring the upper blocks), fix empty / padded coordinates
  const int ty = wid, tx = lane;  // tile row / column
  double Tl[TS][TS];
#pragma unroll
  for (int r = 0; r < TS; ++r)
#pragma unroll
    for (int c = 0; c < TS; ++c) {
      const int i = ty * TS + r, j = tx * TS + c;
      double v = (i <= j) ? A[i * NS + j] : A[j * NS + i];
      if (i == j && (i >= n || !(v > 0.0))) v = 1.0;  // padding / coordinate without curvature: identity
      Tl[r][c] = v;
    }
  // ---- blocked symmetric sweep (SPD, no pivoting), one TS x TS pivot block K per step:
  //   A_KK <- -W,  A_KJ <- W A_KJ,  A_IK <- A_IK W,  A_IJ <- A_IJ - A_IK W A_KJ,      W = A_KK^-1,
  // which is the composition of the TS scalar sweeps of the block, so after all blocks the tiles hold -A^-1.
  // The pivot rows belong to one warp (ty == kt).  It inverts the pivot block with one lane per entry, then
  // publishes two TS x NP panels: the raw rows (by symmetry also the column factors A_IK) and the scaled rows
  // W A_KJ.  Publishing A_KK - I in place of A_KK makes the generic rank-TS update produce the new pivot rows
  // and columns as well; only the pivot block itself is patched.  Two barriers per TS pivots.
  // Layout of the scaled panel: the TS columns of lane tx are split into pairs, pair h of all lanes contiguous, so
  // that every 128-bit shared load of a warp covers one contiguous 512-byte run (a 32-byte lane stride would make
  // each quarter-warp hit every bank twice); the update loop is bound by shared-memory wavefronts, not by FP64.
  auto scl_index = [](int lane_col, int c) {
    return (TS % 2 == 0) ? (c >> 1) * (2 * 32) + lane_col * 2 + (c & 1) : lane_col * TS + c;
  };
  double *panel = stage;                // [2][2][TS][NP], double-buffered; the accumulation staging is idle now
  double *wbuf = panel + 4 * TS * NP;   // [2][TS * TS]
  static_assert(4 * TS * NP + 2 * TS * TS <= NW * CD::STAGE, "pivot panels must fit in the staging area");
  const int nblk = (n + TS - 1) / TS;
  // reciprocal of a pivot: hardware seed + two Newton steps (the quotient routine is 2-3x the latency, and the
  // TS dependent pivots of every block sit on the sweep's critical path)
  auto rcp = [](double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(fma(-x, r, 1.0), r, r);
    r = fma(fma(-x, r, 1.0), r, r);
    return r;
  };
  // pivot warp of block kt: invert the pivot block, publish the raw and scaled row panels
  auto factor_publish = [&](const int kt) {
    double *raw = panel + (kt & 1) * 2 * TS * NP, *scl = raw + TS * NP, *wb = wbuf + (kt & 1) * TS * TS;
    if (tx == kt) {
#pragma unroll
      for (int r = 0; r < TS; ++r) {
#pragma unroll
        for (int c = 0; c < TS; ++c) wb[r * TS + c] = Tl[r][c];
        Tl[r][r] -= 1.0;
      }
    }
    __syncwarp();
    // W = A_KK^-1 by TS scalar sweeps, lane (i, j) holding entry (i, j) (lanes >= TS^2 idle along)
    const int wi = (lane / TS) % TS, wj = lane % TS;
    double w = (lane < TS * TS) ? wb[lane] : 1.0;
#pragma unroll
    for (int k = 0; k < TS; ++k) {
      const double piv = __shfl_sync(0xffffffffu, w, k * TS + k);
      const double wik = __shfl_sync(0xffffffffu, w, wi * TS + k), wkj = __shfl_sync(0xffffffffu, w, k * TS + wj);
      const double inv = rcp(piv);
      if (wi == k && wj == k)
        w = -inv;
      else if (wi == k || wj == k)
        w *= inv;
      else
        w -= wik * wkj * inv;
    }
    __syncwarp();
    if (lane < TS * TS) wb[lane] = -w;
    __syncwarp();
#pragma unroll
    for (int r = 0; r < TS; ++r) {
      double wr[TS];
#pragma unroll
      for (int k = 0; k < TS; ++k) wr[k] = wb[r * TS + k];
#pragma unroll
      for (int c = 0; c < TS; ++c) {
        double sv = 0.0;
#pragma unroll
        for (int k = 0; k < TS; ++k) sv += wr[k] * Tl[k][c];
        raw[r * NP + tx * TS + c] = Tl[r][c];
        scl[r * NP + scl_index(tx, c)] = sv;
      }
    }
  };
  // rank-TS update of this thread's tile with panel kt
  auto apply_panel = [&](const int kt) {
    const double *raw = panel + (kt & 1) * 2 * TS * NP, *scl = raw + TS * NP;
#pragma unroll
    for (int k = 0; k < TS; ++k) {
      double crk[TS], sck[TS];
#pragma unroll
      for (int r = 0; r < TS; ++r) crk[r] = raw[k * NP + ty * TS + r];
#pragma unroll
      for (int c = 0; c < TS; ++c) sck[c] = scl[k * NP + scl_index(tx, c)];
#pragma unroll
      for (int r = 0; r < TS; ++r)
#pragma unroll
        for (int c = 0; c < TS; ++c) Tl[r][c] -= crk[r] * sck[c];
    }
  };
  // Schedule.  The chain "apply panel kt to the rows of block kt+1, factor block kt+1" is the critical path; the
  // other 31 warps' updates are bound by shared-memory wavefronts.  So after panel kt appears the next pivot warp
  // updates its rows alone (second barrier), then factors block kt+1 into the other panel buffer WHILE the other
  // warps apply panel kt.
  if (ty == 0) factor_publish(0);
  for (int kt = 0; kt < nblk; ++kt) {
    __syncthreads();  // panel kt is published; every warp is done with panel kt-1 (its buffer is free again)
    const bool next = (ty == kt + 1) && (kt + 1 < nblk);  // warp-uniform
    if (next) apply_panel(kt);
    __syncthreads();
    if (next) {
      factor_publish(kt + 1);
    } else {
      apply_panel(kt);
      if (ty == kt && tx == kt) {
        const double *wb = wbuf + (kt & 1) * TS * TS;
#pragma unroll
        for (int r = 0; r < TS; ++r)
#pragma unroll
          for (int c = 0; c < TS; ++c) Tl[r][c] = -wb[r * TS + c];
      }
    }
  }
  // ---- write back (negated), symmetrise, store
  __syncthreads();
#pragma unroll
  for (int r = 0; r < TS; ++r)
#pragma unroll
    for (int c = 0; c < TS; ++c) A[(ty * TS + r) * NS + tx * TS + c] = -Tl[r][c];
  __syncthreads();
  double *out = P.c_Ainv + P.c_moff[inst];
  for (int i = wid; i < n; i += NW)
    for (int j = lane; j < n; j += 32) out[i * n + j] = 0.5 * (A[i * NS + j] + A[j * NS + i]);  // exactly symmetric
  if (tid == 0) {
    st[inst].mu_c = st[inst].mu;
    st[inst].c_age = 1;
  }
}

template <int D, int TS>
__global__ void __launch_bounds__(kCoarseThreads) k_coarse_build(DevProblem P, SolverVecs V, InstState *st, double reg,
                                                                int every, WorkLists W) {
  const int *act;
  int n_act;
  wl_get(W, WL_LS, act, n_act);
  for (int ai = blockIdx.x; ai < n_act; ai += gridDim.x) {
    coarse_build_body<D, TS>(P, V, st, reg, every, act[ai]);
    __syncthreads();  // shared matrix / staging are reused by the next instance
  }
}

template <int D>
inline size_t coarse_smem_bytes_d(int ts) {
  const size_t np = 32 * (size_t)ts;
  return sizeof(double) * (np * (np + 1) + (size_t)(kCoarseThreads / 32) * CoarseDims<D>::STAGE);
}
inline size_t coarse_smem_bytes(int d, int ts) { return d == 2 ? coarse_smem_bytes_d<2>(ts) : coarse_smem_bytes_d<3>(ts); }
inline int coarse_tile_size(int nmax) { return (nmax + 31) / 32; }

// y = A_c^-1 c ;  scatter: segment bases -> ytmp (start value of the forward prefix sum), landmarks -> s.
constexpr int kCoarseApplyThreads = 256;
template <int D>
__device__ __forceinline__ void coarse_apply_body(DevProblem P, SolverVecs V, const InstState *st, const int inst) {
  constexpr int BLK = D * (D + 1);
  constexpr int NW = kCoarseApplyThreads / 32;
  __shared__ double red[NW];
  __shared__ double cs[kCoarseMax], ys[kCoarseMax];
  const int n = P.c_n[inst];
  if (n <= 0 || n > kCoarseMax || st[inst].phase == PH_DONE || st[inst].phase == PH_WAIT || st[inst].eval_now) return;
  const double *__restrict__ Ai = P.c_Ainv + P.c_moff[inst];
  const double *c = P.c_rhs + P.c_off[inst];
  double *y = P.c_sol + P.c_off[inst];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < n; i += kCoarseApplyThreads) cs[i] = c[i];
  __syncthreads();
  // four rows per warp step: the loads of all four rows are in flight before the reductions
  for (int i = 4 * wid; i < n; i += 4 * NW) {
    double a[4] = {0.0, 0.0, 0.0, 0.0};
    for (int j = lane; j < n; j += 32) {
      const double cj = cs[j];
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (i + t < n) a[t] += __ldg(Ai + (size_t)(i + t) * n + j) * cj;
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) a[t] = warp_sum(a[t]);
    const double mine = lane == 0 ? a[0] : (lane == 1 ? a[1] : (lane == 2 ? a[2] : a[3]));
    if (lane < 4 && i + lane < n) ys[i + lane] = mine;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += kCoarseApplyThreads) y[i] = ys[i];
  const int nb = P.c_nb[inst];
  const int seg0 = P.seg_begin[inst];
  for (int i = threadIdx.x; i < nb; i += kCoarseApplyThreads) {
    const int sl = i / BLK;
    const int pg = P.seg_ptr[seg0 + 1 + sl];  // base pose of free segment sl
    V.ytmp[P.zoff[inst] + (pg - P.pose_off[inst]) * BLK + (i % BLK)] = ys[i];
  }
  const int Pi = P.pose_off[inst + 1] - P.pose_off[inst];
  const int c0 = P.zoff[inst] + Pi * BLK;
  double acc = 0.0;
  for (int j = threadIdx.x; j < n - nb; j += kCoarseApplyThreads) {
    const double sv = ys[nb + j];
    V.s[c0 + j] = sv;
    acc += sv * V.r[c0 + j];
  }
  const double tot = block_sum<kCoarseApplyThreads>(acc, red);
  if (threadIdx.x == 0) V.part_lm[inst] = tot;
}

template <int D>
__global__ void __launch_bounds__(kCoarseApplyThreads) k_coarse_apply(DevProblem P, SolverVecs V, const InstState *st,
                                                                     WorkLists W) {
  const int *act;
  int n_act;
  wl_get(W, WL_RUN, act, n_act);
  for (int ai = blockIdx.x; ai < n_act; ai += gridDim.x) {
    coarse_apply_body<D>(P, V, st, act[ai]);
    __syncthreads();
  }
}

// ---- large coarse spaces (kCoarseMax < nc <= kCoarseBigMax) -------------------------------------------------
// Same sorted-list accumulation, but into a dense nc x nc work matrix in global memory, inverted by the blocked
// symmetric sweeps of dense.cuh; the application is a dense mat-vec (one warp per row).  Every kernel takes the
// instance it serves, so a batch may hold any number of such instances (launched one after the other).

constexpr int kBigThreads = 256;

template <int D>
__global__ void __launch_bounds__(kBigThreads) k_coarse_big_accum(DevProblem P, SolverVecs V, double reg, double *A, int inst,
                                                                  const InstState *st) {
  __shared__ double stage[(kBigThreads / 32) * CoarseDims<D>::STAGE];
  if (st[inst].phase != PH_LS || st[inst].eval_now) return;
  const int wid = threadIdx.x >> 5;
  const int gw = blockIdx.x * (kBigThreads / 32) + wid, nw = gridDim.x * (kBigThreads / 32);
  coarse_accumulate<D>(P, V, inst, reg, A, P.c_n[inst], gw, nw, false, stage + wid * CoarseDims<D>::STAGE);
}

// priors on the landmark diagonals, identity for coordinates without curvature, lower triangle mirrored from the
// upper one (the accumulation fills the diagonal blocks and the upper off-diagonal blocks)
template <int D>
__global__ void __launch_bounds__(256) k_coarse_big_finish(DevProblem P, double *A, int inst, const InstState *st) {
  if (st[inst].phase != PH_LS || st[inst].eval_now) return;
  const int n = P.c_n[inst], nb = P.c_nb[inst];
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n * n) return;
  const int i = (int)(t / n), j = (int)(t % n);
  if (i == j) {
    double v = A[t];
    if (i >= nb) {
      const int q = (i - nb) / D;
      for (int pl = P.prior_off[inst]; pl < P.prior_off[inst + 1]; ++pl)
        if (P.prior_l[pl] == q) v += 2.0 * P.prior_w[pl];
    }
    if (!(v > 0.0)) v = 1.0;
    A[t] = v;
  } else if (i > j) {
    const int bi = (i < nb) ? i / (D * (D + 1)) : nb + (i - nb) / D, bj = (j < nb) ? j / (D * (D + 1)) : nb + (j - nb) / D;
    if (bi != bj) A[t] = A[(size_t)j * n + i];  // off-diagonal blocks exist in the upper triangle only
  }
}

// y = A_c^-1 c : one warp per row
__global__ void __launch_bounds__(kBigThreads) k_coarse_big_apply(DevProblem P, const InstState *st, int inst) {
  if (st[inst].phase == PH_DONE || st[inst].phase == PH_WAIT || st[inst].eval_now) return;
  const int n = P.c_n[inst], lane = threadIdx.x & 31;
  const int row = blockIdx.x * (kBigThreads / 32) + (threadIdx.x >> 5);
  if (row >= n) return;
  const double *__restrict__ a = P.c_Ainv + P.c_moff[inst] + (size_t)row * n;
  const double *__restrict__ c = P.c_rhs + P.c_off[inst];
  double acc0 = 0.0, acc1 = 0.0;
  int j = lane;
  for (; j + 32 < n; j += 64) {
    acc0 += __ldg(a + j) * c[j];
    acc1 += __ldg(a + j + 32) * c[j + 32];
  }
  if (j < n) acc0 += __ldg(a + j) * c[j];
  const double tot = warp_sum(acc0 + acc1);
  if (lane == 0) P.c_sol[P.c_off[inst] + row] = tot;
}

// scatter: segment bases -> ytmp, landmarks -> s, partial r.s of the landmark block
template <int D>
__global__ void __launch_bounds__(kBigThreads) k_coarse_big_scatter(DevProblem P, SolverVecs V, const InstState *st, int inst) {
  constexpr int BLK = D * (D + 1);
  __shared__ double red[kBigThreads / 32];
  if (st[inst].phase == PH_DONE || st[inst].phase == PH_WAIT || st[inst].eval_now) return;
  const int n = P.c_n[inst], nb = P.c_nb[inst];
  const double *y = P.c_sol + P.c_off[inst];
  const int seg0 = P.seg_begin[inst];
  for (int i = threadIdx.x; i < nb; i += kBigThreads) {
    const int pg = P.seg_ptr[seg0 + 1 + i / BLK];  // base pose of free segment i / BLK
    V.ytmp[(size_t)P.zoff[inst] + (size_t)(pg - P.pose_off[inst]) * BLK + (i % BLK)] = y[i];
  }
  const int c0 = P.zoff[inst] + (P.pose_off[inst + 1] - P.pose_off[inst]) * BLK;
  double acc = 0.0;
  for (int j = threadIdx.x; j < n - nb; j += kBigThreads) {
    const double sv = y[nb + j];
    V.s[c0 + j] = sv;
    acc += sv * V.r[c0 + j];
  }
  const double tot = block_sum<kBigThreads>(acc, red);
  if (threadIdx.x == 0) V.part_lm[inst] = tot;
}

}  // namespace score
