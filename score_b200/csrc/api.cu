// libscore_b200 — C ABI (include/score_b200.h) over the sm_100a kernels.
#include <math.h>
#include <string.h>

#include <vector>

#include "assemble.cuh"
#include "common.cuh"
#include "extract.cuh"
#include "precond.cuh"
#include "solver.cuh"

thread_local std::string g_score_last_error;

using namespace score;

struct ScoreHandle_ {
  int device = 0;
  DevProblem P{};
  SolverVecs V{};
  BlockTables T{};
  InstState *st = nullptr;
  int *d_ndone = nullptr;
  int *h_ndone = nullptr;  // pinned
  double *wsum = nullptr;
  int *nnz_row = nullptr;
  // transpose scratch
  int *sort_keys = nullptr, *sort_idx = nullptr, *sort_perm = nullptr;
  void *sort_tmp = nullptr;
  size_t sort_tmp_bytes = 0;
  // outputs
  double *out_poses = nullptr, *out_lms = nullptr, *out_round = nullptr, *out_dist = nullptr;
  // host copies of the offset tables
  std::vector<int> pose_off, lm_off, edge_off, rng_off, prior_off, zoff, roff, nnzoff, seg_begin;
  std::vector<int> rb_begin, cb_begin;
  std::vector<int> c_off, c_moff, c_n, c_nb;
  int c_nmax = 0;
  std::vector<void *> allocs;
  cudaStream_t own_stream = nullptr;
  cudaGraphExec_t graph_exec = nullptr;
  int graph_ticks = 0;
  SolverCfg graph_cfg{};
  bool solved_once = false;
  int dist_per = 0;
};

namespace {

template <typename T>
int dalloc(ScoreHandle_ *h, T **ptr, size_t n) {
  *ptr = nullptr;
  if (n == 0) n = 1;
  cudaError_t e = cudaMalloc((void **)ptr, n * sizeof(T));
  if (e != cudaSuccess) {
    g_score_last_error = std::string("cudaMalloc failed: ") + cudaGetErrorString(e);
    return SCORE_ERR_ALLOC;
  }
  h->allocs.push_back((void *)*ptr);
  return SCORE_OK;
}

template <typename T>
int upload(ScoreHandle_ *h, T **dst, const T *src, size_t n) {
  int rc = dalloc(h, dst, n);
  if (rc) return rc;
  if (n && src) SCORE_CUDA_CHECK(cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyDefault));
  return SCORE_OK;
}

int fetch_offsets(const int32_t *src, int n_inst, int64_t total, std::vector<int> &dst, const char *name) {
  dst.assign(n_inst + 1, 0);
  if (src == nullptr) {
    if (n_inst != 1) {
      g_score_last_error = std::string(name) + " offsets are required when n_instances > 1";
      return SCORE_ERR_INVALID;
    }
    dst[1] = (int)total;
    return SCORE_OK;
  }
  SCORE_CUDA_CHECK(cudaMemcpy(dst.data(), src, sizeof(int) * (n_inst + 1), cudaMemcpyDefault));
  if (dst[0] != 0 || dst[n_inst] != (int)total) {
    g_score_last_error = std::string(name) + " offsets do not span [0, total]";
    return SCORE_ERR_INVALID;
  }
  for (int i = 0; i < n_inst; ++i)
    if (dst[i + 1] < dst[i]) {
      g_score_last_error = std::string(name) + " offsets are not monotone";
      return SCORE_ERR_INVALID;
    }
  return SCORE_OK;
}

int grid_for(long n, int threads) { return (int)((n + threads - 1) / threads); }

// Bytes one PCG tick of an instance must move (fp64 values, int32 indices; DESIGN.md "algorithmic bytes").
double bytes_cg_tick(int d, double nnz, double m, double nz, double K, double Pn, double nc) {
  const double blk = d * (d + 1), d1 = d + 1;
  double b = 0.0;
  b += 12.0 * nnz + 4.0 * (m + 1) + 8.0 * nz + 8.0 * 2.0 * m + 8.0 * (d * K + 2.0 * K);  // rowpass
  b += 12.0 * nnz + 4.0 * (nz + 1) + 8.0 * m + 8.0 * 5.0 * nz;                            // colpass
  b += 8.0 * 5.0 * nz + 8.0 * Pn * (2.0 * blk + d1 * d1);   // precond rev+fwd (r twice, ytmp w+r, s; G twice, M)
  b += 8.0 * (nc * nc + 3.0 * nc);                           // coarse apply
  b += 8.0 * 3.0 * nz;                                       // pupdate
  return b;
}
double bytes_ls_tick(int d, double nnz, double m, double nz, double K, double Pn, double nc) {
  const double blk = d * (d + 1), d1 = d + 1;
  double b = 0.0;
  b += 12.0 * nnz + 4.0 * (m + 1) + 8.0 * nz + 8.0 * m;                  // rowpass: bdz = B dz
  b += 8.0 * 2.0 * m + 8.0 * (m - d * K) + 16.0 * K;                     // linesearch: res, bdz, w(plain), r~, w(range)
  b += 8.0 * 5.0 * m + 8.0 * 3.0 * K;                                    // rowupdate: res rw, bdz, w, u; r~, ctan, crad
  b += 8.0 * (d * K + 6.0 * K) + 8.0 * nc * nc;                          // coarse build: res, factors, slots, G column; inverse out
  b += 12.0 * nnz + 4.0 * (nz + 1) + 8.0 * m + 8.0 * 5.0 * nz;           // colpass: z rw, dz rw, r w
  b += 8.0 * 5.0 * nz + 8.0 * Pn * (2.0 * blk + d1 * d1);                // precond
  b += 8.0 * (nc * nc + 3.0 * nc);                                       // coarse apply
  b += 8.0 * 3.0 * nz;                                                   // pupdate (p = s)
  return b;
}

}  // namespace

extern "C" const char *score_last_error(void) { return g_score_last_error.c_str(); }
extern "C" const char *score_version(void) { return "score_b200 0.1.0 (sm_100a)"; }

extern "C" void score_destroy(ScoreHandle h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->graph_exec) cudaGraphExecDestroy(h->graph_exec);
  for (void *p : h->allocs) cudaFree(p);
  if (h->sort_tmp) cudaFree(h->sort_tmp);
  if (h->h_ndone) cudaFreeHost(h->h_ndone);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  delete h;
}

static int create_impl(const ScoreProblemDesc *desc, int32_t device, ScoreHandle_ *h) {
  const int d = desc->dim;
  if (d != 2 && d != 3) {
    g_score_last_error = "Value " + std::to_string(d) + " is not 2 or 3";
    return SCORE_ERR_INVALID;
  }
  if (desc->relaxation != SCORE_RELAX_QCQP && desc->relaxation != SCORE_RELAX_SOCP) {
    g_score_last_error = "unknown relaxation";
    return SCORE_ERR_INVALID;
  }
  if (desc->n_instances < 1 || desc->P < 1 || desc->L < 0 || desc->E < 0 || desc->K < 0 || desc->Lp < 0 ||
      desc->n_seg < 1) {
    g_score_last_error = "invalid sizes in ScoreProblemDesc";
    return SCORE_ERR_INVALID;
  }
  const long long blk = d * (d + 1), rpe = d + d * d, npe = d * (d + 2) + d * d * (d + 1);
  const long long nz = desc->P * blk + desc->L * d;
  const long long m = desc->E * rpe + desc->K * d + desc->Lp * d;
  const long long nnz = desc->E * npe + desc->K * 2 * d + desc->Lp * d;
  if (nz >= (1ll << 31) || m >= (1ll << 31) || nnz >= (1ll << 31) || nnz + desc->K * d >= (1ll << 31)) {
    g_score_last_error = "problem too large for 32-bit indexing";
    return SCORE_ERR_INVALID;
  }
  int ndev = 0;
  SCORE_CUDA_CHECK(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) {
    g_score_last_error = "no such CUDA device";
    return SCORE_ERR_CUDA;
  }
  SCORE_CUDA_CHECK(cudaSetDevice(device));
  h->device = device;
  DevProblem &P = h->P;
  P.d = d;
  P.blk = (int)blk;
  P.rpe = (int)rpe;
  P.npe = (int)npe;
  P.relax = desc->relaxation;
  P.n_inst = desc->n_instances;
  P.P = (int)desc->P;
  P.L = (int)desc->L;
  P.E = (int)desc->E;
  P.K = (int)desc->K;
  P.Lp = (int)desc->Lp;
  P.n_seg = (int)desc->n_seg;
  P.nz = (int)nz;
  P.m = (int)m;
  P.nnz = (int)nnz;
  h->dist_per = (desc->relaxation == SCORE_RELAX_QCQP) ? d : 1;
  const int NI = P.n_inst;
  int rc;
  if ((rc = fetch_offsets(desc->pose_off, NI, desc->P, h->pose_off, "pose"))) return rc;
  if ((rc = fetch_offsets(desc->lm_off, NI, desc->L, h->lm_off, "landmark"))) return rc;
  if ((rc = fetch_offsets(desc->edge_off, NI, desc->E, h->edge_off, "edge"))) return rc;
  if ((rc = fetch_offsets(desc->rng_off, NI, desc->K, h->rng_off, "range"))) return rc;
  if ((rc = fetch_offsets(desc->prior_off, NI, desc->Lp, h->prior_off, "prior"))) return rc;
  h->zoff.resize(NI + 1);
  h->roff.resize(NI + 1);
  h->nnzoff.resize(NI + 1);
  for (int i = 0; i <= NI; ++i) {
    h->zoff[i] = h->pose_off[i] * (int)blk + h->lm_off[i] * d;
    h->roff[i] = h->edge_off[i] * (int)rpe + h->rng_off[i] * d + h->prior_off[i] * d;
    h->nnzoff[i] = h->edge_off[i] * (int)npe + h->rng_off[i] * 2 * d + h->prior_off[i] * d;
  }
  for (int i = 0; i < NI; ++i)
    if (h->pose_off[i + 1] == h->pose_off[i]) {
      g_score_last_error = "instance without poses";  // IndexError in the reference (gurobi_utils.py:181)
      return SCORE_ERR_INVALID;
    }
  // segments
  std::vector<int> seg_ptr(P.n_seg + 1), seg_inst(P.n_seg);
  SCORE_CUDA_CHECK(cudaMemcpy(seg_ptr.data(), desc->seg_ptr, sizeof(int) * (P.n_seg + 1), cudaMemcpyDefault));
  SCORE_CUDA_CHECK(cudaMemcpy(seg_inst.data(), desc->seg_inst, sizeof(int) * P.n_seg, cudaMemcpyDefault));
  if (seg_ptr[0] != 0 || seg_ptr[P.n_seg] != P.P) {
    g_score_last_error = "seg_ptr does not span the poses";
    return SCORE_ERR_INVALID;
  }
  h->seg_begin.assign(NI + 1, 0);
  {
    int s = 0;
    for (int i = 0; i < NI; ++i) {
      h->seg_begin[i] = s;
      while (s < P.n_seg && seg_inst[s] == i) {
        if (seg_ptr[s] < h->pose_off[i] || seg_ptr[s + 1] > h->pose_off[i + 1] || seg_ptr[s + 1] <= seg_ptr[s]) {
          g_score_last_error = "segment table inconsistent with pose offsets";
          return SCORE_ERR_INVALID;
        }
        ++s;
      }
      if (s == h->seg_begin[i] || seg_ptr[h->seg_begin[i]] != h->pose_off[i] || seg_ptr[s] != h->pose_off[i + 1]) {
        g_score_last_error = "segments do not tile the poses of an instance";
        return SCORE_ERR_INVALID;
      }
    }
    h->seg_begin[NI] = s;
    if (s != P.n_seg) {
      g_score_last_error = "seg_inst is not sorted by instance";
      return SCORE_ERR_INVALID;
    }
  }
#define UP(field, src, n)                                   \
  if ((rc = upload(h, &P.field, src, (size_t)(n)))) return rc;
  UP(pose_off, h->pose_off.data(), NI + 1)
  UP(lm_off, h->lm_off.data(), NI + 1)
  UP(edge_off, h->edge_off.data(), NI + 1)
  UP(rng_off, h->rng_off.data(), NI + 1)
  UP(prior_off, h->prior_off.data(), NI + 1)
  UP(zoff, h->zoff.data(), NI + 1)
  UP(roff, h->roff.data(), NI + 1)
  UP(nnzoff, h->nnzoff.data(), NI + 1)
  UP(seg_begin, h->seg_begin.data(), NI + 1)
  UP(seg_ptr, seg_ptr.data(), P.n_seg + 1)
  UP(seg_inst, seg_inst.data(), P.n_seg)
  UP(link_edge, desc->link_edge, P.P)
  UP(edge_i, desc->edge_i, P.E)
  UP(edge_j, desc->edge_j, P.E)
  UP(edge_t, desc->edge_t, (size_t)P.E * d)
  UP(edge_R, desc->edge_R, (size_t)P.E * d * d)
  UP(edge_k, desc->edge_k, P.E)
  UP(edge_tau, desc->edge_tau, P.E)
  UP(rng_a, desc->rng_a, P.K)
  UP(rng_b, desc->rng_b, P.K)
  UP(rng_dist, desc->rng_dist, P.K)
  UP(rng_w, desc->rng_w, P.K)
  UP(prior_l, desc->prior_l, P.Lp)
  UP(prior_t, desc->prior_t, (size_t)P.Lp * d)
  UP(prior_w, desc->prior_w, P.Lp)
#undef UP
#define DA(ptr, n) \
  if ((rc = dalloc(h, &(ptr), (size_t)(n)))) return rc;
  DA(P.indptr, P.m + 1)
  DA(P.cols, P.nnz)
  DA(P.vals, P.nnz)
  DA(P.t_indptr, P.nz + 1)
  DA(P.t_rows, P.nnz)
  DA(P.t_vals, P.nnz)
  DA(P.w, P.m)
  DA(P.b, P.m)
  DA(P.G, (size_t)P.P * blk)
  DA(P.M, (size_t)P.P * (d + 1) * (d + 1))
  DA(P.lm_inv, (size_t)P.L * d)
  DA(h->wsum, P.P)
  DA(h->nnz_row, P.nnz)
  DA(h->sort_keys, P.nnz)
  DA(h->sort_idx, P.nnz)
  DA(h->sort_perm, P.nnz)
  SolverVecs &V = h->V;
  DA(V.z, P.nz)
  DA(V.dz, P.nz)
  DA(V.r, P.nz)
  DA(V.s, P.nz)
  DA(V.p, P.nz)
  DA(V.ytmp, P.nz)
  DA(V.res, P.m)
  DA(V.u, P.m)
  DA(V.bdz, P.m)
  DA(V.ctan, P.K)
  DA(V.crad, P.K)
  DA(P.rng_slot, 2 * (size_t)P.K)
  // coarse level: free segment bases + landmarks of every instance (dense, when it fits shared memory)
  {
    h->c_off.assign(NI + 1, 0);
    h->c_moff.assign(NI + 1, 0);
    h->c_n.assign(NI, 0);
    h->c_nb.assign(NI, 0);
    long long moff = 0;
    for (int i = 0; i < NI; ++i) {
      const int nsegfree = h->seg_begin[i + 1] - h->seg_begin[i] - 1;
      const int nb = nsegfree * (int)blk, nc = nb + (h->lm_off[i + 1] - h->lm_off[i]) * d;
      const bool on = nc > 0 && nc <= kCoarseMax && coarse_smem_bytes(d, nc) <= 232448;
      h->c_n[i] = on ? nc : 0;
      h->c_nb[i] = on ? nb : 0;
      h->c_off[i + 1] = h->c_off[i] + (on ? nc : 0);
      moff += on ? (long long)nc * nc : 0;
      if (moff >= (1ll << 31)) {
        g_score_last_error = "coarse matrices too large for 32-bit indexing";
        return SCORE_ERR_INVALID;
      }
      h->c_moff[i + 1] = (int)moff;
      if (on && nc > h->c_nmax) h->c_nmax = nc;
    }
    if ((rc = upload(h, &P.c_off, h->c_off.data(), NI + 1))) return rc;
    if ((rc = upload(h, &P.c_moff, h->c_moff.data(), NI + 1))) return rc;
    if ((rc = upload(h, &P.c_n, h->c_n.data(), NI))) return rc;
    if ((rc = upload(h, &P.c_nb, h->c_nb.data(), NI))) return rc;
    DA(P.c_Ainv, (size_t)moff)
    DA(P.c_rhs, h->c_off[NI])
    DA(P.c_sol, h->c_off[NI])
  }
  // block tables
  std::vector<BlockDesc> rb, cb;
  h->rb_begin.assign(NI + 1, 0);
  h->cb_begin.assign(NI + 1, 0);
  for (int i = 0; i < NI; ++i) {
    h->rb_begin[i] = (int)rb.size();
    for (int r0 = h->roff[i]; r0 < h->roff[i + 1]; r0 += kRowsPerBlock)
      rb.push_back({i, r0, std::min(r0 + kRowsPerBlock, h->roff[i + 1]), 0});
    h->cb_begin[i] = (int)cb.size();
    const int pc0 = h->zoff[i], pc1 = pc0 + (h->pose_off[i + 1] - h->pose_off[i]) * (int)blk;
    for (int c0 = pc0; c0 < pc1; c0 += kColsPerBlock) cb.push_back({i, c0, std::min(c0 + kColsPerBlock, pc1), CB_POSE});
    for (int c0 = pc1; c0 < h->zoff[i + 1]; c0 += 64) cb.push_back({i, c0, std::min(c0 + 64, h->zoff[i + 1]), CB_LANDMARK});
  }
  h->rb_begin[NI] = (int)rb.size();
  h->cb_begin[NI] = (int)cb.size();
  h->T.n_rb = (int)rb.size();
  h->T.n_cb = (int)cb.size();
  if ((rc = upload(h, &h->T.rb, rb.data(), rb.size()))) return rc;
  if ((rc = upload(h, &h->T.cb, cb.data(), cb.size()))) return rc;
  if ((rc = upload(h, &h->T.rb_begin, h->rb_begin.data(), NI + 1))) return rc;
  if ((rc = upload(h, &h->T.cb_begin, h->cb_begin.data(), NI + 1))) return rc;
  DA(V.part_row, rb.size())
  DA(V.part_ls, rb.size() * kLsSums)
  DA(V.part_upd, rb.size() * 2)
  DA(V.part_col, cb.size() * 4)
  DA(V.part_seg, P.n_seg)
  DA(V.part_lm, NI)
  DA(h->st, NI)
  DA(h->d_ndone, 1)
  DA(h->out_poses, (size_t)P.P * blk)
  DA(h->out_lms, (size_t)P.L * d)
  DA(h->out_round, (size_t)P.P * d * d)
  DA(h->out_dist, (size_t)P.K * h->dist_per)
#undef DA
  SCORE_CUDA_CHECK(cudaMallocHost((void **)&h->h_ndone, sizeof(int)));
  // radix-sort scratch for the transpose
  int end_bit = 1;
  while ((1ll << end_bit) <= (long long)P.nz) ++end_bit;
  SCORE_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, h->sort_tmp_bytes, P.cols, h->sort_keys, h->sort_idx,
                                                   h->sort_perm, P.nnz, 0, end_bit, (cudaStream_t)0));
  SCORE_CUDA_CHECK(cudaMalloc(&h->sort_tmp, h->sort_tmp_bytes ? h->sort_tmp_bytes : 1));
  SCORE_CUDA_CHECK(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
  SCORE_CUDA_CHECK(cudaDeviceSynchronize());
  return SCORE_OK;
}

extern "C" int score_create(const ScoreProblemDesc *desc, int32_t device, ScoreHandle *out) {
  if (!desc || !out) {
    g_score_last_error = "null argument";
    return SCORE_ERR_INVALID;
  }
  *out = nullptr;
  ScoreHandle_ *h = new ScoreHandle_();
  int rc = create_impl(desc, device, h);
  if (rc != SCORE_OK) {
    std::string keep = g_score_last_error;
    score_destroy(h);
    g_score_last_error = keep;
    return rc;
  }
  *out = h;
  return SCORE_OK;
}

constexpr int kKernelsPerTick = 11;

// One solver tick.  `ev` (optional, kKernelsPerTick + 1 events) brackets every kernel for profiling.
template <int D>
static void launch_tick(ScoreHandle_ *h, const SolverCfg &cfg, cudaStream_t st, cudaEvent_t *ev = nullptr) {
  const DevProblem &P = h->P;
  int k = 0;
  auto mark = [&]() {
    if (ev) cudaEventRecord(ev[k++], st);
  };
  mark();
  k_rowpass<D><<<h->T.n_rb, kThreads, 0, st>>>(P, h->V, h->T, h->st);
  mark();
  k_linesearch<D><<<h->T.n_rb, kThreads, 0, st>>>(P, h->V, h->T, h->st);
  mark();
  k_ctrl_a<<<P.n_inst, kSegThreads, 0, st>>>(h->V, h->T, h->st, cfg);
  mark();
  k_rowupdate<D><<<h->T.n_rb, kThreads, 0, st>>>(P, h->V, h->T, h->st);
  mark();
  if (h->c_nmax > 0)
    k_coarse_build<D><<<P.n_inst, kCoarseThreads, coarse_smem_bytes_d<D>(h->c_nmax), st>>>(P, h->V, h->st, cfg.coarse_reg);
  mark();
  k_colpass<<<h->T.n_cb, kThreads, 0, st>>>(P, h->V, h->T, h->st);
  mark();
  k_precond_rev<D><<<P.n_seg + P.n_inst, kSegThreads, 0, st>>>(P, h->V, h->st);
  mark();
  if (h->c_nmax > 0) k_coarse_apply<D><<<P.n_inst, kSegThreads, 0, st>>>(P, h->V, h->st);
  mark();
  k_precond_fwd<D><<<P.n_seg, kSegThreads, 0, st>>>(P, h->V, h->st);
  mark();
  k_ctrl_b<<<P.n_inst, kSegThreads, 0, st>>>(P, h->V, h->T, h->st, cfg, h->d_ndone);
  mark();
  k_pupdate<<<h->T.n_cb, kThreads, 0, st>>>(h->V, h->T, h->st);
  mark();
}

extern "C" int score_solve(ScoreHandle h, const ScoreParams *params, ScoreStats *stats, ScoreInstanceStats *inst_stats) {
  if (!h) {
    g_score_last_error = "null handle";
    return SCORE_ERR_INVALID;
  }
  ScoreParams prm{};
  if (params) prm = *params;
  SolverCfg cfg;
  cfg.max_newton = prm.max_newton > 0 ? prm.max_newton : 200;
  if (params && prm.max_newton == -1) cfg.max_newton = 0;  // "evaluate the start point only"
  cfg.max_cg = prm.max_cg > 0 ? prm.max_cg : 100;
  cfg.kkt_tol = prm.kkt_tol > 0 ? prm.kkt_tol : 1e-6;
  cfg.forcing = prm.cg_forcing > 0 ? prm.cg_forcing : 0.1;
  cfg.mu0 = prm.mu0 > 0 ? prm.mu0 : (prm.mu0 < 0 ? 0.0 : 1.0);
  cfg.mu_factor = (prm.mu_factor > 0 && prm.mu_factor < 1) ? prm.mu_factor : 0.1;
  cfg.center_tol = prm.center_tol > 0 ? prm.center_tol : 4.0;
  cfg.mu_min = prm.mu_min > 0 ? prm.mu_min : 1e-16;
  cfg.mu_eval = 1e-5;
  cfg.coarse_reg = 1e-6;
  const int max_ticks = prm.max_ticks > 0 ? prm.max_ticks : 200000;
  const int tpl = prm.ticks_per_launch > 0 ? prm.ticks_per_launch : 32;
  SCORE_CUDA_CHECK(cudaSetDevice(h->device));
  cudaStream_t st = prm.stream ? (cudaStream_t)prm.stream : h->own_stream;
  DevProblem &P = h->P;
  SolverVecs &V = h->V;
  const int d = P.d;
  long launches = 0;

  cudaEvent_t ev[5];
  for (auto &e : ev) SCORE_CUDA_CHECK(cudaEventCreate(&e));
  SCORE_CUDA_CHECK(cudaEventRecord(ev[0], st));

  // ---- 1. assembly: reduced operator, its transpose
  {
    AsmOut out{P.indptr, P.cols, P.vals, P.w, P.b, h->nnz_row};
    const long nf = (long)P.E + P.K + P.Lp;
    k_assemble<<<grid_for(nf, 256), 256, 0, st>>>(P, ASM_REDUCED, 0, P.n_inst, out);
    k_iota<<<grid_for(P.nnz, 256), 256, 0, st>>>(h->sort_idx, P.nnz);
    int end_bit = 1;
    while ((1ll << end_bit) <= (long long)P.nz) ++end_bit;
    SCORE_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(h->sort_tmp, h->sort_tmp_bytes, P.cols, h->sort_keys, h->sort_idx,
                                                     h->sort_perm, P.nnz, 0, end_bit, st));
    SCORE_CUDA_CHECK(cudaMemsetAsync(P.t_indptr, 0, sizeof(int) * (P.nz + 1), st));
    k_transpose_fill<<<grid_for(P.nnz, 256), 256, 0, st>>>(P.nnz, P.nz, h->sort_keys, h->sort_perm, h->nnz_row, P.vals,
                                                           P.t_indptr, P.t_rows, P.t_vals);
    launches += 6;
  }
  SCORE_CUDA_CHECK(cudaEventRecord(ev[1], st));
  // ---- 2. preconditioner, start point, solver state
  {
    k_dead_reckon<<<grid_for(P.n_seg, 64), 64, 0, st>>>(P);
    k_diag_setup<<<grid_for((long)P.P + (long)P.L * d, 256), 256, 0, st>>>(P, h->wsum);
    k_build_M<<<grid_for(P.P, 128), 128, 0, st>>>(P, h->wsum);
    k_init_z<<<grid_for(P.nz, 256), 256, 0, st>>>(P, V.z);
    k_residual<<<grid_for(P.m, kThreads), kThreads, 0, st>>>(P, V.z, V.res);
    if (P.K > 0) k_range_slots<<<grid_for(P.K, 256), 256, 0, st>>>(P);
    launches += 6;
    if (h->c_nmax > 0) {
      const size_t smem = coarse_smem_bytes(d, h->c_nmax);
      if (d == 2)
        SCORE_CUDA_CHECK(cudaFuncSetAttribute(k_coarse_build<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      else
        SCORE_CUDA_CHECK(cudaFuncSetAttribute(k_coarse_build<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    for (double *v : {V.dz, V.r, V.s, V.p, V.ytmp})
      SCORE_CUDA_CHECK(cudaMemsetAsync(v, 0, sizeof(double) * P.nz, st));
    SCORE_CUDA_CHECK(cudaMemsetAsync(V.u, 0, sizeof(double) * P.m, st));
    SCORE_CUDA_CHECK(cudaMemsetAsync(V.bdz, 0, sizeof(double) * P.m, st));
    SCORE_CUDA_CHECK(cudaMemsetAsync(V.part_col, 0, sizeof(double) * 4 * h->T.n_cb, st));
    std::vector<InstState> init(P.n_inst);
    memset(init.data(), 0, sizeof(InstState) * P.n_inst);
    for (auto &s : init) {
      s.phase = PH_LS;
      s.skip_ls = 1;
      s.eta = cfg.forcing;
      s.mu = s.mu_ls = cfg.mu0;
    }
    SCORE_CUDA_CHECK(cudaMemcpyAsync(h->st, init.data(), sizeof(InstState) * P.n_inst, cudaMemcpyHostToDevice, st));
    SCORE_CUDA_CHECK(cudaMemsetAsync(h->d_ndone, 0, sizeof(int), st));
    SCORE_CUDA_CHECK(cudaStreamSynchronize(st));  // init vector is on the host stack
  }
  SCORE_CUDA_CHECK(cudaEventRecord(ev[2], st));
  // ---- 3. solver ticks (CUDA graph of `tpl` ticks, replayed until every instance is done)
  if (h->graph_exec == nullptr || h->graph_ticks != tpl || memcmp(&h->graph_cfg, &cfg, sizeof(cfg)) != 0) {
    if (h->graph_exec) {
      cudaGraphExecDestroy(h->graph_exec);
      h->graph_exec = nullptr;
    }
    cudaGraph_t graph;
    SCORE_CUDA_CHECK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    for (int t = 0; t < tpl; ++t) {
      if (d == 2)
        launch_tick<2>(h, cfg, st);
      else
        launch_tick<3>(h, cfg, st);
    }
    SCORE_CUDA_CHECK(cudaStreamEndCapture(st, &graph));
    SCORE_CUDA_CHECK(cudaGraphInstantiate(&h->graph_exec, graph, 0));
    cudaGraphDestroy(graph);
    h->graph_ticks = tpl;
    h->graph_cfg = cfg;
  }
  long ticks = 0;
  double kernel_ms[12] = {0};
  long profiled = 0;
  if (prm.profile_ticks > 0) {
    // un-graphed ticks with an event between every pair of kernels
    const int nskip = prm.profile_skip > 0 ? prm.profile_skip : 0, nprof = prm.profile_ticks;
    for (int t = 0; t < nskip; ++t) {
      if (d == 2)
        launch_tick<2>(h, cfg, st);
      else
        launch_tick<3>(h, cfg, st);
    }
    std::vector<cudaEvent_t> pev((size_t)nprof * (kKernelsPerTick + 1));
    for (auto &e : pev) SCORE_CUDA_CHECK(cudaEventCreate(&e));
    for (int t = 0; t < nprof; ++t) {
      if (d == 2)
        launch_tick<2>(h, cfg, st, &pev[(size_t)t * (kKernelsPerTick + 1)]);
      else
        launch_tick<3>(h, cfg, st, &pev[(size_t)t * (kKernelsPerTick + 1)]);
    }
    SCORE_CUDA_CHECK(cudaStreamSynchronize(st));
    for (int t = 0; t < nprof; ++t)
      for (int k = 0; k < kKernelsPerTick; ++k) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, pev[(size_t)t * (kKernelsPerTick + 1) + k], pev[(size_t)t * (kKernelsPerTick + 1) + k + 1]);
        kernel_ms[k] += ms;
      }
    for (auto &e : pev) cudaEventDestroy(e);
    ticks += nskip + nprof;
    launches += (long)(nskip + nprof) * kKernelsPerTick;
    profiled = nprof;
  }
  while (ticks < max_ticks) {
    SCORE_CUDA_CHECK(cudaGraphLaunch(h->graph_exec, st));
    ticks += tpl;
    launches += (long)tpl * kKernelsPerTick;
    SCORE_CUDA_CHECK(cudaMemcpyAsync(h->h_ndone, h->d_ndone, sizeof(int), cudaMemcpyDeviceToHost, st));
    SCORE_CUDA_CHECK(cudaStreamSynchronize(st));
    if (*h->h_ndone >= P.n_inst) break;
  }
  SCORE_CUDA_CHECK(cudaEventRecord(ev[3], st));
  // ---- 4. extraction
  k_split_z<<<grid_for(P.nz, 256), 256, 0, st>>>(P, V.z, h->out_poses, h->out_lms);
  k_round_so<<<grid_for(P.P, 128), 128, 0, st>>>(d, P.P, h->out_poses, P.blk, d + 1, h->out_round);
  if (P.K > 0) k_distances<<<grid_for(P.K, 256), 256, 0, st>>>(P, V.z, h->out_dist);
  launches += 3;
  SCORE_CUDA_CHECK(cudaEventRecord(ev[4], st));
  SCORE_CUDA_CHECK(cudaStreamSynchronize(st));
  SCORE_CUDA_CHECK(cudaGetLastError());
  h->solved_once = true;

  std::vector<InstState> fin(P.n_inst);
  SCORE_CUDA_CHECK(cudaMemcpy(fin.data(), h->st, sizeof(InstState) * P.n_inst, cudaMemcpyDeviceToHost));
  int n_solved = 0;
  double bytes = 0.0;
  for (int i = 0; i < P.n_inst; ++i) {
    const InstState &s = fin[i];
    n_solved += (s.phase == PH_DONE && s.solved) ? 1 : 0;
    const double nnz_i = h->nnzoff[i + 1] - h->nnzoff[i], m_i = h->roff[i + 1] - h->roff[i];
    const double nz_i = h->zoff[i + 1] - h->zoff[i], K_i = h->rng_off[i + 1] - h->rng_off[i];
    const double P_i = h->pose_off[i + 1] - h->pose_off[i];
    bytes += s.total_cg * bytes_cg_tick(d, nnz_i, m_i, nz_i, K_i, P_i, h->c_n[i]) +
             (s.newton_it + 1.0) * bytes_ls_tick(d, nnz_i, m_i, nz_i, K_i, P_i, h->c_n[i]);
    if (inst_stats) {
      ScoreInstanceStats &o = inst_stats[i];
      o.solved = (s.phase == PH_DONE && s.solved) ? 1 : 0;
      o.newton_iters = s.newton_it;
      o.cg_iters = s.total_cg;
      o.ls_failures = s.ls_fail;
      o.objective = s.F;
      o.rel_kkt = s.kkt;
      o.r_stat = s.r_stat;
      o.r_gap = s.r_gap;
    }
  }
  if (stats) {
    float ms[4];
    for (int i = 0; i < 4; ++i) cudaEventElapsedTime(&ms[i], ev[i], ev[i + 1]);
    stats->n_instances = P.n_inst;
    stats->n_solved = n_solved;
    stats->ticks = ticks;
    stats->kernel_launches = launches;
    stats->assemble_ms = ms[0];
    stats->setup_ms = ms[1];
    stats->solve_ms = ms[2];
    stats->extract_ms = ms[3];
    stats->total_ms = ms[0] + ms[1] + ms[2] + ms[3];
    stats->nnz_reduced = P.nnz;
    stats->rows = P.m;
    stats->cols = P.nz;
    stats->algorithmic_bytes = bytes;
    stats->profiled_ticks = profiled;
    for (int k = 0; k < 12; ++k) stats->kernel_ms[k] = kernel_ms[k];
    // per-launch algorithmic bytes with every instance in the PCG phase
    const double nnz = P.nnz, m = P.m, nz = P.nz, K = P.K, Pn = P.P, blk = P.blk, d1 = d + 1;
    double cmat = 0.0, cvec = 0.0;
    for (int i = 0; i < P.n_inst; ++i) {
      cmat += (double)h->c_n[i] * h->c_n[i];
      cvec += h->c_n[i];
    }
    for (int k = 0; k < 12; ++k) stats->kernel_bytes[k] = 0.0;
    stats->kernel_bytes[0] = 12.0 * nnz + 4.0 * (m + 1) + 8.0 * nz + 16.0 * m + 8.0 * (d * K + 2.0 * K);  // rowpass
    stats->kernel_bytes[2] = 8.0 * h->T.n_rb;                                                              // ctrl_a
    stats->kernel_bytes[5] = 12.0 * nnz + 4.0 * (nz + 1) + 8.0 * m + 40.0 * nz;                            // colpass
    stats->kernel_bytes[6] = 16.0 * nz + 8.0 * Pn * (blk + d1 * d1);                                       // precond_rev
    stats->kernel_bytes[7] = 8.0 * (cmat + 3.0 * cvec);                                                    // coarse_apply
    stats->kernel_bytes[8] = 24.0 * nz + 8.0 * Pn * blk;                                                   // precond_fwd
    stats->kernel_bytes[9] = 8.0 * (P.n_seg + P.n_inst);                                                   // ctrl_b
    stats->kernel_bytes[10] = 24.0 * nz;                                                                   // pupdate
  }
  for (auto &e : ev) cudaEventDestroy(e);
  return SCORE_OK;
}

extern "C" int score_get_sizes(ScoreHandle h, int64_t *n_cols_full, int64_t *n_rows_full, int64_t *nnz_full) {
  if (!h) {
    g_score_last_error = "null handle";
    return SCORE_ERR_INVALID;
  }
  const DevProblem &P = h->P;
  const bool q = P.relax == SCORE_RELAX_QCQP;
  if (n_cols_full) *n_cols_full = (int64_t)P.nz + (int64_t)P.K * h->dist_per;
  if (n_rows_full) *n_rows_full = (int64_t)P.E * P.rpe + (int64_t)P.K * (q ? P.d : 1) + (int64_t)P.Lp * P.d;
  if (nnz_full) *nnz_full = (int64_t)P.E * P.npe + (int64_t)P.K * (q ? 3 * P.d : 1) + (int64_t)P.Lp * P.d;
  return SCORE_OK;
}

extern "C" int score_get_solution(ScoreHandle h, double *pose_blocks, double *pose_rounded, double *landmarks,
                                  double *dist) {
  if (!h) {
    g_score_last_error = "null handle";
    return SCORE_ERR_INVALID;
  }
  if (!h->solved_once) {
    g_score_last_error = "score_get_solution called before score_solve";
    return SCORE_ERR_STATE;
  }
  SCORE_CUDA_CHECK(cudaSetDevice(h->device));
  const DevProblem &P = h->P;
  const int d = P.d;
  if (pose_blocks)
    SCORE_CUDA_CHECK(cudaMemcpy(pose_blocks, h->out_poses, sizeof(double) * (size_t)P.P * P.blk, cudaMemcpyDefault));
  if (pose_rounded)
    SCORE_CUDA_CHECK(cudaMemcpy(pose_rounded, h->out_round, sizeof(double) * (size_t)P.P * d * d, cudaMemcpyDefault));
  if (landmarks && P.L)
    SCORE_CUDA_CHECK(cudaMemcpy(landmarks, h->out_lms, sizeof(double) * (size_t)P.L * d, cudaMemcpyDefault));
  if (dist && P.K)
    SCORE_CUDA_CHECK(cudaMemcpy(dist, h->out_dist, sizeof(double) * (size_t)P.K * h->dist_per, cudaMemcpyDefault));
  return SCORE_OK;
}

extern "C" int score_get_csr(ScoreHandle h, int32_t which, int32_t inst, int64_t *n_rows, int64_t *n_cols, int64_t *nnz,
                             int32_t *indptr, int32_t *indices, double *values, double *weights, double *rhs) {
  if (!h) {
    g_score_last_error = "null handle";
    return SCORE_ERR_INVALID;
  }
  DevProblem &P = h->P;
  if (inst < 0 || inst >= P.n_inst) {
    g_score_last_error = "instance index out of range";
    return SCORE_ERR_INVALID;
  }
  SCORE_CUDA_CHECK(cudaSetDevice(h->device));
  const int d = P.d;
  const int Ei = h->edge_off[inst + 1] - h->edge_off[inst], Ki = h->rng_off[inst + 1] - h->rng_off[inst];
  const int Li = h->lm_off[inst + 1] - h->lm_off[inst], Pi = h->pose_off[inst + 1] - h->pose_off[inst];
  const int Lpi = h->prior_off[inst + 1] - h->prior_off[inst];
  if (which == SCORE_CSR_FULL) {
    const bool q = P.relax == SCORE_RELAX_QCQP;
    const int rows = Ei * P.rpe + Ki * (q ? d : 1) + Lpi * d;
    const int nn = Ei * P.npe + Ki * (q ? 3 * d : 1) + Lpi * d;
    const int cols = Pi * P.blk + Li * d + Ki * (q ? d : 1);
    if (n_rows) *n_rows = rows;
    if (n_cols) *n_cols = cols;
    if (nnz) *nnz = nn;
    if (!indptr && !indices && !values && !weights && !rhs) return SCORE_OK;
    int *dip = nullptr, *dc = nullptr;
    double *dv = nullptr, *dw = nullptr, *db = nullptr;
    SCORE_CUDA_CHECK(cudaMalloc(&dip, sizeof(int) * (rows + 1)));
    SCORE_CUDA_CHECK(cudaMalloc(&dc, sizeof(int) * (nn + 1)));
    SCORE_CUDA_CHECK(cudaMalloc(&dv, sizeof(double) * (nn + 1)));
    SCORE_CUDA_CHECK(cudaMalloc(&dw, sizeof(double) * (rows + 1)));
    SCORE_CUDA_CHECK(cudaMalloc(&db, sizeof(double) * (rows + 1)));
    AsmOut out{dip, dc, dv, dw, db, nullptr};
    const long nf = (long)Ei + Ki + Lpi;
    k_assemble<<<grid_for(nf, 256), 256, 0, h->own_stream>>>(P, q ? ASM_FULL_QCQP : ASM_FULL_SOCP, inst, inst + 1, out);
    SCORE_CUDA_CHECK(cudaStreamSynchronize(h->own_stream));
    SCORE_CUDA_CHECK(cudaGetLastError());
    if (indptr) SCORE_CUDA_CHECK(cudaMemcpy(indptr, dip, sizeof(int) * (rows + 1), cudaMemcpyDefault));
    if (indices) SCORE_CUDA_CHECK(cudaMemcpy(indices, dc, sizeof(int) * nn, cudaMemcpyDefault));
    if (values) SCORE_CUDA_CHECK(cudaMemcpy(values, dv, sizeof(double) * nn, cudaMemcpyDefault));
    if (weights) SCORE_CUDA_CHECK(cudaMemcpy(weights, dw, sizeof(double) * rows, cudaMemcpyDefault));
    if (rhs) SCORE_CUDA_CHECK(cudaMemcpy(rhs, db, sizeof(double) * rows, cudaMemcpyDefault));
    cudaFree(dip);
    cudaFree(dc);
    cudaFree(dv);
    cudaFree(dw);
    cudaFree(db);
    return SCORE_OK;
  }
  if (which != SCORE_CSR_REDUCED && which != SCORE_CSR_REDUCED_T) {
    g_score_last_error = "unknown matrix selector";
    return SCORE_ERR_INVALID;
  }
  if (!h->solved_once) {
    g_score_last_error = "the reduced operator exists only after score_solve";
    return SCORE_ERR_STATE;
  }
  const int rows = h->roff[inst + 1] - h->roff[inst], cols = h->zoff[inst + 1] - h->zoff[inst];
  const int nn = h->nnzoff[inst + 1] - h->nnzoff[inst];
  const bool tr = which == SCORE_CSR_REDUCED_T;
  if (n_rows) *n_rows = tr ? cols : rows;
  if (n_cols) *n_cols = tr ? rows : cols;
  if (nnz) *nnz = nn;
  if (!indptr && !indices && !values && !weights && !rhs) return SCORE_OK;
  const int nr = tr ? cols : rows;
  const int r0 = tr ? h->zoff[inst] : h->roff[inst];
  const int *ip = tr ? P.t_indptr : P.indptr;
  const int *ix = tr ? P.t_rows : P.cols;
  const double *vv = tr ? P.t_vals : P.vals;
  std::vector<int> hip(nr + 1);
  SCORE_CUDA_CHECK(cudaMemcpy(hip.data(), ip + r0, sizeof(int) * (nr + 1), cudaMemcpyDefault));
  const int base = hip[0], sub = tr ? h->roff[inst] : h->zoff[inst];
  if (hip[nr] - base != nn) {
    g_score_last_error = "internal: instance nnz mismatch";
    return SCORE_ERR_STATE;
  }
  if (indptr)
    for (int i = 0; i <= nr; ++i) indptr[i] = hip[i] - base;
  if (indices) {
    SCORE_CUDA_CHECK(cudaMemcpy(indices, ix + base, sizeof(int) * nn, cudaMemcpyDefault));
    for (int i = 0; i < nn; ++i) indices[i] -= sub;
  }
  if (values) SCORE_CUDA_CHECK(cudaMemcpy(values, vv + base, sizeof(double) * nn, cudaMemcpyDefault));
  if (!tr) {
    if (weights) SCORE_CUDA_CHECK(cudaMemcpy(weights, P.w + r0, sizeof(double) * rows, cudaMemcpyDefault));
    if (rhs) SCORE_CUDA_CHECK(cudaMemcpy(rhs, P.b + r0, sizeof(double) * rows, cudaMemcpyDefault));
  }
  return SCORE_OK;
}

extern "C" int score_round_so(int32_t dim, int64_t n, const double *mats, double *out, int32_t device) {
  if (dim != 2 && dim != 3) {
    g_score_last_error = "Value " + std::to_string(dim) + " is not 2 or 3";
    return SCORE_ERR_INVALID;
  }
  if (n < 0 || (n > 0 && (!mats || !out))) {
    g_score_last_error = "null argument";
    return SCORE_ERR_INVALID;
  }
  if (n == 0) return SCORE_OK;
  SCORE_CUDA_CHECK(cudaSetDevice(device));
  double *din = nullptr, *dout = nullptr;
  const size_t bytes = sizeof(double) * (size_t)n * dim * dim;
  SCORE_CUDA_CHECK(cudaMalloc(&din, bytes));
  SCORE_CUDA_CHECK(cudaMalloc(&dout, bytes));
  SCORE_CUDA_CHECK(cudaMemcpy(din, mats, bytes, cudaMemcpyDefault));
  k_round_so<<<grid_for(n, 128), 128>>>(dim, n, din, dim * dim, dim, dout);
  SCORE_CUDA_CHECK(cudaGetLastError());
  SCORE_CUDA_CHECK(cudaMemcpy(out, dout, bytes, cudaMemcpyDefault));
  cudaFree(din);
  cudaFree(dout);
  return SCORE_OK;
}
