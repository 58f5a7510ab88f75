// libscore_b200 — C ABI (include/score_b200.h) over the sm_100a kernels.
#include <dlfcn.h>
#include <nccl.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <map>
#include <mutex>
#include <thread>
#include <vector>

#include "assemble.cuh"
#include "coarse.cuh"
#include "common.cuh"
#include "dense.cuh"
#include "evaluate.cuh"
#include "extract.cuh"
#include "fused.cuh"
#include "hessvec.cuh"
#include "precond.cuh"
#include "refine.cuh"
#include "solver.cuh"

#ifndef SCORE_EVENT_SYNC_FLAG
#define SCORE_EVENT_SYNC_FLAG cudaEventBlockingSync
#endif

thread_local std::string g_score_last_error;

// NCCL is bound at run time, and only when a row-partitioned solve asks for it: the library then shares the
// process's NCCL (torch's, when torch.distributed brought one) and has no load-time dependency on it.
namespace {
struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  bool ok = false;
};
NcclApi g_nccl;
bool nccl_load() {
  if (g_nccl.ok) return true;
  void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) {
    g_score_last_error = std::string("cannot load NCCL: ") + dlerror();
    return false;
  }
  g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(lib, "ncclGetUniqueId");
  g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(lib, "ncclCommInitRank");
  g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(lib, "ncclAllReduce");
  g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(lib, "ncclCommDestroy");
  g_nccl.ok = g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.AllReduce && g_nccl.CommDestroy;
  if (!g_nccl.ok) g_score_last_error = "NCCL library lacks a required symbol";
  return g_nccl.ok;
}
}  // namespace

using namespace score;

struct ScoreHandle_ {
  int device = 0;
  DevProblem P{};
  SolverVecs V{};
  BlockTables T{};
  WorkLists W{};
  int *wl_mem = nullptr;  // backing store of the work lists
  int n_sm = 148;
  int n_clusters = 16;  // co-resident clusters of the fused PCG kernel
  int *bar_mem = nullptr;  // group barriers of the fused PCG kernel (counter, generation per group)
  InstState *st = nullptr;
  int *d_ndone = nullptr;
  int *h_ndone = nullptr;  // pinned, two slots (double-buffered completion count)
  cudaEvent_t ev_done[2] = {nullptr, nullptr};
  std::map<int, cudaGraphExec_t> graphs;  // cycle length (PCG ticks) -> instantiated cycle graph
  std::map<int, int> graph_kernels;       // cycle length -> kernels in that graph
  double *wsum = nullptr;
  int *nnz_row = nullptr;
  // transpose scratch
  int *sort_keys = nullptr, *sort_idx = nullptr;                       // incidence-list sort
  int *csr_keys = nullptr, *csr_idx = nullptr, *csr_perm = nullptr;    // transpose sort (allocated with the CSR pair)
  size_t sort_cap = 0;  // entries of sort_keys / sort_idx
  void *sort_tmp = nullptr;
  size_t sort_tmp_bytes = 0;
  // outputs
  double *out_poses = nullptr, *out_lms = nullptr, *out_round = nullptr, *out_dist = nullptr;
  // host copies of the offset tables
  std::vector<int> pose_off, lm_off, edge_off, rng_off, prior_off, zoff, roff, nnzoff, seg_begin;
  std::vector<int> rb_begin, cb_begin, pb_begin;
  int n_pbd = 0;  // entries of the dense Hessian-vector item table
  RefVecs R{};    // local refinement (refine.cuh): allocated by the first score_refine
  bool ref_alloc = false, refined = false;
  int *ref_ndone = nullptr;
  void *inc_tmp = nullptr;  // radix-sort scratch of the incidence lists
  uint4 *inc_in = nullptr;  // unsorted incidence records
  size_t inc_tmp_bytes = 0;
  std::vector<int> c_off, c_moff, c_n, c_nb;
  int c_nmax = 0;
  // row-partitioned multi-GPU solve of one instance (score_comm_init)
  ncclComm_t comm = nullptr;
  int n_ranks = 1, rank = 0;
  double *red_send = nullptr, *red_recv = nullptr;  // [part_row n_rb | part_upd 4 n_rb | h nz]
  double *ls_recv = nullptr, *mk_recv = nullptr;
  size_t red_count = 0;
  SolverVecs Vg{};  // V with the reduced (global) partial sums / curvature blocks / h
  // instances with kCoarseMax < nc <= kCoarseBigMax: dense global-memory coarse level (dense.cuh); a handle that is
  // one such instance (c_big) launches its cycles directly and lets the host decide when the level is rebuilt
  std::vector<int> big;
  // per entry of `big`: landmark block eliminated by its Schur complement (dense.cuh) when no range joins two landmarks
  struct BigInst {
    bool schur = false;
    int ldp = 0;
    double *S21 = nullptr, *Y = nullptr, *Dinv = nullptr, *part = nullptr, *w = nullptr;
  };
  std::vector<BigInst> bigx;
  std::vector<char> c_ll;  // per instance: some range joins two landmarks
  bool c_big = false;
  double *cb_work = nullptr, *cb_W = nullptr, *cb_raw = nullptr, *cb_scl = nullptr;  // sweep work matrix / panels
  int cb_ldp = 0;
  std::vector<std::pair<void *, size_t>> allocs;  // arena chunks (ResourceCache)
  char *arena_ptr = nullptr;
  size_t arena_left = 0, arena_hint = 1u << 20;
  cudaStream_t own_stream = nullptr;
  cudaStream_t hi_stream = nullptr;   // high-priority twin of own_stream (sparse last cycles)
  cudaEvent_t ev_switch = nullptr;
  cudaStream_t last_stream = nullptr;  // stream of the latest score_solve (a caller's stream, or own_stream)
  SolverCfg graph_cfg{};
  bool solved_once = false;
  bool csr_valid = false;  // the reduced CSR pair is assembled (not the case after a matrix-free solve)
  int dist_per = 0;
};

constexpr int kMaxFusedGroups = 64;
constexpr int kNumKernels = 13;
constexpr int kStatSlots = 16;  // ScoreStats per-kernel arrays
enum KernelId { KI_ROWPASS = 0, KI_LINESEARCH, KI_CTRL_A, KI_ROWUPDATE, KI_COARSE_BUILD, KI_COLPASS, KI_PRECOND_REV,
                KI_COARSE_APPLY, KI_PRECOND_FWD, KI_CTRL_B, KI_PUPDATE, KI_PCG_FUSED, KI_HESSVEC };

namespace {

// Device memory of a handle comes in large chunks that are sub-allocated linearly; released chunks, streams and
// events go to process-wide caches instead of back to the driver.  In a sweep (handles of similar problems created
// and destroyed continuously, while other handles are solving) score_create / score_destroy then make no
// allocation, stream or event call at all: under concurrent graph launches every such call queues behind the
// launches for the driver's context lock (measured ~1 ms each), and growing the memory pool stalls the whole device.
struct ResourceCache {
  std::mutex mu;
  std::vector<std::pair<void *, size_t>> chunks[16];  // per device
  std::vector<cudaStream_t> streams[16];
  std::vector<cudaEvent_t> events[16];

  // best fit among the cached chunks, else a new allocation (the only path that touches the driver)
  cudaError_t get_chunk(int dev, size_t need, void **ptr, size_t *size) {
    {
      std::lock_guard<std::mutex> lk(mu);
      auto &c = chunks[dev & 15];
      int best = -1;
      for (int i = 0; i < (int)c.size(); ++i)
        if (c[i].second >= need && (best < 0 || c[i].second < c[best].second)) best = i;
      if (best >= 0 && c[best].second <= 2 * need + (1u << 20)) {
        *ptr = c[best].first;
        *size = c[best].second;
        cached_bytes[dev & 15] -= c[best].second;
        c.erase(c.begin() + best);
        return cudaSuccess;
      }
    }
    *size = need;
    cudaMemPool_t mp = nullptr;
    cudaError_t e = pool(dev, &mp);
    if (e != cudaSuccess) return e;
    e = cudaMallocFromPoolAsync(ptr, *size, mp, (cudaStream_t)0);
    if (e == cudaErrorMemoryAllocation) {
      // out of memory while chunks of other sizes sit unused in the cache: give those back to the driver and retry
      cudaGetLastError();
      {
        std::lock_guard<std::mutex> lk(mu);
        for (auto &c : chunks[dev & 15]) cudaFreeAsync(c.first, (cudaStream_t)0);
        chunks[dev & 15].clear();
        cached_bytes[dev & 15] = 0;
      }
      cudaStreamSynchronize((cudaStream_t)0);
      cudaMemPoolTrimTo(mp, 0);
      e = cudaMallocFromPoolAsync(ptr, *size, mp, (cudaStream_t)0);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)0);  // usable on any stream from here on
    return e;
  }
  // Unused chunks are kept up to kCacheCap bytes per device; beyond that the largest ones go back to the driver
  // (sweeps over varying problem sizes would otherwise pile up chunks no later handle fits).
  void put_chunk(int dev, void *ptr, size_t size) {
    std::vector<std::pair<void *, size_t>> evict;
    {
      std::lock_guard<std::mutex> lk(mu);
      auto &c = chunks[dev & 15];
      c.emplace_back(ptr, size);
      cached_bytes[dev & 15] += size;
      while (cached_bytes[dev & 15] > cache_cap() && !c.empty()) {
        int big = 0;
        for (int i = 1; i < (int)c.size(); ++i)
          if (c[i].second > c[big].second) big = i;
        evict.push_back(c[big]);
        cached_bytes[dev & 15] -= c[big].second;
        c.erase(c.begin() + big);
      }
    }
    if (!evict.empty()) {
      for (auto &c : evict) cudaFreeAsync(c.first, (cudaStream_t)0);
      cudaMemPool_t mp = nullptr;
      if (pool(dev, &mp) == cudaSuccess) {
        cudaStreamSynchronize((cudaStream_t)0);
        cudaMemPoolTrimTo(mp, 0);
      }
    }
  }
  static size_t cache_cap() {
    static const size_t cap = getenv("SCORE_CACHE_CAP_MB") ? (size_t)atoll(getenv("SCORE_CACHE_CAP_MB")) << 20 : (size_t)48 << 30;
    return cap;
  }
  // The library's own memory pool (the process's default pool is left alone): freed blocks stay in it, so that
  // create/destroy cycles of similar problems cost no driver allocation; release_all / eviction trim it.
  cudaMemPool_t pools[16] = {};
  size_t cached_bytes[16] = {};
  cudaError_t pool(int dev, cudaMemPool_t *out) {
    std::lock_guard<std::mutex> lk(mu);
    if (!pools[dev & 15]) {
      cudaMemPoolProps props{};
      props.allocType = cudaMemAllocationTypePinned;
      props.handleTypes = cudaMemHandleTypeNone;
      props.location.type = cudaMemLocationTypeDevice;
      props.location.id = dev;
      cudaError_t e = cudaMemPoolCreate(&pools[dev & 15], &props);
      if (e != cudaSuccess) {
        pools[dev & 15] = nullptr;
        return e;
      }
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pools[dev & 15], cudaMemPoolAttrReleaseThreshold, &keep);
    }
    *out = pools[dev & 15];
    return cudaSuccess;
  }
  cudaError_t get_stream(int dev, cudaStream_t *st) {
    {
      std::lock_guard<std::mutex> lk(mu);
      auto &v = streams[dev & 15];
      if (!v.empty()) {
        *st = v.back();
        v.pop_back();
        return cudaSuccess;
      }
    }
    return cudaStreamCreateWithFlags(st, cudaStreamNonBlocking);
  }
  void put_stream(int dev, cudaStream_t st) {
    std::lock_guard<std::mutex> lk(mu);
    streams[dev & 15].push_back(st);
  }
  // high-priority streams (the sparse last cycles of a batch, see score_solve)
  std::vector<cudaStream_t> hi_streams[16];
  cudaError_t get_hi_stream(int dev, cudaStream_t *st) {
    {
      std::lock_guard<std::mutex> lk(mu);
      auto &v = hi_streams[dev & 15];
      if (!v.empty()) {
        *st = v.back();
        v.pop_back();
        return cudaSuccess;
      }
    }
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);  // (numerically lower = higher priority)
    return cudaStreamCreateWithPriority(st, cudaStreamNonBlocking, hi);
  }
  void put_hi_stream(int dev, cudaStream_t st) {
    std::lock_guard<std::mutex> lk(mu);
    hi_streams[dev & 15].push_back(st);
  }
  cudaError_t get_event(int dev, cudaEvent_t *ev) {
    {
      std::lock_guard<std::mutex> lk(mu);
      auto &v = events[dev & 15];
      if (!v.empty()) {
        *ev = v.back();
        v.pop_back();
        return cudaSuccess;
      }
    }
    // blocking sync: a host thread waiting for a cycle to finish sleeps instead of spinning on a core (a sweep keeps
    // several solves in flight per GPU, 8 ranks per host: spinning waiters starve the threads that upload and launch)
    return cudaEventCreateWithFlags(ev, cudaEventDisableTiming | SCORE_EVENT_SYNC_FLAG);
  }
  void put_event(int dev, cudaEvent_t ev) {
    std::lock_guard<std::mutex> lk(mu);
    events[dev & 15].push_back(ev);
  }
  // events with timing (phase times of score_solve, kernel profile)
  std::vector<cudaEvent_t> tevents[16];
  cudaError_t get_tevent(int dev, cudaEvent_t *ev) {
    {
      std::lock_guard<std::mutex> lk(mu);
      auto &v = tevents[dev & 15];
      if (!v.empty()) {
        *ev = v.back();
        v.pop_back();
        return cudaSuccess;
      }
    }
    return cudaEventCreate(ev);
  }
  void put_tevent(int dev, cudaEvent_t ev) {
    std::lock_guard<std::mutex> lk(mu);
    tevents[dev & 15].push_back(ev);
  }
  // Instantiated cycle graphs of destroyed handles.  cudaGraphExecDestroy was seen to block for ~0.2 s now and then
  // and to block every other thread's CUDA calls with it (e2e timelines, profiles/), so a sweep does not call it per
  // handle: the executables are parked here and destroyed in bulk once many have piled up, or on release_all.
  std::vector<cudaGraphExec_t> retired;
  void retire_graph(cudaGraphExec_t g) {
    std::vector<cudaGraphExec_t> victims;
    {
      std::lock_guard<std::mutex> lk(mu);
      retired.push_back(g);
      if (retired.size() >= 512) victims.swap(retired);
    }
    for (auto v : victims) cudaGraphExecDestroy(v);
  }
  // give everything back to the driver (score_release_cached)
  void release_all() {
    std::lock_guard<std::mutex> lk(mu);
    for (auto g : retired) cudaGraphExecDestroy(g);
    retired.clear();
    int cur = 0;
    cudaGetDevice(&cur);
    for (int d = 0; d < 16; ++d) {
      if (chunks[d].empty() && streams[d].empty() && events[d].empty() && tevents[d].empty() && !pools[d]) continue;
      cudaSetDevice(d);
      for (auto &c : chunks[d]) cudaFreeAsync(c.first, (cudaStream_t)0);
      for (auto st : streams[d]) cudaStreamDestroy(st);
      for (auto st : hi_streams[d]) cudaStreamDestroy(st);
      hi_streams[d].clear();
      for (auto ev : events[d]) cudaEventDestroy(ev);
      for (auto ev : tevents[d]) cudaEventDestroy(ev);
      chunks[d].clear();
      streams[d].clear();
      events[d].clear();
      tevents[d].clear();
      cached_bytes[d] = 0;
      if (pools[d]) {  // the frees above are stream-ordered: wait for them, then hand the memory back to the driver
        cudaStreamSynchronize((cudaStream_t)0);
        cudaMemPoolTrimTo(pools[d], 0);
      }
    }
    cudaSetDevice(cur);
  }
};
ResourceCache g_cache;

template <typename T>
int dalloc(ScoreHandle_ *h, T **ptr, size_t n) {
  *ptr = nullptr;
  if (n == 0) n = 1;
  const size_t bytes = (n * sizeof(T) + 255) & ~(size_t)255;
  if (h->arena_left < bytes) {
    void *chunk = nullptr;
    size_t size = 0;
    // first chunk: the handle's estimated footprint (create_impl); further chunks a quarter of it
    const size_t want = std::max(bytes, h->allocs.empty() ? h->arena_hint : h->arena_hint / 4);
    cudaError_t e = g_cache.get_chunk(h->device, want, &chunk, &size);
    if (e != cudaSuccess) {
      g_score_last_error = std::string("cudaMalloc failed: ") + cudaGetErrorString(e);
      return SCORE_ERR_ALLOC;
    }
    h->allocs.emplace_back(chunk, size);
    h->arena_ptr = (char *)chunk;
    h->arena_left = size;
  }
  *ptr = (T *)h->arena_ptr;
  h->arena_ptr += bytes;
  h->arena_left -= bytes;
  return SCORE_OK;
}

template <typename T>
int upload(ScoreHandle_ *h, T **dst, const T *src, size_t n) {
  int rc = dalloc(h, dst, n);
  if (rc) return rc;
  if (n && src) SCORE_CUDA_CHECK(cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyDefault));
  return SCORE_OK;
}

// Copy a caller array (host or device pointer) into host memory.  Host sources — the common case — are copied
// with memcpy: a cudaMemcpy between two host buffers still goes through the driver and synchronises with the
// default stream, which in a pipelined sweep means queueing behind other threads' launches.
cudaError_t fetch_to_host(void *dst, const void *src, size_t bytes) {
  if (bytes == 0) return cudaSuccess;
  cudaPointerAttributes at{};
  const cudaError_t e = cudaPointerGetAttributes(&at, src);
  if (e != cudaSuccess) {
    cudaGetLastError();  // plain (unregistered) host memory on old drivers
    memcpy(dst, src, bytes);
    return cudaSuccess;
  }
  if (at.type == cudaMemoryTypeDevice) return cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost);
  if (at.type == cudaMemoryTypeManaged) return cudaMemcpy(dst, src, bytes, cudaMemcpyDefault);
  memcpy(dst, src, bytes);
  return cudaSuccess;
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
  SCORE_CUDA_CHECK(fetch_to_host(dst.data(), src, sizeof(int) * (n_inst + 1)));
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

// Grid of a work-list kernel: at most `per_sm` CTAs per SM (the CTAs loop over the listed work items).
int wgrid(const ScoreHandle_ *h, long items, int per_sm) {
  static const int mult = getenv("SCORE_WGRID_MULT") ? atoi(getenv("SCORE_WGRID_MULT")) : 2;  // tuning knob; 0: one item per CTA
  if (mult <= 0) return (int)std::max(1l, items);
  return (int)std::max(1l, std::min<long>(items, (long)h->n_sm * per_sm * mult));
}

// Algorithmic bytes one instance moves through tick kernel `k` in tick mode `mode` (fp64 values, int32
// indices, every array read or written once; DESIGN.md "algorithmic bytes").  Control kernels read a few
// partial sums and are counted as zero.
struct InstDims {
  double d, nnz, m, nz, K, P, nc, E = 0, L = 0, Lp = 0;
  bool fused = true;  // coarse application inside k_precond_fwd
  bool mf = false;    // PCG iterations apply the operator matrix-free (k_hessvec)
};
double kernel_bytes_inst(int k, int mode, const InstDims &D) {
  const double d = D.d, blk = d * (d + 1), d1 = d + 1, nnz = D.nnz, m = D.m, nz = D.nz, K = D.K, Pn = D.P, nc = D.nc;
  const bool cg = mode == TM_CG, ls = mode == TM_LS, ev = mode == TM_EVAL;
  switch (k) {
    case KI_HESSVEC: {  // x in, h out, incidence lists, every relative-pose factor once (measurement, 2 precisions, 2 pose
                        // indices), every range term twice (partner index + curvature block)
      if (!cg || !D.mf) return 0.0;
      // x in, h out, list bounds, odometry links, 16-byte incidence records of the ranges / priors, every relative-pose
      // factor once (measurement, 2 precisions), every range term twice (curvature block; the partner gather is in x)
      return 16.0 * nz + 4.0 * (Pn + D.L + 1) + 4.0 * Pn + 16.0 * (2.0 * K + D.Lp) + D.E * 8.0 * (d + d * d + 2) +
             2.0 * K * 8.0 * d * (d + 1) / 2;
    }
    case KI_ROWPASS:  // B (vals+cols+indptr), gather x; CG: w of the plain rows, u out, M_k of the ranges; LS: bdz out
      if (cg && D.mf) return 0.0;
      if (ls && D.mf)  // k_rows_mf: x gathered, every factor once (measurement + 2 pose indices / 2 owner indices), bdz out
        return 8.0 * nz + D.E * (8.0 * (d + d * d) + 8.0) + 8.0 * K + 8.0 * m;
      if (cg) return 12.0 * nnz + 4.0 * (m + 1) + 8.0 * nz + 8.0 * (m - d * K) + 8.0 * m + 8.0 * K * d * (d + 1) / 2;
      if (ls) return 12.0 * nnz + 4.0 * (m + 1) + 8.0 * nz + 8.0 * m;
      return 0.0;
    case KI_LINESEARCH:  // res, bdz, w of the plain rows, (dist, w) of the ranges
      return ls ? 16.0 * m + 8.0 * (m - d * K) + 16.0 * K : 0.0;
    case KI_ROWUPDATE:  // LS: res rw, bdz, w, u out, dist, M_k out; EVAL: res, w, u out, dist
      if (ls) return 40.0 * m + 8.0 * K + 8.0 * K * d * (d + 1) / 2;
      if (ev) return 24.0 * m + 8.0 * K;
      return 0.0;
    case KI_COARSE_BUILD:  // per incidence (2 K): code, 2w, frame, M_k; per pair entry (K): code, 2 frames, M_k; inverse out
      return (ls && nc > 0) ? 2.0 * K * (12.0 + 8.0 * d1 + 4.0 * d * d1) + K * (4.0 + 16.0 * d1 + 4.0 * d * d1) + 8.0 * nc * nc
                            : 0.0;
    case KI_COLPASS:  // B^T, gather u; CG: dz rw, p, r rw; LS: z rw, dz rw, r out; EVAL: z
      if (cg && D.mf) return 48.0 * nz;  // h, p in; dz, r read + written
      if (D.mf)  // k_grad_mf: u in, incidence lists, measurements (base-pose role), then z / dz / r (LS) or z (EVAL)
        return 8.0 * m + 4.0 * (Pn + D.L + 1) + 4.0 * Pn + 16.0 * (2.0 * K + D.Lp) + D.E * 8.0 * (d + d * d) +
               (ls ? 40.0 : 8.0) * nz;
      if (cg || ls) return 12.0 * nnz + 4.0 * (nz + 1) + 8.0 * m + 40.0 * nz;
      return 12.0 * nnz + 4.0 * (nz + 1) + 8.0 * m + 8.0 * nz;
    case KI_PRECOND_REV:  // r, ytmp out, G, M (symmetric: upper triangle)
      return ev ? 0.0 : 16.0 * nz + 8.0 * Pn * (blk + d1 * (d1 + 1) / 2);
    case KI_COARSE_APPLY:  // inverse, rhs, solution out, scatter
      return (ev || D.fused) ? 0.0 : 8.0 * (nc * nc + 3.0 * nc);
    case KI_PRECOND_FWD:  // ytmp, r, s out, G (+ fused coarse application: inverse, rhs, scatter)
      return ev ? 0.0 : 24.0 * nz + 8.0 * Pn * blk + (D.fused ? 8.0 * (nc * nc + 2.0 * nc) : 0.0);
    case KI_PUPDATE:  // s, p rw
      return ev ? 0.0 : 24.0 * nz;
    default:
      return 0.0;
  }
}

}  // namespace

extern "C" const char *score_last_error(void) { return g_score_last_error.c_str(); }
extern "C" void score_release_cached(void) { g_cache.release_all(); }
extern "C" const char *score_version(void) { return "score_b200 0.1.0 (sm_100a)"; }

// Device scratch of the stand-alone entry points: freed on every exit path.
namespace {
struct DevBuf {
  void *p = nullptr;
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
  template <typename T>
  T *as() const {
    return (T *)p;
  }
  ~DevBuf() {
    if (p) cudaFree(p);
  }
  DevBuf() = default;
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
};
}  // namespace

// ---- evaluation after the path: SE(d)-aligned absolute trajectory error, batched (evaluate.cuh) --------------
namespace {
// est / gt / off are device pointers; outputs are host pointers (any may be null)
int ate_launch(int dim, int n_traj, const int *d_off, const double *d_est, long es, int ecs, int ec0, const double *d_gt,
               int align, double *rmse, double *R, double *t, cudaStream_t st) {
  DevBuf out;
  const size_t per = 1 + (size_t)dim * dim + dim;
  SCORE_CUDA_CHECK(out.alloc(sizeof(double) * per * n_traj));
  double *d_out = out.as<double>();
  double *d_rmse = d_out, *d_R = d_out + n_traj, *d_t = d_R + (size_t)n_traj * dim * dim;
  const int grid = std::min(n_traj, 148 * 8);
  if (dim == 2)
    k_ate<2><<<grid, kAteThreads, 0, st>>>(n_traj, d_off, d_est, es, ecs, ec0, d_gt, align, d_rmse, d_R, d_t);
  else
    k_ate<3><<<grid, kAteThreads, 0, st>>>(n_traj, d_off, d_est, es, ecs, ec0, d_gt, align, d_rmse, d_R, d_t);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e == cudaSuccess && rmse) e = cudaMemcpy(rmse, d_rmse, sizeof(double) * n_traj, cudaMemcpyDefault);
  if (e == cudaSuccess && R) e = cudaMemcpy(R, d_R, sizeof(double) * (size_t)n_traj * dim * dim, cudaMemcpyDefault);
  if (e == cudaSuccess && t) e = cudaMemcpy(t, d_t, sizeof(double) * (size_t)n_traj * dim, cudaMemcpyDefault);
  SCORE_CUDA_CHECK(e);
  return SCORE_OK;
}
int ate_check_offsets(int n_traj, const int32_t *off, int64_t n_points) {
  if (n_traj < 0 || (n_traj > 0 && !off)) {
    g_score_last_error = "null trajectory offsets";
    return SCORE_ERR_INVALID;
  }
  for (int i = 0; i < n_traj; ++i)
    if (off[i] < 0 || off[i + 1] < off[i] || (n_points >= 0 && off[i + 1] > n_points)) {
      g_score_last_error = "trajectory offsets must be non-decreasing and within the point array";
      return SCORE_ERR_INVALID;
    }
  return SCORE_OK;
}
}  // namespace

extern "C" int score_trajectory_ate(int32_t dim, int32_t n_traj, const int32_t *traj_off, const double *est,
                                    const double *gt, int32_t align, double *rmse, double *R, double *t, int32_t device) {
  if (dim != 2 && dim != 3) {
    g_score_last_error = "Value " + std::to_string(dim) + " is not 2 or 3";
    return SCORE_ERR_INVALID;
  }
  int rc = ate_check_offsets(n_traj, traj_off, -1);
  if (rc) return rc;
  if (n_traj == 0) return SCORE_OK;
  const size_t n = (size_t)traj_off[n_traj];
  if (n > 0 && (!est || !gt)) {
    g_score_last_error = "null argument";
    return SCORE_ERR_INVALID;
  }
  SCORE_CUDA_CHECK(cudaSetDevice(device));
  DevBuf off_buf, pts_buf;  // [est | gt]
  SCORE_CUDA_CHECK(off_buf.alloc(sizeof(int) * (n_traj + 1)));
  SCORE_CUDA_CHECK(pts_buf.alloc(sizeof(double) * std::max<size_t>(1, 2 * n * dim)));
  int *d_off = off_buf.as<int>();
  double *d_pts = pts_buf.as<double>();
  cudaError_t e = cudaMemcpy(d_off, traj_off, sizeof(int) * (n_traj + 1), cudaMemcpyDefault);
  if (e == cudaSuccess && n) e = cudaMemcpy(d_pts, est, sizeof(double) * n * dim, cudaMemcpyDefault);
  if (e == cudaSuccess && n) e = cudaMemcpy(d_pts + n * dim, gt, sizeof(double) * n * dim, cudaMemcpyDefault);
  rc = SCORE_OK;
  if (e == cudaSuccess)
    rc = ate_launch(dim, n_traj, d_off, d_pts, dim, 1, 0, d_pts + n * dim, align, rmse, R, t, nullptr);
  SCORE_CUDA_CHECK(e);
  return rc;
}

extern "C" int score_eval_ate(ScoreHandle h, int32_t n_traj, const int32_t *traj_off, const double *gt_pos, int32_t align,
                              double *rmse, double *R, double *t) {
  if (!h) {
    g_score_last_error = "null handle";
    return SCORE_ERR_INVALID;
  }
  if (!h->solved_once) {
    g_score_last_error = "score_eval_ate called before score_solve";
    return SCORE_ERR_STATE;
  }
  const DevProblem &P = h->P;
  // default: one trajectory per instance
  const int32_t *off = traj_off ? traj_off : h->pose_off.data();
  if (!traj_off) n_traj = P.n_inst;
  int rc = ate_check_offsets(n_traj, off, P.P);
  if (rc) return rc;
  if (n_traj == 0) return SCORE_OK;
  if (!gt_pos) {
    g_score_last_error = "null ground truth";
    return SCORE_ERR_INVALID;
  }
  SCORE_CUDA_CHECK(cudaSetDevice(h->device));
  DevBuf off_buf, gt_buf;
  SCORE_CUDA_CHECK(off_buf.alloc(sizeof(int) * (n_traj + 1)));
  SCORE_CUDA_CHECK(gt_buf.alloc(sizeof(double) * std::max<size_t>(1, (size_t)P.P * P.d)));
  int *d_off = off_buf.as<int>();
  double *d_gt = gt_buf.as<double>();
  cudaError_t e = cudaMemcpy(d_off, off, sizeof(int) * (n_traj + 1), cudaMemcpyDefault);
  if (e == cudaSuccess) e = cudaMemcpy(d_gt, gt_pos, sizeof(double) * (size_t)P.P * P.d, cudaMemcpyDefault);
  rc = SCORE_OK;
  if (e == cudaSuccess)
    rc = ate_launch(P.d, n_traj, d_off, h->out_poses, P.blk, P.d + 1, P.d, d_gt, align, rmse, R, t, nullptr);
  SCORE_CUDA_CHECK(e);
  return rc;
}

// ---- local refinement (refine.cuh)
template <int D>
static int refine_run(ScoreHandle_ *h, const RefCfg &cfg, cudaStream_t st, int *launches, int *outer_done) {
  const DevProblem &P = h->P;
  const RefVecs &R = h->R;
  const BlockTables &T = h->T;
  const int nb = T.n_pb, ni = P.n_inst, cgrid = (ni + 3) / 4, sgrid = (P.n_seg + 63) / 64;
  const int chain = cfg.chain ? 1 : 0;
  const int *segb = chain ? P.seg_begin : nullptr;
  int n = 0;
  k_ref_visit<D, RM_COST><<<nb, kThreads, 0, st>>>(P, R, T, 0);
  k_ref_ctrl<RC_COST0><<<cgrid, 128, 0, st>>>(T, R, cfg, ni, h->ref_ndone, segb);
  n += 2;
  int done = 0, outer = 0;
  for (; outer < cfg.max_outer && done < ni; ++outer) {
    k_ref_visit<D, RM_LIN><<<nb, kThreads, 0, st>>>(P, R, T, 0);
    k_ref_vec<D, RV_START><<<nb, kThreads, 0, st>>>(P, R, T, chain);
    SCORE_CUDA_CHECK(cudaMemsetAsync(h->ref_ndone + 1, 0, sizeof(int), st));
    n += 2;
    if (chain) {
      k_ref_chain_factor<D><<<sgrid, 64, 0, st>>>(P, R);
      k_ref_chain_apply<D><<<sgrid, 64, 0, st>>>(P, R, 1);
      n += 2;
    }
    k_ref_ctrl<RC_START><<<cgrid, 128, 0, st>>>(T, R, cfg, ni, h->ref_ndone, segb);
    n += 1;
    if (chain) {
      k_ref_vec<D, RV_PUPDATE><<<nb, kThreads, 0, st>>>(P, R, T, chain);  // p = s (beta = 0)
      n += 1;
    }
    int cg_ended = 0;
    for (int it = 0; it < cfg.max_inner; ++it) {
      if (it > 0 && (it & 7) == 0) {  // every 8 PCG iterations: have all instances' solves ended?
        SCORE_CUDA_CHECK(cudaMemcpyAsync(&cg_ended, h->ref_ndone + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
        SCORE_CUDA_CHECK(cudaStreamSynchronize(st));
        if (cg_ended + done >= ni) break;
      }
      k_ref_visit<D, RM_HV><<<nb, kThreads, 0, st>>>(P, R, T, 0);
      k_ref_ctrl<RC_ALPHA><<<cgrid, 128, 0, st>>>(T, R, cfg, ni, h->ref_ndone, segb);
      k_ref_vec<D, RV_UPDATE><<<nb, kThreads, 0, st>>>(P, R, T, chain);
      if (chain) k_ref_chain_apply<D><<<sgrid, 64, 0, st>>>(P, R, 0);
      k_ref_ctrl<RC_BETA><<<cgrid, 128, 0, st>>>(T, R, cfg, ni, h->ref_ndone, segb);
      k_ref_vec<D, RV_PUPDATE><<<nb, kThreads, 0, st>>>(P, R, T, chain);
      n += 5 + chain;
    }
    k_ref_vec<D, RV_TRIAL><<<nb, kThreads, 0, st>>>(P, R, T, chain);
    k_ref_visit<D, RM_COST><<<nb, kThreads, 0, st>>>(P, R, T, 1);
    k_ref_ctrl<RC_ACCEPT><<<cgrid, 128, 0, st>>>(T, R, cfg, ni, h->ref_ndone, segb);
    k_ref_commit<D><<<nb, kThreads, 0, st>>>(P, R, T);
    k_ref_clear<<<grid_for(ni, 128), 128, 0, st>>>(R, ni);
    n += 5;
    SCORE_CUDA_CHECK(cudaMemcpyAsync(&done, h->ref_ndone, sizeof(int), cudaMemcpyDeviceToHost, st));
    SCORE_CUDA_CHECK(cudaStreamSynchronize(st));
  }
  SCORE_CUDA_CHECK(cudaGetLastError());
  *launches = n;
  *outer_done = outer;
  return SCORE_OK;
}

extern "C" int score_refine(ScoreHandle h, const ScoreRefineParams *params, const double *init_poses,
                            const double *init_landmarks, ScoreRefineStats *stats, ScoreRefineInstanceStats *per_instance) {
  if (!h) {
    g_score_last_error = "null handle";
    return SCORE_ERR_INVALID;
  }
  if (!h->solved_once) {
    g_score_last_error = "score_refine called before score_solve (the factor incidence lists are built there)";
    return SCORE_ERR_STATE;
  }
  if ((init_poses == nullptr) != (init_landmarks == nullptr) && h->P.L > 0) {
    g_score_last_error = "score_refine: give both init_poses and init_landmarks, or neither";
    return SCORE_ERR_INVALID;
  }
  ScoreRefineParams prm{};
  if (params) prm = *params;
  RefCfg cfg;
  cfg.max_outer = prm.max_outer > 0 ? prm.max_outer : 100;
  cfg.max_inner = prm.max_inner > 0 ? prm.max_inner : 200;
  cfg.rel_tol = prm.rel_tol > 0 ? prm.rel_tol : 1e-10;
  cfg.lambda0 = prm.lambda0 > 0 ? prm.lambda0 : 1e-3;
  cfg.eta = prm.cg_tol > 0 ? prm.cg_tol : 1e-4;
  cfg.chain = prm.preconditioner != 1;  // 0 (default): odometry-chain block LDL^T; 1: block-Jacobi
  SCORE_CUDA_CHECK(cudaSetDevice(h->device));
  const DevProblem &P = h->P;
  const int d = P.d, dof = d + (d == 2 ? 1 : 3);
  int rc;
  if (!h->ref_alloc) {
    RefVecs &R = h->R;
#define RA(ptr, n) \
  if ((rc = dalloc(h, &(ptr), (size_t)(n)))) return rc;
    RA(R.x, P.nz) RA(R.xt, P.nz) RA(R.g, P.nz) RA(R.dg, P.nz) RA(R.dl, P.nz) RA(R.r, P.nz) RA(R.s, P.nz) RA(R.p, P.nz)
    RA(R.q, P.nz)
    const size_t nblk = (size_t)P.P * dof * dof + (size_t)P.L * d * d;
    RA(R.Db, nblk) RA(R.Mi, nblk) RA(R.Ob, (size_t)P.P * dof * dof) RA(R.Sinv, (size_t)P.P * dof * dof)
    RA(R.Lb, (size_t)P.P * dof * dof) RA(R.part_seg, P.n_seg) RA(R.part_cost, h->T.n_pb) RA(R.part_dot, h->T.n_pb) RA(R.part_aux, 2 * (size_t)h->T.n_pb) RA(R.st, P.n_inst)
    RA(h->ref_ndone, 2)
#undef RA
    h->ref_alloc = true;
  }
  cudaStream_t st = prm.stream ? (cudaStream_t)prm.stream : h->own_stream;
  h->last_stream = st;
  struct TimingPair {  // two pooled timing events, returned to the pool on every exit path
    cudaEvent_t e[2] = {};
    int dev = 0;
    ~TimingPair() {
      for (auto x : e)
        if (x) g_cache.put_tevent(dev, x);
    }
  } tp;
  tp.dev = h->device;
  cudaEvent_t *ev = tp.e;
  for (int i = 0; i < 2; ++i) SCORE_CUDA_CHECK(g_cache.get_tevent(h->device, &ev[i]));
  SCORE_CUDA_CHECK(cudaEventRecord(ev[0], st));
  SCORE_CUDA_CHECK(cudaMemsetAsync(h->ref_ndone, 0, 2 * sizeof(int), st));
  DevBuf ip, il;
  const double *d_poses = h->out_poses, *d_round = h->out_round, *d_lms = h->out_lms;
  if (init_poses) {
    SCORE_CUDA_CHECK(ip.alloc(sizeof(double) * std::max<size_t>(1, (size_t)P.P * P.blk)));
    SCORE_CUDA_CHECK(il.alloc(sizeof(double) * std::max<size_t>(1, (size_t)P.L * d)));
    SCORE_CUDA_CHECK(cudaMemcpyAsync(ip.as<double>(), init_poses, sizeof(double) * (size_t)P.P * P.blk, cudaMemcpyDefault, st));
    if (P.L > 0)
      SCORE_CUDA_CHECK(cudaMemcpyAsync(il.as<double>(), init_landmarks, sizeof(double) * (size_t)P.L * d, cudaMemcpyDefault, st));
    d_poses = ip.as<double>();
    d_round = nullptr;
    d_lms = il.as<double>();
  }
  k_ref_init<<<grid_for(P.nz, 256), 256, 0, st>>>(P, h->R, d_poses, d_round, d_lms, cfg.lambda0);
  int launches = 0, outer = 0;
  rc = (d == 2) ? refine_run<2>(h, cfg, st, &launches, &outer) : refine_run<3>(h, cfg, st, &launches, &outer);
  if (rc) return rc;
  SCORE_CUDA_CHECK(cudaEventRecord(ev[1], st));
  SCORE_CUDA_CHECK(cudaStreamSynchronize(st));
  std::vector<RefState> hs(P.n_inst);
  SCORE_CUDA_CHECK(cudaMemcpy(hs.data(), h->R.st, sizeof(RefState) * P.n_inst, cudaMemcpyDeviceToHost));
  int conv = 0;
  for (int i = 0; i < P.n_inst; ++i) {
    conv += hs[i].converged ? 1 : 0;
    if (per_instance) {
      per_instance[i].cost_initial = hs[i].cost0;
      per_instance[i].cost_final = hs[i].cost;
      per_instance[i].outer_iterations = hs[i].outer;
      per_instance[i].accepted_steps = hs[i].n_accept;
    }
  }
  if (stats) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev[0], ev[1]);
    stats->n_instances = P.n_inst;
    stats->n_converged = conv;
    stats->outer_iterations = outer;
    stats->kernel_launches = launches + 1;
    stats->refine_ms = ms;
  }
  h->refined = true;
  return SCORE_OK;
}

extern "C" int score_get_refined(ScoreHandle h, double *poses, double *landmarks) {
  if (!h) {
    g_score_last_error = "null handle";
    return SCORE_ERR_INVALID;
  }
  if (!h->refined) {
    g_score_last_error = "score_get_refined called before score_refine";
    return SCORE_ERR_STATE;
  }
  SCORE_CUDA_CHECK(cudaSetDevice(h->device));
  const DevProblem &P = h->P;
  // the state is in column-space layout: instance by instance, pose blocks then landmarks
  std::vector<double> x(P.nz);
  SCORE_CUDA_CHECK(cudaMemcpy(x.data(), h->R.x, sizeof(double) * P.nz, cudaMemcpyDeviceToHost));
  for (int i = 0; i < P.n_inst; ++i) {
    const size_t z0 = h->zoff[i], Pi = h->pose_off[i + 1] - h->pose_off[i], Li = h->lm_off[i + 1] - h->lm_off[i];
    if (poses) memcpy(poses + (size_t)h->pose_off[i] * P.blk, x.data() + z0, sizeof(double) * Pi * P.blk);
    if (landmarks && Li) memcpy(landmarks + (size_t)h->lm_off[i] * P.d, x.data() + z0 + Pi * P.blk, sizeof(double) * Li * P.d);
  }
  return SCORE_OK;
}

// Pinned host slots for the per-handle completion counters.  cudaMallocHost / cudaFreeHost are device-wide
// synchronisation points: a handle created or destroyed while another handle's solve is running would stall behind
// it (and stall it), so the slots come from one slab that is pinned once per process.
namespace {
struct PinnedSlots {
  std::mutex mu;
  int *slab = nullptr;
  std::vector<int> free_list;
  static constexpr int kSlots = 256, kInts = 2;
  int *get() {
    std::lock_guard<std::mutex> lk(mu);
    if (!slab) {
      if (cudaMallocHost((void **)&slab, sizeof(int) * kSlots * kInts) != cudaSuccess) {
        slab = nullptr;
        return nullptr;
      }
      for (int i = kSlots - 1; i >= 0; --i) free_list.push_back(i);
    }
    if (free_list.empty()) return nullptr;
    const int i = free_list.back();
    free_list.pop_back();
    return slab + i * kInts;
  }
  bool put(int *p) {  // false: not one of ours
    std::lock_guard<std::mutex> lk(mu);
    if (!slab || p < slab || p >= slab + kSlots * kInts) return false;
    free_list.push_back((int)(p - slab) / kInts);
    return true;
  }
};
PinnedSlots g_pinned;
}  // namespace

extern "C" void score_destroy(ScoreHandle h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->comm && g_nccl.ok) g_nccl.CommDestroy(h->comm);
  if (h->own_stream) cudaStreamSynchronize(h->own_stream);  // nothing of this handle is in flight any more
  if (h->last_stream && h->last_stream != h->own_stream) cudaStreamSynchronize(h->last_stream);
  for (auto &kv : h->graphs) g_cache.retire_graph(kv.second);  // destroyed later, in bulk (see ResourceCache)
  for (auto &e : h->ev_done)
    if (e) g_cache.put_event(h->device, e);
  for (auto &c : h->allocs) g_cache.put_chunk(h->device, c.first, c.second);
  if (h->h_ndone && !g_pinned.put(h->h_ndone)) cudaFreeHost(h->h_ndone);
  if (h->hi_stream) cudaStreamSynchronize(h->hi_stream);
  if (h->hi_stream) g_cache.put_hi_stream(h->device, h->hi_stream);
  if (h->ev_switch) g_cache.put_event(h->device, h->ev_switch);
  if (h->own_stream) g_cache.put_stream(h->device, h->own_stream);
  delete h;
}

// Static sorted lists for the coarse-matrix build (coarse.cuh): per instance, the incidences (range,
// endpoint) sorted by slot and the ranges sorted by (lower slot, higher slot), both by stable counting
// sorts, the run tables, and the split of the off-diagonal runs over the 32 warps of the build CTA.
// Instances are independent: they are processed by a pool of host threads and concatenated afterwards.
struct CoarseInstTables {
  std::vector<int> inc_code, drun_slot, drun_begin, pr_code, orun_lo, orun_hi, orun_begin, ws;
  std::vector<double> inc_w2;
  bool bad = false;
  bool has_ll = false;  // a range between two landmarks (the landmark block of the coarse matrix is then not block diagonal)
};

static void coarse_tables_one(const ScoreHandle_ *h, int i, const int *rng_a, const int *rng_b, const double *rng_w,
                              const int *seg_ptr, CoarseInstTables &out) {
  constexpr int NW = kCoarseThreads / 32;
  const int blk = h->P.blk;
  out.ws.assign(NW + 1, 0);
  if (h->c_n[i] <= 0) return;
  const int Pi = h->pose_off[i + 1] - h->pose_off[i], Li = h->lm_off[i + 1] - h->lm_off[i];
  const int nsegfree = h->c_nb[i] / blk, nslots = nsegfree + Li;
  const int k0 = h->rng_off[i], Ki = h->rng_off[i + 1] - k0;
  std::vector<int> pose_slot(Pi, -1), cnt, pos, sa_v(Ki), sb_v(Ki);
  for (int s = h->seg_begin[i]; s < h->seg_begin[i + 1]; ++s)
    for (int p = seg_ptr[s]; p < seg_ptr[s + 1]; ++p) pose_slot[p - h->pose_off[i]] = s - h->seg_begin[i] - 1;
  auto slot_of = [&](int owner) { return owner < Pi ? pose_slot[owner] : nsegfree + (owner - Pi); };
  // incidences by slot
  cnt.assign(nslots + 1, 0);
  for (int k = 0; k < Ki; ++k) {
    const int a = rng_a[k0 + k], b = rng_b[k0 + k];
    if (a < 0 || b < 0 || a >= Pi + Li || b >= Pi + Li) {
      out.bad = true;
      return;
    }
    const int sa = slot_of(a), sb = slot_of(b);
    if (a >= Pi && b >= Pi) out.has_ll = true;
    sa_v[k] = sa;
    sb_v[k] = sb;
    if (sa >= 0 && sa == sb) {
      cnt[sa + 1]++;
    } else {
      if (sa >= 0) cnt[sa + 1]++;
      if (sb >= 0) cnt[sb + 1]++;
    }
  }
  for (int s = 0; s < nslots; ++s) cnt[s + 1] += cnt[s];
  const int ninc = cnt[nslots];
  out.inc_code.resize(ninc);
  out.inc_w2.resize(ninc);
  for (int s = 0; s < nslots; ++s)
    if (cnt[s + 1] > cnt[s]) {
      out.drun_slot.push_back(s);
      out.drun_begin.push_back(cnt[s]);
    }
  pos.assign(cnt.begin(), cnt.end() - 1);
  auto put_inc = [&](int s, int k, int e) {
    const int j = pos[s]++;
    out.inc_code[j] = (k << 2) | e;
    out.inc_w2[j] = 2.0 * rng_w[k0 + k];
  };
  for (int k = 0; k < Ki; ++k) {
    const int sa = sa_v[k], sb = sb_v[k];
    if (sa >= 0 && sa == sb) {
      put_inc(sa, k, 2);
    } else {
      if (sa >= 0) put_inc(sa, k, 0);
      if (sb >= 0) put_inc(sb, k, 1);
    }
  }
  // ranges by slot pair
  const size_t nbins = (size_t)nslots * nslots;
  cnt.assign(nbins + 1, 0);
  for (int k = 0; k < Ki; ++k) {
    const int sa = sa_v[k], sb = sb_v[k];
    if (sa < 0 || sb < 0 || sa == sb) continue;
    cnt[(size_t)std::min(sa, sb) * nslots + std::max(sa, sb) + 1]++;
  }
  for (size_t q = 0; q < nbins; ++q) cnt[q + 1] += cnt[q];
  const int npair = cnt[nbins];
  out.pr_code.resize(npair);
  for (int lo = 0; lo < nslots; ++lo)
    for (int hi = lo + 1; hi < nslots; ++hi) {
      const size_t q = (size_t)lo * nslots + hi;
      if (cnt[q + 1] > cnt[q]) {
        out.orun_lo.push_back(lo);
        out.orun_hi.push_back(hi);
        out.orun_begin.push_back(cnt[q]);
      }
    }
  pos.assign(cnt.begin(), cnt.end() - 1);
  for (int k = 0; k < Ki; ++k) {
    const int sa = sa_v[k], sb = sb_v[k];
    if (sa < 0 || sb < 0 || sa == sb) continue;
    const size_t q = (size_t)std::min(sa, sb) * nslots + std::max(sa, sb);
    out.pr_code[pos[q]++] = (k << 1) | (sb < sa ? 1 : 0);
  }
  // contiguous, entry-balanced split of the off-diagonal runs over the warps (local run indices)
  const int nrun = (int)out.orun_lo.size();
  int w = 0;
  for (int r = 0; r < nrun; ++r) {
    const long long done = out.orun_begin[r];
    while (w < NW && done * NW >= (long long)(w + 1) * npair) out.ws[++w] = r;
  }
  while (w < NW) out.ws[++w] = nrun;
  out.ws[0] = 0;
}

static int build_coarse_tables(ScoreHandle_ *h, const ScoreProblemDesc *desc) {
  DevProblem &P = h->P;
  const int NI = P.n_inst;
  constexpr int NW = kCoarseThreads / 32;
  std::vector<int> rng_a(P.K), rng_b(P.K), seg_ptr(P.n_seg + 1);
  std::vector<double> rng_w(P.K);
  if (P.K) {
    SCORE_CUDA_CHECK(fetch_to_host(rng_a.data(), desc->rng_a, sizeof(int) * P.K));
    SCORE_CUDA_CHECK(fetch_to_host(rng_b.data(), desc->rng_b, sizeof(int) * P.K));
    SCORE_CUDA_CHECK(fetch_to_host(rng_w.data(), desc->rng_w, sizeof(double) * P.K));
  }
  SCORE_CUDA_CHECK(fetch_to_host(seg_ptr.data(), desc->seg_ptr, sizeof(int) * (P.n_seg + 1)));
  std::vector<CoarseInstTables> tabs(NI);
  h->c_ll.assign(NI, 0);
  {
    std::atomic<int> next(0);
    auto worker = [&]() {
      for (int i = next.fetch_add(1); i < NI; i = next.fetch_add(1))
        coarse_tables_one(h, i, rng_a.data(), rng_b.data(), rng_w.data(), seg_ptr.data(), tabs[i]);
    };
    // SCORE_CREATE_THREADS: upper bound of the table-building threads (default 16); in a pipelined sweep the threads
    // that feed graph launches of the running solves share the host cores with them
    static const int nt_max = getenv("SCORE_CREATE_THREADS") ? std::max(1, atoi(getenv("SCORE_CREATE_THREADS"))) : 16;
    const int nt = std::max(1, std::min<int>({(int)std::thread::hardware_concurrency(), nt_max, NI / 8 + 1}));
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(worker);
    worker();
    for (auto &t : pool) t.join();
  }
  std::vector<int> inc_off(NI + 1, 0), drun_off(NI + 1, 0), pr_off(NI + 1, 0), orun_off(NI + 1, 0);
  for (int i = 0; i < NI; ++i) {
    if (tabs[i].bad) {
      g_score_last_error = "range endpoint out of bounds";
      return SCORE_ERR_INVALID;
    }
    h->c_ll[i] = tabs[i].has_ll ? 1 : 0;
    inc_off[i + 1] = inc_off[i] + (int)tabs[i].inc_code.size();
    drun_off[i + 1] = drun_off[i] + (int)tabs[i].drun_slot.size();
    pr_off[i + 1] = pr_off[i] + (int)tabs[i].pr_code.size();
    orun_off[i + 1] = orun_off[i] + (int)tabs[i].orun_lo.size();
  }
  std::vector<int> inc_code(inc_off[NI]), drun_slot(drun_off[NI]), drun_begin(drun_off[NI] + 1), pr_code(pr_off[NI]);
  std::vector<int> orun_lo(orun_off[NI]), orun_hi(orun_off[NI]), orun_begin(orun_off[NI] + 1), owarp((size_t)NI * (NW + 1));
  std::vector<double> inc_w2(inc_off[NI]);
  for (int i = 0; i < NI; ++i) {
    const CoarseInstTables &t = tabs[i];
    std::copy(t.inc_code.begin(), t.inc_code.end(), inc_code.begin() + inc_off[i]);
    std::copy(t.inc_w2.begin(), t.inc_w2.end(), inc_w2.begin() + inc_off[i]);
    std::copy(t.pr_code.begin(), t.pr_code.end(), pr_code.begin() + pr_off[i]);
    std::copy(t.drun_slot.begin(), t.drun_slot.end(), drun_slot.begin() + drun_off[i]);
    std::copy(t.orun_lo.begin(), t.orun_lo.end(), orun_lo.begin() + orun_off[i]);
    std::copy(t.orun_hi.begin(), t.orun_hi.end(), orun_hi.begin() + orun_off[i]);
    for (size_t r = 0; r < t.drun_begin.size(); ++r) drun_begin[drun_off[i] + r] = inc_off[i] + t.drun_begin[r];
    for (size_t r = 0; r < t.orun_begin.size(); ++r) orun_begin[orun_off[i] + r] = pr_off[i] + t.orun_begin[r];
    for (int w = 0; w <= NW; ++w) owarp[(size_t)i * (NW + 1) + w] = orun_off[i] + t.ws[w];
  }
  drun_begin[drun_off[NI]] = inc_off[NI];
  orun_begin[orun_off[NI]] = pr_off[NI];
  P.c_ninc = inc_off[NI];
  P.c_npair = pr_off[NI];
  int rc;
  const int d1 = P.d + 1;
  if ((rc = upload(h, &P.c_inc_off, inc_off.data(), inc_off.size()))) return rc;
  if ((rc = upload(h, &P.c_inc_code, inc_code.data(), inc_code.size()))) return rc;
  if ((rc = upload(h, &P.c_inc_w2, inc_w2.data(), inc_w2.size()))) return rc;
  if ((rc = upload(h, &P.c_drun_off, drun_off.data(), drun_off.size()))) return rc;
  if ((rc = upload(h, &P.c_drun_slot, drun_slot.data(), drun_slot.size()))) return rc;
  if ((rc = upload(h, &P.c_drun_begin, drun_begin.data(), drun_begin.size()))) return rc;
  if ((rc = upload(h, &P.c_pr_off, pr_off.data(), pr_off.size()))) return rc;
  if ((rc = upload(h, &P.c_pr_code, pr_code.data(), pr_code.size()))) return rc;
  if ((rc = upload(h, &P.c_orun_lo, orun_lo.data(), orun_lo.size()))) return rc;
  if ((rc = upload(h, &P.c_orun_hi, orun_hi.data(), orun_hi.size()))) return rc;
  if ((rc = upload(h, &P.c_orun_begin, orun_begin.data(), orun_begin.size()))) return rc;
  if ((rc = upload(h, &P.c_owarp, owarp.data(), owarp.size()))) return rc;
  if ((rc = upload(h, &P.c_orun_off, orun_off.data(), orun_off.size()))) return rc;
  if ((rc = dalloc(h, &P.c_inc_h, (size_t)P.c_ninc * d1))) return rc;
  if ((rc = dalloc(h, &P.c_pr_h, (size_t)P.c_npair * 2 * d1))) return rc;
  return SCORE_OK;
}

// SCORE_TRACE_CREATE=1: print how long each phase of score_create took (host wall clock)
struct CreateTrace {
  bool on = getenv("SCORE_TRACE_CREATE") && atoi(getenv("SCORE_TRACE_CREATE")) != 0;
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now(), last = t0;
  std::string line;
  void mark(const char *what) {
    if (!on) return;
    const auto now = std::chrono::steady_clock::now();
    char buf[96];
    snprintf(buf, sizeof buf, " %s %.1f", what, std::chrono::duration<double, std::milli>(now - last).count());
    line += buf;
    last = now;
  }
  ~CreateTrace() {
    if (on)
      fprintf(stderr, "[score_create] total %.1f ms:%s\n",
              std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(), line.c_str());
  }
};

static int create_impl(const ScoreProblemDesc *desc, int32_t device, ScoreHandle_ *h) {
  CreateTrace trace;
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
  // footprint estimate for the first arena chunk: operator in both orientations + sort scratch + solver vectors
  h->arena_hint = (size_t)(90 * nz + 60 * m + 40 * (2 * desc->E + 2 * desc->K)) + (1u << 20);
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
  trace.mark("checks");
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
  SCORE_CUDA_CHECK(fetch_to_host(seg_ptr.data(), desc->seg_ptr, sizeof(int) * (P.n_seg + 1)));
  SCORE_CUDA_CHECK(fetch_to_host(seg_inst.data(), desc->seg_inst, sizeof(int) * P.n_seg));
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
  // factor indices are trusted by the assembly / preconditioner kernels: check them here (instance-local pose indices
  // of every relative-pose factor, the odometry link of every pose, the landmark of every prior)
  {
    std::vector<int> ei(P.E), ej(P.E), le(P.P), pl(P.Lp);
    if (P.E) {
      SCORE_CUDA_CHECK(fetch_to_host(ei.data(), desc->edge_i, sizeof(int) * P.E));
      SCORE_CUDA_CHECK(fetch_to_host(ej.data(), desc->edge_j, sizeof(int) * P.E));
    }
    SCORE_CUDA_CHECK(fetch_to_host(le.data(), desc->link_edge, sizeof(int) * P.P));
    if (P.Lp) SCORE_CUDA_CHECK(fetch_to_host(pl.data(), desc->prior_l, sizeof(int) * P.Lp));
    int n_nonlink = 0;
    for (int i = 0; i < NI; ++i) {
      const int Pi = h->pose_off[i + 1] - h->pose_off[i], Li = h->lm_off[i + 1] - h->lm_off[i];
      for (int e = h->edge_off[i]; e < h->edge_off[i + 1]; ++e) {
        if (ei[e] < 0 || ei[e] >= Pi || ej[e] < 0 || ej[e] >= Pi || ei[e] == ej[e]) {
          g_score_last_error = "relative-pose factor " + std::to_string(e) + ": pose index out of range or self-edge";
          return SCORE_ERR_INVALID;
        }
        if (le[h->pose_off[i] + ej[e]] != e) ++n_nonlink;
      }
      for (int q = h->prior_off[i]; q < h->prior_off[i + 1]; ++q)
        if (pl[q] < 0 || pl[q] >= Li) {
          g_score_last_error = "landmark prior " + std::to_string(q) + ": landmark index out of range";
          return SCORE_ERR_INVALID;
        }
      for (int p = h->pose_off[i]; p < h->pose_off[i + 1]; ++p) {
        const int e = le[p];
        if (e == -1) continue;
        const int pl_ = p - h->pose_off[i];
        if (e < h->edge_off[i] || e >= h->edge_off[i + 1] || ej[e] != pl_ || ei[e] != pl_ - 1) {
          g_score_last_error = "link_edge[" + std::to_string(p) + "] is not the odometry factor (p-1 -> p) of its instance";
          return SCORE_ERR_INVALID;
        }
      }
    }
    P.n_nonlink = n_nonlink;
    for (int s = 0; s < P.n_seg; ++s) {
      bool ok = le[seg_ptr[s]] == -1;
      for (int p = seg_ptr[s] + 1; ok && p < seg_ptr[s + 1]; ++p) ok = le[p] >= 0;
      if (!ok) {
        g_score_last_error = "a chain segment must start at a pose without an odometry link and be linked throughout";
        return SCORE_ERR_INVALID;
      }
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
  {  // dense item table of the chain-scan kernels: segment j of instance i at [i * maxseg + j]
    int maxseg = 1;
    for (int i = 0; i < NI; ++i) maxseg = std::max(maxseg, h->seg_begin[i + 1] - h->seg_begin[i]);
    std::vector<int4> tab((size_t)NI * maxseg, make_int4(-1, -1, -1, 0));
    for (int i = 0; i < NI; ++i)
      for (int sg = h->seg_begin[i]; sg < h->seg_begin[i + 1]; ++sg)
        tab[(size_t)i * maxseg + (sg - h->seg_begin[i])] = make_int4(seg_ptr[sg], seg_ptr[sg + 1], sg, 0);
    UP(seg_tab, tab.data(), tab.size())
  }
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
  DA(P.w, P.m)
  DA(P.b, P.m)
  DA(P.G, (size_t)P.P * blk)
  DA(P.M, (size_t)P.P * ((d + 1) * (d + 2) / 2))  // symmetric blocks, upper triangle
  DA(P.lm_inv, (size_t)P.L * d)
  DA(h->wsum, P.P)
  // index scratch of the incidence-list sort; the assembled CSR pair and its sort scratch are allocated on first use
  // (assemble_reduced): a matrix-free solve never needs them — 40 bytes per stored entry, ~2/3 of a handle's footprint
  h->sort_cap = (size_t)std::max<long long>(1, 2 * desc->E + 2 * desc->K + desc->Lp);
  DA(h->sort_keys, h->sort_cap)
  DA(h->sort_idx, h->sort_cap)
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
  DA(V.mk, (size_t)P.K * (d * (d + 1) / 2))
  trace.mark("offsets+uploads");
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
      const bool small = nc > 0 && nc <= kCoarseMax;
      const bool big = nc > kCoarseMax && nc <= kCoarseBigMax;
      const bool on = small || big;
      if (big) h->big.push_back(i);
      if (nc > kCoarseBigMax) {
        static std::atomic<bool> warned(false);
        if (!warned.exchange(true))
          fprintf(stderr, "[score_b200] instance %d: coarse space of %d coordinates exceeds %d; solving it without a coarse "
                          "level (correct, but PCG converges slowly)\n", i, nc, kCoarseBigMax);
      }
      h->c_n[i] = on ? nc : 0;
      h->c_nb[i] = on ? nb : 0;
      h->c_off[i + 1] = h->c_off[i] + (on ? nc : 0);
      moff += on ? (long long)nc * nc : 0;
      if (moff >= (1ll << 31)) {
        g_score_last_error = "coarse matrices too large for 32-bit indexing";
        return SCORE_ERR_INVALID;
      }
      h->c_moff[i + 1] = (int)moff;
      if (small && nc > h->c_nmax) h->c_nmax = nc;
    }
    if ((rc = upload(h, &P.c_off, h->c_off.data(), NI + 1))) return rc;
    if ((rc = upload(h, &P.c_moff, h->c_moff.data(), NI + 1))) return rc;
    if ((rc = upload(h, &P.c_n, h->c_n.data(), NI))) return rc;
    if ((rc = upload(h, &P.c_nb, h->c_nb.data(), NI))) return rc;
    h->c_big = NI == 1 && !h->big.empty();
    if (!h->big.empty()) {
      int nbig = 0;
      for (int i : h->big) nbig = std::max(nbig, h->c_n[i]);
      h->cb_ldp = (nbig + 7) & ~7;
      DA(h->cb_work, (size_t)nbig * nbig)
      DA(h->cb_W, kDB * kDB)
      DA(h->cb_raw, (size_t)kDB * h->cb_ldp)
      DA(h->cb_scl, (size_t)kDB * h->cb_ldp)
    }
    DA(P.c_Ainv, (size_t)moff)
    DA(P.c_rhs, h->c_off[NI])
    DA(P.c_sol, h->c_off[NI])
    trace.mark("coarse-dims");
    if ((rc = build_coarse_tables(h, desc))) return rc;
    trace.mark("coarse-tables");
    static const bool no_schur = getenv("SCORE_NO_SCHUR") && atoi(getenv("SCORE_NO_SCHUR")) != 0;  // A/B knob
    for (int i : h->big) {
      ScoreHandle_::BigInst b;
      const int nb = h->c_nb[i], nl = h->c_n[i] - nb;
      b.schur = !no_schur && nb > 0 && nl > 0 && !h->c_ll[i];
      if (b.schur) {
        b.ldp = (nb + 7) & ~7;
        DA(b.S21, (size_t)nl * b.ldp)
        DA(b.Y, (size_t)nl * b.ldp)
        DA(b.Dinv, (size_t)nl * d)
        DA(b.part, (size_t)kSchurChunks * nb)
        DA(b.w, nb)
      }
      h->bigx.push_back(b);
    }
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
    // landmark columns are processed one warp per column: heavy columns (many ranges per landmark) get one
    // column per warp and CTA-sized blocks of 8 so that they spread over the SMs
    const int Li = h->lm_off[i + 1] - h->lm_off[i], Ki = h->rng_off[i + 1] - h->rng_off[i];
    const int lstep = (Li > 0 && Ki / Li > 128) ? kThreads / 32 : 64;
    for (int c0 = pc1; c0 < h->zoff[i + 1]; c0 += lstep)
      cb.push_back({i, c0, std::min(c0 + lstep, h->zoff[i + 1]), CB_LANDMARK});
  }
  h->rb_begin[NI] = (int)rb.size();
  h->cb_begin[NI] = (int)cb.size();
  // pose / landmark blocks of the matrix-free Hessian-vector kernel (instance-local pose / landmark ranges)
  std::vector<HvBlock> pb;
  h->pb_begin.assign(NI + 1, 0);
  for (int i = 0; i < NI; ++i) {
    h->pb_begin[i] = (int)pb.size();
    const int Pi = h->pose_off[i + 1] - h->pose_off[i], Li = h->lm_off[i + 1] - h->lm_off[i];
    const int z0 = h->zoff[i], pg0 = h->pose_off[i], lg0 = h->lm_off[i];
    for (int p0 = 0; p0 < Pi; p0 += kPosesPerBlock)
      pb.push_back({i, p0, std::min(p0 + kPosesPerBlock, Pi), CB_POSE, z0, pg0, Pi, lg0, (int)pb.size(), 0, 0, 0});
    for (int q0 = 0; q0 < Li; q0 += kLmPerBlock)
      pb.push_back({i, q0, std::min(q0 + kLmPerBlock, Li), CB_LANDMARK, z0, pg0, Pi, lg0, (int)pb.size(), 0, 0, 0});
  }
  h->pb_begin[NI] = (int)pb.size();
  h->T.n_pb = (int)pb.size();
  if ((rc = upload(h, &h->T.pb, pb.data(), pb.size()))) return rc;
  {  // dense item table: block j of instance i at [i * maxpb + j] (jb / je are filled on the device: k_hv_fill)
    int maxpb = 1;
    for (int i = 0; i < NI; ++i) maxpb = std::max(maxpb, h->pb_begin[i + 1] - h->pb_begin[i]);
    HvBlock none{};
    none.kind = -1;
    std::vector<HvBlock> dense((size_t)NI * maxpb, none);
    for (int i = 0; i < NI; ++i)
      for (int b = h->pb_begin[i]; b < h->pb_begin[i + 1]; ++b) dense[(size_t)i * maxpb + (b - h->pb_begin[i])] = pb[b];
    h->n_pbd = (int)dense.size();
    if ((rc = upload(h, &h->T.pbd, dense.data(), dense.size()))) return rc;
  }
  if ((rc = upload(h, &h->T.pb_begin, h->pb_begin.data(), NI + 1))) return rc;
  h->T.n_rb = (int)rb.size();
  h->T.n_cb = (int)cb.size();
  if ((rc = upload(h, &h->T.rb, rb.data(), rb.size()))) return rc;
  if ((rc = upload(h, &h->T.cb, cb.data(), cb.size()))) return rc;
  if ((rc = upload(h, &h->T.rb_begin, h->rb_begin.data(), NI + 1))) return rc;
  if ((rc = upload(h, &h->T.cb_begin, h->cb_begin.data(), NI + 1))) return rc;
  {
    // work lists: [par, ticket, cnt x6, cnt_ev, pad] + 7 lists of n_inst entries
    int maxrb = 1, maxcb = 1, maxseg = 1, maxpb = 1, maxvc = 1;
    for (int i = 0; i < NI; ++i) {
      maxpb = std::max(maxpb, h->pb_begin[i + 1] - h->pb_begin[i]);
      maxvc = std::max(maxvc, (h->zoff[i + 1] - h->zoff[i] + kVecChunk - 1) / kVecChunk);
      maxrb = std::max(maxrb, h->rb_begin[i + 1] - h->rb_begin[i]);
      maxcb = std::max(maxcb, h->cb_begin[i + 1] - h->cb_begin[i]);
      maxseg = std::max(maxseg, h->seg_begin[i + 1] - h->seg_begin[i]);
    }
    DA(h->wl_mem, 16 + 7 * (size_t)NI)
    WorkLists &W = h->W;
    W.par = h->wl_mem;
    W.ticket = h->wl_mem + 1;
    W.cnt = h->wl_mem + 2;
    W.cnt_ev = h->wl_mem + 8;
    W.lists = h->wl_mem + 16;
    W.ev = W.lists + (size_t)6 * NI;
    W.n_inst = NI;
    W.maxrb = maxrb;
    W.maxcb = maxcb;
    W.maxseg = maxseg;
    W.maxpb = maxpb;
    W.maxvc = maxvc;
    // (cudaGetDeviceProperties fills the whole property struct through the driver and takes milliseconds — far
    // longer when other threads are launching work; one attribute is all that is needed)
    SCORE_CUDA_CHECK(cudaDeviceGetAttribute(&h->n_sm, cudaDevAttrMultiProcessorCount, device));
    h->n_clusters = std::max(1, h->n_sm / kClusterSize - 2);  // (GPC boundaries leave a few SMs outside any cluster)
  }
  h->red_count = 5 * rb.size() + (size_t)P.nz;
  DA(h->red_send, h->red_count)
  V.part_row = h->red_send;
  V.part_upd = h->red_send + rb.size();
  V.hloc = h->red_send + 5 * rb.size();
  V.hglob = V.hloc;
  DA(V.part_ls, rb.size() * kLsSums)
  h->W.rb_lo = 0;
  h->W.rb_hi = (int)rb.size();
  DA(V.part_col, cb.size() * 4)
  DA(V.h, P.nz)
  DA(V.part_hv, pb.size())
  DA(V.part_gr, pb.size() * 4)
  P.n_inc = 2 * P.E + 2 * P.K + P.Lp;
  DA(P.inc_ptr, (size_t)P.P + P.L + 1)
  DA(P.inc_rec, P.n_inc)
  DA(h->inc_in, P.n_inc)
  DA(V.part_seg, P.n_seg)
  DA(V.part_lm, NI)
  DA(h->st, NI)
  DA(h->d_ndone, 1)
  DA(h->bar_mem, 2 * kMaxFusedGroups)
  DA(h->out_poses, (size_t)P.P * blk)
  DA(h->out_lms, (size_t)P.L * d)
  DA(h->out_round, (size_t)P.P * d * d)
  DA(h->out_dist, (size_t)P.K * h->dist_per)
#undef DA
  h->h_ndone = g_pinned.get();
  if (!h->h_ndone) SCORE_CUDA_CHECK(cudaMallocHost((void **)&h->h_ndone, 2 * sizeof(int)));
  for (auto &e : h->ev_done) SCORE_CUDA_CHECK(g_cache.get_event(device, &e));
  trace.mark("block-tables+allocs");
  if (P.n_inc > 0) {  // radix-sort scratch of the incidence lists of the matrix-free operator (keys: owners)
    int ob = 1;
    while ((1ll << ob) <= (long long)P.P + P.L + 1) ++ob;  // owners 0 .. P + L (the last one: the odometry-link sentinel)
    static std::mutex mu;
    static std::map<std::pair<long long, int>, size_t> known;
    std::lock_guard<std::mutex> lk(mu);
    const auto key = std::make_pair((long long)P.n_inc, ob);
    auto it = known.find(key);
    if (it == known.end()) {
      size_t bytes = 0;
      SCORE_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, bytes, h->sort_idx, h->sort_keys, h->inc_in, P.inc_rec,
                                                       P.n_inc, 0, ob, (cudaStream_t)0));
      it = known.emplace(key, bytes).first;
    }
    h->inc_tmp_bytes = it->second;
    char *tmp = nullptr;
    if ((rc = dalloc(h, &tmp, h->inc_tmp_bytes))) return rc;
    h->inc_tmp = tmp;
  }
  SCORE_CUDA_CHECK(g_cache.get_stream(device, &h->own_stream));
  SCORE_CUDA_CHECK(g_cache.get_hi_stream(device, &h->hi_stream));
  SCORE_CUDA_CHECK(g_cache.get_event(device, &h->ev_switch));
  // the allocations and uploads above are ordered on the default stream; the handle works on its own (non-blocking)
  // stream.  Wait for the default stream only — a device-wide synchronisation would also wait for (and be delayed
  // by) the solves of other handles that are running concurrently.
  trace.mark("stream-create");
  SCORE_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)0));
  trace.mark("sync");
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


// Optional per-kernel CUDA-event timing of un-graphed ticks (profile mode).  Events come from the process-wide
// cache and go back to it.  Every mark carries the tick it belongs to (tag 0: line-search tick, 1: evaluation
// tick, 2 + i: i-th PCG tick of the cycle), so that launches whose work list is the whole batch — the line-search
// tick's line-search-only kernels and the FIRST PCG tick of an early cycle — can be averaged on their own.
struct TickProfiler {
  std::vector<cudaEvent_t> ev;
  std::vector<int> ids, tags;
  cudaStream_t st = nullptr;
  int device = 0, tag = 0;
  void mark(int id) {
    cudaEvent_t e;
    if (g_cache.get_tevent(device, &e) != cudaSuccess) return;
    cudaEventRecord(e, st);
    ev.push_back(e);
    ids.push_back(id);
    tags.push_back(tag);
  }
  // call after the stream is synchronised
  void collect(double *ms, long long *count, double *ms_full, long long *count_full) {
    for (size_t i = 0; i + 1 < ev.size(); ++i) {
      if (ids[i] < 0) continue;
      float t = 0.f;
      cudaEventElapsedTime(&t, ev[i], ev[i + 1]);
      ms[ids[i]] += t;
      count[ids[i]] += 1;
      const bool ls_only = ids[i] == KI_LINESEARCH || ids[i] == KI_ROWUPDATE || ids[i] == KI_COARSE_BUILD;
      if ((ls_only && tags[i] == 0) || (!ls_only && tags[i] == 2)) {
        ms_full[ids[i]] += t;
        count_full[ids[i]] += 1;
      }
    }
    for (auto &e : ev) g_cache.put_tevent(device, e);
    ev.clear();
    ids.clear();
    tags.clear();
  }
  ~TickProfiler() {
    for (auto &e : ev) g_cache.put_tevent(device, e);
  }
};

// Coarse build kernel for the handle's largest coarse space: tile size TS = ceil(nmax / 32).
template <int D, int TS>
static int coarse_build_attr(size_t smem) {
  SCORE_CUDA_CHECK(cudaFuncSetAttribute(k_coarse_build<D, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  return SCORE_OK;
}
static int coarse_build_prepare(ScoreHandle_ *h) {
  const int ts = coarse_tile_size(h->c_nmax), d = h->P.d;
  const size_t smem = coarse_smem_bytes(d, ts);
  switch (ts * 10 + d) {
    case 12: return coarse_build_attr<2, 1>(smem);
    case 22: return coarse_build_attr<2, 2>(smem);
    case 32: return coarse_build_attr<2, 3>(smem);
    case 42: return coarse_build_attr<2, 4>(smem);
    case 13: return coarse_build_attr<3, 1>(smem);
    case 23: return coarse_build_attr<3, 2>(smem);
    case 33: return coarse_build_attr<3, 3>(smem);
    case 43: return coarse_build_attr<3, 4>(smem);
  }
  g_score_last_error = "internal: unsupported coarse tile size";
  return SCORE_ERR_STATE;
}
template <int D>
static void launch_coarse_build(ScoreHandle_ *h, const SolverCfg &cfg, cudaStream_t st) {
  const DevProblem &P = h->P;
  const int ts = coarse_tile_size(h->c_nmax);
  const size_t smem = coarse_smem_bytes_d<D>(ts);
#define SCORE_CB(TS) \
  k_coarse_build<D, TS><<<wgrid(h, P.n_inst, 1), kCoarseThreads, smem, st>>>(P, h->n_ranks > 1 ? h->Vg : h->V, h->st, cfg.coarse_reg, cfg.coarse_every, h->W)
  switch (ts) {
    case 1: SCORE_CB(1); break;
    case 2: SCORE_CB(2); break;
    case 3: SCORE_CB(3); break;
    default: SCORE_CB(4); break;
  }
#undef SCORE_CB
}

// Dense coarse level of a large instance: accumulate into the global-memory work matrix, invert it by blocked
// symmetric sweeps (dense.cuh; every kernel gated on the instance's phase), symmetrised copy into c_Ainv.
template <int D>
static int launch_coarse_big_build(ScoreHandle_ *h, const SolverCfg &cfg, int bi, cudaStream_t st) {
  const DevProblem &P = h->P;
  const int inst = h->big[bi];
  const ScoreHandle_::BigInst &B = h->bigx[bi];
  const int n = h->c_n[inst], nb = h->c_nb[inst], nl = n - nb;
  double *A = h->cb_work;
  cudaMemsetAsync(A, 0, sizeof(double) * (size_t)n * n, st);
  k_coarse_big_accum<D><<<h->n_sm * 4, kBigThreads, 0, st>>>(P, h->n_ranks > 1 ? h->Vg : h->V, cfg.coarse_reg, A, inst, h->st);
  k_coarse_big_finish<D><<<grid_for((long)n * n, 256), 256, 0, st>>>(P, A, inst, h->st);
  if (!B.schur) {
    const int k = launch_dense_sweep(A, n, n, h->cb_W, h->cb_raw, h->cb_scl, h->cb_ldp, h->st, inst, st);
    k_dense_finish<<<grid_for((long)n * n, 256), 256, 0, st>>>(A, n, n, P.c_Ainv + h->c_moff[inst], n, h->st, inst);
    return 3 + k;
  }
  // landmark block eliminated: T = S11 - S12 Dl^-1 S12^T in place of S11 (the leading nb x nb block of A, stride n)
  k_schur_prep<D><<<grid_for((long)(nl / D) * nb, 256), 256, 0, st>>>(A, n, nb, B.S21, B.Y, B.Dinv, B.ldp, h->st, inst);
  const int nt = (nb + kDT - 1) / kDT;
  k_dense_update<<<dim3(nt, nt), 256, 0, st>>>(A, nb, n, B.S21, B.Y, B.ldp, nl, 0, nullptr, h->st, inst);
  const int k = launch_dense_sweep(A, nb, n, h->cb_W, h->cb_raw, h->cb_scl, h->cb_ldp, h->st, inst, st);
  k_dense_finish<<<grid_for((long)nb * nb, 256), 256, 0, st>>>(A, nb, n, P.c_Ainv + h->c_moff[inst], nb, h->st, inst);
  return 5 + k;
}

// y = A_c^-1 c of a large instance, then scatter (segment bases -> ytmp, landmarks -> s, partial r.s)
template <int D>
static int launch_coarse_big_apply(ScoreHandle_ *h, int bi, cudaStream_t st) {
  const DevProblem &P = h->P;
  const int inst = h->big[bi];
  const ScoreHandle_::BigInst &B = h->bigx[bi];
  const int n = h->c_n[inst], nb = h->c_nb[inst], nl = n - nb;
  int k = 0;
  if (!B.schur) {
    k_coarse_big_apply<<<grid_for(n, kBigThreads / 32), kBigThreads, 0, st>>>(P, h->st, inst);
    k = 1;
  } else {
    const double *crhs = P.c_rhs + h->c_off[inst];
    double *sol = P.c_sol + h->c_off[inst];
    const double *Tinv = P.c_Ainv + h->c_moff[inst];
    k_schur_apply1<D><<<dim3(grid_for(nb, 256), kSchurChunks), 256, 0, st>>>(B.S21, B.Dinv, crhs, nb, nl, B.ldp, B.part, h->st, inst);
    k_schur_apply2<<<grid_for(nb, 256), 256, 0, st>>>(crhs, B.part, nb, B.w, h->st, inst);
    k_schur_matvec<D, false><<<grid_for(nb, 8), 256, 0, st>>>(Tinv, nb, nb, nb, B.w, sol, nullptr, nullptr, h->st, inst);
    k_schur_matvec<D, true><<<grid_for(nl, 8), 256, 0, st>>>(B.Y, nl, nb, B.ldp, sol, sol + nb, B.Dinv, crhs + nb, h->st, inst);
    k = 4;
  }
  k_coarse_big_scatter<D><<<1, kBigThreads, 0, st>>>(P, h->V, h->st, inst);
  return k + 1;
}

static bool coarse_apply_split() {
  static const bool split = getenv("SCORE_SPLIT_COARSE_APPLY") && atoi(getenv("SCORE_SPLIT_COARSE_APPLY")) != 0;
  return split;
}

// Wait for a cycle's completion event.  A single graph's cycle is a few hundred microseconds: polled, the host sees it
// end within a microsecond; a batch's cycle is milliseconds: after SCORE_SPIN_US of polling the thread sleeps on the
// (blocking-sync) event, so the 4-6 solves a sweep keeps in flight per GPU do not each burn a host core.
#ifndef SCORE_SPIN_US
#define SCORE_SPIN_US 150
#endif
static cudaError_t wait_event_hybrid(cudaEvent_t ev) {
  const auto t0 = std::chrono::steady_clock::now();
  for (;;) {
    const cudaError_t e = cudaEventQuery(ev);
    if (e != cudaErrorNotReady) return e;
    if (std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(SCORE_SPIN_US)) break;
  }
  return cudaEventSynchronize(ev);
}

// Preconditioner application s = P r (+ partial r.s), shared by the line-search and PCG ticks.
template <int D>
static int launch_precond(ScoreHandle_ *h, cudaStream_t st, TickProfiler *pf) {
  const DevProblem &P = h->P;
  if (pf) pf->mark(KI_PRECOND_REV);
  k_precond_rev<D><<<wgrid(h, (long)P.n_inst * (h->W.maxseg + 1), 16), kSegThreads, 0, st>>>(P, h->V, h->st, h->W);
  // small coarse spaces: the application A_c^-1 c is fused into the forward pass (SCORE_SPLIT_COARSE_APPLY=1 keeps
  // the stand-alone kernel, for A/B measurements; both give the same bits)
  const bool split = coarse_apply_split();
  const bool fuse = h->c_nmax > 0 && !split;
  int nbig = 0;
  if (h->c_nmax > 0 && split) {
    if (pf) pf->mark(KI_COARSE_APPLY);
    k_coarse_apply<D><<<wgrid(h, P.n_inst, 8), kCoarseApplyThreads, 0, st>>>(P, h->V, h->st, h->W);
  }
  if (!h->big.empty()) {
    if (pf) pf->mark(KI_COARSE_APPLY);
    for (int bi = 0; bi < (int)h->big.size(); ++bi) nbig += launch_coarse_big_apply<D>(h, bi, st);
  }
  if (pf) pf->mark(KI_PRECOND_FWD);
  k_precond_fwd<D><<<wgrid(h, (long)P.n_inst * h->W.maxseg, 16), kSegThreads, 0, st>>>(P, h->V, h->st, h->W, fuse);
  return 2 + ((h->c_nmax > 0 && split) ? 1 : 0) + nbig;
}

// Row-partitioned solve: sum a buffer over the ranks (out of place; every rank contributes its own rows only).
static void dist_allreduce(ScoreHandle_ *h, const double *send, double *recv, size_t count, cudaStream_t st) {
  g_nccl.AllReduce(send, recv, count, ncclDouble, ncclSum, h->comm, st);
}

// h = B^T u and the column-space update.  One fused kernel on a single GPU; with row partitioning the SpMV of
// the local rows, ONE all-reduce of [partial sums | h] over NVLink, then the update from the summed h.
static int launch_colpass(ScoreHandle_ *h, cudaStream_t st, int mode) {
  const DevProblem &P = h->P;
  const int grid = wgrid(h, (long)P.n_inst * h->W.maxcb, 8);
  if (h->n_ranks == 1) {
    k_colpass<<<grid, kThreads, 0, st>>>(P, h->V, h->T, h->st, mode, CS_FUSED, h->W);
    return 1;
  }
  k_colpass<<<grid, kThreads, 0, st>>>(P, h->V, h->T, h->st, mode, CS_SPMV, h->W);
  dist_allreduce(h, h->red_send, h->red_recv, h->red_count, st);
  return 2;
}
static void launch_colapply(ScoreHandle_ *h, cudaStream_t st, int mode) {
  const DevProblem &P = h->P;
  k_colpass<<<wgrid(h, (long)P.n_inst * h->W.maxcb, 8), kThreads, 0, st>>>(P, h->Vg, h->T, h->st, mode, CS_APPLY, h->W);
}

// Line-search tick: step along dz, new residual / gradient / curvature / coarse matrix, first preconditioned
// residual of the next Newton system.  Returns the number of kernels launched.
template <int D>
static int launch_ls_tick(ScoreHandle_ *h, const SolverCfg &cfg, cudaStream_t st, TickProfiler *pf = nullptr,
                          bool big_build = false) {
  const DevProblem &P = h->P;
  const bool dist = h->n_ranks > 1;
  const SolverVecs &Vc = dist ? h->Vg : h->V;  // what the controllers / coarse build read
  int n = 0;
  if (pf) pf->mark(KI_ROWPASS);
  if (h->V.mf)  // bdz = B dz factor by factor
    k_rows_mf<D><<<wgrid(h, (long)P.n_inst * h->W.maxrb, 8), kThreads, 0, st>>>(P, h->V, h->T, h->st, h->W, RM_BDZ);
  else
    k_rowpass<D><<<wgrid(h, (long)P.n_inst * h->W.maxrb, 8), kThreads, 0, st>>>(P, h->V, h->T, h->st, h->W);
  if (h->V.mf) {  // instances still inside a Newton solve take their PCG iteration matrix-free here as well
    if (pf) pf->mark(KI_HESSVEC);
    k_hessvec<D><<<wgrid(h, (long)P.n_inst * h->W.maxpb, 8), kThreads, 0, st>>>(P, h->V, h->T, h->st, h->W);
    n += 1;
  }
  if (pf) pf->mark(KI_LINESEARCH);
  k_linesearch<D><<<wgrid(h, (long)P.n_inst * h->W.maxrb, 3), kThreads, 0, st>>>(P, h->V, h->T, h->st, h->W);
  if (dist) {
    // line-search sums, and p'Hp of an instance that is still inside a Newton solve during this tick
    dist_allreduce(h, h->V.part_ls, h->ls_recv, (size_t)h->T.n_rb * kLsSums, st);
    dist_allreduce(h, h->V.part_row, h->Vg.part_row, (size_t)h->T.n_rb, st);
  }
  if (pf) pf->mark(KI_CTRL_A);
  k_ctrl_a<<<wgrid(h, grid_for(P.n_inst, kSegThreads / 32), 16), kSegThreads, 0, st>>>(Vc, h->T, h->st, cfg, h->W, TM_LS);
  if (pf) pf->mark(KI_ROWUPDATE);
  k_rowupdate<D><<<wgrid(h, (long)P.n_inst * h->W.maxrb, 4), kThreads, 0, st>>>(P, h->V, h->T, h->st, TM_LS, h->W);
  n += 4;
  const bool build_big = !h->big.empty() && (!h->c_big || big_build);  // a single large instance: host-decided
  const bool build = h->c_nmax > 0 || build_big;
  if (dist && build)  // every rank needs the curvature blocks of all ranges
    dist_allreduce(h, h->V.mk, h->mk_recv, (size_t)P.K * (D * (D + 1) / 2), st);
  if (h->c_nmax > 0) {
    if (pf) pf->mark(KI_COARSE_BUILD);
    launch_coarse_build<D>(h, cfg, st);
    n += 1;
  }
  if (build_big) {
    if (pf) pf->mark(KI_COARSE_BUILD);
    for (int bi = 0; bi < (int)h->big.size(); ++bi) n += launch_coarse_big_build<D>(h, cfg, bi, st);
  }
  if (pf) pf->mark(KI_COLPASS);
  if (h->V.mf) {
    // gradient at the new point gathered from u factor by factor (+ the step itself); instances still inside a Newton
    // solve take their PCG update (step length from k_ctrl_a above)
    k_grad_mf<D><<<wgrid(h, (long)P.n_inst * h->W.maxpb, 8), kThreads, 0, st>>>(P, h->V, h->T, h->st, TM_LS, h->W);
    k_cg_update<false><<<wgrid(h, (long)P.n_inst * h->W.maxvc, 8), kThreads, 0, st>>>(P, h->V, h->T, h->st, h->W);
    n += 2;
  } else {
    n += launch_colpass(h, st, TM_LS);
  }
  if (dist) {
    launch_colapply(h, st, TM_LS);
    n += 1;
  }
  n += launch_precond<D>(h, st, pf);
  if (pf) pf->mark(KI_CTRL_B);
  k_ctrl_b<<<wgrid(h, grid_for(P.n_inst, kSegThreads / 32), 16), kSegThreads, 0, st>>>(P, Vc, h->T, h->st, cfg, h->d_ndone, TM_LS, h->W);
  if (pf) pf->mark(KI_PUPDATE);
  k_pupdate_vec<<<wgrid(h, (long)P.n_inst * h->W.maxvc, 8), kThreads, 0, st>>>(P, h->V, h->st, h->W);
  if (pf) pf->mark(-1);
  return n + 2;
}

// Evaluation tick: certificate of the un-smoothed problem for the instances that asked for it.
template <int D>
static int launch_eval_tick(ScoreHandle_ *h, const SolverCfg &cfg, cudaStream_t st, TickProfiler *pf = nullptr) {
  const DevProblem &P = h->P;
  const bool dist = h->n_ranks > 1;
  const SolverVecs &Vc = dist ? h->Vg : h->V;
  int n = 2;
  if (pf) pf->mark(KI_ROWUPDATE);
  k_rowupdate<D><<<wgrid(h, (long)P.n_inst * h->W.maxrb, 4), kThreads, 0, st>>>(P, h->V, h->T, h->st, TM_EVAL, h->W);
  if (pf) pf->mark(KI_COLPASS);
  if (h->V.mf) {
    k_grad_mf<D><<<wgrid(h, (long)P.n_inst * h->W.maxpb, 8), kThreads, 0, st>>>(P, h->V, h->T, h->st, TM_EVAL, h->W);
    n += 1;
  } else {
    n += launch_colpass(h, st, TM_EVAL);
  }
  if (dist) {
    launch_colapply(h, st, TM_EVAL);
    n += 1;
  }
  if (pf) pf->mark(KI_CTRL_B);
  k_ctrl_b<<<wgrid(h, grid_for(P.n_inst, kSegThreads / 32), 16), kSegThreads, 0, st>>>(P, Vc, h->T, h->st, cfg, h->d_ndone, TM_EVAL, h->W);
  if (pf) pf->mark(-1);
  return n;
}

// PCG tick: one preconditioned conjugate-gradient iteration of every instance still solving its Newton system.
// Row-partitioned: the SpMV of B^T runs before the controller so that its partial h and the partial p'Hp travel
// in the same all-reduce (one collective per PCG iteration).
template <int D>
static int launch_cg_tick(ScoreHandle_ *h, const SolverCfg &cfg, cudaStream_t st, bool last, TickProfiler *pf = nullptr) {
  const DevProblem &P = h->P;
  const bool dist = h->n_ranks > 1;
  const SolverVecs &Vc = dist ? h->Vg : h->V;
  int n = 0;
  if (h->V.mf) {
    if (pf) pf->mark(KI_HESSVEC);
    k_hessvec<D><<<wgrid(h, (long)P.n_inst * h->W.maxpb, 8), kThreads, 0, st>>>(P, h->V, h->T, h->st, h->W);
  } else {
    if (pf) pf->mark(KI_ROWPASS);
    k_rowpass<D><<<wgrid(h, (long)P.n_inst * h->W.maxrb, 8), kThreads, 0, st>>>(P, h->V, h->T, h->st, h->W);
  }
  if (dist) {
    if (pf) pf->mark(KI_COLPASS);
    n += launch_colpass(h, st, TM_CG);
  }
  const bool own_alpha = h->V.mf && !dist;  // the element-wise update computes the step length itself: no controller launch
  if (!own_alpha) {
    if (pf) pf->mark(KI_CTRL_A);
    k_ctrl_a<<<wgrid(h, grid_for(P.n_inst, kSegThreads / 32), 16), kSegThreads, 0, st>>>(Vc, h->T, h->st, cfg, h->W, TM_CG);
  }
  if (pf) pf->mark(KI_COLPASS);
  if (dist)
    launch_colapply(h, st, TM_CG);
  else if (h->V.mf)  // the operator was applied by k_hessvec: what is left of the column pass is element-wise
    k_cg_update<true><<<wgrid(h, (long)P.n_inst * h->W.maxvc, 8), kThreads, 0, st>>>(P, h->V, h->T, h->st, h->W);
  else
    launch_colpass(h, st, TM_CG);
  n -= own_alpha ? 1 : 0;
  n += 3 + launch_precond<D>(h, st, pf);
  if (pf) pf->mark(KI_CTRL_B);
  k_ctrl_b<<<wgrid(h, grid_for(P.n_inst, kSegThreads / 32), 16), kSegThreads, 0, st>>>(P, Vc, h->T, h->st, cfg, h->d_ndone, last ? TM_CG_LAST : TM_CG, h->W);
  if (pf) pf->mark(KI_PUPDATE);
  k_pupdate_vec<<<wgrid(h, (long)P.n_inst * h->W.maxvc, 8), kThreads, 0, st>>>(P, h->V, h->st, h->W);
  if (pf) pf->mark(-1);
  return n + 2;
}

// One cycle = line-search tick + evaluation tick + n_cg PCG ticks.
template <int D>
static int launch_cycle(ScoreHandle_ *h, const SolverCfg &cfg, cudaStream_t st, int n_cg, TickProfiler *pf = nullptr,
                        bool big_build = false) {
  if (pf) pf->tag = 0;
  int n = launch_ls_tick<D>(h, cfg, st, pf, big_build);
  if (pf) pf->tag = 1;
  n += launch_eval_tick<D>(h, cfg, st, pf);
  for (int i = 0; i < n_cg; ++i) {
    if (pf) pf->tag = 2 + i;
    n += launch_cg_tick<D>(h, cfg, st, i == n_cg - 1, pf);
  }
  return n;
}
// Tail cycle (few instances left, or a handle that holds few graphs): line-search tick + evaluation tick, then ONE
// fused kernel in which a thread-block cluster per instance runs the whole PCG solve of the next Newton system
// (fused.cuh), then the controller files the work lists.  Same per-instance arithmetic as the lockstep cycle.
template <int D>
static int launch_tail_cycle(ScoreHandle_ *h, const SolverCfg &cfg, cudaStream_t st, TickProfiler *pf = nullptr) {
  const DevProblem &P = h->P;
  if (pf) pf->tag = 0;
  int n = launch_ls_tick<D>(h, cfg, st, pf, false);
  if (pf) pf->tag = 1;
  n += launch_eval_tick<D>(h, cfg, st, pf);
  if (pf) pf->tag = 3;
  if (pf) pf->mark(KI_PCG_FUSED);
  static const bool use_clusters = getenv("SCORE_FUSED_CLUSTERS") && atoi(getenv("SCORE_FUSED_CLUSTERS")) != 0;  // A/B knob
  if (use_clusters) {
    const int ncl = std::max(1, std::min(h->n_clusters, P.n_inst));
    cudaLaunchConfig_t lc{};
    lc.gridDim = dim3(ncl * kClusterSize);
    lc.blockDim = dim3(kThreads);
    lc.dynamicSmemBytes = 0;
    lc.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = kClusterSize;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    lc.attrs = at;
    lc.numAttrs = 1;
    cudaLaunchKernelEx(&lc, k_pcg_fused<D, 0>, P, h->V, h->T, h->st, cfg, h->d_ndone, h->W, cfg.max_cg + 1, h->bar_mem);
  } else {
    // software-barrier groups: the whole grid must be resident at once (spinning CTAs never yield their SM)
    const int groups = std::max(1, std::min({h->n_sm * 4 / kGroupSize, P.n_inst, kMaxFusedGroups}));
    k_pcg_fused<D, 1><<<groups * kGroupSize, kThreads, 0, st>>>(P, h->V, h->T, h->st, cfg, h->d_ndone, h->W, cfg.max_cg + 1, h->bar_mem);
  }
  if (pf) pf->mark(KI_CTRL_B);
  k_ctrl_b<<<wgrid(h, grid_for(P.n_inst, kSegThreads / 32), 16), kSegThreads, 0, st>>>(P, h->V, h->T, h->st, cfg, h->d_ndone, TM_FILE, h->W);
  if (pf) pf->mark(-1);
  return n + 2;
}
static int launch_tail_cycle_d(ScoreHandle_ *h, const SolverCfg &cfg, cudaStream_t st, TickProfiler *pf = nullptr) {
  return h->P.d == 2 ? launch_tail_cycle<2>(h, cfg, st, pf) : launch_tail_cycle<3>(h, cfg, st, pf);
}
static int launch_cycle_d(ScoreHandle_ *h, const SolverCfg &cfg, cudaStream_t st, int n_cg, TickProfiler *pf = nullptr,
                          bool big_build = false) {
  return h->P.d == 2 ? launch_cycle<2>(h, cfg, st, n_cg, pf, big_build) : launch_cycle<3>(h, cfg, st, n_cg, pf, big_build);
}

// PCG ticks of cycle c: `base` early on, doubling every `grow_every` cycles after `grow_after` (the few
// instances still running that late are the ill-conditioned ones that need longer inner solves).
static int cycle_cg_ticks(int c, int base, int grow_after, int grow_every, int max_cg) {
  if (c < grow_after) return base;
  const int k = 1 + (c - grow_after) / grow_every;
  long long n = (long long)base << std::min(k, 20);
  return (int)std::min<long long>(n, max_cg);
}

// Reduced operator B (CSR) and its transpose from the factors (assemble.cuh + a stable radix sort by column).
static int assemble_reduced(ScoreHandle_ *h, cudaStream_t st) {
  DevProblem &P = h->P;
  int end_bit = 1;
  while ((1ll << end_bit) <= (long long)P.nz) ++end_bit;
  if (!P.indptr) {  // first use: the pair, the transpose's scratch (arena: stream-ordered, synchronised by get_chunk)
    int rc;
    if ((rc = dalloc(h, &P.indptr, (size_t)P.m + 1))) return rc;
    if ((rc = dalloc(h, &P.cols, (size_t)P.nnz))) return rc;
    if ((rc = dalloc(h, &P.vals, (size_t)P.nnz))) return rc;
    if ((rc = dalloc(h, &P.t_indptr, (size_t)P.nz + 1))) return rc;
    if ((rc = dalloc(h, &P.t_rows, (size_t)P.nnz))) return rc;
    if ((rc = dalloc(h, &P.t_vals, (size_t)P.nnz))) return rc;
    if ((rc = dalloc(h, &h->nnz_row, (size_t)P.nnz))) return rc;
    if ((rc = dalloc(h, &h->csr_keys, (size_t)P.nnz))) return rc;
    if ((rc = dalloc(h, &h->csr_idx, (size_t)P.nnz))) return rc;
    if ((rc = dalloc(h, &h->csr_perm, (size_t)P.nnz))) return rc;
    {
      // the size query walks through the runtime (device / kernel attribute look-ups); cache it per (nnz, end_bit)
      static std::mutex mu;
      static std::map<std::pair<long long, int>, size_t> known;
      std::lock_guard<std::mutex> lk(mu);
      const auto key = std::make_pair((long long)P.nnz, end_bit);
      auto it = known.find(key);
      if (it == known.end()) {
        size_t bytes = 0;
        SCORE_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, bytes, P.cols, h->csr_keys, h->csr_idx, h->csr_perm, P.nnz, 0,
                                                         end_bit, (cudaStream_t)0));
        it = known.emplace(key, bytes).first;
      }
      h->sort_tmp_bytes = it->second;
    }
    char *tmp = nullptr;
    if ((rc = dalloc(h, &tmp, h->sort_tmp_bytes))) return rc;
    h->sort_tmp = tmp;
  }
  AsmOut out{P.indptr, P.cols, P.vals, P.w, P.b, h->nnz_row};
  const long nf = (long)P.E + P.K + P.Lp;
  k_assemble<<<grid_for(nf, 256), 256, 0, st>>>(P, ASM_REDUCED, 0, P.n_inst, out);
  // transpose of the rows this rank owns (all rows on a single GPU): entries [nnz_lo, nnz_hi)
  int nnz_lo = 0, nnz_hi = P.nnz;
  if (h->n_ranks > 1) {
    const int row_lo = h->W.rb_lo * kRowsPerBlock, row_hi = std::min(P.m, h->W.rb_hi * kRowsPerBlock);
    SCORE_CUDA_CHECK(cudaMemcpyAsync(&nnz_lo, P.indptr + row_lo, sizeof(int), cudaMemcpyDeviceToHost, st));
    SCORE_CUDA_CHECK(cudaMemcpyAsync(&nnz_hi, P.indptr + row_hi, sizeof(int), cudaMemcpyDeviceToHost, st));
    SCORE_CUDA_CHECK(cudaStreamSynchronize(st));
  }
  const int nloc = nnz_hi - nnz_lo;
  k_iota<<<grid_for(std::max(nloc, 1), 256), 256, 0, st>>>(h->csr_idx, nloc);
  SCORE_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(h->sort_tmp, h->sort_tmp_bytes, P.cols + nnz_lo, h->csr_keys, h->csr_idx,
                                                   h->csr_perm, nloc, 0, end_bit, st));
  SCORE_CUDA_CHECK(cudaMemsetAsync(P.t_indptr, 0, sizeof(int) * (P.nz + 1), st));
  if (nloc > 0)
    k_transpose_fill<<<grid_for(nloc, 256), 256, 0, st>>>(nloc, P.nz, h->csr_keys, h->csr_perm, h->nnz_row + nnz_lo,
                                                          P.vals + nnz_lo, P.t_indptr, P.t_rows, P.t_vals);
  h->csr_valid = true;
  return SCORE_OK;
}

extern "C" int score_solve(ScoreHandle h, const ScoreParams *params, ScoreStats *stats, ScoreInstanceStats *inst_stats) {
  if (!h) {
    g_score_last_error = "null handle";
    return SCORE_ERR_INVALID;
  }
  ScoreParams prm{};
  if (params) prm = *params;
  SolverCfg cfg{};
  cfg.max_newton = prm.max_newton > 0 ? prm.max_newton : 200;
  if (params && prm.max_newton == -1) cfg.max_newton = 0;  // "evaluate the start point only"
  cfg.max_cg = prm.max_cg > 0 ? prm.max_cg : 100;
  cfg.kkt_tol = prm.kkt_tol > 0 ? prm.kkt_tol : 1e-6;
  cfg.forcing = prm.cg_forcing > 0 ? prm.cg_forcing : 0.2;
  cfg.mu0 = prm.mu0 > 0 ? prm.mu0 : (prm.mu0 < 0 ? 0.0 : 0.1);
  cfg.mu_factor = (prm.mu_factor > 0 && prm.mu_factor < 1) ? prm.mu_factor : 0.1;
  cfg.center_tol = prm.center_tol > 0 ? prm.center_tol : 16.0;
  cfg.center_tol_late = prm.center_tol_late > 0 ? prm.center_tol_late : 1.0;
  cfg.mu_min = prm.mu_min > 0 ? prm.mu_min : 1e-16;
  cfg.mu_eval = 1e-5;
  cfg.coarse_reg = 1e-6;
  cfg.coarse_every = prm.coarse_every > 0 ? prm.coarse_every : 1;
  // PCG operator: matrix-free (default) unless the caller asks for the assembled CSR pair, the solve is row-partitioned
  // over several GPUs (the all-reduce works on the CSR column pass) or the fused per-instance kernel is requested
  const bool mf = prm.operator_mode == 0 && h->n_ranks == 1 && prm.tail_threshold <= 0;
  cfg.pad = mf ? 1 : 0;  // part of the graph key: captured kernel parameters carry the mode
  h->V.mf = mf ? 1 : 0;
  const int max_ticks = prm.max_ticks > 0 ? prm.max_ticks : 200000;
  int rc;
  SCORE_CUDA_CHECK(cudaSetDevice(h->device));
  cudaStream_t st = prm.stream ? (cudaStream_t)prm.stream : h->own_stream;
  DevProblem &P = h->P;
  SolverVecs &V = h->V;
  const int d = P.d;
  long launches = 0;

  if (prm.verbose >= 2 && !V.trace) {  // diagnostic per-Newton-step trace (score_get_internal SCORE_INT_TRACE)
    V.trace_cap = 256;
    if ((rc = dalloc(h, &V.trace, (size_t)P.n_inst * V.trace_cap * kTraceRec))) return rc;
    for (auto &kv : h->graphs) g_cache.retire_graph(kv.second);  // captured kernel parameters hold the old V
    h->graphs.clear();
  }
  if (V.trace) SCORE_CUDA_CHECK(cudaMemsetAsync(V.trace, 0, sizeof(double) * (size_t)P.n_inst * V.trace_cap * kTraceRec, st));

  // phase-timing events: from the process-wide cache, returned on every exit path
  struct PhaseEvents {
    cudaEvent_t e[5] = {};
    int dev = 0;
    ~PhaseEvents() {
      for (auto x : e)
        if (x) g_cache.put_tevent(dev, x);
    }
  } pev;
  pev.dev = h->device;
  cudaEvent_t *ev = pev.e;
  for (int i = 0; i < 5; ++i) SCORE_CUDA_CHECK(g_cache.get_tevent(h->device, &ev[i]));
  h->last_stream = st;
  SCORE_CUDA_CHECK(cudaEventRecord(ev[0], st));

  // ---- 1. assembly: reduced operator and its transpose (the matrix-free solve never reads them: built on demand by
  //         score_get_csr), row weights, incidence lists of the factor-wise operator
  {
    if (!mf) {
      if ((rc = assemble_reduced(h, st))) return rc;
      launches += 6;
    } else {
      h->csr_valid = false;
      k_row_weights<<<grid_for((long)P.E + P.K + P.Lp, 256), 256, 0, st>>>(P, P.w, P.b);
      launches += 1;
    }
    if (P.n_inc > 0) {
      // incidence lists of the matrix-free operator: (owner, factor) pairs in factor order, stable sort by owner
      // (the transpose's index scratch is free again)
      k_inc_fill<<<grid_for((long)P.E + P.K + P.Lp, 256), 256, 0, st>>>(P, h->sort_idx, h->inc_in);
      int ob = 1;
      while ((1ll << ob) <= (long long)P.P + P.L + 1) ++ob;
      SCORE_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(h->inc_tmp, h->inc_tmp_bytes, h->sort_idx, h->sort_keys, h->inc_in,
                                                       P.inc_rec, P.n_inc, 0, ob, st));
      k_inc_ptr<<<grid_for(P.n_inc, 256), 256, 0, st>>>(P.n_inc, P.P + P.L, h->sort_keys, P.inc_ptr);
      k_hv_fill<<<grid_for(h->T.n_pb + h->n_pbd, 256), 256, 0, st>>>(P, h->T, h->n_pbd);
      launches += 5;
    } else {
      SCORE_CUDA_CHECK(cudaMemsetAsync(P.inc_ptr, 0, sizeof(int) * ((size_t)P.P + P.L + 1), st));
    }
  }
  SCORE_CUDA_CHECK(cudaEventRecord(ev[1], st));
  // ---- 2. preconditioner, start point, solver state
  {
    k_dead_reckon<<<grid_for(P.n_seg, 64), 64, 0, st>>>(P);
    if (mf)
      k_diag_setup_mf<<<grid_for((long)P.P + P.L, 256), 256, 0, st>>>(P, h->wsum);
    else
      k_diag_setup<<<grid_for((long)P.P + (long)P.L * d, 256), 256, 0, st>>>(P, h->wsum);
    if (h->n_ranks > 1) {  // range weights were summed over the local rows only
      g_nccl.AllReduce(h->wsum, h->wsum, P.P, ncclDouble, ncclSum, h->comm, st);
      if (P.L > 0) g_nccl.AllReduce(P.lm_inv, P.lm_inv, (size_t)P.L * d, ncclDouble, ncclSum, h->comm, st);
    }
    if (P.L > 0) k_lm_finish<<<grid_for((long)P.L * d, 256), 256, 0, st>>>(P);
    k_build_M<<<grid_for(P.P, 128), 128, 0, st>>>(P, h->wsum);
    k_init_z<<<grid_for(P.nz, 256), 256, 0, st>>>(P, V.z);
    if (!mf)
      k_residual<<<grid_for(P.m, kThreads), kThreads, 0, st>>>(P, V.z, V.res);
    else if (d == 2)
      k_rows_mf_all<2><<<std::max(1, std::min(h->T.n_rb, h->n_sm * 16)), kThreads, 0, st>>>(P, V, h->T, h->st, RM_RES);
    else
      k_rows_mf_all<3><<<std::max(1, std::min(h->T.n_rb, h->n_sm * 16)), kThreads, 0, st>>>(P, V, h->T, h->st, RM_RES);
    if ((h->c_nmax > 0 || !h->big.empty()) && std::max(P.c_ninc, P.c_npair) > 0) {
      if (d == 2)
        k_coarse_static<2><<<grid_for(std::max(P.c_ninc, P.c_npair), 256), 256, 0, st>>>(P);
      else
        k_coarse_static<3><<<grid_for(std::max(P.c_ninc, P.c_npair), 256), 256, 0, st>>>(P);
    }
    launches += 6;
    if (h->c_nmax > 0 && (rc = coarse_build_prepare(h))) return rc;
    for (double *v : {V.dz, V.r, V.s, V.p, V.ytmp})
      SCORE_CUDA_CHECK(cudaMemsetAsync(v, 0, sizeof(double) * P.nz, st));
    SCORE_CUDA_CHECK(cudaMemsetAsync(V.u, 0, sizeof(double) * P.m, st));
    SCORE_CUDA_CHECK(cudaMemsetAsync(V.bdz, 0, sizeof(double) * P.m, st));
    SCORE_CUDA_CHECK(cudaMemsetAsync(V.part_col, 0, sizeof(double) * 4 * h->T.n_cb, st));
    if (h->n_ranks > 1) {
      // rows of other ranks contribute exact zeros to every all-reduce
      SCORE_CUDA_CHECK(cudaMemsetAsync(h->red_send, 0, sizeof(double) * h->red_count, st));
      SCORE_CUDA_CHECK(cudaMemsetAsync(V.part_ls, 0, sizeof(double) * (size_t)h->T.n_rb * kLsSums, st));
      SCORE_CUDA_CHECK(cudaMemsetAsync(V.mk, 0, sizeof(double) * (size_t)P.K * (d * (d + 1) / 2), st));
    }
    std::vector<InstState> init(P.n_inst);
    memset(init.data(), 0, sizeof(InstState) * P.n_inst);
    for (auto &s : init) {
      s.phase = PH_LS;
      s.skip_ls = 1;
      s.eta = cfg.forcing;
      s.mu = s.mu_ls = cfg.mu0;
      s.mu_c = -1.0;
    }
    SCORE_CUDA_CHECK(cudaMemcpyAsync(h->st, init.data(), sizeof(InstState) * P.n_inst, cudaMemcpyHostToDevice, st));
    SCORE_CUDA_CHECK(cudaMemsetAsync(h->d_ndone, 0, sizeof(int), st));
    SCORE_CUDA_CHECK(cudaMemsetAsync(h->bar_mem, 0, sizeof(int) * 2 * kMaxFusedGroups, st));
    {
      // every instance starts in the line-search phase: run[0] = ls[0] = all instances, parity 0
      std::vector<int> wl(16 + 2 * (size_t)P.n_inst, 0);
      wl[2 + 0 * 3 + WL_RUN] = P.n_inst;
      wl[2 + 0 * 3 + WL_LS] = P.n_inst;
      for (int i = 0; i < P.n_inst; ++i) wl[16 + i] = wl[16 + P.n_inst + i] = i;
      SCORE_CUDA_CHECK(cudaMemcpyAsync(h->wl_mem, wl.data(), sizeof(int) * wl.size(), cudaMemcpyHostToDevice, st));
      SCORE_CUDA_CHECK(cudaStreamSynchronize(st));
    }
    SCORE_CUDA_CHECK(cudaStreamSynchronize(st));  // init vector is on the host stack
  }
  SCORE_CUDA_CHECK(cudaEventRecord(ev[2], st));
  // ---- 3. solver cycles (one CUDA graph per distinct cycle length, replayed until every instance is done)
  if (memcmp(&h->graph_cfg, &cfg, sizeof(cfg)) != 0) {
    for (auto &kv : h->graphs) cudaGraphExecDestroy(kv.second);
    h->graphs.clear();
    h->graph_cfg = cfg;
  }
  // key: PCG ticks of the cycle (< 0: tail cycle with the fused kernel), + 1000 when captured on the high-priority stream
  // (kernel nodes keep the priority of the stream they were captured on)
  auto cycle_graph = [&](int key, cudaGraphExec_t *out) -> int {
    const int n_cg = key >= 500 ? key - 1000 : key;
    auto it = h->graphs.find(key);
    if (it != h->graphs.end()) {
      *out = it->second;
      return SCORE_OK;
    }
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    SCORE_CUDA_CHECK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    const int nk = n_cg < 0 ? launch_tail_cycle_d(h, cfg, st) : launch_cycle_d(h, cfg, st, n_cg);
    // a failed launch inside the capture surfaces here; the capture is always ended so the stream stays usable
    cudaError_t ce = cudaGetLastError();
    const cudaError_t ee = cudaStreamEndCapture(st, &graph);
    if (ce == cudaSuccess) ce = ee;
    if (ce == cudaSuccess) ce = cudaGraphInstantiate(&exec, graph, 0);
    if (graph) cudaGraphDestroy(graph);
    if (ce != cudaSuccess) {
      g_score_last_error = std::string("capturing the solver cycle failed: ") + cudaGetErrorString(ce);
      return SCORE_ERR_CUDA;
    }
    h->graph_kernels[key] = nk;
    h->graphs[key] = exec;
    *out = exec;
    return SCORE_OK;
  };
  const int cg_base = prm.cg_per_cycle > 0 ? prm.cg_per_cycle : 4;
  const int grow_after = prm.cg_grow_after > 0 ? prm.cg_grow_after : (1 << 30);
  const int grow_every = prm.cg_grow_every > 0 ? prm.cg_grow_every : 8;
  long ticks = 0, cycles = 0, tail_cycles = 0;
  double kernel_ms[kStatSlots] = {0}, kernel_ms_full[kStatSlots] = {0};
  long long kernel_count[kStatSlots] = {0}, kernel_count_full[kStatSlots] = {0};
  long profiled = 0;
  const int prof_skip = prm.profile_cycles > 0 ? std::max(0, prm.profile_skip) : 0;
  const int prof_end = prm.profile_cycles > 0 ? prof_skip + prm.profile_cycles : 0;
  TickProfiler pf;
  pf.st = st;
  pf.device = h->device;
  h->h_ndone[0] = h->h_ndone[1] = 0;
  if (h->c_big || h->n_ranks > 1) {
    // large single instance: cuSOLVER / NCCL are not captured into graphs; cycles are launched directly and the host
    // decides per cycle whether the line-search tick will rebuild the coarse level (the instance is in PH_LS)
    const int n_cg = cg_base;
    while (ticks < max_ticks) {
      InstState s0;
      SCORE_CUDA_CHECK(cudaMemcpyAsync(&s0, h->st, sizeof(InstState), cudaMemcpyDeviceToHost, st));
      SCORE_CUDA_CHECK(cudaStreamSynchronize(st));
      if (s0.phase == PH_DONE) break;
      const bool prof = cycles >= prof_skip && cycles < prof_end;
      launches += launch_cycle_d(h, cfg, st, n_cg, prof ? &pf : nullptr, s0.phase == PH_LS);
      if (prof) profiled += 1;
      ticks += 1 + n_cg;
      cycles += 1;
    }
  }
  // tail mode (opt-in, ScoreParams.tail_threshold): once at most `tail_thresh` instances are unfinished (known one
  // cycle late: the completion count is read back asynchronously), cycles run the fused per-instance PCG kernel
  // instead of lockstep PCG ticks.  Measured (profiles/fused_pcg_r2.txt): bit-identical, but at 8 CTAs per instance an
  // iteration takes 60-80 us — no faster than a lockstep tick of a nearly empty batch (58 us) — so it is off by default.
  const int tail_thresh = prm.tail_threshold > 0 ? prm.tail_threshold : -1;
  const bool tail_ok = h->big.empty() && !coarse_apply_split();
  int last_done = 0;
  // Sparse last cycles at high stream priority: once at most `hi_thresh` instances are unfinished the batch no longer
  // fills the GPU and every tick costs the latency of its dependent kernels.  When another handle is solving at the same
  // time (sub-batches of a sweep on their own streams) those small kernels would queue behind the other handle's large
  // ones block by block; on the high-priority stream their blocks are dispatched first, so one handle's tail runs UNDER
  // the other handle's dense cycles instead of being stretched by them.  (Library-owned stream only.)
  // Measured (profiles/pipeline_probe_r2.txt): no effect — what stretches a tail under another handle's dense cycles is the
  // loaded memory latency of its dependent loads, not block dispatch order — so it is opt-in.
  const int hi_thresh = (prm.stream || prm.hi_prio_threshold <= 0) ? -1 : prm.hi_prio_threshold;
  bool on_hi = false;
  while (!(h->c_big || h->n_ranks > 1) && ticks < max_ticks) {
    const bool tail = tail_ok && P.n_inst - last_done <= tail_thresh;
    const int n_cg = tail ? -1 : cycle_cg_ticks((int)cycles, cg_base, grow_after, grow_every, cfg.max_cg);
    if (!on_hi && hi_thresh > 0 && cycles >= 2 && P.n_inst - last_done <= hi_thresh && P.n_inst > hi_thresh) {
      SCORE_CUDA_CHECK(cudaEventRecord(h->ev_switch, st));
      SCORE_CUDA_CHECK(cudaStreamWaitEvent(h->hi_stream, h->ev_switch, 0));
      st = h->hi_stream;
      h->last_stream = st;
      pf.st = st;
      on_hi = true;
    }
    const bool prof = cycles >= prof_skip && cycles < prof_end;
    if (prof) {
      launches += tail ? launch_tail_cycle_d(h, cfg, st, &pf) : launch_cycle_d(h, cfg, st, n_cg, &pf);
      profiled += 1;
    } else {
      cudaGraphExec_t exec;
      const int key = n_cg + (on_hi ? 1000 : 0);
      if ((rc = cycle_graph(key, &exec))) return rc;
      SCORE_CUDA_CHECK(cudaGraphLaunch(exec, st));
      launches += h->graph_kernels[key];
    }
    ticks += tail ? 2 : 1 + n_cg;
    tail_cycles += tail ? 1 : 0;
    const int slot = (int)(cycles & 1);
    SCORE_CUDA_CHECK(cudaMemcpyAsync(&h->h_ndone[slot], h->d_ndone, sizeof(int), cudaMemcpyDeviceToHost, st));
    SCORE_CUDA_CHECK(cudaEventRecord(h->ev_done[slot], st));
    cycles += 1;
    // the completion count of the previous cycle is read while this one runs (no host bubble between cycles)
    if (cycles >= 2) {
      SCORE_CUDA_CHECK(wait_event_hybrid(h->ev_done[slot ^ 1]));
      last_done = h->h_ndone[slot ^ 1];
      if (last_done >= P.n_inst) break;
    }
    if (prof && cycles == prof_end) {
      SCORE_CUDA_CHECK(cudaStreamSynchronize(st));
      pf.collect(kernel_ms, kernel_count, kernel_ms_full, kernel_count_full);
    }
  }
  SCORE_CUDA_CHECK(cudaStreamSynchronize(st));
  if (!pf.ev.empty()) pf.collect(kernel_ms, kernel_count, kernel_ms_full, kernel_count_full);
  SCORE_CUDA_CHECK(cudaEventRecord(ev[3], st));
  // ---- 4. extraction
  k_split_z<<<grid_for(P.nz, 256), 256, 0, st>>>(P, V.z, h->out_poses, h->out_lms);
  k_round_so<<<grid_for(P.P, 128), 128, 0, st>>>(d, P.P, h->out_poses, P.blk, d + 1, h->out_round);
  if (P.K > 0) k_distances<<<grid_for(P.K, 256), 256, 0, st>>>(P, V.z, h->st, h->out_dist);
  launches += 3;
  SCORE_CUDA_CHECK(cudaEventRecord(ev[4], st));
  SCORE_CUDA_CHECK(cudaStreamSynchronize(st));
  SCORE_CUDA_CHECK(cudaGetLastError());
  h->solved_once = true;

  std::vector<InstState> fin(P.n_inst);
  SCORE_CUDA_CHECK(cudaMemcpy(fin.data(), h->st, sizeof(InstState) * P.n_inst, cudaMemcpyDeviceToHost));
  int n_solved = 0;
  double bytes = 0.0, kbytes_total[kStatSlots] = {0}, kbytes_launch[kStatSlots] = {0};
  for (int i = 0; i < P.n_inst; ++i) {
    const InstState &s = fin[i];
    n_solved += (s.phase == PH_DONE && s.solved) ? 1 : 0;
    InstDims D;
    D.d = d;
    D.nnz = h->nnzoff[i + 1] - h->nnzoff[i];
    D.m = h->roff[i + 1] - h->roff[i];
    D.nz = h->zoff[i + 1] - h->zoff[i];
    D.K = h->rng_off[i + 1] - h->rng_off[i];
    D.P = h->pose_off[i + 1] - h->pose_off[i];
    D.nc = h->c_n[i];
    D.E = h->edge_off[i + 1] - h->edge_off[i];
    D.L = h->lm_off[i + 1] - h->lm_off[i];
    D.Lp = h->prior_off[i + 1] - h->prior_off[i];
    D.mf = mf;
    D.fused = h->c_nmax > 0 && !coarse_apply_split();
    double fused_iter = 0.0;  // bytes of one PCG iteration = the PCG-tick bytes of all its kernels
    for (int k = 0; k < kNumKernels; ++k) {
      if (k == KI_PCG_FUSED) continue;
      const double bcg = kernel_bytes_inst(k, TM_CG, D), bls = kernel_bytes_inst(k, TM_LS, D);
      // the first line-search tick only evaluates the start point (no direction yet)
      const double n_ls = (k == KI_ROWPASS || k == KI_LINESEARCH) ? s.newton_it : s.newton_it + 1.0;
      // iterations done inside the fused kernel are booked on the fused kernel, not on the stand-alone ones
      const double tot = (s.total_cg - s.fused_cg) * bcg + n_ls * bls + s.n_eval * kernel_bytes_inst(k, TM_EVAL, D);
      kbytes_total[k] += tot;
      bytes += tot;
      kbytes_launch[k] += (bcg > 0.0) ? bcg : bls;
      fused_iter += bcg;
    }
    kbytes_total[KI_PCG_FUSED] += s.fused_cg * fused_iter;
    bytes += s.fused_cg * fused_iter;
    kbytes_launch[KI_PCG_FUSED] += fused_iter;
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
    stats->cycles = cycles;
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
    stats->profiled_cycles = profiled;
    for (int k = 0; k < kStatSlots; ++k) {
      stats->kernel_ms[k] = kernel_ms[k];
      stats->kernel_count[k] = kernel_count[k];
      stats->kernel_bytes[k] = kbytes_launch[k];
      stats->kernel_bytes_total[k] = kbytes_total[k];
      stats->kernel_ms_full[k] = kernel_ms_full[k];
      stats->kernel_count_full[k] = kernel_count_full[k];
    }
  }
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
    DevBuf bip, bc, bv, bw, bb;
    SCORE_CUDA_CHECK(bip.alloc(sizeof(int) * (rows + 1)));
    SCORE_CUDA_CHECK(bc.alloc(sizeof(int) * (nn + 1)));
    SCORE_CUDA_CHECK(bv.alloc(sizeof(double) * (nn + 1)));
    SCORE_CUDA_CHECK(bw.alloc(sizeof(double) * (rows + 1)));
    SCORE_CUDA_CHECK(bb.alloc(sizeof(double) * (rows + 1)));
    int *dip = bip.as<int>(), *dc = bc.as<int>();
    double *dv = bv.as<double>(), *dw = bw.as<double>(), *db = bb.as<double>();
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
  if (!h->csr_valid) {  // a matrix-free solve does not assemble the pair: do it now
    int rc = assemble_reduced(h, h->own_stream);
    if (rc) return rc;
    SCORE_CUDA_CHECK(cudaStreamSynchronize(h->own_stream));
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

extern "C" int score_nccl_unique_id(char *out128) {
  if (!out128) {
    g_score_last_error = "null argument";
    return SCORE_ERR_INVALID;
  }
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  if (!nccl_load()) return SCORE_ERR_CUDA;
  ncclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != ncclSuccess) {
    g_score_last_error = "ncclGetUniqueId failed";
    return SCORE_ERR_CUDA;
  }
  memcpy(out128, &id, 128);
  return SCORE_OK;
}

extern "C" int score_comm_init(ScoreHandle h, int32_t n_ranks, int32_t rank, const char *id128) {
  if (!h || !id128 || n_ranks < 1 || rank < 0 || rank >= n_ranks) {
    g_score_last_error = "bad argument";
    return SCORE_ERR_INVALID;
  }
  DevProblem &P = h->P;
  if (P.n_inst != 1) {
    g_score_last_error = "row partitioning applies to a single instance (batches are sharded by instance)";
    return SCORE_ERR_INVALID;
  }
  if (h->comm) {
    g_score_last_error = "communicator already initialised";
    return SCORE_ERR_STATE;
  }
  SCORE_CUDA_CHECK(cudaSetDevice(h->device));
  if (n_ranks == 1) return SCORE_OK;
  if (!nccl_load()) return SCORE_ERR_CUDA;
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  if (g_nccl.CommInitRank(&h->comm, n_ranks, id, rank) != ncclSuccess) {
    g_score_last_error = "ncclCommInitRank failed";
    h->comm = nullptr;
    return SCORE_ERR_CUDA;
  }
  h->n_ranks = n_ranks;
  h->rank = rank;
  const int n_rb = h->T.n_rb, d = P.d;
  h->W.rb_lo = (int)((long long)rank * n_rb / n_ranks);
  h->W.rb_hi = (int)((long long)(rank + 1) * n_rb / n_ranks);
  int rc;
  if ((rc = dalloc(h, &h->red_recv, h->red_count))) return rc;
  if ((rc = dalloc(h, &h->ls_recv, (size_t)n_rb * kLsSums))) return rc;
  if ((rc = dalloc(h, &h->mk_recv, (size_t)P.K * (d * (d + 1) / 2)))) return rc;
  h->Vg = h->V;
  h->Vg.part_row = h->red_recv;
  h->Vg.part_upd = h->red_recv + n_rb;
  h->Vg.hglob = h->red_recv + 5 * (size_t)n_rb;
  h->Vg.part_ls = h->ls_recv;
  h->Vg.mk = h->mk_recv;
  SCORE_CUDA_CHECK(cudaDeviceSynchronize());
  return SCORE_OK;
}

extern "C" int score_get_internal(ScoreHandle h, int32_t which, int32_t inst, double *out, int64_t capacity,
                                  int64_t *count) {
  if (!h) {
    g_score_last_error = "null handle";
    return SCORE_ERR_INVALID;
  }
  const DevProblem &P = h->P;
  if (inst < 0 || inst >= P.n_inst) {
    g_score_last_error = "instance index out of range";
    return SCORE_ERR_INVALID;
  }
  if (!h->solved_once) {
    g_score_last_error = "solver internals exist only after score_solve";
    return SCORE_ERR_STATE;
  }
  SCORE_CUDA_CHECK(cudaSetDevice(h->device));
  const double *src = nullptr;
  int64_t n = 0;
  const int nm = P.d * (P.d + 1) / 2;
  switch (which) {
    case SCORE_INT_COARSE_INV:
      n = (int64_t)h->c_n[inst] * h->c_n[inst];
      for (size_t bi = 0; bi < h->big.size(); ++bi)  // landmark block eliminated: what is stored is T^-1 = (A_c^-1)_ss
        if (h->big[bi] == inst && h->bigx[bi].schur) n = (int64_t)h->c_nb[inst] * h->c_nb[inst];
      src = P.c_Ainv + h->c_moff[inst];
      break;
    case SCORE_INT_RANGE_CURV:
      n = (int64_t)(h->rng_off[inst + 1] - h->rng_off[inst]) * nm;
      src = h->V.mk + (size_t)h->rng_off[inst] * nm;
      break;
    case SCORE_INT_FRAMES:
      n = (int64_t)(h->pose_off[inst + 1] - h->pose_off[inst]) * P.blk;
      src = P.G + (size_t)h->pose_off[inst] * P.blk;
      break;
    case SCORE_INT_TRACE:
      n = h->V.trace ? (int64_t)h->V.trace_cap * kTraceRec : 0;
      src = h->V.trace ? h->V.trace + (size_t)inst * h->V.trace_cap * kTraceRec : nullptr;
      break;
    case SCORE_INT_REF_GRAD:
    case SCORE_INT_REF_DIR:
    case SCORE_INT_REF_HDIR:
    case SCORE_INT_REF_DIAG:
      if (!h->refined) {
        g_score_last_error = "refinement internals exist only after score_refine";
        return SCORE_ERR_STATE;
      }
      n = h->zoff[inst + 1] - h->zoff[inst];
      src = (which == SCORE_INT_REF_GRAD ? h->R.g : which == SCORE_INT_REF_DIR ? h->R.p : which == SCORE_INT_REF_HDIR ? h->R.q : h->R.dg) +
            h->zoff[inst];
      break;
    default:
      g_score_last_error = "unknown internal array selector";
      return SCORE_ERR_INVALID;
  }
  if (count) *count = n;
  if (!out) return SCORE_OK;
  if (capacity < n) {
    g_score_last_error = "buffer too small";
    return SCORE_ERR_INVALID;
  }
  if (n) SCORE_CUDA_CHECK(cudaMemcpy(out, src, sizeof(double) * n, cudaMemcpyDefault));
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
  DevBuf bin, bout;
  const size_t bytes = sizeof(double) * (size_t)n * dim * dim;
  SCORE_CUDA_CHECK(bin.alloc(bytes));
  SCORE_CUDA_CHECK(bout.alloc(bytes));
  double *din = bin.as<double>(), *dout = bout.as<double>();
  SCORE_CUDA_CHECK(cudaMemcpy(din, mats, bytes, cudaMemcpyDefault));
  k_round_so<<<grid_for(n, 128), 128>>>(dim, n, din, dim * dim, dim, dout);
  SCORE_CUDA_CHECK(cudaGetLastError());
  SCORE_CUDA_CHECK(cudaMemcpy(out, dout, bytes, cudaMemcpyDefault));
  return SCORE_OK;
}
