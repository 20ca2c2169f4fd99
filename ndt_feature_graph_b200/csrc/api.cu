// api.cu — the extern "C" ABI of include/ndtb.h: contexts, HBM-resident maps, batched map build and
// batched registration.  Host code only orchestrates (allocation, job tables, launches); all arithmetic
// is in map_build.cu (kernel i) and d2d.cu (kernel ii + optimiser).  There is no CPU compute path here.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <unordered_map>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/ndtb.h"
#include "engine.cuh"
#include "map_build.cuh"

namespace ndtb {
// d2d.cu
size_t match_smem_bytes(int table_entries);
size_t opt_state_bytes();
cudaError_t match_kernel_prepare(int smem_optin_bytes);
cudaError_t launch_match(const MatchJob *d_jobs, const int *d_job_ids, int n_slots, int cluster, const MatchConfig &cfg,
                         ndtb_result *d_out, void *d_states, int resume, int pass_budget, int *d_unfinished, int *d_yielded,
                         cudaStream_t stream);
cudaError_t launch_derivatives(const MatchJob *d_job, const MatchConfig &cfg, bool hess, int n_ctas, double *d_partial,
                               double *d_out29, cudaStream_t stream);
cudaError_t launch_covariance(const MatchJob *d_jobs, int n_jobs, const MatchConfig &cfg, const ndtb_result *d_res,
                              const long long *d_gt_off, double *d_gt, double *d_partial, int n_chunks, double *d_cov36,
                              int *d_status, const int *d_yielded, int mode, cudaStream_t stream);
int cov_partial_width();
int acc_total();
}  // namespace ndtb

using namespace ndtb;

struct ndtb_ctx {
  int device = 0;
  cudaMemPool_t pool = nullptr;  // the context's own stream-ordered pool: no reuse dependencies between contexts
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int sm_count = 0;
  int smem_optin = 0;
  int64_t launches = 0;
  std::string last_error;
  bool timing = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timed;  // event pairs around match-kernel launches
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timed_build;  // event pairs around batched map builds
  // host-buffer path: scans are uploaded on a second stream in chunks while the previous chunk's maps are being built
  cudaStream_t copy_stream = nullptr;
  std::vector<cudaEvent_t> copy_events;
  // batched match: the covariance of the registrations that finished in the first launch runs on this stream while the
  // second launch (stragglers, about half of the SMs) is still working
  cudaStream_t aux_stream = nullptr;
  cudaEvent_t aux_ev[2] = {nullptr, nullptr};
  // small host->device uploads (job descriptors, poses, single scans) bypass the copy engine: see h2d_small
  static constexpr int STG_SEGS = 8;
  static constexpr size_t STG_SEG_BYTES = (size_t)4 << 20, STG_MAX = (size_t)1 << 20;
  // slab cache: see slab_alloc
  struct CachedSlab {
    char *p;
    uint64_t stamp;
  };
  std::mutex cache_mu;
  std::unordered_map<size_t, std::vector<CachedSlab>> slab_cache;  // by capacity class
  size_t cached_bytes = 0, cache_limit = (size_t)16 << 30;
  uint64_t cache_clock = 0;
  char *stg_host = nullptr, *stg_dev = nullptr;
  cudaEvent_t stg_ev[STG_SEGS] = {};
  bool stg_used[STG_SEGS] = {};
  int stg_seg = 0;
  size_t stg_head = 0;
};

#define CU_TRY(ctx, expr)                                                                           \
  do {                                                                                              \
    cudaError_t e__ = (expr);                                                                       \
    if (e__ != cudaSuccess) {                                                                       \
      (ctx)->last_error = std::string(#expr) + ": " + cudaGetErrorString(e__);                      \
      return NDTB_ERR_CUDA;                                                                         \
    }                                                                                               \
  } while (0)

// CUDA's current device is per host thread: a context created on device 1 and then driven from a fresh thread (whose
// current device is 0) would create its events and slabs on the wrong device.  Every entry point that touches CUDA
// therefore switches to the context's device and restores the caller's on return.
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(const ndtb_ctx *ctx) {
    if (!ctx) return;
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != ctx->device) cudaSetDevice(ctx->device);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

namespace {

// NDTB_SLOWLOG=<ms>: calls of ndtb_register_scans slower than that print where their host time went (stderr)
struct SlowLog {
  double alloc_ms = 0, sync_ms = 0, up_ms = 0, enq_ms = 0;
  int allocs = 0, syncs = 0;
};
thread_local SlowLog g_slow;
inline double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
inline cudaError_t timed_sync(cudaStream_t st) {
  const double t0 = now_ms();
  const cudaError_t e = cudaStreamSynchronize(st);
  g_slow.sync_ms += now_ms() - t0, g_slow.syncs++;
  return e;
}

// ---- small uploads without the copy engine -----------------------------------------------------------------------
// The H2D copy engine serves its queue in submission order ACROSS streams: a 470 KB job-descriptor upload issued while
// 1.9 GB of scans are queued on the copy stream (this context's or another lane's) waits for all of them — measured on the
// B200 box: the first chunk's build started 34 ms late, after the last scan had arrived.  Uploads of up to 1 MB therefore go
// through a pinned, device-mapped staging ring and a copy kernel on the consuming stream, which depends on nothing but
// that stream.  A ring segment is reused only after the event behind its last copy kernel has completed.
__global__ void k_upload16(const uint4 *__restrict__ src, uint4 *__restrict__ dst, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
__global__ void k_upload1(const unsigned char *__restrict__ src, unsigned char *__restrict__ dst, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

cudaError_t h2d_small(ndtb_ctx *ctx, void *dst, const void *src, size_t bytes, cudaStream_t st) {
  if (bytes == 0) return cudaSuccess;
  static const bool no_staging = std::getenv("NDTB_NO_STAGING") != nullptr;  // A/B switch for measurements
  if (no_staging || bytes > ndtb_ctx::STG_MAX || st != ctx->stream) return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st);
  if (!ctx->stg_host) {
    cudaError_t e = cudaHostAlloc((void **)&ctx->stg_host, ndtb_ctx::STG_SEGS * ndtb_ctx::STG_SEG_BYTES, cudaHostAllocMapped);
    if (e != cudaSuccess) return e;
    if ((e = cudaHostGetDevicePointer((void **)&ctx->stg_dev, ctx->stg_host, 0)) != cudaSuccess) return e;
    for (auto &ev : ctx->stg_ev)
      if ((e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) return e;
  }
  const size_t need = (bytes + 15) & ~(size_t)15;
  if (ctx->stg_head + need > ndtb_ctx::STG_SEG_BYTES) {  // next segment: wait until its previous contents were consumed
    ctx->stg_seg = (ctx->stg_seg + 1) % ndtb_ctx::STG_SEGS, ctx->stg_head = 0;
    if (ctx->stg_used[ctx->stg_seg]) {
      const cudaError_t e = cudaEventSynchronize(ctx->stg_ev[ctx->stg_seg]);
      if (e != cudaSuccess) return e;
    }
  }
  const size_t at = (size_t)ctx->stg_seg * ndtb_ctx::STG_SEG_BYTES + ctx->stg_head;
  ctx->stg_head += need;
  std::memcpy(ctx->stg_host + at, src, bytes);
  if ((((uintptr_t)dst | bytes) & 15) == 0) {
    const size_t n = bytes / 16;
    k_upload16<<<(unsigned)std::min<size_t>((n + 255) / 256, 296), 256, 0, st>>>((const uint4 *)(ctx->stg_dev + at), (uint4 *)dst, n);
  } else {
    k_upload1<<<(unsigned)std::min<size_t>((bytes + 255) / 256, 296), 256, 0, st>>>((const unsigned char *)(ctx->stg_dev + at),
                                                                                  (unsigned char *)dst, bytes);
  }
  ctx->launches++;
  ctx->stg_used[ctx->stg_seg] = true;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  return cudaEventRecord(ctx->stg_ev[ctx->stg_seg], st);
}

// NDTB_PROFILE=1: host-side phase timer (synchronises the stream at every mark; debugging aid, never on in benchmarks)
struct PhaseTimer {
  ndtb_ctx *ctx;
  bool on;
  const char *what;
  std::chrono::steady_clock::time_point t0;
  PhaseTimer(ndtb_ctx *c, const char *w);
  void mark(const char *label);
};

// Slabs: stream-ordered device allocations shared by the maps of a batch.  A released slab goes to a per-context cache
// keyed by its capacity class (sizes rounded up to 1/8 octave) and the next request of that class takes it back: every
// use of a context's slabs is ordered on the context's stream (the copy and covariance streams join it through events), so
// a block released on the host may be handed to work that is enqueued later.  CUDA's own stream-ordered pool sits below
// the cache and only sees the misses: in steady state it re-mapped physical memory when three contexts kept releasing
// and requesting ~2 GB slabs of slightly different sizes — single cudaMallocFromPoolAsync calls of 0.4-2.4 s, during which
// the other contexts' synchronisations stalled as well (measured with NDTB_SLOWLOG on the B200 box).
struct Slab {
  ndtb_ctx *ctx;
  char *p = nullptr;
  size_t bytes = 0;  // requested
  size_t cap = 0;    // capacity class actually allocated
  Slab(ndtb_ctx *c) : ctx(c) {}
  ~Slab();
};
using SlabP = std::shared_ptr<Slab>;

inline size_t slab_class(size_t b) {
  if (b <= 4096) return 4096;
  const int e = 63 - __builtin_clzll((unsigned long long)b);
  const size_t step = (size_t)1 << (e - 3);
  return (b + step - 1) & ~(step - 1);
}
Slab::~Slab() {
  if (!p) return;
  std::vector<char *> evict;
  {
    std::lock_guard<std::mutex> lk(ctx->cache_mu);
    ctx->slab_cache[cap].push_back({p, ++ctx->cache_clock});
    ctx->cached_bytes += cap;
    while (ctx->cached_bytes > ctx->cache_limit) {  // evict the blocks released longest ago
      size_t best_cls = 0, best_i = 0;
      uint64_t best = ~0ull;
      for (auto &kv : ctx->slab_cache)
        for (size_t i = 0; i < kv.second.size(); i++)
          if (kv.second[i].stamp < best) best = kv.second[i].stamp, best_cls = kv.first, best_i = i;
      if (best == ~0ull) break;
      auto &v = ctx->slab_cache[best_cls];
      evict.push_back(v[best_i].p);
      v.erase(v.begin() + (long)best_i);
      ctx->cached_bytes -= best_cls;
    }
  }
  for (char *q : evict) cudaFreeAsync(q, ctx->stream);
}

int slab_alloc(ndtb_ctx *ctx, size_t bytes, SlabP &out) {
  out = std::make_shared<Slab>(ctx);
  out->bytes = bytes ? bytes : 256;
  out->cap = slab_class(out->bytes);
  {
    std::lock_guard<std::mutex> lk(ctx->cache_mu);
    auto it = ctx->slab_cache.find(out->cap);
    if (it != ctx->slab_cache.end() && !it->second.empty()) {
      out->p = it->second.back().p;
      it->second.pop_back();
      ctx->cached_bytes -= out->cap;
      return NDTB_OK;
    }
  }
  const double t0 = now_ms();
  cudaError_t e = cudaMallocFromPoolAsync((void **)&out->p, out->cap, ctx->pool, ctx->stream);
  if (e != cudaSuccess) {  // out of memory with blocks parked in the cache: release them and try once more
    cudaGetLastError();
    std::vector<char *> all;
    {
      std::lock_guard<std::mutex> lk(ctx->cache_mu);
      for (auto &kv : ctx->slab_cache) {
        for (auto &c : kv.second) all.push_back(c.p);
        kv.second.clear();
      }
      ctx->cached_bytes = 0;
    }
    for (char *q : all) cudaFreeAsync(q, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    out->p = nullptr;
    CU_TRY(ctx, cudaMallocFromPoolAsync((void **)&out->p, out->cap, ctx->pool, ctx->stream));
  }
  g_slow.alloc_ms += now_ms() - t0, g_slow.allocs++;
  return NDTB_OK;
}

PhaseTimer::PhaseTimer(ndtb_ctx *c, const char *w) : ctx(c), what(w) {
  static const bool env = std::getenv("NDTB_PROFILE") != nullptr;
  on = env;
  if (on) {
    cudaStreamSynchronize(ctx->stream);
    t0 = std::chrono::steady_clock::now();
  }
}
void PhaseTimer::mark(const char *label) {
  if (!on) return;
  cudaStreamSynchronize(ctx->stream);
  const auto t1 = std::chrono::steady_clock::now();
  std::fprintf(stderr, "[ndtb] %s/%s: %.3f ms\n", what, label, std::chrono::duration<double, std::milli>(t1 - t0).count());
  t0 = t1;
}

struct Carver {
  size_t off = 0;
  size_t take(size_t bytes) {
    const size_t o = off;
    off += (bytes + 255) & ~(size_t)255;
    return o;
  }
};

}  // namespace

struct ndtb_map {
  ndtb_ctx *ctx;
  double cell[3];
  // NDTMap members [upstream]
  bool guess_size = true;
  double centerx = 0, centery = 0, centerz = 0;
  double map_sizex = -1, map_sizey = -1, map_sizez = -1;
  bool is_first_load = true;
  bool grid_ready = false;
  GridDesc g;
  int nblk = 0;
  // storage box of the dense block arrays (map_build.cuh BuildJob::boff / nbs); the whole grid unless allow_box
  int boff[3] = {0, 0, 0}, nbs[3] = {0, 0, 0};
  bool allow_box = false;  // temporaries of ndtb_register_scans: store only the box the points can occupy
  double ext_maxdist = 0, ext_dz_min = 0, ext_dz_max = 0;  // extents found by the guess-size pass (about the centroid)
  // pending points (loadPointCloud / addPointCloud before computeNDTCells)
  struct Chunk {
    SlabP buf;
    int n;
    bool trace = false;  // addPointCloud with the free-space ray trace: the points are rays from `origin`
    double origin[3] = {0, 0, 0};
    double maxz = 100.0, sensor_noise = 0.25;
    float occ_limit = 255.f;
  };
  std::vector<Chunk> pending;
  bool pending_load = false;  // pending points came from loadPointCloud (fresh map)
  bool pending_built = false; // ... and their cells were already computed (with occupancy limit pending_occ)
  float pending_occ = 255.f;
  double pending_range = -1.0;
  double range_origin[3] = {0, 0, 0};  // reference point of the range filter (loadPointCloudCentroid: the sensor origin)
  // all-cells structure
  SlabP s_blocks, s_cells;
  unsigned long long *amask = nullptr;
  int *abase = nullptr, *tb_list = nullptr, *counts = nullptr;
  int n_all = 0, ntb = 0;
  double *cmean = nullptr, *ccov = nullptr;
  int *cn = nullptr, *chas = nullptr;
  float *cocc = nullptr;
  // Gaussian view
  double *gcell = nullptr;
  int *g2c = nullptr;
  HashEntry *table = nullptr;
  int tsize = 0, ng = 0, ngb = 0;
  int64_t last_binned = 0;

  void set_grid(double cx, double cy, double cz, double sx, double sy, double sz) {
    g.center[0] = cx, g.center[1] = cy, g.center[2] = cz;
    const double sm[3] = {sx, sy, sz};
    for (int i = 0; i < 3; i++) {
      g.cell[i] = cell[i];
      g.size[i] = (int32_t)std::abs(std::ceil(sm[i] / cell[i]));  // LazyGrid::setSize
      g.nb[i] = (g.size[i] + 3) / 4;
      boff[i] = 0, nbs[i] = g.nb[i];
    }
    grid_ready = true;
    drop_cells();
  }
  void fill_box(BuildJob &j) const {
    for (int i = 0; i < 3; i++) j.boff[i] = boff[i], j.nbs[i] = nbs[i];
  }
  // guess-size grid: the points lie within maxDist of the centre (z: within the recorded extent) -> store that box only
  void shrink_box() {
    const double lo[3] = {-ext_maxdist, -ext_maxdist, -ext_dz_max}, hi[3] = {ext_maxdist, ext_maxdist, -ext_dz_min};
    for (int i = 0; i < 3; i++) {
      const double a = std::floor(lo[i] / g.cell[i] + 0.5) + g.size[i] / 2.0, b = std::floor(hi[i] / g.cell[i] + 0.5) + g.size[i] / 2.0;
      int b0 = ((int)a >> 2) - 1, b1 = ((int)b >> 2) + 1;
      b0 = std::max(b0, 0), b1 = std::min(b1, g.nb[i] - 1);
      if (b1 < b0) b0 = 0, b1 = g.nb[i] - 1;
      boff[i] = b0, nbs[i] = b1 - b0 + 1;
    }
  }
  int64_t nblocks() const { return (int64_t)nbs[0] * nbs[1] * nbs[2]; }
  void drop_cells() {
    s_blocks.reset(), s_cells.reset();
    amask = nullptr, abase = nullptr, tb_list = nullptr, counts = nullptr;
    cmean = ccov = nullptr, cn = chas = nullptr, cocc = nullptr, gcell = nullptr, g2c = nullptr, table = nullptr;
    n_all = ntb = tsize = ng = ngb = 0;
  }
};

namespace {

struct PointSrc {
  const float4 *dev;  // device pointer
  int n;
};

// fill a MapView for the matcher
MapView view_of(const ndtb_map *m) {
  MapView v;
  v.g = m->g;
  v.gcell = m->gcell;
  v.table = m->table;
  v.ng = m->ng;
  v.tsize = m->tsize;
  return v;
}

// NDTMap::loadPointCloud [upstream], grid part: guess_size -> centre = centroid of the usable points, size = setMapSize
// values or (4 maxDist, 4 maxDist, 3 (maxz - minz)); else the guessSize()/initialize() values.  empty[i] = no usable point.
int define_grids(ndtb_ctx *ctx, const std::vector<ndtb_map *> &maps, const std::vector<PointSrc> &pts,
                 const std::vector<double> &range, std::vector<char> &empty) {
  const int M = (int)maps.size();
  cudaStream_t st = ctx->stream;
  empty.assign(M, 0);
  std::vector<int> which;
  for (int i = 0; i < M; i++)
    if (maps[i]->guess_size) which.push_back(i);
  if (!which.empty()) {
    const int W = (int)which.size();
    std::vector<BuildJob> jobs(W);
    std::memset(jobs.data(), 0, sizeof(BuildJob) * W);
    std::vector<int> ident(W);
    int max_pts = 1;
    for (int w = 0; w < W; w++) {
      jobs[w].pts = pts[which[w]].dev, jobs[w].npts = pts[which[w]].n, jobs[w].range_limit = range[which[w]];
      ident[w] = w;
      max_pts = std::max(max_pts, jobs[w].npts);
    }
    Carver c;
    const size_t o_j = c.take(sizeof(BuildJob) * W), o_gs = c.take(64 * (size_t)W), o_w = c.take(4 * (size_t)W);
    std::vector<long long> rec_off((size_t)W + 1, 0);  // centroid chunk records (map_build.cu k_centroid_chunks)
    for (int w = 0; w < W; w++) rec_off[w + 1] = rec_off[w] + (jobs[w].npts + centroid_chunk_points() - 1) / centroid_chunk_points();
    const size_t o_ro = c.take(8 * (size_t)(W + 1)), o_rec = c.take(8 * (size_t)centroid_record_doubles() * (size_t)std::max<long long>(rec_off[W], 1));
    SlabP s;
    if (int rc = slab_alloc(ctx, c.off, s)) return rc;
    std::vector<double> gs(8 * (size_t)W);
    CU_TRY(ctx, h2d_small(ctx, s->p + o_j, jobs.data(), sizeof(BuildJob) * W, st));
    CU_TRY(ctx, h2d_small(ctx, s->p + o_w, ident.data(), 4 * (size_t)W, st));
    CU_TRY(ctx, h2d_small(ctx, s->p + o_ro, rec_off.data(), 8 * (size_t)(W + 1), st));
    CU_TRY(ctx, cudaMemsetAsync(s->p + o_gs, 0, 64 * (size_t)W, st));
    ctx->launches += launch_guess((const BuildJob *)(s->p + o_j), (const int *)(s->p + o_w), W, max_pts, (const long long *)(s->p + o_ro),
                                  (double *)(s->p + o_rec), (double *)(s->p + o_gs), st);
    CU_TRY(ctx, cudaGetLastError());
    CU_TRY(ctx, cudaMemcpyAsync(gs.data(), s->p + o_gs, 64 * (size_t)W, cudaMemcpyDeviceToHost, st));
    CU_TRY(ctx, timed_sync(st));
    auto unkey = [](double bits) {
      unsigned long long k;
      std::memcpy(&k, &bits, 8);
      const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
      double v;
      std::memcpy(&v, &b, 8);
      return v;
    };
    for (int w = 0; w < W; w++) {
      ndtb_map *m = maps[which[w]];
      const double *o = &gs[8 * (size_t)w];
      if (o[3] <= 0) {
        empty[which[w]] = 1;
        m->grid_ready = false;
        m->drop_cells();
        continue;
      }
      const double maxDist = o[4], maxz = unkey(o[5]), minz = unkey(o[6]);
      if (m->map_sizex > 0 && m->map_sizey > 0 && m->map_sizez > 0)
        m->set_grid(o[0], o[1], o[2], m->map_sizex, m->map_sizey, m->map_sizez);
      else
        m->set_grid(o[0], o[1], o[2], 4 * maxDist, 4 * maxDist, 3 * (maxz - minz));
      m->ext_maxdist = maxDist, m->ext_dz_min = minz, m->ext_dz_max = maxz;
      if (m->allow_box) m->shrink_box();
      m->is_first_load = false;
    }
  }
  for (int i = 0; i < M; i++) {
    ndtb_map *m = maps[i];
    if (!m->guess_size) {
      m->set_grid(m->centerx, m->centery, m->centerz, m->map_sizex, m->map_sizey, m->map_sizez);
      m->is_first_load = false;
    }
  }
  return NDTB_OK;
}

// Batched (load|add)PointCloud + computeNDTCells.  pts[i] are device pointers that stay valid during the call.
// load[i]: 1 = loadPointCloud semantics (grid (re)defined, map emptied), 2 = fresh map on the grid it already has,
//          0 = addPointCloud (merge into the existing cells).
using TraceSegs = std::vector<TraceSeg>;
int build_batch_slice(ndtb_ctx *ctx, const std::vector<ndtb_map *> &maps, const std::vector<PointSrc> &pts,
                      const std::vector<char> &load, const std::vector<double> &range, uint32_t maxnumpoints, float occ_limit,
                      const std::vector<TraceSegs> *trace);

// The build kernels put the map index on gridDim.y (limit 65535): larger batches are built slice by slice.
constexpr int BUILD_SLICE = 32768;
int build_batch(ndtb_ctx *ctx, const std::vector<ndtb_map *> &maps, const std::vector<PointSrc> &pts,
                const std::vector<char> &load, const std::vector<double> &range, uint32_t maxnumpoints, float occ_limit,
                const std::vector<TraceSegs> *trace = nullptr) {
  const size_t M = maps.size();
  if (M <= (size_t)BUILD_SLICE) return build_batch_slice(ctx, maps, pts, load, range, maxnumpoints, occ_limit, trace);
  for (size_t b = 0; b < M; b += BUILD_SLICE) {
    const size_t e = std::min(M, b + BUILD_SLICE);
    std::vector<ndtb_map *> m(maps.begin() + b, maps.begin() + e);
    std::vector<PointSrc> p(pts.begin() + b, pts.begin() + e);
    std::vector<char> l(load.begin() + b, load.begin() + e);
    std::vector<double> r(range.begin() + b, range.begin() + e);
    std::vector<TraceSegs> t;
    if (trace) t.assign(trace->begin() + b, trace->begin() + e);
    if (int rc = build_batch_slice(ctx, m, p, l, r, maxnumpoints, occ_limit, trace ? &t : nullptr)) return rc;
  }
  return NDTB_OK;
}

int build_batch_slice(ndtb_ctx *ctx, const std::vector<ndtb_map *> &maps, const std::vector<PointSrc> &pts,
                      const std::vector<char> &load, const std::vector<double> &range, uint32_t maxnumpoints, float occ_limit,
                      const std::vector<TraceSegs> *trace) {
  const int M = (int)maps.size();
  if (M == 0) return NDTB_OK;
  cudaStream_t st = ctx->stream;
  PhaseTimer pt(ctx, "build");
  std::vector<BuildJob> jobs(M);
  std::memset(jobs.data(), 0, sizeof(BuildJob) * M);
  int max_pts = 1;
  bool any_trace = false;
  for (int i = 0; i < M; i++) {
    jobs[i].pts = pts[i].dev;
    jobs[i].npts = pts[i].n;
    jobs[i].range_limit = load[i] ? range[i] : -1.0;
    for (int a = 0; a < 3; a++) jobs[i].range_origin[a] = load[i] ? maps[i]->range_origin[a] : 0.0;
    jobs[i].maxnumpoints = maxnumpoints;
    jobs[i].occ_limit = occ_limit;
    jobs[i].log_occ = std::log(0.6 / (1.0 - 0.6));
    max_pts = std::max(max_pts, pts[i].n);
    if (trace && !(*trace)[i].empty()) {
      if ((*trace)[i].size() > (size_t)MAX_TRACE_SEG) return NDTB_ERR_ARG;
      jobs[i].n_seg = (int)(*trace)[i].size();
      for (int q = 0; q < jobs[i].n_seg; q++) jobs[i].seg[q] = (*trace)[i][q];
      any_trace = true;
    }
  }
  SlabP s_jobs;
  if (int rc = slab_alloc(ctx, sizeof(BuildJob) * M, s_jobs)) return rc;
  BuildJob *d_jobs = (BuildJob *)s_jobs->p;

  // ---- phase A: grids
  std::vector<char> empty(M, 0);
  {
    std::vector<ndtb_map *> lm;
    std::vector<PointSrc> lp;
    std::vector<double> lr;
    std::vector<int> li;
    for (int i = 0; i < M; i++)
      if (load[i] == 1) lm.push_back(maps[i]), lp.push_back(pts[i]), lr.push_back(range[i]), li.push_back(i);
    std::vector<char> le;
    if (!lm.empty()) {
      if (int rc = define_grids(ctx, lm, lp, lr, le)) return rc;
      for (size_t q = 0; q < li.size(); q++) empty[li[q]] = le[q];
    }
  }
  for (int i = 0; i < M; i++) {
    ndtb_map *m = maps[i];
    if (load[i] == 2) {
      if (!m->grid_ready) empty[i] = 1;
      else m->drop_cells();
    }
    if (empty[i]) {
      m->drop_cells();
      m->last_binned = 0;
      continue;
    }
    if (!m->grid_ready) return NDTB_ERR_GRID;
    if (m->nblocks() <= 0 || m->nblocks() * 64 >= ((int64_t)1 << 31)) return NDTB_ERR_GRID;
    m->nblk = (int)m->nblocks();
  }

  pt.mark("grids");
  // ---- phase B: mark touched voxels, number the cells
  std::vector<SlabP> keep_old;  // previous storage of merged maps stays alive until the end of the build
  Carver cb, ct;
  struct OffB {
    size_t amask, abase, tbl, counts, ptc, seg, seg2, key2, rshist, viscnt, visoff;
  };
  std::vector<OffB> ob(M);
  const size_t o_counts_all = cb.take(32 * (size_t)M);  // contiguous: one D2H copy for the whole batch
  for (int i = 0; i < M; i++) {
    if (empty[i]) continue;
    ndtb_map *m = maps[i];
    // touched blocks: at most one new block per point, or — with the ray trace — any block of the grid
    const int64_t tb_cap = jobs[i].n_seg ? (int64_t)m->nblk : std::min<int64_t>(m->nblk, (int64_t)m->n_all + pts[i].n);
    ob[i].amask = cb.take(8 * (size_t)m->nblk);
    ob[i].abase = cb.take(4 * (size_t)m->nblk);
    ob[i].tbl = cb.take(4 * (size_t)std::max<int64_t>(tb_cap, 1));
    ob[i].counts = o_counts_all + 32 * (size_t)i;
    ob[i].ptc = ct.take(4 * (size_t)pts[i].n);
    ob[i].seg = ct.take(4 * (size_t)pts[i].n);
    ob[i].seg2 = ct.take(4 * (size_t)pts[i].n);
    ob[i].key2 = ct.take(4 * (size_t)pts[i].n);
    ob[i].rshist = ct.take(4 * 256 * (size_t)((pts[i].n + sort_tile_points() - 1) / sort_tile_points() + 1));
    if (jobs[i].n_seg) ob[i].viscnt = ct.take(4 * (size_t)pts[i].n), ob[i].visoff = ct.take(4 * (size_t)pts[i].n);
  }
  SlabP s_b, s_t;
  if (int rc = slab_alloc(ctx, cb.off, s_b)) return rc;
  if (int rc = slab_alloc(ctx, ct.off, s_t)) return rc;
  CU_TRY(ctx, cudaMemsetAsync(s_b->p, 0, s_b->bytes, st));
  for (int i = 0; i < M; i++) {
    if (empty[i]) continue;
    ndtb_map *m = maps[i];
    BuildJob &j = jobs[i];
    j.g = m->g;
    m->fill_box(j);
    j.nblk = m->nblk;
    j.amask = (unsigned long long *)(s_b->p + ob[i].amask);
    j.abase = (int *)(s_b->p + ob[i].abase);
    j.tb_list = (int *)(s_b->p + ob[i].tbl);
    j.counts = (int *)(s_b->p + ob[i].counts);
    j.pt_cell = (int *)(s_t->p + ob[i].ptc);
    j.seg_idx = (int *)(s_t->p + ob[i].seg);
    j.seg2 = (int *)(s_t->p + ob[i].seg2);
    j.key2 = (int *)(s_t->p + ob[i].key2);
    j.rs_hist = (int *)(s_t->p + ob[i].rshist);
    if (j.n_seg) j.vis_cnt = (int *)(s_t->p + ob[i].viscnt), j.vis_off = (int *)(s_t->p + ob[i].visoff);
    if (m->n_all > 0) {  // merge: start from the cells the map already has
      j.o_amask = m->amask, j.o_abase = m->abase, j.o_cmean = m->cmean, j.o_ccov = m->ccov;
      j.o_cn = m->cn, j.o_chas = m->chas, j.o_cocc = m->cocc;
      CU_TRY(ctx, cudaMemcpyAsync(j.amask, m->amask, 8 * (size_t)m->nblk, cudaMemcpyDeviceToDevice, st));
      keep_old.push_back(m->s_blocks), keep_old.push_back(m->s_cells);
    }
  }
  std::vector<BuildJob> live;  // only non-empty maps are launched
  std::vector<int> live_idx;
  for (int i = 0; i < M; i++)
    if (!empty[i]) live.push_back(jobs[i]), live_idx.push_back(i);
  const int L = (int)live.size();
  if (L == 0) return NDTB_OK;
  CU_TRY(ctx, h2d_small(ctx, d_jobs, live.data(), sizeof(BuildJob) * L, st));
  bool fast_mark = !any_trace;  // power-of-two cells, no range limit, no trace segment: k_mark<true>
  for (const BuildJob &lj : live) {
    for (int a = 0; a < 3; a++) {
      int ex = 0;
      fast_mark = fast_mark && std::frexp(lj.g.cell[a], &ex) == 0.5 && ex > -900 && ex < 900;
    }
    fast_mark = fast_mark && !(lj.range_limit > 0) && lj.n_seg == 0;
  }
  ctx->launches += launch_mark(d_jobs, L, max_pts, any_trace, fast_mark, st);
  CU_TRY(ctx, cudaGetLastError());
  std::vector<int> cnts_all(8 * (size_t)M), cnts(8 * (size_t)L);
  CU_TRY(ctx, cudaMemcpyAsync(cnts_all.data(), s_b->p + o_counts_all, 32 * (size_t)M, cudaMemcpyDeviceToHost, st));
  CU_TRY(ctx, timed_sync(st));
  for (int l = 0; l < L; l++) std::memcpy(&cnts[8 * l], &cnts_all[8 * (size_t)live_idx[l]], 32);
  pt.mark("mark+scan");

  // ---- phase C: cell records + Gaussian view
  Carver cc, ct2;
  struct OffC {
    size_t mean, cov, n, has, occ, gcell, g2c, table, cnt, segoff, cursor, gmask, gbase, ckey;
    size_t vkey, vray, vidx, vseg2, vkey2, vhist, vcnt, vsegoff;
  };
  std::vector<OffC> oc(L);
  int max_ntb = 1, max_cells = 1, max_vis = 1;
  for (int l = 0; l < L; l++) {  // the hash tables first, contiguous: one memset initialises them all
    int tsize = 2;
    while (tsize < 2 * cnts[8 * l + 1]) tsize <<= 1;
    live[l].tsize = tsize;
    oc[l].table = cc.take(sizeof(HashEntry) * (size_t)tsize);
  }
  const size_t tables_bytes = cc.off;
  for (int l = 0; l < L; l++) {
    const int n_all = cnts[8 * l], ntb = cnts[8 * l + 1];
    max_ntb = std::max(max_ntb, ntb);
    max_cells = std::max(max_cells, n_all);
    live[l].n_all = n_all;
    const size_t na = (size_t)std::max(n_all, 1);
    oc[l].mean = cc.take(24 * na), oc[l].cov = cc.take(72 * na), oc[l].n = cc.take(4 * na), oc[l].has = cc.take(4 * na);
    oc[l].occ = cc.take(4 * na), oc[l].gcell = cc.take(72 * na), oc[l].g2c = cc.take(4 * na);
    oc[l].cnt = ct2.take(4 * na), oc[l].segoff = ct2.take(4 * na), oc[l].cursor = ct2.take(4 * na), oc[l].ckey = ct2.take(4 * na);
    oc[l].gmask = ct2.take(8 * (size_t)std::max(ntb, 1)), oc[l].gbase = ct2.take(4 * (size_t)std::max(ntb, 1));
    if (live[l].n_seg) {
      const size_t nv = (size_t)std::max(cnts[8 * l + 7], 1);
      max_vis = std::max(max_vis, cnts[8 * l + 7]);
      oc[l].vkey = ct2.take(4 * nv), oc[l].vray = ct2.take(4 * nv), oc[l].vidx = ct2.take(4 * nv), oc[l].vseg2 = ct2.take(4 * nv);
      oc[l].vkey2 = ct2.take(4 * nv), oc[l].vhist = ct2.take(4 * 256 * (nv / (size_t)sort_tile_points() + 2));
      oc[l].vcnt = ct2.take(4 * na), oc[l].vsegoff = ct2.take(4 * na);
    }
  }
  SlabP s_c, s_t2;
  if (int rc = slab_alloc(ctx, cc.off, s_c)) return rc;
  if (int rc = slab_alloc(ctx, ct2.off, s_t2)) return rc;
  CU_TRY(ctx, cudaMemsetAsync(s_t2->p, 0, s_t2->bytes, st));
  CU_TRY(ctx, cudaMemsetAsync(s_c->p, 0xFF, tables_bytes, st));
  for (int l = 0; l < L; l++) {
    BuildJob &j = live[l];
    j.cmean = (double *)(s_c->p + oc[l].mean), j.ccov = (double *)(s_c->p + oc[l].cov);
    j.cn = (int *)(s_c->p + oc[l].n), j.chas = (int *)(s_c->p + oc[l].has), j.cocc = (float *)(s_c->p + oc[l].occ);
    j.gcell = (double *)(s_c->p + oc[l].gcell), j.g2c = (int *)(s_c->p + oc[l].g2c);
    j.table = (HashEntry *)(s_c->p + oc[l].table);
    j.cnt = (int *)(s_t2->p + oc[l].cnt), j.seg_off = (int *)(s_t2->p + oc[l].segoff), j.cursor = (int *)(s_t2->p + oc[l].cursor);
    j.gmask_t = (unsigned long long *)(s_t2->p + oc[l].gmask), j.gbase_t = (int *)(s_t2->p + oc[l].gbase);
    j.cell_key = (int *)(s_t2->p + oc[l].ckey);
    // the stable radix sort ping-pongs between (pt_cell, seg_idx) and (key2, seg2): an odd number of passes ends in seg2
    const bool odd = sort_passes(max_cells) & 1;
    j.sorted_ids = odd ? j.seg2 : j.seg_idx;
    if (j.n_seg) {
      j.vis_key = (int *)(s_t2->p + oc[l].vkey), j.vis_ray = (int *)(s_t2->p + oc[l].vray);
      j.v_cnt = (int *)(s_t2->p + oc[l].vcnt), j.v_seg_off = (int *)(s_t2->p + oc[l].vsegoff);
      j.v_seg2 = (int *)(s_t2->p + (odd ? oc[l].vseg2 : oc[l].vidx));
    }
  }
  CU_TRY(ctx, h2d_small(ctx, d_jobs, live.data(), sizeof(BuildJob) * L, st));
  pt.mark("alloc C");
  SlabP s_vjobs;
  if (any_trace) {  // per-cell visit lists through the points' counting sort: the same kernels on the per-visit arrays
    std::vector<BuildJob> vj(live);
    for (int l = 0; l < L; l++) {
      BuildJob &v = vj[l];
      v.pts = nullptr;
      v.npts = live[l].n_seg ? cnts[8 * l + 7] : 0;
      v.pt_cell = live[l].vis_key;
      v.seg_idx = live[l].n_seg ? (int *)(s_t2->p + oc[l].vidx) : nullptr, v.seg2 = live[l].n_seg ? (int *)(s_t2->p + oc[l].vseg2) : nullptr;
      v.key2 = live[l].n_seg ? (int *)(s_t2->p + oc[l].vkey2) : nullptr, v.rs_hist = live[l].n_seg ? (int *)(s_t2->p + oc[l].vhist) : nullptr;
      v.cnt = live[l].v_cnt, v.seg_off = live[l].v_seg_off;
      if (!live[l].n_seg) v.n_all = 0, v.cnt = nullptr;  // nothing to do for a map without traced rays
    }
    if (int rc = slab_alloc(ctx, sizeof(BuildJob) * L, s_vjobs)) return rc;
    CU_TRY(ctx, h2d_small(ctx, s_vjobs->p, vj.data(), sizeof(BuildJob) * L, st));
    ctx->launches += launch_trace_lists(d_jobs, (const BuildJob *)s_vjobs->p, L, max_pts, max_vis, max_ntb, max_cells, st);
    CU_TRY(ctx, cudaGetLastError());
  }
  ctx->launches += launch_cells(d_jobs, L, max_pts, max_ntb, max_cells, st);
  CU_TRY(ctx, cudaGetLastError());
  pt.mark("cells");
  ctx->launches += launch_gview(d_jobs, L, max_ntb, max_cells, st);
  CU_TRY(ctx, cudaGetLastError());
  CU_TRY(ctx, cudaMemcpyAsync(cnts_all.data(), s_b->p + o_counts_all, 32 * (size_t)M, cudaMemcpyDeviceToHost, st));
  CU_TRY(ctx, timed_sync(st));
  for (int l = 0; l < L; l++) std::memcpy(&cnts[8 * l], &cnts_all[8 * (size_t)live_idx[l]], 32);
  pt.mark("gview");
  for (int l = 0; l < L; l++) {
    ndtb_map *m = maps[live_idx[l]];
    const BuildJob &j = live[l];
    m->s_blocks = s_b, m->s_cells = s_c;
    m->amask = j.amask, m->abase = j.abase, m->tb_list = j.tb_list, m->counts = j.counts;
    m->n_all = cnts[8 * l], m->ntb = cnts[8 * l + 1], m->ng = cnts[8 * l + 2], m->ngb = cnts[8 * l + 3];
    m->last_binned = cnts[8 * l + 4];
    m->cmean = j.cmean, m->ccov = j.ccov, m->cn = j.cn, m->chas = j.chas, m->cocc = j.cocc;
    m->gcell = j.gcell, m->g2c = j.g2c, m->table = j.table, m->tsize = j.tsize;
  }
  pt.mark("publish");
  s_t.reset(), s_t2.reset(), s_jobs.reset(), s_vjobs.reset(), keep_old.clear();
  pt.mark("free temps");
  return NDTB_OK;
}

// device copy of host/device points
int stage_points(ndtb_ctx *ctx, const float *pts, int64_t n, int mem, SlabP &out) {
  if (int rc = slab_alloc(ctx, 16 * (size_t)std::max<int64_t>(n, 1), out)) return rc;
  if (n > 0)
    CU_TRY(ctx, mem == NDTB_MEM_DEVICE ? cudaMemcpyAsync(out->p, pts, 16 * (size_t)n, cudaMemcpyDeviceToDevice, ctx->stream)
                                       : h2d_small(ctx, out->p, pts, 16 * (size_t)n, ctx->stream));
  return NDTB_OK;
}

MatchConfig make_config(const ndtb_ctx *ctx, const ndtb_params *p, int max_tsize) {
  MatchConfig c;
  c.n_neighbours = p->n_neighbours;
  c.itr_max = p->itr_max, c.step_control = p->step_control, c.regularize = p->regularize, c.planar = p->planar;
  c.soft = p->use_soft_constraints, c.tik = p->use_tikhonov;
  c.delta_score = p->delta_score, c.lfd1 = p->lfd1, c.lfd2 = p->lfd2;
  // stage the block table in shared memory when it fits next to the optimiser state
  int cap = (int)((ctx->smem_optin - (int)match_smem_bytes(0) - 1024) / (int)sizeof(HashEntry));
  if (cap < 0) cap = 0;
  c.table_smem_entries = max_tsize <= cap ? max_tsize : 0;
  if (max_tsize > cap) {  // largest power of two that fits: smaller maps of the batch still get staged
    int t = 1;
    while (2 * t <= cap) t <<= 1;
    c.table_smem_entries = cap >= 2 ? t : 0;
  }
  return c;
}

void fill_job(MatchJob &j, const ndtb_map *tgt, const ndtb_map *src, const double *T0, const double *Q36) {
  std::memset(&j, 0, sizeof j);
  j.tgt = view_of(tgt);
  j.src_gcell = src->gcell;
  j.src_ng = src->ng;
  std::memcpy(j.T0, T0, sizeof j.T0);
  j.fusion = Q36 != nullptr;
  if (Q36) std::memcpy(j.Q, Q36, sizeof j.Q);
}

bool map_ok(const ndtb_map *m) { return m && m->grid_ready; }

// a map without any cell still needs valid pointers for the kernels: give it a 2-entry empty table
int ensure_view(ndtb_ctx *ctx, ndtb_map *m) {
  if (m->table) return NDTB_OK;
  SlabP s;
  if (int rc = slab_alloc(ctx, 256, s)) return rc;
  CU_TRY(ctx, cudaMemsetAsync(s->p, 0xFF, 256, ctx->stream));
  m->s_cells = s;
  m->table = (HashEntry *)s->p;
  m->tsize = 2;
  m->gcell = (double *)(s->p + 64);
  m->ng = 0;
  return NDTB_OK;
}

int match_batch_impl(ndtb_ctx *ctx, int64_t n, const ndtb_map *const *tgt, const ndtb_map *const *src, const double *T0s,
                     const double *Q36s /*n x 36 or null*/, const ndtb_params *p, int with_cov, int out_mem, ndtb_result *res,
                     double *cov36s) {
  if (n <= 0) return NDTB_OK;
  cudaStream_t st = ctx->stream;
  std::vector<MatchJob> jobs((size_t)n);
  int max_tsize = 2;
  size_t gt_total = 0;
  std::vector<long long> gt_off((size_t)n);
  for (int64_t e = 0; e < n; e++) {
    if (!map_ok(tgt[e]) || !map_ok(src[e])) return NDTB_ERR_GRID;
    if (tgt[e]->ctx != ctx || src[e]->ctx != ctx) return NDTB_ERR_ARG;  // storage is released in the order of its own context's stream
    if (int rc = ensure_view(ctx, const_cast<ndtb_map *>(tgt[e]))) return rc;
    if (int rc = ensure_view(ctx, const_cast<ndtb_map *>(src[e]))) return rc;
    fill_job(jobs[e], tgt[e], src[e], T0s + 16 * e, Q36s ? Q36s + 36 * e : nullptr);
    max_tsize = std::max(max_tsize, tgt[e]->tsize);
    gt_off[e] = (long long)gt_total;
    gt_total += (size_t)std::max(tgt[e]->ng, 1);
  }
  const MatchConfig cfg = make_config(ctx, p, max_tsize);
  // CTAs per registration in the covariance pass: a small batch wants them all for latency, a large one fewer and longer
  // ones (12 warps share the rounds of a chunk: 2-3 rounds per warp leave a third of them idle at the end)
  static const int env_chunks = std::getenv("NDTB_COV_CHUNKS") ? std::atoi(std::getenv("NDTB_COV_CHUNKS")) : 0;
  const int n_chunks = env_chunks > 0 ? env_chunks : (n >= 2 * ctx->sm_count ? 3 : 8);
  Carver c;
  const size_t o_jobs = c.take(sizeof(MatchJob) * n), o_res = c.take(sizeof(ndtb_result) * n);
  const size_t o_goff = c.take(8 * n), o_gt = c.take(with_cov ? 48 * gt_total : 0);
  const size_t o_part = c.take(with_cov ? 8 * (size_t)cov_partial_width() * n_chunks * n : 0);
  const size_t o_cov = c.take(with_cov ? 288 * (size_t)n : 0), o_stat = c.take(4 * n);
  // cluster width and straggler policy.  A registration that does not converge runs ITR_MAX iterations, ~5x the mean
  // work: with one CTA per registration a handful of them would set the kernel time of a whole batch, so in large
  // batches every registration gets a pass budget and the unfinished ones are completed on 8-CTA clusters.
  const int sms = std::max(ctx->sm_count, 1);
  auto pow2_floor = [](int v) { int p = 1; while (2 * p <= v) p <<= 1; return p; };
  int G1 = p->ctas_per_match > 0 ? pow2_floor(std::min(p->ctas_per_match, 8)) : ((int64_t)2 * n >= sms ? 1 : pow2_floor((int)std::min<int64_t>(8, sms / n)));
  // hand-over also for batches that fill only half of the GPU (an edge shard of a multi-GPU run): their stragglers would
  // otherwise finish alone on one CTA each
  int budget = p->pass_budget > 0 ? p->pass_budget : (p->pass_budget < 0 ? 0 : (((int64_t)2 * n * G1 >= sms) ? 64 : 0));
  const size_t o_states = c.take(budget > 0 ? opt_state_bytes() * (size_t)n : 0), o_unf = c.take(4 * (size_t)(2 * n + 1));
  SlabP s;
  if (int rc = slab_alloc(ctx, c.off, s)) return rc;
  MatchJob *d_jobs = (MatchJob *)(s->p + o_jobs);
  ndtb_result *d_res = out_mem == NDTB_MEM_DEVICE ? res : (ndtb_result *)(s->p + o_res);
  const bool do_cov = with_cov && cov36s;
  double *d_cov = do_cov ? (out_mem == NDTB_MEM_DEVICE ? cov36s : (double *)(s->p + o_cov)) : nullptr;
  CU_TRY(ctx, h2d_small(ctx, d_jobs, jobs.data(), sizeof(MatchJob) * n, st));
  if (do_cov) {
    CU_TRY(ctx, h2d_small(ctx, s->p + o_goff, gt_off.data(), 8 * n, st));
    CU_TRY(ctx, cudaMemsetAsync(s->p + o_gt, 0, 48 * gt_total, st));
  }
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  if (ctx->timing) {
    CU_TRY(ctx, cudaEventCreate(&ev0));
    CU_TRY(ctx, cudaEventCreate(&ev1));
    CU_TRY(ctx, cudaEventRecord(ev0, st));
  }
  int *d_unf = (int *)(s->p + o_unf), *d_yielded = d_unf + 1 + n;
  if (budget > 0) CU_TRY(ctx, cudaMemsetAsync(d_unf, 0, 4 * (size_t)(2 * n + 1), st));
  CU_TRY(ctx, launch_match(d_jobs, nullptr, (int)n, G1, cfg, d_res, s->p + o_states, 0, budget, d_unf, budget > 0 ? d_yielded : nullptr, st));
  ctx->launches += 1;
  bool cov_split = false;
  if (budget > 0) {
    int n_unf = 0;
    CU_TRY(ctx, cudaMemcpyAsync(&n_unf, d_unf, 4, cudaMemcpyDeviceToHost, st));
    CU_TRY(ctx, timed_sync(st));
    if (n_unf > 0) {
      if (do_cov) {
        if (!ctx->aux_stream) {
          CU_TRY(ctx, cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
          for (cudaEvent_t &e : ctx->aux_ev) CU_TRY(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        }
        CU_TRY(ctx, cudaEventRecord(ctx->aux_ev[0], st));  // first launch complete
      }
      // Always the widest cluster, even when n_unf x 8 exceeds the SMs: most of the handed-over registrations need only
      // a few more passes and leave quickly; the wall time is set by the few that need hundreds of passes, i.e. by the
      // latency of one pass (measured: 62 ms vs 72 ms for 592 pairs with 30 hand-overs).
      int G2 = 8;
      if (const char *g2 = std::getenv("NDTB_G2")) G2 = pow2_floor(std::max(1, std::min(8, std::atoi(g2))));  // A/B knob
      CU_TRY(ctx, launch_match(d_jobs, d_unf + 1, n_unf, G2, cfg, d_res, s->p + o_states, 1, 0, d_unf, nullptr, st));
      ctx->launches += 1;
      if (do_cov) {
        // Covariance of everything that is already final, on the second stream, concurrently with the finishing launch.
        // Enqueued AFTER it: the finishing launch's clusters (whole SMs) are placed first, the covariance CTAs fill
        // the SMs it leaves free.
        CU_TRY(ctx, cudaStreamWaitEvent(ctx->aux_stream, ctx->aux_ev[0], 0));
        CU_TRY(ctx, launch_covariance(d_jobs, (int)n, cfg, d_res, (const long long *)(s->p + o_goff), (double *)(s->p + o_gt),
                                      (double *)(s->p + o_part), n_chunks, d_cov, (int *)(s->p + o_stat), d_yielded, 1,
                                      ctx->aux_stream));
        CU_TRY(ctx, cudaEventRecord(ctx->aux_ev[1], ctx->aux_stream));
        ctx->launches += 2;
        cov_split = true;
      }
    }
  }
  if (ctx->timing) {
    CU_TRY(ctx, cudaEventRecord(ev1, st));
    ctx->timed.push_back({ev0, ev1});
  }
  if (do_cov) {
    CU_TRY(ctx, launch_covariance(d_jobs, (int)n, cfg, d_res, (const long long *)(s->p + o_goff), (double *)(s->p + o_gt),
                                  (double *)(s->p + o_part), n_chunks, d_cov, (int *)(s->p + o_stat), d_yielded,
                                  cov_split ? 2 : 0, st));
    ctx->launches += 2;
    if (cov_split) CU_TRY(ctx, cudaStreamWaitEvent(st, ctx->aux_ev[1], 0));
    if (out_mem != NDTB_MEM_DEVICE)
      CU_TRY(ctx, cudaMemcpyAsync(cov36s, d_cov, 288 * (size_t)n, cudaMemcpyDeviceToHost, st));
  }
  if (out_mem != NDTB_MEM_DEVICE) {
    CU_TRY(ctx, cudaMemcpyAsync(res, d_res, sizeof(ndtb_result) * n, cudaMemcpyDeviceToHost, st));
    CU_TRY(ctx, timed_sync(st));
  }
  return NDTB_OK;
}

}  // namespace

// =========================================================================================== C ABI
extern "C" {

int ndtb_version(void) { return NDTB_VERSION; }

const char *ndtb_strerror(int code) {
  switch (code) {
    case NDTB_OK: return "ok";
    case NDTB_ERR_CUDA: return "CUDA error or no CUDA device (this engine has no CPU path)";
    case NDTB_ERR_ARG: return "bad argument";
    case NDTB_ERR_GRID: return "grid undefined or too large";
    case NDTB_ERR_EMPTY: return "no usable points / cells";
    case NDTB_ERR_SINGULAR: return "singular 6x6 system";
    case NDTB_ERR_NOMEM: return "out of memory";
    default: return "unknown error";
  }
}

const char *ndtb_last_error(const ndtb_ctx *ctx) { return ctx ? ctx->last_error.c_str() : ""; }

int ndtb_ctx_create(int device, void *stream, ndtb_ctx **out) {
  if (!out) return NDTB_ERR_ARG;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return NDTB_ERR_CUDA;
  if (cudaSetDevice(device) != cudaSuccess) return NDTB_ERR_CUDA;
  ndtb_ctx *c = new ndtb_ctx();
  c->device = device;
  if (stream) {
    c->stream = (cudaStream_t)stream;
  } else {
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
      delete c;
      return NDTB_ERR_CUDA;
    }
    c->own_stream = true;
  }
  cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
  cudaDeviceGetAttribute(&c->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  // A private pool per context: contexts driven concurrently (one per host thread) would otherwise hand each other's
  // freed slabs around through the device's default pool, which makes the driver insert cross-stream dependencies (or fall
  // back to a synchronising cudaMalloc) and serialises them.  Freed slabs stay cached: after warm-up an allocation is a
  // pointer bump.
  cudaMemPoolProps props = {};
  props.allocType = cudaMemAllocationTypePinned;
  props.handleTypes = cudaMemHandleTypeNone;
  props.location.type = cudaMemLocationTypeDevice;
  props.location.id = device;
  if (cudaMemPoolCreate(&c->pool, &props) != cudaSuccess) {
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
    return NDTB_ERR_CUDA;
  }
  uint64_t thr = UINT64_MAX;
  cudaMemPoolSetAttribute(c->pool, cudaMemPoolAttrReleaseThreshold, &thr);
  {
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && total_b > 0) c->cache_limit = total_b / 8;  // released slabs parked per context
    if (const char *e = std::getenv("NDTB_SLAB_CACHE_MB")) c->cache_limit = (size_t)std::atoll(e) << 20;
  }
  if (match_kernel_prepare(c->smem_optin) != cudaSuccess) {
    cudaMemPoolDestroy(c->pool);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
    return NDTB_ERR_CUDA;
  }
  *out = c;
  return NDTB_OK;
}

void ndtb_ctx_destroy(ndtb_ctx *ctx) {
  DeviceGuard dev_guard(ctx);
  if (!ctx) return;
  cudaStreamSynchronize(ctx->stream);
  if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream), cudaStreamDestroy(ctx->copy_stream);
  for (auto &kv : ctx->slab_cache)
    for (auto &c : kv.second) cudaFreeAsync(c.p, ctx->stream);
  ctx->slab_cache.clear();
  cudaStreamSynchronize(ctx->stream);
  if (ctx->stg_host) {
    for (cudaEvent_t e : ctx->stg_ev)
      if (e) cudaEventDestroy(e);
    cudaFreeHost(ctx->stg_host);
  }
  for (cudaEvent_t e : ctx->copy_events) cudaEventDestroy(e);
  if (ctx->aux_stream) cudaStreamSynchronize(ctx->aux_stream), cudaStreamDestroy(ctx->aux_stream);
  for (cudaEvent_t e : ctx->aux_ev)
    if (e) cudaEventDestroy(e);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  if (ctx->pool) cudaMemPoolDestroy(ctx->pool);
  delete ctx;
}

int ndtb_ctx_synchronize(ndtb_ctx *ctx) {
  DeviceGuard dev_guard(ctx);
  if (!ctx) return NDTB_ERR_ARG;
  CU_TRY(ctx, timed_sync(ctx->stream));
  return NDTB_OK;
}
int64_t ndtb_ctx_launch_count(const ndtb_ctx *ctx) { return ctx ? ctx->launches : 0; }
int ndtb_ctx_enable_timing(ndtb_ctx *ctx, int on) {
  if (!ctx) return NDTB_ERR_ARG;
  ctx->timing = on != 0;
  return NDTB_OK;
}
int ndtb_ctx_match_time(ndtb_ctx *ctx, double *ms, int64_t *launches) {
  DeviceGuard dev_guard(ctx);
  if (!ctx || !ms) return NDTB_ERR_ARG;
  CU_TRY(ctx, timed_sync(ctx->stream));
  double tot = 0;
  for (auto &pr : ctx->timed) {
    float t = 0;
    CU_TRY(ctx, cudaEventElapsedTime(&t, pr.first, pr.second));
    tot += t;
    cudaEventDestroy(pr.first), cudaEventDestroy(pr.second);
  }
  *ms = tot;
  if (launches) *launches = (int64_t)ctx->timed.size();
  ctx->timed.clear();
  return NDTB_OK;
}
int ndtb_ctx_build_time(ndtb_ctx *ctx, double *ms, int64_t *calls) {
  DeviceGuard dev_guard(ctx);
  if (!ctx || !ms) return NDTB_ERR_ARG;
  CU_TRY(ctx, timed_sync(ctx->stream));
  double tot = 0;
  for (auto &pr : ctx->timed_build) {
    float t = 0;
    CU_TRY(ctx, cudaEventElapsedTime(&t, pr.first, pr.second));
    tot += t;
    cudaEventDestroy(pr.first), cudaEventDestroy(pr.second);
  }
  *ms = tot;
  if (calls) *calls = (int64_t)ctx->timed_build.size();
  ctx->timed_build.clear();
  return NDTB_OK;
}
int ndtb_ctx_sm_count(const ndtb_ctx *ctx) { return ctx ? ctx->sm_count : 0; }

void ndtb_default_params(ndtb_params *p) {
  p->n_neighbours = 2;
  p->itr_max = 30;
  p->step_control = 1;
  p->regularize = 1;
  p->delta_score = 10e-3 * 0.1;  // upstream init(): DELTA_SCORE = 10e-3 * current_resolution(0.1)
  p->lfd1 = 1;
  p->lfd2 = 0.05;
  p->use_soft_constraints = 0;
  p->use_tikhonov = 0;
  p->ctas_per_match = 0;
  p->pass_budget = 0;
  p->planar = 0;
  p->reserved_ = 0;
}

// ---- maps
int ndtb_map_create(ndtb_ctx *ctx, double cx, double cy, double cz, ndtb_map **out) {
  if (!ctx || !out || !(cx > 0 && cy > 0 && cz > 0)) return NDTB_ERR_ARG;
  ndtb_map *m = new ndtb_map();
  m->ctx = ctx;
  m->cell[0] = cx, m->cell[1] = cy, m->cell[2] = cz;
  std::memset(&m->g, 0, sizeof m->g);
  *out = m;
  return NDTB_OK;
}
void ndtb_map_destroy(ndtb_map *m) {
  DeviceGuard dev_guard(m ? m->ctx : nullptr);
  delete m;
}

int ndtb_map_guess_size(ndtb_map *m, double cx, double cy, double cz, double sx, double sy, double sz) {
  if (!m) return NDTB_ERR_ARG;
  m->guess_size = false;  // NDTMap::guessSize takes floats upstream
  m->centerx = (float)cx, m->centery = (float)cy, m->centerz = (float)cz;
  m->map_sizex = (float)sx, m->map_sizey = (float)sy, m->map_sizez = (float)sz;
  return NDTB_OK;
}
int ndtb_map_set_map_size(ndtb_map *m, double sx, double sy, double sz) {
  if (!m) return NDTB_ERR_ARG;
  m->map_sizex = (float)sx, m->map_sizey = (float)sy, m->map_sizez = (float)sz;
  return NDTB_OK;
}
int ndtb_map_initialize(ndtb_map *m, double cx, double cy, double cz, double sx, double sy, double sz) {
  if (!m) return NDTB_ERR_ARG;
  m->is_first_load = false;
  m->guess_size = false;
  m->centerx = cx, m->centery = cy, m->centerz = cz;
  m->map_sizex = sx, m->map_sizey = sy, m->map_sizez = sz;
  m->set_grid(cx, cy, cz, sx, sy, sz);
  if (m->nblocks() <= 0 || m->nblocks() * 64 >= ((int64_t)1 << 31)) return NDTB_ERR_GRID;
  return NDTB_OK;
}

int ndtb_map_load_point_cloud(ndtb_map *m, const float *pts, int64_t n, double range_limit, int mem, int64_t *n_binned) {
  DeviceGuard dev_guard(m ? m->ctx : nullptr);
  if (!m || n < 0 || n > 0x7fffffff || (n > 0 && !pts)) return NDTB_ERR_ARG;
  ndtb_ctx *ctx = m->ctx;
  SlabP buf;
  if (int rc = stage_points(ctx, pts, n, mem, buf)) return rc;
  m->pending.clear();
  m->pending.push_back({buf, (int)n});
  m->pending_load = true;
  m->pending_range = range_limit;
  m->pending_built = false;
  m->range_origin[0] = m->range_origin[1] = m->range_origin[2] = 0.0;
  std::vector<ndtb_map *> maps{m};
  std::vector<PointSrc> ps{{(const float4 *)buf->p, (int)n}};
  std::vector<char> empty;
  if (int rc = define_grids(ctx, maps, ps, {range_limit}, empty)) return rc;
  if (n_binned) {
    // the count needs the binning: build now with the default limits; computeNDTCells with the same
    // occupancy limit is then a no-op (maxnumpoints only matters when merging into existing cells)
    if (int rc = build_batch(ctx, maps, ps, {2}, {range_limit}, 0xffffffffu, 255.f)) return rc;
    m->pending_built = true;
    m->pending_occ = 255.f;
    *n_binned = m->last_binned;
  }
  return NDTB_OK;
}

int ndtb_map_load_point_cloud_centroid(ndtb_map *m, const float *pts, int64_t n, int mem, const double *origin, const double *old_centroid,
                                       const double *map_size, double range_limit) {
  DeviceGuard dev_guard(m ? m->ctx : nullptr);
  if (!m || !origin || !old_centroid || !map_size || n < 0 || n > 0x7fffffff || (n > 0 && !pts)) return NDTB_ERR_ARG;
  ndtb_ctx *ctx = m->ctx;
  SlabP buf;
  if (int rc = stage_points(ctx, pts, n, mem, buf)) return rc;
  // how many cells towards the new origin the centre moves from the old centroid [upstream]
  double c[3];
  for (int a = 0; a < 3; a++) c[a] = old_centroid[a] + std::floor((origin[a] - old_centroid[a]) / m->cell[a]) * m->cell[a];
  m->guess_size = false, m->is_first_load = false;
  m->centerx = c[0], m->centery = c[1], m->centerz = c[2];
  m->map_sizex = map_size[0], m->map_sizey = map_size[1], m->map_sizez = map_size[2];
  m->set_grid(c[0], c[1], c[2], map_size[0], map_size[1], map_size[2]);
  if (m->nblocks() <= 0 || m->nblocks() * 64 >= ((int64_t)1 << 31)) return NDTB_ERR_GRID;
  m->pending.clear();
  m->pending.push_back({buf, (int)n});
  m->pending_load = true, m->pending_built = false;
  m->pending_range = range_limit;
  for (int a = 0; a < 3; a++) m->range_origin[a] = origin[a];
  return NDTB_OK;
}

int ndtb_map_add_points(ndtb_map *m, const float *pts, int64_t n, int mem, int64_t *n_binned) {
  DeviceGuard dev_guard(m ? m->ctx : nullptr);
  if (!m || n < 0 || n > 0x7fffffff || (n > 0 && !pts)) return NDTB_ERR_ARG;
  if (!m->grid_ready) return NDTB_ERR_GRID;
  ndtb_ctx *ctx = m->ctx;
  SlabP buf;
  if (int rc = stage_points(ctx, pts, n, mem, buf)) return rc;
  if (m->pending_load && m->pending_built) {  // a loadPointCloud whose cells were already computed
    m->pending.clear();
    m->pending_load = false;
  }
  m->pending.push_back({buf, (int)n});
  if (n_binned) {
    *n_binned = 0;
    if (n > 0) {
      const int64_t r = ndtb_map_point_indices(m, (const float *)buf->p, n, NDTB_MEM_DEVICE + 1, nullptr);
      if (r < 0) return (int)r;
      *n_binned = r;
    }
  }
  return NDTB_OK;
}

int ndtb_map_add_point_cloud(ndtb_map *m, const double *origin, const float *pts, int64_t n, int mem, double classifier_th,
                             double maxz, double sensor_noise, double occupancy_limit) {
  DeviceGuard dev_guard(m ? m->ctx : nullptr);
  (void)classifier_th;  // unused by the upstream branch restated here
  if (!m || !origin || n < 0 || n > 0x7fffffff || (n > 0 && !pts)) return NDTB_ERR_ARG;
  if (m->is_first_load) return ndtb_map_load_point_cloud(m, pts, n, -1.0, mem, nullptr);  // NDTMap::addPointCloud: isFirstLoad_
  if (!m->grid_ready) return NDTB_ERR_GRID;
  ndtb_ctx *ctx = m->ctx;
  SlabP buf;
  if (int rc = stage_points(ctx, pts, n, mem, buf)) return rc;
  if (m->pending_load && m->pending_built) {
    m->pending.clear();
    m->pending_load = false;
  }
  int traced = 0;
  for (auto &c : m->pending) traced += c.trace;
  if (traced >= MAX_TRACE_SEG) return NDTB_ERR_ARG;
  ndtb_map::Chunk c;
  c.buf = buf, c.n = (int)n, c.trace = true;
  for (int i = 0; i < 3; i++) c.origin[i] = origin[i];
  c.maxz = maxz, c.sensor_noise = sensor_noise, c.occ_limit = (float)occupancy_limit;
  m->pending.push_back(c);
  return NDTB_OK;
}

int ndtb_map_compute_cells(ndtb_map *m, uint32_t maxnumpoints, float occupancy_limit) {
  DeviceGuard dev_guard(m ? m->ctx : nullptr);
  if (!m) return NDTB_ERR_ARG;
  ndtb_ctx *ctx = m->ctx;
  if (m->pending.empty()) return NDTB_OK;
  if (m->pending_load && m->pending_built && occupancy_limit == m->pending_occ) {
    m->pending.clear();
    m->pending_load = false;
    return NDTB_OK;
  }
  // concatenate the pending chunks in arrival order (= NDTCell::points_ insertion order)
  int64_t total = 0;
  for (auto &c : m->pending) total += c.n;
  if (total > 0x7fffffff) return NDTB_ERR_ARG;
  SlabP all;
  const float4 *ptr;
  if (m->pending.size() == 1) {
    all = m->pending[0].buf;
    ptr = (const float4 *)all->p;
  } else {
    if (int rc = slab_alloc(ctx, 16 * (size_t)std::max<int64_t>(total, 1), all)) return rc;
    size_t off = 0;
    for (auto &c : m->pending) {
      CU_TRY(ctx, cudaMemcpyAsync(all->p + off, c.buf->p, 16 * (size_t)c.n, cudaMemcpyDeviceToDevice, ctx->stream));
      off += 16 * (size_t)c.n;
    }
    ptr = (const float4 *)all->p;
  }
  const char load = m->pending_load ? 2 : 0;
  std::vector<ndtb_map *> maps{m};
  std::vector<PointSrc> ps{{ptr, (int)total}};
  std::vector<TraceSegs> segs(1);
  {
    int at = 0;
    for (auto &c : m->pending) {
      if (c.trace) {
        TraceSeg sg;
        for (int i = 0; i < 3; i++) sg.origin[i] = c.origin[i];
        sg.maxz = c.maxz, sg.sensor_noise = c.sensor_noise, sg.occ_limit = c.occ_limit, sg.begin = at, sg.end = at + c.n;
        segs[0].push_back(sg);
      }
      at += c.n;
    }
  }
  const int rc = build_batch(ctx, maps, ps, {load}, {m->pending_range}, maxnumpoints, occupancy_limit, &segs);
  m->pending.clear();
  m->pending_load = false;
  return rc;
}

int ndtb_map_build_batch(ndtb_ctx *ctx, int64_t n_maps, ndtb_map *const *maps, const float *const *pts, const int64_t *n_pts,
                         double range_limit, int mem, uint32_t maxnumpoints, float occupancy_limit) {
  DeviceGuard dev_guard(ctx);
  if (!ctx || n_maps < 0 || (n_maps > 0 && (!maps || !pts || !n_pts))) return NDTB_ERR_ARG;
  std::vector<ndtb_map *> mv((size_t)n_maps);
  std::vector<PointSrc> ps((size_t)n_maps);
  for (int64_t i = 0; i < n_maps; i++) {
    if (!maps[i] || n_pts[i] < 0 || n_pts[i] > 0x7fffffff) return NDTB_ERR_ARG;
    mv[i] = maps[i];
    if (mem == NDTB_MEM_DEVICE) ps[i] = {(const float4 *)pts[i], (int)n_pts[i]};
    maps[i]->pending.clear();
    maps[i]->pending_load = false;
  }
  PhaseTimer pt(ctx, "map_build_batch");
  if (mem == NDTB_MEM_DEVICE) {
    std::vector<char> load((size_t)n_maps, 1);
    std::vector<double> range((size_t)n_maps, range_limit);
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (ctx->timing) {
      CU_TRY(ctx, cudaEventCreate(&ev0));
      CU_TRY(ctx, cudaEventCreate(&ev1));
      CU_TRY(ctx, cudaEventRecord(ev0, ctx->stream));
    }
    const int rc = build_batch(ctx, mv, ps, load, range, maxnumpoints, occupancy_limit);
    if (ctx->timing) {
      CU_TRY(ctx, cudaEventRecord(ev1, ctx->stream));
      ctx->timed_build.push_back({ev0, ev1});
    }
    pt.mark("build_batch");
    return rc;
  }
  // Host scans: one slab; the H2D copies run on the copy stream in chunks of maps, chunk c+1 is in flight on PCIe
  // while the kernels (and the host synchronisations) of chunk c's build run on the context's stream.
  Carver c;
  std::vector<size_t> off((size_t)n_maps);
  size_t total = 0;
  for (int64_t i = 0; i < n_maps; i++) off[i] = c.take(16 * (size_t)n_pts[i]), total += 16 * (size_t)n_pts[i];
  SlabP big;
  if (int rc = slab_alloc(ctx, c.off, big)) return rc;
  for (int64_t i = 0; i < n_maps; i++) ps[i] = {(const float4 *)(big->p + off[i]), (int)n_pts[i]};
  int n_chunks = 1;
  static const int env_h2d_chunks = std::getenv("NDTB_H2D_CHUNKS") ? std::atoi(std::getenv("NDTB_H2D_CHUNKS")) : 0;
  if (n_maps >= 16 && total >= ((size_t)32 << 20)) n_chunks = (int)std::min<int64_t>(env_h2d_chunks > 0 ? env_h2d_chunks : (n_maps >= 512 ? 4 : 8), n_maps / 8);  // B200, 1184 maps, 3 lanes: 2 chunks 9.06 k, 4: 9.33 k, 8: 9.15 k, 16: 8.46 k reg/s
  if (n_chunks > 1 && !ctx->copy_stream) CU_TRY(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  while ((int)ctx->copy_events.size() < n_chunks + 1) {
    cudaEvent_t e;
    CU_TRY(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ctx->copy_events.push_back(e);
  }
  cudaStream_t cs = n_chunks > 1 ? ctx->copy_stream : ctx->stream;
  if (n_chunks > 1) {  // the slab is stream-ordered on ctx->stream: the copy stream may touch it only after its allocation
    CU_TRY(ctx, cudaEventRecord(ctx->copy_events[n_chunks], ctx->stream));
    CU_TRY(ctx, cudaStreamWaitEvent(cs, ctx->copy_events[n_chunks], 0));
  }
  std::vector<int64_t> cbeg((size_t)n_chunks + 1, 0);
  {  // chunk boundaries at (roughly) equal bytes
    size_t acc = 0;
    int cc = 1;
    for (int64_t i = 0; i < n_maps && cc < n_chunks; i++) {
      acc += 16 * (size_t)n_pts[i];
      if (acc * n_chunks >= total * cc) cbeg[cc++] = i + 1;
    }
    for (; cc <= n_chunks; cc++) cbeg[cc] = n_maps;
  }
  static const bool prof_copies = std::getenv("NDTB_PROFILE") != nullptr;
  std::vector<cudaEvent_t> tev;
  if (prof_copies) {
    tev.resize((size_t)n_chunks + 1);
    for (auto &e : tev) cudaEventCreate(&e);
    cudaEventRecord(tev[0], cs);
  }
  const double t_enq = now_ms();
  for (int k = 0; k < n_chunks; k++) {
    for (int64_t i = cbeg[k]; i < cbeg[k + 1]; i++)
      if (n_pts[i] > 0) CU_TRY(ctx, cudaMemcpyAsync(big->p + off[i], pts[i], 16 * (size_t)n_pts[i], cudaMemcpyHostToDevice, cs));
    if (n_chunks > 1) CU_TRY(ctx, cudaEventRecord(ctx->copy_events[k], cs));
    if (prof_copies) cudaEventRecord(tev[(size_t)k + 1], cs);
  }
  g_slow.enq_ms += now_ms() - t_enq;
  pt.mark("enqueue H2D");
  int rc = NDTB_OK;
  for (int k = 0; k < n_chunks && rc == NDTB_OK; k++) {
    const int64_t b = cbeg[k], e = cbeg[k + 1];
    if (e <= b) continue;
    if (n_chunks > 1) CU_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->copy_events[k], 0));
    std::vector<ndtb_map *> cm(mv.begin() + b, mv.begin() + e);
    std::vector<PointSrc> cp(ps.begin() + b, ps.begin() + e);
    std::vector<char> load((size_t)(e - b), 1);
    std::vector<double> range((size_t)(e - b), range_limit);
    rc = build_batch(ctx, cm, cp, load, range, maxnumpoints, occupancy_limit);
  }
  if (rc != NDTB_OK && n_chunks > 1) cudaStreamSynchronize(cs);  // the slab must outlive the copies in flight
  pt.mark("build_batch");
  if (prof_copies) {
    cudaStreamSynchronize(cs);
    for (int k = 0; k < n_chunks; k++) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, tev[0], tev[(size_t)k + 1]);
      std::fprintf(stderr, "[ndtb] H2D chunk %d done at %.3f ms (maps %lld..%lld)\n", k, ms, (long long)cbeg[k], (long long)cbeg[k + 1]);
    }
    for (auto &e : tev) cudaEventDestroy(e);
  }
  return rc;
}

int ndtb_map_from_cells(ndtb_map *m, const ndtb_grid *g, const ndtb_cell *cells, int64_t n, int use_idx) {
  DeviceGuard dev_guard(m ? m->ctx : nullptr);
  if (!m || !g || n < 0 || (n > 0 && !cells)) return NDTB_ERR_ARG;
  ndtb_ctx *ctx = m->ctx;
  cudaStream_t st = ctx->stream;
  for (int i = 0; i < 3; i++) {
    m->cell[i] = g->cell[i];
    m->g.cell[i] = g->cell[i], m->g.center[i] = g->center[i], m->g.size[i] = g->size[i], m->g.nb[i] = (g->size[i] + 3) / 4;
  }
  for (int i = 0; i < 3; i++) m->boff[i] = 0, m->nbs[i] = m->g.nb[i];
  m->guess_size = false, m->is_first_load = false, m->grid_ready = true;
  m->drop_cells();
  m->pending.clear();
  if (m->nblocks() <= 0 || m->nblocks() * 64 >= ((int64_t)1 << 31)) return NDTB_ERR_GRID;
  m->nblk = (int)m->nblocks();
  Carver cb, ct;
  const size_t o_amask = cb.take(8 * (size_t)m->nblk), o_abase = cb.take(4 * (size_t)m->nblk);
  const size_t o_tbl = cb.take(4 * (size_t)std::max<int64_t>(std::min<int64_t>(m->nblk, n), 1)), o_counts = cb.take(32);
  const size_t o_cells = ct.take(sizeof(ndtb_cell) * (size_t)std::max<int64_t>(n, 1)), o_vox = ct.take(4 * (size_t)std::max<int64_t>(n, 1));
  const size_t o_err = ct.take(4), o_job = ct.take(sizeof(BuildJob));
  SlabP s_b, s_t;
  if (int rc = slab_alloc(ctx, cb.off, s_b)) return rc;
  if (int rc = slab_alloc(ctx, ct.off, s_t)) return rc;
  CU_TRY(ctx, cudaMemsetAsync(s_b->p, 0, s_b->bytes, st));
  CU_TRY(ctx, cudaMemsetAsync(s_t->p + o_err, 0, 4, st));
  if (n > 0) CU_TRY(ctx, h2d_small(ctx, s_t->p + o_cells, cells, sizeof(ndtb_cell) * (size_t)n, st));
  BuildJob j;
  std::memset(&j, 0, sizeof j);
  j.g = m->g, j.nblk = m->nblk;
  m->fill_box(j);
  j.amask = (unsigned long long *)(s_b->p + o_amask), j.abase = (int *)(s_b->p + o_abase);
  j.tb_list = (int *)(s_b->p + o_tbl), j.counts = (int *)(s_b->p + o_counts);
  BuildJob *d_job = (BuildJob *)(s_t->p + o_job);
  const ndtb_cell *d_cells = (const ndtb_cell *)(s_t->p + o_cells);
  int *d_vox = (int *)(s_t->p + o_vox), *d_err = (int *)(s_t->p + o_err);
  CU_TRY(ctx, h2d_small(ctx, d_job, &j, sizeof j, st));
  ctx->launches += launch_from_cells_voxel(d_job, d_cells, (int)n, use_idx, d_vox, d_err, st);
  ctx->launches += launch_blockscan(d_job, 1, st);
  int cnts[8], err = 0;
  CU_TRY(ctx, cudaMemcpyAsync(cnts, j.counts, 32, cudaMemcpyDeviceToHost, st));
  CU_TRY(ctx, cudaMemcpyAsync(&err, d_err, 4, cudaMemcpyDeviceToHost, st));
  CU_TRY(ctx, timed_sync(st));
  if (err) return NDTB_ERR_ARG;  // a cell outside the grid (the CPU restatement returns -1)
  const int n_all = cnts[0], ntb = cnts[1];
  int tsize = 2;
  while (tsize < 2 * ntb) tsize <<= 1;
  Carver cc;
  const size_t na = (size_t)std::max(n_all, 1), nt = (size_t)std::max(ntb, 1);
  const size_t o_mean = cc.take(24 * na), o_cov = cc.take(72 * na), o_n = cc.take(4 * na), o_has = cc.take(4 * na);
  const size_t o_occ = cc.take(4 * na), o_gcell = cc.take(72 * na), o_g2c = cc.take(4 * na);
  const size_t o_table = cc.take(sizeof(HashEntry) * (size_t)tsize), o_gm = cc.take(8 * nt), o_gb = cc.take(4 * nt);
  SlabP s_c;
  if (int rc = slab_alloc(ctx, cc.off, s_c)) return rc;
  j.n_all = n_all, j.tsize = tsize;
  j.cmean = (double *)(s_c->p + o_mean), j.ccov = (double *)(s_c->p + o_cov), j.cn = (int *)(s_c->p + o_n);
  j.chas = (int *)(s_c->p + o_has), j.cocc = (float *)(s_c->p + o_occ), j.gcell = (double *)(s_c->p + o_gcell);
  j.g2c = (int *)(s_c->p + o_g2c), j.table = (HashEntry *)(s_c->p + o_table);
  j.gmask_t = (unsigned long long *)(s_c->p + o_gm), j.gbase_t = (int *)(s_c->p + o_gb);
  CU_TRY(ctx, cudaMemsetAsync(j.table, 0xFF, sizeof(HashEntry) * (size_t)tsize, st));
  CU_TRY(ctx, h2d_small(ctx, d_job, &j, sizeof j, st));
  ctx->launches += launch_from_cells_place(d_job, d_cells, (int)n, d_vox, st);
  ctx->launches += launch_gview(d_job, 1, std::max(ntb, 1), std::max(n_all, 1), st);
  CU_TRY(ctx, cudaMemcpyAsync(cnts, j.counts, 32, cudaMemcpyDeviceToHost, st));
  CU_TRY(ctx, timed_sync(st));
  m->s_blocks = s_b, m->s_cells = s_c;
  m->amask = j.amask, m->abase = j.abase, m->tb_list = j.tb_list, m->counts = j.counts;
  m->n_all = n_all, m->ntb = ntb, m->ng = cnts[2], m->ngb = cnts[3];
  m->cmean = j.cmean, m->ccov = j.ccov, m->cn = j.cn, m->chas = j.chas, m->cocc = j.cocc;
  m->gcell = j.gcell, m->g2c = j.g2c, m->table = j.table, m->tsize = tsize;
  return NDTB_OK;
}

int ndtb_map_grid(const ndtb_map *m, ndtb_grid *g) {
  if (!m || !g) return NDTB_ERR_ARG;
  if (!m->grid_ready) return NDTB_ERR_GRID;
  for (int i = 0; i < 3; i++) g->center[i] = m->g.center[i], g->cell[i] = m->g.cell[i], g->size[i] = m->g.size[i];
  return NDTB_OK;
}

int64_t ndtb_map_num_cells(const ndtb_map *m, int gaussian_only) {
  if (!m) return NDTB_ERR_ARG;
  return gaussian_only ? m->ng : m->n_all;
}

int64_t ndtb_map_export_cells(const ndtb_map *m, ndtb_cell *out, int64_t cap, int gaussian_only) {
  DeviceGuard dev_guard(m ? m->ctx : nullptr);
  if (!m || cap < 0 || (cap > 0 && !out)) return NDTB_ERR_ARG;
  ndtb_ctx *ctx = m->ctx;
  if (m->n_all == 0) return 0;
  SlabP tmp;
  if (int rc = slab_alloc(ctx, sizeof(ndtb_cell) * (size_t)m->n_all + sizeof(BuildJob) + 256, tmp)) return rc;
  BuildJob j;
  std::memset(&j, 0, sizeof j);
  j.g = m->g, j.amask = m->amask, j.abase = m->abase, j.tb_list = m->tb_list, j.counts = m->counts;
  m->fill_box(j);
  j.cmean = m->cmean, j.ccov = m->ccov, j.cn = m->cn, j.chas = m->chas, j.cocc = m->cocc;
  BuildJob *d_job = (BuildJob *)(tmp->p + ((sizeof(ndtb_cell) * (size_t)m->n_all + 255) & ~(size_t)255));
  CU_TRY(ctx, h2d_small(ctx, d_job, &j, sizeof j, ctx->stream));
  ctx->launches += launch_export(d_job, m->ntb, (ndtb_cell *)tmp->p, ctx->stream);
  std::vector<ndtb_cell> h((size_t)m->n_all);
  CU_TRY(ctx, cudaMemcpyAsync(h.data(), tmp->p, sizeof(ndtb_cell) * (size_t)m->n_all, cudaMemcpyDeviceToHost, ctx->stream));
  CU_TRY(ctx, timed_sync(ctx->stream));
  std::vector<std::pair<int64_t, int>> ord;
  for (int i = 0; i < m->n_all; i++) {
    if (gaussian_only && !h[i].has_gaussian) continue;
    ord.push_back({((int64_t)h[i].idx[0] * m->g.size[1] + h[i].idx[1]) * m->g.size[2] + h[i].idx[2], i});
  }
  std::sort(ord.begin(), ord.end());
  for (size_t k = 0; k < ord.size() && (int64_t)k < cap; k++) out[k] = h[ord[k].second];
  return (int64_t)ord.size();
}

int64_t ndtb_map_point_indices(const ndtb_map *m, const float *pts, int64_t n, int mem, int32_t *out) {
  DeviceGuard dev_guard(m ? m->ctx : nullptr);
  const bool count_only = mem == NDTB_MEM_DEVICE + 1;  // internal: device points, no index output
  if (count_only) mem = NDTB_MEM_DEVICE;
  if (!m || n < 0 || n > 0x7fffffff || (n > 0 && (!pts || (!out && !count_only)))) return NDTB_ERR_ARG;
  if (!m->grid_ready) return NDTB_ERR_GRID;
  ndtb_ctx *ctx = m->ctx;
  if (n == 0) return 0;
  SlabP pbuf, obuf;
  PhaseTimer pt(ctx, "point_indices");
  const float4 *d_pts = (const float4 *)pts;
  if (mem != NDTB_MEM_DEVICE) {
    if (int rc = stage_points(ctx, pts, n, mem, pbuf)) return rc;
    d_pts = (const float4 *)pbuf->p;
  }
  pt.mark("stage");
  if (int rc = slab_alloc(ctx, 12 * (size_t)n + 256, obuf)) return rc;
  int *d_nin = (int *)(obuf->p + ((12 * (size_t)n + 255) & ~(size_t)255));
  int *d_out = (mem == NDTB_MEM_DEVICE && !count_only) ? out : (int *)obuf->p;
  CU_TRY(ctx, cudaMemsetAsync(d_nin, 0, 4, ctx->stream));
  ctx->launches += launch_point_indices(m->g, d_pts, (int)n, d_out, d_nin, ctx->stream);
  pt.mark("kernel");
  int nin = 0;
  if (mem != NDTB_MEM_DEVICE) CU_TRY(ctx, cudaMemcpyAsync(out, d_out, 12 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  CU_TRY(ctx, cudaMemcpyAsync(&nin, d_nin, 4, cudaMemcpyDeviceToHost, ctx->stream));
  CU_TRY(ctx, timed_sync(ctx->stream));
  return nin;
}

// ---- matcher
int ndtb_d2d_derivatives(ndtb_ctx *ctx, const ndtb_map *tgt, const ndtb_map *src, const double *T, const ndtb_params *p,
                         int want_hessian, double *out43, int64_t *n_pairs) {
  DeviceGuard dev_guard(ctx);
  if (!ctx || !tgt || !src || !T || !p || !out43) return NDTB_ERR_ARG;
  if (!map_ok(tgt) || !map_ok(src)) return NDTB_ERR_GRID;
  if (tgt->ctx != ctx || src->ctx != ctx) return NDTB_ERR_ARG;  // storage is released in the order of its own context's stream
  if (int rc = ensure_view(ctx, const_cast<ndtb_map *>(tgt))) return rc;
  if (int rc = ensure_view(ctx, const_cast<ndtb_map *>(src))) return rc;
  MatchJob j;
  fill_job(j, tgt, src, T, nullptr);
  MatchConfig cfg = make_config(ctx, p, 2);
  const int W = acc_total();
  int n_ctas = std::max(1, std::min(ctx->sm_count, (src->ng + 255) / 256));
  SlabP s;
  if (int rc = slab_alloc(ctx, sizeof(MatchJob) + 256 + 8 * (size_t)W * (n_ctas + 1), s)) return rc;
  MatchJob *d_job = (MatchJob *)s->p;
  double *d_part = (double *)(s->p + ((sizeof(MatchJob) + 255) & ~(size_t)255));
  double *d_out = d_part + (size_t)W * n_ctas;
  CU_TRY(ctx, h2d_small(ctx, d_job, &j, sizeof j, ctx->stream));
  CU_TRY(ctx, launch_derivatives(d_job, cfg, want_hessian != 0, n_ctas, d_part, d_out, ctx->stream));
  ctx->launches += 2;
  std::vector<double> h((size_t)W);
  CU_TRY(ctx, cudaMemcpyAsync(h.data(), d_out, 8 * (size_t)W, cudaMemcpyDeviceToHost, ctx->stream));
  CU_TRY(ctx, timed_sync(ctx->stream));
  out43[0] = h[0];
  for (int i = 0; i < 6; i++) out43[1 + i] = h[ACC_G + i];
  for (int a = 0; a < 6; a++)
    for (int b = a; b < 6; b++) out43[7 + a * 6 + b] = out43[7 + b * 6 + a] = want_hessian ? h[hidx(a, b)] : 0.0;
  if (n_pairs) *n_pairs = (int64_t)h[28];
  return NDTB_OK;
}

namespace {
// host NDTCell copies -> a source-only pseudo map (compact Gaussian records, no table: sources are never probed)
int cells_source(ndtb_ctx *ctx, const ndtb_cell *cells, int64_t n, std::unique_ptr<ndtb_map> &out) {
  if (n < 0 || n > 0x7fffffff || (n > 0 && !cells)) return NDTB_ERR_ARG;
  std::vector<double> g;
  g.reserve(9 * (size_t)n);
  for (int64_t i = 0; i < n; i++) {
    if (!cells[i].has_gaussian) continue;
    for (int q = 0; q < 3; q++) g.push_back(cells[i].mean[q]);
    for (int q = 0; q < 6; q++) g.push_back(cells[i].cov[q]);
  }
  const size_t ng = g.size() / 9;
  SlabP cbuf;
  if (int rc = slab_alloc(ctx, 256 + 72 * std::max<size_t>(ng, 1), cbuf)) return rc;
  CU_TRY(ctx, cudaMemsetAsync(cbuf->p, 0xFF, 256, ctx->stream));
  if (ng > 0) {
    CU_TRY(ctx, h2d_small(ctx, cbuf->p + 256, g.data(), 72 * ng, ctx->stream));
    CU_TRY(ctx, timed_sync(ctx->stream));  // g is a local vector
  }
  out.reset(new ndtb_map());
  ndtb_map *m = out.get();
  m->ctx = ctx;
  m->cell[0] = m->cell[1] = m->cell[2] = 1.0;
  std::memset(&m->g, 0, sizeof m->g);
  m->grid_ready = true;
  m->s_cells = cbuf;
  m->table = (HashEntry *)cbuf->p, m->tsize = 2;
  m->gcell = (double *)(cbuf->p + 256), m->ng = (int)ng;
  return NDTB_OK;
}
}  // namespace

int ndtb_d2d_derivatives_cells(ndtb_ctx *ctx, const ndtb_map *tgt, const ndtb_cell *src, int64_t n, const double *T,
                               const ndtb_params *p, int want_hessian, double *out43, int64_t *n_pairs) {
  DeviceGuard dev_guard(ctx);
  if (!ctx || !tgt || !p || !out43) return NDTB_ERR_ARG;
  std::unique_ptr<ndtb_map> s;
  if (int rc = cells_source(ctx, src, n, s)) return rc;
  const double I16[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  return ndtb_d2d_derivatives(ctx, tgt, s.get(), T ? T : I16, p, want_hessian, out43, n_pairs);
}

int ndtb_d2d_line_search_cells(ndtb_ctx *ctx, const ndtb_map *tgt, const ndtb_cell *src, int64_t n, double *increment6,
                               const ndtb_params *p, double *step) {
  DeviceGuard dev_guard(ctx);
  if (!ctx || !tgt || !p || !increment6 || !step) return NDTB_ERR_ARG;
  std::unique_ptr<ndtb_map> sm;
  if (int rc = cells_source(ctx, src, n, sm)) return rc;
  // the optimiser's own More-Thuente state machine (csrc/optimizer.h), driven from the host: one gradient pass per trial
  OptParams prm;
  std::memset(&prm, 0, sizeof prm);
  prm.itr_max = p->itr_max, prm.step_control = 1, prm.regularize = p->regularize, prm.delta_score = p->delta_score;
  OptState st;
  const double I16[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  opt_begin(st, prm, I16);
  st.ls_only = 1;
  for (int i = 0; i < 6; i++) st.incr[i] = increment6[i];
  double out43[43], sums[29];
  auto eval = [&](const Pose &P) -> int {
    double T16[16];
    pose_to_cm(P, T16);
    if (int rc = ndtb_d2d_derivatives(ctx, tgt, sm.get(), T16, p, 0, out43, nullptr)) return rc;
    for (int i = 0; i < 29; i++) sums[i] = 0.0;
    for (int i = 0; i < 7; i++) sums[i] = out43[i];
    return NDTB_OK;
  };
  if (int rc = eval(st.T)) return rc;
  st.ls_soft = 0;
  on_ls_init(st, prm, sums);
  int guard = 0;
  while (st.phase == PH_LS_EVAL && guard++ < 64) {
    if (int rc = eval(st.Peval)) return rc;
    opt_advance(st, prm, sums);
  }
  for (int i = 0; i < 6; i++) increment6[i] = st.incr[i];
  *step = st.ls_result;
  return NDTB_OK;
}

int ndtb_mt_cstep(double *stx, double *fx, double *dx, double *sty, double *fy, double *dy, double *stp, double fp, double dp,
                  int *brackt, double stmin, double stmax) {
  if (!stx || !fx || !dx || !sty || !fy || !dy || !stp || !brackt) return NDTB_ERR_ARG;
  int b = *brackt != 0;
  const int info = mt_cstep(*stx, *fx, *dx, *sty, *fy, *dy, *stp, fp, dp, b, stmin, stmax);
  *brackt = b;
  return info;
}

int ndtb_eig_sym3(const double *A9, int stop_at_fixed_point, double *evals3, double *V9, int32_t *sweeps) {
  if (!A9 || !evals3 || !V9) return NDTB_ERR_ARG;
  int n = 0;
  const bool ok = stop_at_fixed_point ? eig_sym_n<3, true>(A9, evals3, V9, 64, &n) : eig_sym_n<3, false>(A9, evals3, V9, 64, &n);
  if (sweeps) *sweeps = n;
  return ok ? 1 : 0;
}

int ndtb_d2d_match(ndtb_ctx *ctx, const ndtb_map *tgt, const ndtb_map *src, const double *T0, const ndtb_params *p,
                   ndtb_result *res) {
  DeviceGuard dev_guard(ctx);
  if (!ctx || !tgt || !src || !T0 || !p || !res) return NDTB_ERR_ARG;
  return match_batch_impl(ctx, 1, &tgt, &src, T0, nullptr, p, 0, NDTB_MEM_HOST, res, nullptr);
}

int ndtb_fusion_match(ndtb_ctx *ctx, const ndtb_map *tgt, const ndtb_map *src, const double *T0, const double *Tcov36,
                      const ndtb_params *p, ndtb_result *res) {
  DeviceGuard dev_guard(ctx);
  if (!ctx || !tgt || !src || !T0 || !Tcov36 || !p || !res) return NDTB_ERR_ARG;
  double Q[36];
  if (!inv6(Tcov36, Q)) return NDTB_ERR_SINGULAR;
  return match_batch_impl(ctx, 1, &tgt, &src, T0, Q, p, 0, NDTB_MEM_HOST, res, nullptr);
}

int ndtb_d2d_match_batch(ndtb_ctx *ctx, int64_t n_edges, const ndtb_map *const *tgt, const ndtb_map *const *src,
                         const double *T0s, const ndtb_params *p, int with_covariance, int out_mem, ndtb_result *res,
                         double *cov36s) {
  DeviceGuard dev_guard(ctx);
  if (!ctx || n_edges < 0 || (n_edges > 0 && (!tgt || !src || !T0s || !res)) || !p) return NDTB_ERR_ARG;
  return match_batch_impl(ctx, n_edges, tgt, src, T0s, nullptr, p, with_covariance && cov36s, out_mem, res, cov36s);
}

int ndtb_d2d_covariance(ndtb_ctx *ctx, const ndtb_map *tgt, const ndtb_map *src, const double *T, const ndtb_params *p,
                        double *cov36) {
  DeviceGuard dev_guard(ctx);
  if (!ctx || !tgt || !src || !T || !p || !cov36) return NDTB_ERR_ARG;
  if (!map_ok(tgt) || !map_ok(src)) return NDTB_ERR_GRID;
  if (tgt->ctx != ctx || src->ctx != ctx) return NDTB_ERR_ARG;  // storage is released in the order of its own context's stream
  if (int rc = ensure_view(ctx, const_cast<ndtb_map *>(tgt))) return rc;
  if (int rc = ensure_view(ctx, const_cast<ndtb_map *>(src))) return rc;
  cudaStream_t st = ctx->stream;
  MatchJob j;
  fill_job(j, tgt, src, T, nullptr);
  MatchConfig cfg = make_config(ctx, p, 2);
  const int n_chunks = std::max(1, std::min(ctx->sm_count, (src->ng + 127) / 128));
  Carver c;
  const size_t o_job = c.take(sizeof j), o_goff = c.take(8), o_gt = c.take(48 * (size_t)std::max(tgt->ng, 1));
  const size_t o_part = c.take(8 * (size_t)cov_partial_width() * n_chunks), o_cov = c.take(288), o_stat = c.take(4);
  SlabP s;
  if (int rc = slab_alloc(ctx, c.off, s)) return rc;
  CU_TRY(ctx, cudaMemsetAsync(s->p, 0, s->bytes, st));
  CU_TRY(ctx, h2d_small(ctx, s->p + o_job, &j, sizeof j, st));
  CU_TRY(ctx, launch_covariance((MatchJob *)(s->p + o_job), 1, cfg, nullptr, (const long long *)(s->p + o_goff),
                                (double *)(s->p + o_gt), (double *)(s->p + o_part), n_chunks, (double *)(s->p + o_cov),
                                (int *)(s->p + o_stat), nullptr, 0, st));
  ctx->launches += 2;
  int status = 0;
  CU_TRY(ctx, cudaMemcpyAsync(cov36, s->p + o_cov, 288, cudaMemcpyDeviceToHost, st));
  CU_TRY(ctx, cudaMemcpyAsync(&status, s->p + o_stat, 4, cudaMemcpyDeviceToHost, st));
  CU_TRY(ctx, timed_sync(st));
  return status;
}

// ---- NDTMatcherP2D: the cloud becomes a list of zero-covariance cells that the D2D kernels consume
namespace {
int points_source(ndtb_ctx *ctx, const float *pts, int64_t n, int mem, std::unique_ptr<ndtb_map> &out) {
  if (n < 0 || n > 0x7fffffff || (n > 0 && !pts)) return NDTB_ERR_ARG;
  SlabP pbuf, cbuf;
  const float4 *d_pts = (const float4 *)pts;
  if (mem != NDTB_MEM_DEVICE) {
    if (int rc = stage_points(ctx, pts, n, mem, pbuf)) return rc;
    d_pts = (const float4 *)pbuf->p;
  }
  if (int rc = slab_alloc(ctx, 256 + 8 * (size_t)GC * (size_t)std::max<int64_t>(n, 1), cbuf)) return rc;
  CU_TRY(ctx, cudaMemsetAsync(cbuf->p, 0xFF, 256, ctx->stream));  // a 2-entry empty block table (never probed: sources have no table)
  if (n > 0) ctx->launches += launch_points_as_cells(d_pts, (int)n, (double *)(cbuf->p + 256), ctx->stream);
  out.reset(new ndtb_map());
  ndtb_map *m = out.get();
  m->ctx = ctx;
  m->cell[0] = m->cell[1] = m->cell[2] = 1.0;
  std::memset(&m->g, 0, sizeof m->g);
  m->grid_ready = true;
  m->s_cells = cbuf;
  m->table = (HashEntry *)cbuf->p, m->tsize = 2;
  m->gcell = (double *)(cbuf->p + 256), m->ng = (int)n;
  return NDTB_OK;
}
}  // namespace

int ndtb_p2d_derivatives(ndtb_ctx *ctx, const ndtb_map *tgt, const float *pts, int64_t n, int mem, const double *T,
                         const ndtb_params *p, int want_hessian, double *out43, int64_t *n_pairs) {
  DeviceGuard dev_guard(ctx);
  if (!ctx || !tgt || !T || !p || !out43) return NDTB_ERR_ARG;
  std::unique_ptr<ndtb_map> src;
  if (int rc = points_source(ctx, pts, n, mem, src)) return rc;
  return ndtb_d2d_derivatives(ctx, tgt, src.get(), T, p, want_hessian, out43, n_pairs);
}

int ndtb_p2d_match(ndtb_ctx *ctx, const ndtb_map *tgt, const float *pts, int64_t n, int mem, const double *T0,
                   const ndtb_params *p, ndtb_result *res) {
  DeviceGuard dev_guard(ctx);
  if (!ctx || !tgt || !T0 || !p || !res) return NDTB_ERR_ARG;
  std::unique_ptr<ndtb_map> src;
  if (int rc = points_source(ctx, pts, n, mem, src)) return rc;
  const ndtb_map *sp = src.get();
  return match_batch_impl(ctx, 1, &tgt, &sp, T0, nullptr, p, 0, NDTB_MEM_HOST, res, nullptr);
}

int ndtb_register_scans(ndtb_ctx *ctx, int64_t n_pairs, const float *const *tgt_pts, const int64_t *n_tgt,
                        const float *const *src_pts, const int64_t *n_src, const double *T0s, double cell,
                        const double *map_size, double range_limit, const ndtb_params *p, int with_covariance, int in_mem,
                        int out_mem, ndtb_result *res, double *cov36s) {
  DeviceGuard dev_guard(ctx);
  if (!ctx || n_pairs < 0 || !p || !(cell > 0)) return NDTB_ERR_ARG;
  if (n_pairs == 0) return NDTB_OK;
  if (!tgt_pts || !n_tgt || !src_pts || !n_src || !T0s || !res) return NDTB_ERR_ARG;
  PhaseTimer pt0(ctx, "register_scans");
  static const double slow_ms = std::getenv("NDTB_SLOWLOG") ? std::atof(std::getenv("NDTB_SLOWLOG")) : 0.0;
  g_slow = SlowLog();
  const double t_call = now_ms();
  std::vector<std::unique_ptr<ndtb_map>> own((size_t)(2 * n_pairs));
  std::vector<ndtb_map *> maps((size_t)(2 * n_pairs));
  std::vector<const float *> pts((size_t)(2 * n_pairs));
  std::vector<int64_t> npts((size_t)(2 * n_pairs));
  for (int64_t i = 0; i < 2 * n_pairs; i++) {
    own[i].reset(new ndtb_map());
    ndtb_map *m = own[i].get();
    m->ctx = ctx;
    m->cell[0] = m->cell[1] = m->cell[2] = cell;
    std::memset(&m->g, 0, sizeof m->g);
    m->allow_box = true;  // never merged into, exported or scored: only the Gaussian view outlives the build
    if (map_size && map_size[0] > 0 && map_size[1] > 0 && map_size[2] > 0)
      m->map_sizex = (float)map_size[0], m->map_sizey = (float)map_size[1], m->map_sizez = (float)map_size[2];
    maps[i] = m;
    const int64_t e = i >> 1;
    pts[i] = (i & 1) ? src_pts[e] : tgt_pts[e];
    npts[i] = (i & 1) ? n_src[e] : n_tgt[e];
  }
  pt0.mark("create maps");
  PhaseTimer pt(ctx, "register_scans");
  if (int rc = ndtb_map_build_batch(ctx, 2 * n_pairs, maps.data(), pts.data(), npts.data(), range_limit, in_mem, 0xffffffffu, 255.f))
    return rc;
  pt.mark("build_batch");
  const double t_built = now_ms();
  const SlowLog at_build = g_slow;
  std::vector<const ndtb_map *> tg((size_t)n_pairs), sr((size_t)n_pairs);
  for (int64_t e = 0; e < n_pairs; e++) {
    tg[e] = maps[2 * e], sr[e] = maps[2 * e + 1];
    // A scan without a single usable point defines no grid.  One bad scan must not fail the whole batch (the reference
    // would run the other edges): it becomes an empty one-voxel map, its registration leaves the pose unchanged and
    // reports NDTB_ST_NO_CELLS.
    for (int q = 0; q < 2; q++) {
      ndtb_map *m = maps[2 * e + q];
      if (!m->grid_ready) m->set_grid(0.0, 0.0, 0.0, cell, cell, cell);
    }
  }
  const int rc = match_batch_impl(ctx, n_pairs, tg.data(), sr.data(), T0s, nullptr, p, with_covariance && cov36s, out_mem, res, cov36s);
  pt.mark("match+cov");
  // outputs in device memory stay there and the temporary maps are released stream-ordered: no host sync needed
  const double t_matched = now_ms();
  own.clear();
  pt.mark("release maps");
  if (slow_ms > 0 && now_ms() - t_call > slow_ms)
    std::fprintf(stderr,
                 "[ndtb slow] register_scans %.1f ms: build %.1f (alloc %.1f in %d, sync %.1f in %d, enqueue H2D %.1f) match %.1f "
                 "(alloc %.1f, sync %.1f in %d) release %.1f\n",
                 now_ms() - t_call, t_built - t_call, at_build.alloc_ms, at_build.allocs, at_build.sync_ms, at_build.syncs,
                 at_build.enq_ms, t_matched - t_built, g_slow.alloc_ms - at_build.alloc_ms, g_slow.sync_ms - at_build.sync_ms,
                 g_slow.syncs - at_build.syncs, now_ms() - t_matched);
  return rc;
}

int ndtb_map_write_jff(const ndtb_map *m, const char *path) {
  if (!m || !path) return NDTB_ERR_ARG;
  ndtb_grid g;
  if (int rc = ndtb_map_grid(m, &g)) return rc;
  const int64_t n = ndtb_map_num_cells(m, 0);
  std::vector<ndtb_cell> cells((size_t)std::max<int64_t>(n, 1));
  const int64_t k = n > 0 ? ndtb_map_export_cells(m, cells.data(), n, 0) : 0;
  if (k < 0) return (int)k;
  return ndtb_jff_write_cells(path, &g, cells.data(), k);
}

int ndtb_map_load_jff(ndtb_map *m, const char *path) {
  if (!m || !path) return NDTB_ERR_ARG;
  ndtb_grid g;
  int64_t n = 0;
  if (int rc = ndtb_jff_read_cells(path, &g, nullptr, 0, &n)) return rc;
  std::vector<ndtb_cell> cells((size_t)std::max<int64_t>(n, 1));
  if (int rc = ndtb_jff_read_cells(path, &g, cells.data(), n, &n)) return rc;
  return ndtb_map_from_cells(m, &g, cells.data(), n, 1);
}

// NDTFeatureGraph::updateLinkUsingNDTRegistration scores every refined link with overlapNDTOccupancyScore when !keepScore
// (ndt_feature_graph.cpp:335-342): one launch over all links.  T = n x 16 doubles (or the T field of n ndtb_result records:
// T_stride_bytes = sizeof(ndtb_result)) in `T_mem` memory; scores = n doubles in `out_mem` memory.
int ndtb_overlap_score_batch(ndtb_ctx *ctx, int64_t n, const ndtb_map *const *ref, const ndtb_map *const *mov, const void *T,
                             int64_t T_stride_bytes, int T_mem, int out_mem, double *scores) {
  DeviceGuard dev_guard(ctx);
  if (!ctx || n < 0 || (n > 0 && (!ref || !mov || !T || !scores)) || T_stride_bytes < 128 || T_stride_bytes % 8) return NDTB_ERR_ARG;
  if (n == 0) return NDTB_OK;
  std::vector<BuildJob> j((size_t)(2 * n));
  std::memset(j.data(), 0, sizeof(BuildJob) * j.size());
  for (int64_t l = 0; l < n; l++) {
    const ndtb_map *mm[2] = {ref[l], mov[l]};
    for (int i = 0; i < 2; i++) {
      if (!map_ok(mm[i])) return NDTB_ERR_GRID;
      if (mm[i]->ctx != ctx) return NDTB_ERR_ARG;
      BuildJob &b = j[(size_t)(2 * l + i)];
      b.g = mm[i]->g;
      mm[i]->fill_box(b);
      if (mm[i]->n_all > 0)
        b.amask = mm[i]->amask, b.abase = mm[i]->abase, b.tb_list = mm[i]->tb_list, b.counts = mm[i]->counts, b.cocc = mm[i]->cocc;
    }
  }
  cudaStream_t st = ctx->stream;
  Carver c;
  const size_t o_j = c.take(sizeof(BuildJob) * j.size()), o_T = c.take(T_mem == NDTB_MEM_DEVICE ? 0 : (size_t)T_stride_bytes * n);
  const size_t o_out = c.take(8 * (size_t)n);
  SlabP s;
  if (int rc = slab_alloc(ctx, c.off, s)) return rc;
  CU_TRY(ctx, h2d_small(ctx, s->p + o_j, j.data(), sizeof(BuildJob) * j.size(), st));
  const double *d_T = (const double *)T;
  if (T_mem != NDTB_MEM_DEVICE) {
    CU_TRY(ctx, h2d_small(ctx, s->p + o_T, T, (size_t)T_stride_bytes * n, st));
    d_T = (const double *)(s->p + o_T);
  }
  double *d_out = out_mem == NDTB_MEM_DEVICE ? scores : (double *)(s->p + o_out);
  ctx->launches += launch_overlap((const BuildJob *)(s->p + o_j), (int)n, d_T, (int)(T_stride_bytes / 8), d_out, st);
  CU_TRY(ctx, cudaGetLastError());
  if (out_mem != NDTB_MEM_DEVICE) CU_TRY(ctx, cudaMemcpyAsync(scores, d_out, 8 * (size_t)n, cudaMemcpyDeviceToHost, st));
  // the job table and a host T are pageable stack / vector memory: they must have been consumed before returning
  CU_TRY(ctx, timed_sync(st));
  return NDTB_OK;
}

int ndtb_overlap_score(ndtb_ctx *ctx, const ndtb_map *ref, const ndtb_map *mov, const double *T, double *score) {
  if (!ctx || !ref || !mov || !T || !score) return NDTB_ERR_ARG;
  return ndtb_overlap_score_batch(ctx, 1, &ref, &mov, T, 128, NDTB_MEM_HOST, NDTB_MEM_HOST, score);
}

int ndtb_transform_point_cloud(ndtb_ctx *ctx, const double *T16, const float *in, int64_t n, int in_mem, float *out, int out_mem) {
  DeviceGuard dev_guard(ctx);
  if (!ctx || !T16 || n < 0 || n > 0x7fffffff || (n > 0 && (!in || !out))) return NDTB_ERR_ARG;
  if (n == 0) return NDTB_OK;
  cudaStream_t st = ctx->stream;
  SlabP ibuf, obuf, tbuf;
  const float4 *d_in = (const float4 *)in;
  if (in_mem != NDTB_MEM_DEVICE) {
    if (int rc = stage_points(ctx, in, n, in_mem, ibuf)) return rc;
    d_in = (const float4 *)ibuf->p;
  }
  float4 *d_out = (float4 *)out;
  if (out_mem != NDTB_MEM_DEVICE) {
    if (int rc = slab_alloc(ctx, 16 * (size_t)n, obuf)) return rc;
    d_out = (float4 *)obuf->p;
  }
  float T12[12];  // Eigen::Affine3d::cast<float>(): every coefficient rounded to float
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 3; c++) T12[r * 3 + c] = (float)T16[c * 4 + r];
    T12[9 + r] = (float)T16[12 + r];
  }
  if (int rc = slab_alloc(ctx, sizeof T12, tbuf)) return rc;
  CU_TRY(ctx, h2d_small(ctx, tbuf->p, T12, sizeof T12, st));
  ctx->launches += launch_transform_points(d_in, d_out, (int)n, (const float *)tbuf->p, st);
  if (out_mem != NDTB_MEM_DEVICE) {
    CU_TRY(ctx, cudaMemcpyAsync(out, d_out, 16 * (size_t)n, cudaMemcpyDeviceToHost, st));
    CU_TRY(ctx, timed_sync(st));
  } else {
    CU_TRY(ctx, timed_sync(st));  // T12 is a stack array: the copy must have been consumed before returning
  }
  return NDTB_OK;
}

// device scratch for the front end (fuser.cu): a stream-ordered buffer owned by the context's pool
int ndtb_internal_alloc(ndtb_ctx *ctx, size_t bytes, void **out) {
  DeviceGuard dev_guard(ctx);
  if (!ctx || !out) return NDTB_ERR_ARG;
  CU_TRY(ctx, cudaMallocFromPoolAsync(out, bytes ? bytes : 256, ctx->pool, ctx->stream));
  return NDTB_OK;
}
void ndtb_internal_free(ndtb_ctx *ctx, void *p) {
  DeviceGuard dev_guard(ctx);
  if (ctx && p) cudaFreeAsync(p, ctx->stream);
}
int ndtb_internal_upload(ndtb_ctx *ctx, void *dst, const void *src, size_t bytes, int src_mem) {
  DeviceGuard dev_guard(ctx);
  if (!ctx) return NDTB_ERR_ARG;
  CU_TRY(ctx, src_mem == NDTB_MEM_DEVICE ? cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream)
                                         : h2d_small(ctx, dst, src, bytes, ctx->stream));
  if (src_mem != NDTB_MEM_DEVICE) CU_TRY(ctx, timed_sync(ctx->stream));  // the caller may reuse its buffer
  return NDTB_OK;
}

// ---- NCCL result gather (loaded lazily: the library itself has no link-time dependency on NCCL)
namespace {
struct NcclId128 {  // ncclUniqueId (passed by value)
  char b[128];
};
struct NcclApi {
  void *h = nullptr;
  int (*GetUniqueId)(void *) = nullptr;
  int (*CommInitRank)(void **, int, NcclId128, int) = nullptr;
  int (*CommDestroy)(void *) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  bool ok = false;
};
NcclApi &nccl() {
  static NcclApi a;
  static bool tried = false;
  if (tried) return a;
  tried = true;
  for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
    a.h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
    if (a.h) break;
  }
  if (!a.h) return a;
  a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(a.h, "ncclGetUniqueId");
  a.CommInitRank = (decltype(a.CommInitRank))dlsym(a.h, "ncclCommInitRank");
  a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.h, "ncclCommDestroy");
  a.AllGather = (decltype(a.AllGather))dlsym(a.h, "ncclAllGather");
  a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.h, "ncclGetErrorString");
  a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllGather;
  return a;
}
}  // namespace

struct ndtb_comm {
  ndtb_ctx *ctx;
  void *comm = nullptr;
  int rank = 0, world = 1;
};

int ndtb_comm_unique_id(char id128[128]) {
  if (!id128) return NDTB_ERR_ARG;
  NcclApi &a = nccl();
  if (!a.ok) return NDTB_ERR_CUDA;
  return a.GetUniqueId(id128) == 0 ? NDTB_OK : NDTB_ERR_CUDA;
}

int ndtb_comm_create(ndtb_ctx *ctx, const char id128[128], int rank, int world, ndtb_comm **out) {
  DeviceGuard dev_guard(ctx);
  if (!ctx || !id128 || !out || world < 1 || rank < 0 || rank >= world) return NDTB_ERR_ARG;
  NcclApi &a = nccl();
  if (!a.ok) {
    ctx->last_error = "libnccl.so.2 not found";
    return NDTB_ERR_CUDA;
  }
  NcclId128 id;
  std::memcpy(id.b, id128, 128);
  ndtb_comm *c = new ndtb_comm();
  c->ctx = ctx, c->rank = rank, c->world = world;
  // The gather moves 192-byte records: one CTA is plenty.  NCCL's default (up to 32 CTAs per collective) matters here because
  // an all-gather kernel spins on its SMs until the slowest rank arrives, and the registration kernels want every SM: with 8
  // ranks and 3 lanes the default cost 7 % of the step (B200 x 8: 59.1 -> 55.3 ms).  A value set by the caller is kept.
  setenv("NCCL_MAX_CTAS", "1", 0);
  const int rc = a.CommInitRank(&c->comm, world, id, rank);
  if (rc != 0) {
    ctx->last_error = std::string("ncclCommInitRank: ") + (a.GetErrorString ? a.GetErrorString(rc) : "error");
    delete c;
    return NDTB_ERR_CUDA;
  }
  *out = c;
  return NDTB_OK;
}

void ndtb_comm_destroy(ndtb_comm *c) {
  if (!c) return;
  DeviceGuard dev_guard(c->ctx);
  cudaStreamSynchronize(c->ctx->stream);
  if (c->comm) nccl().CommDestroy(c->comm);
  delete c;
}

int ndtb_gather_results(ndtb_comm *c, const ndtb_result *local_dev, int64_t n_local, ndtb_result *all_dev) {
  if (!c || n_local < 0 || (n_local > 0 && (!local_dev || !all_dev))) return NDTB_ERR_ARG;
  if (n_local == 0) return NDTB_OK;
  DeviceGuard dev_guard(c->ctx);
  const int rc = nccl().AllGather(local_dev, all_dev, sizeof(ndtb_result) * (size_t)n_local, /*ncclChar*/ 0, c->comm, c->ctx->stream);
  if (rc != 0) {
    c->ctx->last_error = std::string("ncclAllGather: ") + (nccl().GetErrorString ? nccl().GetErrorString(rc) : "error");
    return NDTB_ERR_CUDA;
  }
  return NDTB_OK;
}

}  // extern "C"
