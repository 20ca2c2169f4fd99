// map_build.cu — kernel (i): voxel binning of a scan and per-voxel mean / covariance (NDT cell) computation,
// batched over many maps per launch.  COMPILED WITH -fmad=false: every floating-point operation below is a
// separately rounded IEEE operation, in the same order as the CPU restatement, so voxel indices, N, means,
// covariances and eigen-clamped covariances are BIT-IDENTICAL to it (exact-parity claim of DESIGN.md).
//
// Reference path replaced:
//   a12 NDTMap::loadPointCloud / addPointCloud end-point binning + LazyGrid::getIndexForPoint/addPoint [upstream]
//       call sites ndt_feature/src/ndt_feature_src/ndt_feature_fuser_hmt.cpp:92,201-225,485
//   a13 NDTMap::computeNDTCells -> NDTCell::computeGaussian(SAMPLE_VARIANCE) + rescaleCovariance [upstream]
//       call sites ndt_feature_fuser_hmt.cpp:94,227,486
//
// Pipeline (all order-independent pieces are parallel; the order-dependent sums are done in point-index order):
//   centroid/extent (guess-size grids only) -> mark touched voxels in a 1-bit-per-voxel block mask (dense over the
//   storage box) [-> trace the rays of addPointCloud: mark + list the cells every ray meets] ->
//   popcount scan (cell numbering = (block, bit) order, touched-block list) -> cell id per point -> stable LSD radix
//   sort of the point ids by cell id (ids stay ascending inside a cell = insertion order of NDTCell::points_) ->
//   per-cell segments [-> per cell: free-space evidence of the rays in tracing order] -> per cell: sequential mean,
//   sequential scatter matrix, merge with the previous (N, mean, cov), occupancy, 3x3 Jacobi eigen clamp ->
//   Gaussian view (compact cells + block hash table).
#include "map_build.cuh"

namespace ndtb {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ bool pt_skip(const float4 p, double range_limit, const double *o) {
  if (isnan(p.x) || isnan(p.y) || isnan(p.z)) return true;
  if (range_limit > 0) {
    const double d0 = (double)p.x - o[0], d1 = (double)p.y - o[1], d2 = (double)p.z - o[2];  // exact for o = 0
    const double d = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
    if (d > range_limit) return true;
  }
  return false;
}

// ---- guess-size grids: centroid in point order (one warp per map), then extents -------------------------
// out[map*8 + {0,1,2}] = centroid sum / count, [3] = count, [4] = maxDist bits, [5] = max dz key, [6] = min dz key
//
// The reference adds the coordinates one by one in point order (every addition rounds), so the sum is order dependent.
// It is reproduced bit for bit without walking the points one at a time: a chunk of 256 points is added in one go
// whenever NO addition of the sequential walk through that chunk can round — then the walk's result is the exact sum,
// which a tree computes just as well.  Sufficient condition per axis: with S the running sum, every partial sum is
// bounded by M = |S| + sum|v| < 2^K and every operand is a multiple of 2^q (q = lowest set bit of S, float ulp of the
// smallest non-zero |v|), so all partial sums are multiples of 2^q below 2^K: exactly representable when K - q <= 53.
// Chunks that fail the test (a coordinate within millimetres of zero next to a large running sum) are walked in order.
constexpr int CEN_U = 8;                 // points per lane per chunk
constexpr int CEN_CHUNK = 32 * CEN_U;    // 256 points
constexpr int CEN_REC = 10;              // doubles per chunk record: 3 x (L, AB, mn) + count

__device__ __forceinline__ int lowbit_exp(double s) {  // exponent of the lowest set bit of a finite double; 4096 for 0
  const unsigned long long b = (unsigned long long)__double_as_longlong(s);
  const int e = (int)((b >> 52) & 0x7ffull);
  unsigned long long m = b & 0xfffffffffffffull;
  if (e == 0 && m == 0ull) return 4096;
  if (e != 0) m |= 1ull << 52;
  return (e == 0 ? -1074 : e - 1075) + (__ffsll((long long)m) - 1);
}
__device__ __forceinline__ int dexp(double x) {  // floor(log2 |x|) of a normal double (the operands here are float-derived)
  return (int)(((unsigned long long)__double_as_longlong(x) >> 52) & 0x7ffull) - 1023;
}

// (1) fully parallel: one warp per 256-point chunk.  Per axis: L = tree sum of the chunk (exact whenever the test in
// (2) passes), AB = sum |v| rounded up, mn = smallest non-zero |v|; plus the number of usable points.
__global__ void __launch_bounds__(128) k_centroid_chunks(const BuildJob *__restrict__ jobs, const int *__restrict__ which,
                                                         const long long *__restrict__ rec_off, double *__restrict__ recs) {
  const BuildJob &j = jobs[which[blockIdx.y]];
  const int lane = threadIdx.x & 31;
  const int nchunks = (j.npts + CEN_CHUNK - 1) / CEN_CHUNK;
  for (int ch = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); ch < nchunks; ch += gridDim.x * (blockDim.x >> 5)) {
    const int base = ch * CEN_CHUNK;
    double L[3] = {0, 0, 0}, AB[3] = {0, 0, 0}, mn[3] = {1.7976931348623157e308, 1.7976931348623157e308, 1.7976931348623157e308};
    int used = 0;
    float4 p[CEN_U];
#pragma unroll
    for (int u = 0; u < CEN_U; u++) {
      const int i = base + u * 32 + lane;
      p[u] = i < j.npts ? j.pts[i] : make_float4(nanf(""), 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < CEN_U; u++) {
      const bool ok = !pt_skip(p[u], j.range_limit, j.range_origin);
      used += ok;
      const double v[3] = {ok ? (double)p[u].x : 0.0, ok ? (double)p[u].y : 0.0, ok ? (double)p[u].z : 0.0};
#pragma unroll
      for (int a = 0; a < 3; a++) {
        const double ax = fabs(v[a]);
        L[a] += v[a];
        AB[a] = __dadd_ru(AB[a], ax);  // rounded up: a true upper bound
        mn[a] = (ax != 0.0 && ax < mn[a]) ? ax : mn[a];
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
      for (int a = 0; a < 3; a++) {
        L[a] += __shfl_xor_sync(FULL, L[a], off);
        AB[a] = __dadd_ru(AB[a], __shfl_xor_sync(FULL, AB[a], off));
        const double o = __shfl_xor_sync(FULL, mn[a], off);
        mn[a] = o < mn[a] ? o : mn[a];
      }
    }
    used = __reduce_add_sync(FULL, used);
    if (lane == 0) {
      double *r = recs + (rec_off[blockIdx.y] + ch) * CEN_REC;
#pragma unroll
      for (int a = 0; a < 3; a++) r[a * 3] = L[a], r[a * 3 + 1] = AB[a], r[a * 3 + 2] = mn[a];
      r[9] = (double)used;
    }
  }
}

// (2) one warp per map walks its chunk records in order (lanes prefetch 32 records at a time, the recurrence on S is a
// few dozen instructions per chunk); a chunk that fails the exactness test is walked point by point.
__global__ void __launch_bounds__(32) k_centroid(const BuildJob *__restrict__ jobs, const int *__restrict__ which,
                                                 const long long *__restrict__ rec_off, const double *__restrict__ recs,
                                                 double *__restrict__ out) {
  const BuildJob &j = jobs[which[blockIdx.x]];
  const int lane = threadIdx.x;
  const int nchunks = (j.npts + CEN_CHUNK - 1) / CEN_CHUNK;
  const double *rbase = recs + rec_off[blockIdx.x] * CEN_REC;
  double S[3] = {0.0, 0.0, 0.0};
  long long cnt = 0;
  for (int c0 = 0; c0 < nchunks; c0 += 32) {
    double r[CEN_REC];
#pragma unroll
    for (int q = 0; q < CEN_REC; q++) r[q] = (c0 + lane < nchunks) ? rbase[(size_t)(c0 + lane) * CEN_REC + q] : 0.0;
    const int m = min(32, nchunks - c0);
    for (int l = 0; l < m; l++) {
      cnt += (long long)__shfl_sync(FULL, r[9], l);
#pragma unroll
      for (int a = 0; a < 3; a++) {
        const double L = __shfl_sync(FULL, r[a * 3], l), AB = __shfl_sync(FULL, r[a * 3 + 1], l), mn = __shfl_sync(FULL, r[a * 3 + 2], l);
        const double M = __dadd_ru(fabs(S[a]), AB);  // >= |any partial sum| of the walk through this chunk
        bool exact = M < 1e300;
        if (exact && AB != 0.0) {
          int q = dexp(mn) - 23;  // float ulp of the smallest non-zero addend
          const int qs = lowbit_exp(S[a]);
          q = qs < q ? qs : q;
          exact = dexp(M) - q <= 52;  // M < 2^(q+53): every multiple of 2^q up to M is a double
        }
        if (exact) {
          S[a] += L;  // = the sequential result (no rounding anywhere)
        } else {      // walk the chunk in point order
          const int base = (c0 + l) * CEN_CHUNK;
          double s = S[a];
          for (int u = 0; u < CEN_U; u++) {
            const int i = base + u * 32 + lane;
            float4 p = make_float4(nanf(""), 0.f, 0.f, 0.f);
            if (i < j.npts) p = j.pts[i];
            const bool ok = !pt_skip(p, j.range_limit, j.range_origin);
            const double x = a == 0 ? (double)p.x : (a == 1 ? (double)p.y : (double)p.z);
            const unsigned um = __ballot_sync(FULL, ok);
            for (int ll = 0; ll < 32; ll++) {
              const double xx = __shfl_sync(FULL, x, ll);
              if ((um >> ll) & 1u) s += xx;
            }
          }
          S[a] = s;
        }
      }
    }
  }
  if (lane == 0) {
    double *o = out + (size_t)blockIdx.x * 8;
    o[3] = (double)cnt;
    if (cnt > 0) o[0] = S[0] / (double)cnt, o[1] = S[1] / (double)cnt, o[2] = S[2] / (double)cnt;
    unsigned long long *k = reinterpret_cast<unsigned long long *>(o);
    k[4] = 0ull;   // maxDist = +0.0
    k[5] = 0ull;   // ordered key of the smallest value
    k[6] = ~0ull;  // ordered key of the largest value
  }
}
__device__ __forceinline__ unsigned long long ord_key(double v) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__global__ void k_extent(const BuildJob *__restrict__ jobs, const int *__restrict__ which, double *__restrict__ out) {
  const BuildJob &j = jobs[which[blockIdx.y]];
  double *o = out + (size_t)blockIdx.y * 8;
  const double cx = o[0], cy = o[1], cz = o[2];
  // maxDist = max_i sqrt(s_i) = sqrt(max_i s_i): a correctly rounded square root is monotone, so one sqrt per warp gives
  // the bits the reference's per-point sqrt gives
  double ms = 0.0;
  unsigned long long kmax = 0ull, kmin = ~0ull;
  const int stride = gridDim.x * blockDim.x;
  for (int i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < j.npts; i0 += 4 * stride) {
    float4 p[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int i = i0 + u * stride;
      p[u] = i < j.npts ? j.pts[i] : make_float4(nanf(""), 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      if (pt_skip(p[u], j.range_limit, j.range_origin)) continue;
      const double d0 = cx - (double)p[u].x, d1 = cy - (double)p[u].y, d2 = cz - (double)p[u].z;
      const double sq = d0 * d0 + d1 * d1 + d2 * d2;
      ms = sq > ms ? sq : ms;
      const unsigned long long kk = ord_key(d2);
      kmax = kk > kmax ? kk : kmax;
      kmin = kk < kmin ? kk : kmin;
    }
  }
  for (int off = 16; off > 0; off >>= 1) {
    const double m2 = __shfl_xor_sync(FULL, ms, off);
    ms = m2 > ms ? m2 : ms;
    const unsigned long long a = __shfl_xor_sync(FULL, kmax, off), b = __shfl_xor_sync(FULL, kmin, off);
    kmax = a > kmax ? a : kmax;
    kmin = b < kmin ? b : kmin;
  }
  if ((threadIdx.x & 31) == 0) {
    unsigned long long *k = reinterpret_cast<unsigned long long *>(o);
    atomicMax(k + 4, (unsigned long long)__double_as_longlong(sqrt(ms)));  // >= 0: the bit pattern is monotone
    atomicMax(k + 5, kmax);
    atomicMin(k + 6, kmin);
  }
}

__device__ __forceinline__ int cta_excl_scan(int v, int *tot, int *wsum);

// storage block id of a voxel inside the job's storage box, -1 outside
__device__ __forceinline__ int sblock_id(const BuildJob &j, int ix, int iy, int iz) {
  const int bx = (ix >> 2) - j.boff[0], by = (iy >> 2) - j.boff[1], bz = (iz >> 2) - j.boff[2];
  if (bx < 0 || by < 0 || bz < 0 || bx >= j.nbs[0] || by >= j.nbs[1] || bz >= j.nbs[2]) return -1;
  return (bx * j.nbs[1] + by) * j.nbs[2] + bz;
}
// block coordinates in the full grid of a storage block id
__device__ __forceinline__ void sblock_coords(const BuildJob &j, int b, int &bx, int &by, int &bz) {
  bz = b % j.nbs[2] + j.boff[2], by = (b / j.nbs[2]) % j.nbs[1] + j.boff[1], bx = b / (j.nbs[2] * j.nbs[1]) + j.boff[0];
}

// ---- free-space ray trace (NDTMap::addPointCloud + LazyGrid::traceLine [upstream]) ---------------------------
// call sites ndt_feature_fuser_hmt.cpp:92 (initialize) and :485 (update).  A pending point that belongs to a trace
// segment is a ray from the segment's origin: the ray is dropped (end point included) when it is longer than 200 m or
// ends above maxz; otherwise N = int(l / min cell size) and the N-2 samples origin + (i+1) diff/N, rounded to FLOAT
// points, are looked up with the LazyGrid index rule; a sample in the voxel of the previous sample is skipped.
__device__ __forceinline__ int seg_of(const BuildJob &j, int i) {
  for (int s = 0; s < j.n_seg; s++)
    if (i >= j.seg[s].begin && i < j.seg[s].end) return s;
  return -1;
}
// false: the ray (and its end point) is ignored
__device__ __forceinline__ bool ray_setup(const BuildJob &j, const TraceSeg &sg, const float4 p, double *d, int &N, double &l) {
  if (isnan(p.x) || isnan(p.y) || isnan(p.z)) return false;
  const double f0 = (double)p.x - sg.origin[0], f1 = (double)p.y - sg.origin[1], f2 = (double)p.z - sg.origin[2];
  l = sqrt(f0 * f0 + f1 * f1 + f2 * f2);
  if (l > 200.) return false;
  if ((double)p.z > sg.maxz) return false;
  const double m1 = j.g.cell[0] < j.g.cell[1] ? j.g.cell[0] : j.g.cell[1];
  const double m2 = j.g.cell[2] < j.g.cell[1] ? j.g.cell[2] : j.g.cell[1];
  const double res = m1 < m2 ? m1 : m2;
  if (res < 0.01) return false;
  N = __double2int_rz(l / res);
  const double fN = (double)(float)N;
  d[0] = f0 / fN, d[1] = f1 / fN, d[2] = f2 / fN;
  return true;
}
// calls visit(key) for every cell the ray meets, in order
template <class F>
__device__ __forceinline__ void ray_walk(const BuildJob &j, const TraceSeg &sg, const double *d, int N, F visit) {
  int xo = 0, yo = 0, zo = 0;
  for (int s = 0; s < N - 2; s++) {
    const double f = (double)(float)(s + 1);
    const float qx = (float)(sg.origin[0] + f * d[0]), qy = (float)(sg.origin[1] + f * d[1]), qz = (float)(sg.origin[2] + f * d[2]);
    int x, y, z;
    if (!voxel_index(j.g, (double)qx, (double)qy, (double)qz, x, y, z)) continue;
    if (x == xo && y == yo && z == zo) continue;
    xo = x, yo = y, zo = z;
    if (!in_grid(j.g, x, y, z)) continue;
    const int sb = sblock_id(j, x, y, z);
    if (sb >= 0) visit(sb * 64 + block_bit(x, y, z));
  }
}

// thread per ray: mark the voxels it meets (they become cells) and count them
__global__ void k_trace_count(const BuildJob *__restrict__ jobs) {
  const BuildJob &j = jobs[blockIdx.y];
  if (j.n_seg == 0) return;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < j.npts; i += gridDim.x * blockDim.x) {
    int cnt = 0;
    const int s = seg_of(j, i);
    if (s >= 0) {
      double d[3], l;
      int N;
      if (ray_setup(j, j.seg[s], j.pts[i], d, N, l))
        ray_walk(j, j.seg[s], d, N, [&](int key) {
          const unsigned long long bit = 1ull << (key & 63);
          if (!(__ldcg(j.amask + (key >> 6)) & bit)) atomicOr(j.amask + (key >> 6), bit);
          cnt++;
        });
    }
    j.vis_cnt[i] = cnt;
  }
}

// one CTA per map: exclusive scan of the per-ray visit counts
__global__ void __launch_bounds__(1024) k_rayscan(const BuildJob *__restrict__ jobs) {
  const BuildJob &j = jobs[blockIdx.x];
  if (j.n_seg == 0) {
    if (threadIdx.x == 0) j.counts[7] = 0;
    return;
  }
  __shared__ int wsum[32];
  int run = 0;
  for (int base = 0; base < j.npts; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < j.npts ? j.vis_cnt[i] : 0;
    int t;
    const int e = cta_excl_scan(v, &t, wsum);
    if (i < j.npts) j.vis_off[i] = run + e;
    run += t;
  }
  if (threadIdx.x == 0) j.counts[7] = run;
}

// thread per ray: the same walk again, writing (voxel key, ray) of every visit at its ray-major position
__global__ void k_trace_fill(const BuildJob *__restrict__ jobs) {
  const BuildJob &j = jobs[blockIdx.y];
  if (j.n_seg == 0) return;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < j.npts; i += gridDim.x * blockDim.x) {
    if (j.vis_cnt[i] == 0) continue;
    const int s = seg_of(j, i);
    double d[3], l;
    int N;
    if (!ray_setup(j, j.seg[s], j.pts[i], d, N, l)) continue;
    int at = j.vis_off[i];
    ray_walk(j, j.seg[s], d, N, [&](int key) {
      j.vis_key[at] = key;
      j.vis_ray[at] = i;
      at++;
    });
  }
}

// ---- binning -----------------------------------------------------------------------------------------
// Four points per thread and iteration, all loads of a stage issued before the first use: the kernel is bound by the
// latency of the dependent mask look-up, not by bytes.
// FAST: every map of the launch has power-of-two cell sizes, no range limit and no trace segment (the host checks) — the
// general kernel carries the IEEE division and the square root of the range test as predicated code even where a job does
// not need them (60 % of its issued instructions on the C2 workload, ncu source counters).
template <bool FAST>
__global__ void k_mark(const BuildJob *__restrict__ jobs) {
  const BuildJob &j = jobs[blockIdx.y];
  const int stride = gridDim.x * blockDim.x;
  double inv[3] = {0, 0, 0}, half[3] = {0, 0, 0};
  if (FAST) {
#pragma unroll
    for (int a = 0; a < 3; a++) inv[a] = __drcp_rn(j.g.cell[a]), half[a] = (double)j.g.size[a] * 0.5;  // exact: power of two
  }
  for (int i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < j.npts; i0 += 4 * stride) {
    float4 p[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int i = i0 + u * stride;
      p[u] = i < j.npts ? j.pts[i] : make_float4(nanf(""), 0.f, 0.f, 0.f);
    }
    int key[4], b[4], bit[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      int ix, iy, iz;
      key[u] = -1, b[u] = 0, bit[u] = 0;
      bool ok;
      if (FAST) {
        // same operations as voxel_axis with x / cell replaced by the exact x * (1 / cell); NaN fails the range test
        const double v0 = __dadd_rn(floor(__dadd_rn(__dmul_rn(__dsub_rn((double)p[u].x, j.g.center[0]), inv[0]), 0.5)), half[0]);
        const double v1 = __dadd_rn(floor(__dadd_rn(__dmul_rn(__dsub_rn((double)p[u].y, j.g.center[1]), inv[1]), 0.5)), half[1]);
        const double v2 = __dadd_rn(floor(__dadd_rn(__dmul_rn(__dsub_rn((double)p[u].z, j.g.center[2]), inv[2]), 0.5)), half[2]);
        ok = v0 > -2147483000.0 && v0 < 2147483000.0 && v1 > -2147483000.0 && v1 < 2147483000.0 && v2 > -2147483000.0 &&
             v2 < 2147483000.0;
        ix = __double2int_rz(v0), iy = __double2int_rz(v1), iz = __double2int_rz(v2);
      } else {
        bool live = !pt_skip(p[u], j.range_limit, j.range_origin);
        if (live && j.n_seg) {  // the end point of a ray that addPointCloud ignores is not binned either
          const int sgi = seg_of(j, i0 + u * stride);
          if (sgi >= 0) {
            double dd[3], l;
            int N;
            live = ray_setup(j, j.seg[sgi], p[u], dd, N, l);
          }
        }
        ok = live && voxel_index(j.g, (double)p[u].x, (double)p[u].y, (double)p[u].z, ix, iy, iz);
      }
      if (ok && in_grid(j.g, ix, iy, iz)) {
        b[u] = sblock_id(j, ix, iy, iz), bit[u] = block_bit(ix, iy, iz);
        key[u] = b[u] >= 0 ? b[u] * 64 + bit[u] : -1;  // (outside the storage box: cannot happen, the box bounds the points)
        b[u] = b[u] >= 0 ? b[u] : 0;
      }
    }
    // ~90 % of the points fall into a voxel that is already marked: look before taking the (contended) atomic.  A stale
    // read only costs a redundant atomicOr.
    unsigned long long m[4];
#pragma unroll
    for (int u = 0; u < 4; u++) m[u] = key[u] >= 0 ? __ldcg(j.amask + b[u]) : ~0ull;
#pragma unroll
    for (int u = 0; u < 4; u++)
      if (key[u] >= 0 && !((m[u] >> bit[u]) & 1ull)) atomicOr(j.amask + b[u], 1ull << bit[u]);
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int i = i0 + u * stride;
      if (i < j.npts) j.pt_cell[i] = key[u];
    }
  }
}

// CTA-wide exclusive scan of one int per thread (blockDim.x == 1024); returns the exclusive prefix, total in *tot
__device__ __forceinline__ int cta_excl_scan(int v, int *tot, int *wsum /*[32] shared*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
  for (int off = 1; off < 32; off <<= 1) {
    const int t = __shfl_up_sync(FULL, inc, off);
    if (lane >= off) inc += t;
  }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = wsum[lane];
    for (int off = 1; off < 32; off <<= 1) {
      const int t = __shfl_up_sync(FULL, w, off);
      if (lane >= off) w += t;
    }
    wsum[lane] = w;  // inclusive
  }
  __syncthreads();
  const int before = warp ? wsum[warp - 1] : 0;
  *tot = wsum[31];
  __syncthreads();
  return before + inc - v;
}

// one CTA per map: abase = exclusive popcount scan of amask, compact list of touched blocks
__global__ void __launch_bounds__(1024) k_blockscan(const BuildJob *__restrict__ jobs) {
  const BuildJob &j = jobs[blockIdx.x];
  __shared__ int wsum[32];
  int run_c = 0, run_b = 0;
  for (int base = 0; base < j.nblk; base += 1024) {
    const int b = base + threadIdx.x;
    const unsigned long long m = b < j.nblk ? j.amask[b] : 0ull;
    const int c = __popcll(m), nz = m != 0ull;
    int tc, tb;
    const int ec = cta_excl_scan(c, &tc, wsum);
    const int eb = cta_excl_scan(nz, &tb, wsum);
    if (b < j.nblk) j.abase[b] = run_c + ec;
    if (nz) j.tb_list[run_b + eb] = b;
    run_c += tc, run_b += tb;
  }
  if (threadIdx.x == 0) j.counts[0] = run_c, j.counts[1] = run_b;
}

__global__ void k_count(const BuildJob *__restrict__ jobs) {
  const BuildJob &j = jobs[blockIdx.y];
  const int stride = gridDim.x * blockDim.x;
  for (int i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < j.npts; i0 += 4 * stride) {
    int key[4], base[4];
    unsigned long long m[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int i = i0 + u * stride;
      key[u] = i < j.npts ? j.pt_cell[i] : -1;
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      base[u] = 0, m[u] = 0ull;
      if (key[u] >= 0) base[u] = j.abase[key[u] >> 6], m[u] = j.amask[key[u] >> 6];
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      if (key[u] < 0) continue;
      const int i = i0 + u * stride;
      j.pt_cell[i] = base[u] + __popcll(m[u] & ((1ull << (key[u] & 63)) - 1ull));  // cell id: the sort key
    }
  }
}

// ---- stable LSD radix sort of the point ids by cell id (8-bit digits), batched over the maps of a build -------------
// Replaces "atomic rank + scatter + per-cell sort": a stable sort keeps the ids of a cell ascending (= the insertion order
// of NDTCell::points_ the per-cell sums are taken in) for free, without one atomic and without per-cell work.
// Dropped points (key -1) sort to the end as key n_all.  Ping-pong: X = (pt_cell, seg_idx), Y = (key2, seg2); pass p reads
// X when p is even; the ids of the first pass are implicit (iota).
constexpr int RS_TILE = 2048, RS_THREADS = 256, RS_WARPS = RS_THREADS / 32, RS_PER_WARP = RS_TILE / RS_WARPS;
__global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const BuildJob *__restrict__ jobs, int shift, int parity) {
  const BuildJob &j = jobs[blockIdx.y];
  const int n = j.npts, tile = blockIdx.x;
  if (tile * RS_TILE >= n || j.pt_cell == nullptr) return;
  const int T = (n + RS_TILE - 1) / RS_TILE;
  const int *keys = parity ? j.key2 : j.pt_cell;
  __shared__ int hist[256];
  hist[threadIdx.x] = 0;
  __syncthreads();
  int kk[RS_TILE / RS_THREADS];
#pragma unroll
  for (int u = 0; u < RS_TILE / RS_THREADS; u++) {
    const int i = tile * RS_TILE + u * RS_THREADS + threadIdx.x;
    kk[u] = i < n ? keys[i] : 0;
  }
#pragma unroll
  for (int u = 0; u < RS_TILE / RS_THREADS; u++) {
    const int i = tile * RS_TILE + u * RS_THREADS + threadIdx.x;
    if (i < n) atomicAdd(&hist[((kk[u] < 0 ? j.n_all : kk[u]) >> shift) & 255], 1);
  }
  __syncthreads();
  j.rs_hist[threadIdx.x * T + tile] = hist[threadIdx.x];
}
// one CTA per map: exclusive scan of the 256 x tiles histogram in digit-major order
__global__ void __launch_bounds__(1024) k_rs_scan(const BuildJob *__restrict__ jobs) {
  const BuildJob &j = jobs[blockIdx.x];
  if (j.pt_cell == nullptr || j.npts <= 0) return;
  const int total = 256 * ((j.npts + RS_TILE - 1) / RS_TILE);
  __shared__ int wsum[32];
  int run = 0;
  for (int base = 0; base < total; base += 1024) {
    const int e = base + threadIdx.x;
    const int v = e < total ? j.rs_hist[e] : 0;
    int t;
    const int ex = cta_excl_scan(v, &t, wsum);
    if (e < total) j.rs_hist[e] = run + ex;
    run += t;
  }
}
__global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(const BuildJob *__restrict__ jobs, int shift, int parity, int first) {
  const BuildJob &j = jobs[blockIdx.y];
  const int n = j.npts, tile = blockIdx.x;
  if (tile * RS_TILE >= n || j.pt_cell == nullptr) return;
  const int T = (n + RS_TILE - 1) / RS_TILE;
  const int *keys = parity ? j.key2 : j.pt_cell, *vals = parity ? j.seg2 : j.seg_idx;
  int *okeys = parity ? j.pt_cell : j.key2, *ovals = parity ? j.seg_idx : j.seg2;
  __shared__ int wh[RS_WARPS][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  for (int q = threadIdx.x; q < RS_WARPS * 256; q += RS_THREADS) (&wh[0][0])[q] = 0;
  const int w0 = tile * RS_TILE + warp * RS_PER_WARP;
  constexpr int U = RS_PER_WARP / 32;
  int k[U], v[U], rank[U];
#pragma unroll
  for (int u = 0; u < U; u++) {  // all loads of the slice in flight together
    const int i = w0 + u * 32 + lane;
    k[u] = i < n ? keys[i] : 0;
    v[u] = (i < n && !first) ? vals[i] : i;
  }
  __syncthreads();
  // (a) one counting pass over the warp's contiguous slice: the running count of a digit when an element arrives, plus the
  // equal digits in lower lanes of its chunk, is the element's stable rank among the slice's elements with that digit
#pragma unroll
  for (int u = 0; u < U; u++) {
    const int i = w0 + u * 32 + lane;
    const bool act = i < n;
    const int d = ((k[u] < 0 ? j.n_all : k[u]) >> shift) & 255;
    const unsigned am = __ballot_sync(FULL, act);
    rank[u] = 0;
    if (act) {
      const unsigned peers = __match_any_sync(am, d);
      const int leader = __ffs(peers) - 1;
      int o = 0;
      if (lane == leader) o = wh[warp][d], wh[warp][d] = o + __popc(peers);
      o = __shfl_sync(peers, o, leader);
      rank[u] = o + __popc(peers & lt);
    }
    __syncwarp();
  }
  __syncthreads();
  // (b) first output position of (warp, digit): global offset of (digit, tile) + the slices before this warp
  {
    const int d = threadIdx.x;
    int base = j.rs_hist[d * T + tile];
    for (int w = 0; w < RS_WARPS; w++) {
      const int c = wh[w][d];
      wh[w][d] = base;
      base += c;
    }
  }
  __syncthreads();
  // (c) stable scatter
#pragma unroll
  for (int u = 0; u < U; u++) {
    const int i = w0 + u * 32 + lane;
    if (i < n) {
      const int d = ((k[u] < 0 ? j.n_all : k[u]) >> shift) & 255;
      const int pos = wh[warp][d] + rank[u];
      okeys[pos] = k[u];
      ovals[pos] = v[u];
    }
  }
}
// segment of every cell in the sorted order: seg_off[c] = first position, cnt[c] = number of points; counts[4] = points binned
__global__ void k_seg_bounds(const BuildJob *__restrict__ jobs, int final_parity) {
  const BuildJob &j = jobs[blockIdx.y];
  if (j.pt_cell == nullptr) return;
  const int n = j.npts;
  const int *keys = final_parity ? j.key2 : j.pt_cell;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int k = keys[i];
    const int kp = i > 0 ? keys[i - 1] : -2, kn = i + 1 < n ? keys[i + 1] : -2;
    if (k < 0) {
      if (kp >= 0 || i == 0) j.counts[4] = i;  // the dropped points follow the binned ones
      continue;
    }
    if (k != kp) j.seg_off[k] = i;
    if (k != kn) {
      j.cursor[k] = i + 1;  // segment end (cursor is free until k_cells builds its eigen list: consumed by k_seg_counts)
      if (i == n - 1) j.counts[4] = n;
    }
  }
}
__global__ void k_seg_counts(const BuildJob *__restrict__ jobs) {
  const BuildJob &j = jobs[blockIdx.y];
  if (j.pt_cell == nullptr) return;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < j.n_all; c += gridDim.x * blockDim.x) {
    const int e = j.cursor[c];
    j.cnt[c] = e > 0 ? e - j.seg_off[c] : 0;
    j.cursor[c] = 0;
  }
}
// voxel key (block * 64 + bit) of every cell, cells of a block in bit order
__global__ void k_cell_keys(const BuildJob *__restrict__ jobs) {
  const BuildJob &j = jobs[blockIdx.y];
  const int ntb = j.counts[1];
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < ntb; t += gridDim.x * blockDim.x) {
    const int b = j.tb_list[t];
    unsigned long long m = j.amask[b];
    int c = j.abase[b];
    for (; m; m &= m - 1ull, c++) j.cell_key[c] = b * 64 + (__ffsll((long long)m) - 1);
  }
}

// ---- per-cell Gaussians --------------------------------------------------------------------------------
// NDTCell::rescaleCovariance [upstream]: any eigenvalue <= 0 -> no Gaussian; clamp to >= max/1000 (fixture-pinned)
// max_sweeps < 64: *deferred is set (and cov left untouched) when the Jacobi iteration needs more sweeps than that
__device__ bool rescale_covariance(double *cov, int max_sweeps = 64, bool *deferred = nullptr) {
  double ev[3], V[9];
  // the deferred cells (second pass, full sweep cap) stop at the fixed point of the iteration: see eig_sym_n
  const bool done = deferred ? eig_sym_n<3, false>(cov, ev, V, max_sweeps) : eig_sym_n<3, true>(cov, ev, V, max_sweeps);
  if (deferred) *deferred = !done;
  if (!done) return false;
  if (ev[0] <= 0 || ev[1] <= 0 || ev[2] <= 0) return false;
  double maxe = ev[0] > ev[1] ? ev[0] : ev[1];
  maxe = maxe > ev[2] ? maxe : ev[2];
  bool recalc = false;
  for (int i = 0; i < 3; i++)
    if (maxe > ev[i] * 1000.0) ev[i] = maxe / 1000.0, recalc = true;
  if (recalc)
    for (int i = 0; i < 3; i++)
      for (int k = 0; k < 3; k++) {
        double s = 0;
        for (int q = 0; q < 3; q++) s += V[i * 3 + q] * ev[q] * V[k * 3 + q];
        cov[i * 3 + k] = s;
      }
  return true;
}

// Occupancy evidence of one traced ray for a cell that holds a Gaussian (NDTMap::addPointCloud [upstream], REFACTORED
// branch): maximum-likelihood point X of the cell's Gaussian on the line (NDTCell::computeMaximumLikelihoodAlongLine,
// origin rounded to float like pcl::PointXYZ po), its likelihood (getLikelihood of the float-rounded X), damped by the
// probability that X is the end point itself.  Returns false when the ray leaves the cell untouched.
__device__ __forceinline__ bool inv3_cofactor(const double *m, double *inv) {  // Eigen computeInverseAndDetWithCheck order
  const double c00 = m[4] * m[8] - m[5] * m[7];
  const double c10 = m[2] * m[7] - m[1] * m[8];
  const double c20 = m[1] * m[5] - m[2] * m[4];
  const double det = c00 * m[0] + c10 * m[3] + c20 * m[6];
  if (!(fabs(det) > 1e-12)) return false;
  const double id = 1.0 / det;
  inv[0] = c00 * id, inv[1] = c10 * id, inv[2] = c20 * id;
  inv[3] = (m[5] * m[6] - m[3] * m[8]) * id;
  inv[4] = (m[0] * m[8] - m[2] * m[6]) * id;
  inv[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  inv[6] = (m[3] * m[7] - m[4] * m[6]) * id;
  inv[7] = (m[1] * m[6] - m[0] * m[7]) * id;
  inv[8] = (m[0] * m[4] - m[1] * m[3]) * id;
  return true;
}
__device__ bool ray_gaussian_evidence(const double *mean, const double *icov, bool icov_ok, const TraceSeg &sg, const float4 p,
                                      float &logodd_out) {
  const double po[3] = {(double)(float)sg.origin[0], (double)(float)sg.origin[1], (double)(float)sg.origin[2]};
  const double pe[3] = {(double)p.x, (double)p.y, (double)p.z};
  const double f0 = pe[0] - sg.origin[0], f1 = pe[1] - sg.origin[1], f2 = pe[2] - sg.origin[2];
  const double l = sqrt(f0 * f0 + f1 * f1 + f2 * f2);
  double X[3] = {pe[0], pe[1], pe[2]};
  double lik = 1.0;
  if (icov_ok) {
    double L[3] = {pe[0] - po[0], pe[1] - po[1], pe[2] - po[2]};
    const double nrm = sqrt(L[0] * L[0] + L[1] * L[1] + L[2] * L[2]);
    L[0] /= nrm, L[1] /= nrm, L[2] /= nrm;
    const double A[3] = {icov[0] * L[0] + icov[1] * L[1] + icov[2] * L[2], icov[3] * L[0] + icov[4] * L[1] + icov[5] * L[2],
                         icov[6] * L[0] + icov[7] * L[1] + icov[8] * L[2]};
    const double B[3] = {pe[0] - mean[0], pe[1] - mean[1], pe[2] - mean[2]};
    const double sigma = A[0] * L[0] + A[1] * L[1] + A[2] * L[2];
    if (sigma != 0) {
      const double t = -(A[0] * B[0] + A[1] * B[1] + A[2] * B[2]) / sigma;
      X[0] = L[0] * t + pe[0], X[1] = L[1] * t + pe[1], X[2] = L[2] * t + pe[2];
      const double v[3] = {(double)(float)X[0] - mean[0], (double)(float)X[1] - mean[1], (double)(float)X[2] - mean[2]};
      const double iv[3] = {icov[0] * v[0] + icov[1] * v[1] + icov[2] * v[2], icov[3] * v[0] + icov[4] * v[1] + icov[5] * v[2],
                            icov[6] * v[0] + icov[7] * v[1] + icov[8] * v[2]};
      const double q = v[0] * iv[0] + v[1] * iv[1] + v[2] * iv[2];
      lik = isnan(q) ? -1.0 : exp(-q / 2);
    }
  }
  const double dist = sqrt((sg.origin[0] - X[0]) * (sg.origin[0] - X[0]) + (sg.origin[1] - X[1]) * (sg.origin[1] - X[1]) +
                           (sg.origin[2] - X[2]) * (sg.origin[2] - X[2]));
  if (dist > l) return false;
  const double l2 = sqrt((X[0] - pe[0]) * (X[0] - pe[0]) + (X[1] - pe[1]) * (X[1] - pe[1]) + (X[2] - pe[2]) * (X[2] - pe[2]));
  const double snoise = 0.5 * (dist / 30.0) + sg.sensor_noise;
  const double thr = exp(-0.5 * (l2 * l2) / (snoise * snoise));
  lik *= (1.0 - thr);
  if (lik < 0.3) return false;
  lik = 0.1 * lik + 0.5;
  logodd_out = (float)log((1.0 - lik) / lik);
  return true;
}

// (a') ray-traced builds only, one thread per cell: the free-space evidence of the rays that met the cell, applied to
// its stored occupancy in the order the rays were traced; a cell whose occupancy drops to <= 0 loses its Gaussian at
// once (later rays of the same scan see an empty cell).  Leaves (occupancy, hasGaussian) in the new record for k_cells.
__global__ void __launch_bounds__(128) k_cell_trace(const BuildJob *__restrict__ jobs) {
  const BuildJob &j = jobs[blockIdx.y];
  if (j.n_seg == 0) return;
  const float4 *__restrict__ pts = j.pts;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < j.n_all; c += gridDim.x * blockDim.x) {
    const int key = j.cell_key[c];
    const int b = key >> 6, bit = key & 63;
    double mean[3] = {0, 0, 0}, cov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    int has = 0;
    float occ = 0.f;
    if (j.o_amask) {
      const unsigned long long om = j.o_amask[b];
      if (om >> bit & 1ull) {
        const int oc = j.o_abase[b] + __popcll(om & ((1ull << bit) - 1ull));
        has = j.o_chas[oc], occ = j.o_cocc[oc];
        if (has) {
          for (int q = 0; q < 3; q++) mean[q] = j.o_cmean[(size_t)oc * 3 + q];
          for (int q = 0; q < 9; q++) cov[q] = j.o_ccov[(size_t)oc * 9 + q];
        }
      }
    }
    {
      const int nv = j.v_cnt[c];
      const int *__restrict__ vids = j.v_seg2 + j.v_seg_off[c];  // (host: v_seg2 = the visit sort's final buffer)
      double icov[9];
      bool icov_ok = false, icov_done = false;
      for (int q = 0; q < nv; q++) {
        const int ray = j.vis_ray[vids[q]];
        const TraceSeg &sg = j.seg[seg_of(j, ray)];
        float ev = (float)-0.2;
        bool apply = true;
        if (has) {
          if (!icov_done) icov_ok = inv3_cofactor(cov, icov), icov_done = true;
          apply = ray_gaussian_evidence(mean, icov, icov_ok, sg, pts[ray], ev);
        }
        if (apply) {
          occ += ev;
          occ = occ > sg.occ_limit ? sg.occ_limit : occ;
          occ = occ < -sg.occ_limit ? -sg.occ_limit : occ;
          if (occ <= 0.f) has = 0;
        }
      }
    }
    j.cocc[c] = occ, j.chas[c] = has;
  }
}

// (b) one THREAD per cell: sequential mean and scatter matrix over its points in id order (the operation order of
// NDTCell::computeGaussian), merge with the stored (N, mean, cov), occupancy.  Cells whose covariance has to go through
// rescaleCovariance are appended to the map's eigen list (the dead `cursor` array) and finished by k_eigen: the 3x3
// Jacobi iteration is a long serial fp64 chain that wants many resident warps, the gather loops here want registers.
#ifndef NDTB_KCELLS_MINBLOCKS
#define NDTB_KCELLS_MINBLOCKS 8  // resident CTAs per SM the register allocation of k_cells is sized for (64 registers; measured 1 -> 8: -0.7 ms per 1184 maps)
#endif
__global__ void __launch_bounds__(128, NDTB_KCELLS_MINBLOCKS) k_cells(const BuildJob *__restrict__ jobs) {
  const BuildJob &j = jobs[blockIdx.y];
  const float4 *__restrict__ pts = j.pts;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < j.n_all; c += gridDim.x * blockDim.x) {
    const int key = j.cell_key[c];
    const int b = key >> 6, bit = key & 63;
    double mean[3] = {0, 0, 0}, cov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    int N = 0, has = 0;
    float occ = 0.f;
    if (j.o_amask) {  // previous record of this voxel
      const unsigned long long om = j.o_amask[b];
      if (om >> bit & 1ull) {
        const int oc = j.o_abase[b] + __popcll(om & ((1ull << bit) - 1ull));
        for (int q = 0; q < 3; q++) mean[q] = j.o_cmean[(size_t)oc * 3 + q];
        for (int q = 0; q < 9; q++) cov[q] = j.o_ccov[(size_t)oc * 9 + q];
        N = j.o_cn[oc], has = j.o_chas[oc], occ = j.o_cocc[oc];
      }
    }
    if (j.n_seg) occ = j.cocc[c], has = j.chas[c];  // occupancy / Gaussian flag after the ray trace (k_cell_trace)
    const int n = j.cnt[c];
    bool eigen = false;
    bool emptied = false;
    if (n > 0) {  // occupancy: += n*log(0.6/0.4), clamped (NDTCell::updateOccupancy); a cell whose occupancy is still <= 0
                  // (free-space evidence outweighs the hits) loses its Gaussian and drops the points [upstream computeGaussian]
      float o2 = occ + (float)((double)n * j.log_occ);
      o2 = o2 > j.occ_limit ? j.occ_limit : o2;
      o2 = o2 < -j.occ_limit ? -j.occ_limit : o2;
      occ = o2;
      if (occ <= 0.f) has = 0, emptied = true;
    }
    if (n > 0 && !emptied) {
      const int *__restrict__ ids = j.sorted_ids + j.seg_off[c];
      if (has || n >= 3) {
        // the additions stay in id order; the gathers of four points are issued together
        double ms0 = 0, ms1 = 0, ms2 = 0;
        int q = 0;
        for (; q + 4 <= n; q += 4) {
          const float4 p0 = pts[ids[q]], p1 = pts[ids[q + 1]], p2 = pts[ids[q + 2]], p3 = pts[ids[q + 3]];
          ms0 += (double)p0.x, ms1 += (double)p0.y, ms2 += (double)p0.z;
          ms0 += (double)p1.x, ms1 += (double)p1.y, ms2 += (double)p1.z;
          ms0 += (double)p2.x, ms1 += (double)p2.y, ms2 += (double)p2.z;
          ms0 += (double)p3.x, ms1 += (double)p3.y, ms2 += (double)p3.z;
        }
        for (; q < n; q++) {
          const float4 p = pts[ids[q]];
          ms0 += (double)p.x, ms1 += (double)p.y, ms2 += (double)p.z;
        }
        const double ml0 = ms0 / (double)n, ml1 = ms1 / (double)n, ml2 = ms2 / (double)n;
        double c00 = 0, c01 = 0, c02 = 0, c11 = 0, c12 = 0, c22 = 0;
        q = 0;
        for (; q + 2 <= n; q += 2) {
          const float4 p0 = pts[ids[q]], p1 = pts[ids[q + 1]];
          const double d0 = (double)p0.x - ml0, d1 = (double)p0.y - ml1, d2 = (double)p0.z - ml2;
          const double e0 = (double)p1.x - ml0, e1 = (double)p1.y - ml1, e2 = (double)p1.z - ml2;
          c00 += d0 * d0, c01 += d0 * d1, c02 += d0 * d2, c11 += d1 * d1, c12 += d1 * d2, c22 += d2 * d2;
          c00 += e0 * e0, c01 += e0 * e1, c02 += e0 * e2, c11 += e1 * e1, c12 += e1 * e2, c22 += e2 * e2;
        }
        for (; q < n; q++) {
          const float4 p = pts[ids[q]];
          const double d0 = (double)p.x - ml0, d1 = (double)p.y - ml1, d2 = (double)p.z - ml2;
          c00 += d0 * d0, c01 += d0 * d1, c02 += d0 * d2, c11 += d1 * d1, c12 += d1 * d2, c22 += d2 * d2;
        }
        const double ms[3] = {ms0, ms1, ms2}, ml[3] = {ml0, ml1, ml2};
        const double csum[9] = {c00, c01, c02, c01, c11, c12, c02, c12, c22};
        if (!has) {
          for (int q2 = 0; q2 < 3; q2++) mean[q2] = ml[q2];
          for (int q2 = 0; q2 < 9; q2++) cov[q2] = csum[q2] / (double)(n - 1);
          N = n;
        } else {  // pairwise (Chan) merge with the stored (N, mean, cov)
          const double N0 = (double)N, n1 = (double)n;
          double mS[3], cS[9], tv[3];
          for (int q2 = 0; q2 < 3; q2++) mS[q2] = mean[q2] * N0;
          for (int q2 = 0; q2 < 9; q2++) cS[q2] = cov[q2] * (N0 - 1.0);
          const double w = N0 / (n1 * (N0 + n1));
          for (int q2 = 0; q2 < 3; q2++) tv[q2] = (n1 / N0) * mS[q2] - ms[q2];
          for (int a = 0; a < 3; a++)
            for (int bb = 0; bb < 3; bb++) cS[a * 3 + bb] += csum[a * 3 + bb] + w * tv[a] * tv[bb];
          for (int q2 = 0; q2 < 3; q2++) mS[q2] += ms[q2];
          double Nt = N0 + n1;
          for (int q2 = 0; q2 < 3; q2++) mean[q2] = mS[q2] / Nt;
          for (int q2 = 0; q2 < 9; q2++) cov[q2] = cS[q2] / (Nt - 1.0);
          if (Nt > (double)j.maxnumpoints) Nt = (double)j.maxnumpoints;
          N = (int)Nt;
        }
        eigen = true;
        has = 0;  // decided by k_eigen
      }
    }
    for (int q = 0; q < 3; q++) j.cmean[(size_t)c * 3 + q] = mean[q];
    for (int q = 0; q < 9; q++) j.ccov[(size_t)c * 9 + q] = cov[q];
    j.cn[c] = N, j.chas[c] = has, j.cocc[c] = occ;
    if (eigen) j.cursor[atomicAdd(j.counts + 5, 1)] = c;
  }
}

// (c) one thread per listed cell: NDTCell::rescaleCovariance (eigen clamp), decides hasGaussian_.  Almost every cell
// converges in 2-4 Jacobi sweeps; the ~2 % with an exactly singular covariance (collinear points) never meet the
// stopping test and run the full 64 sweeps — in a warp of 32 cells one of them would make the other 31 wait 20x longer,
// so the first pass stops after EIG_FAST_SWEEPS and defers those cells to a compact second list (the dead `cell_key`
// array) that k_eigen_hard works through with full warps.  Same operations per cell either way: bit-identical results.
constexpr int EIG_FAST_SWEEPS = 8;
#ifndef NDTB_EIGEN_MINBLOCKS
#define NDTB_EIGEN_MINBLOCKS 6  // B200 A/B of k_eigen: 6 (80 registers) 1.19 ms, 5 (96) 1.29 ms, 4 (114, no spills) 1.46 ms
#endif

__global__ void __launch_bounds__(128, NDTB_EIGEN_MINBLOCKS) k_eigen(const BuildJob *__restrict__ jobs) {
  const BuildJob &j = jobs[blockIdx.y];
  const int n = j.counts[5];
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const int c = j.cursor[t];
    double cov[9];
#pragma unroll
    for (int q = 0; q < 9; q++) cov[q] = j.ccov[(size_t)c * 9 + q];
    bool deferred;
    const bool ok = rescale_covariance(cov, EIG_FAST_SWEEPS, &deferred);
    if (deferred) {
      j.cell_key[atomicAdd(j.counts + 6, 1)] = c;
      continue;
    }
#pragma unroll
    for (int q = 0; q < 9; q++) j.ccov[(size_t)c * 9 + q] = cov[q];
    j.chas[c] = ok ? 1 : 0;
  }
}

__global__ void __launch_bounds__(128, 4) k_eigen_hard(const BuildJob *__restrict__ jobs) {
  const BuildJob &j = jobs[blockIdx.y];
  const int n = j.counts[6];
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const int c = j.cell_key[t];
    double cov[9];
#pragma unroll
    for (int q = 0; q < 9; q++) cov[q] = j.ccov[(size_t)c * 9 + q];
    const bool ok = rescale_covariance(cov);
#pragma unroll
    for (int q = 0; q < 9; q++) j.ccov[(size_t)c * 9 + q] = cov[q];
    j.chas[c] = ok ? 1 : 0;
  }
}

// ---- Gaussian view ---------------------------------------------------------------------------------------
// one CTA per map over the touched blocks: per-block Gaussian mask, exclusive scan -> base of the block in gcell
__global__ void __launch_bounds__(1024) k_gscan(const BuildJob *__restrict__ jobs) {
  const BuildJob &j = jobs[blockIdx.x];
  __shared__ int wsum[32];
  const int ntb = j.counts[1];
  int run_c = 0, run_b = 0;
  for (int base = 0; base < ntb; base += 1024) {
    const int t = base + threadIdx.x;
    unsigned long long gm = 0ull;
    if (t < ntb) {
      const int b = j.tb_list[t];
      unsigned long long m = j.amask[b];
      int c = j.abase[b];
      for (; m; m &= m - 1ull, c++)
        if (j.chas[c]) gm |= 1ull << (__ffsll((long long)m) - 1);
    }
    int tc, tb;
    const int ec = cta_excl_scan(__popcll(gm), &tc, wsum);
    cta_excl_scan(gm != 0ull, &tb, wsum);
    if (t < ntb) j.gmask_t[t] = gm, j.gbase_t[t] = run_c + ec;
    run_c += tc, run_b += tb;
  }
  if (threadIdx.x == 0) j.counts[2] = run_c, j.counts[3] = run_b;
}

// thread per touched block: publish the block in the hash table, copy its Gaussian cells to the compact view
__global__ void k_gfill(const BuildJob *__restrict__ jobs) {
  const BuildJob &j = jobs[blockIdx.y];
  const int ntb = j.counts[1];
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < ntb; t += gridDim.x * blockDim.x) {
    const unsigned long long gm = j.gmask_t[t];
    if (!gm) continue;
    const int b = j.tb_list[t];
    int gx, gy, gz;
    sblock_coords(j, b, gx, gy, gz);
    const int gkey = (gx * j.g.nb[1] + gy) * j.g.nb[2] + gz;  // the matcher probes with block ids of the full grid
    unsigned h = hash_block(gkey, j.tsize);
    for (;;) {
      const int prev = atomicCAS(&j.table[h].key, -1, gkey);
      if (prev == -1) break;
      h = (h + 1) & (unsigned)(j.tsize - 1);
    }
    j.table[h].base = j.gbase_t[t];
    j.table[h].mask = gm;
    unsigned long long m = j.amask[b];
    int c = j.abase[b], gs = j.gbase_t[t];
    for (; m; m &= m - 1ull, c++) {
      const int bit = __ffsll((long long)m) - 1;
      if (!(gm >> bit & 1ull)) continue;
      j.g2c[gs] = c;  // the record itself is copied by k_gcopy, one thread per double
      gs++;
    }
  }
}
// compact Gaussian view: gcell[g] = mean(3) + upper triangle of cov(6) of cell g2c[g]; one thread per output double
__global__ void k_gcopy(const BuildJob *__restrict__ jobs) {
  const BuildJob &j = jobs[blockIdx.y];
  const int total = j.counts[2] * GC;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int g = e / GC, q = e - g * GC;
    const int c = j.g2c[g];
    const int cq = q == 3 ? 0 : (q == 4 ? 1 : (q == 5 ? 2 : (q == 6 ? 4 : (q == 7 ? 5 : 8))));
    j.gcell[e] = q < 3 ? j.cmean[(size_t)c * 3 + q] : j.ccov[(size_t)c * 9 + cq];
  }
}

// ---- export / from_cells / parity hooks ------------------------------------------------------------------
__global__ void k_export(const BuildJob *__restrict__ jobs, ndtb_cell *__restrict__ out) {
  const BuildJob &j = jobs[0];
  const int ntb = j.counts[1];
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < ntb; t += gridDim.x * blockDim.x) {
    const int b = j.tb_list[t];
    int bx, by, bz;
    sblock_coords(j, b, bx, by, bz);
    unsigned long long m = j.amask[b];
    int c = j.abase[b];
    for (; m; m &= m - 1ull, c++) {
      const int bit = __ffsll((long long)m) - 1;
      ndtb_cell r;
      for (int q = 0; q < 3; q++) r.mean[q] = j.cmean[(size_t)c * 3 + q];
      const double *cv = j.ccov + (size_t)c * 9;
      r.cov[0] = cv[0], r.cov[1] = cv[1], r.cov[2] = cv[2], r.cov[3] = cv[4], r.cov[4] = cv[5], r.cov[5] = cv[8];
      r.n = j.cn[c], r.has_gaussian = j.chas[c];
      r.idx[0] = bx * 4 + (bit >> 4), r.idx[1] = by * 4 + ((bit >> 2) & 3), r.idx[2] = bz * 4 + (bit & 3);
      r.occ = j.cocc[c];
      out[c] = r;
    }
  }
}

// from_cells: mark + place records given voxel indices (host resolved or from the mean)
__global__ void k_cells_voxel(const BuildJob *__restrict__ jobs, const ndtb_cell *__restrict__ cells, int n, int use_idx,
                              int *__restrict__ vox /*[n]: key or -1*/, int *__restrict__ err) {
  const BuildJob &j = jobs[0];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int ix, iy, iz;
    bool ok = true;
    if (use_idx)
      ix = cells[i].idx[0], iy = cells[i].idx[1], iz = cells[i].idx[2];
    else
      ok = voxel_index(j.g, (double)(float)cells[i].mean[0], (double)(float)cells[i].mean[1], (double)(float)cells[i].mean[2],
                       ix, iy, iz);
    if (!ok || !in_grid(j.g, ix, iy, iz)) {
      vox[i] = -1;
      atomicExch(err, 1);
      continue;
    }
    const int b = sblock_id(j, ix, iy, iz), bit = block_bit(ix, iy, iz);
    if (b < 0) {
      vox[i] = -1;
      atomicExch(err, 1);
      continue;
    }
    vox[i] = b * 64 + bit;
    atomicOr(j.amask + b, 1ull << bit);
  }
}
__global__ void k_cells_place(const BuildJob *__restrict__ jobs, const ndtb_cell *__restrict__ cells, int n,
                              const int *__restrict__ vox) {
  const BuildJob &j = jobs[0];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int key = vox[i];
    if (key < 0) continue;
    const int b = key >> 6, bit = key & 63;
    const int c = j.abase[b] + __popcll(j.amask[b] & ((1ull << bit) - 1ull));
    for (int q = 0; q < 3; q++) j.cmean[(size_t)c * 3 + q] = cells[i].mean[q];
    const double *t = cells[i].cov;
    const double full[9] = {t[0], t[1], t[2], t[1], t[3], t[4], t[2], t[4], t[5]};
    for (int q = 0; q < 9; q++) j.ccov[(size_t)c * 9 + q] = full[q];
    j.cn[c] = cells[i].n, j.chas[c] = cells[i].has_gaussian != 0, j.cocc[c] = cells[i].occ;
  }
}

__global__ void k_point_indices(GridDesc g, const float4 *__restrict__ pts, int n, int *__restrict__ out,
                                int *__restrict__ n_in) {
  int local = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 p = pts[i];
    int ix, iy, iz;
    const bool ok = !(isnan(p.x) || isnan(p.y) || isnan(p.z)) && voxel_index(g, (double)p.x, (double)p.y, (double)p.z, ix, iy, iz);
    if (!ok) ix = iy = iz = INT32_MIN;
    out[3 * i] = ix, out[3 * i + 1] = iy, out[3 * i + 2] = iz;
    if (ok && in_grid(g, ix, iy, iz)) local++;
  }
  local = __reduce_add_sync(FULL, local);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(n_in, local);
}

// NDTMatcherP2D: a point is handed to the D2D kernels as a cell with mean p and zero covariance (DESIGN.md "P2D").
// NaN points keep a NaN mean: the matcher's voxel lookup rejects them, like the CPU restatement which drops them.
__global__ void k_points_as_cells(const float4 *__restrict__ pts, int n, double *__restrict__ gcell) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 p = pts[i];
    double *o = gcell + (size_t)i * GC;
    o[0] = (double)p.x, o[1] = (double)p.y, o[2] = (double)p.z;
#pragma unroll
    for (int q = 3; q < GC; q++) o[q] = 0.0;
  }
}

// ndt_feature::overlapNDTOccupancyScore (ndt_feature/include/ndt_feature/ndt_feature_node.h:213-252):
// mean squared difference of rescaled occupancies over the initialised cells of `mov` mapped into `ref`.
// Integer-free sums are accumulated per cell in (block, bit) order by ONE thread block with a fixed tree.
__device__ __forceinline__ double occ_rescaled(float occ) { return 1.0 - 1.0 / (1.0 + exp((double)occ)); }
// one CTA per link: jobs[2*link] = ref, jobs[2*link+1] = mov, T16 + 16*link (T_stride doubles between links), out[link]
__global__ void __launch_bounds__(256) k_overlap(const BuildJob *__restrict__ jobs /*[0]=ref,[1]=mov*/, const double *__restrict__ T16,
                                                  int T_stride, double *__restrict__ out_all) {
  const BuildJob &ref = jobs[2 * blockIdx.x], &mov = jobs[2 * blockIdx.x + 1];
  const Pose T = pose_from_cm(T16 + (size_t)T_stride * blockIdx.x);
  double *out = out_all + blockIdx.x;
  if (mov.counts == nullptr || ref.counts == nullptr) {  // a map without cells: the reference returns 1
    if (threadIdx.x == 0) out[0] = 1.0;
    return;
  }
  double sum = 0.0;
  int nb = 0;
  const int ntb = mov.counts[1];
  for (int t = threadIdx.x; t < ntb; t += blockDim.x) {
    const int b = mov.tb_list[t];
    int bx, by, bz;
    sblock_coords(mov, b, bx, by, bz);
    unsigned long long m = mov.amask[b];
    int c = mov.abase[b];
    for (; m; m &= m - 1ull, c++) {
      const int bit = __ffsll((long long)m) - 1;
      const double mo = occ_rescaled(mov.cocc[c]);
      if (mo == 0.5) continue;
      const int idx[3] = {bx * 4 + (bit >> 4), by * 4 + ((bit >> 2) & 3), bz * 4 + (bit & 3)};
      double e[3];
      for (int a = 0; a < 3; a++) {
        const int idc = (int)(mov.g.size[a] / 2.0);
        e[a] = (double)(float)(mov.g.center[a] + (idx[a] - idc) * mov.g.cell[a]);  // NDTCell::getCenter is a float point
      }
      float pt[3];
      for (int a = 0; a < 3; a++) pt[a] = (float)((T.R[a * 3] * e[0] + T.R[a * 3 + 1] * e[1] + T.R[a * 3 + 2] * e[2]) + T.t[a]);
      int ix, iy, iz;
      if (!voxel_index(ref.g, (double)pt[0], (double)pt[1], (double)pt[2], ix, iy, iz) || !in_grid(ref.g, ix, iy, iz)) continue;
      const int rb = sblock_id(ref, ix, iy, iz), rbit = block_bit(ix, iy, iz);
      if (rb < 0) continue;
      const unsigned long long rm = ref.amask[rb];
      if (!(rm >> rbit & 1ull)) continue;
      const double ro = occ_rescaled(ref.cocc[ref.abase[rb] + __popcll(rm & ((1ull << rbit) - 1ull))]);
      if (ro != 0.5) nb++, sum += (mo - ro) * (mo - ro);
    }
  }
  __shared__ double ssum[8];
  __shared__ int scnt[8];
  for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(FULL, sum, off), nb += __shfl_xor_sync(FULL, nb, off);
  if ((threadIdx.x & 31) == 0) ssum[threadIdx.x >> 5] = sum, scnt[threadIdx.x >> 5] = nb;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0;
    int n = 0;
    for (int w = 0; w < 8; w++) s += ssum[w], n += scnt[w];
    out[0] = n == 0 ? 1.0 : s / (1.0 * n);
  }
}

// lslgeneric::transformPointCloudInPlace [upstream]: the pose is cast to FLOAT and applied in float, column by column
// (Eigen's 3x3 matrix * vector accumulates col0*x + col1*y + col2*z), then the translation is added.  T12 = R row-major, t.
__global__ void k_transform_points(const float4 *__restrict__ in, float4 *__restrict__ out, int n, const float *__restrict__ T12) {
  float T[12];
#pragma unroll
  for (int q = 0; q < 12; q++) T[q] = T12[q];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 p = in[i];
    float4 o;
    o.x = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[0], p.x), __fmul_rn(T[1], p.y)), __fmul_rn(T[2], p.z)), T[9]);
    o.y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[3], p.x), __fmul_rn(T[4], p.y)), __fmul_rn(T[5], p.z)), T[10]);
    o.z = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[6], p.x), __fmul_rn(T[7], p.y)), __fmul_rn(T[8], p.z)), T[11]);
    o.w = p.w;
    out[i] = o;
  }
}

// ---- launch wrappers (host) ---------------------------------------------------------------------------
#ifndef NDTB_PTS_PER_CTA
#define NDTB_PTS_PER_CTA 16384  // points per CTA of the streaming per-point kernels (grid-stride loops; B200 A/B per 1184 maps: 1024 -> 14.3 ms, 4096 -> 13.0 ms, 16384 -> 12.6 ms)
#endif
static inline int chunks_for(int n, int per) {
  int c = (n + per - 1) / per;
  return c < 1 ? 1 : (c > 1024 ? 1024 : c);
}

int centroid_chunk_points() { return CEN_CHUNK; }
int centroid_record_doubles() { return CEN_REC; }
int launch_guess(const BuildJob *d_jobs, const int *d_which, int n_which, int max_pts, const long long *d_rec_off, double *d_recs,
                 double *d_out, cudaStream_t s) {
  const int max_chunks = (max_pts + CEN_CHUNK - 1) / CEN_CHUNK;
  k_centroid_chunks<<<dim3(chunks_for(max_chunks, 4), n_which), 128, 0, s>>>(d_jobs, d_which, d_rec_off, d_recs);
  k_centroid<<<n_which, 32, 0, s>>>(d_jobs, d_which, d_rec_off, d_recs, d_out);
  k_extent<<<dim3(chunks_for(max_pts, NDTB_PTS_PER_CTA), n_which), 256, 0, s>>>(d_jobs, d_which, d_out);
  return 3;
}
static int launch_group_by_cell(const BuildJob *d_jobs, int n, int max_items, int max_cells, cudaStream_t s);
int launch_mark(const BuildJob *d_jobs, int n, int max_pts, bool trace, bool fast, cudaStream_t s) {
  if (fast && !trace) k_mark<true><<<dim3(chunks_for(max_pts, NDTB_PTS_PER_CTA), n), 256, 0, s>>>(d_jobs);
  else k_mark<false><<<dim3(chunks_for(max_pts, NDTB_PTS_PER_CTA), n), 256, 0, s>>>(d_jobs);
  if (trace) {
    k_trace_count<<<dim3(chunks_for(max_pts, 128), n), 128, 0, s>>>(d_jobs);
    k_rayscan<<<n, 1024, 0, s>>>(d_jobs);
  }
  k_blockscan<<<n, 1024, 0, s>>>(d_jobs);
  return trace ? 4 : 2;
}
// per-cell visit lists: fill the ray-major visit arrays, then the counting sort the points go through (d_vjobs: npts =
// number of visits, pt_cell = vis_key, cnt / seg_off / seg_idx / seg2 = the visit arrays)
int launch_trace_lists(const BuildJob *d_jobs, const BuildJob *d_vjobs, int n, int max_pts, int max_vis, int max_ntb, int max_cells,
                       cudaStream_t s) {
  k_trace_fill<<<dim3(chunks_for(max_pts, 128), n), 128, 0, s>>>(d_jobs);
  const int ns = launch_group_by_cell(d_vjobs, n, max_vis, max_cells, s);
  k_cell_keys<<<dim3(chunks_for(max_ntb, 128), n), 128, 0, s>>>(d_jobs);
  k_cell_trace<<<dim3(chunks_for(max_cells, 128), n), 128, 0, s>>>(d_jobs);
  return ns + 3;
}
int launch_transform_points(const float4 *d_in, float4 *d_out, int n, const float *d_T12, cudaStream_t s) {
  k_transform_points<<<chunks_for(n, 1024), 256, 0, s>>>(d_in, d_out, n, d_T12);
  return 1;
}
int sort_tile_points() { return RS_TILE; }
int sort_passes(int max_cells) {
  int bits = 1;
  while ((1ll << bits) <= (long long)max_cells) bits++;
  return (bits + 7) / 8;
}
// key -> cell id, stable radix sort of the ids by cell id, per-cell segments.  Returns the number of launches.
static int launch_group_by_cell(const BuildJob *d_jobs, int n, int max_items, int max_cells, cudaStream_t s) {
  const int tiles = (max_items + RS_TILE - 1) / RS_TILE > 0 ? (max_items + RS_TILE - 1) / RS_TILE : 1;
  const int P = sort_passes(max_cells);
  k_count<<<dim3(chunks_for(max_items, NDTB_PTS_PER_CTA), n), 256, 0, s>>>(d_jobs);
  for (int p = 0; p < P; p++) {
    k_rs_hist<<<dim3(tiles, n), RS_THREADS, 0, s>>>(d_jobs, 8 * p, p & 1);
    k_rs_scan<<<n, 1024, 0, s>>>(d_jobs);
    k_rs_scatter<<<dim3(tiles, n), RS_THREADS, 0, s>>>(d_jobs, 8 * p, p & 1, p == 0);
  }
  k_seg_bounds<<<dim3(chunks_for(max_items, NDTB_PTS_PER_CTA), n), 256, 0, s>>>(d_jobs, P & 1);
  k_seg_counts<<<dim3(chunks_for(max_cells, 256), n), 256, 0, s>>>(d_jobs);
  return 3 + 3 * P;
}
int launch_cells(const BuildJob *d_jobs, int n, int max_pts, int max_ntb, int max_cells, cudaStream_t s) {
  const int ns = launch_group_by_cell(d_jobs, n, max_pts, max_cells, s);
  k_cell_keys<<<dim3(chunks_for(max_ntb, 128), n), 128, 0, s>>>(d_jobs);
  k_cells<<<dim3(chunks_for(max_cells, 128), n), 128, 0, s>>>(d_jobs);
  k_eigen<<<dim3(chunks_for(max_cells, 128), n), 128, 0, s>>>(d_jobs);
  k_eigen_hard<<<dim3(chunks_for(max_cells / 16 + 1, 128), n), 128, 0, s>>>(d_jobs);
  return ns + 4;
}
int launch_gview(const BuildJob *d_jobs, int n, int max_ntb, int max_cells, cudaStream_t s) {
  k_gscan<<<n, 1024, 0, s>>>(d_jobs);
  k_gfill<<<dim3(chunks_for(max_ntb, 128), n), 128, 0, s>>>(d_jobs);
  k_gcopy<<<dim3(chunks_for(max_cells * GC, 1024), n), 256, 0, s>>>(d_jobs);
  return 3;
}
int launch_blockscan(const BuildJob *d_jobs, int n, cudaStream_t s) {
  k_blockscan<<<n, 1024, 0, s>>>(d_jobs);
  return 1;
}
int launch_export(const BuildJob *d_job, int ntb, ndtb_cell *d_out, cudaStream_t s) {
  k_export<<<chunks_for(ntb, 128), 128, 0, s>>>(d_job, d_out);
  return 1;
}
int launch_from_cells_voxel(const BuildJob *d_job, const ndtb_cell *d_cells, int n, int use_idx, int *d_vox, int *d_err,
                            cudaStream_t s) {
  k_cells_voxel<<<chunks_for(n, 256), 256, 0, s>>>(d_job, d_cells, n, use_idx, d_vox, d_err);
  return 1;
}
int launch_from_cells_place(const BuildJob *d_job, const ndtb_cell *d_cells, int n, const int *d_vox, cudaStream_t s) {
  k_cells_place<<<chunks_for(n, 256), 256, 0, s>>>(d_job, d_cells, n, d_vox);
  return 1;
}
int launch_point_indices(const GridDesc &g, const float4 *d_pts, int n, int *d_out, int *d_nin, cudaStream_t s) {
  k_point_indices<<<chunks_for(n, 1024), 256, 0, s>>>(g, d_pts, n, d_out, d_nin);
  return 1;
}
int launch_points_as_cells(const float4 *d_pts, int n, double *d_gcell, cudaStream_t s) {
  k_points_as_cells<<<chunks_for(n, 1024), 256, 0, s>>>(d_pts, n, d_gcell);
  return 1;
}
int launch_overlap(const BuildJob *d_jobs2, int n_links, const double *d_T16, int T_stride, double *d_out, cudaStream_t s) {
  k_overlap<<<n_links, 256, 0, s>>>(d_jobs2, d_T16, T_stride, d_out);
  return 1;
}

}  // namespace ndtb
