// d2d.cu — kernel (ii): NDT-D2D score / gradient / Hessian accumulation, the device-resident Newton +
// More-Thuente loop around it, and the covariance pass.
//
// Reference path replaced:
//   a9  NDTMatcherD2D::derivativesNDT [upstream]   called at ndt_feature/include/ndt_feature/ndt_matcher_d2d_fusion.h:856,617,444
//   a14 NDTMap::pseudoTransformNDT / per-iteration cell move (fusion.h:840, :1047-1056) — fused: every pass
//       applies the current pose to the immutable source cells on the fly, nothing is written back
//   a15 LazyGrid::getClosestNDTCells [upstream] — (2k+1)^3 cube probe, done per 4x4x4 block (8 hash probes for k=2)
//   a8/a5 NDTMatcherD2D::match / matchFusion (fusion.h:797-1155) — optimizer.h state machine, one thread
//   a11 NDTMatcherD2D::covariance [upstream] (ndt_feature_graph.cpp:298)
//
// Work decomposition of one derivative pass: a warp takes 32 source cells per round (one per lane): each lane
// moves its cell, stages (mu, C) in shared memory, probes the target's block table and pushes its hits
// (source lane, target slot) into a per-warp queue; the queue is drained 32 pairs at a time, ONE PAIR PER LANE,
// so the fp64 pair arithmetic always runs on full warps although hit counts per source cell vary from 0 to 125.
// All reductions use fixed trees (shuffle butterflies, ordered partial sums): results are run-to-run deterministic.
#include <cooperative_groups.h>

#include "../../include/ndtb.h"
#include "engine.cuh"

namespace cg = cooperative_groups;

namespace ndtb {

#ifndef NDTB_MATCH_THREADS
#define NDTB_MATCH_THREADS 384  // 12 warps at 168 registers (B200 A/B, 592 C2 pairs: 256 thr 62.1 ms, 384 thr 56.8 ms, 512 thr + Hessian sums in shared memory 57.1 ms)
#endif
#ifndef NDTB_GRAD2
#define NDTB_GRAD2 0  // gradient pass: two pairs per lane interleaved (measured slower on B200: spills)
#endif
#ifndef NDTB_PUSH_SCAN
#define NDTB_PUSH_SCAN 0  // 1: one warp prefix scan per probe column; 0: one ballot per pushed hit (faster from 384 threads up)
#endif
#ifndef NDTB_HESS_SMEM
#define NDTB_HESS_SMEM 0  // 1: Hessian sums in per-lane shared-memory slots (needed at 512 threads / 128 registers); 0: in registers
#endif
#ifndef NDTB_MATCH_MINBLOCKS
#define NDTB_MATCH_MINBLOCKS 1
#endif
constexpr int MATCH_THREADS = NDTB_MATCH_THREADS;
constexpr int MATCH_WARPS = MATCH_THREADS / 32;
constexpr int QCAP = 256;       // per-warp pair queue capacity (entries); >= 64
constexpr int MAXSPAN = 3;      // 4-blocks per axis touched by a (2k+1)-voxel probe window, k <= 4
constexpr int ACC_PAIRS = 28;   // slot counting contributing pairs
constexpr int ACC_TOTAL = 29;
constexpr unsigned FULL = 0xffffffffu;

struct WarpScratch {
  double stage[GC][32];  // moved source cells of the current round, SoA: stage[k][lane]
  unsigned queue[QCAP];  // (source lane << 27) | target slot
};

// Hessian sums of the registration kernel: one shared-memory slot per lane and entry, hs[i * MATCH_THREADS + tid] (conflict
// free), so that the 21 running sums do not compete with the pair arithmetic for registers.
struct SmemAcc {
  static constexpr bool kPairHook = false;
  double *slot;  // &hs[tid]
  __device__ __forceinline__ void add(int i, double v) const { slot[i * MATCH_THREADS] += v; }
};

// Covariance pass (NDTMatcherD2D::covariance [upstream], ndt_feature_graph.cpp:298): a Hessian pass that also needs
// the rows of J^T J — the gradient sum of every SOURCE cell (all its pairs) and of every TARGET cell (all pairs that hit
// it).  Pairs are spread over the lanes, so both kinds of rows are summed with 64-bit FIXED-POINT atomics (2^-36
// resolution): integer addition is associative, the result does not depend on the order the pairs arrive in.
constexpr double COV_FIX = 68719476736.0;  // 2^36
#ifndef NDTB_COV_THREADS
#define NDTB_COV_THREADS 384  // B200 A/B per 592 C2 pairs: 256 thr (227 regs) 5.38 ms, 384 thr (168 regs) 4.37 ms, 512 thr (128 regs, spills) 4.64 ms
#endif
constexpr int COV2_THREADS = NDTB_COV_THREADS;
struct CovAcc {
  static constexpr bool kPairHook = true;
  double *acc;                // Hessian sums (21 updates per pair): the thread's registers, like RegAcc
  double *gg;                 // sum g g^T over the source rows (21 updates per SOURCE CELL): per-lane shared-memory slots
  unsigned long long *gs;     // [32][6] source rows of the warp's current round (shared)
  unsigned long long *gt;     // [n target cells][6] target rows of this registration (global)
  __device__ __forceinline__ void add(int i, double v) const { acc[ACC_H + i] += v; }
  __device__ __forceinline__ void pair(int src_lane, int tgt_slot, const double *g6) const {
#pragma unroll
    for (int a = 0; a < 6; a++) {
      const unsigned long long v = (unsigned long long)__double2ll_rn(g6[a] * COV_FIX);
      atomicAdd(gs + src_lane * 6 + a, v);
      atomicAdd(gt + (size_t)tgt_slot * 6 + a, v);
    }
  }
};

struct PassCtx {
  const GridDesc *g;  // target grid (shared memory)
  const HashEntry *table;
  int tsize;
  const double *tcell;
  const double *scell;
  int ns;
  int k;
  double lfd1, lfd2;
};

__device__ __forceinline__ bool table_find(const HashEntry *table, int tsize, int key, int &base,
                                           unsigned long long &mask) {
  unsigned h = hash_block(key, tsize);
#pragma unroll 1
  for (;;) {
    const int4 e = *reinterpret_cast<const int4 *>(table + h);
    if (e.x == key) {
      base = e.y;
      mask = ((unsigned long long)(unsigned)e.w << 32) | (unsigned)e.z;
      return true;
    }
    if (e.x == -1) return false;
    h = (h + 1) & (unsigned)(tsize - 1);
  }
}

// 4-bit mask of local coordinates [lo,hi] of block b (4 voxels) clipped to the probe range [c-k, c+k]
__device__ __forceinline__ unsigned axis_bits(int c, int k, int b) {
  int lo = c - k - 4 * b, hi = c + k - 4 * b;
  lo = lo < 0 ? 0 : lo;
  hi = hi > 3 ? 3 : hi;
  return ((2u << hi) - 1u) & ~((1u << lo) - 1u);
}
__device__ __forceinline__ unsigned long long expand_x(unsigned a) {
  return ((a & 1u) ? 0xFFFFull : 0ull) | ((a & 2u) ? 0xFFFFull << 16 : 0ull) | ((a & 4u) ? 0xFFFFull << 32 : 0ull) |
         ((a & 8u) ? 0xFFFFull << 48 : 0ull);
}
__device__ __forceinline__ unsigned long long expand_y(unsigned a) {
  const unsigned long long row = ((a & 1u) ? 0xFull : 0ull) | ((a & 2u) ? 0xF0ull : 0ull) | ((a & 4u) ? 0xF00ull : 0ull) |
                                 ((a & 8u) ? 0xF000ull : 0ull);
  return row * 0x0001000100010001ull;
}
__device__ __forceinline__ unsigned long long expand_z(unsigned a) { return (unsigned long long)a * 0x1111111111111111ull; }

template <bool HESS, class HA>
__device__ __forceinline__ void process_entry(const PassCtx &c, const WarpScratch &ws, unsigned e, double *acc, const HA &ha) {
  const int sl = e >> 27;
  const int slot = e & 0x7FFFFFF;
  double C[6], m[3], S[6];
  const double mu0 = ws.stage[0][sl], mu1 = ws.stage[1][sl], mu2 = ws.stage[2][sl];
#pragma unroll
  for (int j = 0; j < 6; j++) C[j] = ws.stage[3 + j][sl];
  const double *t = c.tcell + (size_t)slot * GC;
#pragma unroll
  for (int j = 0; j < 3; j++) m[j] = __ldg(t + j);
#pragma unroll
  for (int j = 0; j < 6; j++) S[j] = __ldg(t + 3 + j);
  double g6[6];
  if (pair_contrib_acc<HESS, HA>(mu0, mu1, mu2, C, m, S, c.lfd1, c.lfd2, acc, ha, HA::kPairHook ? g6 : nullptr)) {
    acc[ACC_PAIRS] += 1.0;
    if constexpr (HA::kPairHook) ha.pair(sl, slot, g6);
  }
}

// gradient pass: two pairs per lane at once (independent dependency chains interleave, hiding the fp64 / exp / load
// latencies that two resident warps per scheduler cannot hide)
__device__ __forceinline__ void process_entry2(const PassCtx &c, const WarpScratch &ws, unsigned e0, unsigned e1, bool live1,
                                               double *acc) {
  const int sl0 = e0 >> 27, sl1 = e1 >> 27;
  const double *t0 = c.tcell + (size_t)(e0 & 0x7FFFFFF) * GC, *t1 = c.tcell + (size_t)(e1 & 0x7FFFFFF) * GC;
  double C0[6], m0[3], S0[6], C1[6], m1[3], S1[6];
#pragma unroll
  for (int j = 0; j < 3; j++) m0[j] = __ldg(t0 + j), m1[j] = __ldg(t1 + j);
#pragma unroll
  for (int j = 0; j < 6; j++) S0[j] = __ldg(t0 + 3 + j), S1[j] = __ldg(t1 + 3 + j);
#pragma unroll
  for (int j = 0; j < 6; j++) C0[j] = ws.stage[3 + j][sl0], C1[j] = ws.stage[3 + j][sl1];
  pair_grad_nb(ws.stage[0][sl0], ws.stage[1][sl0], ws.stage[2][sl0], C0, m0, S0, c.lfd1, c.lfd2, true, acc, acc + ACC_PAIRS);
  pair_grad_nb(ws.stage[0][sl1], ws.stage[1][sl1], ws.stage[2][sl1], C1, m1, S1, c.lfd1, c.lfd2, live1, acc, acc + ACC_PAIRS);
}

template <bool HESS, class HA>
__device__ __forceinline__ void drain_batches(const PassCtx &c, WarpScratch &ws, int &qcount, int keep, int lane, double *acc,
                                              const HA &ha) {
  // process full batches of 32 pairs from the top of the queue until fewer than `keep` + 32 entries are left
  if (!HESS && NDTB_GRAD2) {
    while (qcount >= keep + 64) {
      qcount -= 64;
      process_entry2(c, ws, ws.queue[qcount + lane], ws.queue[qcount + 32 + lane], true, acc);
    }
  }
  while (qcount >= keep + 32) {
    qcount -= 32;
    process_entry<HESS, HA>(c, ws, ws.queue[qcount + lane], acc, ha);
  }
}

// moved source cell: mean <- R mean + t (bit-exact operation order of the CPU restatement, the voxel of the
// moved mean selects the neighbourhood), cov <- R cov R^T
__device__ __forceinline__ void move_cell(const double *P, const double *s, double *mu, double *C) {
  const double m0 = s[0], m1 = s[1], m2 = s[2];
#pragma unroll
  for (int r = 0; r < 3; r++)
    mu[r] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(P[r * 3], m0), __dmul_rn(P[r * 3 + 1], m1)), __dmul_rn(P[r * 3 + 2], m2)),
                      P[9 + r]);
  const double s00 = s[3], s01 = s[4], s02 = s[5], s11 = s[6], s12 = s[7], s22 = s[8];
  double A[9];  // A = R * Sigma
#pragma unroll
  for (int r = 0; r < 3; r++) {
    const double r0 = P[r * 3], r1 = P[r * 3 + 1], r2 = P[r * 3 + 2];
    A[r * 3 + 0] = r0 * s00 + r1 * s01 + r2 * s02;
    A[r * 3 + 1] = r0 * s01 + r1 * s11 + r2 * s12;
    A[r * 3 + 2] = r0 * s02 + r1 * s12 + r2 * s22;
  }
  C[0] = A[0] * P[0] + A[1] * P[1] + A[2] * P[2];
  C[1] = A[0] * P[3] + A[1] * P[4] + A[2] * P[5];
  C[2] = A[0] * P[6] + A[1] * P[7] + A[2] * P[8];
  C[3] = A[3] * P[3] + A[4] * P[4] + A[5] * P[5];
  C[4] = A[3] * P[6] + A[4] * P[7] + A[5] * P[8];
  C[5] = A[6] * P[6] + A[7] * P[7] + A[8] * P[8];
}

// One derivative pass over the source cells assigned to this warp (rounds wg, wg+nwg, ...).
template <bool HESS, class HA>
__device__ void d2d_pass(const PassCtx &c, const double *P, WarpScratch &ws, int wg, int nwg, double *acc, const HA &ha) {
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const int k = c.k;
  const int span = (2 * k + 3) / 4 + 1;  // max 4-blocks touched by 2k+1 consecutive voxels
  const int nb0 = c.g->nb[0], nb1 = c.g->nb[1], nb2 = c.g->nb[2];
  int qcount = 0;
  for (int base = wg * 32; base < c.ns; base += nwg * 32) {
    const int i = base + lane;
    bool valid = i < c.ns;
    int ix = 0, iy = 0, iz = 0;
    if (valid) {
      double mu[3], C[6];
      move_cell(P, c.scell + (size_t)i * GC, mu, C);
      // pcl::PointXYZ is float: the lookup point is the float-rounded mean
      valid = voxel_index(*c.g, (double)__double2float_rn(mu[0]), (double)__double2float_rn(mu[1]),
                          (double)__double2float_rn(mu[2]), ix, iy, iz);
#pragma unroll
      for (int j = 0; j < 3; j++) ws.stage[j][lane] = mu[j];
#pragma unroll
      for (int j = 0; j < 6; j++) ws.stage[3 + j][lane] = C[j];
    }
    __syncwarp();
    const int bx0 = (ix - k) >> 2, by0 = (iy - k) >> 2, bz0 = (iz - k) >> 2;
    const int bx1 = (ix + k) >> 2, by1 = (iy + k) >> 2, bz1 = (iz + k) >> 2;
    for (int dx = 0; dx < span; dx++) {
      const int bx = bx0 + dx;
      const bool okx = valid && bx <= bx1 && bx >= 0 && bx < nb0;
      const unsigned long long mx = expand_x(axis_bits(ix, k, bx));
      for (int dy = 0; dy < span; dy++) {
        const int by = by0 + dy;
        const bool oky = okx && by <= by1 && by >= 0 && by < nb1;
        const unsigned long long mxy = mx & expand_y(axis_bits(iy, k, by));
#if NDTB_PUSH_SCAN
        // one column of probes (all z-blocks): look up, count, one warp scan, every lane writes its own hits
        unsigned long long hitsv[MAXSPAN], bmaskv[MAXSPAN];
        int basev[MAXSPAN];
        int cnt = 0;
#pragma unroll
        for (int dz = 0; dz < MAXSPAN; dz++) {
          hitsv[dz] = 0ull, bmaskv[dz] = 0ull, basev[dz] = 0;
          const int bz = bz0 + dz;
          if (dz < span && oky && bz <= bz1 && bz >= 0 && bz < nb2) {
            if (table_find(c.table, c.tsize, (bx * nb1 + by) * nb2 + bz, basev[dz], bmaskv[dz]))
              hitsv[dz] = bmaskv[dz] & mxy & expand_z(axis_bits(iz, k, bz));
          }
          cnt += __popcll(hitsv[dz]);
        }
        int incl = cnt;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          const int t = __shfl_up_sync(FULL, incl, off);
          if (lane >= off) incl += t;
        }
        const int total = __shfl_sync(FULL, incl, 31);
        if (total == 0) continue;
        if (qcount + total > QCAP) {  // make room: drain the full batches
          __syncwarp();
          drain_batches<HESS, HA>(c, ws, qcount, 0, lane, acc, ha);
          __syncwarp();
        }
        if (qcount + total <= QCAP) {
          int pos = qcount + incl - cnt;
#pragma unroll
          for (int dz = 0; dz < MAXSPAN; dz++) {
            unsigned long long hits = hitsv[dz];
            while (hits) {
              const int b = __ffsll((long long)hits) - 1;
              hits &= hits - 1ull;
              ws.queue[pos++] = ((unsigned)lane << 27) | (unsigned)(basev[dz] + __popcll(bmaskv[dz] & ((1ull << b) - 1ull)));
            }
          }
          qcount += total;
        } else {
          // a column with more hits than the queue holds (very dense maps): one hit per lane per iteration
#pragma unroll
          for (int dz = 0; dz < MAXSPAN; dz++) {
            unsigned long long hits = hitsv[dz];
            for (;;) {
              const unsigned active = __ballot_sync(FULL, hits != 0ull);
              if (!active) break;
              if (qcount + 32 > QCAP) {
                __syncwarp();
                while (qcount >= 32) {
                  qcount -= 32;
                  process_entry<HESS, HA>(c, ws, ws.queue[qcount + lane], acc, ha);
                }
                __syncwarp();
              }
              if (hits) {
                const int b = __ffsll((long long)hits) - 1;
                hits &= hits - 1ull;
                ws.queue[qcount + __popc(active & lt)] =
                    ((unsigned)lane << 27) | (unsigned)(basev[dz] + __popcll(bmaskv[dz] & ((1ull << b) - 1ull)));
              }
              qcount += __popc(active);
            }
          }
        }
#else
        for (int dz = 0; dz < span; dz++) {
          const int bz = bz0 + dz;
          unsigned long long hits = 0, bmask = 0;
          int cbase = 0;
          if (oky && bz <= bz1 && bz >= 0 && bz < nb2) {
            if (table_find(c.table, c.tsize, (bx * nb1 + by) * nb2 + bz, cbase, bmask))
              hits = bmask & mxy & expand_z(axis_bits(iz, k, bz));
          }
          // push this probe's hits, one per lane per iteration
          for (;;) {
            const unsigned active = __ballot_sync(FULL, hits != 0ull);
            if (!active) break;
            if (qcount + 32 > QCAP) {
              __syncwarp();
              while (qcount >= 32) {
                qcount -= 32;
                process_entry<HESS, HA>(c, ws, ws.queue[qcount + lane], acc, ha);
              }
              __syncwarp();
            }
            if (hits) {
              const int b = __ffsll((long long)hits) - 1;
              hits &= hits - 1ull;
              const int slot = cbase + __popcll(bmask & ((1ull << b) - 1ull));
              ws.queue[qcount + __popc(active & lt)] = ((unsigned)lane << 27) | (unsigned)slot;
            }
            qcount += __popc(active);
          }
        }
#endif
      }
    }
    // end of round: the staged cells are about to be overwritten — drain everything
    __syncwarp();
    drain_batches<HESS, HA>(c, ws, qcount, 0, lane, acc, ha);
    if (qcount > 0) {
      if (lane < qcount) process_entry<HESS, HA>(c, ws, ws.queue[lane], acc, ha);
      qcount = 0;
    }
    __syncwarp();
    if constexpr (HA::kPairHook) {  // covariance: this lane's source cell is complete — its row enters sum g g^T
      double g[6];
#pragma unroll
      for (int a = 0; a < 6; a++) {
        g[a] = (double)(long long)ha.gs[lane * 6 + a] * (1.0 / COV_FIX);
        ha.gs[lane * 6 + a] = 0ull;
      }
      int n = 0;
#pragma unroll
      for (int a = 0; a < 6; a++)
#pragma unroll
        for (int b = a; b < 6; b++) ha.gg[(n++) * COV2_THREADS] += g[a] * g[b];
      __syncwarp();
    }
  }
}

// deterministic block reduction of n per-thread accumulators -> sums[n] (shared)
// hs != nullptr: the Hessian entries (j >= ACC_H) are read from the per-lane shared-memory slots hs[(j - ACC_H) * stride + tid]
__device__ __forceinline__ void block_reduce(const double *acc, int n, double *red, double *sums, int nwarps,
                                             const double *hs = nullptr, int stride = 0) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < ACC_TOTAL; j++) {
    if (j < n || j == ACC_PAIRS) {
      double v = (hs && j >= ACC_H && j < ACC_PAIRS) ? hs[(j - ACC_H) * stride + threadIdx.x] : acc[j];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(FULL, v, off);
      if (lane == 0) red[warp * ACC_TOTAL + j] = v;
    }
  }
  __syncthreads();
  if (threadIdx.x < ACC_TOTAL) {
    double s = 0.0;
    if (threadIdx.x < n || threadIdx.x == ACC_PAIRS)
      for (int w = 0; w < nwarps; w++) s += red[w * ACC_TOTAL + threadIdx.x];
    sums[threadIdx.x] = s;
  }
  __syncthreads();
}

// ------------------------------------------------------------------ TMA bulk copy of the block table
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void stage_table_bulk(HashEntry *dst, const HashEntry *src, int entries, uint64_t *mbar) {
  const unsigned bar = smem_u32(mbar);
  const unsigned bytes = (unsigned)entries * (unsigned)sizeof(HashEntry);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    const unsigned chunk = 16384u;
    for (unsigned off = 0; off < bytes; off += chunk) {
      const unsigned n = bytes - off < chunk ? bytes - off : chunk;
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
              smem_u32(reinterpret_cast<const char *>(dst) + off)),
          "l"(reinterpret_cast<const char *>(src) + off), "r"(n), "r"(bar)
          : "memory");
    }
  }
  // every thread waits for phase 0 of the barrier
  unsigned done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar)
        : "memory");
  }
}

// ------------------------------------------------------------------ K5: whole registration in one launch
constexpr int MAX_CLUSTER = 8;  // portable thread-block-cluster size limit

struct MatchCtl {  // what every CTA of a registration's cluster needs to start the next pass
  double P[12];    // pose of the pending evaluation: R row-major, t
  int phase;       // PH_* of the optimiser, or PH_YIELD
  int want_hess;
};
constexpr int PH_YIELD = 100;  // pass budget exhausted: state saved, the registration continues in a later launch

struct MatchShared {
  OptState st;  // rank 0 only
  OptParams prm;
  GridDesc grid;
  MatchCtl ctl;
  double sums[ACC_TOTAL];
  double csums[MAX_CLUSTER][ACC_TOTAL];  // rank 0 only: per-CTA partial sums written through DSMEM
  double red[MATCH_WARPS * ACC_TOTAL];
  uint64_t mbar;
};

extern __shared__ __align__(16) unsigned char dyn_smem[];

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// One registration per thread-block CLUSTER (G = 1..8 CTAs).  Every pass is split over the G*MATCH_WARPS warps;
// the per-CTA sums go to rank 0 through distributed shared memory, rank 0's thread 0 advances the optimiser and
// broadcasts the next pose the same way.  job_ids (optional) selects the jobs of this launch; states/pass_budget
// implement the straggler hand-over: a registration that has used pass_budget derivative passes saves its optimiser
// state and is finished by a second launch with a wider cluster.
__global__ void __launch_bounds__(MATCH_THREADS, NDTB_MATCH_MINBLOCKS)
match_kernel(const MatchJob *__restrict__ jobs, const int *__restrict__ job_ids, MatchConfig cfg, ndtb_result *__restrict__ out,
             OptState *__restrict__ states, int resume, int pass_budget, int *__restrict__ unfinished /*[0]=count, ids follow*/,
             int *__restrict__ yielded /*[n jobs] set to 1 for a registration handed to the finishing launch*/) {
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned G = cluster.num_blocks();
  const unsigned rank = cluster.block_rank();
  // dynamic smem: [table staging][MatchShared][WarpScratch x warps]
  HashEntry *stab = reinterpret_cast<HashEntry *>(dyn_smem);
  MatchShared &sh = *reinterpret_cast<MatchShared *>(dyn_smem + (size_t)cfg.table_smem_entries * sizeof(HashEntry));
  WarpScratch *wsp = reinterpret_cast<WarpScratch *>(reinterpret_cast<unsigned char *>(&sh) + ((sizeof(MatchShared) + 15) & ~(size_t)15));
  double *hs = reinterpret_cast<double *>(wsp + MATCH_WARPS);  // [21][MATCH_THREADS] Hessian sums, one slot per lane
  const SmemAcc hacc{hs + threadIdx.x};
  const int slot = blockIdx.x / G;
  const int jid = job_ids ? job_ids[slot] : slot;
  const MatchJob &job = jobs[jid];
  const int warp = threadIdx.x >> 5;
  const unsigned long long t_start = globaltimer_ns();

  PassCtx c;
  c.g = &sh.grid;
  c.tsize = job.tgt.tsize;
  c.tcell = job.tgt.gcell;
  c.scell = job.src_gcell;
  c.ns = job.src_ng;
  c.k = cfg.n_neighbours;
  c.lfd1 = cfg.lfd1, c.lfd2 = cfg.lfd2;
  const bool staged = job.tgt.tsize <= cfg.table_smem_entries;
  c.table = staged ? stab : job.tgt.table;

  if (G > 1) cluster.sync();  // every CTA of the cluster has started: its shared memory may be written remotely
  if (threadIdx.x == 0) {
    sh.grid = job.tgt.g;
    if (rank == 0) {
      OptParams &p = sh.prm;
      p.itr_max = cfg.itr_max, p.step_control = cfg.step_control, p.regularize = cfg.regularize, p.planar = cfg.planar;
      p.fusion = job.fusion, p.soft = job.fusion && cfg.soft, p.tik = job.fusion && cfg.tik;
      p.delta_score = cfg.delta_score;
      for (int i = 0; i < 36; i++) p.Q[i] = job.Q[i];
      if (resume)
        sh.st = states[jid];
      else
        opt_begin(sh.st, p, job.T0);
      MatchCtl ctl;
      for (int i = 0; i < 9; i++) ctl.P[i] = sh.st.Peval.R[i];
      for (int i = 0; i < 3; i++) ctl.P[9 + i] = sh.st.Peval.t[i];
      ctl.phase = sh.st.phase, ctl.want_hess = sh.st.want_hess;
      for (unsigned r = 0; r < G; r++) *cluster.map_shared_rank(&sh.ctl, r) = ctl;
    }
  }
  if (staged)
    stage_table_bulk(stab, job.tgt.table, job.tgt.tsize, &sh.mbar);
  if (G > 1)
    cluster.sync();
  else
    __syncthreads();

  int passes = 0;
  for (;;) {
    const int phase = sh.ctl.phase;  // uniform over the cluster: written before the last barrier
    if (phase == PH_DONE || phase == PH_YIELD) break;
    const bool hess = sh.ctl.want_hess != 0;
    double acc[ACC_TOTAL];
#pragma unroll
    for (int j = 0; j < ACC_TOTAL; j++) acc[j] = 0.0;
    if (hess) {
#if NDTB_HESS_SMEM
#pragma unroll
      for (int j = 0; j < 21; j++) hs[j * MATCH_THREADS + threadIdx.x] = 0.0;
      d2d_pass<true, SmemAcc>(c, sh.ctl.P, wsp[warp], rank * MATCH_WARPS + warp, G * MATCH_WARPS, acc, hacc);
      block_reduce(acc, 28, sh.red, sh.sums, MATCH_WARPS, hs, MATCH_THREADS);
#else
      d2d_pass<true, RegAcc>(c, sh.ctl.P, wsp[warp], rank * MATCH_WARPS + warp, G * MATCH_WARPS, acc, RegAcc{acc});
      block_reduce(acc, 28, sh.red, sh.sums, MATCH_WARPS);
#endif
    } else {
      d2d_pass<false, SmemAcc>(c, sh.ctl.P, wsp[warp], rank * MATCH_WARPS + warp, G * MATCH_WARPS, acc, hacc);
      block_reduce(acc, 7, sh.red, sh.sums, MATCH_WARPS);
    }
    passes++;
    if (G > 1) {
      if (threadIdx.x < ACC_TOTAL) cluster.map_shared_rank(&sh.csums[0][0], 0)[rank * ACC_TOTAL + threadIdx.x] = sh.sums[threadIdx.x];
      cluster.sync();
    }
    if (rank == 0 && threadIdx.x == 0) {
      if (G > 1)
        for (int j = 0; j < ACC_TOTAL; j++) {
          double v = 0.0;
          for (unsigned r = 0; r < G; r++) v += sh.csums[r][j];  // fixed order: deterministic
          sh.sums[j] = v;
        }
      opt_advance(sh.st, sh.prm, sh.sums);
      MatchCtl ctl;
      for (int i = 0; i < 9; i++) ctl.P[i] = sh.st.Peval.R[i];
      for (int i = 0; i < 3; i++) ctl.P[9 + i] = sh.st.Peval.t[i];
      ctl.phase = sh.st.phase, ctl.want_hess = sh.st.want_hess;
      if (ctl.phase != PH_DONE && pass_budget > 0 && passes >= pass_budget) ctl.phase = PH_YIELD;
      for (unsigned r = 0; r < G; r++) *cluster.map_shared_rank(&sh.ctl, r) = ctl;
    }
    if (G > 1)
      cluster.sync();
    else
      __syncthreads();
  }

  if (rank == 0 && threadIdx.x == 0) {
    const OptState &s = sh.st;
    const float ms = (float)((double)(globaltimer_ns() - t_start) * 1e-6);
    if (sh.ctl.phase == PH_YIELD) {
      states[jid] = s;
      out[jid].kernel_ms = ms;  // carried over to the finishing launch
      const int at = atomicAdd(unfinished, 1);
      unfinished[1 + at] = jid;
      if (yielded) yielded[jid] = 1;
    } else {
      ndtb_result r;
      pose_to_cm(s.T, r.T);
      r.score = s.score_here, r.score_best = s.score_best;
      r.converged = s.ret, r.iterations = s.itr, r.n_hess_passes = s.n_hess, r.n_grad_passes = s.n_grad;
      r.exit_code = s.exit_code;
      int changed = 0;
      for (int i = 0; i < 16; i++) changed |= (r.T[i] != job.T0[i]);
      r.pose_changed = changed;
      r.status = (s.ret ? NDTB_ST_CONVERGED : NDTB_ST_ITR_MAX) | (changed ? NDTB_ST_POSE_CHANGED : 0) |
                 (s.nonfinite ? NDTB_ST_NONFINITE : 0) | ((job.src_ng == 0 || job.tgt.ng == 0) ? NDTB_ST_NO_CELLS : 0);
      r.n_src_cells = job.src_ng, r.n_tgt_cells = job.tgt.ng, r.tgt_table_entries = job.tgt.tsize;
      r.kernel_ms = ms * (float)G + (resume ? out[jid].kernel_ms : 0.f);  // SM-milliseconds spent on this registration
      r.n_exec_passes = s.n_exec;
      out[jid] = r;
    }
  }
}

size_t match_smem_bytes(int table_entries) {
  return (size_t)table_entries * sizeof(HashEntry) + ((sizeof(MatchShared) + 15) & ~(size_t)15) +
         MATCH_WARPS * sizeof(WarpScratch) + (NDTB_HESS_SMEM ? (size_t)21 * MATCH_THREADS * sizeof(double) : 0);
}
size_t opt_state_bytes() { return sizeof(OptState); }

// The dynamic shared memory limit is an attribute of the FUNCTION on a device, shared by every context / host thread that
// launches it: it is raised once per context creation to the device's opt-in maximum and never lowered per launch (two
// threads with different table sizes would otherwise race on it).
size_t cov_smem_bytes();
__global__ void cov_pass_kernel(const MatchJob *, MatchConfig, const ndtb_result *, const long long *, double *, double *, const int *, int);
cudaError_t match_kernel_prepare(int smem_optin_bytes) {
  cudaError_t e = cudaFuncSetAttribute(match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin_bytes);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(cov_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cov_smem_bytes());
}

// n_slots registrations (job_ids[slot] or slot itself), each on a cluster of `cluster` CTAs
cudaError_t launch_match(const MatchJob *d_jobs, const int *d_job_ids, int n_slots, int cluster, const MatchConfig &cfg,
                         ndtb_result *d_out, void *d_states, int resume, int pass_budget, int *d_unfinished, int *d_yielded,
                         cudaStream_t stream) {
  const size_t smem = match_smem_bytes(cfg.table_smem_entries);
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3((unsigned)(n_slots * cluster));
  lc.blockDim = dim3(MATCH_THREADS);
  lc.dynamicSmemBytes = smem;
  lc.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)cluster, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
  lc.attrs = at, lc.numAttrs = 1;
  return cudaLaunchKernelEx(&lc, match_kernel, d_jobs, d_job_ids, cfg, d_out, (OptState *)d_states, resume, pass_budget,
                            d_unfinished, d_yielded);
}

// ------------------------------------------------------------------ stand-alone derivativesNDT (API + tests)
constexpr int DERIV_WARPS = 8;
struct DerivShared {
  GridDesc grid;
  double P[12];
  double sums[ACC_TOTAL];
  double red[DERIV_WARPS * ACC_TOTAL];
};

template <bool HESS>
__global__ void __launch_bounds__(DERIV_WARPS * 32, 1)
deriv_kernel(const MatchJob *__restrict__ job_p, MatchConfig cfg, double *__restrict__ partial /*[grid][ACC_TOTAL]*/) {
  __shared__ DerivShared sh;
  __shared__ WarpScratch ws[DERIV_WARPS];
  const MatchJob &job = *job_p;
  if (threadIdx.x == 0) {
    sh.grid = job.tgt.g;
    const Pose T = pose_from_cm(job.T0);
    for (int i = 0; i < 9; i++) sh.P[i] = T.R[i];
    for (int i = 0; i < 3; i++) sh.P[9 + i] = T.t[i];
  }
  __syncthreads();
  PassCtx c;
  c.g = &sh.grid;
  c.table = job.tgt.table, c.tsize = job.tgt.tsize, c.tcell = job.tgt.gcell;
  c.scell = job.src_gcell, c.ns = job.src_ng;
  c.k = cfg.n_neighbours, c.lfd1 = cfg.lfd1, c.lfd2 = cfg.lfd2;
  double acc[ACC_TOTAL];
#pragma unroll
  for (int j = 0; j < ACC_TOTAL; j++) acc[j] = 0.0;
  const int warp = threadIdx.x >> 5;
  d2d_pass<HESS, RegAcc>(c, sh.P, ws[warp], blockIdx.x * DERIV_WARPS + warp, gridDim.x * DERIV_WARPS, acc, RegAcc{acc});
  block_reduce(acc, HESS ? 28 : 7, sh.red, sh.sums, DERIV_WARPS);
  if (threadIdx.x < ACC_TOTAL) partial[blockIdx.x * ACC_TOTAL + threadIdx.x] = sh.sums[threadIdx.x];
}

__global__ void sum_partials_kernel(const double *__restrict__ partial, int n_parts, int width, double *__restrict__ out) {
  const int j = threadIdx.x;
  if (j >= width) return;
  double s = 0.0;
  for (int p = 0; p < n_parts; p++) s += partial[p * width + j];
  out[j] = s;
}

cudaError_t launch_derivatives(const MatchJob *d_job, const MatchConfig &cfg, bool hess, int n_ctas, double *d_partial,
                               double *d_out29, cudaStream_t stream) {
  if (hess)
    deriv_kernel<true><<<n_ctas, DERIV_WARPS * 32, 0, stream>>>(d_job, cfg, d_partial);
  else
    deriv_kernel<false><<<n_ctas, DERIV_WARPS * 32, 0, stream>>>(d_job, cfg, d_partial);
  sum_partials_kernel<<<1, 32, 0, stream>>>(d_partial, n_ctas, ACC_TOTAL, d_out29);
  return cudaGetLastError();
}

// ------------------------------------------------------------------ covariance
// Definition (fixed by the CPU restatement): H = Hessian at T; rows = per-source-cell gradient
// sums g_i and per-target-cell gradient sums g_t; cov = H^-1 (sigma^2 sum_c g_c g_c^T) H^-1, sigma = 0.03.
constexpr int COV_THREADS = 128;      // cov_finalize_kernel
constexpr int COV_W = 28 + 21;  // H pass sums + upper triangle of sum g g^T

__device__ __forceinline__ void outer_upper(const double *g, double *o) {
  int n = 0;
#pragma unroll
  for (int a = 0; a < 6; a++)
#pragma unroll
    for (int b = a; b < 6; b++) o[n++] += g[a] * g[b];
}

// grid (jobs, chunks): the pair-per-lane machinery of the registration kernel (d2d_pass) with the CovAcc policy.
// dynamic smem: [WarpScratch x warps][gs rows x warps][21 x COV2_THREADS Hessian slots]
struct CovShared {
  GridDesc grid;
  double P[12];
  double red[(COV2_THREADS / 32) * 49];
};
__global__ void __launch_bounds__(COV2_THREADS, 1)
cov_pass_kernel(const MatchJob *__restrict__ jobs, MatchConfig cfg, const ndtb_result *__restrict__ res,
                const long long *__restrict__ gt_off, double *__restrict__ gt, double *__restrict__ partial,
                const int *__restrict__ yielded, int mode /*0 all, 1 not yielded, 2 yielded only*/) {
  if (mode && (yielded[blockIdx.x] != 0) != (mode == 2)) return;
  constexpr int NW = COV2_THREADS / 32;
  const MatchJob &job = jobs[blockIdx.x];
  __shared__ CovShared sh;
  WarpScratch *wsp = reinterpret_cast<WarpScratch *>(dyn_smem);
  unsigned long long *gs_all = reinterpret_cast<unsigned long long *>(wsp + NW);
  double *hs = reinterpret_cast<double *>(gs_all + NW * 32 * 6);
  const bool skip = res && !res[blockIdx.x].pose_changed;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    sh.grid = job.tgt.g;
    const Pose T = pose_from_cm(res ? res[blockIdx.x].T : job.T0);
    for (int i = 0; i < 9; i++) sh.P[i] = T.R[i];
    for (int i = 0; i < 3; i++) sh.P[9 + i] = T.t[i];
  }
#pragma unroll
  for (int j = 0; j < 21; j++) hs[j * COV2_THREADS + threadIdx.x] = 0.0;
  for (int a = 0; a < 6; a++) gs_all[(warp * 32 + lane) * 6 + a] = 0ull;
  __syncthreads();
  double acc[ACC_TOTAL];  // [0] score, [1..6] g, [7..27] H, [28] pairs; sum g g^T over the source rows: shared slots hs
#pragma unroll
  for (int j = 0; j < ACC_TOTAL; j++) acc[j] = 0.0;
  if (!skip) {
    PassCtx c;
    c.g = &sh.grid;
    c.table = job.tgt.table, c.tsize = job.tgt.tsize, c.tcell = job.tgt.gcell;
    c.scell = job.src_gcell, c.ns = job.src_ng;
    c.k = cfg.n_neighbours, c.lfd1 = cfg.lfd1, c.lfd2 = cfg.lfd2;
    const CovAcc ha{acc, hs + threadIdx.x, gs_all + warp * 32 * 6, reinterpret_cast<unsigned long long *>(gt + gt_off[blockIdx.x] * 6)};
    d2d_pass<true, CovAcc>(c, sh.P, wsp[warp], blockIdx.y * NW + warp, gridDim.y * NW, acc, ha);
  }
  // block reduce the 49 values (fixed tree) -> partial[job][chunk][COV_W]: [0..6] score+g, [7..27] H, [28..48] sum g g^T
#pragma unroll
  for (int j = 0; j < 49; j++) {
    double v = j < 28 ? acc[j] : hs[(j - 28) * COV2_THREADS + threadIdx.x];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(FULL, v, off);
    if (lane == 0) sh.red[warp * 49 + j] = v;
  }
  __syncthreads();
  if (threadIdx.x < 49) {
    double s = 0.0;
    for (int w = 0; w < NW; w++) s += sh.red[w * 49 + threadIdx.x];
    partial[((size_t)blockIdx.x * gridDim.y + blockIdx.y) * 49 + threadIdx.x] = s;
  }
}

size_t cov_smem_bytes() {
  return (COV2_THREADS / 32) * sizeof(WarpScratch) + (size_t)(COV2_THREADS / 32) * 32 * 6 * 8 + (size_t)21 * COV2_THREADS * 8;
}

// one CTA per job: ordered sum of the partials, sum over target rows, 6x6 sandwich
__global__ void __launch_bounds__(COV_THREADS)
cov_finalize_kernel(const MatchJob *__restrict__ jobs, const ndtb_result *__restrict__ res, const long long *__restrict__ gt_off,
                    const double *__restrict__ gt, const double *__restrict__ partial, int n_chunks,
                    double *__restrict__ cov36, int *__restrict__ status, const int *__restrict__ yielded, int mode) {
  if (mode && (yielded[blockIdx.x] != 0) != (mode == 2)) return;
  const MatchJob &job = jobs[blockIdx.x];
  __shared__ double tot[COV_W];
  __shared__ double red[(COV_THREADS / 32) * 21];
  double *out = cov36 + (size_t)blockIdx.x * 36;
  if (res && !res[blockIdx.x].pose_changed) {  // ndt_feature_graph.cpp:300-310
    if (threadIdx.x < 36) out[threadIdx.x] = (threadIdx.x % 7 == 0) ? 0.02 : 0.0;
    if (threadIdx.x == 0 && status) status[blockIdx.x] = 0;
    return;
  }
  if (threadIdx.x < COV_W) {
    double s = 0.0;
    for (int p = 0; p < n_chunks; p++) s += partial[((size_t)blockIdx.x * n_chunks + p) * COV_W + threadIdx.x];
    tot[threadIdx.x] = s;
  }
  double o[21];
#pragma unroll
  for (int j = 0; j < 21; j++) o[j] = 0.0;
  const double *gtj = gt + gt_off[blockIdx.x] * 6;
  for (int t = threadIdx.x; t < job.tgt.ng; t += COV_THREADS) {
    double g[6];
#pragma unroll
    for (int a = 0; a < 6; a++) g[a] = (double)reinterpret_cast<const long long *>(gtj)[(size_t)t * 6 + a] * (1.0 / COV_FIX);
    outer_upper(g, o);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 21; j++) {
    double v = o[j];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(FULL, v, off);
    if (lane == 0) red[warp * 21 + j] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double H[36], JtJ[36], Hinv[36], tmp[36];
    int n = 0;
    for (int a = 0; a < 6; a++)
      for (int b = a; b < 6; b++, n++) {
        H[a * 6 + b] = H[b * 6 + a] = tot[ACC_H + n];
        double s = tot[28 + n];
        for (int w = 0; w < COV_THREADS / 32; w++) s += red[w * 21 + n];
        JtJ[a * 6 + b] = JtJ[b * 6 + a] = s;
      }
    const bool ok = inv6(H, Hinv);
    const double sigmaS = 0.03 * 0.03;
    for (int i = 0; i < 6; i++)
      for (int j = 0; j < 6; j++) {
        double s = 0;
        for (int q = 0; q < 6; q++) s += Hinv[i * 6 + q] * sigmaS * JtJ[q * 6 + j];
        tmp[i * 6 + j] = s;
      }
    for (int i = 0; i < 6; i++)
      for (int j = 0; j < 6; j++) {
        double s = 0;
        for (int q = 0; q < 6; q++) s += tmp[i * 6 + q] * Hinv[q * 6 + j];
        out[i * 6 + j] = ok ? s : 0.0;
      }
    if (status) status[blockIdx.x] = ok ? 0 : NDTB_ERR_SINGULAR;
  }
}

cudaError_t launch_covariance(const MatchJob *d_jobs, int n_jobs, const MatchConfig &cfg, const ndtb_result *d_res,
                              const long long *d_gt_off, double *d_gt, double *d_partial, int n_chunks,
                              double *d_cov36, int *d_status, const int *d_yielded, int mode, cudaStream_t stream) {
  dim3 grid(n_jobs, n_chunks);  // jobs on x: no 65535 limit on the batch size
  cov_pass_kernel<<<grid, COV2_THREADS, cov_smem_bytes(), stream>>>(d_jobs, cfg, d_res, d_gt_off, d_gt, d_partial, d_yielded, mode);
  cov_finalize_kernel<<<n_jobs, COV_THREADS, 0, stream>>>(d_jobs, d_res, d_gt_off, d_gt, d_partial, n_chunks, d_cov36,
                                                         d_status, d_yielded, mode);
  return cudaGetLastError();
}

int cov_partial_width() { return COV_W; }
int acc_total() { return ACC_TOTAL; }

}  // namespace ndtb
