// engine.cuh — device-side data layout shared by the map-build (K1) and D2D (K2/K5) kernels.
//
// HBM layout of one NDT map (lslgeneric::NDTMap on a LazyGrid [upstream]):
//   * the voxel grid is tiled in 4x4x4 BLOCKS; bit (lx*16 + ly*4 + lz) of a block's 64-bit mask is one voxel.
//   * "all cells"  (every voxel that ever received a point): dense per-block arrays amask[nblk] (u64) and
//     abase[nblk] (i32, exclusive popcount scan) + SoA cell records indexed abase[b] + rank(bit).
//   * "Gaussian view" used by the matcher: compact AoS gcell[ng][9] = mean(3) + cov(xx,xy,xz,yy,yz,zz),
//     ordered by (block id, bit), and an open-addressing hash table block id -> {base, mask} (16-byte
//     entries) that the D2D kernel stages in shared memory with one bulk (TMA) copy.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "optimizer.h"

namespace ndtb {

struct GridDesc {
  double center[3];
  double cell[3];
  int size[3];  // voxels per axis
  int nb[3];    // 4x4x4 blocks per axis = (size+3)/4
};

struct __align__(16) HashEntry {
  int key;   // block linear id (bx*nb1 + by)*nb2 + bz, -1 = empty
  int base;  // index of the block's first Gaussian cell in gcell
  unsigned long long mask;  // Gaussian voxels of the block
};

struct MapView {
  GridDesc g;
  const double *gcell;     // [ng][9]
  const HashEntry *table;  // [tsize], tsize power of two (>= 2)
  int ng;
  int tsize;
};

constexpr int GC = 9;  // doubles per Gaussian cell record

__host__ __device__ __forceinline__ unsigned hash_block(int key, int tsize) {
  return ((unsigned)key * 2654435761u >> 7) & (unsigned)(tsize - 1);
}

#ifdef __CUDACC__
// LazyGrid::getIndexForPoint [upstream]: ind = floor((p - center)/cell + 0.5) + size/2.0 truncated to int.
// Written with explicit round-to-nearest intrinsics so no FMA contraction can change a voxel index
// relative to the CPU restatement (exact-index parity, SURVEY.md §7 hard part c).
// A cell size that is a power of two (0.5 m in every shipped configuration) divides exactly: x / cell == x * (1/cell)
// bit for bit, and the multiplication is one instruction where IEEE division is ~30.
__device__ __forceinline__ bool voxel_axis(double p, double c, double cell, int size, int &out) {
  const double d = __dsub_rn(p, c);
  const bool pow2 = (__double_as_longlong(cell) & 0x000fffffffffffffll) == 0ll && cell > 1e-300 && cell < 1e300;
  const double qd = pow2 ? __dmul_rn(d, __drcp_rn(cell)) : __ddiv_rn(d, cell);
  const double v = __dadd_rn(floor(__dadd_rn(qd, 0.5)), (double)size * 0.5);
  if (!(v > -2147483000.0 && v < 2147483000.0)) return false;  // also rejects NaN
  out = __double2int_rz(v);
  return true;
}
__device__ __forceinline__ bool voxel_index(const GridDesc &g, double px, double py, double pz, int &ix, int &iy,
                                            int &iz) {
  return voxel_axis(px, g.center[0], g.cell[0], g.size[0], ix) &
         voxel_axis(py, g.center[1], g.cell[1], g.size[1], iy) &
         voxel_axis(pz, g.center[2], g.cell[2], g.size[2], iz);
}
__device__ __forceinline__ bool in_grid(const GridDesc &g, int ix, int iy, int iz) {
  return ix >= 0 && iy >= 0 && iz >= 0 && ix < g.size[0] && iy < g.size[1] && iz < g.size[2];
}
__device__ __forceinline__ int block_id(const GridDesc &g, int ix, int iy, int iz) {
  return ((ix >> 2) * g.nb[1] + (iy >> 2)) * g.nb[2] + (iz >> 2);
}
__device__ __forceinline__ int block_bit(int ix, int iy, int iz) { return ((ix & 3) << 4) | ((iy & 3) << 2) | (iz & 3); }
#endif

// one registration handed to the match kernel
struct MatchJob {
  MapView tgt;
  const double *src_gcell;
  int src_ng;
  int fusion;
  double T0[16];   // column-major
  double Q[36];    // Tcov^-1 when fusion
};

struct MatchConfig {  // uniform over a batch
  int n_neighbours;
  int itr_max, step_control, regularize, soft, tik, planar;
  double delta_score, lfd1, lfd2;
  int table_smem_entries;  // capacity of the shared-memory staging area (hash entries); 0 = probe in global
};

}  // namespace ndtb
