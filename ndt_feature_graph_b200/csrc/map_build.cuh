// map_build.cuh — job descriptor and launchers of the map-build (K1) kernels; see map_build.cu.
#pragma once

#include "../../include/ndtb.h"
#include "engine.cuh"

namespace ndtb {

constexpr int MAX_TRACE_SEG = 4;  // addPointCloud calls (each with its own origin) between two computeNDTCells calls
struct TraceSeg {
  double origin[3];
  double maxz, sensor_noise;
  float occ_limit;
  int begin, end;  // range of pending points traced from this origin
};

// One map of a batched build.  Pointers address this map's slice of the batch slabs.
struct BuildJob {
  GridDesc g;
  // Storage box of the dense per-block arrays (amask / abase), in 4x4x4 blocks: boff = first stored block per axis, nbs =
  // stored blocks per axis.  Normally the whole grid (boff = 0, nbs = g.nb).  A guess-size grid spans 4 x maxDist per axis
  // although its points only occupy the central +-maxDist: the temporaries of ndtb_register_scans store that box only
  // (1/12 of the blocks at C2: 12x less to clear and to scan).  Cell order ((bx, by, bz) lexicographic, then bit) and the
  // block ids of the matcher's hash table are those of the full grid either way.
  int boff[3], nbs[3];
  const float4 *pts;   // pending points (pcl::PointXYZ layout), device memory
  int npts;
  double range_limit;  // loadPointCloud range filter; <= 0: none
  double range_origin[3];  // the filter's reference point ((0,0,0) for loadPointCloud, the sensor origin for loadPointCloudCentroid)
  // all-cells structure (new layout)
  unsigned long long *amask;  // [nblk] touched voxels per 4x4x4 block
  int *abase;                 // [nblk] exclusive popcount scan
  int nblk;
  int *tb_list;  // compact list of touched blocks (ascending)
  int *counts;   // [8]: 0 n_all, 1 n touched blocks, 2 n gaussian cells, 3 n gaussian blocks, 4 points binned
  // per-point temporaries
  int *pt_cell;  // [npts] voxel key (block*64+bit) then cell id; -1 = dropped
  int *seg_idx;  // [npts] radix-sort ping-pong buffer (point ids)
  int *seg2;     // [npts] point ids grouped by cell, ascending inside each cell (= insertion order of NDTCell::points_)
  const int *sorted_ids;  // seg2 or seg_idx: where the last sort pass left the ids (depends on the number of passes)
  int *key2;     // [npts] radix-sort ping-pong buffer (cell ids)
  int *rs_hist;  // [256 * tiles] digit histograms of the sort tiles, digit-major
  // per-cell (valid after the popcount scan)
  int n_all;
  int *cnt, *seg_off, *cursor;
  int *cell_key;  // [n_all] block*64+bit of every cell
  double *cmean;  // [n_all][3]
  double *ccov;   // [n_all][9]
  int *cn, *chas;
  float *cocc;
  // previous content of the map (merge path of computeNDTCells); null when fresh
  const unsigned long long *o_amask;
  const int *o_abase;
  const double *o_cmean, *o_ccov;
  const int *o_cn, *o_chas;
  const float *o_cocc;
  // Gaussian view
  double *gcell;  // [ng][9]
  int *g2c;       // [ng] -> all-cells index
  HashEntry *table;
  int tsize;
  unsigned long long *gmask_t;  // per touched block
  int *gbase_t;
  unsigned maxnumpoints;
  float occ_limit;
  double log_occ;  // log(0.6/0.4) evaluated on the host
  // ---- free-space ray trace of NDTMap::addPointCloud (LazyGrid::traceLine): the pending points [begin, end) of a
  // segment are rays from that segment's origin.  n_seg == 0: plain end-point binning.
  int n_seg;
  TraceSeg seg[MAX_TRACE_SEG];
  int *vis_cnt;   // [npts] cells met by ray i (deduplicated, inside the grid)
  int *vis_off;   // [npts] exclusive scan of vis_cnt; total in counts[7]
  int *vis_key;   // [n_vis] voxel key (block*64+bit) of visit v, then the cell id; visits are numbered ray-major
  int *vis_ray;   // [n_vis] ray (= point index) of visit v
  int *v_cnt, *v_seg_off;  // [n_all] visits per cell, segment offsets
  int *v_seg2;    // [n_vis] visit ids grouped by cell, ascending inside a cell (= the order the rays were traced in)
};

int centroid_chunk_points();
int centroid_record_doubles();
// d_rec_off[w] = first chunk record of map w in d_recs (records of centroid_record_doubles() doubles, one per
// centroid_chunk_points() points)
int launch_guess(const BuildJob *d_jobs, const int *d_which, int n_which, int max_pts, const long long *d_rec_off, double *d_recs,
                 double *d_out, cudaStream_t s);
int launch_mark(const BuildJob *d_jobs, int n, int max_pts, bool trace, bool fast, cudaStream_t s);
// visit lists of the ray trace: d_vjobs = the jobs with the per-point arrays replaced by the per-visit arrays
int launch_trace_lists(const BuildJob *d_jobs, const BuildJob *d_vjobs, int n, int max_pts, int max_vis, int max_ntb, int max_cells,
                       cudaStream_t s);
int launch_transform_points(const float4 *d_in, float4 *d_out, int n, const float *d_T12, cudaStream_t s);
int launch_cells(const BuildJob *d_jobs, int n, int max_pts, int max_ntb, int max_cells, cudaStream_t s);
int sort_passes(int max_cells);  // 8-bit radix passes needed for cell ids 0..max_cells
int sort_tile_points();  // points per radix-sort tile (sizing of BuildJob::rs_hist: 256 ints per tile)
int launch_gview(const BuildJob *d_jobs, int n, int max_ntb, int max_cells, cudaStream_t s);
int launch_blockscan(const BuildJob *d_jobs, int n, cudaStream_t s);
int launch_export(const BuildJob *d_job, int ntb, ndtb_cell *d_out, cudaStream_t s);
int launch_from_cells_voxel(const BuildJob *d_job, const ndtb_cell *d_cells, int n, int use_idx, int *d_vox, int *d_err,
                            cudaStream_t s);
int launch_from_cells_place(const BuildJob *d_job, const ndtb_cell *d_cells, int n, const int *d_vox, cudaStream_t s);
int launch_point_indices(const GridDesc &g, const float4 *d_pts, int n, int *d_out, int *d_nin, cudaStream_t s);
int launch_points_as_cells(const float4 *d_pts, int n, double *d_gcell, cudaStream_t s);
int launch_overlap(const BuildJob *d_jobs2, int n_links, const double *d_T16, int T_stride, double *d_out, cudaStream_t s);

}  // namespace ndtb
