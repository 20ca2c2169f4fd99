/*
 * ndtb.h — C ABI of the B200-native NDT registration engine (libndtb.so).
 *
 * Drop-in boundary for the NDT hot path of MalcolmMielle/ndt_feature_graph.  Every entry point
 * names the reference interface it replaces (paths relative to the reference checkout; "[upstream]"
 * = perception_oru symbol the reference calls but does not vendor):
 *
 *   ndtb_map_*           lslgeneric::NDTMap(new LazyGrid(res)) + loadPointCloud/addPointCloud/
 *                        computeNDTCells [upstream]; call sites
 *                        ndt_feature/src/ndt_feature_src/ndt_feature_fuser_hmt.cpp:87-94,195-227,485-486
 *   ndtb_d2d_derivatives NDTMatcherD2D::derivativesNDT [upstream]; call sites
 *                        ndt_feature/include/ndt_feature/ndt_matcher_d2d_fusion.h:80,238,444,617,856,1085
 *   ndtb_d2d_match       NDTMatcherD2D::match(NDTMap&,NDTMap&,Affine3d&,bool) [upstream]; call site
 *                        ndt_feature/src/ndt_feature_src/ndt_feature_graph.cpp:273
 *   ndtb_fusion_match    ndt_feature::matchFusion, ndt_matcher_d2d_fusion.h:797-1155 (useNDT only)
 *   ndtb_d2d_covariance  NDTMatcherD2D::covariance [upstream]; call sites ndt_feature_graph.cpp:298,
 *                        ndt_feature_fuser_hmt.cpp:405
 *   ndtb_d2d_match_batch NDTFeatureGraph::updateLinksUsingNDTRegistration, ndt_feature_graph.cpp:347-353
 *                        (the serial loop over links becomes one batched launch)
 *   ndtb_register_scans  NDTFeatureFuserHMT::update front-end step, ndt_feature_fuser_hmt.cpp:195-227
 *                        (local map of the scan) + :356-357 (match) + :399-420 (covariance), batched
 *   ndtb_p2d_match       NDTMatcherP2D::match [upstream]; no call site in the reference (BASELINE.json config C3)
 *   ndtb_overlap_score   NDTFeatureNode::overlapNDTOccupancyScore, ndt_feature/include/ndt_feature/ndt_feature_node.h:213-252
 *
 * Conventions
 *   - plain C types only; poses are 16 doubles, column-major 4x4 (Eigen::Affine3d::matrix().data()).
 *   - every function returns an int: 0 = NDTB_OK, <0 = error (ndtb_strerror).  Nothing throws across
 *     the ABI, nothing blocks on stdin (cf. ndt_feature_graph.cpp:318-328), nothing prints.
 *   - a ndtb_ctx is single-owner (one per host thread / per GPU); calls are synchronous at return
 *     unless every output lives in device memory (NDTB_MEM_DEVICE), in which case work is only
 *     enqueued on the context's stream.  Maps are used with the context that created them: a matcher /
 *     overlap call with a map of another context returns NDTB_ERR_ARG (map storage is recycled in the
 *     order of its own context's stream).
 *   - there is NO CPU implementation behind this ABI: without a CUDA device every compute entry
 *     point fails with NDTB_ERR_CUDA.
 */
#ifndef NDTB_H
#define NDTB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NDTB_VERSION 200

enum {
  NDTB_OK = 0,
  NDTB_ERR_CUDA = -1,      /* CUDA runtime error / no device */
  NDTB_ERR_ARG = -2,       /* bad argument */
  NDTB_ERR_GRID = -3,      /* grid undefined or too large (> 2^31 voxels) */
  NDTB_ERR_EMPTY = -4,     /* no usable points / no Gaussian cells */
  NDTB_ERR_SINGULAR = -5,  /* 6x6 system not invertible (covariance, Tcov) */
  NDTB_ERR_NOMEM = -6
};

enum { NDTB_MEM_HOST = 0, NDTB_MEM_DEVICE = 1 };

typedef struct ndtb_ctx ndtb_ctx;
typedef struct ndtb_map ndtb_map; /* lslgeneric::NDTMap on a LazyGrid, resident in HBM */

typedef struct ndtb_grid { /* LazyGrid geometry */
  double center[3];        /* centerX/Y/Z */
  double cell[3];          /* cellSizeX/Y/Z */
  int32_t size[3];         /* sizeX/Y/Z in cells = |ceil(size_m / cell)| */
} ndtb_grid;

typedef struct ndtb_cell { /* NDTCell snapshot (same layout as the oracle's orc_cell) */
  double mean[3];
  double cov[6];           /* xx, xy, xz, yy, yz, zz */
  int32_t n;               /* NDTCell::N */
  int32_t has_gaussian;    /* NDTCell::hasGaussian_ */
  int32_t idx[3];          /* voxel index in the LazyGrid */
  float occ;               /* log-odds occupancy */
} ndtb_cell;

typedef struct ndtb_params { /* NDTMatcherD2D public knobs + matchFusion flags */
  int32_t n_neighbours;    /* NDTMatcherD2D::n_neighbours (2) */
  int32_t itr_max;         /* ITR_MAX (30) */
  int32_t step_control;    /* 1 */
  int32_t regularize;      /* 1 */
  double delta_score;      /* DELTA_SCORE (1e-3 for a default-constructed matcher) */
  double lfd1, lfd2;       /* 1, 0.05 */
  int32_t use_soft_constraints; /* matchFusion only */
  int32_t use_tikhonov;         /* matchFusion only */
  int32_t ctas_per_match;  /* engine knob: CTAs (one thread-block cluster, 1/2/4/8) cooperating on one
                              registration; 0 = auto (1 for batches >= #SMs, up to 8 for small batches) */
  int32_t pass_budget;     /* engine knob: derivative passes a registration may use in the first (1-CTA) launch before
                              it is handed to a second launch on 8-CTA clusters (stragglers); 0 = auto, <0 = never */
  int32_t planar;          /* 1 = NDTMatcherD2D_2D [upstream] (matchFusion2d, ndt_matcher_d2d_fusion.h:1159-1176):
                              the same loop estimating (x, y, yaw) only */
  int32_t reserved_;
} ndtb_params;

/* ndtb_result.status bits (SURVEY.md §5 "failure detection") */
enum {
  NDTB_ST_CONVERGED = 1,   /* match() returned true */
  NDTB_ST_ITR_MAX = 2,     /* ITR_MAX exceeded (match() returned false) */
  NDTB_ST_POSE_CHANGED = 4,/* T differs bitwise from the initial guess (ndt_feature_graph.cpp:286-294) */
  NDTB_ST_NONFINITE = 8,   /* a non-finite score/gradient was met */
  NDTB_ST_NO_CELLS = 16    /* source or target has no Gaussian cells */
};

typedef struct ndtb_result {
  double T[16];            /* column-major 4x4 */
  double score;            /* final score_here */
  double score_best;
  int32_t converged;       /* return value of match() */
  int32_t iterations;
  int32_t n_hess_passes;   /* derivativesNDT calls with Hessian */
  int32_t n_grad_passes;   /* gradient-only derivativesNDT calls the reference makes */
  int32_t pose_changed;
  int32_t exit_code;       /* 0 loop end, 1 gradient vanished, 2 dginit>0, 3 itr_max, 4 no regularisation */
  int32_t status;          /* NDTB_ST_* */
  int32_t n_src_cells;     /* Gaussian cells of the source map (Cs) */
  int32_t n_tgt_cells;     /* Gaussian cells of the target map (Ct) */
  int32_t tgt_table_entries; /* 16-byte entries of the target's block table (probe window W) */
  float kernel_ms;         /* SM-milliseconds the registration kernel spent on this edge (CTAs x device time) */
  int32_t n_exec_passes;   /* derivative passes actually run on the device (the reference repeats some evaluations
                              at an already evaluated pose; those are answered from the previous pass) */
} ndtb_result;

/* ---- context ---------------------------------------------------------------------------------- */
int ndtb_version(void);
const char *ndtb_strerror(int code);
const char *ndtb_last_error(const ndtb_ctx *ctx); /* text of the last CUDA error on this context */
/* device = CUDA ordinal. stream = cudaStream_t to enqueue on (NULL: the context creates its own). */
int ndtb_ctx_create(int device, void *stream, ndtb_ctx **out);
void ndtb_ctx_destroy(ndtb_ctx *ctx);
int ndtb_ctx_synchronize(ndtb_ctx *ctx);
/* number of kernels this context has launched so far (bench.py "gpu_launches") */
int64_t ndtb_ctx_launch_count(const ndtb_ctx *ctx);
int ndtb_ctx_sm_count(const ndtb_ctx *ctx);
/* Device-time accounting of the dominant kernel (bench.py roofline): when enabled, every launch of the
 * registration kernel is bracketed by CUDA events on the context's stream.  ndtb_ctx_match_time synchronises,
 * returns the accumulated milliseconds and launch count since the last call, and resets both. */
int ndtb_ctx_enable_timing(ndtb_ctx *ctx, int on);
int ndtb_ctx_match_time(ndtb_ctx *ctx, double *ms, int64_t *launches);
/* the same for batched map builds from device-resident points (all kernels of kernel (i), host waits included) */
int ndtb_ctx_build_time(ndtb_ctx *ctx, double *ms, int64_t *calls);
void ndtb_default_params(ndtb_params *p);

/* ---- maps: lslgeneric::NDTMap(new LazyGrid(cell)) --------------------------------------------- */
int ndtb_map_create(ndtb_ctx *ctx, double cell_x, double cell_y, double cell_z, ndtb_map **out);
void ndtb_map_destroy(ndtb_map *m);
/* NDTMap::guessSize(cx,cy,cz,sx,sy,sz) (floats upstream) — ndt_feature_fuser_hmt.cpp:222 */
int ndtb_map_guess_size(ndtb_map *m, double cx, double cy, double cz, double sx, double sy, double sz);
/* NDTMap::setMapSize(sx,sy,sz) */
int ndtb_map_set_map_size(ndtb_map *m, double sx, double sy, double sz);
/* NDTMap::initialize(cx,cy,cz,sx,sy,sz) — ndt_feature_fuser_hmt.cpp:89 */
int ndtb_map_initialize(ndtb_map *m, double cx, double cy, double cz, double sx, double sy, double sz);
/* NDTMap::loadPointCloud(pc, range_limit): (re)defines the grid, bins the points.  pts = n x 4 float
 * (pcl::PointXYZ: x,y,z,pad), in host or device memory.  *n_binned (optional) = points kept. */
int ndtb_map_load_point_cloud(ndtb_map *m, const float *pts, int64_t n, double range_limit, int mem,
                              int64_t *n_binned);
/* NDTMap::loadPointCloudCentroid(pc, origin, old_centroid, map_size, range_limit) [upstream] — the loadCentroid branch
 * of the local-map build, ndt_feature_fuser_hmt.cpp:199-217: the grid centre is old_centroid moved by a whole number of
 * cells towards origin (floor((origin - old_centroid) / cell) * cell per axis), its size is map_size, and points farther
 * than range_limit from ORIGIN are dropped. */
int ndtb_map_load_point_cloud_centroid(ndtb_map *m, const float *pts, int64_t n, int mem, const double *origin3,
                                       const double *old_centroid3, const double *map_size3, double range_limit);
/* NDTMap::addPointCloud, end-point binning only (no ray tracing; see ndtb_map_add_point_cloud) */
int ndtb_map_add_points(ndtb_map *m, const float *pts, int64_t n, int mem, int64_t *n_binned);
/* NDTMap::addPointCloud(origin, pc, classifierTh, maxz, sensor_noise, occupancy_limit) [upstream] WITH the free-space ray
 * trace (LazyGrid::traceLine) and the occupancy update of the cells every ray meets — the call the fuser makes on the node
 * map, ndt_feature_fuser_hmt.cpp:92 (initialize: 0.1, 100.0, 0.1) and :485 (update: 0.06, 25).  origin = 3 doubles
 * (sensor position in the map frame).  The rays are traced when ndtb_map_compute_cells runs (the reference calls the two
 * back to back); at most 4 addPointCloud calls may be pending.  Cells met by a ray exist afterwards (occupancy < 0) like
 * the cells of an initialize()d LazyGrid; on a lazily allocated grid upstream would additionally drop a fake point into
 * every cell a ray creates, which is not restated (never reached from the reference's call sites). */
int ndtb_map_add_point_cloud(ndtb_map *m, const double *origin, const float *pts, int64_t n, int mem, double classifier_th,
                             double maxz, double sensor_noise, double occupancy_limit);
/* NDTMap::computeNDTCells(CELL_UPDATE_MODE_SAMPLE_VARIANCE, maxnumpoints, occupancy_limit) */
int ndtb_map_compute_cells(ndtb_map *m, uint32_t maxnumpoints, float occupancy_limit);
/* Batched loadPointCloud + computeNDTCells over n_maps maps in a handful of launches (front-end and
 * bench path).  pts[i] / n_pts[i] per map; all pointers in the same memory space `mem`. */
int ndtb_map_build_batch(ndtb_ctx *ctx, int64_t n_maps, ndtb_map *const *maps, const float *const *pts,
                         const int64_t *n_pts, double range_limit, int mem, uint32_t maxnumpoints,
                         float occupancy_limit);
/* Build directly from cells (JFF fixtures / NDTMapMsg): placed by the voxel of their mean unless use_idx. */
int ndtb_map_from_cells(ndtb_map *m, const ndtb_grid *g, const ndtb_cell *cells, int64_t n, int use_idx);
int ndtb_map_grid(const ndtb_map *m, ndtb_grid *g);
int64_t ndtb_map_num_cells(const ndtb_map *m, int gaussian_only);
/* cells sorted by linear voxel index ((ix*sy)+iy)*sz+iz; returns the number available */
int64_t ndtb_map_export_cells(const ndtb_map *m, ndtb_cell *out, int64_t cap, int gaussian_only);
/* LazyGrid::getIndexForPoint for n points (parity hook): out = n x 3 int32, INT32_MIN for NaN points;
 * returns the number of in-bounds points or <0 */
int64_t ndtb_map_point_indices(const ndtb_map *m, const float *pts, int64_t n, int mem, int32_t *out);

/* ---- NDTMatcherD2D ---------------------------------------------------------------------------- */
/* derivativesNDT of the source cells moved by T against the target map.
 * out43 = score, g[6], H[36] row-major (H zero when !want_hessian); n_pairs optional. */
int ndtb_d2d_derivatives(ndtb_ctx *ctx, const ndtb_map *tgt, const ndtb_map *src, const double *T,
                         const ndtb_params *p, int want_hessian, double *out43, int64_t *n_pairs);
/* The cell-vector overload the reference's optimiser calls (ndt_matcher_d2d_fusion.h:856,617,444):
 * derivativesNDT(const std::vector<NDTCell*> &sourceNDT, const NDTMap &targetNDT, g, H, computeHessian) on host copies of
 * cells that were already moved (pseudoTransformNDT, :840); T (optional, NULL = identity) moves them further. */
int ndtb_d2d_derivatives_cells(ndtb_ctx *ctx, const ndtb_map *tgt, const ndtb_cell *src, int64_t n, const double *T,
                               const ndtb_params *p, int want_hessian, double *out43, int64_t *n_pairs);
/* NDTMatcherD2D::lineSearchMT(increment, sourceNDT, targetNDT) [upstream] (ndt_matcher_d2d_fusion.h:1013; the in-repo twin
 * is lineSearchMTFusion :390-793): More-Thuente step length along `increment6` for the cells as they are; increment6 may be
 * negated in place (wrong-direction case, :462).  Every trial evaluation is one gradient pass on the device. */
int ndtb_d2d_line_search_cells(ndtb_ctx *ctx, const ndtb_map *tgt, const ndtb_cell *src, int64_t n, double *increment6,
                               const ndtb_params *p, double *step);
/* NDTMatcherD2D::MoreThuente::cstep [upstream] == MINPACK dcstep (ndt_matcher_d2d_fusion.h:347,366,756,775): returns info */
int ndtb_mt_cstep(double *stx, double *fx, double *dx, double *sty, double *fy, double *dy, double *stp, double fp, double dp,
                  int *brackt, double stmin, double stmax);
/* Host hook of the engine's 3x3 symmetric eigen-solver (cyclic Jacobi: the decomposition behind NDTCell::rescaleCovariance in
 * the map build; csrc/optimizer.h eig_sym_n<3>).  stop_at_fixed_point = 1 is what the build runs on the covariances that
 * never meet the stopping test: it returns when a sweep changes no bit — the CPU tests check that this is bit-identical to
 * the full 64 sweeps.  A9 row-major; evals ascending; V9 eigenvectors in columns; sweeps (optional) = sweeps actually run. */
int ndtb_eig_sym3(const double *A9, int stop_at_fixed_point, double *evals3, double *V9, int32_t *sweeps);
int ndtb_d2d_match(ndtb_ctx *ctx, const ndtb_map *tgt, const ndtb_map *src, const double *T0,
                   const ndtb_params *p, ndtb_result *res);
/* matchFusion with useNDT=true, useFeat=false: soft constraint Q = Tcov^-1 (Tcov36 row-major 6x6) */
int ndtb_fusion_match(ndtb_ctx *ctx, const ndtb_map *tgt, const ndtb_map *src, const double *T0,
                      const double *Tcov36, const ndtb_params *p, ndtb_result *res);
/* covariance(target, source, T, cov) -> cov36 row-major */
int ndtb_d2d_covariance(ndtb_ctx *ctx, const ndtb_map *tgt, const ndtb_map *src, const double *T,
                        const ndtb_params *p, double *cov36);
/* n_edges independent registrations in one launch.  T0s = n_edges x 16.  with_covariance follows
 * ndt_feature_graph.cpp:286-310 (covariance() only if the pose changed, else 0.02*I).
 * res / cov36s (n_edges x 36, may be NULL) live in `out_mem` memory. */
int ndtb_d2d_match_batch(ndtb_ctx *ctx, int64_t n_edges, const ndtb_map *const *tgt,
                         const ndtb_map *const *src, const double *T0s, const ndtb_params *p,
                         int with_covariance, int out_mem, ndtb_result *res, double *cov36s);

/* ---- NDTMatcherP2D [upstream; no call site in the reference — BASELINE config C3] ----------------- */
/* Point-to-distribution NDT of a cloud (n x 4 float, host or device memory) against a map: the D2D score /
 * gradient / Hessian with a zero source covariance, the neighbourhood and the Newton / More-Thuente driver of
 * NDTMatcherD2D (DESIGN.md "P2D"; parity unpinned, defined by the oracle). */
int ndtb_p2d_derivatives(ndtb_ctx *ctx, const ndtb_map *tgt, const float *pts, int64_t n, int mem, const double *T,
                         const ndtb_params *p, int want_hessian, double *out43, int64_t *n_pairs);
int ndtb_p2d_match(ndtb_ctx *ctx, const ndtb_map *tgt, const float *pts, int64_t n, int mem, const double *T0,
                   const ndtb_params *p, ndtb_result *res);

/* Batched front-end step: for each pair build the NDT map of the target scan and of the source scan
 * (cell size `cell`, guess-size grids unless map_size[0..2] > 0: then setMapSize), register source onto
 * target from T0s, optionally compute the covariance.  Point pointers in `in_mem`, outputs in `out_mem`. */
int ndtb_register_scans(ndtb_ctx *ctx, int64_t n_pairs, const float *const *tgt_pts, const int64_t *n_tgt,
                        const float *const *src_pts, const int64_t *n_src, const double *T0s, double cell,
                        const double *map_size, double range_limit, const ndtb_params *p,
                        int with_covariance, int in_mem, int out_mem, ndtb_result *res, double *cov36s);

/* ---- JFF map files: NDTMap::writeToJFF / loadFromJFF [upstream]; call sites ndt_feature_fuser_hmt.cpp:15,24,39 ----
 * Host-side format code (no GPU needed): the grid + the cells that carry information (hasGaussian_, N > 0 or a
 * non-zero occupancy), ndtb_cell.idx = voxel index.  Byte layout in csrc/jff.cpp, decoded from the maps the reference
 * ships.  ndtb_jff_read_cells with cells == NULL returns the count in *n. */
int ndtb_jff_write_cells(const char *path, const ndtb_grid *g, const ndtb_cell *cells, int64_t n);
int ndtb_jff_read_cells(const char *path, ndtb_grid *g, ndtb_cell *cells, int64_t cap, int64_t *n);
/* the same for a map resident in HBM (export + write / read + ndtb_map_from_cells); 0 on success like upstream */
int ndtb_map_write_jff(const ndtb_map *m, const char *path);
int ndtb_map_load_jff(ndtb_map *m, const char *path);

/* lslgeneric::transformPointCloudInPlace [upstream]: the pose is cast to float and applied in float arithmetic
 * (ndt_feature_fuser_hmt.cpp:75-76,191,479).  in / out: n x 4 float in `in_mem` / `out_mem` memory (may alias). */
int ndtb_transform_point_cloud(ndtb_ctx *ctx, const double *T16, const float *in, int64_t n, int in_mem, float *out,
                               int out_mem);

/* ---- front end: NDTFeatureFuserHMT (ndt_feature/src/ndt_feature_src/ndt_feature_fuser_hmt.cpp) ----------------------
 * The per-scan step of the reference with useNDT only (useFeat = useOdom = false, loadCentroid = false: what every
 * offline configuration sets — ndt_graph_offline.cpp:306-317): the node map lives in HBM and is updated in place.
 *   initialize  :65-102   map->initialize + addPointCloud(origin, cloud, 0.1, 100, 0.1) + computeNDTCells(1e5, 255)
 *   update      :108-512  local map of the scan (:195-227) -> matchFusion vs the node map (:352-358) -> covariance
 *                         (:399-420) -> pose bookkeeping (:436-474) -> ray-traced addPointCloud(origin, cloud, 0.06, 25)
 *                         + computeNDTCells(1e5, 255) (:482-487)                                                      */
typedef struct ndtb_fuser ndtb_fuser;
typedef struct ndtb_fuser_params { /* NDTFeatureFuserHMT::Params (ndt_feature_fuser_hmt.h:58-207) + sensor pose + motion model */
  double resolution, map_size_x, map_size_y, map_size_z, sensor_range;
  double max_translation_norm, max_rotation_norm;
  double delta_score;            /* DELTA_SCORE */
  int32_t neighbours, itr_max;   /* neighbours, ITR_MAX */
  int32_t step_control;          /* stepcontrol */
  int32_t global_transf;         /* globalTransf */
  int32_t use_soft_constraints, use_tikhonov, compute_cov, fusion2d;
  int32_t all_matches_valid, fuse_incomplete, check_consistency, force_odom_as_est;
  double sensor_pose[16];        /* column-major 4x4 (setSensorPose) */
  double motion[6];              /* MotionModel2d::Params Cd, Ct, Dd, Dt, Td, Tt (motion_model.hpp:123-163) */
} ndtb_fuser_params;
void ndtb_fuser_default_params(ndtb_fuser_params *p);
int ndtb_fuser_create(ndtb_ctx *ctx, const ndtb_fuser_params *p, ndtb_fuser **out);
void ndtb_fuser_destroy(ndtb_fuser *f);
/* cloud = n x 4 float in the SENSOR frame (host or device memory) */
int ndtb_fuser_initialize(ndtb_fuser *f, const double *init_pose16, const float *cloud, int64_t n, int mem);
/* Tnow16 (out), res (optional: the registration record), cov36 (optional, row-major; zero when !compute_cov) */
int ndtb_fuser_update(ndtb_fuser *f, const double *Tmotion16, const float *cloud, int64_t n, int mem, int update_ndt_map,
                      double *Tnow16, ndtb_result *res, double *cov36);
ndtb_map *ndtb_fuser_map(ndtb_fuser *f); /* the node map (owned by the fuser) */
int ndtb_fuser_pose(const ndtb_fuser *f, double *Tnow16);
int ndtb_fuser_set_pose(ndtb_fuser *f, const double *Tnow16); /* NDTFeatureFuserHMT::Tnow is a public member (fuser_hmt.h:37) */

/* ---- NDTFeatureGraph front end (ndt_feature/src/ndt_feature_src/ndt_feature_graph.cpp:24-144): a chain of nodes, each
 * a fuser with its own map in its own frame; a new node is spawned when the odometric path length inside the current
 * node exceeds new_node_transl_dist (the last scan is registered against the old node without being fused). */
typedef struct ndtb_graph ndtb_graph;
int ndtb_graph_create(ndtb_ctx *ctx, const ndtb_fuser_params *p, double new_node_transl_dist, ndtb_graph **out);
void ndtb_graph_destroy(ndtb_graph *g);
int ndtb_graph_set_new_node_dist(ndtb_graph *g, double new_node_transl_dist);
int ndtb_graph_initialize(ndtb_graph *g, const double *init_pose16, const float *cloud, int64_t n, int mem);
int ndtb_graph_update(ndtb_graph *g, const double *Tmotion16, const float *cloud, int64_t n, int mem, double *Tnow16);
int64_t ndtb_graph_num_nodes(const ndtb_graph *g);
/* node k: pose T, Tlocal_odom, Tlocal_fuse (16 doubles each, optional), its map (borrowed), number of fused scans */
int ndtb_graph_node(ndtb_graph *g, int64_t k, double *T16, double *Tlocal_odom16, double *Tlocal_fuse16, ndtb_map **map,
                    int32_t *nb_updates);

/* ---- result hand-off formats (host code, no GPU; csrc/formats.cpp) ---------------------------------------------------
 * ndt_feature/NDTEdgeMsg: the ROS1 wire bytes of one refined link, as edgeToMsg builds it
 * (ndt_feature/include/ndt_feature/ndtgraph_conversion.h:17-34; ndt_feature/msg/NDTEdgeMsg.msg): u32 ref_idx, u32 mov_idx,
 * geometry_msgs/Pose T, Float64MultiArray cov (3x3), Float64MultiArray cov_3d (6x6, or empty when cov36 == NULL), f64 score.
 * cov9 / cov36 row-major.  Returns the message length (write happens only when cap suffices) or < 0. */
int64_t ndtb_edge_msg_pack(uint32_t ref_idx, uint32_t mov_idx, const double *T16, const double *cov9, const double *cov36,
                           double score, uint8_t *out, int64_t cap);
/* msgToEdge (ndtgraph_conversion.h:104-145) */
int ndtb_edge_msg_unpack(const uint8_t *buf, int64_t len, uint32_t *ref_idx, uint32_t *mov_idx, double *T16, double *cov9,
                         double *cov36, int32_t *has_cov36, double *score);
/* The whole-graph message, NDTGraphToMsg / msgToNDTGraph (ndtgraph_conversion.h:59-83,189-216; ndt_feature/msg/NDTGraphMsg.msg,
 * NDTNodeMsg.msg, NDTFeatureFuserHMTMsg.msg), composed from its parts.  Every *_pack returns the message length (bytes are
 * written only when cap suffices) or < 0; every *_unpack reports offsets into the caller's buffer.
 * ndt_map/NDTMapMsg is an upstream perception_oru type that the reference does not vendor: its field order is restated from
 * the published message definitions (lslgeneric::toMessage writes the Gaussian cells; fromMessage re-inserts each cell at the
 * voxel of its mean — ndtb_map_from_cells with use_idx = 0).  cells: all cells of a map (ndtb_map_export_cells). */
int64_t ndtb_map_msg_pack(uint32_t seq, uint32_t sec, uint32_t nsec, const char *frame_id, const ndtb_grid *g, const ndtb_cell *cells,
                          int64_t n_cells, uint8_t *out, int64_t cap);
int ndtb_map_msg_unpack(const uint8_t *buf, int64_t len, uint32_t *stamp3, char *frame_id, int32_t frame_cap, ndtb_grid *g,
                        ndtb_cell *cells, int64_t cells_cap, int64_t *n_cells, int64_t *consumed);
typedef struct ndtb_node_fields { /* NDTNodeMsg + its NDTFeatureFuserHMTMsg without the map (poses column-major 4x4) */
  double Tnow[16], Tlast_fuse[16], Todom[16]; /* fuserHMTToMsg, ndtgraph_conversion.h:36-45 */
  uint32_t ctr;
  uint32_t nb_updates;                        /* nodeToMsg, :47-57 */
  double T[16];
  double cov9[9];                             /* row-major 3x3 */
  double Tlocal_odom[16], Tlocal_fuse[16];
  double time_last_update;
} ndtb_node_fields;
int64_t ndtb_node_msg_pack(const ndtb_node_fields *f, const uint8_t *map_msg, int64_t map_len, uint8_t *out, int64_t cap);
int ndtb_node_msg_unpack(const uint8_t *buf, int64_t len, ndtb_node_fields *f, int64_t *map_off, int64_t *map_len, int64_t *consumed);
int64_t ndtb_graph_msg_pack(uint32_t seq, uint32_t sec, uint32_t nsec, const char *frame_id, const double *sensor_pose16,
                            const double *Tnow16, double distance_moved, int64_t n_nodes, const uint8_t *const *node_msgs,
                            const int64_t *node_lens, int64_t n_edges, const uint8_t *const *edge_msgs, const int64_t *edge_lens,
                            uint8_t *out, int64_t cap);
int ndtb_graph_msg_unpack(const uint8_t *buf, int64_t len, uint32_t *stamp3, char *frame_id, int32_t frame_cap, double *sensor_pose16,
                          double *Tnow16, double *distance_moved, int64_t *n_nodes, int64_t *node_off, int64_t *node_len,
                          int64_t nodes_cap, int64_t *n_edges, int64_t *edge_off, int64_t *edge_len, int64_t edges_cap);
/* saveAffine3d / loadAffine3d of NDTFeatureNode::save / load (ndt_feature_node.h:100-152): the boost text archives
 * mapping{k}.T, ...local_odom.T, ...local_fuse.T the reference ships (byte-compatible) */
int ndtb_pose_archive_write(const char *path, const double *T16);
int ndtb_pose_archive_read(const char *path, double *T16);
/* transformToEvalString (planar = 0) / transformToEval2dString (planar = 1), utils.h:243-259: one line of the est / gt
 * trajectory files ("x y z qx qy qz qw\n", 15 significant digits).  Returns the length or < 0. */
int ndtb_eval_string(const double *T16, int planar, char *out, int32_t cap);

/* ---- multi-GPU: one process (one ndtb_ctx) per GPU.  Edges / scan pairs are independent units, so the only cross-GPU
 * step is the gather of the fixed-size result records after a sharded batch — the serial loop over links of
 * NDTFeatureGraph::updateLinksUsingNDTRegistration (ndt_feature_graph.cpp:347-353) split over ranks.  NCCL (all-gather
 * over NVLink / NVSwitch) is loaded at run time (libnccl.so.2); without it these calls return NDTB_ERR_CUDA. */
typedef struct ndtb_comm ndtb_comm;
/* rank 0 creates the id and shares it with the other ranks out of band (file, socket, MPI, the launcher's store) */
int ndtb_comm_unique_id(char id128[128]);
/* Sets NCCL_MAX_CTAS=1 for the process unless the caller has set it: the records are tiny, and a collective kernel with NCCL's
 * default CTA count holds SMs the registration kernels need while it waits for the slowest rank (7 % of a step on 8 GPUs). */
int ndtb_comm_create(ndtb_ctx *ctx, const char id128[128], int rank, int world, ndtb_comm **out);
void ndtb_comm_destroy(ndtb_comm *c);
/* every rank contributes n_local records (device memory, the same n_local on every rank: pad the last shard);
 * all_dev (device memory, world * n_local records) receives rank r's records at [r * n_local, (r+1) * n_local).
 * Enqueued on the context's stream; synchronise the context before reading all_dev from the host. */
int ndtb_gather_results(ndtb_comm *c, const ndtb_result *local_dev, int64_t n_local, ndtb_result *all_dev);

/* ndt_feature::overlapNDTOccupancyScore(ref, mov, T) */
int ndtb_overlap_score(ndtb_ctx *ctx, const ndtb_map *ref, const ndtb_map *mov, const double *T,
                       double *score);
/* the same for n links in one launch (updateLinksUsingNDTRegistration with !keepScore, ndt_feature_graph.cpp:335-342).
 * T: n poses T_stride_bytes apart (128 for packed poses, sizeof(ndtb_result) to read the T field of result records) in
 * `T_mem` memory; scores: n doubles in `out_mem` memory. */
int ndtb_overlap_score_batch(ndtb_ctx *ctx, int64_t n, const ndtb_map *const *ref, const ndtb_map *const *mov, const void *T,
                             int64_t T_stride_bytes, int T_mem, int out_mem, double *scores);

#ifdef __cplusplus
}
#endif
#endif /* NDTB_H */
