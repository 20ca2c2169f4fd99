// ndtb_lslgeneric.hpp — header-only C++ host façade over the C ABI (ndtb.h).
//
// It carries the class and method names the reference's L3/L4 sources use for the NDT hot path, so that
// ndt_feature_graph.cpp / ndt_feature_fuser_hmt.cpp compile against it instead of perception_oru's
// ndt_map / ndt_registration headers:
//
//   lslgeneric::LazyGrid(double cellSize), lslgeneric::SpatialIndex       ndt_feature_fuser_hmt.cpp:87,195
//   lslgeneric::NDTMap(SpatialIndex*, bool dealloc=false)                  ndt_feature_fuser_hmt.cpp:87,196
//     initialize / guessSize / setMapSize / loadPointCloud / addPointCloud / computeNDTCells
//                                                                         ndt_feature_fuser_hmt.cpp:89-94,201-227,485-486
//     writeToJFF / loadFromJFF                                            ndt_feature_fuser_hmt.cpp:15,24,39
//     numberOfActiveCells / getAllCells / getAllInitializedCells / pseudoTransformNDT / getCentroid
//                                                                         ndt_matcher_d2d_fusion.h:840, ndt_feature_node.h:216
//   lslgeneric::NDTCell {getMean,getCov,setMean,setCov,getCenter,getOccupancy,hasGaussian_}
//   lslgeneric::NDTMatcherD2D {n_neighbours, ITR_MAX, DELTA_SCORE, step_control; match; covariance; derivativesNDT}
//                                                                         ndt_feature_graph.cpp:261-298, ndt_matcher_d2d_fusion.h:856
//   lslgeneric::NDTMatcherD2D_2D, ndt_feature::matchFusion2d              ndt_matcher_d2d_fusion.h:1159-1176
//   lslgeneric::NDTMatcherP2D {match}                                      (no call site in the reference; BASELINE config C3)
//   ndt_feature::matchFusion (NDT term + soft constraint / Tikhonov)       ndt_matcher_d2d_fusion.h:797-1155
//   ndt_feature::overlapNDTOccupancyScore                                  ndt_feature_node.h:213-252
//   ndtb::GraphRegistrar::updateLinksUsingNDTRegistration                  ndt_feature_graph.cpp:347-353 (one batched launch)
//
// Eigen and PCL are used when present (__has_include); otherwise minimal stand-ins with the same member names
// (Eigen::Affine3d::matrix()/data()/operator(), Eigen::MatrixXd, pcl::PointXYZ, pcl::PointCloud<T>::points) are
// defined so that the façade and its tests build in containers without those libraries.
//
// Error behaviour mirrors the reference: match() returns bool (false <=> ITR_MAX exceeded or engine error),
// covariance() returns bool, JFF-style int==0 for success elsewhere; nothing throws on the hot path, nothing
// prints, nothing blocks on stdin (cf. ndt_feature_graph.cpp:318-328).  The last engine error is kept in
// ndtb::last_status().  There is no CPU implementation: without a CUDA device every call fails.
#pragma once

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <vector>

#include "ndtb.h"

#if defined(__has_include)
#if __has_include(<Eigen/Geometry>) && !defined(NDTB_NO_EIGEN)
#include <Eigen/Dense>
#include <Eigen/Geometry>
#define NDTB_HAVE_EIGEN 1
#endif
#if __has_include(<pcl/point_cloud.h>) && !defined(NDTB_NO_PCL)
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#define NDTB_HAVE_PCL 1
#endif
#endif

#ifndef NDTB_HAVE_EIGEN
namespace Eigen {  // stand-ins: just enough surface for the reference's call sites on this path
struct Vector3d {
  double v[3] = {0, 0, 0};
  Vector3d() {}
  Vector3d(double x, double y, double z) : v{x, y, z} {}
  double &operator()(int i) { return v[i]; }
  double operator()(int i) const { return v[i]; }
  double &operator[](int i) { return v[i]; }
  double operator[](int i) const { return v[i]; }
  const double *data() const { return v; }
  double *data() { return v; }
};
struct Matrix3d {
  double m[9] = {0};  // column-major
  double &operator()(int r, int c) { return m[c * 3 + r]; }
  double operator()(int r, int c) const { return m[c * 3 + r]; }
  const double *data() const { return m; }
  double *data() { return m; }
};
struct MatrixXd {
  int r = 0, c = 0;
  std::vector<double> m;  // column-major
  MatrixXd() {}
  MatrixXd(int rows, int cols) : r(rows), c(cols), m((size_t)rows * cols, 0.0) {}
  void resize(int rows, int cols) { r = rows, c = cols, m.assign((size_t)rows * cols, 0.0); }
  void setZero() { std::fill(m.begin(), m.end(), 0.0); }
  int rows() const { return r; }
  int cols() const { return c; }
  double &operator()(int i, int j) { return m[(size_t)j * r + i]; }
  double operator()(int i, int j) const { return m[(size_t)j * r + i]; }
  const double *data() const { return m.data(); }
  double *data() { return m.data(); }
};
template <typename S, int R, int C>
struct Matrix {  // fixed-size column-major matrix / vector (Eigen::Matrix<double,6,1> increment of the line searches)
  S v[R * C] = {};
  S &operator()(int i) { return v[i]; }
  S operator()(int i) const { return v[i]; }
  S &operator()(int i, int j) { return v[j * R + i]; }
  S operator()(int i, int j) const { return v[j * R + i]; }
  void setZero() {
    for (int i = 0; i < R * C; i++) v[i] = S(0);
  }
  const S *data() const { return v; }
  S *data() { return v; }
};
struct Affine3d {
  double m[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};  // column-major 4x4
  struct MatrixRef {
    double *p;
    double *data() { return p; }
    double &operator()(int r, int c) { return p[c * 4 + r]; }
  };
  struct ConstMatrixRef {
    const double *p;
    const double *data() const { return p; }
    double operator()(int r, int c) const { return p[c * 4 + r]; }
  };
  MatrixRef matrix() { return MatrixRef{m}; }
  ConstMatrixRef matrix() const { return ConstMatrixRef{m}; }
  double &operator()(int r, int c) { return m[c * 4 + r]; }
  double operator()(int r, int c) const { return m[c * 4 + r]; }
  const double *data() const { return m; }
  double *data() { return m; }
  Vector3d translation() const { return Vector3d(m[12], m[13], m[14]); }
  void setIdentity() {
    for (int i = 0; i < 16; i++) m[i] = (i % 5 == 0) ? 1.0 : 0.0;
  }
  static Affine3d Identity() { return Affine3d(); }
};
}  // namespace Eigen
#endif

#ifndef NDTB_HAVE_PCL
namespace pcl {
struct alignas(16) PointXYZ {
  float x = 0, y = 0, z = 0, pad_ = 1.f;
  PointXYZ() {}
  PointXYZ(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
};
template <class P>
struct PointCloud {
  std::vector<P> points;
  uint32_t width = 0, height = 1;
  bool is_dense = true;
  size_t size() const { return points.size(); }
  void push_back(const P &p) { points.push_back(p), width = (uint32_t)points.size(); }
  const P &front() const { return points.front(); }
  const P &back() const { return points.back(); }
  typename std::vector<P>::const_iterator begin() const { return points.begin(); }
  typename std::vector<P>::const_iterator end() const { return points.end(); }
};
}  // namespace pcl
#endif
static_assert(sizeof(pcl::PointXYZ) == 16, "pcl::PointXYZ must be float4-compatible (ndtb.h point layout)");

namespace ndtb {

// One engine context per host thread and device (ndtb_ctx is single-owner).  Created on first use.
inline int &current_device() {
  static thread_local int d = 0;
  return d;
}
inline int &last_status() {
  static thread_local int s = NDTB_OK;
  return s;
}
// Contexts are created on first use, one per (host thread, device), and are NEVER destroyed before process exit: every
// NDTMap keeps the raw context pointer (its storage is freed stream-ordered on that context), so a map may outlive the
// thread that created it or a change of current_device().
struct CtxHolder {
  std::vector<std::pair<int, ndtb_ctx *>> ctxs;  // (device, context) of this thread; intentionally leaked at thread exit
};
// Returns nullptr (and sets last_status) when no CUDA device / library problem: callers fail like the reference
// fails, by returning false / non-zero.
inline ndtb_ctx *context() {
  static thread_local CtxHolder h;
  for (auto &dc : h.ctxs)
    if (dc.first == current_device()) return dc.second;
  ndtb_ctx *c = nullptr;
  last_status() = ndtb_ctx_create(current_device(), nullptr, &c);
  if (last_status() != NDTB_OK) return nullptr;
  h.ctxs.push_back({current_device(), c});
  return c;
}

template <class Affine>
inline const double *pose_data(const Affine &T) {
  return T.matrix().data();
}
template <class Affine>
inline double *pose_data(Affine &T) {
  return T.matrix().data();
}

}  // namespace ndtb

#ifndef CELL_UPDATE_MODE_SAMPLE_VARIANCE
// NDTCell update modes [upstream ndt_map/ndt_cell.h]; only SAMPLE_VARIANCE is reachable from the reference
// (ndt_feature_fuser_hmt.cpp:94,227,486)
#define CELL_UPDATE_MODE_COVARIANCE_INTERSECTION 0
#define CELL_UPDATE_MODE_SAMPLE_VARIANCE 1
#define CELL_UPDATE_MODE_ERROR_REFINEMENT 2
#define CELL_UPDATE_MODE_SAMPLE_VARIANCE_SURFACE_ESTIMATION 3
#define CELL_UPDATE_MODE_STUDENT_T 4
#endif

namespace lslgeneric {

class SpatialIndex {
 public:
  virtual ~SpatialIndex() {}
  virtual void getCellSize(double &cx, double &cy, double &cz) const = 0;
};

class LazyGrid : public SpatialIndex {
 public:
  explicit LazyGrid(double cellSize) : cx_(cellSize), cy_(cellSize), cz_(cellSize) {}
  LazyGrid(double cx, double cy, double cz) : cx_(cx), cy_(cy), cz_(cz) {}
  void getCellSize(double &cx, double &cy, double &cz) const override { cx = cx_, cy = cy_, cz = cz_; }

 private:
  double cx_, cy_, cz_;
};

class NDTCell;
// lslgeneric::CellVector [upstream]: the index of the feature / odometry NDT maps (ndt_feature_fuser_hmt.cpp:281-325).
// Those maps hold at most a few dozen correspondence cells and are empty in every shipped configuration
// (useFeat = useOdom = false); kept as a host container so that the reference's call sites compile.
class CellVector : public SpatialIndex {
 public:
  ~CellVector() override;
  void getCellSize(double &cx, double &cy, double &cz) const override { cx = cy = cz = 1.0; }
  void addCell(NDTCell *c) { cells_.push_back(c); }
  void addNDTCell(NDTCell *c) { cells_.push_back(c); }
  int size() const { return (int)cells_.size(); }
  NDTCell *getCellIdx(unsigned int i) const { return i < cells_.size() ? cells_[i] : nullptr; }

 private:
  std::vector<NDTCell *> cells_;
};

// Host snapshot of one cell (what pseudoTransformNDT / getAllCells hand to legacy code; caller deletes).
class NDTCell {
 public:
  bool hasGaussian_ = false;
  NDTCell() {}
  explicit NDTCell(const ndtb_cell &c, const ndtb_grid &g) { from(c, g); }
  Eigen::Vector3d getMean() const { return Eigen::Vector3d(mean_[0], mean_[1], mean_[2]); }
  Eigen::Matrix3d getCov() const {
    Eigen::Matrix3d C;
    C(0, 0) = cov_[0], C(0, 1) = C(1, 0) = cov_[1], C(0, 2) = C(2, 0) = cov_[2];
    C(1, 1) = cov_[3], C(1, 2) = C(2, 1) = cov_[4], C(2, 2) = cov_[5];
    return C;
  }
  void setMean(const Eigen::Vector3d &m) { mean_[0] = m(0), mean_[1] = m(1), mean_[2] = m(2); }
  void setCov(const Eigen::Matrix3d &C) {
    cov_[0] = C(0, 0), cov_[1] = C(0, 1), cov_[2] = C(0, 2), cov_[3] = C(1, 1), cov_[4] = C(1, 2), cov_[5] = C(2, 2);
  }
  pcl::PointXYZ getCenter() const { return pcl::PointXYZ((float)center_[0], (float)center_[1], (float)center_[2]); }
  float getOccupancy() const { return occ_; }
  double getOccupancyRescaled() const { return 1.0 - 1.0 / (1.0 + std::exp((double)occ_)); }
  int getN() const { return n_; }
  NDTCell *copy() const { return new NDTCell(*this); }
  NDTCell *clone() const {
    NDTCell *c = new NDTCell();
    std::memcpy(c->center_, center_, sizeof center_);
    return c;
  }
  const double *mean_data() const { return mean_; }
  const double *cov_data() const { return cov_; }
  void to(ndtb_cell &c) const {
    std::memcpy(c.mean, mean_, sizeof mean_), std::memcpy(c.cov, cov_, sizeof cov_);
    c.n = n_, c.has_gaussian = hasGaussian_, c.occ = occ_;
    c.idx[0] = idx_[0], c.idx[1] = idx_[1], c.idx[2] = idx_[2];
  }

 private:
  void from(const ndtb_cell &c, const ndtb_grid &g) {
    std::memcpy(mean_, c.mean, sizeof mean_), std::memcpy(cov_, c.cov, sizeof cov_);
    n_ = c.n, hasGaussian_ = c.has_gaussian != 0, occ_ = c.occ;
    for (int a = 0; a < 3; a++) {
      idx_[a] = c.idx[a];
      center_[a] = g.center[a] + (c.idx[a] - (int)(g.size[a] / 2.0)) * g.cell[a];  // LazyGrid cell centre
    }
  }
  double mean_[3] = {0, 0, 0}, cov_[6] = {0, 0, 0, 0, 0, 0}, center_[3] = {0, 0, 0};
  int n_ = 0, idx_[3] = {0, 0, 0};
  float occ_ = 0.f;
};

inline CellVector::~CellVector() {
  for (NDTCell *c : cells_) delete c;
}

class NDTMap {
 public:
  // NDTMap(new LazyGrid(res)) / NDTMap(idx, true): the index only carries the resolution, the grid lives in HBM
  explicit NDTMap(SpatialIndex *idx, bool dealloc = false) {
    double cx = 0.5, cy = 0.5, cz = 0.5;
    if (idx) idx->getCellSize(cx, cy, cz);
    if (CellVector *cv = dynamic_cast<CellVector *>(idx)) {  // a feature / odometry map: host cells only
      cell_vector_ = cv, owns_index_ = dealloc;
      return;
    }
    if (dealloc) delete idx;  // upstream keeps it as prototype and frees it in ~NDTMap; nothing else to keep here
    if (ndtb_ctx *c = ndtb::context()) ndtb::last_status() = ndtb_map_create(c, cx, cy, cz, &h_);
  }
  ~NDTMap() {
    if (h_) ndtb_map_destroy(h_);
    if (cell_vector_ && owns_index_) delete cell_vector_;
  }
  SpatialIndex *getMyIndex() const { return cell_vector_; }
  // pseudoTransformNDTMap(T) of a CellVector map (ndt_feature_fuser_hmt.cpp:297-305): a new host map, caller deletes
  template <class Affine>
  NDTMap *pseudoTransformNDTMap(const Affine & /*T*/) const {
    return new NDTMap(new CellVector(), true);  // the feature branch is outside the engine's scope: its maps stay empty
  }
  // loadPointCloudCentroid(pc, origin, old_centroid, map_size, range_limit) (ndt_feature_fuser_hmt.cpp:199-217)
  template <class Cloud>
  void loadPointCloudCentroid(const Cloud &pc, const Eigen::Vector3d &origin, const Eigen::Vector3d &old_centroid,
                              const Eigen::Vector3d &map_size, double range_limit) {
    if (!h_) return;
    const double o[3] = {origin(0), origin(1), origin(2)}, c[3] = {old_centroid(0), old_centroid(1), old_centroid(2)};
    const double ms[3] = {map_size(0), map_size(1), map_size(2)};
    ndtb::last_status() = ndtb_map_load_point_cloud_centroid(h_, reinterpret_cast<const float *>(pc.points.data()),
                                                             (int64_t)pc.points.size(), NDTB_MEM_HOST, o, c, ms, range_limit);
  }
  NDTMap(const NDTMap &) = delete;
  NDTMap &operator=(const NDTMap &) = delete;

  void initialize(double cenx, double ceny, double cenz, double sizex, double sizey, double sizez) {
    if (h_) ndtb::last_status() = ndtb_map_initialize(h_, cenx, ceny, cenz, sizex, sizey, sizez);
  }
  void guessSize(float cenx, float ceny, float cenz, float sizex, float sizey, float sizez) {
    if (h_) ndtb::last_status() = ndtb_map_guess_size(h_, cenx, ceny, cenz, sizex, sizey, sizez);
  }
  void setMapSize(float sx, float sy, float sz) {
    if (h_) ndtb::last_status() = ndtb_map_set_map_size(h_, sx, sy, sz);
  }
  template <class Cloud>
  void loadPointCloud(const Cloud &pc, double range_limit = -1) {
    if (h_)
      ndtb::last_status() = ndtb_map_load_point_cloud(h_, reinterpret_cast<const float *>(pc.points.data()),
                                                      (int64_t)pc.points.size(), range_limit, NDTB_MEM_HOST, nullptr);
  }
  // addPointCloud(origin, pc, classifierTh, maxz, sensor_noise, occupancy_limit) [upstream]: end points binned, every
  // ray from `origin` traced through the grid and the occupancy of the cells it meets updated (LazyGrid::traceLine);
  // call sites ndt_feature_fuser_hmt.cpp:92,485.  Takes effect at the next computeNDTCells, as the reference calls them.
  template <class Cloud>
  void addPointCloud(const Eigen::Vector3d &origin, const Cloud &pc, double classifierTh = 0.06, double maxz = 100.0,
                     double sensor_noise = 0.25, double occupancy_limit = 255) {
    if (!h_) return;
    const double o[3] = {origin(0), origin(1), origin(2)};
    ndtb::last_status() = ndtb_map_add_point_cloud(h_, o, reinterpret_cast<const float *>(pc.points.data()),
                                                   (int64_t)pc.points.size(), NDTB_MEM_HOST, classifierTh, maxz, sensor_noise,
                                                   occupancy_limit);
  }
  void computeNDTCells(int cellupdatemode = CELL_UPDATE_MODE_SAMPLE_VARIANCE, unsigned int maxnumpoints = 1000000000u,
                       float occupancy_limit = 255, Eigen::Vector3d /*origin*/ = Eigen::Vector3d(0, 0, 0),
                       double /*sensor_noise*/ = 0.1) {
    if (!h_) return;
    ndtb::last_status() = cellupdatemode == CELL_UPDATE_MODE_SAMPLE_VARIANCE
                              ? ndtb_map_compute_cells(h_, maxnumpoints, occupancy_limit)
                              : NDTB_ERR_ARG;  // other update modes are not reachable from the reference
  }
  // writeToJFF / loadFromJFF (ndt_feature_fuser_hmt.cpp:15,24,39): 0 on success, like upstream
  int writeToJFF(const char *filename) { return h_ ? ndtb_map_write_jff(h_, filename) : NDTB_ERR_CUDA; }
  int loadFromJFF(const char *filename) { return h_ ? ndtb_map_load_jff(h_, filename) : NDTB_ERR_CUDA; }
  int numberOfActiveCells() const { return h_ ? (int)ndtb_map_num_cells(h_, 1) : 0; }
  bool getGridSizeInMeters(double &cx, double &cy, double &cz) const {
    ndtb_grid g;
    if (!h_ || ndtb_map_grid(h_, &g) != NDTB_OK) return false;
    cx = g.size[0] * g.cell[0], cy = g.size[1] * g.cell[1], cz = g.size[2] * g.cell[2];
    return true;
  }
  bool getCentroid(double &cx, double &cy, double &cz) const {
    ndtb_grid g;
    if (!h_ || ndtb_map_grid(h_, &g) != NDTB_OK) return false;
    cx = g.center[0], cy = g.center[1], cz = g.center[2];
    return true;
  }
  // cells with a Gaussian, as heap copies the CALLER deletes (upstream ownership convention)
  std::vector<NDTCell *> getAllCells() const { return snapshot(true, nullptr); }
  std::vector<NDTCell *> getAllInitializedCells() const { return snapshot(false, nullptr); }
  // pseudoTransformNDT(T): Gaussian cells moved by T (mean <- T mean, cov <- R cov R^T); caller deletes.
  // The matcher never needs this (the kernels move cells on the fly); it exists for legacy callers
  // (ndt_matcher_d2d_fusion.h:840).
  template <class Affine>
  std::vector<NDTCell *> pseudoTransformNDT(const Affine &T) const {
    return snapshot(true, ndtb::pose_data(T));
  }
  ndtb_map *handle() const { return h_; }

 private:
  std::vector<NDTCell *> snapshot(bool gaussian_only, const double *T) const {
    std::vector<NDTCell *> out;
    ndtb_grid g;
    if (!h_ || ndtb_map_grid(h_, &g) != NDTB_OK) return out;
    const int64_t n = ndtb_map_num_cells(h_, 0);
    if (n <= 0) return out;
    std::vector<ndtb_cell> cells((size_t)n);
    const int64_t k = ndtb_map_export_cells(h_, cells.data(), n, gaussian_only ? 1 : 0);
    for (int64_t i = 0; i < k; i++) {
      NDTCell *c = new NDTCell(cells[i], g);
      if (T) {
        double R[9], m[3], S[9], A[9];
        for (int r = 0; r < 3; r++)
          for (int q = 0; q < 3; q++) R[r * 3 + q] = T[q * 4 + r];
        const double *mu = cells[i].mean, *cv = cells[i].cov;
        for (int r = 0; r < 3; r++) m[r] = R[r * 3] * mu[0] + R[r * 3 + 1] * mu[1] + R[r * 3 + 2] * mu[2] + T[12 + r];
        const double full[9] = {cv[0], cv[1], cv[2], cv[1], cv[3], cv[4], cv[2], cv[4], cv[5]};
        for (int r = 0; r < 3; r++)
          for (int q = 0; q < 3; q++) A[r * 3 + q] = R[r * 3] * full[q] + R[r * 3 + 1] * full[3 + q] + R[r * 3 + 2] * full[6 + q];
        for (int r = 0; r < 3; r++)
          for (int q = 0; q < 3; q++) S[r * 3 + q] = A[r * 3] * R[q * 3] + A[r * 3 + 1] * R[q * 3 + 1] + A[r * 3 + 2] * R[q * 3 + 2];
        Eigen::Matrix3d C;
        for (int r = 0; r < 3; r++)
          for (int q = 0; q < 3; q++) C(r, q) = S[r * 3 + q];
        c->setMean(Eigen::Vector3d(m[0], m[1], m[2]));
        c->setCov(C);
      }
      out.push_back(c);
    }
    return out;
  }
  ndtb_map *h_ = nullptr;
  CellVector *cell_vector_ = nullptr;
  bool owns_index_ = false;
};

class NDTMatcherD2D {
 public:
  // public knobs of upstream's matcher, same names and defaults (init(): ITR_MAX 30, DELTA_SCORE 10e-3*0.1,
  // step_control true, n_neighbours 2)
  int n_neighbours = 2;
  int ITR_MAX = 30;
  double DELTA_SCORE = 10e-3 * 0.1;
  bool step_control = true;
  bool regularize = true;
  int iteration_counter_internal = 0;  // iterations of the last match() (upstream member of the same name)
  double finalscore = 0;               // score of the last match()

  NDTMatcherD2D() {}
  NDTMatcherD2D(bool /*isIrregularGrid*/, bool /*useDefaultGridResolutions*/, std::vector<double> /*resolutions*/) {}

  // NDTMatcherD2D::MoreThuente [upstream]: the helpers the in-repo line searches call (ndt_matcher_d2d_fusion.h:154-165,
  // 320,347,366,529-540,729,756,775)
  struct MoreThuente {
    static double min(double a, double b) { return a < b ? a : b; }
    static double max(double a, double b) { return a > b ? a : b; }
    static int cstep(double &stx, double &fx, double &dx, double &sty, double &fy, double &dy, double &stp, double &fp,
                     double &dp, bool &brackt, double stmin, double stmax) {
      int b = brackt ? 1 : 0;
      const int info = ndtb_mt_cstep(&stx, &fx, &dx, &sty, &fy, &dy, &stp, fp, dp, &b, stmin, stmax);
      brackt = b != 0;
      return info;
    }
  };

  // derivativesNDT(sourceNDT, targetNDT, score_gradient, Hessian, computeHessian) on a vector of already moved cells
  // (ndt_matcher_d2d_fusion.h:856,617,444)
  template <class Mat>
  double derivativesNDT(const std::vector<NDTCell *> &sourceNDT, const NDTMap &targetNDT, Mat &score_gradient, Mat &Hessian,
                        bool computeHessian) {
    ndtb_ctx *c = ndtb::context();
    if (!c || !targetNDT.handle()) return 0.0;
    std::vector<ndtb_cell> cells(sourceNDT.size());
    for (size_t i = 0; i < sourceNDT.size(); i++) sourceNDT[i]->to(cells[i]);
    double out[43];
    const ndtb_params p = params();
    ndtb::last_status() = ndtb_d2d_derivatives_cells(c, targetNDT.handle(), cells.data(), (int64_t)cells.size(), nullptr, &p,
                                                     computeHessian, out, nullptr);
    if (ndtb::last_status() != NDTB_OK) return 0.0;
    score_gradient.resize(6, 1);
    Hessian.resize(6, 6);
    for (int i = 0; i < 6; i++) {
      score_gradient(i, 0) = out[1 + i];
      for (int j = 0; j < 6; j++) Hessian(i, j) = out[7 + i * 6 + j];
    }
    return out[0];
  }
  // lineSearchMT(increment, sourceNDT, targetNDT) -> step (ndt_matcher_d2d_fusion.h:1013); increment may be negated
  template <class Vec6>
  double lineSearchMT(Vec6 &increment, std::vector<NDTCell *> &sourceNDT, NDTMap &targetNDT) {
    ndtb_ctx *c = ndtb::context();
    if (!c || !targetNDT.handle()) return 0.0;
    std::vector<ndtb_cell> cells(sourceNDT.size());
    for (size_t i = 0; i < sourceNDT.size(); i++) sourceNDT[i]->to(cells[i]);
    double inc[6], step = 0.0;
    for (int i = 0; i < 6; i++) inc[i] = increment(i);
    const ndtb_params p = params();
    ndtb::last_status() = ndtb_d2d_line_search_cells(c, targetNDT.handle(), cells.data(), (int64_t)cells.size(), inc, &p, &step);
    for (int i = 0; i < 6; i++) increment(i) = inc[i];
    return ndtb::last_status() == NDTB_OK ? step : 0.0;
  }

  ndtb_params params() const {
    ndtb_params p;
    ndtb_default_params(&p);
    p.n_neighbours = n_neighbours, p.itr_max = ITR_MAX, p.delta_score = DELTA_SCORE;
    p.step_control = step_control, p.regularize = regularize;
    p.planar = planar_ ? 1 : 0;
    return p;
  }

  // match(target, source, T, useInitialGuess) -> converged; T is refined in place (ndt_feature_graph.cpp:273)
  template <class Affine>
  bool match(NDTMap &target, NDTMap &source, Affine &T, bool useInitialGuess = false) {
    ndtb_ctx *c = ndtb::context();
    if (!c || !target.handle() || !source.handle()) return false;
    double T0[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    if (useInitialGuess) std::memcpy(T0, ndtb::pose_data(T), sizeof T0);
    const ndtb_params p = params();
    ndtb_result r;
    ndtb::last_status() = ndtb_d2d_match(c, target.handle(), source.handle(), T0, &p, &r);
    if (ndtb::last_status() != NDTB_OK) return false;
    std::memcpy(ndtb::pose_data(T), r.T, sizeof r.T);
    iteration_counter_internal = r.iterations, finalscore = r.score;
    last = r;
    return r.converged != 0;
  }
  // covariance(target, source, T, cov) (ndt_feature_graph.cpp:298, ndt_feature_fuser_hmt.cpp:405)
  template <class Affine, class Mat>
  bool covariance(NDTMap &target, NDTMap &source, Affine &T, Mat &cov) {
    ndtb_ctx *c = ndtb::context();
    if (!c || !target.handle() || !source.handle()) return false;
    double out[36];
    const ndtb_params p = params();
    ndtb::last_status() = ndtb_d2d_covariance(c, target.handle(), source.handle(), ndtb::pose_data(T), &p, out);
    if (ndtb::last_status() != NDTB_OK) return false;
    cov.resize(6, 6);
    for (int i = 0; i < 6; i++)
      for (int j = 0; j < 6; j++) cov(i, j) = out[i * 6 + j];
    return true;
  }
  // derivativesNDT on a map pair at pose T (the cell-vector overload of upstream is replaced by the map + pose:
  // the engine moves the source cells on the fly).  score_gradient 6x1, Hessian 6x6.
  template <class Affine, class Mat>
  double derivativesNDT(NDTMap &source, const Affine &T, NDTMap &target, Mat &score_gradient, Mat &Hessian,
                        bool computeHessian) {
    ndtb_ctx *c = ndtb::context();
    if (!c) return 0.0;
    double out[43];
    const ndtb_params p = params();
    ndtb::last_status() =
        ndtb_d2d_derivatives(c, target.handle(), source.handle(), ndtb::pose_data(T), &p, computeHessian, out, nullptr);
    if (ndtb::last_status() != NDTB_OK) return 0.0;
    score_gradient.resize(6, 1);
    Hessian.resize(6, 6);
    for (int i = 0; i < 6; i++) {
      score_gradient(i, 0) = out[1 + i];
      for (int j = 0; j < 6; j++) Hessian(i, j) = out[7 + i * 6 + j];
    }
    return out[0];
  }
  ndtb_result last = {};

 protected:
  bool planar_ = false;
};

// NDTMatcherD2D_2D [upstream]: the same matcher estimating (x, y, yaw) only; the reference reaches it through
// matchFusion2d (ndt_matcher_d2d_fusion.h:1159-1176) when NDTFeatureFuserHMT::Params::fusion2d is set
class NDTMatcherD2D_2D : public NDTMatcherD2D {
 public:
  NDTMatcherD2D_2D() { planar_ = true; }
};

// NDTMatcherP2D [upstream]: point cloud against an NDT map.  Not referenced by ndt_feature_graph itself (BASELINE
// config C3 names it); same knobs as NDTMatcherD2D.
class NDTMatcherP2D {
 public:
  int n_neighbours = 2;
  int ITR_MAX = 30;
  double DELTA_SCORE = 10e-3 * 0.1;
  bool step_control = true;
  ndtb_result last = {};

  // match(target map, source cloud, T, useInitialGuess) -> converged; T refined in place
  template <class Cloud, class Affine>
  bool match(NDTMap &target, const Cloud &source, Affine &T, bool useInitialGuess = false) {
    ndtb_ctx *c = ndtb::context();
    if (!c || !target.handle()) return false;
    double T0[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    if (useInitialGuess) std::memcpy(T0, ndtb::pose_data(T), sizeof T0);
    ndtb_params p;
    ndtb_default_params(&p);
    p.n_neighbours = n_neighbours, p.itr_max = ITR_MAX, p.delta_score = DELTA_SCORE, p.step_control = step_control;
    ndtb_result r;
    ndtb::last_status() = ndtb_p2d_match(c, target.handle(), reinterpret_cast<const float *>(source.points.data()),
                                         (int64_t)source.points.size(), NDTB_MEM_HOST, T0, &p, &r);
    if (ndtb::last_status() != NDTB_OK) return false;
    std::memcpy(ndtb::pose_data(T), r.T, sizeof r.T);
    last = r;
    return r.converged != 0;
  }
};

// NDTMatcherFeatureD2D(corr) [upstream]: D2D between cells with known correspondences (the FLIRT feature / odometry term of
// matchFusion, ndt_matcher_d2d_fusion.h:812,858,1016,1087).  With useFeat = useOdom = false — every shipped configuration —
// its maps are empty, its derivatives are zero and its line search returns 0: that trivial host behaviour is all the
// engine provides (out of scope as a kernel).
class NDTMatcherFeatureD2D {
 public:
  explicit NDTMatcherFeatureD2D(const std::vector<std::pair<int, int>> &corr) : corr_(corr) {}
  template <class Mat>
  double derivativesNDT(const std::vector<NDTCell *> &, const NDTMap &, Mat &score_gradient, Mat &Hessian, bool) {
    score_gradient.resize(6, 1), Hessian.resize(6, 6);
    score_gradient.setZero(), Hessian.setZero();
    return 0.0;
  }
  template <class Vec6>
  double lineSearchMT(Vec6 &, std::vector<NDTCell *> &, NDTMap &) { return 0.0; }
  size_t correspondences() const { return corr_.size(); }

 private:
  std::vector<std::pair<int, int>> corr_;
};

}  // namespace lslgeneric

namespace ndt_feature {

// matchFusion with the reference's exact parameter list (ndt_matcher_d2d_fusion.h:797-804, call site
// ndt_feature_fuser_hmt.cpp:356-357).  The NDT term runs on the engine; the feature maps are accepted for source
// compatibility and must be empty (useFeat = false, or no correspondences): a non-empty feature term is refused (returns
// false with last_status() == NDTB_ERR_ARG) rather than silently ignored.
template <class Affine, class Mat>
inline bool matchFusion(lslgeneric::NDTMap &targetNDT, lslgeneric::NDTMap &sourceNDT, lslgeneric::NDTMap & /*targetNDT_feat*/,
                        lslgeneric::NDTMap & /*sourceNDT_feat*/, const std::vector<std::pair<int, int>> &corr_feat, Affine &T,
                        const Mat &Tcov, bool useInitialGuess, bool useNDT, bool useFeat, bool step_control, int ITR_MAX = 30,
                        int n_neighbours = 2, double DELTA_SCORE = 10e-4, bool useSoftConstraints = true,
                        bool step_control_fusion = true, bool useTikhonovRegularization = false);

// matchFusion2d with the reference's parameter list (ndt_matcher_d2d_fusion.h:1159-1176, call site fuser_hmt.cpp:353)
template <class Affine>
inline bool matchFusion2d(lslgeneric::NDTMap &targetNDT, lslgeneric::NDTMap &sourceNDT, lslgeneric::NDTMap & /*targetNDT_feat*/,
                          lslgeneric::NDTMap & /*sourceNDT_feat*/, const std::vector<std::pair<int, int>> & /*corr_feat*/,
                          Affine &T, bool useInitialGuess, bool /*useNDT*/, bool /*useFeat*/, bool step_control,
                          int ITR_MAX = 30, int n_neighbours = 2, double DELTA_SCORE = 10e-4) {
  lslgeneric::NDTMatcherD2D_2D matcher_d2d_2d;
  matcher_d2d_2d.n_neighbours = n_neighbours;
  matcher_d2d_2d.step_control = step_control;
  matcher_d2d_2d.ITR_MAX = ITR_MAX;
  matcher_d2d_2d.DELTA_SCORE = DELTA_SCORE;
  return matcher_d2d_2d.match(targetNDT, sourceNDT, T, useInitialGuess);
}

// matchFusion with useNDT = true and no feature / odometry cell term (useFeat = false; the configuration of every
// offline driver, ndt_graph_offline.cpp:308): ndt_matcher_d2d_fusion.h:797-1155.
template <class Affine, class Mat>
inline bool matchFusion(lslgeneric::NDTMap &targetNDT, lslgeneric::NDTMap &sourceNDT, Affine &T, const Mat &Tcov,
                        bool /*useInitialGuess*/, bool step_control, int ITR_MAX, int n_neighbours, double DELTA_SCORE,
                        bool useSoftConstraints, bool useTikhonovRegularization) {
  ndtb_ctx *c = ndtb::context();
  if (!c || !targetNDT.handle() || !sourceNDT.handle()) return false;
  ndtb_params p;
  ndtb_default_params(&p);
  p.step_control = step_control, p.itr_max = ITR_MAX, p.n_neighbours = n_neighbours, p.delta_score = DELTA_SCORE;
  p.use_soft_constraints = useSoftConstraints, p.use_tikhonov = useTikhonovRegularization;
  double cov36[36];
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 6; j++) cov36[i * 6 + j] = Tcov(i, j);
  ndtb_result r;
  ndtb::last_status() = ndtb_fusion_match(c, targetNDT.handle(), sourceNDT.handle(), ndtb::pose_data(T), cov36, &p, &r);
  if (ndtb::last_status() != NDTB_OK) return false;
  std::memcpy(ndtb::pose_data(T), r.T, sizeof r.T);
  return r.converged != 0;
}

template <class Affine, class Mat>
inline bool matchFusion(lslgeneric::NDTMap &targetNDT, lslgeneric::NDTMap &sourceNDT, lslgeneric::NDTMap &, lslgeneric::NDTMap &,
                        const std::vector<std::pair<int, int>> &corr_feat, Affine &T, const Mat &Tcov, bool useInitialGuess,
                        bool useNDT, bool useFeat, bool step_control, int ITR_MAX, int n_neighbours, double DELTA_SCORE,
                        bool useSoftConstraints, bool /*step_control_fusion*/, bool useTikhonovRegularization) {
  if (!useNDT || (useFeat && !corr_feat.empty())) {  // feature / odometry-cell term: outside the engine's scope
    ndtb::last_status() = NDTB_ERR_ARG;
    return false;
  }
  return matchFusion(targetNDT, sourceNDT, T, Tcov, useInitialGuess, step_control, ITR_MAX, n_neighbours, DELTA_SCORE,
                     useSoftConstraints, useTikhonovRegularization);
}

// matchFusion2d (ndt_matcher_d2d_fusion.h:1159-1176): NDT-only, through NDTMatcherD2D_2D
template <class Affine>
inline bool matchFusion2d(lslgeneric::NDTMap &targetNDT, lslgeneric::NDTMap &sourceNDT, Affine &T, bool useInitialGuess,
                          bool step_control, int ITR_MAX = 30, int n_neighbours = 2, double DELTA_SCORE = 10e-4) {
  lslgeneric::NDTMatcherD2D_2D matcher_d2d_2d;
  matcher_d2d_2d.n_neighbours = n_neighbours;
  matcher_d2d_2d.step_control = step_control;
  matcher_d2d_2d.ITR_MAX = ITR_MAX;
  matcher_d2d_2d.DELTA_SCORE = DELTA_SCORE;
  return matcher_d2d_2d.match(targetNDT, sourceNDT, T, useInitialGuess);
}

// overlapNDTOccupancyScore(ref, mov, T) on the maps of two nodes (ndt_feature_node.h:213-252)
template <class Affine>
inline double overlapNDTOccupancyScore(lslgeneric::NDTMap &ref, lslgeneric::NDTMap &mov, const Affine &T) {
  ndtb_ctx *c = ndtb::context();
  double s = 1.0;
  if (c) ndtb::last_status() = ndtb_overlap_score(c, ref.handle(), mov.handle(), ndtb::pose_data(T), &s);
  return s;
}

}  // namespace ndt_feature

namespace ndtb {

// NDTFeatureLink's numeric payload (ndt_feature_link.h:9-56)
struct Link {
  size_t ref_idx = 0, mov_idx = 0;
  Eigen::Affine3d T;
  Eigen::MatrixXd cov_3d;
  double score = -1.0;
  bool converged = false;
};

// NDTFeatureGraph::updateLinksUsingNDTRegistration(links, nb_neighbours, keepScore) (ndt_feature_graph.cpp:347-353)
// with the serial loop over links turned into ONE batched launch; per link: match from link.T, covariance() only if
// the pose changed bitwise else 0.02*I (ndt_feature_graph.cpp:286-310), overlap score unless keepScore (:335-342).
class GraphRegistrar {
 public:
  explicit GraphRegistrar(const std::vector<lslgeneric::NDTMap *> &node_maps) : nodes_(node_maps) {}
  int updateLinksUsingNDTRegistration(std::vector<Link> &links, int nb_neighbours, bool keepScore) {
    ndtb_ctx *c = context();
    if (!c) return last_status();
    const size_t n = links.size();
    if (n == 0) return NDTB_OK;
    std::vector<const ndtb_map *> tg(n), sr(n);
    std::vector<double> T0(16 * n), cov(36 * n);
    std::vector<ndtb_result> res(n);
    for (size_t i = 0; i < n; i++) {
      tg[i] = nodes_[links[i].ref_idx]->handle(), sr[i] = nodes_[links[i].mov_idx]->handle();
      std::memcpy(&T0[16 * i], pose_data(links[i].T), 128);
    }
    ndtb_params p;
    ndtb_default_params(&p);
    p.n_neighbours = nb_neighbours;
    const int rc = ndtb_d2d_match_batch(c, (int64_t)n, tg.data(), sr.data(), T0.data(), &p, 1, NDTB_MEM_HOST, res.data(), cov.data());
    last_status() = rc;
    if (rc != NDTB_OK) return rc;
    for (size_t i = 0; i < n; i++) {
      std::memcpy(pose_data(links[i].T), res[i].T, 128);
      links[i].converged = res[i].converged != 0;
      links[i].cov_3d.resize(6, 6);
      for (int a = 0; a < 6; a++)
        for (int b = 0; b < 6; b++) links[i].cov_3d(a, b) = cov[36 * i + a * 6 + b];
    }
    if (!keepScore) {  // ndt_feature_graph.cpp:335-342, one launch over all links
      std::vector<double> sc(n);
      const int rs = ndtb_overlap_score_batch(c, (int64_t)n, tg.data(), sr.data(), res.data(), (int64_t)sizeof(ndtb_result),
                                              NDTB_MEM_HOST, NDTB_MEM_HOST, sc.data());
      last_status() = rs;
      if (rs != NDTB_OK) return rs;
      for (size_t i = 0; i < n; i++) links[i].score = sc[i];
    }
    return NDTB_OK;
  }

 private:
  std::vector<lslgeneric::NDTMap *> nodes_;
};

}  // namespace ndtb
