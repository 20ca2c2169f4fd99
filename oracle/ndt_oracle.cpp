/*
 * ndt_oracle.cpp — CPU ORACLE (test infrastructure, NOT product code).  See ndt_oracle.h.
 *
 * PARITY STATUS: "parity unpinned" for the perception_oru arithmetic (see header).
 *
 * What each block restates (reference file:line where the reference has the code,
 * "[upstream]" where the code lives in OrebroUniversity/perception_oru, unpinned, and is
 * restated from its published algorithm / SURVEY.md Appendix A):
 *
 *   grid index / binning      [upstream] LazyGrid::getIndexForPoint/addPoint, NDTMap::loadPointCloud
 *                             call sites ndt_feature_fuser_hmt.cpp:195-227, :87-94
 *   cell Gaussians            [upstream] NDTCell::computeGaussian(SAMPLE_VARIANCE)+rescaleCovariance
 *                             pinned by fixture invariants (eig ratio 1000, occ = n*ln1.5)
 *   derivativesNDT            [upstream] NDTMatcherD2D::derivativesNDT/computeDerivativesLocal/
 *                             update_gradient_hessian_local; call sites ndt_matcher_d2d_fusion.h:856,617,444
 *   lineSearchMT              ndt_matcher_d2d_fusion.h:390-793 (lineSearchMTFusion, NDT terms only)
 *   MoreThuente::cstep        [upstream] (MINPACK dcstep), SURVEY.md A4
 *   match                     ndt_matcher_d2d_fusion.h:797-1155 minus feature terms == upstream match()
 *   matchFusion soft/Tikhonov ndt_matcher_d2d_fusion.h:11-32, :37-385, :873-911, :1008-1023
 *   covariance                [upstream] NDTMatcherD2D::covariance, SURVEY.md A5 (least certain)
 *   overlap score             ndt_feature_node.h:213-252
 */
#include "ndt_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <unordered_map>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ------------------------------------------------------------------ small linear algebra
// 3x3 matrices are row-major double[9]; 6x6 row-major double[36].

inline void mat3_mul(const double *A, const double *B, double *C) {
  double t[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) t[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
  std::memcpy(C, t, sizeof t);
}
inline void mat3_mulT(const double *A, const double *B, double *C) {  // C = A * B^T
  double t[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      t[i * 3 + j] = A[i * 3] * B[j * 3] + A[i * 3 + 1] * B[j * 3 + 1] + A[i * 3 + 2] * B[j * 3 + 2];
  std::memcpy(C, t, sizeof t);
}
inline void mat3_vec(const double *A, const double *x, double *y) {
  double t0 = A[0] * x[0] + A[1] * x[1] + A[2] * x[2];
  double t1 = A[3] * x[0] + A[4] * x[1] + A[5] * x[2];
  double t2 = A[6] * x[0] + A[7] * x[1] + A[8] * x[2];
  y[0] = t0, y[1] = t1, y[2] = t2;
}
inline double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// Eigen's computeInverseAndDetWithCheck for 3x3: cofactor expansion along column 0,
// invertible iff |det| > dummy_precision (1e-12).
inline bool inv3_check(const double *m, double *inv, double &det) {
  double c00 = m[4] * m[8] - m[5] * m[7];
  double c10 = m[2] * m[7] - m[1] * m[8];
  double c20 = m[1] * m[5] - m[2] * m[4];
  det = c00 * m[0] + c10 * m[3] + c20 * m[6];
  if (!(std::fabs(det) > 1e-12)) return false;
  double id = 1.0 / det;
  inv[0] = c00 * id;
  inv[1] = c10 * id;
  inv[2] = c20 * id;
  inv[3] = (m[5] * m[6] - m[3] * m[8]) * id;
  inv[4] = (m[0] * m[8] - m[2] * m[6]) * id;
  inv[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  inv[6] = (m[3] * m[7] - m[4] * m[6]) * id;
  inv[7] = (m[1] * m[6] - m[0] * m[7]) * id;
  inv[8] = (m[0] * m[4] - m[1] * m[3]) * id;
  return true;
}

// Cyclic Jacobi for a symmetric n x n matrix (n <= 6). evals ascending, V columns = eigenvectors.
void jacobi_eig(int n, const double *Ain, double *evals, double *V) {
  double A[36];
  for (int i = 0; i < n * n; i++) A[i] = Ain[i];
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) V[i * n + j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 64; sweep++) {
    double off = 0, diag = 0;
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++) (i == j ? diag : off) += A[i * n + j] * A[i * n + j];
    if (off <= 1e-32 * diag || off == 0.0) break;
    for (int p = 0; p < n - 1; p++)
      for (int q = p + 1; q < n; q++) {
        double apq = A[p * n + q];
        if (apq == 0.0) continue;
        double app = A[p * n + p], aqq = A[q * n + q];
        double theta = (aqq - app) / (2.0 * apq);
        double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; k++) {  // A <- A J
          double akp = A[k * n + p], akq = A[k * n + q];
          A[k * n + p] = c * akp - s * akq;
          A[k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; k++) {  // A <- J^T A
          double apk = A[p * n + k], aqk = A[q * n + k];
          A[p * n + k] = c * apk - s * aqk;
          A[q * n + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; k++) {
          double vkp = V[k * n + p], vkq = V[k * n + q];
          V[k * n + p] = c * vkp - s * vkq;
          V[k * n + q] = s * vkp + c * vkq;
        }
      }
  }
  int order[6];
  for (int i = 0; i < n; i++) order[i] = i;
  std::sort(order, order + n, [&](int a, int b) { return A[a * n + a] < A[b * n + b]; });
  double Vt[36];
  for (int j = 0; j < n; j++) {
    evals[j] = A[order[j] * n + order[j]];
    for (int i = 0; i < n; i++) Vt[i * n + j] = V[i * n + order[j]];
  }
  for (int i = 0; i < n * n; i++) V[i] = Vt[i];
}

// x = A^{-1} b via LDL^T with symmetric diagonal pivoting (Eigen::LDLT semantics), n = 6.
void ldlt_solve6(const double *Ain, const double *b, double *x) {
  const int n = 6;
  double A[36];
  std::memcpy(A, Ain, sizeof A);
  int perm[6];
  for (int i = 0; i < n; i++) perm[i] = i;
  for (int k = 0; k < n; k++) {
    int p = k;
    double best = std::fabs(A[k * n + k]);
    for (int i = k + 1; i < n; i++)
      if (std::fabs(A[i * n + i]) > best) best = std::fabs(A[i * n + i]), p = i;
    if (p != k) {
      for (int j = 0; j < n; j++) std::swap(A[k * n + j], A[p * n + j]);
      for (int i = 0; i < n; i++) std::swap(A[i * n + k], A[i * n + p]);
      std::swap(perm[k], perm[p]);
    }
    double d = A[k * n + k];
    if (d == 0.0) continue;
    for (int i = k + 1; i < n; i++) A[i * n + k] /= d;  // L(i,k)
    for (int i = k + 1; i < n; i++)
      for (int j = k + 1; j <= i; j++) {
        A[i * n + j] -= A[i * n + k] * d * A[j * n + k];
        A[j * n + i] = A[i * n + j];
      }
  }
  double y[6];
  for (int i = 0; i < n; i++) y[i] = b[perm[i]];
  for (int i = 0; i < n; i++)
    for (int j = 0; j < i; j++) y[i] -= A[i * n + j] * y[j];
  for (int i = 0; i < n; i++) {
    double d = A[i * n + i];
    y[i] = (std::fabs(d) > std::numeric_limits<double>::min()) ? y[i] / d : 0.0;
  }
  for (int i = n - 1; i >= 0; i--)
    for (int j = i + 1; j < n; j++) y[i] -= A[j * n + i] * y[j];
  for (int i = 0; i < n; i++) x[perm[i]] = y[i];
}

// general 6x6 inverse, Gauss-Jordan with partial pivoting (Eigen MatrixXd::inverse() = PartialPivLU)
bool inv6(const double *Ain, double *Ainv) {
  const int n = 6;
  double M[6][12];
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) M[i][j] = Ain[i * n + j], M[i][n + j] = (i == j);
  for (int c = 0; c < n; c++) {
    int p = c;
    for (int r = c + 1; r < n; r++)
      if (std::fabs(M[r][c]) > std::fabs(M[p][c])) p = r;
    if (M[p][c] == 0.0) return false;
    if (p != c)
      for (int j = 0; j < 2 * n; j++) std::swap(M[p][j], M[c][j]);
    double id = 1.0 / M[c][c];
    for (int j = 0; j < 2 * n; j++) M[c][j] *= id;
    for (int r = 0; r < n; r++)
      if (r != c) {
        double f = M[r][c];
        if (f != 0.0)
          for (int j = 0; j < 2 * n; j++) M[r][j] -= f * M[c][j];
      }
  }
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) Ainv[i * n + j] = M[i][n + j];
  return true;
}

// ------------------------------------------------------------------ poses (column-major 4x4)
struct Pose {
  double R[9];  // row-major rotation
  double t[3];
};
Pose pose_from_cm(const double *T) {
  Pose P;
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) P.R[i * 3 + j] = T[j * 4 + i];
    P.t[i] = T[12 + i];
  }
  return P;
}
void pose_to_cm(const Pose &P, double *T) {
  for (int i = 0; i < 16; i++) T[i] = 0;
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) T[j * 4 + i] = P.R[i * 3 + j];
    T[12 + i] = P.t[i];
  }
  T[15] = 1;
}
Pose pose_identity() {
  Pose P;
  for (int i = 0; i < 9; i++) P.R[i] = (i % 4 == 0);
  P.t[0] = P.t[1] = P.t[2] = 0;
  return P;
}
// TR = Translation(p0,p1,p2) * AngleAxis(p3,X) * AngleAxis(p4,Y) * AngleAxis(p5,Z)
// (ndt_matcher_d2d_fusion.h:1036-1039, :558-561)
Pose pose_from_vec(const double *p) {
  double cx = std::cos(p[3]), sx = std::sin(p[3]);
  double cy = std::cos(p[4]), sy = std::sin(p[4]);
  double cz = std::cos(p[5]), sz = std::sin(p[5]);
  double Rx[9] = {1, 0, 0, 0, cx, -sx, 0, sx, cx};
  double Ry[9] = {cy, 0, sy, 0, 1, 0, -sy, 0, cy};
  double Rz[9] = {cz, -sz, 0, sz, cz, 0, 0, 0, 1};
  Pose P;
  double t[9];
  mat3_mul(Rx, Ry, t);
  mat3_mul(t, Rz, P.R);
  P.t[0] = p[0], P.t[1] = p[1], P.t[2] = p[2];
  return P;
}
Pose pose_mul(const Pose &A, const Pose &B) {  // A*B
  Pose C;
  mat3_mul(A.R, B.R, C.R);
  double t[3];
  mat3_vec(A.R, B.t, t);
  for (int i = 0; i < 3; i++) C.t[i] = t[i] + A.t[i];
  return C;
}

// ------------------------------------------------------------------ NDT cells and the LazyGrid map
const double EVAL_FACTOR = 1000.0;  // upstream NDTCell EVAL_FACTOR; fixture-confirmed

struct Gauss {  // what the matcher needs from an NDTCell
  double mean[3];
  double cov[9];
};

struct Cell {
  double mean[3] = {0, 0, 0};
  double cov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  int32_t N = 0;
  bool has_gaussian = false;
  float occ = 0.f;
  int32_t idx[3] = {0, 0, 0};
  std::vector<float> pts;  // NDTCell::points_ (xyz triples)
};

}  // namespace

struct orc_map {
  double cell[3];
  double center[3] = {0, 0, 0};
  double size_m[3] = {0, 0, 0};
  int32_t size[3] = {0, 0, 0};
  bool grid_ready = false;
  // NDTMap members
  bool guess_size = true;
  double centerx = 0, centery = 0, centerz = 0;
  double map_sizex = -1, map_sizey = -1, map_sizez = -1;
  bool is_first_load = true;
  // storage
  std::vector<int32_t> dense;                 // voxel -> cell id (or -1)
  std::unordered_map<int64_t, int32_t> sparse;  // used when the grid is too large for `dense`
  bool use_dense = true;
  std::vector<Cell> cells;
  std::vector<int32_t> update_set;
  std::vector<uint8_t> in_update;

  int64_t nvox() const { return (int64_t)size[0] * size[1] * size[2]; }
  int64_t lin(int x, int y, int z) const { return ((int64_t)x * size[1] + y) * size[2] + z; }
  bool inb(int x, int y, int z) const {
    return x >= 0 && y >= 0 && z >= 0 && x < size[0] && y < size[1] && z < size[2];
  }
  void set_grid(double cx, double cy, double cz, double sx, double sy, double sz) {
    center[0] = cx, center[1] = cy, center[2] = cz;
    size_m[0] = sx, size_m[1] = sy, size_m[2] = sz;
    for (int i = 0; i < 3; i++) size[i] = (int32_t)std::abs(std::ceil(size_m[i] / cell[i]));  // LazyGrid::setSize
    reset_storage();
  }
  void reset_storage() {
    cells.clear();
    update_set.clear();
    in_update.clear();
    sparse.clear();
    use_dense = nvox() <= ((int64_t)1 << 26);
    if (use_dense)
      dense.assign((size_t)nvox(), -1);
    else
      dense.clear();
    grid_ready = true;
  }
  // LazyGrid::getIndexForPoint: ind = floor((p - center)/cell + 0.5) + size/2.0, truncated to int
  bool index_of(double px, double py, double pz, int &ix, int &iy, int &iz) const {
    double v[3] = {std::floor((px - center[0]) / cell[0] + 0.5) + size[0] / 2.0,
                   std::floor((py - center[1]) / cell[1] + 0.5) + size[1] / 2.0,
                   std::floor((pz - center[2]) / cell[2] + 0.5) + size[2] / 2.0};
    int o[3];
    for (int i = 0; i < 3; i++) {
      if (!(v[i] > -2147483000.0 && v[i] < 2147483000.0)) return false;  // also rejects NaN
      o[i] = (int)v[i];
    }
    ix = o[0], iy = o[1], iz = o[2];
    return true;
  }
  int32_t find(int x, int y, int z) const {
    if (!inb(x, y, z)) return -1;
    if (use_dense) return dense[(size_t)lin(x, y, z)];
    auto it = sparse.find(lin(x, y, z));
    return it == sparse.end() ? -1 : it->second;
  }
  int32_t find_or_create(int x, int y, int z) {
    int64_t l = lin(x, y, z);
    int32_t id = use_dense ? dense[(size_t)l] : -1;
    if (!use_dense) {
      auto it = sparse.find(l);
      if (it != sparse.end()) id = it->second;
    }
    if (id >= 0) return id;
    id = (int32_t)cells.size();
    cells.emplace_back();
    cells.back().idx[0] = x, cells.back().idx[1] = y, cells.back().idx[2] = z;
    in_update.push_back(0);
    if (use_dense)
      dense[(size_t)l] = id;
    else
      sparse[l] = id;
    return id;
  }
  // LazyGrid::addPoint + update_set.insert
  bool add_point(const float *p) {
    if (std::isnan(p[0]) || std::isnan(p[1]) || std::isnan(p[2])) return false;
    int x, y, z;
    if (!index_of(p[0], p[1], p[2], x, y, z)) return false;
    if (!inb(x, y, z)) return false;
    int32_t id = find_or_create(x, y, z);
    Cell &c = cells[id];
    c.pts.push_back(p[0]), c.pts.push_back(p[1]), c.pts.push_back(p[2]);
    if (!in_update[id]) in_update[id] = 1, update_set.push_back(id);
    return true;
  }
  void gaussians(std::vector<Gauss> &out) const {  // cells with hasGaussian_, linear-index order
    std::vector<std::pair<int64_t, int32_t>> ord;
    for (size_t i = 0; i < cells.size(); i++)
      if (cells[i].has_gaussian) ord.push_back({lin(cells[i].idx[0], cells[i].idx[1], cells[i].idx[2]), (int32_t)i});
    std::sort(ord.begin(), ord.end());
    out.resize(ord.size());
    for (size_t k = 0; k < ord.size(); k++) {
      const Cell &c = cells[ord[k].second];
      std::memcpy(out[k].mean, c.mean, sizeof c.mean);
      std::memcpy(out[k].cov, c.cov, sizeof c.cov);
    }
  }
};

namespace {

// NDTCell::rescaleCovariance [upstream]: eig-decompose, any eval <= 0 -> no Gaussian,
// clamp evals to >= max/EVAL_FACTOR, rebuild cov.
void rescale_covariance(Cell &c) {
  double evals[3], V[9];
  jacobi_eig(3, c.cov, evals, V);
  if (evals[0] <= 0 || evals[1] <= 0 || evals[2] <= 0) {
    c.has_gaussian = false;
    return;
  }
  double maxe = std::max(evals[0], std::max(evals[1], evals[2]));
  bool recalc = false;
  for (int i = 0; i < 3; i++)
    if (maxe > evals[i] * EVAL_FACTOR) evals[i] = maxe / EVAL_FACTOR, recalc = true;
  if (recalc) {
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        double s = 0;
        for (int k = 0; k < 3; k++) s += V[i * 3 + k] * evals[k] * V[j * 3 + k];
        c.cov[i * 3 + j] = s;
      }
  }
  c.has_gaussian = true;
}

// NDTCell::computeGaussian(CELL_UPDATE_MODE_SAMPLE_VARIANCE, maxnumpoints, occupancy_limit) [upstream]
void compute_gaussian(Cell &c, uint32_t maxnumpoints, float occupancy_limit) {
  size_t n = c.pts.size() / 3;
  // occupancy: += n * log(0.6/0.4), clamped. Fixture-pinned: cells hold exactly k*ln1.5 incl. k=1,2.
  if (n > 0) {
    double lo = (double)n * std::log(0.6 / (1.0 - 0.6));
    float occ = c.occ + (float)lo;
    if (occ > occupancy_limit) occ = occupancy_limit;
    if (occ < -occupancy_limit) occ = -occupancy_limit;
    c.occ = occ;
  }
  // [upstream] after the occupancy update: if(occ<=0){ hasGaussian_=false; return; } -- the cell keeps its old
  // (N, mean, cov) and computeNDTCells moves its leftover points_ to conflictPoints.  Only reachable after the
  // free-space ray trace of addPointCloud has driven the occupancy negative.
  if (c.occ <= 0.f) {
    c.has_gaussian = false;
    c.pts.clear();
    return;
  }
  if ((!c.has_gaussian && n < 3) || n == 0) {
    c.pts.clear();
    return;
  }
  double msum[3] = {0, 0, 0};
  for (size_t i = 0; i < n; i++)
    for (int k = 0; k < 3; k++) msum[k] += (double)c.pts[3 * i + k];
  double mloc[3] = {msum[0] / (double)n, msum[1] / (double)n, msum[2] / (double)n};
  double csum[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (size_t i = 0; i < n; i++) {
    double d[3] = {(double)c.pts[3 * i] - mloc[0], (double)c.pts[3 * i + 1] - mloc[1], (double)c.pts[3 * i + 2] - mloc[2]};
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++) csum[a * 3 + b] += d[a] * d[b];
  }
  if (!c.has_gaussian) {
    for (int k = 0; k < 3; k++) c.mean[k] = mloc[k];
    for (int k = 0; k < 9; k++) c.cov[k] = csum[k] / (double)(n - 1);
    c.N = (int32_t)n;
    rescale_covariance(c);
  } else {
    // pairwise (Chan et al.) update of (N, mean, cov) with the new batch
    double N0 = (double)c.N, n1 = (double)n;
    double mS[3], cS[9];
    for (int k = 0; k < 3; k++) mS[k] = c.mean[k] * N0;
    for (int k = 0; k < 9; k++) cS[k] = c.cov[k] * (N0 - 1.0);
    double w = N0 / (n1 * (N0 + n1));
    double tv[3];
    for (int k = 0; k < 3; k++) tv[k] = (n1 / N0) * mS[k] - msum[k];
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++) cS[a * 3 + b] += csum[a * 3 + b] + w * tv[a] * tv[b];
    for (int k = 0; k < 3; k++) mS[k] += msum[k];
    double Nt = N0 + n1;
    for (int k = 0; k < 3; k++) c.mean[k] = mS[k] / Nt;
    for (int k = 0; k < 9; k++) c.cov[k] = cS[k] / (Nt - 1.0);
    if (Nt > (double)maxnumpoints) {
      // upstream rescales the running sums so that the cell "forgets": N capped at maxnumpoints
      Nt = (double)maxnumpoints;
    }
    c.N = (int32_t)Nt;
    rescale_covariance(c);
  }
  c.pts.clear();
}

// ------------------------------------------------------------------ derivativesNDT
struct Deriv {
  double score = 0;
  double g[6] = {0, 0, 0, 0, 0, 0};
  double H[36];
  int64_t pairs = 0;
  Deriv() {
    for (int i = 0; i < 36; i++) H[i] = 0;
  }
};

// computeDerivativesLocal [upstream]: per-source-cell Jest (3x6), Hest (18x6, blocks 3x1),
// Zest (3x18, blocks 3x3), ZHest (18x18, blocks 3x3).  Blocks are generated from the rotation
// generators G_r (v -> e_r x v):  J_r = G_r mu, Z_r = G_r C + C G_r^T,
// H_ab = G_a G_b mu, ZH_ab = G_a G_b C + G_a C G_b^T + G_b C G_a^T + C (G_a G_b)^T  for a<=b,
// mirrored for a>b (upstream fills the lower blocks by copying the upper ones).
struct Local {
  double J[3][3];      // J[r] = rotation column r (translation columns are identity)
  double Z[3][9];      // Z[r]
  double Hh[3][3][3];  // Hh[a][b] = 3-vector
  double ZH[3][3][9];
};
const double GEN[3][9] = {{0, 0, 0, 0, 0, -1, 0, 1, 0}, {0, 0, 1, 0, 0, 0, -1, 0, 0}, {0, -1, 0, 1, 0, 0, 0, 0, 0}};

void compute_local(const double *mu, const double *C, bool hess, Local &L) {
  for (int r = 0; r < 3; r++) {
    mat3_vec(GEN[r], mu, L.J[r]);
    double GC[9];
    mat3_mul(GEN[r], C, GC);
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) L.Z[r][i * 3 + j] = GC[i * 3 + j] + GC[j * 3 + i];
  }
  if (!hess) return;
  for (int a = 0; a < 3; a++)
    for (int b = a; b < 3; b++) {
      double GG[9], t1[9], t2[9], t3[9], t4[9], tmp[9];
      mat3_mul(GEN[a], GEN[b], GG);
      mat3_vec(GG, mu, L.Hh[a][b]);
      mat3_mul(GG, C, t1);
      mat3_mul(GEN[a], C, tmp);
      mat3_mulT(tmp, GEN[b], t2);
      mat3_mul(GEN[b], C, tmp);
      mat3_mulT(tmp, GEN[a], t3);
      mat3_mulT(C, GG, t4);
      for (int k = 0; k < 9; k++) L.ZH[a][b][k] = t1[k] + t2[k] + t3[k] + t4[k];
      if (b != a) {
        std::memcpy(L.Hh[b][a], L.Hh[a][b], sizeof L.Hh[a][b]);
        std::memcpy(L.ZH[b][a], L.ZH[a][b], sizeof L.ZH[a][b]);
      }
    }
}

// update_gradient_hessian_local [upstream]
inline void update_local(Deriv &D, const double *x, const double *B, double likelihood, const Local &L,
                         bool hess, double lfd2) {
  double xtB[3] = {x[0] * B[0] + x[1] * B[3] + x[2] * B[6], x[0] * B[1] + x[1] * B[4] + x[2] * B[7],
                   x[0] * B[2] + x[1] * B[5] + x[2] * B[8]};
  double xtBJ[6], xtBZBx[6] = {0, 0, 0, 0, 0, 0}, Q[6];
  double TMP1[3][3];  // rows r: xtB * Z_r * B
  for (int k = 0; k < 3; k++) xtBJ[k] = xtB[k];
  for (int r = 0; r < 3; r++) {
    xtBJ[3 + r] = dot3(xtB, L.J[r]);
    double t[3] = {xtB[0] * L.Z[r][0] + xtB[1] * L.Z[r][3] + xtB[2] * L.Z[r][6],
                   xtB[0] * L.Z[r][1] + xtB[1] * L.Z[r][4] + xtB[2] * L.Z[r][7],
                   xtB[0] * L.Z[r][2] + xtB[1] * L.Z[r][5] + xtB[2] * L.Z[r][8]};
    for (int k = 0; k < 3; k++) TMP1[r][k] = t[0] * B[k] + t[1] * B[3 + k] + t[2] * B[6 + k];
    xtBZBx[3 + r] = dot3(TMP1[r], x);
  }
  for (int k = 0; k < 6; k++) Q[k] = 2 * xtBJ[k] - xtBZBx[k];
  double factor = -(lfd2 / 2) * likelihood;
  for (int k = 0; k < 6; k++) D.g[k] += Q[k] * factor;
  if (!hess) return;

  // J^T B J (6x6), J = [I | J_r]
  double BJ[6][3];  // B * J_col
  for (int k = 0; k < 3; k++) BJ[k][0] = B[k], BJ[k][1] = B[3 + k], BJ[k][2] = B[6 + k];
  for (int r = 0; r < 3; r++) mat3_vec(B, L.J[r], BJ[3 + r]);
  auto Jcol = [&](int p, double *v) {
    if (p < 3) {
      v[0] = v[1] = v[2] = 0;
      v[p] = 1;
    } else {
      v[0] = L.J[p - 3][0], v[1] = L.J[p - 3][1], v[2] = L.J[p - 3][2];
    }
  };
  double Bx[3];
  mat3_vec(B, x, Bx);
  for (int p = 0; p < 6; p++) {
    double Jp[3];
    Jcol(p, Jp);
    for (int q = 0; q < 6; q++) {
      double Jq[3];
      Jcol(q, Jq);
      double JtBJ = dot3(Jp, BJ[q]);
      double xtBH = 0, xtBZhBx = 0, xtBZBJ_pq = 0, xtBZBJ_qp = 0, zbz = 0, zbzT = 0;
      if (p >= 3 && q >= 3) {
        xtBH = dot3(xtB, L.Hh[p - 3][q - 3]);
        double t[3];
        mat3_vec(L.ZH[p - 3][q - 3], Bx, t);
        xtBZhBx = dot3(xtB, t);
        // xtBZBZBx(p,q) = TMP1_p * Z_q * B * x
        double u[3];
        mat3_vec(L.Z[q - 3], Bx, u);
        zbz = dot3(TMP1[p - 3], u);
        mat3_vec(L.Z[p - 3], Bx, u);
        zbzT = dot3(TMP1[q - 3], u);
      }
      // _xtBZBJ.col(i) = (TMP1_i * Jest)^T  => _xtBZBJ(q,p) = TMP1_p . J_q
      if (p >= 3) xtBZBJ_qp = dot3(TMP1[p - 3], Jq);  // element (q,p) -> contributes to transpose at (p,q)
      if (q >= 3) xtBZBJ_pq = dot3(TMP1[q - 3], Jp);  // element (p,q)
      D.H[p * 6 + q] += factor * (2 * JtBJ + 2 * xtBH - xtBZhBx - 2 * xtBZBJ_qp - 2 * xtBZBJ_pq + zbz + zbzT -
                                  lfd2 * Q[p] * Q[q] / 2);
    }
  }
}

// LazyGrid::getClosestNDTCells order: 0, +1, -1, +2, -2 per axis (x outer, z inner) [upstream]
inline int nb_off(int t) { return (t % 2 == 0) ? t / 2 : -(t / 2); }

void derivatives_cells(const std::vector<Gauss> &src, const Pose &P, const orc_map &tgt, const orc_params &prm,
                       bool hess, Deriv &out, std::vector<double> *per_source_g = nullptr,
                       std::vector<double> *per_target_g = nullptr) {
  const int k = prm.n_neighbours;
  const int64_t ns = (int64_t)src.size();
  int nthreads = prm.n_threads > 1 ? prm.n_threads : 1;
  if (per_target_g) nthreads = 1;
  std::vector<Deriv> parts((size_t)nthreads);
#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads) if (nthreads > 1)
#endif
  {
    int tid = 0;
#ifdef _OPENMP
    tid = omp_get_thread_num();
#endif
    Deriv &D = parts[(size_t)tid];
    Local L;
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
    for (int64_t i = 0; i < ns; i++) {
      // pseudoTransformNDT / per-iteration cell move: mean <- T mean, cov <- R cov R^T
      double mu[3], C[9], tmp[9];
      mat3_vec(P.R, src[(size_t)i].mean, mu);
      for (int a = 0; a < 3; a++) mu[a] += P.t[a];
      mat3_mul(P.R, src[(size_t)i].cov, tmp);
      mat3_mulT(tmp, P.R, C);
      compute_local(mu, C, hess, L);
      int ix, iy, iz;
      if (!tgt.index_of((double)(float)mu[0], (double)(float)mu[1], (double)(float)mu[2], ix, iy, iz)) continue;
      double g_before[6] = {0, 0, 0, 0, 0, 0};
      if (per_source_g)
        for (int a = 0; a < 6; a++) g_before[a] = D.g[a];
      for (int xx = 1; xx < 2 * k + 2; xx++) {
        int X = ix + nb_off(xx);
        for (int yy = 1; yy < 2 * k + 2; yy++) {
          int Y = iy + nb_off(yy);
          for (int zz = 1; zz < 2 * k + 2; zz++) {
            int Z = iz + nb_off(zz);
            int32_t id = tgt.find(X, Y, Z);
            if (id < 0) continue;
            const Cell &tc = tgt.cells[(size_t)id];
            if (!tc.has_gaussian) continue;
            double x[3] = {mu[0] - tc.mean[0], mu[1] - tc.mean[1], mu[2] - tc.mean[2]};
            double CS[9], B[9], det;
            for (int a = 0; a < 9; a++) CS[a] = tc.cov[a] + C[a];
            if (!inv3_check(CS, B, det)) continue;
            double Bx[3];
            mat3_vec(B, x, Bx);
            double l = dot3(x, Bx);
            if (l * 0 != 0) continue;
            double sh = -prm.lfd1 * std::exp(-prm.lfd2 * l / 2);
            double gt_before[6] = {0, 0, 0, 0, 0, 0};
            if (per_target_g)
              for (int a = 0; a < 6; a++) gt_before[a] = D.g[a];
            update_local(D, x, B, sh, L, hess, prm.lfd2);
            if (per_target_g)
              for (int a = 0; a < 6; a++) (*per_target_g)[(size_t)id * 6 + a] += D.g[a] - gt_before[a];
            D.score += sh;
            D.pairs++;
          }
        }
      }
      if (per_source_g)
        for (int a = 0; a < 6; a++) (*per_source_g)[(size_t)i * 6 + a] = D.g[a] - g_before[a];
    }
  }
  out = Deriv();
  for (auto &D : parts) {
    out.score += D.score;
    out.pairs += D.pairs;
    for (int a = 0; a < 6; a++) out.g[a] += D.g[a];
    if (hess)
      for (int a = 0; a < 36; a++) out.H[a] += D.H[a];
  }
}

// ------------------------------------------------------------------ More-Thuente
inline double mt_absmax(double a, double b, double c) { return std::max(std::fabs(a), std::max(std::fabs(b), std::fabs(c))); }

// MoreThuente::cstep [upstream] == MINPACK dcstep
int cstep(double &stx, double &fx, double &dx, double &sty, double &fy, double &dy, double &stp, double fp, double dp,
          bool &brackt, double stmin, double stmax) {
  int info = 0;
  if ((brackt && ((stp <= std::min(stx, sty)) || (stp >= std::max(stx, sty)))) || (dx * (stp - stx) >= 0.0) ||
      (stmax < stmin))
    return info;
  double sgnd = dp * (dx / std::fabs(dx));
  bool bound;
  double theta, s, gamma, p, q, r, stpc, stpq, stpf;
  if (fp > fx) {
    info = 1;
    bound = true;
    theta = 3 * (fx - fp) / (stp - stx) + dx + dp;
    s = mt_absmax(theta, dx, dp);
    gamma = s * std::sqrt(((theta / s) * (theta / s)) - (dx / s) * (dp / s));
    if (stp < stx) gamma = -gamma;
    p = (gamma - dx) + theta;
    q = ((gamma - dx) + gamma) + dp;
    r = p / q;
    stpc = stx + r * (stp - stx);
    stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2) * (stp - stx);
    if (std::fabs(stpc - stx) < std::fabs(stpq - stx))
      stpf = stpc;
    else
      stpf = stpc + (stpq - stpc) / 2;
    brackt = true;
  } else if (sgnd < 0.0) {
    info = 2;
    bound = false;
    theta = 3 * (fx - fp) / (stp - stx) + dx + dp;
    s = mt_absmax(theta, dx, dp);
    gamma = s * std::sqrt(((theta / s) * (theta / s)) - (dx / s) * (dp / s));
    if (stp > stx) gamma = -gamma;
    p = (gamma - dp) + theta;
    q = ((gamma - dp) + gamma) + dx;
    r = p / q;
    stpc = stp + r * (stx - stp);
    stpq = stp + (dp / (dp - dx)) * (stx - stp);
    if (std::fabs(stpc - stp) > std::fabs(stpq - stp))
      stpf = stpc;
    else
      stpf = stpq;
    brackt = true;
  } else if (std::fabs(dp) < std::fabs(dx)) {
    info = 3;
    bound = true;
    theta = 3 * (fx - fp) / (stp - stx) + dx + dp;
    s = mt_absmax(theta, dx, dp);
    gamma = s * std::sqrt(std::max(0.0, (theta / s) * (theta / s) - (dx / s) * (dp / s)));
    if (stp > stx) gamma = -gamma;
    p = (gamma - dp) + theta;
    q = (gamma + (dx - dp)) + gamma;
    r = p / q;
    if ((r < 0.0) && (gamma != 0.0))
      stpc = stp + r * (stx - stp);
    else if (stp > stx)
      stpc = stmax;
    else
      stpc = stmin;
    stpq = stp + (dp / (dp - dx)) * (stx - stp);
    if (brackt) {
      stpf = (std::fabs(stp - stpc) < std::fabs(stp - stpq)) ? stpc : stpq;
    } else {
      stpf = (std::fabs(stp - stpc) > std::fabs(stp - stpq)) ? stpc : stpq;
    }
  } else {
    info = 4;
    bound = false;
    if (brackt) {
      theta = 3 * (fp - fy) / (sty - stp) + dy + dp;
      s = mt_absmax(theta, dy, dp);
      gamma = s * std::sqrt(((theta / s) * (theta / s)) - (dy / s) * (dp / s));
      if (stp > sty) gamma = -gamma;
      p = (gamma - dp) + theta;
      q = ((gamma - dp) + gamma) + dy;
      r = p / q;
      stpc = stp + r * (sty - stp);
      stpf = stpc;
    } else if (stp > stx)
      stpf = stmax;
    else
      stpf = stmin;
  }
  if (fp > fx) {
    sty = stp, fy = fp, dy = dp;
  } else {
    if (sgnd < 0.0) sty = stx, fy = fx, dy = dx;
    stx = stp, fx = fp, dx = dp;
  }
  stpf = std::min(stmax, stpf);
  stpf = std::max(stmin, stpf);
  stp = stpf;
  if (brackt && bound) {
    if (sty > stx)
      stp = std::min(stx + 0.66 * (sty - stx), stp);
    else
      stp = std::max(stx + 0.66 * (sty - stx), stp);
  }
  return info;
}

// soft odometry prior (ndt_matcher_d2d_fusion.h:11-32); C = Tcov^-1 row-major 6x6
double maha_score(const double *x, const double *C) {
  double s = 0;
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 6; j++) s += x[i] * C[i * 6 + j] * x[j];
  return s;
}
void maha_hessian(const double *C, double *H) {
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 6; j++) H[i * 6 + j] = C[j * 6 + i] + C[i * 6 + j];
}
void maha_gradient(const double *x, const double *C, double *g) {
  double H[36];
  maha_hessian(C, H);
  for (int i = 0; i < 6; i++) {
    g[i] = 0;
    for (int j = 0; j < 6; j++) g[i] += H[i * 6 + j] * x[j];
  }
}

struct Counters {
  int hess = 0, grad = 0;
};

// lineSearchMT == ndt_matcher_d2d_fusion.h:390-793 with the feature terms removed.
// `soft`: lineSearchMTFusionTcov variant (:37-385) adding the Mahalanobis prior on X=localpose.
double line_search_mt(double *increment, const std::vector<Gauss> &src, const Pose &Pcur, const orc_map &tgt,
                      const orc_params &prm, Counters &cnt, bool soft = false, const double *localpose = nullptr,
                      const double *Cq = nullptr) {
  double stp = 1.0;
  const double recoverystep = 0.1, ftol = 0.11111, gtol = 0.99999, stpmax = 4.0, stpmin = 0.001, xtol = 0.01;
  const int maxfev = 40;
  double dginit = 0.0;
  int info = 0, infoc = 1;
  double X[6] = {0, 0, 0, 0, 0, 0};
  if (soft)
    for (int i = 0; i < 6; i++) X[i] = localpose[i];

  Deriv D;
  derivatives_cells(src, Pcur, tgt, prm, false, D);
  cnt.grad++;
  double score_init = D.score;
  double gh[6];
  for (int i = 0; i < 6; i++) gh[i] = D.g[i];
  if (soft) {
    score_init += maha_score(X, Cq);
    double gm[6];
    maha_gradient(X, Cq, gm);
    for (int i = 0; i < 6; i++) gh[i] += gm[i];
  }
  for (int i = 0; i < 6; i++) dginit += increment[i] * gh[i];
  if (dginit >= 0.0) {
    for (int i = 0; i < 6; i++) increment[i] = -increment[i];
    dginit = -dginit;
    if (dginit >= 0.0) return recoverystep;
  }
  bool brackt = false, stage1 = true;
  int nfev = 0;
  double dgtest = ftol * dginit;
  double width = stpmax - stpmin, width1 = 2 * width;
  double finit = score_init;
  double stx = 0.0, fx = finit, dgx = dginit, sty = 0.0, fy = finit, dgy = dginit;
  double stmin, stmax, fm, fxm, fym, dgm, dgxm, dgym;
  while (1) {
    if (brackt) {
      stmin = std::min(stx, sty);
      stmax = std::max(stx, sty);
    } else {
      stmin = stx;
      stmax = stp + 4 * (stp - stx);
    }
    stp = std::max(stp, stpmin);
    stp = std::min(stp, stpmax);
    if ((brackt && ((stp <= stmin) || (stp >= stmax))) || (nfev >= maxfev - 1) || (infoc == 0) ||
        (brackt && (stmax - stmin <= xtol * stmax)))
      stp = stx;
    double pincr[6];
    for (int i = 0; i < 6; i++) pincr[i] = stp * increment[i];
    if (soft)
      for (int i = 0; i < 6; i++) X[i] += pincr[i];  // (sic) accumulates across evaluations, :183
    Pose ps = pose_from_vec(pincr);
    Pose Pe = pose_mul(ps, Pcur);
    derivatives_cells(src, Pe, tgt, prm, false, D);
    cnt.grad++;
    double f = D.score;
    for (int i = 0; i < 6; i++) gh[i] = D.g[i];
    if (soft) {
      f += maha_score(X, Cq);
      double gm[6];
      maha_gradient(X, Cq, gm);
      for (int i = 0; i < 6; i++) gh[i] += gm[i];
    }
    double dg = 0;
    for (int i = 0; i < 6; i++) dg += increment[i] * gh[i];
    nfev++;
    double ftest1 = finit + stp * dgtest;
    if ((brackt && ((stp <= stmin) || (stp >= stmax))) || (infoc == 0)) info = 6;
    if ((stp == stpmax) && (f <= ftest1) && (dg <= dgtest)) info = 5;
    if ((stp == stpmin) && ((f > ftest1) || (dg >= dgtest))) info = 4;
    if (nfev >= maxfev) info = 3;
    if (brackt && (stmax - stmin <= xtol * stmax)) info = 2;
    bool sufficient = (f <= ftest1);
    if (sufficient && (std::fabs(dg) <= gtol * (-dginit))) info = 1;
    if (info != 0) {
      if (info != 1) stp = recoverystep;
      return stp;
    }
    if (stage1 && (f <= ftest1) && (dg >= std::min(ftol, gtol) * dginit)) stage1 = false;
    if (stage1 && (f <= fx) && (f > ftest1)) {
      fm = f - stp * dgtest;
      fxm = fx - stx * dgtest;
      fym = fy - sty * dgtest;
      dgm = dg - dgtest;
      dgxm = dgx - dgtest;
      dgym = dgy - dgtest;
      infoc = cstep(stx, fxm, dgxm, sty, fym, dgym, stp, fm, dgm, brackt, stmin, stmax);
      fx = fxm + stx * dgtest;
      fy = fym + sty * dgtest;
      dgx = dgxm + dgtest;
      dgy = dgym + dgtest;
    } else {
      infoc = cstep(stx, fx, dgx, sty, fy, dgy, stp, f, dg, brackt, stmin, stmax);
    }
    if (brackt) {
      if (std::fabs(sty - stx) >= 0.66 * width1) stp = stx + 0.5 * (sty - stx);
      width1 = width;
      width = std::fabs(sty - stx);
    }
  }
}

// NDTMatcherD2D::match == matchFusion (ndt_matcher_d2d_fusion.h:797-1155) with useNDT only.
// fusion != nullptr adds the soft-constraint / Tikhonov terms of matchFusion.
struct Fusion {
  const double *Tcov;  // 6x6 row-major
};

int match_cells(const orc_map &tgt, const std::vector<Gauss> &src, const double *T0, const orc_params &prm,
                const Fusion *fusion, orc_result &res);

int match_impl(const orc_map &tgt, const orc_map &srcmap, const double *T0, const orc_params &prm, const Fusion *fusion,
               orc_result &res) {
  std::vector<Gauss> src;
  srcmap.gaussians(src);
  return match_cells(tgt, src, T0, prm, fusion, res);
}

// NDTMatcherP2D [upstream, no call site in the reference: parity unpinned, defined here].  Point-to-distribution NDT
// (Magnusson): score = sum_points sum_{cells in the (2k+1)^3 neighbourhood} -lfd1 exp(-lfd2/2 x^T S^-1 x), x = T p - m.
// Its gradient / Hessian are exactly the D2D expressions with a zero source covariance (every Z term vanishes), so a
// point is handed to the D2D machinery as a cell with mean p and covariance 0; the Newton / More-Thuente driver is the
// one of NDTMatcherD2D::match.
void points_as_cells(const float *pts, int64_t n, std::vector<Gauss> &out) {
  out.clear();
  out.reserve((size_t)n);
  for (int64_t i = 0; i < n; i++) {
    const float *p = pts + 4 * i;
    if (std::isnan(p[0]) || std::isnan(p[1]) || std::isnan(p[2])) continue;
    Gauss g;
    g.mean[0] = p[0], g.mean[1] = p[1], g.mean[2] = p[2];
    for (int k = 0; k < 9; k++) g.cov[k] = 0.0;
    out.push_back(g);
  }
}

int match_cells(const orc_map &tgt, const std::vector<Gauss> &src, const double *T0, const orc_params &prm,
                const Fusion *fusion, orc_result &res) {
  Pose T = pose_from_cm(T0), Tbest = T;
  const Pose Tinit = T;
  double score_best = fusion ? std::numeric_limits<double>::max() : (double)2147483647;  // INT_MAX upstream
  int itr = 0;
  bool convergence = false, ret = true;
  Counters cnt;
  double Q[36];
  double pose_local[6] = {0, 0, 0, 0, 0, 0};
  double x0[6] = {0, 0, 0, 0, 0, 0};
  const bool soft = fusion && prm.use_soft_constraints;
  const bool tik = fusion && prm.use_tikhonov;
  if (fusion) {
    if (!inv6(fusion->Tcov, Q)) return -2;
  }
  int exit_code = 0;
  double score_here = 0;
  Deriv D;
  // incremental mode (the reference's own data flow): nextNDT = pseudoTransformNDT(T) once (:840), every accepted increment
  // moves that list in place (:1047-1056), derivatives and line searches are evaluated on it with no further pose
  const bool inc = prm.incremental_cells != 0;
  std::vector<Gauss> cur;
  const Pose Pid = pose_identity();
  auto move_list = [&](const Pose &M) {
    for (Gauss &c : cur) {
      double mu[3], tmp[9], C[9];
      mat3_vec(M.R, c.mean, mu);
      for (int a = 0; a < 3; a++) c.mean[a] = mu[a] + M.t[a];
      mat3_mul(M.R, c.cov, tmp);
      mat3_mulT(tmp, M.R, C);
      std::memcpy(c.cov, C, sizeof C);
    }
  };
  if (inc) {
    cur = src;
    move_list(T);
  }
  const std::vector<Gauss> &cells = inc ? cur : src;
  while (!convergence) {
    derivatives_cells(cells, inc ? Pid : T, tgt, prm, true, D);
    cnt.hess++;
    score_here = D.score;
    double g[6], H[36];
    for (int i = 0; i < 6; i++) g[i] = D.g[i];
    for (int i = 0; i < 36; i++) H[i] = D.H[i];
    if (soft) {
      score_here += maha_score(pose_local, Q);
      double Hm[36], gm[6];
      maha_hessian(Q, Hm);
      maha_gradient(pose_local, Q, gm);
      for (int i = 0; i < 36; i++) H[i] += Hm[i];
      for (int i = 0; i < 6; i++) g[i] += gm[i];
    }
    if (tik) {
      // :894-911  g <- H^T P g + Q x0 ; H <- H^T P H + Q ; x0 = vec(force2d(T * Tinit^-1))
      // x0 uses eulerAngles of a yaw-only rotation => (x, y, 0, 0, 0, yaw)
      Pose Ti;  // Tinit^-1
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) Ti.R[i * 3 + j] = Tinit.R[j * 3 + i];
      double tt[3];
      mat3_vec(Ti.R, Tinit.t, tt);
      for (int i = 0; i < 3; i++) Ti.t[i] = -tt[i];
      Pose X0 = pose_mul(T, Ti);
      double T16[16];
      pose_to_cm(X0, T16);
      double yaw = orc_robust_yaw(T16);
      x0[0] = X0.t[0], x0[1] = X0.t[1], x0[2] = 0, x0[3] = 0, x0[4] = 0, x0[5] = yaw;
      double Hn[36], gn[6];
      for (int i = 0; i < 6; i++) {
        gn[i] = 0;
        for (int k = 0; k < 6; k++) gn[i] += H[k * 6 + i] * g[k] + Q[i * 6 + k] * x0[k];
        for (int j = 0; j < 6; j++) {
          double s = 0;
          for (int k = 0; k < 6; k++) s += H[k * 6 + i] * H[k * 6 + j];
          Hn[i * 6 + j] = s + Q[i * 6 + j];
        }
      }
      std::memcpy(H, Hn, sizeof H);
      std::memcpy(g, gn, sizeof g);
      score_here += maha_score(x0, Q);
    }
    if (prm.planar) {
      // NDTMatcherD2D_2D [upstream] (reached through matchFusion2d, ndt_matcher_d2d_fusion.h:1159-1176): the same
      // Newton / More-Thuente loop on (x, y, yaw) only.  Restated as the 6-DoF system with z, roll and pitch decoupled
      // (zero gradient, unit diagonal): their increments are exactly 0 and the 3x3 (x, y, yaw) block is solved and
      // eigen-regularised exactly as the full system would be.
      const int drop[3] = {2, 3, 4};
      for (int d = 0; d < 3; d++) {
        g[drop[d]] = 0.0;
        for (int j = 0; j < 6; j++) H[drop[d] * 6 + j] = H[j * 6 + drop[d]] = 0.0;
        H[drop[d] * 6 + drop[d]] = 1.0;
      }
    }
    double scg[6];
    for (int i = 0; i < 6; i++) scg[i] = g[i];
    if (score_here < score_best) {
      Tbest = T;
      score_best = score_here;
    }
    // eigen-regularisation :922-940
    double evals[6], evecs[36];
    jacobi_eig(6, H, evals, evecs);
    double minC = evals[0], maxC = evals[5];
    double gnorm = 0;
    for (int i = 0; i < 6; i++) gnorm += g[i] * g[i];
    gnorm = std::sqrt(gnorm);
    if (minC < 0) {
      if (prm.regularize || fusion) {
        double reg = gnorm;
        reg = reg + minC > 0 ? reg : 0.001 * maxC - minC;
        for (int i = 0; i < 6; i++) evals[i] += reg;
        for (int i = 0; i < 6; i++)
          for (int j = 0; j < 6; j++) {
            double s = 0;
            for (int k = 0; k < 6; k++) s += evecs[i * 6 + k] * evals[k] * evecs[j * 6 + k];
            H[i * 6 + j] = s;
          }
      } else {
        if (score_here > score_best) T = Tbest;
        exit_code = 4;
        goto done_early;
      }
    }
    if (gnorm <= prm.delta_score) {
      if (score_here > score_best) T = Tbest;
      exit_code = 1;
      goto done_early;
    }
    {
      double incr[6], ng[6];
      ldlt_solve6(H, g, ng);
      for (int i = 0; i < 6; i++) incr[i] = -ng[i];
      double dginit = 0;
      for (int i = 0; i < 6; i++) dginit += incr[i] * scg[i];
      if (dginit > 0) {
        if (score_here > score_best) T = Tbest;
        exit_code = 2;
        goto done_early;
      }
      double step = 1;
      if (prm.step_control) {
        if (soft) {
          // :1008-1010 result is overwritten by :1012-1023, but `increment` may be flipped in place
          (void)line_search_mt(incr, cells, inc ? Pid : T, tgt, prm, cnt, true, pose_local, Q);
        }
        step = line_search_mt(incr, cells, inc ? Pid : T, tgt, prm, cnt);
        // :1018-1023 with step_size_feat == 0  =>  step_size = max(step_ndt, 0)
        if (fusion) step = std::max(step, 0.0);
      }
      for (int i = 0; i < 6; i++) incr[i] *= step;
      Pose TR = pose_from_vec(incr);
      T = pose_mul(TR, T);
      if (inc) move_list(TR);
      for (int i = 0; i < 6; i++) pose_local[i] += incr[i];
      double nrm = 0;
      for (int i = 0; i < 6; i++) nrm += incr[i] * incr[i];
      nrm = std::sqrt(nrm);
      if (itr > 0) convergence = nrm < prm.delta_score;
      if (itr > prm.itr_max) {
        convergence = true;
        ret = false;
        exit_code = 3;
      }
      itr++;
    }
  }
  {
    derivatives_cells(cells, inc ? Pid : T, tgt, prm, false, D);
    cnt.grad++;
    score_here = D.score;
    if (soft) score_here += maha_score(pose_local, Q);
    if (tik) score_here += maha_score(x0, Q);
    if (score_here > score_best) T = Tbest;
  }
done_early:
  pose_to_cm(T, res.T);
  res.score = score_here;
  res.score_best = score_best;
  res.converged = ret ? 1 : 0;
  res.iterations = itr;
  res.n_hess_passes = cnt.hess;
  res.n_grad_passes = cnt.grad;
  res.exit_code = exit_code;
  res.pose_changed = 0;
  for (int i = 0; i < 16; i++)
    if (res.T[i] != T0[i]) res.pose_changed = 1;
  return 0;
}

// NDTMatcherD2D::covariance [upstream, least certain recall -- SURVEY.md A5]. Definition fixed here:
//   H     = Hessian of the score at T
//   rows  = per-source-cell gradient sums g_i (cell i vs. its target neighbourhood) and
//           per-target-cell gradient sums g_t (all pairs that hit target cell t)
//   cov   = H^-1 * (sigma^2 * sum_c g_c g_c^T) * H^-1,  sigma = 0.03
int covariance_impl(const orc_map &tgt, const orc_map &srcmap, const double *T16, const orc_params &prm, double *cov36) {
  std::vector<Gauss> src;
  srcmap.gaussians(src);
  Pose T = pose_from_cm(T16);
  Deriv D;
  orc_params p1 = prm;
  p1.n_threads = 1;
  std::vector<double> gs(src.size() * 6, 0.0), gt(tgt.cells.size() * 6, 0.0);
  derivatives_cells(src, T, tgt, p1, true, D, &gs, &gt);
  double JtJ[36];
  for (int i = 0; i < 36; i++) JtJ[i] = 0;
  auto acc = [&](const std::vector<double> &rows) {
    for (size_t r = 0; r < rows.size() / 6; r++)
      for (int a = 0; a < 6; a++)
        for (int b = 0; b < 6; b++) JtJ[a * 6 + b] += rows[r * 6 + a] * rows[r * 6 + b];
  };
  acc(gs);
  acc(gt);
  const double sigmaS = 0.03 * 0.03;
  double Hinv[36];
  if (!inv6(D.H, Hinv)) return -1;
  double tmp[36];
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 6; j++) {
      double s = 0;
      for (int k = 0; k < 6; k++) s += Hinv[i * 6 + k] * sigmaS * JtJ[k * 6 + j];
      tmp[i * 6 + j] = s;
    }
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 6; j++) {
      double s = 0;
      for (int k = 0; k < 6; k++) s += tmp[i * 6 + k] * Hinv[k * 6 + j];
      cov36[i * 6 + j] = s;
    }
  return 0;
}

}  // namespace

// ====================================================================== C API
extern "C" {

void orc_default_params(orc_params *p) {
  p->n_neighbours = 2;
  p->itr_max = 30;
  p->step_control = 1;
  p->regularize = 1;
  p->delta_score = 10e-3 * 0.1;  // upstream init(): DELTA_SCORE = 10e-3*current_resolution, current_resolution=0.1
  p->lfd1 = 1;
  p->lfd2 = 0.05;
  p->use_soft_constraints = 0;
  p->use_tikhonov = 0;
  p->n_threads = 1;
  p->planar = 0;
  p->incremental_cells = 0;
  p->reserved_ = 0;
}

orc_map *orc_map_create(double cx, double cy, double cz) {
  orc_map *m = new orc_map();
  m->cell[0] = cx, m->cell[1] = cy, m->cell[2] = cz;
  return m;
}
void orc_map_destroy(orc_map *m) { delete m; }

void orc_map_guess_size(orc_map *m, double cx, double cy, double cz, double sx, double sy, double sz) {
  m->guess_size = false;
  // NDTMap::guessSize takes floats upstream
  m->centerx = (float)cx, m->centery = (float)cy, m->centerz = (float)cz;
  m->map_sizex = (float)sx, m->map_sizey = (float)sy, m->map_sizez = (float)sz;
}
void orc_map_set_map_size(orc_map *m, double sx, double sy, double sz) {
  m->map_sizex = (float)sx, m->map_sizey = (float)sy, m->map_sizez = (float)sz;
}
void orc_map_initialize(orc_map *m, double cx, double cy, double cz, double sx, double sy, double sz) {
  m->is_first_load = false;
  m->guess_size = false;
  m->centerx = cx, m->centery = cy, m->centerz = cz;
  m->map_sizex = sx, m->map_sizey = sy, m->map_sizez = sz;
  m->set_grid(cx, cy, cz, sx, sy, sz);
}

int64_t orc_map_load_point_cloud(orc_map *m, const float *pts, int64_t n, double range_limit) {
  auto skip = [&](const float *p) {
    if (std::isnan(p[0]) || std::isnan(p[1]) || std::isnan(p[2])) return true;
    if (range_limit > 0) {
      double d = std::sqrt((double)p[0] * p[0] + (double)p[1] * p[1] + (double)p[2] * p[2]);
      if (d > range_limit) return true;
    }
    return false;
  };
  if (m->guess_size) {
    double cen[3] = {0, 0, 0};
    int64_t npts = 0;
    for (int64_t i = 0; i < n; i++) {
      const float *p = pts + 4 * i;
      if (skip(p)) continue;
      cen[0] += p[0], cen[1] += p[1], cen[2] += p[2];
      npts++;
    }
    if (npts == 0) return 0;
    for (int k = 0; k < 3; k++) cen[k] /= (double)npts;
    double maxDist = 0, maxz = -1000, minz = 10000;
    for (int64_t i = 0; i < n; i++) {
      const float *p = pts + 4 * i;
      if (skip(p)) continue;
      double d[3] = {cen[0] - p[0], cen[1] - p[1], cen[2] - p[2]};
      double dist = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
      maxDist = dist > maxDist ? dist : maxDist;
      maxz = d[2] > maxz ? d[2] : maxz;
      minz = d[2] < minz ? d[2] : minz;
    }
    if (m->map_sizex > 0 && m->map_sizey > 0 && m->map_sizez > 0)
      m->set_grid(cen[0], cen[1], cen[2], m->map_sizex, m->map_sizey, m->map_sizez);
    else
      m->set_grid(cen[0], cen[1], cen[2], 4 * maxDist, 4 * maxDist, 3 * (maxz - minz));
  } else {
    m->set_grid(m->centerx, m->centery, m->centerz, m->map_sizex, m->map_sizey, m->map_sizez);
  }
  m->is_first_load = false;
  int64_t added = 0;
  for (int64_t i = 0; i < n; i++) {
    const float *p = pts + 4 * i;
    if (skip(p)) continue;
    if (m->add_point(p)) added++;
  }
  return added;
}

int64_t orc_map_add_points(orc_map *m, const float *pts, int64_t n) {
  if (!m->grid_ready) return -1;
  int64_t added = 0;
  for (int64_t i = 0; i < n; i++)
    if (m->add_point(pts + 4 * i)) added++;
  return added;
}

// NDTMap::addPointCloud(origin, pc, classifierTh, maxz, sensor_noise, occupancy_limit) [upstream, REFACTORED branch]
// with LazyGrid::traceLine: the free-space ray trace + occupancy update the fuser runs on the node map
// (call sites ndt_feature_fuser_hmt.cpp:92 and :485).  Per point, in cloud order:
//   diff = p - origin, l = |diff| (skip NaN points and l > 200 m);
//   traceLine: skip the ray if p.z > maxz; N = int(l / min cell size); N-2 samples origin + (i+1) diff/N, stored as
//   FLOAT points, voxel index by the LazyGrid rule; a sample in the same voxel as the previous one is skipped, out of
//   grid samples are skipped;
//   every cell met:  no Gaussian -> occupancy -= 0.2;  Gaussian -> maximum-likelihood point X of the cell's Gaussian on
//   the line (NDTCell::computeMaximumLikelihoodAlongLine), lik = exp(-d_M(X)^2/2); ignored when X lies beyond the end
//   point; lik *= 1 - exp(-|X-p|^2 / (2 (0.5 dist/30 + sensor_noise)^2)); ignored when < 0.3; occupancy +=
//   log((1-q)/q), q = 0.1 lik + 0.5;  occupancy clamped to +-limit, and a cell whose occupancy drops to <= 0 loses its
//   Gaussian at once (later rays of the same scan see it as an empty cell);
//   finally the end point is binned like loadPointCloud does (LazyGrid::addPoint, update_set).
// `classifierTh` is unused by this branch upstream.  The cells of an initialize()d map all exist ("initializeAll"); for
// a lazily allocated grid upstream would add the sample as a fake point to the new cell — here (and in the engine) a
// cell met by a ray on such a map is simply created empty, documented deviation outside the reference's call sites.
namespace {
inline double cell_line_likelihood(const Cell &c, const double *po /*float-rounded origin*/, const double *pe, double *X) {
  double L[3] = {pe[0] - po[0], pe[1] - po[1], pe[2] - po[2]};
  const double nrm = std::sqrt(L[0] * L[0] + L[1] * L[1] + L[2] * L[2]);
  for (int a = 0; a < 3; a++) L[a] /= nrm;
  double icov[9], det;
  if (!inv3_check(c.cov, icov, det)) {  // icov_ undefined upstream in that case; treat like sigma == 0
    for (int a = 0; a < 3; a++) X[a] = pe[a];
    return 1.0;
  }
  double A[3];
  mat3_vec(icov, L, A);
  const double B[3] = {pe[0] - c.mean[0], pe[1] - c.mean[1], pe[2] - c.mean[2]};
  const double sigma = A[0] * L[0] + A[1] * L[1] + A[2] * L[2];
  if (sigma == 0) {
    for (int a = 0; a < 3; a++) X[a] = pe[a];  // `out` is left untouched upstream; the caller's value is unspecified
    return 1.0;
  }
  const double t = -(A[0] * B[0] + A[1] * B[1] + A[2] * B[2]) / sigma;
  for (int a = 0; a < 3; a++) X[a] = L[a] * t + pe[a];
  // getLikelihood(pcl::PointXYZ): the point is rounded to float
  const double v[3] = {(double)(float)X[0] - c.mean[0], (double)(float)X[1] - c.mean[1], (double)(float)X[2] - c.mean[2]};
  double iv[3];
  mat3_vec(icov, v, iv);
  const double lik = v[0] * iv[0] + v[1] * iv[1] + v[2] * iv[2];
  if (std::isnan(lik)) return -1;
  return std::exp(-lik / 2);
}
inline void update_occupancy(Cell &c, float v, float limit) {  // NDTCell::updateOccupancy
  c.occ += v;
  if (c.occ > limit) c.occ = limit;
  if (c.occ < -limit) c.occ = -limit;
}
}  // namespace

int64_t orc_map_add_point_cloud(orc_map *m, const double *origin, const float *pts, int64_t n, double classifier_th,
                                double maxz, double sensor_noise, double occupancy_limit) {
  (void)classifier_th;
  if (m->is_first_load) return orc_map_load_point_cloud(m, pts, n, -1.0);
  if (!m->grid_ready) return -1;
  const double po[3] = {(double)(float)origin[0], (double)(float)origin[1], (double)(float)origin[2]};
  const double max_range = 200.;
  const double resolution = std::min(std::min(m->cell[0], m->cell[1]), std::min(m->cell[2], m->cell[1]));
  int64_t added = 0;
  for (int64_t i = 0; i < n; i++) {
    const float *p = pts + 4 * i;
    if (std::isnan(p[0]) || std::isnan(p[1]) || std::isnan(p[2])) continue;
    const double pe[3] = {p[0], p[1], p[2]};
    const double diff[3] = {pe[0] - origin[0], pe[1] - origin[1], pe[2] - origin[2]};
    const double l = std::sqrt(diff[0] * diff[0] + diff[1] * diff[1] + diff[2] * diff[2]);
    if (l > max_range) continue;
    // ---- LazyGrid::traceLine
    if (pe[2] > maxz) continue;
    if (resolution < 0.01) continue;
    const int N = (int)(l / resolution);
    const double fN = (double)(float)N;
    const double d[3] = {diff[0] / fN, diff[1] / fN, diff[2] / fN};
    int xo = 0, yo = 0, zo = 0;
    for (int s = 0; s < N - 2; s++) {
      const double f = (double)(float)(s + 1);
      const float q[3] = {(float)(origin[0] + f * d[0]), (float)(origin[1] + f * d[1]), (float)(origin[2] + f * d[2])};
      int x, y, z;
      if (!m->index_of(q[0], q[1], q[2], x, y, z)) continue;
      if (x == xo && y == yo && z == zo) continue;
      xo = x, yo = y, zo = z;
      if (!m->inb(x, y, z)) continue;
      Cell &c = m->cells[(size_t)m->find_or_create(x, y, z)];
      if (c.has_gaussian) {
        double X[3];
        double lik = cell_line_likelihood(c, po, pe, X);
        const double dist = std::sqrt((origin[0] - X[0]) * (origin[0] - X[0]) + (origin[1] - X[1]) * (origin[1] - X[1]) +
                                      (origin[2] - X[2]) * (origin[2] - X[2]));
        if (dist > l) continue;  // the maximum-likelihood point lies beyond the measurement
        const double l2 = std::sqrt((X[0] - pe[0]) * (X[0] - pe[0]) + (X[1] - pe[1]) * (X[1] - pe[1]) +
                                    (X[2] - pe[2]) * (X[2] - pe[2]));
        const double snoise = 0.5 * (dist / 30.0) + sensor_noise;
        const double thr = std::exp(-0.5 * (l2 * l2) / (snoise * snoise));
        lik *= (1.0 - thr);
        if (lik < 0.3) continue;
        lik = 0.1 * lik + 0.5;
        const double logodd = std::log((1.0 - lik) / lik);
        update_occupancy(c, (float)logodd, (float)occupancy_limit);
        if (c.occ <= 0) c.has_gaussian = false;
      } else {
        update_occupancy(c, (float)-0.2, (float)occupancy_limit);
        if (c.occ <= 0) c.has_gaussian = false;
      }
    }
    if (m->add_point(p)) added++;
  }
  return added;
}

void orc_map_compute_cells(orc_map *m, uint32_t maxnumpoints, float occupancy_limit) {
  for (int32_t id : m->update_set) {
    compute_gaussian(m->cells[(size_t)id], maxnumpoints, occupancy_limit);
    m->in_update[(size_t)id] = 0;
  }
  m->update_set.clear();
}

int orc_map_from_cells(orc_map *m, const orc_grid *g, const orc_cell *cells, int64_t n, int use_idx) {
  for (int i = 0; i < 3; i++) {
    m->cell[i] = g->cell[i];
    m->center[i] = g->center[i];
    m->size[i] = g->size[i];
    m->size_m[i] = g->size[i] * g->cell[i];
  }
  m->guess_size = false;
  m->is_first_load = false;
  m->reset_storage();
  for (int64_t i = 0; i < n; i++) {
    int x, y, z;
    if (use_idx) {
      x = cells[i].idx[0], y = cells[i].idx[1], z = cells[i].idx[2];
    } else {
      if (!m->index_of((float)cells[i].mean[0], (float)cells[i].mean[1], (float)cells[i].mean[2], x, y, z)) return -1;
    }
    if (!m->inb(x, y, z)) return -1;
    int32_t id = m->find_or_create(x, y, z);
    Cell &c = m->cells[(size_t)id];
    for (int k = 0; k < 3; k++) c.mean[k] = cells[i].mean[k];
    const double *t = cells[i].cov;
    double full[9] = {t[0], t[1], t[2], t[1], t[3], t[4], t[2], t[4], t[5]};
    std::memcpy(c.cov, full, sizeof full);
    c.N = cells[i].n;
    c.has_gaussian = cells[i].has_gaussian != 0;
    c.occ = cells[i].occ;
  }
  return 0;
}

void orc_map_grid(const orc_map *m, orc_grid *g) {
  for (int i = 0; i < 3; i++) g->center[i] = m->center[i], g->cell[i] = m->cell[i], g->size[i] = m->size[i];
}

int64_t orc_map_num_cells(const orc_map *m, int gaussian_only) {
  int64_t c = 0;
  for (auto &cl : m->cells)
    if (!gaussian_only || cl.has_gaussian) c++;
  return c;
}

int64_t orc_map_export_cells(const orc_map *m, orc_cell *out, int64_t cap, int gaussian_only) {
  std::vector<std::pair<int64_t, int32_t>> ord;
  for (size_t i = 0; i < m->cells.size(); i++) {
    const Cell &c = m->cells[i];
    if (gaussian_only && !c.has_gaussian) continue;
    ord.push_back({m->lin(c.idx[0], c.idx[1], c.idx[2]), (int32_t)i});
  }
  std::sort(ord.begin(), ord.end());
  int64_t k = 0;
  for (auto &o : ord) {
    if (k >= cap) break;
    const Cell &c = m->cells[(size_t)o.second];
    orc_cell &e = out[k++];
    for (int a = 0; a < 3; a++) e.mean[a] = c.mean[a], e.idx[a] = c.idx[a];
    e.cov[0] = c.cov[0], e.cov[1] = c.cov[1], e.cov[2] = c.cov[2];
    e.cov[3] = c.cov[4], e.cov[4] = c.cov[5], e.cov[5] = c.cov[8];
    e.n = c.N;
    e.has_gaussian = c.has_gaussian;
    e.occ = c.occ;
  }
  return (int64_t)ord.size();
}

int64_t orc_map_point_indices(const orc_map *m, const float *pts, int64_t n, int32_t *out) {
  int64_t inb = 0;
  for (int64_t i = 0; i < n; i++) {
    int x = -1, y = -1, z = -1;
    const float *p = pts + 4 * i;
    bool ok = !(std::isnan(p[0]) || std::isnan(p[1]) || std::isnan(p[2])) && m->index_of(p[0], p[1], p[2], x, y, z);
    if (!ok) x = y = z = INT32_MIN;
    out[3 * i] = x, out[3 * i + 1] = y, out[3 * i + 2] = z;
    if (ok && m->inb(x, y, z)) inb++;
  }
  return inb;
}

int orc_d2d_derivatives(const orc_map *tgt, const orc_map *src, const double *T, const orc_params *p, int want_hessian,
                        double *out43, int64_t *n_pairs) {
  std::vector<Gauss> s;
  src->gaussians(s);
  Deriv D;
  derivatives_cells(s, pose_from_cm(T), *tgt, *p, want_hessian != 0, D);
  out43[0] = D.score;
  for (int i = 0; i < 6; i++) out43[1 + i] = D.g[i];
  for (int i = 0; i < 36; i++) out43[7 + i] = want_hessian ? D.H[i] : 0.0;
  if (n_pairs) *n_pairs = D.pairs;
  return 0;
}

// NDTMatcherD2D::lineSearchMT(increment, sourceNDT, targetNDT) [upstream] == ndt_matcher_d2d_fusion.h:390-793 without the
// feature terms: the source cells moved by T are the `sourceNDT` vector; incr6 may be negated in place (:462)
double orc_d2d_line_search(const orc_map *tgt, const orc_map *src, const double *T, double *incr6, const orc_params *p) {
  std::vector<Gauss> s;
  src->gaussians(s);
  Counters cnt;
  return line_search_mt(incr6, s, pose_from_cm(T), *tgt, *p, cnt);
}

// NDTMap::loadPointCloudCentroid(pc, origin, old_centroid, map_size, range_limit) [upstream], the loadCentroid branch of
// the fuser's local-map build (ndt_feature_fuser_hmt.cpp:199-217): centre = old_centroid + floor((origin - old_centroid) /
// cell) * cell, size = map_size, range filter around `origin`
int64_t orc_map_load_point_cloud_centroid(orc_map *m, const float *pts, int64_t n, const double *origin, const double *old_centroid,
                                          const double *map_size, double range_limit) {
  double c[3];
  for (int a = 0; a < 3; a++) c[a] = old_centroid[a] + std::floor((origin[a] - old_centroid[a]) / m->cell[a]) * m->cell[a];
  m->guess_size = false;
  m->is_first_load = false;
  m->centerx = c[0], m->centery = c[1], m->centerz = c[2];
  m->map_sizex = map_size[0], m->map_sizey = map_size[1], m->map_sizez = map_size[2];
  m->set_grid(c[0], c[1], c[2], map_size[0], map_size[1], map_size[2]);
  int64_t added = 0;
  for (int64_t i = 0; i < n; i++) {
    const float *p = pts + 4 * i;
    if (std::isnan(p[0]) || std::isnan(p[1]) || std::isnan(p[2])) continue;
    if (range_limit > 0) {
      const double d0 = (double)p[0] - origin[0], d1 = (double)p[1] - origin[1], d2 = (double)p[2] - origin[2];
      if (std::sqrt(d0 * d0 + d1 * d1 + d2 * d2) > range_limit) continue;
    }
    if (m->add_point(p)) added++;
  }
  return added;
}

int orc_d2d_match(const orc_map *tgt, const orc_map *src, const double *T0, const orc_params *p, orc_result *res) {
  return match_impl(*tgt, *src, T0, *p, nullptr, *res);
}

int orc_fusion_match(const orc_map *tgt, const orc_map *src, const double *T0, const double *Tcov36,
                     const orc_params *p, orc_result *res) {
  Fusion f{Tcov36};
  return match_impl(*tgt, *src, T0, *p, &f, *res);
}

int orc_p2d_derivatives(const orc_map *tgt, const float *pts, int64_t n, const double *T, const orc_params *p,
                        int want_hessian, double *out43, int64_t *n_pairs) {
  std::vector<Gauss> s;
  points_as_cells(pts, n, s);
  Deriv D;
  derivatives_cells(s, pose_from_cm(T), *tgt, *p, want_hessian != 0, D);
  out43[0] = D.score;
  for (int i = 0; i < 6; i++) out43[1 + i] = D.g[i];
  for (int i = 0; i < 36; i++) out43[7 + i] = want_hessian ? D.H[i] : 0.0;
  if (n_pairs) *n_pairs = D.pairs;
  return 0;
}

int orc_p2d_match(const orc_map *tgt, const float *pts, int64_t n, const double *T0, const orc_params *p, orc_result *res) {
  std::vector<Gauss> s;
  points_as_cells(pts, n, s);
  return match_cells(*tgt, s, T0, *p, nullptr, *res);
}

int orc_d2d_covariance(const orc_map *tgt, const orc_map *src, const double *T, const orc_params *p, double *cov36) {
  return covariance_impl(*tgt, *src, T, *p, cov36);
}

int orc_d2d_match_batch(int64_t n_edges, const orc_map *const *tgt, const orc_map *const *src, const double *T0s,
                        const orc_params *p, int with_covariance, int n_threads_edges, orc_result *res,
                        double *cov36s) {
  int rc = 0;
  int nt = n_threads_edges > 1 ? n_threads_edges : 1;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(nt) if (nt > 1)
#endif
  for (int64_t e = 0; e < n_edges; e++) {
    orc_params pe = *p;
    if (nt > 1) pe.n_threads = 1;
    int r = match_impl(*tgt[e], *src[e], T0s + 16 * e, pe, nullptr, res[e]);
    if (r == 0 && with_covariance && cov36s) {
      // ndt_feature_graph.cpp:286-310: covariance only if the pose changed, else 0.02*I
      if (res[e].pose_changed)
        r = covariance_impl(*tgt[e], *src[e], res[e].T, pe, cov36s + 36 * e);
      else
        for (int i = 0; i < 36; i++) cov36s[36 * e + i] = (i % 7 == 0) ? 0.02 : 0.0;
    }
    if (r != 0) rc = r;
  }
  return rc;
}

// NDTCell::getOccupancyRescaled [upstream]: 1 - 1/(1+exp(occ))
static inline double occ_rescaled(float occ) { return 1.0 - 1.0 / (1.0 + std::exp((double)occ)); }

double orc_overlap_occupancy_score(const orc_map *ref, const orc_map *mov, const double *T16) {
  Pose T = pose_from_cm(T16);
  double diff_sum = 0;
  size_t nb = 0;
  for (const Cell &c : mov->cells) {
    double mov_occ = occ_rescaled(c.occ);
    if (mov_occ == 0.5) continue;
    // NDTCell::getCenter: pcl::PointXYZ (float) centre of the voxel
    float ctr[3];
    for (int a = 0; a < 3; a++) {
      int idc = (int)(mov->size[a] / 2.0);
      ctr[a] = (float)(mov->center[a] + (c.idx[a] - idc) * mov->cell[a]);
    }
    double e[3] = {ctr[0], ctr[1], ctr[2]}, t[3];
    mat3_vec(T.R, e, t);
    float pt[3] = {(float)(t[0] + T.t[0]), (float)(t[1] + T.t[1]), (float)(t[2] + T.t[2])};
    int x, y, z;
    if (!ref->index_of(pt[0], pt[1], pt[2], x, y, z)) continue;
    int32_t id = ref->find(x, y, z);
    if (id < 0) continue;
    double ref_occ = occ_rescaled(ref->cells[(size_t)id].occ);
    if (ref_occ != 0.5) {
      nb++;
      double d = mov_occ - ref_occ;
      diff_sum += d * d;
    }
  }
  if (nb == 0) return 1.0;
  return diff_sum / (1.0 * nb);
}

int orc_cstep(double *stx, double *fx, double *dx, double *sty, double *fy, double *dy, double *stp, double fp,
              double dp, int *brackt, double stmin, double stmax) {
  bool b = *brackt != 0;
  int info = cstep(*stx, *fx, *dx, *sty, *fy, *dy, *stp, fp, dp, b, stmin, stmax);
  *brackt = b;
  return info;
}

void orc_eig_sym(int n, const double *A, double *evals, double *evecs) { jacobi_eig(n, A, evals, evecs); }

void orc_pose_from_vec(const double *p6, double *T16) { pose_to_cm(pose_from_vec(p6), T16); }

double orc_robust_yaw(const double *T) {
  // v2 = R * (1,0,0) = first column; angle in the xy-plane (utils.h:30-40)
  double v2x = T[0], v2y = T[1];
  double d = v2x;
  if (d > 1) d = 1;  // acos domain guard (reference would return NaN)
  if (d < -1) d = -1;
  double angle = std::acos(d);
  return (v2y > 0) ? angle : -angle;
}

int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"
