// fuser.cu — host side of the front end behind the C ABI: NDTFeatureFuserHMT (per-scan step on one node map) and the
// NDTFeatureGraph node chain.  Pure orchestration over the map / matcher entry points of api.cu: every cloud transform,
// map build, ray trace, registration and covariance runs on the device; nothing here touches point or cell data.
//
// Reference path replaced:
//   NDTFeatureFuserHMT::initialize   ndt_feature/src/ndt_feature_src/ndt_feature_fuser_hmt.cpp:65-102
//   NDTFeatureFuserHMT::update       ndt_feature_fuser_hmt.cpp:108-512 (useFeat = useOdom = false, loadCentroid = false)
//   NDTFeatureGraph::initialize      ndt_feature/src/ndt_feature_src/ndt_feature_graph.cpp:24-55
//   NDTFeatureGraph::update          ndt_feature_graph.cpp:60-144
//   MotionModel2d::getCovMatrix6     ndt_feature/src/ndt_feature_src/motion_model.cpp:175-207
#include <cmath>
#include <cstring>
#include <memory>
#include <vector>

#include "../../include/ndtb.h"
#include "optimizer.h"

extern "C" {
int ndtb_internal_alloc(ndtb_ctx *ctx, size_t bytes, void **out);
void ndtb_internal_free(ndtb_ctx *ctx, void *p);
int ndtb_internal_upload(ndtb_ctx *ctx, void *dst, const void *src, size_t bytes, int src_mem);
}

using ndtb::Pose;

namespace {

// Eigen::Matrix3d::eulerAngles(0, 1, 2) (Eigen 3.2: first angle in [0, pi]); R row-major
void euler_xyz(const double *R, double *res) {
  auto at = [&](int r, int c) { return R[r * 3 + c]; };
  const int i = 0, j = 1, k = 2;
  res[0] = std::atan2(at(j, k), at(k, k));
  const double c2 = std::sqrt(at(i, i) * at(i, i) + at(i, j) * at(i, j));
  if (res[0] > 0.0) {
    res[0] = res[0] - M_PI;
    res[1] = std::atan2(-at(i, k), -c2);
  } else {
    res[1] = std::atan2(-at(i, k), c2);
  }
  const double s1 = std::sin(res[0]), c1 = std::cos(res[0]);
  res[2] = std::atan2(s1 * at(k, i) - c1 * at(j, i), c1 * at(j, j) - s1 * at(k, j));
  for (int q = 0; q < 3; q++) res[q] = -res[q];
}

Pose identity_pose() {
  Pose P;
  for (int i = 0; i < 9; i++) P.R[i] = (i % 4 == 0) ? 1.0 : 0.0;
  P.t[0] = P.t[1] = P.t[2] = 0.0;
  return P;
}

struct DevBuf {  // grow-only device scratch
  ndtb_ctx *ctx = nullptr;
  void *p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return NDTB_OK;
    if (p) ndtb_internal_free(ctx, p);
    p = nullptr, cap = 0;
    const size_t want = bytes + bytes / 2;
    if (int rc = ndtb_internal_alloc(ctx, want, &p)) return rc;
    cap = want;
    return NDTB_OK;
  }
  ~DevBuf() {
    if (p) ndtb_internal_free(ctx, p);
  }
};

}  // namespace

struct ndtb_fuser {
  ndtb_ctx *ctx;
  ndtb_fuser_params p;
  ndtb_map *map = nullptr;
  Pose Tnow, sensor;
  bool is_init = false;
  DevBuf raw, tmp;
  ~ndtb_fuser() {
    if (map) ndtb_map_destroy(map);
  }
};

namespace {

// device copy of the caller's cloud (or the caller's device pointer itself)
int stage_cloud(ndtb_fuser *f, const float *cloud, int64_t n, int mem, const float **dev) {
  if (mem == NDTB_MEM_DEVICE) {
    *dev = cloud;
    return NDTB_OK;
  }
  if (int rc = f->raw.reserve(16 * (size_t)(n > 0 ? n : 1))) return rc;
  if (n > 0)
    if (int rc = ndtb_internal_upload(f->ctx, f->raw.p, cloud, 16 * (size_t)n, mem)) return rc;
  *dev = (const float *)f->raw.p;
  return NDTB_OK;
}

void match_params(const ndtb_fuser_params &fp, ndtb_params *mp) {
  ndtb_default_params(mp);
  mp->n_neighbours = fp.neighbours, mp->itr_max = fp.itr_max, mp->step_control = fp.step_control;
  mp->delta_score = fp.delta_score;
  mp->use_soft_constraints = fp.use_soft_constraints, mp->use_tikhonov = fp.use_tikhonov, mp->planar = fp.fusion2d;
}

// TmotionCov of update(): MotionModel2d::getCovMatrix6(relpose) with z / roll / pitch variances set to 1 (:137-143)
void motion_cov6(const ndtb_fuser_params &fp, const Pose &Tm, double *cov36) {
  double e[3];
  euler_xyz(Tm.R, e);
  const double d2 = Tm.t[0] * Tm.t[0] + Tm.t[1] * Tm.t[1], rot = e[2];
  const double Cd = fp.motion[0], Ct = fp.motion[1], Dd = fp.motion[2], Dt = fp.motion[3], Td = fp.motion[4], Tt = fp.motion[5];
  for (int i = 0; i < 36; i++) cov36[i] = (i % 7 == 0) ? 1.0 : 0.0;
  cov36[0] = Dd * d2 + Dt * rot * rot;
  cov36[7] = Cd * d2 + Ct * rot * rot;
  cov36[35] = Td * d2 + Tt * rot * rot;
}

}  // namespace

extern "C" {

void ndtb_fuser_default_params(ndtb_fuser_params *p) {  // NDTFeatureFuserHMT::Params() (ndt_feature_fuser_hmt.h:60-95)
  std::memset(p, 0, sizeof *p);
  p->resolution = 1.0, p->map_size_x = 40.0, p->map_size_y = 40.0, p->map_size_z = 10.0, p->sensor_range = 3.0;
  p->max_translation_norm = 1.0, p->max_rotation_norm = M_PI / 4.0;
  p->delta_score = 10e-4, p->neighbours = 0, p->itr_max = 30, p->step_control = 1;
  p->global_transf = 1, p->use_soft_constraints = 1, p->use_tikhonov = 1, p->compute_cov = 1, p->fusion2d = 0;
  p->all_matches_valid = 0, p->fuse_incomplete = 0, p->check_consistency = 0, p->force_odom_as_est = 0;
  for (int i = 0; i < 16; i++) p->sensor_pose[i] = (i % 5 == 0) ? 1.0 : 0.0;
  const double m[6] = {0.001, 0.001, 0.005, 0.005, 0.001, 0.001};  // MotionModel2d::Params()
  for (int i = 0; i < 6; i++) p->motion[i] = m[i];
}

int ndtb_fuser_create(ndtb_ctx *ctx, const ndtb_fuser_params *p, ndtb_fuser **out) {
  if (!ctx || !p || !out || !(p->resolution > 0)) return NDTB_ERR_ARG;
  ndtb_fuser *f = new ndtb_fuser();
  f->ctx = ctx, f->p = *p;
  f->Tnow = identity_pose();
  f->sensor = ndtb::pose_from_cm(p->sensor_pose);
  f->raw.ctx = f->tmp.ctx = ctx;
  *out = f;
  return NDTB_OK;
}

void ndtb_fuser_destroy(ndtb_fuser *f) { delete f; }

ndtb_map *ndtb_fuser_map(ndtb_fuser *f) { return f ? f->map : nullptr; }

int ndtb_fuser_pose(const ndtb_fuser *f, double *Tnow16) {
  if (!f || !Tnow16) return NDTB_ERR_ARG;
  ndtb::pose_to_cm(f->Tnow, Tnow16);
  return NDTB_OK;
}

int ndtb_fuser_set_pose(ndtb_fuser *f, const double *Tnow16) {
  if (!f || !Tnow16) return NDTB_ERR_ARG;
  f->Tnow = ndtb::pose_from_cm(Tnow16);
  return NDTB_OK;
}

int ndtb_fuser_initialize(ndtb_fuser *f, const double *init_pose16, const float *cloud, int64_t n, int mem) {
  if (!f || !init_pose16 || n < 0 || (n > 0 && !cloud)) return NDTB_ERR_ARG;
  const float *d_raw;
  if (int rc = stage_cloud(f, cloud, n, mem, &d_raw)) return rc;
  if (int rc = f->tmp.reserve(16 * (size_t)(n > 0 ? n : 1))) return rc;
  float *d_tmp = (float *)f->tmp.p;
  // :75-76 the cloud goes to the vehicle frame, then to the initial pose (two float transforms)
  if (int rc = ndtb_transform_point_cloud(f->ctx, f->p.sensor_pose, d_raw, n, NDTB_MEM_DEVICE, d_tmp, NDTB_MEM_DEVICE)) return rc;
  if (int rc = ndtb_transform_point_cloud(f->ctx, init_pose16, d_tmp, n, NDTB_MEM_DEVICE, d_tmp, NDTB_MEM_DEVICE)) return rc;
  f->Tnow = ndtb::pose_from_cm(init_pose16);
  if (f->map) ndtb_map_destroy(f->map), f->map = nullptr;
  if (int rc = ndtb_map_create(f->ctx, f->p.resolution, f->p.resolution, f->p.resolution, &f->map)) return rc;
  if (int rc = ndtb_map_initialize(f->map, f->Tnow.t[0], f->Tnow.t[1], 0.0, f->p.map_size_x, f->p.map_size_y, f->p.map_size_z)) return rc;
  const Pose sp = ndtb::pose_mul(f->Tnow, f->sensor);
  if (int rc = ndtb_map_add_point_cloud(f->map, sp.t, d_tmp, n, NDTB_MEM_DEVICE, 0.1, 100.0, 0.1, 255.0)) return rc;
  if (int rc = ndtb_map_compute_cells(f->map, 100000u, 255.f)) return rc;
  f->is_init = true;
  return NDTB_OK;
}

int ndtb_fuser_update(ndtb_fuser *f, const double *Tmotion16, const float *cloud, int64_t n, int mem, int update_ndt_map,
                      double *Tnow16, ndtb_result *res, double *cov36) {
  if (!f || !Tmotion16 || n < 0 || (n > 0 && !cloud)) return NDTB_ERR_ARG;
  if (!f->is_init) return NDTB_ERR_GRID;  // "Call Initialize first!!" (:110-113)
  const ndtb_fuser_params &fp = f->p;
  const Pose Tmotion = ndtb::pose_from_cm(Tmotion16);
  double Tcov[36];
  motion_cov6(fp, Tmotion, Tcov);
  Pose Tinit, Test;
  if (fp.global_transf)
    Tinit = f->Tnow, Test = Tmotion;
  else
    Tinit = identity_pose(), Test = ndtb::pose_mul(f->Tnow, Tmotion);
  const float *d_raw;
  if (int rc = stage_cloud(f, cloud, n, mem, &d_raw)) return rc;
  if (int rc = f->tmp.reserve(16 * (size_t)(n > 0 ? n : 1))) return rc;
  float *d_tmp = (float *)f->tmp.p;
  double T16[16];
  ndtb::pose_to_cm(ndtb::pose_mul(Tinit, f->sensor), T16);
  if (int rc = ndtb_transform_point_cloud(f->ctx, T16, d_raw, n, NDTB_MEM_DEVICE, d_tmp, NDTB_MEM_DEVICE)) return rc;
  // :195-227 local map of the scan
  ndtb_map *local = nullptr;
  if (int rc = ndtb_map_create(f->ctx, fp.resolution, fp.resolution, fp.resolution, &local)) return rc;
  std::unique_ptr<ndtb_map, void (*)(ndtb_map *)> local_guard(local, ndtb_map_destroy);
  if (!fp.global_transf) ndtb_map_guess_size(local, 0, 0, 0, fp.sensor_range, fp.sensor_range, fp.map_size_z);
  if (int rc = ndtb_map_load_point_cloud(local, d_tmp, n, fp.sensor_range, NDTB_MEM_DEVICE, nullptr)) return rc;
  if (int rc = ndtb_map_compute_cells(local, 0xffffffffu, 255.f)) return rc;
  // :352-358 matchFusion (NDT term; the soft / Tikhonov prior needs Tcov^-1, otherwise Tcov is not read)
  ndtb_params mp;
  match_params(fp, &mp);
  double Tin[16];
  ndtb::pose_to_cm(Test, Tin);
  double ident[36];
  for (int i = 0; i < 36; i++) ident[i] = (i % 7 == 0) ? 1.0 : 0.0;
  ndtb_result r;
  if (int rc = ndtb_fusion_match(f->ctx, f->map, local, Tin, (fp.use_soft_constraints || fp.use_tikhonov) ? Tcov : ident, &mp, &r))
    return rc;
  bool match_ok = r.converged != 0 || fp.fuse_incomplete;
  if (fp.all_matches_valid) match_ok = true;
  Test = ndtb::pose_from_cm(r.T);
  if (cov36)
    for (int i = 0; i < 36; i++) cov36[i] = 0.0;
  if (match_ok) {
    if (fp.compute_cov && cov36) {  // :399-420
      const int rc = ndtb_d2d_covariance(f->ctx, f->map, local, r.T, &mp, cov36);
      if (rc != NDTB_OK && rc != NDTB_ERR_SINGULAR) return rc;
    }
    const Pose diff = ndtb::pose_mul(ndtb::pose_inverse(Test), Tmotion);  // :436-441
    double e[3];
    euler_xyz(diff.R, e);
    const double tn = std::sqrt(diff.t[0] * diff.t[0] + diff.t[1] * diff.t[1] + diff.t[2] * diff.t[2]);
    const double rn = std::sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
    if ((tn > fp.max_translation_norm || rn > fp.max_rotation_norm) && fp.check_consistency)
      f->Tnow = ndtb::pose_mul(f->Tnow, Tmotion);
    else if (fp.force_odom_as_est)
      f->Tnow = ndtb::pose_mul(f->Tnow, Tmotion);
    else
      f->Tnow = fp.global_transf ? ndtb::pose_mul(f->Tnow, Test) : Test;
  } else {
    f->Tnow = ndtb::pose_mul(f->Tnow, Tmotion);  // :471-474
  }
  if (update_ndt_map) {  // :476-487
    const Pose sp = ndtb::pose_mul(f->Tnow, f->sensor);
    ndtb::pose_to_cm(sp, T16);
    if (int rc = ndtb_transform_point_cloud(f->ctx, T16, d_raw, n, NDTB_MEM_DEVICE, d_tmp, NDTB_MEM_DEVICE)) return rc;
    if (int rc = ndtb_map_add_point_cloud(f->map, sp.t, d_tmp, n, NDTB_MEM_DEVICE, 0.06, 25.0, 0.25, 255.0)) return rc;
    if (int rc = ndtb_map_compute_cells(f->map, 100000u, 255.f)) return rc;
  }
  if (Tnow16) ndtb::pose_to_cm(f->Tnow, Tnow16);
  if (res) *res = r;
  return NDTB_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------ graph
struct ndtb_graph {
  struct Node {
    ndtb_fuser *fuser = nullptr;
    Pose T, Tlocal_odom, Tlocal_fuse;
    int nb_updates = 0;
  };
  ndtb_ctx *ctx;
  ndtb_fuser_params p;
  double new_node_transl_dist = 1.0;
  double distance_moved = 0.0;
  Pose Tnow;
  std::vector<Node> nodes;
  ~ndtb_graph() {
    for (auto &n : nodes) ndtb_fuser_destroy(n.fuser);
  }
};

namespace {
int graph_new_node(ndtb_graph *g, const Pose &T, const float *cloud, int64_t n, int mem) {
  ndtb_graph::Node nd;
  if (int rc = ndtb_fuser_create(g->ctx, &g->p, &nd.fuser)) return rc;
  double I16[16];
  ndtb::pose_to_cm(identity_pose(), I16);  // every node map lives in its own frame (ndt_feature_graph.cpp:34,100-103)
  if (int rc = ndtb_fuser_initialize(nd.fuser, I16, cloud, n, mem)) {
    ndtb_fuser_destroy(nd.fuser);
    return rc;
  }
  nd.T = T, nd.Tlocal_odom = identity_pose(), nd.Tlocal_fuse = identity_pose();
  g->nodes.push_back(nd);
  return NDTB_OK;
}
}  // namespace

extern "C" {

int ndtb_graph_create(ndtb_ctx *ctx, const ndtb_fuser_params *p, double new_node_transl_dist, ndtb_graph **out) {
  if (!ctx || !p || !out) return NDTB_ERR_ARG;
  ndtb_graph *g = new ndtb_graph();
  g->ctx = ctx, g->p = *p, g->new_node_transl_dist = new_node_transl_dist;
  g->Tnow = identity_pose();
  *out = g;
  return NDTB_OK;
}
void ndtb_graph_destroy(ndtb_graph *g) { delete g; }
int ndtb_graph_set_new_node_dist(ndtb_graph *g, double d) {
  if (!g) return NDTB_ERR_ARG;
  g->new_node_transl_dist = d;
  return NDTB_OK;
}
int64_t ndtb_graph_num_nodes(const ndtb_graph *g) { return g ? (int64_t)g->nodes.size() : 0; }

int ndtb_graph_initialize(ndtb_graph *g, const double *init_pose16, const float *cloud, int64_t n, int mem) {
  if (!g || !init_pose16) return NDTB_ERR_ARG;
  g->Tnow = ndtb::pose_from_cm(init_pose16);
  return graph_new_node(g, g->Tnow, cloud, n, mem);
}

int ndtb_graph_update(ndtb_graph *g, const double *Tmotion16, const float *cloud, int64_t n, int mem, double *Tnow16) {
  if (!g || !Tmotion16 || g->nodes.empty()) return NDTB_ERR_ARG;
  const Pose Tm = ndtb::pose_from_cm(Tmotion16);
  g->distance_moved += std::sqrt(Tm.t[0] * Tm.t[0] + Tm.t[1] * Tm.t[1] + Tm.t[2] * Tm.t[2]);
  const bool spawn = g->distance_moved > g->new_node_transl_dist;
  ndtb_graph::Node &node = g->nodes.back();
  double Tl16[16];
  if (int rc = ndtb_fuser_update(node.fuser, Tmotion16, cloud, n, mem, spawn ? 0 : 1, Tl16, nullptr, nullptr)) return rc;
  const Pose Tl = ndtb::pose_from_cm(Tl16);
  g->Tnow = ndtb::pose_mul(node.T, Tl);
  node.Tlocal_odom = ndtb::pose_mul(node.Tlocal_odom, Tm);
  node.Tlocal_fuse = Tl;
  if (spawn) {
    g->distance_moved = 0.0;
    if (int rc = graph_new_node(g, g->Tnow, cloud, n, mem)) return rc;  // invalidates `node`
  } else {
    node.nb_updates++;
  }
  if (Tnow16) ndtb::pose_to_cm(g->Tnow, Tnow16);
  return NDTB_OK;
}

int ndtb_graph_node(ndtb_graph *g, int64_t k, double *T16, double *Tlocal_odom16, double *Tlocal_fuse16, ndtb_map **map,
                    int32_t *nb_updates) {
  if (!g || k < 0 || k >= (int64_t)g->nodes.size()) return NDTB_ERR_ARG;
  ndtb_graph::Node &nd = g->nodes[(size_t)k];
  if (T16) ndtb::pose_to_cm(nd.T, T16);
  if (Tlocal_odom16) ndtb::pose_to_cm(nd.Tlocal_odom, Tlocal_odom16);
  if (Tlocal_fuse16) ndtb::pose_to_cm(nd.Tlocal_fuse, Tlocal_fuse16);
  if (map) *map = ndtb_fuser_map(nd.fuser);
  if (nb_updates) *nb_updates = nd.nb_updates;
  return NDTB_OK;
}

}  // extern "C"
