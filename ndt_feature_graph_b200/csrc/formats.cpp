// formats.cpp — the hand-off formats of the hot path's results (host code behind the C ABI, no GPU):
//   ndt_feature/NDTEdgeMsg wire bytes     ndt_feature/msg/NDTEdgeMsg.msg + edgeToMsg / msgToEdge,
//                                         ndt_feature/include/ndt_feature/ndtgraph_conversion.h:17-34,104-145
//                                         (ROS1 serialisation: little endian, u32 length prefixes for strings and arrays;
//                                          geometry_msgs/Pose via tf::poseEigenToMsg, matrices via tf::matrixEigenToMsg)
//   pose archives (*.T)                   saveAffine3d / loadAffine3d used by NDTFeatureNode::save / load,
//                                         ndt_feature/include/ndt_feature/ndt_feature_node.h:100-152 (boost text archive,
//                                          byte-compatible with the files in ndt_feature/data/FULL GRAPH/)
//   ndt_feature/NDTGraphMsg / NDTNodeMsg /  nodeToMsg / fuserHMTToMsg / NDTGraphToMsg and their inverses,
//   NDTFeatureFuserHMTMsg wire bytes      ndtgraph_conversion.h:36-83,147-216 + ndt_feature/msg/*.msg.  The node message embeds
//                                         ndt_map/NDTMapMsg of lslgeneric::toMessage [upstream perception_oru, NOT vendored:
//                                          field order restated from the published ndt_map message definitions — unpinned,
//                                          no bag in the reference holds such a message]
//   evaluation trajectory lines           transformToEvalString / transformToEval2dString, ndt_feature/include/ndt_feature/utils.h:243-259
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/ndtb.h"

namespace {

struct Writer {
  std::vector<uint8_t> b;
  void u32(uint32_t v) { put(&v, 4); }
  void f64(double v) { put(&v, 8); }
  void put(const void *p, size_t n) {
    const uint8_t *q = (const uint8_t *)p;
    b.insert(b.end(), q, q + n);
  }
};
struct Reader {
  const uint8_t *p;
  int64_t len, at = 0;
  bool ok = true;
  bool get(void *o, size_t n) {
    if (!ok || at + (int64_t)n > len) return ok = false;
    std::memcpy(o, p + at, n);
    at += (int64_t)n;
    return true;
  }
  uint32_t u32() {
    uint32_t v = 0;
    get(&v, 4);
    return v;
  }
  double f64() {
    double v = 0;
    get(&v, 8);
    return v;
  }
};

// Eigen::Quaterniond(Matrix3d) (RotationBase -> quaternion, Eigen/src/Geometry/Quaternion.h); R row-major
void quat_from_rot(const double *R, double *q /*x y z w*/) {
  auto m = [&](int r, int c) { return R[r * 3 + c]; };
  double t = m(0, 0) + m(1, 1) + m(2, 2);
  if (t > 0.0) {
    t = std::sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (m(2, 1) - m(1, 2)) * t;
    q[1] = (m(0, 2) - m(2, 0)) * t;
    q[2] = (m(1, 0) - m(0, 1)) * t;
  } else {
    int i = 0;
    if (m(1, 1) > m(0, 0)) i = 1;
    if (m(2, 2) > m(i, i)) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (m(k, j) - m(j, k)) * t;
    q[j] = (m(j, i) + m(i, j)) * t;
    q[k] = (m(k, i) + m(i, k)) * t;
  }
}
// Eigen::Quaterniond::toRotationMatrix
void rot_from_quat(const double *q, double *R) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1 - (tyy + tzz), R[1] = txy - twz, R[2] = txz + twy;
  R[3] = txy + twz, R[4] = 1 - (txx + tzz), R[5] = tyz - twx;
  R[6] = txz - twy, R[7] = tyz + twx, R[8] = 1 - (txx + tyy);
}
void rot_of(const double *T16, double *R) {
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) R[r * 3 + c] = T16[c * 4 + r];
}

// geometry_msgs/Pose via tf::poseEigenToMsg: position, orientation x y z w (flipped so that w >= 0)
void put_pose(Writer &w, const double *T16) {
  double R[9], q[4];
  rot_of(T16, R);
  quat_from_rot(R, q);
  if (q[3] < 0)
    for (double &v : q) v = -v;
  w.f64(T16[12]), w.f64(T16[13]), w.f64(T16[14]);
  for (double v : q) w.f64(v);
}
bool get_pose(Reader &r, double *T16) {
  double t[3], q[4], R[9];
  for (double &v : t) v = r.f64();
  for (double &v : q) v = r.f64();
  if (!r.ok) return false;
  rot_from_quat(q, R);  // tf::poseMsgToEigen: Translation * Quaternion
  for (int i = 0; i < 16; i++) T16[i] = 0;
  for (int rr = 0; rr < 3; rr++)
    for (int c = 0; c < 3; c++) T16[c * 4 + rr] = R[rr * 3 + c];
  T16[12] = t[0], T16[13] = t[1], T16[14] = t[2], T16[15] = 1;
  return true;
}
// std_msgs/Float64MultiArray via tf::matrixEigenToMsg: two dimensions (empty labels), row-major data
void put_matrix(Writer &w, const double *m, int rows, int cols) {
  if (!m) rows = cols = 0;  // an unset Eigen::MatrixXd (0 x 0)
  w.u32(2);
  w.u32(0), w.u32((uint32_t)rows), w.u32((uint32_t)(rows * cols));
  w.u32(0), w.u32((uint32_t)cols), w.u32((uint32_t)cols);
  w.u32(0);  // data_offset
  w.u32((uint32_t)(rows * cols));
  for (int i = 0; i < rows * cols; i++) w.f64(m[i]);
}
bool get_matrix(Reader &r, std::vector<double> &data) {
  const uint32_t nd = r.u32();
  for (uint32_t d = 0; d < nd && r.ok; d++) {
    const uint32_t ls = r.u32();
    r.at += ls;  // label
    r.u32(), r.u32();
  }
  r.u32();
  const uint32_t n = r.u32();
  if (!r.ok || (int64_t)n * 8 > r.len - r.at) return r.ok = false;
  data.resize(n);
  for (uint32_t i = 0; i < n; i++) data[i] = r.f64();
  return r.ok;
}

void put_string(Writer &w, const char *s) {
  const size_t n = s ? std::strlen(s) : 0;
  w.u32((uint32_t)n);
  if (n) w.put(s, n);
}
bool get_string(Reader &r, std::string &out) {
  const uint32_t n = r.u32();
  if (!r.ok || (int64_t)n > r.len - r.at) return r.ok = false;
  out.assign((const char *)r.p + r.at, n);
  r.at += n;
  return true;
}
int64_t finish(const Writer &w, uint8_t *out, int64_t cap) {
  if (out && cap >= (int64_t)w.b.size()) std::memcpy(out, w.b.data(), w.b.size());
  return (int64_t)w.b.size();
}
// skips one NDTEdgeMsg / NDTMapMsg / NDTNodeMsg and reports where it ended
bool skip_matrix(Reader &r) {
  std::vector<double> d;
  return get_matrix(r, d);
}
bool skip_map_msg(Reader &r) {
  std::string f;
  r.u32(), r.u32(), r.u32();
  if (!get_string(r, f)) return false;
  r.at += 9 * 8;
  const uint32_t n = r.u32();
  for (uint32_t i = 0; i < n && r.ok; i++) {
    r.at += 4 * 8;
    const uint32_t nc = r.u32();
    r.at += (int64_t)nc * 8 + 8;
    if (r.at > r.len) r.ok = false;
  }
  return r.ok && r.at <= r.len;
}

std::string g15(double v) {
  char buf[64];
  std::snprintf(buf, sizeof buf, "%.15g", v);
  return buf;
}

}  // namespace

extern "C" {

int64_t ndtb_edge_msg_pack(uint32_t ref_idx, uint32_t mov_idx, const double *T16, const double *cov9, const double *cov36, double score,
                           uint8_t *out, int64_t cap) {
  if (!T16 || !cov9) return NDTB_ERR_ARG;
  Writer w;
  w.u32(ref_idx), w.u32(mov_idx);
  put_pose(w, T16);
  put_matrix(w, cov9, 3, 3);
  put_matrix(w, cov36, 6, 6);
  w.f64(score);
  if (out && cap >= (int64_t)w.b.size()) std::memcpy(out, w.b.data(), w.b.size());
  return (int64_t)w.b.size();
}

int ndtb_edge_msg_unpack(const uint8_t *buf, int64_t len, uint32_t *ref_idx, uint32_t *mov_idx, double *T16, double *cov9, double *cov36,
                         int32_t *has_cov36, double *score) {
  if (!buf || len < 0 || !T16) return NDTB_ERR_ARG;
  Reader r{buf, len};
  const uint32_t a = r.u32(), b = r.u32();
  if (!get_pose(r, T16)) return NDTB_ERR_ARG;
  std::vector<double> c3, c6;
  if (!get_matrix(r, c3) || !get_matrix(r, c6)) return NDTB_ERR_ARG;
  const double s = r.f64();
  if (!r.ok || c3.size() != 9 || (c6.size() != 0 && c6.size() != 36)) return NDTB_ERR_ARG;  // msgToEdge asserts the same sizes
  if (ref_idx) *ref_idx = a;
  if (mov_idx) *mov_idx = b;
  if (cov9) std::copy(c3.begin(), c3.end(), cov9);
  if (cov36 && c6.size() == 36) std::copy(c6.begin(), c6.end(), cov36);
  if (has_cov36) *has_cov36 = c6.size() == 36;
  if (score) *score = s;
  return NDTB_OK;
}

// ---- ndt_map/NDTMapMsg [upstream]: Header, x/y/z_size (metres), x/y/z_cen, x/y/z_cell_size, NDTCellMsg[] cells with
// NDTCellMsg = mean_x, mean_y, mean_z, occupancy, float64[] cov_matrix (9, row-major), N.  toMessage writes the cells that
// hold a Gaussian; fromMessage re-inserts every cell at the voxel of its mean (ndtb_map_from_cells with use_idx = 0).
int64_t ndtb_map_msg_pack(uint32_t seq, uint32_t sec, uint32_t nsec, const char *frame_id, const ndtb_grid *g, const ndtb_cell *cells,
                          int64_t n_cells, uint8_t *out, int64_t cap) {
  if (!g || n_cells < 0 || (n_cells > 0 && !cells)) return NDTB_ERR_ARG;
  Writer w;
  w.u32(seq), w.u32(sec), w.u32(nsec);
  put_string(w, frame_id);
  for (int a = 0; a < 3; a++) w.f64((double)g->size[a] * g->cell[a]);
  for (int a = 0; a < 3; a++) w.f64(g->center[a]);
  for (int a = 0; a < 3; a++) w.f64(g->cell[a]);
  uint32_t ng = 0;
  for (int64_t i = 0; i < n_cells; i++) ng += cells[i].has_gaussian != 0;
  w.u32(ng);
  for (int64_t i = 0; i < n_cells; i++) {
    const ndtb_cell &c = cells[i];
    if (!c.has_gaussian) continue;
    w.f64(c.mean[0]), w.f64(c.mean[1]), w.f64(c.mean[2]);
    w.f64((double)c.occ);
    w.u32(9);
    const double m[9] = {c.cov[0], c.cov[1], c.cov[2], c.cov[1], c.cov[3], c.cov[4], c.cov[2], c.cov[4], c.cov[5]};
    for (double v : m) w.f64(v);
    w.f64((double)c.n);
  }
  return finish(w, out, cap);
}

int ndtb_map_msg_unpack(const uint8_t *buf, int64_t len, uint32_t *stamp3, char *frame_id, int32_t frame_cap, ndtb_grid *g,
                        ndtb_cell *cells, int64_t cells_cap, int64_t *n_cells, int64_t *consumed) {
  if (!buf || len < 0 || !g) return NDTB_ERR_ARG;
  Reader r{buf, len};
  const uint32_t a = r.u32(), b = r.u32(), c = r.u32();
  std::string frame;
  if (!get_string(r, frame)) return NDTB_ERR_ARG;
  double sz[3], cen[3], cs[3];
  for (double &v : sz) v = r.f64();
  for (double &v : cen) v = r.f64();
  for (double &v : cs) v = r.f64();
  const uint32_t n = r.u32();
  if (!r.ok) return NDTB_ERR_ARG;
  for (int q = 0; q < 3; q++) {
    if (!(cs[q] > 0)) return NDTB_ERR_ARG;
    g->center[q] = cen[q], g->cell[q] = cs[q];
    g->size[q] = (int32_t)std::fabs(std::ceil(sz[q] / cs[q]));  // LazyGrid::initialize
  }
  for (uint32_t i = 0; i < n; i++) {
    ndtb_cell cell;
    std::memset(&cell, 0, sizeof cell);
    for (double &v : cell.mean) v = r.f64();
    cell.occ = (float)r.f64();
    const uint32_t nc = r.u32();
    if (!r.ok || nc != 9) return NDTB_ERR_ARG;
    double m[9];
    for (double &v : m) v = r.f64();
    cell.cov[0] = m[0], cell.cov[1] = m[1], cell.cov[2] = m[2], cell.cov[3] = m[4], cell.cov[4] = m[5], cell.cov[5] = m[8];
    cell.n = (int32_t)r.f64();
    cell.has_gaussian = 1;
    if (!r.ok) return NDTB_ERR_ARG;
    if (cells && (int64_t)i < cells_cap) cells[i] = cell;
  }
  if (stamp3) stamp3[0] = a, stamp3[1] = b, stamp3[2] = c;
  if (frame_id && frame_cap > 0) std::snprintf(frame_id, (size_t)frame_cap, "%s", frame.c_str());
  if (n_cells) *n_cells = n;
  if (consumed) *consumed = r.at;
  return NDTB_OK;
}

// ---- NDTNodeMsg: NDTFeatureFuserHMTMsg map {Pose Tnow, Tlast_fuse, Todom, NDTMapMsg map, u32 ctr}, Pose T,
// Float64MultiArray cov (3x3), Pose Tlocal_odom, Pose Tlocal_fuse, u32 nbUpdates, f64 time_last_update
int64_t ndtb_node_msg_pack(const ndtb_node_fields *f, const uint8_t *map_msg, int64_t map_len, uint8_t *out, int64_t cap) {
  if (!f || !map_msg || map_len <= 0) return NDTB_ERR_ARG;
  Writer w;
  put_pose(w, f->Tnow), put_pose(w, f->Tlast_fuse), put_pose(w, f->Todom);
  w.put(map_msg, (size_t)map_len);
  w.u32(f->ctr);
  put_pose(w, f->T);
  put_matrix(w, f->cov9, 3, 3);
  put_pose(w, f->Tlocal_odom), put_pose(w, f->Tlocal_fuse);
  w.u32(f->nb_updates);
  w.f64(f->time_last_update);
  return finish(w, out, cap);
}

int ndtb_node_msg_unpack(const uint8_t *buf, int64_t len, ndtb_node_fields *f, int64_t *map_off, int64_t *map_len, int64_t *consumed) {
  if (!buf || len < 0 || !f) return NDTB_ERR_ARG;
  Reader r{buf, len};
  if (!get_pose(r, f->Tnow) || !get_pose(r, f->Tlast_fuse) || !get_pose(r, f->Todom)) return NDTB_ERR_ARG;
  const int64_t m0 = r.at;
  if (!skip_map_msg(r)) return NDTB_ERR_ARG;
  const int64_t m1 = r.at;
  f->ctr = r.u32();
  if (!get_pose(r, f->T)) return NDTB_ERR_ARG;
  std::vector<double> c3;
  if (!get_matrix(r, c3) || c3.size() < 9) return NDTB_ERR_ARG;  // msgToNode reads nine values
  std::copy(c3.begin(), c3.begin() + 9, f->cov9);
  if (!get_pose(r, f->Tlocal_odom) || !get_pose(r, f->Tlocal_fuse)) return NDTB_ERR_ARG;
  f->nb_updates = r.u32();
  f->time_last_update = r.f64();
  if (!r.ok) return NDTB_ERR_ARG;
  if (map_off) *map_off = m0;
  if (map_len) *map_len = m1 - m0;
  if (consumed) *consumed = r.at;
  return NDTB_OK;
}

// ---- NDTGraphMsg: Header, Pose sensor_pose_, Pose Tnow, f64 distance_moved_in_last_node_, NDTNodeMsg[] nodes, NDTEdgeMsg[] edges
int64_t ndtb_graph_msg_pack(uint32_t seq, uint32_t sec, uint32_t nsec, const char *frame_id, const double *sensor_pose16,
                            const double *Tnow16, double distance_moved, int64_t n_nodes, const uint8_t *const *node_msgs,
                            const int64_t *node_lens, int64_t n_edges, const uint8_t *const *edge_msgs, const int64_t *edge_lens,
                            uint8_t *out, int64_t cap) {
  if (!sensor_pose16 || !Tnow16 || n_nodes < 0 || n_edges < 0 || (n_nodes > 0 && (!node_msgs || !node_lens)) ||
      (n_edges > 0 && (!edge_msgs || !edge_lens)))
    return NDTB_ERR_ARG;
  Writer w;
  w.u32(seq), w.u32(sec), w.u32(nsec);
  put_string(w, frame_id);
  put_pose(w, sensor_pose16), put_pose(w, Tnow16);
  w.f64(distance_moved);
  w.u32((uint32_t)n_nodes);
  for (int64_t i = 0; i < n_nodes; i++) w.put(node_msgs[i], (size_t)node_lens[i]);
  w.u32((uint32_t)n_edges);
  for (int64_t i = 0; i < n_edges; i++) w.put(edge_msgs[i], (size_t)edge_lens[i]);
  return finish(w, out, cap);
}

int ndtb_graph_msg_unpack(const uint8_t *buf, int64_t len, uint32_t *stamp3, char *frame_id, int32_t frame_cap, double *sensor_pose16,
                          double *Tnow16, double *distance_moved, int64_t *n_nodes, int64_t *node_off, int64_t *node_len,
                          int64_t nodes_cap, int64_t *n_edges, int64_t *edge_off, int64_t *edge_len, int64_t edges_cap) {
  if (!buf || len < 0) return NDTB_ERR_ARG;
  Reader r{buf, len};
  const uint32_t a = r.u32(), b = r.u32(), c = r.u32();
  std::string frame;
  if (!get_string(r, frame)) return NDTB_ERR_ARG;
  double sp[16], tn[16];
  if (!get_pose(r, sp) || !get_pose(r, tn)) return NDTB_ERR_ARG;
  const double dist = r.f64();
  const uint32_t nn = r.u32();
  if (!r.ok) return NDTB_ERR_ARG;
  for (uint32_t i = 0; i < nn; i++) {
    ndtb_node_fields f;
    int64_t used = 0;
    if (ndtb_node_msg_unpack(buf + r.at, len - r.at, &f, nullptr, nullptr, &used) != NDTB_OK) return NDTB_ERR_ARG;
    if (node_off && (int64_t)i < nodes_cap) node_off[i] = r.at;
    if (node_len && (int64_t)i < nodes_cap) node_len[i] = used;
    r.at += used;
  }
  const uint32_t ne = r.u32();
  if (!r.ok) return NDTB_ERR_ARG;
  for (uint32_t i = 0; i < ne; i++) {
    const int64_t e0 = r.at;
    r.u32(), r.u32();
    double T[16];
    if (!get_pose(r, T) || !skip_matrix(r) || !skip_matrix(r)) return NDTB_ERR_ARG;
    r.f64();
    if (!r.ok) return NDTB_ERR_ARG;
    if (edge_off && (int64_t)i < edges_cap) edge_off[i] = e0;
    if (edge_len && (int64_t)i < edges_cap) edge_len[i] = r.at - e0;
  }
  if (stamp3) stamp3[0] = a, stamp3[1] = b, stamp3[2] = c;
  if (frame_id && frame_cap > 0) std::snprintf(frame_id, (size_t)frame_cap, "%s", frame.c_str());
  if (sensor_pose16) std::copy(sp, sp + 16, sensor_pose16);
  if (Tnow16) std::copy(tn, tn + 16, Tnow16);
  if (distance_moved) *distance_moved = dist;
  if (n_nodes) *n_nodes = nn;
  if (n_edges) *n_edges = ne;
  return NDTB_OK;
}

int ndtb_pose_archive_write(const char *path, const double *T16) {
  if (!path || !T16) return NDTB_ERR_ARG;
  FILE *f = std::fopen(path, "wb");
  if (!f) return NDTB_ERR_ARG;
  std::string s = "22 serialization::archive 12 0 0";
  char buf[64];
  for (int i = 0; i < 16; i++) {
    std::snprintf(buf, sizeof buf, " %.17e", T16[i]);
    s += buf;
  }
  s += "\n";
  const bool ok = std::fwrite(s.data(), 1, s.size(), f) == s.size();
  return (std::fclose(f) == 0 && ok) ? NDTB_OK : NDTB_ERR_ARG;
}

int ndtb_pose_archive_read(const char *path, double *T16) {
  if (!path || !T16) return NDTB_ERR_ARG;
  FILE *f = std::fopen(path, "rb");
  if (!f) return NDTB_ERR_ARG;
  char head[64] = {0};
  int ver = 0, a = 0, b = 0;
  bool ok = std::fscanf(f, "%*d %63s %d %d %d", head, &ver, &a, &b) == 4 && std::strcmp(head, "serialization::archive") == 0;
  for (int i = 0; ok && i < 16; i++) ok = std::fscanf(f, "%lf", &T16[i]) == 1;
  std::fclose(f);
  return ok ? NDTB_OK : NDTB_ERR_ARG;
}

int ndtb_eval_string(const double *T16, int planar, char *out, int32_t cap) {
  if (!T16 || !out || cap <= 0) return NDTB_ERR_ARG;
  double R[9], q[4];
  std::string s;
  if (!planar) {
    rot_of(T16, R);
    quat_from_rot(R, q);
    // Eigen prints the transposed translation with its columns right-aligned to the widest coefficient
    const std::string c[3] = {g15(T16[12]), g15(T16[13]), g15(T16[14])};
    const size_t wd = std::max(c[0].size(), std::max(c[1].size(), c[2].size()));
    for (int i = 0; i < 3; i++) s += std::string(wd - c[i].size(), ' ') + c[i] + (i < 2 ? " " : "");
    s += " " + g15(q[0]) + " " + g15(q[1]) + " " + g15(q[2]) + " " + g15(q[3]) + "\n";
  } else {
    // getRobustYawFromAffine3d (utils.h:30-40), then Translation(x, y, 0) * AngleAxis(yaw, Z)
    double d = T16[0];
    d = d > 1.0 ? 1.0 : (d < -1.0 ? -1.0 : d);
    const double a = std::acos(d), yaw = T16[1] > 0 ? a : -a;
    const double cy = std::cos(yaw), sy = std::sin(yaw);
    const double R2[9] = {cy, -sy, 0, sy, cy, 0, 0, 0, 1};
    quat_from_rot(R2, q);
    s = g15(T16[12]) + " " + g15(T16[13]) + " 0.  " + g15(q[0]) + " " + g15(q[1]) + " " + g15(q[2]) + " " + g15(q[3]) + "\n";
  }
  if ((int)s.size() + 1 > cap) return NDTB_ERR_ARG;
  std::memcpy(out, s.c_str(), s.size() + 1);
  return (int)s.size();
}

}  // extern "C"
