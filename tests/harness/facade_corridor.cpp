// facade_corridor.cpp — host-side test program written against the C++ façade (include/ndtb_lslgeneric.hpp), in the
// shape of the reference's own experiment ndt_feature/src/ndt_odom_debug.cpp:94-256: a synthetic corridor (two
// walls at y = -1 / +1 plus extra clusters, x = tan(pi j / 2n) |c_y|, Gaussian noise), the moving cloud is the
// ground-truth transform of the static one, registration starts from an "odometry" guess.  Then the graph entry
// point (updateLinksUsingNDTRegistration, ndt_feature_graph.cpp:347-353) is exercised on three nodes.
//
//   facade_corridor --abi-check        : no GPU needed; checks the library answers and refuses to compute without a device
//   facade_corridor <outdir>           : runs on cuda:0, writes static.bin / moving.bin / third.bin (float4 clouds) and
//                                        prints one JSON object with every result, which tests/test_facade.py compares
//                                        with the oracle run on the same clouds.
#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>

#include "../../include/ndtb_lslgeneric.hpp"

using lslgeneric::LazyGrid;
using lslgeneric::NDTMap;
using lslgeneric::NDTMatcherD2D;
using lslgeneric::SpatialIndex;

static Eigen::Affine3d pose_from(double x, double y, double z, double rx, double ry, double rz) {
  // Translation * AngleAxis(rx, X) * AngleAxis(ry, Y) * AngleAxis(rz, Z)   (ndt_odom_debug.cpp:124-127)
  const double cx = std::cos(rx), sx = std::sin(rx), cy = std::cos(ry), sy = std::sin(ry), cz = std::cos(rz), sz = std::sin(rz);
  const double Rx[9] = {1, 0, 0, 0, cx, -sx, 0, sx, cx}, Ry[9] = {cy, 0, sy, 0, 1, 0, -sy, 0, cy}, Rz[9] = {cz, -sz, 0, sz, cz, 0, 0, 0, 1};
  double A[9], R[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      A[i * 3 + j] = 0;
      for (int k = 0; k < 3; k++) A[i * 3 + j] += Rx[i * 3 + k] * Ry[k * 3 + j];
    }
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      R[i * 3 + j] = 0;
      for (int k = 0; k < 3; k++) R[i * 3 + j] += A[i * 3 + k] * Rz[k * 3 + j];
    }
  Eigen::Affine3d T;
  T.setIdentity();
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) T(i, j) = R[i * 3 + j];
  T(0, 3) = x, T(1, 3) = y, T(2, 3) = z;
  return T;
}

static pcl::PointCloud<pcl::PointXYZ> transform_cloud(const Eigen::Affine3d &T, const pcl::PointCloud<pcl::PointXYZ> &pc) {
  pcl::PointCloud<pcl::PointXYZ> out;
  for (const auto &p : pc.points) {
    const double x = p.x, y = p.y, z = p.z;
    out.push_back(pcl::PointXYZ((float)(T(0, 0) * x + T(0, 1) * y + T(0, 2) * z + T(0, 3)),
                                (float)(T(1, 0) * x + T(1, 1) * y + T(1, 2) * z + T(1, 3)),
                                (float)(T(2, 0) * x + T(2, 1) * y + T(2, 2) * z + T(2, 3))));
  }
  return out;
}

static void dump(const std::string &path, const pcl::PointCloud<pcl::PointXYZ> &pc) {
  FILE *f = std::fopen(path.c_str(), "wb");
  if (!f) std::exit(3);
  std::fwrite(pc.points.data(), sizeof(pcl::PointXYZ), pc.points.size(), f);
  std::fclose(f);
}

static void print_pose(const char *name, const Eigen::Affine3d &T, bool comma = true) {
  std::printf("\"%s\": [", name);
  for (int i = 0; i < 16; i++) std::printf("%.17g%s", T.matrix().data()[i], i < 15 ? ", " : "");
  std::printf("]%s\n", comma ? "," : "");
}
static void print_mat(const char *name, const Eigen::MatrixXd &M, bool comma = true) {
  std::printf("\"%s\": [", name);
  for (int i = 0; i < M.rows(); i++)
    for (int j = 0; j < M.cols(); j++) std::printf("%.17g%s", M(i, j), (i == M.rows() - 1 && j == M.cols() - 1) ? "" : ", ");
  std::printf("]%s\n", comma ? "," : "");
}

int main(int argc, char **argv) {
  if (argc < 2) {
    std::fprintf(stderr, "usage: facade_corridor --abi-check | <outdir>\n");
    return 2;
  }
  if (std::string(argv[1]) == "--abi-check") {
    if (ndtb_version() != NDTB_VERSION) return 1;
    ndtb_ctx *c = nullptr;
    const int rc = ndtb_ctx_create(0, nullptr, &c);
    if (rc == NDTB_OK) {  // a GPU is present: fine, the façade can compute
      ndtb_ctx_destroy(c);
      std::printf("abi ok (device present)\n");
      return 0;
    }
    // no device: the façade must fail like the reference fails (false / empty), never fall back to a CPU path
    NDTMap a(new LazyGrid(0.5), true), b(new LazyGrid(0.5), true);
    NDTMatcherD2D m;
    Eigen::Affine3d T;
    T.setIdentity();
    if (m.match(a, b, T, true)) return 1;
    if (ndtb::last_status() != NDTB_ERR_CUDA) return 1;
    std::printf("abi ok (no device: %s)\n", ndtb_strerror(ndtb::last_status()));
    return 0;
  }
  const std::string out = argv[1];
  const double resolution = 0.5, std_dev = 0.03, max_range = 30.0;
  const int nb_clusters = 6, nb_points = 400, n_neighbours = 2;
  std::mt19937 rng(12345);
  std::uniform_real_distribution<double> ud(-6.0, 6.0);
  std::normal_distribution<double> nd(0.0, std_dev);

  pcl::PointCloud<pcl::PointXYZ> static_pc;
  for (int i = 0; i < nb_clusters; i++) {
    double c_y = ud(rng);
    if (i == 0) c_y = -1;
    if (i == 1) c_y = 1;
    for (int j = 0; j < nb_points; j++) {
      const double angle = M_PI * j / (2. * nb_points);
      const double x = std::tan(angle) * std::fabs(c_y);
      if (std::fabs(x) > max_range) continue;
      const double px = nd(rng) + x, py = c_y + nd(rng), pz = nd(rng);
      static_pc.push_back(pcl::PointXYZ((float)px, (float)py, (float)pz));
      static_pc.push_back(pcl::PointXYZ((float)(-px * 0.7), (float)(py + 0.02), (float)pz));  // something behind the sensor too
    }
  }
  // two cross walls close the corridor: without them x is unobservable (the aperture problem ndt_odom_debug studies)
  for (int j = 0; j < 600; j++) {
    const double y = -6.0 + 12.0 * j / 600.0;
    static_pc.push_back(pcl::PointXYZ((float)(9.0 + nd(rng)), (float)(y + nd(rng)), (float)nd(rng)));
    static_pc.push_back(pcl::PointXYZ((float)(-6.5 + nd(rng)), (float)(y + nd(rng)), (float)nd(rng)));
  }
  const Eigen::Affine3d gt_transform = pose_from(0.12, -0.05, 0, 0, 0, 0.03);
  // match(target, source, T) finds T with T * source ~ target, i.e. gt^-1; the "odometry" guess is a rough version of it
  const Eigen::Affine3d odom_transform = pose_from(-0.08, 0.02, 0, 0, 0, -0.01);
  const pcl::PointCloud<pcl::PointXYZ> moving_pc = transform_cloud(gt_transform, static_pc);
  const Eigen::Affine3d gt2 = pose_from(-0.2, 0.08, 0, 0, 0, -0.05);
  const pcl::PointCloud<pcl::PointXYZ> third_pc = transform_cloud(gt2, static_pc);
  dump(out + "/static.bin", static_pc);
  dump(out + "/moving.bin", moving_pc);
  dump(out + "/third.bin", third_pc);

  // ndt_odom_debug.cpp:168-192: two LazyGrid maps, the static one with setMapSize
  SpatialIndex *index1_lazzy = new LazyGrid(resolution);
  SpatialIndex *index2_lazzy = new LazyGrid(resolution);
  NDTMap ndt(index1_lazzy, true);
  ndt.setMapSize(80., 80., 2.);
  ndt.loadPointCloud(static_pc);
  ndt.computeNDTCells();
  NDTMap mov(index2_lazzy, true);
  mov.setMapSize(80., 80., 2.);
  mov.loadPointCloud(moving_pc);
  mov.computeNDTCells();
  NDTMap third(new LazyGrid(resolution), true);
  third.setMapSize(80., 80., 2.);
  third.loadPointCloud(third_pc);
  third.computeNDTCells();
  if (ndtb::last_status() != NDTB_OK) {
    std::fprintf(stderr, "map build failed: %s\n", ndtb_strerror(ndtb::last_status()));
    return 1;
  }

  NDTMatcherD2D matcher;
  matcher.n_neighbours = n_neighbours;
  Eigen::Affine3d T_d2d = odom_transform;
  const bool converged = matcher.match(ndt, mov, T_d2d, true);  // ndt_odom_debug.cpp:206
  Eigen::MatrixXd cov(6, 6);
  const bool cov_ok = matcher.covariance(ndt, mov, T_d2d, cov);  // ndt_feature_graph.cpp:298
  Eigen::MatrixXd g, H;
  const double score = matcher.derivativesNDT(mov, T_d2d, ndt, g, H, true);

  Eigen::MatrixXd Tcov(6, 6);
  for (int i = 0; i < 6; i++) Tcov(i, i) = i < 2 ? 0.01 : (i == 5 ? 0.001 : 1e-6);
  Eigen::Affine3d T_fusion = odom_transform;
  const bool fusion_ok = ndt_feature::matchFusion(ndt, mov, T_fusion, Tcov, true, true, 30, n_neighbours, 1e-6, true, false);

  lslgeneric::NDTMatcherP2D p2d;
  Eigen::Affine3d T_p2d = odom_transform;
  const bool p2d_ok = p2d.match(ndt, moving_pc, T_p2d, true);

  // graph edge refinement over three nodes (ndt_feature_graph.cpp:347-353)
  std::vector<NDTMap *> nodes = {&ndt, &mov, &third};
  ndtb::GraphRegistrar graph(nodes);
  std::vector<ndtb::Link> links(3);
  links[0].ref_idx = 0, links[0].mov_idx = 1, links[0].T = odom_transform;
  links[1].ref_idx = 0, links[1].mov_idx = 2, links[1].T = pose_from(0.15, 0.0, 0, 0, 0, 0);
  links[2].ref_idx = 1, links[2].mov_idx = 2, links[2].T = pose_from(900, 900, 0, 0, 0, 0);  // no overlap: pose unchanged -> 0.02 I
  const int rc = graph.updateLinksUsingNDTRegistration(links, n_neighbours, false);

  auto cells = ndt.pseudoTransformNDT(gt_transform);
  const size_t n_pseudo = cells.size();
  double pseudo_mean0[3] = {0, 0, 0};
  if (!cells.empty()) {
    const auto m = cells[0]->getMean();
    pseudo_mean0[0] = m(0), pseudo_mean0[1] = m(1), pseudo_mean0[2] = m(2);
  }
  for (auto *c : cells) delete c;  // caller deletes, as in ndt_matcher_d2d_fusion.h:953-962

  std::printf("{\n");
  std::printf("\"n_points\": %zu, \"cells_static\": %d, \"cells_moving\": %d, \"n_pseudo\": %zu,\n", static_pc.size(),
              ndt.numberOfActiveCells(), mov.numberOfActiveCells(), n_pseudo);
  std::printf("\"pseudo_mean0\": [%.17g, %.17g, %.17g],\n", pseudo_mean0[0], pseudo_mean0[1], pseudo_mean0[2]);
  print_pose("gt", gt_transform);
  print_pose("odom", odom_transform);
  print_pose("T_d2d", T_d2d);
  std::printf("\"converged\": %d, \"iterations\": %d, \"cov_ok\": %d, \"score\": %.17g,\n", (int)converged,
              matcher.iteration_counter_internal, (int)cov_ok, score);
  print_mat("cov", cov);
  print_mat("gradient", g);
  print_mat("hessian", H);
  print_pose("T_p2d", T_p2d);
  std::printf("\"p2d_ok\": %d, \"p2d_iterations\": %d,\n", (int)p2d_ok, p2d.last.iterations);
  print_pose("T_fusion", T_fusion);
  std::printf("\"fusion_ok\": %d, \"links_rc\": %d,\n", (int)fusion_ok, rc);
  for (int i = 0; i < 3; i++) {
    std::string n = "link" + std::to_string(i);
    print_pose((n + "_T").c_str(), links[i].T);
    print_mat((n + "_cov").c_str(), links[i].cov_3d);
    std::printf("\"%s_score\": %.17g, \"%s_converged\": %d,\n", n.c_str(), links[i].score, n.c_str(), (int)links[i].converged);
  }
  std::printf("\"status\": %d\n}\n", ndtb::last_status());
  return 0;
}
