// refine_edges.cpp — what the reference's back-end tool does with a saved graph (ndt_feature/src/ndt_feature_graph_opt.cpp:
// 142-144 → NDTFeatureGraph::updateLinksUsingNDTRegistration, ndt_feature_graph.cpp:347-353), written against the façade:
// load the node maps the fuser saved (mapping{k}.jff, ndt_feature_fuser_hmt.cpp:20-49), take the odometry between
// consecutive nodes as the initial link transforms (mapping{k}local_odom.T), refine every link on the GPU in one batched
// call and print pose, convergence, overlap score and the diagonal of the 6x6 covariance.
//
//   g++ -std=c++17 -O2 -I include examples/refine_edges.cpp -L ndt_feature_graph_b200/lib -lndtb -o refine_edges
//   ./refine_edges "<reference>/ndt_feature/data/FULL GRAPH/mapping" 8
#include <cstdio>
#include <fstream>
#include <memory>
#include <sstream>
#include <string>

#include "ndtb_lslgeneric.hpp"

// boost text archive of an Eigen::Affine3d as the reference writes it: the last 16 tokens are the column-major 4x4
static bool read_pose(const std::string &path, Eigen::Affine3d &T) {
  std::ifstream f(path);
  if (!f) return false;
  std::vector<std::string> tok;
  for (std::string s; f >> s;) tok.push_back(s);
  if (tok.size() < 16) return false;
  for (int i = 0; i < 16; i++) T.matrix().data()[i] = std::stod(tok[tok.size() - 16 + i]);
  return true;
}

int main(int argc, char **argv) {
  if (argc < 3) {
    std::fprintf(stderr, "usage: %s <prefix of mapping{k}.jff / mapping{k}local_odom.T> <number of nodes>\n", argv[0]);
    return 2;
  }
  const std::string prefix = argv[1];
  const int n_nodes = std::atoi(argv[2]);
  std::vector<std::unique_ptr<lslgeneric::NDTMap>> nodes;
  std::vector<lslgeneric::NDTMap *> maps;
  for (int k = 0; k < n_nodes; k++) {
    nodes.emplace_back(new lslgeneric::NDTMap(new lslgeneric::LazyGrid(0.5), true));
    const std::string file = prefix + std::to_string(k) + ".jff";
    if (nodes.back()->loadFromJFF(file.c_str()) != 0) {
      std::fprintf(stderr, "cannot load %s (%s)\n", file.c_str(), ndtb_strerror(ndtb::last_status()));
      return 1;
    }
    maps.push_back(nodes.back().get());
    std::printf("node %d: %d Gaussian cells\n", k, nodes.back()->numberOfActiveCells());
  }
  std::vector<ndtb::Link> links;
  for (int k = 0; k + 1 < n_nodes; k++) {
    ndtb::Link l;
    l.ref_idx = (size_t)k, l.mov_idx = (size_t)k + 1;
    if (!read_pose(prefix + std::to_string(k) + "local_odom.T", l.T)) l.T.setIdentity();
    links.push_back(l);
  }
  ndtb::GraphRegistrar graph(maps);
  const int rc = graph.updateLinksUsingNDTRegistration(links, /*nb_neighbours=*/2, /*keepScore=*/false);
  if (rc != NDTB_OK) {
    std::fprintf(stderr, "registration failed: %s\n", ndtb_strerror(rc));
    return 1;
  }
  for (const ndtb::Link &l : links) {
    const double yaw = std::atan2(l.T(1, 0), l.T(0, 0));
    std::printf("link %zu -> %zu: x %.4f y %.4f yaw %.4f  converged %d  score %.4f  cov diag", l.ref_idx, l.mov_idx, l.T(0, 3),
                l.T(1, 3), yaw, (int)l.converged, l.score);
    for (int i = 0; i < 6; i++) std::printf(" %.2e", l.cov_3d(i, i));
    std::printf("\n");
  }
  return 0;
}
