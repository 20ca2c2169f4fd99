// callsites.cpp — the reference's own call sites of the NDT hot path, compiled against the façade
// (include/ndtb_lslgeneric.hpp) to prove that its L3/L4 sources keep compiling when perception_oru's ndt_map /
// ndt_registration headers are swapped for it.  The statements marked [ref file:line] are the reference's call
// expressions token for token (declarations around them are scaffolding with the reference's member names):
//   ndt_feature/src/ndt_feature_src/ndt_feature_fuser_hmt.cpp:87-94,195-227,352-358,399-405,485-486
//   ndt_feature/src/ndt_feature_src/ndt_feature_graph.cpp:261-298
//   ndt_feature/include/ndt_feature/ndt_matcher_d2d_fusion.h:810-814,840,856,1013,154-165,347
// Run (GPU box): prints the poses so that tests/test_facade.py can compare them with the oracle.
#include <cstdio>
#include <cstdlib>
#include <utility>
#include <vector>

#include "ndtb_lslgeneric.hpp"

namespace ndt_feature {
struct Params {  // NDTFeatureFuserHMT::Params (ndt_feature_fuser_hmt.h:58-207), the fields these call sites read
  double resolution = 0.5, map_size_x = 100, map_size_y = 100, map_size_z = 1.0, sensor_range = 30.;
  bool useNDT = true, stepcontrol = true, globalTransf = false, fusion2d = false, fuseIncomplete = false;
  int ITR_MAX = 30, neighbours = 2;
  double DELTA_SCORE = 1e-6;
  bool useSoftConstraints = false, stepControlFusion = true, useTikhonovRegularization = false;
};
struct NDTFeatureNode {
  lslgeneric::NDTMap *m;
  lslgeneric::NDTMap &getNDTMap() { return *m; }
};
struct NDTFeatureLink {
  size_t ref = 0, mov = 1;
  Eigen::Affine3d T;
  size_t getRefIdx() const { return ref; }
  size_t getMovIdx() const { return mov; }
};
}  // namespace ndt_feature

static pcl::PointCloud<pcl::PointXYZ> corridor(double shift, unsigned seed) {
  pcl::PointCloud<pcl::PointXYZ> pc;
  srand(seed);
  for (int i = 0; i < 1800; i++) {
    const double a = -2.3 + 4.6 * i / 1800.0;
    double r = 6.0 + 2.0 * std::sin(3 * a) + 0.5 * std::cos(11 * a);
    const double n = 0.01 * (rand() / (double)RAND_MAX - 0.5);
    pc.push_back(pcl::PointXYZ((float)((r + n) * std::cos(a) - shift), (float)((r + n) * std::sin(a)), (float)(0.02 * rand() / (double)RAND_MAX)));
  }
  return pc;
}

int main() {
  ndt_feature::Params params_;
  Eigen::Affine3d Tnow;
  Tnow.setIdentity();
  pcl::PointCloud<pcl::PointXYZ> cloud = corridor(0.0, 1), cloud2 = corridor(0.15, 2);
  lslgeneric::NDTMap *map;
  // ---- NDTFeatureFuserHMT::initialize
  map = new lslgeneric::NDTMap(new lslgeneric::LazyGrid(params_.resolution));                                                  // [ref fuser_hmt.cpp:87]
  map->initialize(Tnow.translation()(0),Tnow.translation()(1),0./*Tnow.translation()(2)*/,params_.map_size_x,params_.map_size_y,params_.map_size_z);  // [ref :89]
  Eigen::Affine3d Tnow_sensor = Tnow;
  map->addPointCloud(Tnow_sensor.translation(),cloud, 0.1, 100.0, 0.1);                                                        // [ref :92]
  map->computeNDTCells(CELL_UPDATE_MODE_SAMPLE_VARIANCE, 1e5, 255, Tnow_sensor.translation(), 0.1);                            // [ref :94]
  // ---- NDTFeatureFuserHMT::update: local map
  lslgeneric::SpatialIndex* ndglobal_idx = new lslgeneric::LazyGrid(params_.resolution);                                       // [ref :195]
  lslgeneric::NDTMap ndglobal(ndglobal_idx, true);                                                                             // [ref :196]
  ndglobal.guessSize(0,0,0,params_.sensor_range,params_.sensor_range,params_.map_size_z);                                      // [ref :222]
  ndglobal.loadPointCloud(cloud2, params_.sensor_range);                                                                       // [ref :225]
  ndglobal.computeNDTCells(CELL_UPDATE_MODE_SAMPLE_VARIANCE);                                                                  // [ref :227]
  // ---- feature / odometry maps (empty: useFeat = useOdom = false)
  lslgeneric::CellVector* cv_prev_sensor_frame = new lslgeneric::CellVector();                                                 // [ref :281]
  lslgeneric::CellVector* cv_curr_sensor_frame = new lslgeneric::CellVector();                                                 // [ref :282]
  lslgeneric::NDTMap ndt_feat_prev_sensor_frame(cv_prev_sensor_frame, true);                                                   // [ref :291]
  lslgeneric::NDTMap ndt_feat_curr_sensor_frame(cv_curr_sensor_frame, true);                                                   // [ref :292]
  Eigen::Affine3d sensor_pose;
  sensor_pose.setIdentity();
  lslgeneric::NDTMap* ndt_feat_prev = ndt_feat_prev_sensor_frame.pseudoTransformNDTMap(sensor_pose);
  lslgeneric::NDTMap* ndt_feat_curr = ndt_feat_curr_sensor_frame.pseudoTransformNDTMap(sensor_pose);
  std::vector<std::pair<int, int> > corr;
  bool use_odom_or_features = false;
  Eigen::Affine3d Tmotion_est;
  Tmotion_est.setIdentity();
  Eigen::MatrixXd TmotionCov(6, 6);
  for (int i = 0; i < 6; i++) TmotionCov(i, i) = 1.0;
  bool match_ok = true;
  if (params_.fusion2d) {
        match_ok = ndt_feature::matchFusion2d(*map, ndglobal, *ndt_feat_prev, *ndt_feat_curr, corr, Tmotion_est, true, params_.useNDT, use_odom_or_features, params_.stepcontrol, params_.ITR_MAX, params_.neighbours, params_.DELTA_SCORE) || params_.fuseIncomplete;  // [ref :353]
  }
  else {
        match_ok = ndt_feature::matchFusion(*map, ndglobal, *ndt_feat_prev, *ndt_feat_curr, corr, Tmotion_est, TmotionCov,
                                           true, params_.useNDT, use_odom_or_features, params_.stepcontrol, params_.ITR_MAX, params_.neighbours, params_.DELTA_SCORE, params_.useSoftConstraints, params_.stepControlFusion, params_.useTikhonovRegularization) || params_.fuseIncomplete;  // [ref :356-357]
  }
  std::printf("matchFusion %d %.12f %.12f %.12f\n", (int)match_ok, Tmotion_est(0, 3), Tmotion_est(1, 3), Tmotion_est(0, 0));
  {
          lslgeneric::NDTMatcherD2D matcher_d2d;                                                                              // [ref :403]
            Eigen::MatrixXd matching_cov(6,6);                                                                                 // [ref :404]
            matcher_d2d.covariance(*map, ndglobal, Tmotion_est, matching_cov);                                                 // [ref :405]
    std::printf("covariance %.6e %.6e\n", matching_cov(0, 0), matching_cov(5, 5));
  }
  Eigen::Affine3d spose = Tmotion_est;
        map->addPointCloud(spose.translation(),cloud2, 0.06, 25);                                                              // [ref :485]
        map->computeNDTCells(CELL_UPDATE_MODE_SAMPLE_VARIANCE, 1e5, 255, spose.translation(), 0.1);                            // [ref :486]
  std::printf("cells %d\n", map->numberOfActiveCells());

  // ---- NDTFeatureGraph::updateLinkUsingNDTRegistration
  std::vector<ndt_feature::NDTFeatureNode> nodes_ = {{map}, {&ndglobal}};
  ndt_feature::NDTFeatureLink link;
  link.T.setIdentity();
  int nb_neighbours = 2;
    lslgeneric::NDTMatcherD2D matcher_d2d;                                                                                    // [ref graph.cpp:261]
    matcher_d2d.n_neighbours = nb_neighbours;                                                                                  // [ref :262]
	Eigen::Affine3d before_T = link.T;                                                                                         // [ref :271]
    bool converged = matcher_d2d.match(nodes_[link.getRefIdx()].getNDTMap(), nodes_[link.getMovIdx()].getNDTMap(), link.T, true);  // [ref :273]
	Eigen::MatrixXd cov(6,6);                                                                                                  // [ref :284]
	bool same = true;
	for(size_t i = 0; i < 4 ; ++i){
		for(size_t  j = 0 ; j < 4 ; ++j){
			if(before_T(i,j) != link.T(i,j)){                                                                                  // [ref :290]
				same = false;
			}
		}
	}
	if(same == false){
		cov.setZero();                                                                                                         // [ref :297]
		matcher_d2d.covariance(nodes_[link.getRefIdx()].getNDTMap(), nodes_[link.getMovIdx()].getNDTMap(), link.T, cov);       // [ref :298]
	}
  std::printf("link %d %d %.12f %.12f %.6e\n", (int)converged, (int)same, link.T(0, 3), link.T(1, 3), cov(0, 0));

  // ---- the optimiser's building blocks as matchFusion uses them (ndt_matcher_d2d_fusion.h)
  {
    Eigen::Affine3d T;
    T.setIdentity();
    int n_neighbours = 2;
    lslgeneric::NDTMap &targetNDT = *map, &sourceNDT = ndglobal;
    const std::vector<std::pair<int, int> > &corr_feat = corr;
  lslgeneric::NDTMatcherD2D matcher_d2d;                                                                                      // [ref fusion.h:810]
  lslgeneric::NDTMatcherFeatureD2D matcher_feat_d2d(corr_feat);                                                                // [ref :811]
  matcher_d2d.n_neighbours = n_neighbours;                                                                                     // [ref :813]
    std::vector<lslgeneric::NDTCell*> nextNDT = sourceNDT.pseudoTransformNDT(T);                                               // [ref :840]
    Eigen::MatrixXd score_gradient_ndt(6,1), Hessian_ndt(6,6);
      double score_here_ndt = matcher_d2d.derivativesNDT(nextNDT,targetNDT,score_gradient_ndt,Hessian_ndt,true);               // [ref :856]
    Eigen::Matrix<double,6,1> pose_increment_v;
    for (int i = 0; i < 6; i++) pose_increment_v(i) = -1e-3 * score_gradient_ndt(i, 0);
        double step_size_ndt = matcher_d2d.lineSearchMT(pose_increment_v,nextNDT,targetNDT);                                   // [ref :1013]
    double stx = 0, sty = 1, stp = 0.5, stmin, stmax;
            stmin = lslgeneric::NDTMatcherD2D::MoreThuente::min(stx, sty);                                                     // [ref :154]
            stmax = lslgeneric::NDTMatcherD2D::MoreThuente::max(stx, sty);                                                     // [ref :155]
    double fxm = 0, dgxm = -1, fym = 0.1, dgym = 0.2, fm = -0.2, dgm = -0.1;
    bool brackt = true;
            int infoc = lslgeneric::NDTMatcherD2D::MoreThuente::cstep(stx,fxm,dgxm,sty,fym,dgym,stp,fm,dgm,
                        brackt,stmin,stmax);                                                                                   // [ref :347-348]
    std::printf("blocks %.12f %.6f %d %zu\n", score_here_ndt, step_size_ndt, infoc, nextNDT.size());
    for (unsigned int i = 0; i < nextNDT.size(); i++) delete nextNDT[i];                                                       // caller deletes (:953-962)
  }
  delete ndt_feat_prev;
  delete ndt_feat_curr;
  delete map;
  return 0;
}
