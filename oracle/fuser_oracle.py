"""CPU ORACLE of the front-end step (test infrastructure, NOT product code).

Restates, on top of the C oracle's map / matcher primitives (oracle/ndt_oracle.cpp):

  NDTFeatureFuserHMT::initialize   ndt_feature/src/ndt_feature_src/ndt_feature_fuser_hmt.cpp:65-102
  NDTFeatureFuserHMT::update       ndt_feature_fuser_hmt.cpp:108-512   (useFeat = useOdom = false: the NDT branch, the only
                                   one the shipped configurations enable — ndt_graph_offline.cpp:306-308,
                                   launch/henrik_replay_mapperbag_fuser.launch)
  NDTFeatureGraph::initialize      ndt_feature/src/ndt_feature_src/ndt_feature_graph.cpp:24-55
  NDTFeatureGraph::update          ndt_feature_graph.cpp:60-144 (node spawning every newNodeTranslDist metres)
  MotionModel2d::getCovMatrix6     ndt_feature/src/ndt_feature_src/motion_model.cpp:175-207
  lslgeneric::transformPointCloudInPlace [upstream perception_oru]: the cloud is moved by the pose cast to FLOAT.

Only tests/, scripts/ run as checkers and bench.py's CPU legs import this module.
"""
import math

import numpy as np

import oracle_py as O


def transform_cloud_f32(T, pts):
    """lslgeneric::transformPointCloudInPlace [upstream]: T.cast<float>() * p, evaluated column by column in float
    (Eigen's 3x3 matrix-vector product accumulates col0*x + col1*y + col2*z), then + translation."""
    Tf = np.asarray(T, dtype=np.float64).astype(np.float32)
    p = np.asarray(pts, dtype=np.float32)
    out = p.copy()
    x, y, z = p[:, 0], p[:, 1], p[:, 2]
    for r in range(3):
        acc = Tf[r, 0] * x
        acc = acc + Tf[r, 1] * y
        acc = acc + Tf[r, 2] * z
        out[:, r] = acc + Tf[r, 3]
    return out


def pmul(A, B):
    """A * B for 4x4 rigid poses with the operation order of the engine's host code (csrc/optimizer.h pose_mul): every
    product and sum separately rounded, left to right — numpy's matmul may use FMA kernels."""
    A = np.asarray(A, dtype=np.float64)
    B = np.asarray(B, dtype=np.float64)
    C = np.eye(4)
    for i in range(3):
        for j in range(3):
            C[i, j] = A[i, 0] * B[0, j] + A[i, 1] * B[1, j] + A[i, 2] * B[2, j]
        C[i, 3] = (A[i, 0] * B[0, 3] + A[i, 1] * B[1, 3] + A[i, 2] * B[2, 3]) + A[i, 3]
    return C


def pinv(A):
    """inverse of a rigid pose, order of csrc/optimizer.h pose_inverse"""
    A = np.asarray(A, dtype=np.float64)
    C = np.eye(4)
    C[:3, :3] = A[:3, :3].T
    for i in range(3):
        C[i, 3] = -(C[i, 0] * A[0, 3] + C[i, 1] * A[1, 3] + C[i, 2] * A[2, 3])
    return C


def euler_angles_xyz(R):
    """Eigen::Matrix3d::eulerAngles(0, 1, 2) (Eigen 3.2, the ROS-indigo era the reference targets)."""
    i, j, k = 0, 1, 2
    res = [0.0, 0.0, 0.0]
    res[0] = math.atan2(R[j, k], R[k, k])
    c2 = math.sqrt(R[i, i] * R[i, i] + R[i, j] * R[i, j])
    if res[0] > 0.0:
        res[0] = res[0] - math.pi
        res[1] = math.atan2(-R[i, k], -c2)
    else:
        res[1] = math.atan2(-R[i, k], c2)
    s1, c1 = math.sin(res[0]), math.cos(res[0])
    res[2] = math.atan2(s1 * R[k, i] - c1 * R[j, i], c1 * R[j, j] - s1 * R[k, j])
    return [-res[0], -res[1], -res[2]]


class MotionParams:
    """MotionModel2d::Params (motion_model.hpp:123-163)."""

    def __init__(self, Cd=0.001, Ct=0.001, Dd=0.005, Dt=0.005, Td=0.001, Tt=0.001):
        self.Cd, self.Ct, self.Dd, self.Dt, self.Td, self.Tt = Cd, Ct, Dd, Dt, Td, Tt


def motion_cov6(mp, Tmotion):
    """TmotionCov of NDTFeatureFuserHMT::update (fuser_hmt.cpp:125-143): getCovMatrix6(relpose), z/roll/pitch set to 1."""
    x, y = Tmotion[0, 3], Tmotion[1, 3]
    rot = euler_angles_xyz(Tmotion[:3, :3])[2]
    d2 = x * x + y * y
    cov = np.eye(6)
    cov[0, 0] = mp.Dd * d2 + mp.Dt * rot * rot
    cov[1, 1] = mp.Cd * d2 + mp.Ct * rot * rot
    cov[5, 5] = mp.Td * d2 + mp.Tt * rot * rot
    cov[2, 2] = cov[3, 3] = cov[4, 4] = 1.0
    return cov


class FuserParams:
    """NDTFeatureFuserHMT::Params (ndt_feature_fuser_hmt.h:58-207), the fields the NDT branch reads."""

    def __init__(self, **kw):
        self.resolution = 1.0
        self.map_size_x = 40.0
        self.map_size_y = 40.0
        self.map_size_z = 10.0
        self.sensor_range = 3.0
        self.neighbours = 0
        self.stepcontrol = True
        self.ITR_MAX = 30
        self.DELTA_SCORE = 10e-4
        self.globalTransf = True
        self.useSoftConstraints = True
        self.useTikhonovRegularization = True
        self.computeCov = True
        self.fusion2d = False
        for k, v in kw.items():
            assert hasattr(self, k), k
            setattr(self, k, v)


class FuserOracle:
    """NDTFeatureFuserHMT with useNDT only."""

    def __init__(self, params, sensor_pose=None, motion_params=None):
        self.p = params
        self.sensor_pose = np.eye(4) if sensor_pose is None else np.asarray(sensor_pose, dtype=np.float64)
        self.mp = motion_params or MotionParams()
        self.Tnow = np.eye(4)
        self.map = None
        self.last_result = None
        self.last_cov = None
        self.n_local_cells = 0

    def _match_params(self):
        return O.default_params(n_neighbours=self.p.neighbours, itr_max=self.p.ITR_MAX, step_control=int(self.p.stepcontrol),
                                delta_score=self.p.DELTA_SCORE, use_soft_constraints=int(self.p.useSoftConstraints),
                                use_tikhonov=int(self.p.useTikhonovRegularization), planar=int(self.p.fusion2d))

    def initialize(self, init_pose, cloud):
        """fuser_hmt.cpp:65-102"""
        c = transform_cloud_f32(self.sensor_pose, cloud)
        c = transform_cloud_f32(init_pose, c)
        self.Tnow = np.array(init_pose, dtype=np.float64)
        self.map = O.OracleMap(self.p.resolution)
        self.map.initialize(self.Tnow[0, 3], self.Tnow[1, 3], 0.0, self.p.map_size_x, self.p.map_size_y, self.p.map_size_z)
        origin = pmul(self.Tnow, self.sensor_pose)[:3, 3]
        self.map.add_point_cloud(origin, c, 0.1, 100.0, 0.1, 255.0)
        self.map.compute_cells(int(1e5), 255.0)

    def update(self, Tmotion, cloud, update_ndt_map=True):
        """fuser_hmt.cpp:108-512 with useFeat = useOdom = false, loadCentroid = false, checkConsistency = false,
        allMatchesValid / fuseIncomplete irrelevant (the pose estimate is used whatever match() returns only when
        allMatchesValid; otherwise a failed match falls back to odometry, :471-474)."""
        Tmotion = np.asarray(Tmotion, dtype=np.float64)
        Tcov = motion_cov6(self.mp, Tmotion)
        if self.p.globalTransf:
            Tinit = self.Tnow.copy()
            Tmotion_est = Tmotion.copy()
        else:
            Tinit = np.eye(4)
            Tmotion_est = pmul(self.Tnow, Tmotion)
        local = transform_cloud_f32(pmul(Tinit, self.sensor_pose), cloud)
        nd = O.OracleMap(self.p.resolution)
        if not self.p.globalTransf:
            nd.guess_size(0, 0, 0, self.p.sensor_range, self.p.sensor_range, self.p.map_size_z)
        nd.load_point_cloud(local, self.p.sensor_range)
        nd.compute_cells()
        self.n_local_cells = nd.num_cells(True)
        prm = self._match_params()
        # matchFusion without soft / Tikhonov terms still runs the fusion control flow (fusion.h:797-1155): score_best
        # starts at DBL_MAX and the Hessian is always regularised; Tcov is then never read (identity passed)
        soft_or_tik = self.p.useSoftConstraints or self.p.useTikhonovRegularization
        r = O.fusion_match(self.map, nd, Tmotion_est, Tcov if soft_or_tik else np.eye(6), prm)
        self.last_result = r
        Test = r.pose()
        if self.p.computeCov:
            rc, cov = O.d2d_covariance(self.map, nd, Test, prm)
            self.last_cov = cov if rc == 0 else None
        match_ok = True  # allMatchesValid in every shipped configuration
        if match_ok:
            self.Tnow = pmul(self.Tnow, Test) if self.p.globalTransf else Test
        else:
            self.Tnow = pmul(self.Tnow, Tmotion)
        if update_ndt_map:
            spose = pmul(self.Tnow, self.sensor_pose)
            world = transform_cloud_f32(spose, cloud)
            self.map.add_point_cloud(spose[:3, 3], world, 0.06, 25.0, 0.25, 255.0)
            self.map.compute_cells(int(1e5), 255.0)
        return self.Tnow.copy()


class GraphNode:
    def __init__(self, fuser, T):
        self.map = fuser
        self.T = np.array(T, dtype=np.float64)
        self.Tlocal_odom = np.eye(4)
        self.Tlocal_fuse = np.eye(4)
        self.nbUpdates = 0


class GraphOracle:
    """NDTFeatureGraph front end (ndt_feature_graph.cpp:24-144)."""

    def __init__(self, fuser_params, sensor_pose, motion_params=None, new_node_transl_dist=1.0):
        self.fp, self.sensor_pose, self.mp = fuser_params, sensor_pose, motion_params
        self.new_node_transl_dist = new_node_transl_dist
        self.nodes = []
        self.distance_moved_in_last_node = 0.0
        self.Tnow = np.eye(4)
        self.n_registrations = 0

    def _new_node(self, T, cloud):
        f = FuserOracle(self.fp, self.sensor_pose, self.mp)
        f.initialize(np.eye(4), cloud)  # every node map lives in its own frame (node.T)
        return GraphNode(f, T)

    def initialize(self, init_pose, cloud):
        self.nodes.append(self._new_node(init_pose, cloud))
        self.Tnow = np.array(init_pose, dtype=np.float64)

    def update(self, Tmotion, cloud):
        node = self.nodes[-1]
        self.distance_moved_in_last_node += float(np.linalg.norm(Tmotion[:3, 3]))
        self.n_registrations += 1
        if self.distance_moved_in_last_node > self.new_node_transl_dist:
            self.distance_moved_in_last_node = 0.0
            Tl = node.map.update(Tmotion, cloud, update_ndt_map=False)
            self.Tnow = pmul(node.T, Tl)
            node.Tlocal_odom = pmul(node.Tlocal_odom, Tmotion)
            node.Tlocal_fuse = Tl
            self.nodes.append(self._new_node(self.Tnow, cloud))
            return self.Tnow.copy()
        Tl = node.map.update(Tmotion, cloud)
        self.Tnow = pmul(node.T, Tl)
        node.Tlocal_odom = pmul(node.Tlocal_odom, Tmotion)
        node.Tlocal_fuse = Tl
        node.nbUpdates += 1
        return self.Tnow.copy()
