import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ndt_feature_graph_b200 as N, oracle_py as O
from ndt_feature_graph_b200 import synth
O.lib(); eng = N.Engine(0)
scene = synth.velodyne_scene(4242)
poses = [synth.pose_from_xyzrpy(2.0 * k, 0.3 * np.sin(k), 1.8, 0, 0, 0.05 * k) for k in range(10)]
clouds = []
for k, T in enumerate(poses):
    c = synth.velodyne_scan(scene, T, 900 + k)
    w = np.zeros_like(c); w[:, :3] = (c[:, :3].astype(np.float64) @ T[:3, :3].T + T[:3, 3]).astype(np.float32); clouds.append(w)
allpts = np.concatenate(clouds)
center, size = (9.0, 0.0, 3.0), (160.0, 160.0, 14.0)
om = O.OracleMap(0.5); om.initialize(*center, *size); om.add_points(allpts); om.compute_cells()
gm = N.NDTMap(eng, 0.5); gm.initialize(*center, *size); gm.addPointCloud(allpts, want_count=False); gm.computeNDTCells()
Tq = synth.pose_from_xyzrpy(9.3, 0.4, 1.8, 0, 0, 0.21)
scan = synth.velodyne_scan(scene, Tq, 999)
T0 = synth.perturb_pose(Tq, 5, dt=0.03, dr=0.004)
p2d = N.NDTMatcherP2D(eng)
for T in (T0, Tq):
    so, go, Ho, no = O.p2d_derivatives(om, scan, T)
    sg, gg, Hg, ng = p2d.derivativesPointCloud(gm, scan, T)
    print("pairs", no, ng, "score", so, sg, "g rel", np.abs(go-gg).max()/np.abs(go).max(), "H rel", np.abs(Ho-Hg).max()/np.abs(Ho).max())
for itr in (0, 1, 2, 4, 8, 30):
    ro = O.p2d_match(om, scan, T0, O.default_params(itr_max=itr))
    m = N.NDTMatcherP2D(eng, itr_max=itr)
    rg = m.match(gm, scan, T0)
    print("itr_max", itr, "oracle it/h/g", ro.iterations, ro.n_hess_passes, ro.n_grad_passes, "gpu", rg.iterations, rg.n_hess_passes, rg.n_grad_passes, "err", synth.pose_error(ro.pose(), rg.pose()), "score", ro.score, rg.score)
