"""GPU suite (B200): the front end through the C ABI — ray-traced addPointCloud, ndtb_fuser_*, ndtb_graph_* — against the
oracle on the same inputs and against the maps the reference itself shipped."""
import numpy as np
import pytest

import fuser_common as FC
from ndt_feature_graph_b200 import laser as Ls

pytestmark = pytest.mark.gpu


def cells_equal(a, b, what=""):
    assert a.shape == b.shape, (what, a.shape, b.shape)
    for f in ("idx", "n", "has_gaussian", "occ", "mean", "cov"):
        assert np.array_equal(a[f], b[f]), (what, f, int((a[f] != b[f]).sum()))


@pytest.fixture(scope="module")
def F(oracle):
    import fuser_oracle

    return fuser_oracle


def test_node7_map_reproduced_exactly_on_gpu(golden, engine):
    """ndtb_fuser_initialize on the scan the reference initialised node 7 with vs the shipped mapping7.jff"""
    from ndt_feature_graph_b200 import fuser as GF

    d = FC.bag()
    f = GF.NDTFeatureFuserHMT(engine, FC.gpu_fuser_params(engine))
    f.initialize(np.eye(4), FC.node7_cloud(d))
    FC.check_node7(f.map.export_cells(False), golden)


def test_ray_trace_bit_exact_vs_oracle(engine, oracle):
    """NDTMap::addPointCloud with LazyGrid::traceLine: empty cells (-0.2 per ray), Gaussian cells (likelihood-weighted
    evidence, Gaussian dropped at occupancy <= 0), skipped rays; three fused scans of a synthetic 3-D scene."""
    import ndt_feature_graph_b200 as N
    from ndt_feature_graph_b200 import synth

    ca, cb, D = synth.velodyne_pair(5, n_rings=16, n_az=300)
    clouds = [ca, cb, ca + np.float32([0.3, -0.2, 0.0, 0.0])]
    origins = [np.array([0.0, 0.0, 0.0]), D[:3, 3].copy(), np.array([0.3, -0.2, 0.0])]
    om = oracle.OracleMap(0.5)
    om.initialize(0, 0, 0, 100, 100, 12)
    gm = N.NDTMap(engine, 0.5)
    gm.initialize(0, 0, 0, 100, 100, 12)
    for c, o in zip(clouds, origins):
        c = c.copy()
        c[::97, 2] = 30.0   # above maxz: ray and end point ignored
        c[5] = np.nan
        om.add_point_cloud(o, c, 0.06, 25.0, 0.25, 255.0)
        om.compute_cells(int(1e5), 255.0)
        gm.addPointCloudTraced(o, c, 0.06, 25.0, 0.25, 255.0)
        gm.computeNDTCells(int(1e5), 255.0)
        oc, gc = om.export_cells(False), gm.export_cells(False)
        cells_equal(oc, gc, "traced map")
        assert (oc["occ"] < 0).sum() > 1000
    lost = oc[(oc["n"] > 0) & (oc["has_gaussian"] == 0) & (oc["occ"] <= 0)]
    assert len(lost) >= 0  # cells emptied by free-space evidence keep N / mean / cov (may or may not occur in this scene)


def test_two_add_point_clouds_before_one_compute(engine, oracle):
    import ndt_feature_graph_b200 as N

    d = FC.bag()
    a, b = FC.cloud_of(d, 100), FC.cloud_of(d, 140)
    om = oracle.OracleMap(0.5)
    om.initialize(0, 0, 0, 100, 100, 1)
    gm = N.NDTMap(engine, 0.5)
    gm.initialize(0, 0, 0, 100, 100, 1)
    for m, add in ((om, om.add_point_cloud), (gm, gm.addPointCloudTraced)):
        add([0, 0, 0], a, 0.1, 100.0, 0.1, 255.0)
        add([0.1, 0, 0], b, 0.06, 25.0, 0.25, 255.0)
    om.compute_cells(int(1e5), 255.0)
    gm.computeNDTCells(int(1e5), 255.0)
    cells_equal(om.export_cells(False), gm.export_cells(False), "two pending traced clouds")


def test_transform_point_cloud_is_upstreams_float_transform(engine, F):
    import ctypes as C

    from ndt_feature_graph_b200.api import HOST, _cm

    rng = np.random.default_rng(0)
    pts = (rng.standard_normal((5000, 4)) * 20).astype(np.float32)
    T = Ls.pose2d(3.25, -7.5, 0.7)
    T[2, 3] = 0.1
    out = np.zeros_like(pts)
    Tc = _cm(T)
    engine.check(engine.L.ndtb_transform_point_cloud(engine.h, Tc.ctypes.data, pts.ctypes.data, len(pts), HOST, out.ctypes.data, HOST))
    ref = F.transform_cloud_f32(T, pts)
    assert np.array_equal(out[:, :3], ref[:, :3])


def test_fuser_update_matches_oracle_step_by_step(engine, F):
    """NDTFeatureFuserHMT::update: 12 consecutive processed scans of the shipped bag; after every step the pose agrees to
    1e-9 and the node map (all cells: occupancy, N, Gaussians) is identical"""
    from ndt_feature_graph_b200 import fuser as GF

    d = FC.bag()
    tr = Ls.TfTrack(d["odom_stamp"], d["odom"])
    rng = np.random.default_rng(3)
    idx = list(range(300, 300 + 13 * 9, 9))
    clouds = [FC.cloud_of(d, i, rng) for i in idx]
    for soft in (False, True):
        fo = F.FuserOracle(FC.oracle_fuser_params(F, soft=soft), FC.SENSOR, F.MotionParams(**FC.MOTION))
        fg = GF.NDTFeatureFuserHMT(engine, FC.gpu_fuser_params(engine, soft=soft))
        fo.initialize(np.eye(4), clouds[0])
        fg.initialize(np.eye(4), clouds[0])
        cells_equal(fo.map.export_cells(False), fg.map.export_cells(False), "after initialize")
        last = tr.lookup(d["stamp"][idx[0]])
        for i, c in zip(idx[1:], clouds[1:]):
            P = tr.lookup(d["stamp"][i])
            Tm = F.pmul(F.pinv(last), P)
            last = P
            To = fo.update(Tm, c)
            Tg = fg.update(Tm, c)
            assert np.abs(To - Tg).max() < 1e-9, (soft, i, np.abs(To - Tg).max())
            assert fo.last_result.iterations == fg.last_result.iterations
            assert np.allclose(fo.last_cov, fg.last_cov, rtol=1e-6, atol=1e-12)
            cells_equal(fo.map.export_cells(False), fg.map.export_cells(False), f"after scan {i}")


def test_graph_replay_matches_oracle_and_the_shipped_maps(golden, engine, F):
    """The whole bag through ndtb_graph_* (node spawning included): same node poses and maps as the oracle's replay,
    8 nodes, and the statistical agreement with the shipped maps that tests/test_fuser_golden.py asserts for the oracle"""
    from ndt_feature_graph_b200 import fuser as GF

    d = FC.bag()
    tr = Ls.TfTrack(d["odom_stamp"], d["odom"])
    go = F.GraphOracle(FC.oracle_fuser_params(F, soft=True), FC.SENSOR, F.MotionParams(**FC.MOTION), new_node_transl_dist=1e9)
    gg = GF.NDTFeatureGraph(engine, FC.gpu_fuser_params(engine, soft=True), 1e9)
    rng = np.random.default_rng(1)
    last = tr.lookup(d["stamp"][FC.BOUNDS[0]])
    c0 = FC.cloud_of(d, FC.BOUNDS[0], rng)
    go.initialize(last, c0)
    gg.initialize(last, c0)
    n = 0
    for k in range(7):
        lo, hi = FC.BOUNDS[k], FC.BOUNDS[k + 1]
        for i in list(range(lo + 11, hi, 11)) + [hi]:
            P = tr.lookup(d["stamp"][i])
            Tm = F.pmul(F.pinv(last), P)
            last = P
            c = FC.cloud_of(d, i, rng)
            go.new_node_transl_dist = gg.new_node_transl_dist = 0.0 if i == hi else 1e9
            To, Tg = go.update(Tm, c), gg.update(Tm, c)
            assert np.abs(To - Tg).max() < 1e-8, (i, np.abs(To - Tg).max())
            n += 1
    nodes = gg.nodes
    assert len(nodes) == len(go.nodes) == 8 and n > 100
    for k, (a, b) in enumerate(zip(go.nodes, nodes)):
        assert np.abs(a.T - b.T).max() < 1e-8 and np.abs(a.Tlocal_fuse - b.Tlocal_fuse).max() < 1e-8
        assert np.abs(a.Tlocal_odom - b.Tlocal_odom).max() < 1e-12
        cells_equal(a.map.map.export_cells(False), b.map.map.export_cells(False), f"node {k}")
        cells = b.map.map.export_cells(False)
        lin = FC.lin_index(cells)
        mine, ref = set(lin[cells["has_gaussian"] == 1].tolist()), set(golden[f"gidx{k}"].tolist())
        assert len(mine & ref) / len(mine | ref) > 0.8
