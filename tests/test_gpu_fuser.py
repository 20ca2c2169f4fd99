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


def cells_close(a, b, what=""):
    """Maps fused from clouds that were moved by poses agreeing to ~1e-9 (not bitwise): a point may round to the other
    float, so single points may change cell and a rank-deficient 3-point cell may flip hasGaussian_."""
    ka = {tuple(i): k for k, i in enumerate(a["idx"].tolist())}
    kb = {tuple(i): k for k, i in enumerate(b["idx"].tolist())}
    common = sorted(set(ka) & set(kb))
    assert len(common) >= 0.99 * max(len(ka), len(kb)), (what, len(ka), len(kb))
    ia, ib = np.array([ka[c] for c in common]), np.array([kb[c] for c in common])
    same_n = a["n"][ia] == b["n"][ib]
    assert same_n.mean() > 0.98, (what, same_n.mean())
    assert (np.abs(a["occ"][ia] - b["occ"][ib]) < 1e-3).mean() > 0.98, what
    both = same_n & (a["has_gaussian"][ia] == 1) & (b["has_gaussian"][ib] == 1)
    assert np.abs(a["mean"][ia][both] - b["mean"][ib][both]).max() < 1e-6, what
    flips = (a["has_gaussian"][ia] != b["has_gaussian"][ib]) & same_n
    assert (a["n"][ia][flips] <= 4).all() and flips.sum() <= 3, (what, int(flips.sum()))


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
    """NDTFeatureFuserHMT::update on 12 processed scans of the shipped bag, with and without the soft odometry prior.
    Before every step the engine's fuser is given the oracle's state (node map cell by cell, pose): from identical
    inputs the step must return the same pose (1e-8; a registration with DELTA_SCORE 1e-6 amplifies summation-order
    differences) and, whenever the pose is bitwise the same, a bit-identical updated map.  Left alone, the two chains
    drift apart in z / roll / pitch, which a planar scan hardly constrains: one rank-deficient 3-point cell flipping
    hasGaussian_ moves the next pose by 0.4 mm in z (measured), so a chain comparison would test chaos, not parity."""
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
        n_bitwise = 0
        for i, c in zip(idx[1:], clouds[1:]):
            P = tr.lookup(d["stamp"][i])
            Tm = F.pmul(F.pinv(last), P)
            last = P
            center, cell, size = fo.map.grid()
            state = fo.map.export_cells(False)
            fg.map.from_cells(center, cell, size, state, use_idx=True)
            fo.map.from_cells(center, cell, size, state, use_idx=True)  # (the exported record holds the upper triangle of
            fg.Tnow = fo.Tnow                                            # a covariance whose two halves may differ by an ulp)
            To = fo.update(Tm, c)
            Tg = fg.update(Tm, c)
            assert np.abs(To - Tg).max() < 1e-8, (soft, i, np.abs(To - Tg).max())
            assert fo.last_result.iterations == fg.last_result.iterations
            assert np.allclose(fo.last_cov, fg.last_cov, rtol=1e-5, atol=1e-12)
            # the poses agree to ~1e-15, not bitwise; cast to float for the cloud transform they almost always round to the
            # same matrix, and then the updated maps must be bit-identical
            a, b = fo.map.export_cells(False), fg.map.export_cells(False)
            try:
                cells_equal(a, b, f"after scan {i}")
                n_bitwise += 1
            except AssertionError:
                cells_close(a, b, f"after scan {i}")
        assert n_bitwise >= 3, n_bitwise  # (typically 8-12 of 12; a pose difference of 1e-9 already moves a float coefficient)


def test_graph_replay_matches_oracle_and_the_shipped_maps(golden, engine, F):
    """The whole bag through ndtb_graph_* (node spawning included): 8 nodes, node poses next to the oracle's replay, and
    the statistical agreement with the shipped maps that tests/test_fuser_golden.py asserts for the oracle"""
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
            # two free-running chains: x, y, yaw stay together, z / roll / pitch of a planar scan drift (see above)
            assert np.hypot(*(To[:2, 3] - Tg[:2, 3])) < 0.02 and abs(Ls.yaw_of(To) - Ls.yaw_of(Tg)) < 5e-3, i
            n += 1
    nodes = gg.nodes
    assert len(nodes) == len(go.nodes) == 8 and n > 100
    for k, (a, b) in enumerate(zip(go.nodes, nodes)):
        assert np.hypot(*(a.T[:2, 3] - b.T[:2, 3])) < 0.02
        assert np.abs(a.Tlocal_odom - b.Tlocal_odom).max() < 1e-12
        cells = b.map.map.export_cells(False)
        lin = FC.lin_index(cells)
        mine, ref = set(lin[cells["has_gaussian"] == 1].tolist()), set(golden[f"gidx{k}"].tolist())
        assert len(mine & ref) / len(mine | ref) > 0.8
    # hand-off: the graph as ndt_feature/NDTGraphMsg bytes (NDTGraphToMsg) and back (msgToNDTGraph): every node map that
    # comes out of the message holds the Gaussian cells of the resident map, at the same voxels
    from ndt_feature_graph_b200 import api
    import ndt_feature_graph_b200 as N

    node_msgs = []
    for b in nodes:
        cen, cs, sz = b.map.map.grid()
        g = api.Grid((api.C.c_double * 3)(*cen), (api.C.c_double * 3)(*cs), (api.C.c_int32 * 3)(*[int(s) for s in sz]))
        f = api.NodeFields.make(T=b.T, Tlocal_odom=b.Tlocal_odom, Tlocal_fuse=b.Tlocal_fuse, Tnow=b.Tlocal_fuse, nb_updates=b.nbUpdates)
        node_msgs.append(api.node_msg_pack(f, api.map_msg_pack(g, b.map.map.export_cells(False))))
    edge_msgs = [api.edge_msg_pack(k, k + 1, nodes[k].Tlocal_fuse, np.eye(3), None, 0.0) for k in range(7)]
    msg = api.graph_msg_pack(FC.SENSOR, Tg, 0.0, node_msgs, edge_msgs)
    back = api.graph_msg_unpack(msg)
    assert len(back["nodes"]) == 8 and len(back["edges"]) == 7 and np.abs(back["Tnow"] - Tg).max() < 1e-9
    for b, nm in zip(nodes, back["nodes"]):
        f, mm, _ = api.node_msg_unpack(nm)
        assert np.abs(f.pose("T") - b.T).max() < 1e-9 and f.nb_updates == b.nbUpdates
        g, cells, frame, _, _ = api.map_msg_unpack(mm)
        m2 = N.NDTMap(engine, float(g.cell[0])).from_cells(list(g.center), list(g.cell), list(g.size), cells, use_idx=False)
        a0, a1 = b.map.map.export_cells(True), m2.export_cells(True)
        assert len(a0) == len(a1) > 20
        o0, o1 = np.argsort(FC.lin_index(a0)), np.argsort(FC.lin_index(a1))
        for fld in ("idx", "mean", "cov", "n"):
            assert np.array_equal(a0[fld][o0], a1[fld][o1]), fld
