"""Shared set-up of the front-end tests: the reference's own run recovered from its shipped fixtures.

`ndt_feature/data/mapping.bag` (tests/golden/mapping_bag.npz) replayed by launch/henrik_replay_mapperbag_fuser.launch
produced `FULL GRAPH/mapping{0..7}.jff` (tests/golden/full_graph.npz).  What the fixtures pin (found by
tests/golden/make_mapping_bag.py + the searches described in DESIGN.md §4):
  * the run started at scan 57 and spawned its nodes at scans 260, 434, 641, 831, 1019, 1189, 1319: the /tf poses
    interpolated at those stamps chain to mapping{k}.T / mapping{k}local_odom.T within 1e-14;
  * the z jitter of the node (publish_graph_message.cpp:1377, glibc rand(), never seeded) stood at 51548 drawn values
    when the scan that initialised node 7 was converted: with that offset node 7's map is reproduced bit for bit.
"""
import os

import numpy as np

from ndt_feature_graph_b200 import laser as Ls

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BOUNDS = [57, 260, 434, 641, 831, 1019, 1189, 1319]
RAND_OFFSET_NODE7 = 51548
SENSOR = Ls.pose2d(0.695, -0.01, -0.0069813)  # launch/henrik_replay_mapperbag_fuser.launch sensor_pose_*
MOTION = dict(Cd=1, Ct=1, Dd=1, Dt=1, Td=10, Tt=10)


def bag():
    return np.load(os.path.join(ROOT, "tests", "golden", "mapping_bag.npz"))


def cloud_of(d, i, rng=None, varz=0.02):
    return Ls.scan_to_cloud(d["ranges"][i], d["angle_min"], d["angle_inc"], d["range_min"], d["range_max"], 0.5, varz, rng)


def node7_cloud(d):
    """the cloud node 7 was initialised with, z jitter included"""
    return cloud_of(d, BOUNDS[7], Ls.GlibcRand(skip=RAND_OFFSET_NODE7))


def lin_index(cells):
    return (cells["idx"][:, 0].astype(np.int64) * 200 + cells["idx"][:, 1]) * 2 + cells["idx"][:, 2]


def oracle_fuser_params(F, soft=False):
    return F.FuserParams(resolution=0.5, map_size_x=100, map_size_y=100, map_size_z=1.0, sensor_range=30.0, neighbours=2,
                         ITR_MAX=30, DELTA_SCORE=1e-6, globalTransf=False, useSoftConstraints=soft,
                         useTikhonovRegularization=False)


def gpu_fuser_params(engine, soft=False):
    from ndt_feature_graph_b200 import fuser as GF

    return GF.fuser_params(engine, sensor_pose=SENSOR, motion=(1, 1, 1, 1, 10, 10), resolution=0.5, map_size_x=100, map_size_y=100,
                           map_size_z=1.0, sensor_range=30.0, neighbours=2, itr_max=30, delta_score=1e-6, global_transf=0,
                           use_soft_constraints=int(soft), use_tikhonov=0, all_matches_valid=1)


def check_node7(cells, g):
    """cells (all cells of a map built by initialize() from node7_cloud) against the shipped mapping7.jff"""
    lin = lin_index(cells)
    # occupancy: every cell the reference touched, bit for bit (free-space ray trace + N ln 1.5)
    occ_ref = dict(zip(g["occidx7"].tolist(), g["occ7"].tolist()))
    occ_mine = {int(l): float(o) for l, o in zip(lin, cells["occ"]) if o != 0}
    assert set(occ_mine) == set(occ_ref)
    assert all(np.float32(occ_mine[k]) == np.float32(occ_ref[k]) for k in occ_ref)
    assert sum(1 for v in occ_ref.values() if v < 0) == 120
    pos = {int(l): i for i, l in enumerate(lin)}
    n_diff = 0
    for j, l in enumerate(g["gidx7"].tolist()):
        c = cells[pos[l]]
        assert c["n"] == g["n7"][j]
        assert np.array_equal(c["mean"], g["mean7"][j])  # bit-exact in x, y and z
        if c["has_gaussian"] != 1:
            # rank-deficient cell (3 points): the sign of a rounding-noise eigenvalue decides hasGaussian_; upstream's
            # Eigen::SelfAdjointEigenSolver and the Jacobi sweeps used here round differently
            assert c["n"] == 3
            n_diff += 1
            continue
        assert np.abs(c["cov"] - g["cov7"][j]).max() < 1e-15
    mine = set(lin[cells["has_gaussian"] == 1].tolist())
    extra = mine - set(g["gidx7"].tolist())
    assert all(cells[pos[l]]["n"] == 3 for l in extra)
    assert n_diff <= 1 and len(extra) <= 1
