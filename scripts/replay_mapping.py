#!/usr/bin/env python3
"""Replay of the reference's own bag `ndt_feature/data/mapping.bag` (tests/golden/mapping_bag.npz) through the complete
front end — NDTFeatureGraph::update -> NDTFeatureFuserHMT::update (local map, matchFusion, covariance, ray-traced
addPointCloud + computeNDTCells) with node spawning — and comparison of every node map with the maps the reference ships
(`FULL GRAPH/mapping{0..7}.jff`, tests/golden/full_graph.npz).

What is known about the shipped run (recovered from the fixtures): it started at scan 57, spawned its nodes at
scans 260, 434, 641, 831, 1019, 1189, 1319 (the interpolated /tf pose chains reproduce mapping{k}local_odom.T to 1e-15)
and processed only ~1 scan in 11 (the z-jitter rand() stream stood at 51548 points when node 7 was initialised: a live
node fed by `rosbag play` drops scans while it registers).  WHICH scans it processed in between is not recoverable, so
nodes 0..6 can only be compared statistically; node 7 (one scan) is reproduced exactly (tests/test_fuser_golden.py).

  --backend oracle|gpu     CPU oracle (oracle/fuser_oracle.py) or the engine (ndt_feature_graph_b200.fuser)
  --stride N               process every N-th scan between the known node boundaries (default 11)
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from ndt_feature_graph_b200 import laser as Ls  # noqa: E402

BOUNDS = [57, 260, 434, 641, 831, 1019, 1189, 1319]
SENSOR = (0.695, -0.01, -0.0069813)


def compare_node(cells, g, k):
    """cells: exported cell records of the replayed node map; g: full_graph.npz"""
    lin = (cells["idx"][:, 0].astype(np.int64) * 200 + cells["idx"][:, 1]) * 2 + cells["idx"][:, 2]
    has = cells["has_gaussian"] == 1
    mine, ref = set(lin[has].tolist()), set(g[f"gidx{k}"].tolist())
    common = sorted(mine & ref)
    pos = {int(l): i for i, l in enumerate(lin)}
    rpos = {int(l): i for i, l in enumerate(g[f"gidx{k}"].tolist())}
    dm = np.array([np.hypot(*(cells["mean"][pos[l]][:2] - g[f"mean{k}"][rpos[l]][:2])) for l in common]) if common else np.zeros(0)
    nr = np.array([cells["n"][pos[l]] / max(1, g[f"n{k}"][rpos[l]]) for l in common]) if common else np.zeros(0)
    occ_ref = dict(zip(g[f"occidx{k}"].tolist(), g[f"occ{k}"].tolist()))
    occ_mine = {int(l): float(o) for l, o in zip(lin, cells["occ"]) if o != 0}
    free_ref = {l for l, o in occ_ref.items() if o < 0}
    free_mine = {l for l, o in occ_mine.items() if o < 0}
    return {
        "node": k, "gauss_mine": len(mine), "gauss_ref": len(ref), "gauss_iou": len(common) / max(1, len(mine | ref)),
        "mean_xy_median_m": float(np.median(dm)) if dm.size else None, "mean_xy_p90_m": float(np.percentile(dm, 90)) if dm.size else None,
        "n_ratio_median": float(np.median(nr)) if nr.size else None,
        "free_mine": len(free_mine), "free_ref": len(free_ref), "free_iou": len(free_mine & free_ref) / max(1, len(free_mine | free_ref)),
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", default="oracle")
    ap.add_argument("--stride", type=int, default=11)
    ap.add_argument("--soft", type=int, default=1)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    d = np.load(os.path.join(ROOT, "tests", "golden", "mapping_bag.npz"))
    g = np.load(os.path.join(ROOT, "tests", "golden", "full_graph.npz"))
    track = Ls.TfTrack(d["odom_stamp"], d["odom"])
    st = d["stamp"]
    sensor = Ls.pose2d(*SENSOR)
    rng = np.random.default_rng(1)
    if a.backend == "oracle":
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import fuser_oracle as F

        fp = F.FuserParams(resolution=0.5, map_size_x=100, map_size_y=100, map_size_z=1.0, sensor_range=30.0, neighbours=2,
                           ITR_MAX=30, DELTA_SCORE=1e-6, globalTransf=False, useSoftConstraints=bool(a.soft),
                           useTikhonovRegularization=False)
        graph = F.GraphOracle(fp, sensor, F.MotionParams(Cd=1, Ct=1, Dd=1, Dt=1, Td=10, Tt=10), new_node_transl_dist=1e9)
    else:
        from ndt_feature_graph_b200 import fuser as GF

        graph = GF.make_graph(resolution=0.5, map_size=(100, 100, 1.0), sensor_range=30.0, neighbours=2, itr_max=30,
                              delta_score=1e-6, soft=bool(a.soft), tikhonov=False, sensor_pose=sensor,
                              motion=(1, 1, 1, 1, 10, 10), new_node_transl_dist=1e9)

    def cloud_of(i):
        return Ls.scan_to_cloud(d["ranges"][i], d["angle_min"], d["angle_inc"], d["range_min"], d["range_max"], 0.5, 0.02, rng)

    t0 = time.time()
    last = track.lookup(st[BOUNDS[0]])
    graph.initialize(last, cloud_of(BOUNDS[0]))
    n_reg = 0
    for k in range(7):
        lo, hi = BOUNDS[k], BOUNDS[k + 1]
        seq = list(range(lo + a.stride, hi, a.stride)) + [hi]
        for i in seq:
            P = track.lookup(st[i])
            Tm = np.linalg.inv(last) @ P
            if i != hi and np.linalg.norm(Tm[:3, 3]) < 0.02 and abs(Ls.yaw_of(Tm)) < 0.02:
                continue
            last = P
            if i == hi:
                graph.new_node_transl_dist = 0.0  # spawn exactly where the shipped run did
            graph.update(Tm, cloud_of(i))
            graph.new_node_transl_dist = 1e9
            n_reg += 1
    dt = time.time() - t0
    rep = {"backend": a.backend, "stride": a.stride, "soft": a.soft, "registrations": n_reg, "seconds": dt, "nodes": []}
    for k, node in enumerate(graph.nodes):
        cells = node.map.map.export_cells(False)
        r = compare_node(cells, g, k)
        r["T_err_m"] = float(np.hypot(*(node.T[:2, 3] - g[f"T{k}"][:2, 3])))
        if k < 7:
            r["Tfuse_err_m"] = float(np.hypot(*(node.Tlocal_fuse[:2, 3] - g[f"Tfuse{k}"][:2, 3])))
            r["Tfuse_err_yaw"] = float(abs(Ls.yaw_of(node.Tlocal_fuse) - Ls.yaw_of(g[f"Tfuse{k}"])))
        rep["nodes"].append(r)
        print(json.dumps(r))
    print(json.dumps({k: v for k, v in rep.items() if k != "nodes"}))
    if a.out:
        json.dump(rep, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
