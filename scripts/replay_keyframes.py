#!/usr/bin/env python3
"""Front-end replay on real data (BASELINE config C5 in miniature): the 160 laser keyframes
extracted from the reference's Kyl1.bag (tests/golden/kyl1_scans.npz) are registered pair by pair from the odometry
guess — the registration part of ndt_graph_offline's loop (ndt_offline_ndt_feature/src/ndt_graph_offline.cpp:479-672 →
NDTFeatureFuserHMT::update) — on the GPU in one batched call and on the CPU oracle, and the chained trajectories are
compared.  usage: python scripts/replay_keyframes.py"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ndt_feature_graph_b200 as N  # noqa: E402
import oracle_py as O  # noqa: E402
from ndt_feature_graph_b200 import synth  # noqa: E402

z = np.load(os.path.join(ROOT, "tests", "golden", "kyl1_scans.npz"))
ang = z["angle_min"] + z["angle_inc"] * np.arange(z["ranges"].shape[1])
clouds = []
for k, r in enumerate(z["ranges"]):
    ok = np.isfinite(r) & (r > max(z["range_min"], 0.05)) & (r < z["range_max"] - 1e-3)
    c = np.zeros((int(ok.sum()), 4), np.float32)
    c[:, 0], c[:, 1] = r[ok] * np.cos(ang[ok]), r[ok] * np.sin(ang[ok])
    c[:, 2] = np.random.default_rng(500 + k).uniform(0.0, 1.0, c.shape[0]) * 0.02  # publish_graph_message.cpp:1373-1381
    clouds.append(c)
poses = [synth.pose2d(*p) for p in z["odom"]]
n = len(clouds) - 1
T0s = [np.linalg.inv(poses[k]) @ poses[k + 1] for k in range(n)]

eng = N.Engine(0)
kw = dict(cell=0.5, map_size=(40.0, 40.0, 1.0), range_limit=16.0, with_covariance=True)
eng.register_scans(clouds[:n], clouds[1:], T0s, **kw)  # warm-up
t0 = time.perf_counter()
res, cov = eng.register_scans(clouds[:n], clouds[1:], T0s, **kw)
t_gpu = time.perf_counter() - t0

O.lib()
t0 = time.perf_counter()
ores = []
for k in range(n):
    om = []
    for c in (clouds[k], clouds[k + 1]):
        m = O.OracleMap(0.5)
        m.set_map_size(40.0, 40.0, 1.0)
        m.load_point_cloud(c, 16.0)
        m.compute_cells()
        om.append(m)
    r = O.d2d_match(om[0], om[1], T0s[k])
    if r.pose_changed:
        O.d2d_covariance(om[0], om[1], r.pose())
    ores.append(r)
t_cpu = time.perf_counter() - t0
err = np.array([synth.pose_error(ores[k].pose(), res["T"][k].reshape(4, 4).T) for k in range(n)])
Tg, To, Tod = np.eye(4), np.eye(4), np.eye(4)
for k in range(n):
    Tg, To, Tod = Tg @ res["T"][k].reshape(4, 4).T, To @ ores[k].pose(), Tod @ T0s[k]
print(json.dumps({
    "workload": "159 consecutive keyframe pairs of ndt_feature/data/Kyl1.bag (361-ray laser, 0.5 m cells), map build x2 + match + covariance",
    "gpu_ms_batch_host_buffers": 1e3 * t_gpu, "gpu_pairs_per_s": n / t_gpu,
    "cpu_oracle_ms_1_thread": 1e3 * t_cpu, "cpu_pairs_per_s": n / t_cpu,
    "converged": int(res["converged"].sum()), "pose_err_vs_oracle_max": float(err.max()),
    "end_pose_gpu_xyyaw": [float(Tg[0, 3]), float(Tg[1, 3]), float(synth.robust_yaw(Tg))],
    "end_pose_oracle_xyyaw": [float(To[0, 3]), float(To[1, 3]), float(synth.robust_yaw(To))],
    "end_pose_odometry_xyyaw": [float(Tod[0, 3]), float(Tod[1, 3]), float(synth.robust_yaw(Tod))]}))
