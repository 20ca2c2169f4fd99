"""Synthetic workloads of BASELINE.json configs C4 and C5, shared by bench.py, the scripts and the tests.

  C4  256 graph-edge D2D registrations between 64 resident node maps: nodes = C2-shaped scans (100k points) taken along a
      loop through one seeded scene, edges = consecutive + second neighbours + seeded loop closures, every edge with an
      odometry-noised initial pose (the input of NDTFeatureGraph::updateLinksUsingNDTRegistration, ndt_feature_graph.cpp:347-353)
  C5  the front end of ndt_offline_ndt_feature on a 2000-scan trajectory (ndt_graph_offline.cpp:479-672): a planar laser
      (541 rays over 270 deg like the shipped mapping.bag) driven along a closed path in a seeded room, with drifting odometry
Scans are generated once per (config, seed) and cached under the system temp directory.
"""
import os
import pickle
import tempfile

import numpy as np

from . import synth


def _cache(name, make):
    path = os.path.join(tempfile.gettempdir(), f"ndtb_workload_{name}.pkl")
    if os.path.exists(path):
        try:
            with open(path, "rb") as f:
                return pickle.load(f)
        except Exception:
            pass
    out = make()
    try:
        tmp = path + f".{os.getpid()}"
        with open(tmp, "wb") as f:
            pickle.dump(out, f)
        os.replace(tmp, path)
    except Exception:
        pass
    return out


def _c4_scan(args):
    seed, k, T = args
    return synth.velodyne_scan(synth.velodyne_scene(seed), T, 3000 + k)


def c4_graph(n_nodes=64, n_edges=256, seed=777, workers=None):
    """Returns (node clouds, edges [(ref, mov)], initial poses T0, true relative poses D)."""

    def make():
        from concurrent.futures import ProcessPoolExecutor

        poses = [synth.pose_from_xyzrpy(25 * np.cos(2 * np.pi * k / n_nodes), 25 * np.sin(2 * np.pi * k / n_nodes), 1.8, 0, 0,
                                        2 * np.pi * k / n_nodes + np.pi / 2) for k in range(n_nodes)]
        w = workers or max(1, min(16, len(os.sched_getaffinity(0))))
        with ProcessPoolExecutor(max_workers=w) as ex:
            clouds = list(ex.map(_c4_scan, [(seed, k, T) for k, T in enumerate(poses)]))
        rng = np.random.default_rng(5)
        edges = [(k, (k + 1) % n_nodes) for k in range(n_nodes)] + [(k, (k + 2) % n_nodes) for k in range(n_nodes)]
        while len(edges) < n_edges:
            a = int(rng.integers(n_nodes))
            edges.append((a, (a + int(rng.integers(1, 4))) % n_nodes))
        edges = edges[:n_edges]
        Ds = [np.linalg.inv(poses[a]) @ poses[b] for a, b in edges]
        T0s = [synth.odometry_guess(D, 50 + i) for i, D in enumerate(Ds)]
        return clouds, edges, T0s, Ds

    return _cache(f"c4_{n_nodes}_{n_edges}_{seed}", make)


def c5_trajectory(n_scans=2000, seed=4100, n_rays=541, fov_deg=270.0, step=0.06):
    """Returns (clouds in the sensor frame, true poses, odometry increments Tm[i] from scan i-1 to scan i).
    The vehicle follows a rounded rectangle inside the room at `step` metres per scan (keyframes of the offline driver are
    0.2 m apart; 0.06 m is the spacing of consecutive scans of the shipped bags at their driving speed)."""

    def make():
        scene = synth.room2d_scene(seed, half=18.0)
        # closed path: superellipse |x/a|^4 + |y/b|^4 = 1, traversed at constant arc length
        a, b = 11.0, 9.0
        u = np.linspace(0, 2 * np.pi, 20001)
        px = a * np.sign(np.cos(u)) * np.abs(np.cos(u)) ** 0.5
        py = b * np.sign(np.sin(u)) * np.abs(np.sin(u)) ** 0.5
        seg = np.hypot(np.diff(px), np.diff(py))
        s = np.concatenate([[0], np.cumsum(seg)])
        want = (np.arange(n_scans) * step) % s[-1]
        x, y = np.interp(want, s, px), np.interp(want, s, py)
        ahead = (want + 0.05) % s[-1]
        yaw = np.arctan2(np.interp(ahead, s, py) - y, np.interp(ahead, s, px) - x)
        poses = [synth.pose2d(x[i], y[i], yaw[i]) for i in range(n_scans)]
        clouds = [synth.laser2d_scan(scene, poses[i], seed * 13 + i, n_rays=n_rays, fov=np.deg2rad(fov_deg)) for i in range(n_scans)]
        rng = np.random.default_rng(seed + 1)
        Tm = [np.eye(4)]
        for i in range(1, n_scans):
            D = np.linalg.inv(poses[i - 1]) @ poses[i]
            d = np.hypot(D[0, 3], D[1, 3])
            e = synth.pose2d(rng.normal(0, 0.02 * d + 1e-4), rng.normal(0, 0.01 * d + 1e-4), rng.normal(0, 0.01 * d + 2e-4))
            Tm.append(e @ D)
        return clouds, poses, Tm

    return _cache(f"c5_{n_scans}_{seed}_{n_rays}", make)
