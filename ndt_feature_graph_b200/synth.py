"""Synthetic scans for the BASELINE.json configs.

Seeded numpy generators (PCG64); no reference code involved.  Shapes:
  C1  2-D room, 10 000 rays over 360 deg, range <= 30 m, sigma 0.02, z = U[0,1)*0.02
      (z jitter mirrors publish_graph_message.cpp:1373-1381, laser_variance_z=0.02)
  C2  Velodyne-like: 64 rings (-24.8..+2 deg) x 1563 azimuth steps = 100 032 rays into
      ground plane + boxes + walls within 70 m, sigma 0.02
All clouds are float32 [n,4] (x,y,z,0) in the SENSOR frame, i.e. pcl::PointXYZ layout.
"""
import numpy as np


# ------------------------------------------------------------------ poses
def pose_from_xyzrpy(x, y, z, rx, ry, rz):
    """Trans(x,y,z) * Rx(rx) * Ry(ry) * Rz(rz) -- the reference's increment convention
    (ndt_matcher_d2d_fusion.h:1036-1039)."""
    cx, sx, cy, sy, cz, sz = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry), np.cos(rz), np.sin(rz)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    T = np.eye(4)
    T[:3, :3] = Rx @ Ry @ Rz
    T[:3, 3] = (x, y, z)
    return T


def pose2d(x, y, yaw):
    return pose_from_xyzrpy(x, y, 0.0, 0.0, 0.0, yaw)


def se3_log(T):
    """6-vector (translation, rotation-vector) of a 4x4 pose; the SE(3) parity metric of SURVEY.md §8d."""
    R = T[:3, :3]
    c = np.clip((np.trace(R) - 1.0) / 2.0, -1.0, 1.0)
    ang = np.arccos(c)
    w = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]]) / 2.0
    if ang < 1e-9:
        rv = w
    else:
        rv = w * (ang / np.sin(ang))
    return np.concatenate([T[:3, 3], rv])


def pose_error(Ta, Tb):
    """|| log(Ta^-1 Tb) || (translation and rotation-vector stacked)."""
    return float(np.linalg.norm(se3_log(np.linalg.inv(Ta) @ Tb)))


def robust_yaw(T):
    """ndt_feature::getRobustYawFromAffine3d (utils.h:30-40)."""
    v = T[:3, :3] @ np.array([1.0, 0, 0])
    ang = np.arccos(np.clip(v[0], -1, 1))
    return ang if v[1] > 0 else -ang


def pose_error_2d(Ta, Tb):
    D = np.linalg.inv(Ta) @ Tb
    return float(np.linalg.norm([D[0, 3], D[1, 3], robust_yaw(D)]))


# ------------------------------------------------------------------ ray casting against AABBs (+ ground)
def _cast(origin, dirs, boxes, ground_z=None, max_range=70.0):
    """origin [3], dirs [n,3] unit, boxes [m,2,3] -> hit distance [n] (inf = miss)."""
    n = dirs.shape[0]
    t_hit = np.full(n, np.inf)
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / dirs
        if boxes is not None and len(boxes):
            lo = (boxes[None, :, 0, :] - origin[None, None, :]) * inv[:, None, :]
            hi = (boxes[None, :, 1, :] - origin[None, None, :]) * inv[:, None, :]
            tmin = np.nanmax(np.minimum(lo, hi), axis=2)
            tmax = np.nanmin(np.maximum(lo, hi), axis=2)
            ok = (tmax >= np.maximum(tmin, 0.0)) & (tmin > 1e-6)
            t = np.where(ok, tmin, np.inf)
            t_hit = np.minimum(t_hit, t.min(axis=1))
        if ground_z is not None:
            tg = (ground_z - origin[2]) * inv[:, 2]
            tg = np.where((dirs[:, 2] < -1e-9) & (tg > 1e-6), tg, np.inf)
            t_hit = np.minimum(t_hit, tg)
    t_hit[t_hit > max_range] = np.inf
    return t_hit


def _scan(scene, T_ws, dirs_s, rng, noise, max_range, chunk=20000):
    R, o = T_ws[:3, :3], T_ws[:3, 3]
    out = []
    for s in range(0, dirs_s.shape[0], chunk):
        d_s = dirs_s[s : s + chunk]
        d_w = d_s @ R.T
        t = _cast(o, d_w, scene["boxes"], scene.get("ground_z"), max_range)
        ok = np.isfinite(t)
        r = t[ok] + rng.normal(0.0, noise, ok.sum())
        out.append(d_s[ok] * r[:, None])
    return np.concatenate(out, axis=0)


def _to_cloud(p):
    c = np.zeros((p.shape[0], 4), np.float32)
    c[:, :3] = p.astype(np.float32)
    return c


# ------------------------------------------------------------------ C2: Velodyne-like 3-D scans
def velodyne_scene(seed, extent=60.0):
    rng = np.random.default_rng(seed)
    boxes = []
    nb = int(rng.integers(20, 41))
    for _ in range(nb):
        c = rng.uniform(-extent * 0.8, extent * 0.8, 2)
        if np.linalg.norm(c) < 4.0:
            c += 6.0 * c / max(np.linalg.norm(c), 1e-3) + 1.0
        sz = rng.uniform(1.0, 10.0, 3)
        sz[2] = rng.uniform(1.0, 8.0)
        boxes.append([[c[0] - sz[0] / 2, c[1] - sz[1] / 2, 0.0], [c[0] + sz[0] / 2, c[1] + sz[1] / 2, sz[2]]])
    w, h, th = extent, 6.0, 0.5
    boxes += [
        [[-w, w, 0], [w, w + th, h]],
        [[-w, -w - th, 0], [w, -w, h]],
        [[w, -w, 0], [w + th, w, h]],
        [[-w - th, -w, 0], [-w, w, h]],
    ]
    return {"boxes": np.array(boxes, float), "ground_z": 0.0}


def velodyne_dirs(n_rings=64, n_az=1563, elev_lo=-24.8, elev_hi=2.0):
    el = np.deg2rad(np.linspace(elev_lo, elev_hi, n_rings))
    az = np.linspace(0.0, 2 * np.pi, n_az, endpoint=False)
    # azimuth-major (all rings fire per azimuth step), like a spinning head
    A, E = np.meshgrid(az, el, indexing="ij")
    d = np.stack([np.cos(E) * np.cos(A), np.cos(E) * np.sin(A), np.sin(E)], axis=-1).reshape(-1, 3)
    return d


def velodyne_scan(scene, T_ws, seed, n_rings=64, n_az=1563, noise=0.02, max_range=70.0):
    rng = np.random.default_rng(seed)
    return _to_cloud(_scan(scene, T_ws, velodyne_dirs(n_rings, n_az), rng, noise, max_range))


def velodyne_pair(i, n_rings=64, n_az=1563, sensor_height=1.8):
    """Scan pair i of config C2.  Returns (cloud_target, cloud_source, T_true) where T_true maps the
    source (moving) scan into the target (fixed) frame."""
    seed = 2000 + i
    rng = np.random.default_rng(seed)
    scene = velodyne_scene(seed)
    Ta = pose_from_xyzrpy(rng.uniform(-3, 3), rng.uniform(-3, 3), sensor_height, 0, 0, rng.uniform(-np.pi, np.pi))
    d = np.concatenate([rng.uniform(-1, 1, 2), rng.uniform(-0.1, 0.1, 1), rng.uniform(-0.02, 0.02, 2), rng.uniform(-0.1, 0.1, 1)])
    D = pose_from_xyzrpy(*d)
    Tb = Ta @ D
    ca = velodyne_scan(scene, Ta, seed * 7 + 1, n_rings, n_az)
    cb = velodyne_scan(scene, Tb, seed * 7 + 2, n_rings, n_az)
    return ca, cb, D


# ------------------------------------------------------------------ C1: 2-D laser in a room
def room2d_scene(seed, half=20.0):
    rng = np.random.default_rng(seed)
    boxes = []
    th, h = 0.3, 2.0
    # outer walls with a few jogs (8-12 segments)
    boxes += [
        [[-half, half, -1], [half, half + th, h]],
        [[-half, -half - th, -1], [half, -half, h]],
        [[half, -half, -1], [half + th, half, h]],
        [[-half - th, -half, -1], [-half, half, h]],
    ]
    for _ in range(int(rng.integers(4, 9))):
        c = rng.uniform(-half * 0.9, half * 0.9, 2)
        L = rng.uniform(2.0, 10.0)
        if rng.random() < 0.5:
            boxes.append([[c[0] - L / 2, c[1] - th / 2, -1], [c[0] + L / 2, c[1] + th / 2, h]])
        else:
            boxes.append([[c[0] - th / 2, c[1] - L / 2, -1], [c[0] + th / 2, c[1] + L / 2, h]])
    for _ in range(6):
        c = rng.uniform(-half * 0.8, half * 0.8, 2)
        if np.linalg.norm(c) < 2.5:
            c += 4.0
        s = rng.uniform(0.5, 3.0, 2)
        boxes.append([[c[0] - s[0] / 2, c[1] - s[1] / 2, -1], [c[0] + s[0] / 2, c[1] + s[1] / 2, h]])
    return {"boxes": np.array(boxes, float)}


def laser2d_scan(scene, T_ws, seed, n_rays=10000, fov=2 * np.pi, noise=0.02, max_range=30.0, z_jitter=0.02):
    rng = np.random.default_rng(seed)
    a = np.linspace(-fov / 2, fov / 2, n_rays, endpoint=False)
    d = np.stack([np.cos(a), np.sin(a), np.zeros_like(a)], axis=-1)
    p = _scan(scene, T_ws, d, rng, noise, max_range)
    p[:, 2] = rng.random(p.shape[0]) * z_jitter
    return _to_cloud(p)


def laser2d_pair(i, n_rays=10000):
    """Scan pair i of config C1."""
    seed = 1000 + i
    rng = np.random.default_rng(seed)
    scene = room2d_scene(seed)
    Ta = pose2d(rng.uniform(-2, 2), rng.uniform(-2, 2), rng.uniform(-np.pi, np.pi))
    D = pose2d(rng.uniform(-0.5, 0.5), rng.uniform(-0.5, 0.5), rng.uniform(-0.1, 0.1))
    Tb = Ta @ D
    ca = laser2d_scan(scene, Ta, seed * 7 + 1, n_rays)
    cb = laser2d_scan(scene, Tb, seed * 7 + 2, n_rays)
    return ca, cb, D


def perturb_pose(T, seed, dt=0.15, dr=0.03, planar=False):
    """Odometry-like initial guess: T_true composed with a small seeded offset."""
    rng = np.random.default_rng(seed)
    if planar:
        e = pose2d(rng.uniform(-dt, dt), rng.uniform(-dt, dt), rng.uniform(-dr, dr))
    else:
        e = pose_from_xyzrpy(*rng.uniform(-dt, dt, 3), *rng.uniform(-dr, dr, 3))
    return e @ T


# ------------------------------------------------------------------ batches of C2-shaped pairs (bench.py, tests)
def odometry_guess(T, seed, dxy=0.15, dz=0.02, drp=0.005, dyaw=0.03):
    """Vehicle-odometry-like initial guess: planar error dominates (x, y, yaw), small z / roll / pitch error."""
    rng = np.random.default_rng(seed)
    e = pose_from_xyzrpy(rng.uniform(-dxy, dxy), rng.uniform(-dxy, dxy), rng.uniform(-dz, dz), rng.uniform(-drp, drp),
                         rng.uniform(-drp, drp), rng.uniform(-dyaw, dyaw))
    return e @ T


def _apply(G, cloud):
    out = np.zeros_like(cloud)
    out[:, :3] = (cloud[:, :3].astype(np.float64) @ G[:3, :3].T + G[:3, 3]).astype(np.float32)
    return out


def velodyne_batch(n_pairs, n_base=8, seed=0, n_rings=64, n_az=1563, start=0, indices=None):
    """Pairs [start, start + n_pairs) — or the pairs `indices` — of the C2 workload `seed`.  n_base scenes are ray-cast (slow, numpy); pair i is base
    pair i % n_base moved by its own rigid transform G_i (both scans), which changes the voxelisation of both maps, so
    all pairs are distinct data.  G_i and the initial guess depend only on (seed, i): ranks of a multi-GPU run take
    disjoint index ranges of ONE workload.  Returns lists (target clouds, source clouds, initial guesses T0, true D)."""
    idx = list(range(start, start + n_pairs)) if indices is None else [int(i) for i in indices]
    base = [velodyne_pair(100 * seed + b, n_rings, n_az) for b in range(min(n_base, (max(idx) + 1) if idx else 0))]
    tg, sr, T0s, Ds = [], [], [], []
    for i in idx:
        ca, cb, D = base[i % len(base)]
        if i < len(base):
            G = np.eye(4)
        else:
            rng = np.random.default_rng([9000 + seed, i])
            G = pose_from_xyzrpy(rng.uniform(-2, 2), rng.uniform(-2, 2), rng.uniform(-0.2, 0.2), 0.0, 0.0, rng.uniform(-np.pi, np.pi))
        Gi = np.linalg.inv(G)
        Di = G @ D @ Gi
        tg.append(_apply(G, ca) if i >= len(base) else ca)
        sr.append(_apply(G, cb) if i >= len(base) else cb)
        Ds.append(Di)
        T0s.append(odometry_guess(Di, 7000 + 31 * seed + i))
    return tg, sr, T0s, Ds
