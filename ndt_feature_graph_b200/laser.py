"""Host-side laser front end of the replay drivers: LaserScan -> pcl::PointXYZ cloud, tf pose at a stamp, keyframe gate.

Mirrors what the reference's drivers do before they call `graph->update` (no registration arithmetic here):

  laser_geometry::LaserProjection::projectLaser   publish_graph_message.cpp:1345 (x = r cos a, y = r sin a in double,
                                                   stored as float; ranges outside [range_min, range_max) dropped)
  z jitter + min range filter                      publish_graph_message.cpp:1371-1381 (z += varz * rand()/INT_MAX, points
                                                   closer than min_laser_range dropped)
  tf lookup at the scan stamp                      publish_graph_message.cpp:1283-1300 (linear interpolation between the
                                                   two bracketing /tf samples, as tf::Transformer does)
  incremental motion gate                          publish_graph_message.cpp:1325-1336 (min_incr_dist / min_incr_rot) and
                                                   ndt_graph_offline.cpp:586-588 (min_dist 0.2 m / min_rot 5 deg)
"""
import math

import numpy as np


class GlibcRand:
    """glibc rand() (TYPE_3 additive feedback generator, default seed 1): the stream publish_graph_message.cpp:1377 draws
    its z jitter from — `pt.z += varz*((double)rand())/(double)INT_MAX` — so a replay can reproduce the reference's
    clouds bit for bit.  skip = number of values already consumed."""

    def __init__(self, seed=1, skip=0):
        r = [0] * 34
        r[0] = seed
        for i in range(1, 31):
            r[i] = (16807 * r[i - 1]) % 2147483647
        for i in range(31, 34):
            r[i] = r[i - 31]
        self.r = r
        self.i = 34
        for _ in range(310):
            self._next()
        if skip:
            self.take(skip)

    def _next(self):
        r = self.r
        v = (r[self.i - 31] + r[self.i - 3]) & 0xFFFFFFFF
        r.append(v)
        self.i += 1
        if len(r) > 4096:  # keep the last 34 values only
            del r[: len(r) - 34]
            self.i = 34
        return v >> 1

    def take(self, n):
        return np.array([self._next() for _ in range(int(n))], dtype=np.int64)

    def jitter_z(self, n, varz):
        return (varz * self.take(n).astype(np.float64) / 2147483647.0).astype(np.float32)


def scan_to_cloud(ranges, angle_min, angle_inc, range_min, range_max, min_laser_range=0.5, varz=0.02, rng=None):
    """One LaserScan -> n x 4 float32 cloud in the laser frame (w = 0)."""
    r = np.asarray(ranges, dtype=np.float64)
    ang = float(angle_min) + np.arange(r.shape[0], dtype=np.float64) * float(angle_inc)
    ok = (r >= float(range_min)) & (r < float(range_max))
    x = (r * np.cos(ang)).astype(np.float32)
    y = (r * np.sin(ang)).astype(np.float32)
    keep = ok & (np.sqrt(x.astype(np.float32) * x + y * y) > np.float32(min_laser_range))
    n = int(keep.sum())
    out = np.zeros((n, 4), np.float32)
    out[:, 0], out[:, 1] = x[keep], y[keep]
    if varz > 0:
        if isinstance(rng, GlibcRand):
            out[:, 2] = rng.jitter_z(n, varz)
        else:
            u = (rng or np.random.default_rng(0)).random(n)
            out[:, 2] = (varz * u).astype(np.float32)
    return out


def pose2d(x, y, yaw):
    c, s = math.cos(yaw), math.sin(yaw)
    T = np.eye(4)
    T[0, 0], T[0, 1], T[1, 0], T[1, 1] = c, -s, s, c
    T[0, 3], T[1, 3] = x, y
    return T


def yaw_of(T):
    return math.atan2(T[1, 0], T[0, 0])


class TfTrack:
    """/world -> /odom_base_link samples; lookup(stamp) interpolates like tf (lerp translation, slerp yaw)."""

    def __init__(self, stamps, xyyaw):
        self.t = np.asarray(stamps, dtype=np.float64)
        self.p = np.asarray(xyyaw, dtype=np.float64)

    def lookup(self, stamp):
        i = int(np.searchsorted(self.t, stamp))
        if i <= 0:
            return pose2d(*self.p[0])
        if i >= len(self.t):
            return pose2d(*self.p[-1])
        t0, t1 = self.t[i - 1], self.t[i]
        a = 0.0 if t1 == t0 else (stamp - t0) / (t1 - t0)
        p0, p1 = self.p[i - 1], self.p[i]
        dyaw = (p1[2] - p0[2] + math.pi) % (2 * math.pi) - math.pi
        return pose2d(p0[0] + a * (p1[0] - p0[0]), p0[1] + a * (p1[1] - p0[1]), p0[2] + a * dyaw)


def rot_norm_xyz(T):
    """|eulerAngles(0,1,2)| of a planar pose = |yaw| (the gate of publish_graph_message.cpp:1328)."""
    return abs(yaw_of(T))


def keyframes(track, stamps, min_dist, min_rot, start=0):
    """Indices of the scans a driver hands to graph->update, with the motion since the previous processed scan.
    Gate: skip a scan while translation < min_dist AND rotation < min_rot (publish_graph_message.cpp:1328)."""
    out = []
    last = None
    for i in range(start, len(stamps)):
        P = track.lookup(stamps[i])
        if last is None:
            out.append((i, P))
            last = P
            continue
        Tm = np.linalg.inv(last) @ P
        if np.linalg.norm(Tm[:3, 3]) < min_dist and rot_norm_xyz(Tm) < min_rot:
            continue
        out.append((i, Tm))
        last = P
    return out
