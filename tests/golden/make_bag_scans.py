#!/usr/bin/env python3
"""Generate tests/golden/kyl1_scans.npz from the reference's shipped ROS bag ndt_feature/data/Kyl1.bag.

Runs ONLY in the authoring container (needs /root/reference); tests use the committed .npz.  A minimal ROSBAG v2.0
walker (uncompressed chunks): records are <u32 header_len><header fields "name=value" each with u32 len><u32 data_len>
<data>; op=0x05 chunk (recurse into its data), op=0x07 connection (conn id -> topic), op=0x02 message.  Extracted:
every /laserscan message (sensor_msgs/LaserScan: stamp, angle_min, angle_increment, range_max, ranges[361]) and every
/odom message (nav_msgs/Odometry: stamp, x, y, yaw), then keyframes are selected the way the offline driver does
(ndt_offline_ndt_feature/src/ndt_graph_offline.cpp:588: moved > 0.2 m or > 5 degrees).  Only data is extracted."""
import os
import struct
import sys

import numpy as np

BAG = "/root/reference/ndt_feature/data/Kyl1.bag"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "kyl1_scans.npz")


def fields(h):
    out, i = {}, 0
    while i < len(h):
        (n,) = struct.unpack_from("<I", h, i)
        kv = h[i + 4:i + 4 + n]
        k, v = kv.split(b"=", 1)
        out[k.decode()] = v
        i += 4 + n
    return out


def walk(buf, conns, msgs):
    i = 0
    while i + 8 <= len(buf):
        (hl,) = struct.unpack_from("<I", buf, i)
        h = fields(buf[i + 4:i + 4 + hl])
        (dl,) = struct.unpack_from("<I", buf, i + 4 + hl)
        data = buf[i + 8 + hl:i + 8 + hl + dl]
        i += 8 + hl + dl
        op = h.get("op", b"\xff")[0]
        if op == 0x05:
            assert h.get("compression", b"none") == b"none"
            walk(data, conns, msgs)
        elif op == 0x07:
            conns[struct.unpack("<I", h["conn"])[0]] = h["topic"].decode()
        elif op == 0x02:
            msgs.append((conns.get(struct.unpack("<I", h["conn"])[0], "?"), data))


def parse_scan(d):
    seq, sec, nsec, fl = struct.unpack_from("<IIII", d, 0)
    o = 16 + fl
    amin, amax, ainc, tinc, stime, rmin, rmax = struct.unpack_from("<7f", d, o)
    o += 28
    (n,) = struct.unpack_from("<I", d, o)
    r = np.frombuffer(d, "<f4", n, o + 4).copy()
    return sec + 1e-9 * nsec, amin, ainc, rmin, rmax, r


def parse_odom(d):
    seq, sec, nsec, fl = struct.unpack_from("<IIII", d, 0)
    o = 16 + fl
    (cl,) = struct.unpack_from("<I", d, o)
    o += 4 + cl
    px, py, pz, qx, qy, qz, qw = struct.unpack_from("<7d", d, o)
    yaw = np.arctan2(2 * (qw * qz + qx * qy), 1 - 2 * (qy * qy + qz * qz))
    return sec + 1e-9 * nsec, px, py, yaw


def main():
    raw = open(BAG, "rb").read()
    assert raw[:13] == b"#ROSBAG V2.0\n"
    conns, msgs = {}, []
    walk(raw[13:], conns, msgs)
    scans = [parse_scan(d) for t, d in msgs if t == "/laserscan"]
    odom = np.array([parse_odom(d) for t, d in msgs if t == "/odom"])
    print(len(scans), "scans", len(odom), "odom", sorted(set(t for t, _ in msgs)))
    st = np.array([s[0] for s in scans])
    # odometry pose at every scan stamp (nearest earlier sample)
    idx = np.clip(np.searchsorted(odom[:, 0], st) - 1, 0, len(odom) - 1)
    pose = odom[idx, 1:4]
    keep, last = [0], pose[0]
    for i in range(1, len(scans)):
        d = pose[i] - last
        dyaw = (d[2] + np.pi) % (2 * np.pi) - np.pi
        if np.hypot(d[0], d[1]) > 0.2 or abs(dyaw) > np.deg2rad(5):
            keep.append(i)
            last = pose[i]
    keep = keep[:160]
    ranges = np.stack([scans[i][5] for i in keep]).astype(np.float32)
    np.savez_compressed(OUT, ranges=ranges, angle_min=np.float64(scans[0][1]), angle_inc=np.float64(scans[0][2]),
                        range_min=np.float64(scans[0][3]), range_max=np.float64(scans[0][4]), odom=pose[keep].astype(np.float64),
                        stamp=st[keep])
    print("wrote", OUT, ranges.shape, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
