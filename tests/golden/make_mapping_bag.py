#!/usr/bin/env python3
"""Generate tests/golden/mapping_bag.npz from the reference's shipped ROS bag ndt_feature/data/mapping.bag — the bag
launch/henrik_replay_mapperbag_fuser.launch replays and the probable source of `FULL GRAPH/mapping{0..7}.jff`.

Runs ONLY in the authoring container (needs /root/reference); tests use the committed .npz.  Extracted: every
/laserscan message (stamp, 541 ranges as float32; angle_min / angle_increment / range_min / range_max) and every
/tf sample /world -> /odom_base_link (stamp, x, y, yaw): the pose source of publish_graph_message.cpp:1283-1300
(lookupTransform(world_frame, tf_odom_frame_, stamp), tf_odom_frame_ default "/odom_base_link").  Only data is extracted."""
import os
import struct
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_bag_scans as M  # noqa: E402

BAG = "/root/reference/ndt_feature/data/mapping.bag"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mapping_bag.npz")


def parse_tf(d):
    out = []
    (n,) = struct.unpack_from("<I", d, 0)
    o = 4
    for _ in range(n):
        seq, sec, nsec, fl = struct.unpack_from("<IIII", d, o)
        o += 16
        fr = d[o:o + fl].decode()
        o += fl
        (cl,) = struct.unpack_from("<I", d, o)
        o += 4
        ch = d[o:o + cl].decode()
        o += cl
        v = struct.unpack_from("<7d", d, o)
        o += 56
        yaw = np.arctan2(2 * (v[6] * v[5] + v[3] * v[4]), 1 - 2 * (v[4] ** 2 + v[5] ** 2))
        out.append((sec + 1e-9 * nsec, fr, ch, v[0], v[1], yaw))
    return out


def main():
    raw = open(BAG, "rb").read()
    assert raw[:13] == b"#ROSBAG V2.0\n"
    conns, msgs = {}, []
    M.walk(raw[13:], conns, msgs)
    scans = [M.parse_scan(d) for t, d in msgs if t == "/laserscan"]
    tfs = [x for t, d in msgs if t == "/tf" for x in parse_tf(d)]
    odom = np.array([[x[0], x[3], x[4], x[5]] for x in tfs if x[2] == "/odom_base_link"])
    t0 = np.floor(min(odom[0, 0], scans[0][0]))
    ranges = np.stack([s[5] for s in scans]).astype(np.float32)
    np.savez_compressed(OUT, ranges=ranges, stamp=np.array([s[0] for s in scans]) - t0, t0=np.float64(t0),
                        angle_min=np.float64(scans[0][1]), angle_inc=np.float64(scans[0][2]),
                        range_min=np.float64(scans[0][3]), range_max=np.float64(scans[0][4]),
                        odom_stamp=odom[:, 0] - t0, odom=odom[:, 1:4])
    print("wrote", OUT, ranges.shape, odom.shape, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
