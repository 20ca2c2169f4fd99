"""CPU suite: the result hand-off formats (csrc/formats.cpp) — NDTEdgeMsg wire bytes, *.T pose archives, eval lines."""
import json
import os
import struct

import numpy as np

from ndt_feature_graph_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_pose_archives_byte_exact_against_the_shipped_files(tmp_path):
    """ndt_feature/data/FULL GRAPH/mapping{k}.T, ...local_odom.T, ...local_fuse.T (tests/golden/pose_archives.json holds
    their text): read -> write reproduces every file byte for byte."""
    arch = json.load(open(os.path.join(ROOT, "tests", "golden", "pose_archives.json")))
    assert len(arch) == 24
    for name, text in arch.items():
        src = tmp_path / ("in_" + name)
        src.write_text(text)
        T = api.pose_archive_read(src)
        tok = text.split()
        assert np.array_equal(T, np.array(tok[-16:], float).reshape(4, 4).T)
        out = tmp_path / ("out_" + name)
        api.pose_archive_write(out, T)
        assert out.read_text() == text, name


def _ros_edge(ref, mov, T, c3, c6, score):
    """independent restatement of the ROS1 serialisation of ndt_feature/NDTEdgeMsg (msg/NDTEdgeMsg.msg)"""
    from scipy.spatial.transform import Rotation

    q = Rotation.from_matrix(T[:3, :3]).as_quat()  # x y z w
    if q[3] < 0:
        q = -q
    b = struct.pack("<II", ref, mov) + struct.pack("<7d", *T[:3, 3], *q)
    for m in (c3, c6):
        rows, cols = (0, 0) if m is None else m.shape
        b += struct.pack("<I", 2) + struct.pack("<III", 0, rows, rows * cols) + struct.pack("<III", 0, cols, cols) + struct.pack("<I", 0)
        b += struct.pack("<I", rows * cols) + (m.astype("<f8").tobytes() if m is not None else b"")
    return b + struct.pack("<d", score)


def test_edge_msg_wire_bytes_and_round_trip(golden):
    rng = np.random.default_rng(0)
    for k in range(7):
        T = golden[f"Tfuse{k}"]
        c3 = rng.normal(size=(3, 3))
        c6 = rng.normal(size=(6, 6)) if k % 2 == 0 else None
        msg = api.edge_msg_pack(k, k + 1, T, c3, c6, 0.25 * k)
        ref = _ros_edge(k, k + 1, T, c3, c6, 0.25 * k)
        assert len(msg) == len(ref)
        # header / matrices / score byte-exact; the quaternion may differ in the last bits (scipy vs Eigen's formula)
        assert msg[:32] == ref[:32] and msg[88:] == ref[88:]
        assert np.allclose(np.frombuffer(msg[32:88], "<f8"), np.frombuffer(ref[32:88], "<f8"), atol=1e-15)
        a, b, T2, d3, d6, s = api.edge_msg_unpack(msg)
        assert (a, b, s) == (k, k + 1, 0.25 * k) and np.array_equal(d3, c3)
        assert (d6 is None) == (c6 is None) and (c6 is None or np.array_equal(d6, c6))
        assert np.abs(T2 - T).max() < 1e-14
    # msgToEdge refuses a 3x3 covariance of the wrong size
    import pytest

    with pytest.raises(api.NdtbError):
        api.edge_msg_unpack(msg[:-20])


def test_eval_strings(golden):
    T = golden["T1"]
    line = api.eval_string(T)
    v = [float(x) for x in line.split()]
    assert len(v) == 7 and line.endswith("\n")
    assert np.allclose(v[:3], T[:3, 3], rtol=1e-14) and abs(np.linalg.norm(v[3:]) - 1) < 1e-12
    cols = line.rstrip("\n").split(" ")
    # Eigen right-aligns the three translation coefficients to a common width
    w = max(len(f"{x:.15g}") for x in T[:3, 3])
    assert line.startswith(f"{T[0, 3]:.15g}".rjust(w) + " " + f"{T[1, 3]:.15g}".rjust(w) + " " + f"{T[2, 3]:.15g}".rjust(w) + " ")
    l2 = api.eval_string(T, planar=True)
    v2 = l2.split()
    assert v2[2] == "0." and abs(float(v2[0]) - T[0, 3]) < 1e-12 and float(v2[3]) == 0 and float(v2[4]) == 0
    yaw = np.arccos(np.clip(T[0, 0], -1, 1)) * (1 if T[1, 0] > 0 else -1)  # getRobustYawFromAffine3d (utils.h:30-40)
    assert abs(2 * np.arctan2(float(v2[5]), float(v2[6])) - yaw) < 1e-9 or abs(abs(2 * np.arctan2(float(v2[5]), float(v2[6])) - yaw) - 2 * np.pi) < 1e-9
    assert cols
