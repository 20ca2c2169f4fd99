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


def _ros_pose(T):
    """geometry_msgs/Pose bytes: the library's own (its quaternion follows Eigen's algorithm and differs from scipy's in the
    last bit; the edge-message test above checks it against scipy to 1e-15), cut out of an NDTEdgeMsg"""
    return api.edge_msg_pack(0, 0, T, np.eye(3), None, 0.0)[8:64]


def _ros_matrix(m):
    rows, cols = m.shape
    return (struct.pack("<I", 2) + struct.pack("<III", 0, rows, rows * cols) + struct.pack("<III", 0, cols, cols) + struct.pack("<II", 0, rows * cols)
            + m.astype("<f8").tobytes())


def _ros_map(grid, cells, frame, stamp):
    """independent restatement of ndt_map/NDTMapMsg [upstream]: Header, sizes in metres, centre, cell sizes, NDTCellMsg[]"""
    b = struct.pack("<III", *stamp) + struct.pack("<I", len(frame)) + frame.encode()
    b += struct.pack("<3d", *[grid.size[a] * grid.cell[a] for a in range(3)]) + struct.pack("<3d", *grid.center) + struct.pack("<3d", *grid.cell)
    g = [c for c in cells if c["has_gaussian"]]
    b += struct.pack("<I", len(g))
    for c in g:
        xx, xy, xz, yy, yz, zz = c["cov"]
        b += struct.pack("<4d", *c["mean"], float(c["occ"])) + struct.pack("<I", 9) + struct.pack("<9d", xx, xy, xz, xy, yy, yz, xz, yz, zz)
        b += struct.pack("<d", float(c["n"]))
    return b


def _some_cells(rng, n):
    cells = np.zeros(n, api.CELL_DTYPE)
    cells["mean"] = rng.uniform(-20, 20, (n, 3))
    a = rng.normal(size=(n, 3, 3))
    cov = a @ a.transpose(0, 2, 1) * 0.01
    cells["cov"] = np.stack([cov[:, 0, 0], cov[:, 0, 1], cov[:, 0, 2], cov[:, 1, 1], cov[:, 1, 2], cov[:, 2, 2]], 1)
    cells["n"] = rng.integers(3, 500, n)
    cells["has_gaussian"] = rng.random(n) < 0.7
    cells["occ"] = rng.uniform(-5, 50, n).astype(np.float32)
    return cells


def test_graph_message_bytes_and_round_trip():
    """NDTGraphMsg = Header + sensor pose + Tnow + distance + NDTNodeMsg[] + NDTEdgeMsg[] (ndtgraph_conversion.h:36-83): the bytes
    equal an independent struct.pack restatement of the .msg definitions, and unpacking returns every field."""
    from ndt_feature_graph_b200 import synth

    rng = np.random.default_rng(5)
    grid = api.Grid()
    grid.center[:], grid.cell[:], grid.size[:] = [1.0, -2.0, 0.0], [0.5, 0.5, 0.5], [200, 200, 2]
    nodes, node_ref = [], []
    for k in range(3):
        cells = _some_cells(rng, 40 + 7 * k)
        mm = api.map_msg_pack(grid, cells, "/world", (k, 100 + k, 5))
        assert mm == _ros_map(grid, cells, "/world", (k, 100 + k, 5))
        g2, c2, frame, stamp, used = api.map_msg_unpack(mm)
        assert used == len(mm) and frame == "/world" and stamp == (k, 100 + k, 5)
        assert list(g2.size) == [200, 200, 2] and list(g2.center) == list(grid.center) and list(g2.cell) == list(grid.cell)
        gc = cells[cells["has_gaussian"] != 0]
        assert len(c2) == len(gc) and np.array_equal(c2["mean"], gc["mean"]) and np.array_equal(c2["cov"], gc["cov"])
        assert np.array_equal(c2["n"], gc["n"]) and np.array_equal(c2["occ"], gc["occ"])
        P = {name: synth.pose_from_xyzrpy(*rng.uniform(-1, 1, 3), *rng.uniform(-0.5, 0.5, 3)) for name in api.NodeFields.POSES}
        cov = rng.normal(size=(3, 3))
        f = api.NodeFields.make(cov=cov, ctr=10 + k, nb_updates=3 * k, time_last_update=12.5 + k, **P)
        nm = api.node_msg_pack(f, mm)
        ref = (_ros_pose(P["Tnow"]) + _ros_pose(P["Tlast_fuse"]) + _ros_pose(P["Todom"]) + mm + struct.pack("<I", 10 + k) + _ros_pose(P["T"])
               + _ros_matrix(cov) + _ros_pose(P["Tlocal_odom"]) + _ros_pose(P["Tlocal_fuse"]) + struct.pack("<Id", 3 * k, 12.5 + k))
        assert nm == ref
        f2, mm2, used = api.node_msg_unpack(nm)
        assert used == len(nm) and mm2 == mm and f2.ctr == 10 + k and f2.nb_updates == 3 * k and f2.time_last_update == 12.5 + k
        for name in api.NodeFields.POSES:
            assert np.allclose(f2.pose(name), P[name], atol=1e-12)
        assert np.array_equal(np.array(f2.cov9[:]).reshape(3, 3), cov)
        nodes.append(nm)
    edges = []
    for k in range(2):
        T = synth.pose_from_xyzrpy(*rng.uniform(-1, 1, 3), *rng.uniform(-0.5, 0.5, 3))
        edges.append(api.edge_msg_pack(k, k + 1, T, rng.normal(size=(3, 3)), rng.normal(size=(6, 6)), -100.0 - k))
    S, Tn = synth.pose_from_xyzrpy(0.695, -0.01, 0, 0, 0, -0.0069813), synth.pose_from_xyzrpy(3, 4, 0, 0, 0, 1.0)
    gm = api.graph_msg_pack(S, Tn, 0.73, nodes, edges, "/world", (7, 8, 9))
    ref = (struct.pack("<IIII", 7, 8, 9, 6) + b"/world" + _ros_pose(S) + _ros_pose(Tn) + struct.pack("<d", 0.73) + struct.pack("<I", 3)
           + b"".join(nodes) + struct.pack("<I", 2) + b"".join(edges))
    assert gm == ref
    out = api.graph_msg_unpack(gm)
    assert out["nodes"] == nodes and out["edges"] == edges and out["frame_id"] == "/world" and out["stamp"] == (7, 8, 9)
    assert out["distance_moved"] == 0.73 and np.allclose(out["sensor_pose"], S, atol=1e-12) and np.allclose(out["Tnow"], Tn, atol=1e-12)
    # truncated or corrupt input is rejected, never read past the end
    for cut in (0, 10, len(gm) // 2, len(gm) - 1):
        try:
            api.graph_msg_unpack(gm[:cut])
            assert False, cut
        except api.NdtbError:
            pass
    # an empty graph
    e = api.graph_msg_unpack(api.graph_msg_pack(S, Tn, 0.0, [], []))
    assert e["nodes"] == [] and e["edges"] == []
