"""The C++ host façade (include/ndtb_lslgeneric.hpp): the reference's class names over the C ABI.

CPU: the façade test program compiles against the header + libndtb.so and fails loudly without a device.
GPU: tests/harness/facade_corridor.cpp (the shape of ndt_feature/src/ndt_odom_debug.cpp:94-256 plus the graph entry
point of ndt_feature_graph.cpp:347-353) runs on cuda:0 and every number it prints is compared with the oracle run on
the clouds it dumped."""
import json
import os
import subprocess

import numpy as np
import pytest

from ndt_feature_graph_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
POSE_TOL = 1e-4  # north_star tolerance


@pytest.fixture(scope="module")
def facade_bin():
    import __graft_entry__ as g

    return g.build_facade_test()


def test_facade_compiles_and_refuses_without_device(facade_bin):
    out = subprocess.run([facade_bin, "--abi-check"], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "abi ok" in out.stdout


def _cm(v):
    return np.array(v, dtype=np.float64).reshape(4, 4).T


@pytest.mark.gpu
def test_facade_corridor_matches_oracle(facade_bin, oracle, tmp_path):
    out = subprocess.run([facade_bin, str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    r = json.loads(out.stdout)
    assert r["status"] == 0 and r["links_rc"] == 0
    clouds = [np.fromfile(tmp_path / f, dtype=np.float32).reshape(-1, 4) for f in ("static.bin", "moving.bin", "third.bin")]
    assert clouds[0].shape[0] == r["n_points"]
    om = []
    for c in clouds:
        m = oracle.OracleMap(0.5)
        m.set_map_size(80.0, 80.0, 2.0)
        m.load_point_cloud(c, -1.0)
        m.compute_cells()
        om.append(m)
    assert r["cells_static"] == om[0].num_cells(True) and r["cells_moving"] == om[1].num_cells(True)
    assert r["n_pseudo"] == r["cells_static"]
    odom, gt = _cm(r["odom"]), _cm(r["gt"])
    # pseudoTransformNDT: first Gaussian cell (linear voxel order) moved by gt
    c0 = om[0].export_cells(True)[0]
    np.testing.assert_allclose(r["pseudo_mean0"], gt[:3, :3] @ c0["mean"] + gt[:3, 3], rtol=0, atol=1e-12)
    # NDTMatcherD2D::match from the odometry guess
    ro = oracle.d2d_match(om[0], om[1], odom)
    Tg = _cm(r["T_d2d"])
    assert synth.pose_error(ro.pose(), Tg) < POSE_TOL
    assert (r["converged"], r["iterations"]) == (ro.converged, ro.iterations)
    assert synth.pose_error(Tg, np.linalg.inv(gt)) < 0.05  # and it registers: moving = gt * static, so T -> gt^-1
    # covariance and derivatives at the refined pose
    rc, co = oracle.d2d_covariance(om[0], om[1], Tg)
    assert rc == 0 and r["cov_ok"] == 1
    np.testing.assert_allclose(np.array(r["cov"]).reshape(6, 6), co, rtol=1e-6, atol=1e-9 * np.abs(co).max())
    so, go, Ho, _ = oracle.d2d_derivatives(om[0], om[1], Tg, want_hessian=True)
    assert abs(r["score"] - so) <= 1e-10 * abs(so)
    np.testing.assert_allclose(np.array(r["gradient"]), go, rtol=1e-8, atol=1e-9 * np.abs(go).max())
    np.testing.assert_allclose(np.array(r["hessian"]).reshape(6, 6), Ho, rtol=1e-8, atol=1e-9 * np.abs(Ho).max())
    # matchFusion with the soft constraint
    Tcov = np.diag([0.01, 0.01, 1e-6, 1e-6, 1e-6, 0.001])
    rf = oracle.fusion_match(om[0], om[1], odom, Tcov, oracle.default_params(delta_score=1e-6, use_soft_constraints=1))
    assert synth.pose_error(rf.pose(), _cm(r["T_fusion"])) < POSE_TOL and r["fusion_ok"] == rf.converged
    # NDTMatcherP2D: the moving cloud against the static map
    rp = oracle.p2d_match(om[0], clouds[1], odom)
    assert synth.pose_error(rp.pose(), _cm(r["T_p2d"])) < POSE_TOL
    assert (r["p2d_ok"], r["p2d_iterations"]) == (rp.converged, rp.iterations)
    # graph edges: updateLinksUsingNDTRegistration
    T0s = [odom, synth.pose2d(0.15, 0.0, 0.0), synth.pose2d(900.0, 900.0, 0.0)]
    pairs = [(0, 1), (0, 2), (1, 2)]
    for i, ((a, b), T0) in enumerate(zip(pairs, T0s)):
        rl = oracle.d2d_match(om[a], om[b], T0)
        Tl = _cm(r[f"link{i}_T"])
        assert synth.pose_error(rl.pose(), Tl) < POSE_TOL
        assert r[f"link{i}_converged"] == rl.converged
        cl = np.array(r[f"link{i}_cov"]).reshape(6, 6)
        if rl.pose_changed:
            _, col = oracle.d2d_covariance(om[a], om[b], rl.pose())
            np.testing.assert_allclose(cl, col, rtol=1e-6, atol=1e-9 * np.abs(col).max())
        else:
            assert np.array_equal(cl, 0.02 * np.eye(6))  # ndt_feature_graph.cpp:300-310
        so = oracle.overlap_occupancy_score(om[a], om[b], Tl)
        assert abs(r[f"link{i}_score"] - so) <= 1e-12 * max(1.0, abs(so))
    assert np.array_equal(_cm(r["link2_T"]), T0s[2])


def _write_saved_graph(golden, tmp_path):
    """What the fuser saves per node (ndt_feature_fuser_hmt.cpp:20-49): mapping{k}.jff + mapping{k}local_odom.T, rebuilt
    from the golden extraction of the shipped files with the library's own JFF writer."""
    from conftest import fixture_cells
    from ndt_feature_graph_b200 import api

    for k in range(8):
        center, cell, size, cells = fixture_cells(golden, k, api.CELL_DTYPE)
        api.jff_write_cells(tmp_path / f"mapping{k}.jff", center, cell, size, cells)
        if k < 7:
            T = golden[f"Todom{k}"]
            with open(tmp_path / f"mapping{k}local_odom.T", "w") as f:  # boost text archive: header tokens, then 16 numbers
                f.write("22 serialization::archive 10 0 0 " + " ".join(repr(float(v)) for v in T.T.ravel()) + "\n")
    return str(tmp_path / "mapping")


def test_example_compiles_and_fails_loudly_without_device(golden, tmp_path):
    import torch

    import __graft_entry__ as g

    exe = g.build_example()
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by test_example_refines_saved_graph")
    out = subprocess.run([exe, _write_saved_graph(golden, tmp_path), "8"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 1 and "no CUDA device" in out.stderr


@pytest.mark.gpu
def test_example_refines_saved_graph(golden, oracle, oracle_fixture_maps, tmp_path):
    """examples/refine_edges.cpp end to end: JFF node maps from disk -> facade -> one batched edge refinement."""
    import re

    import __graft_entry__ as g

    out = subprocess.run([g.build_example(), _write_saved_graph(golden, tmp_path), "8"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    links = re.findall(r"link (\d+) -> (\d+): x (\S+) y (\S+) yaw (\S+)  converged (\d)  score (\S+)", out.stdout)
    assert len(links) == 7
    for k, (a, b, x, y, yaw, conv, score) in enumerate(links):
        ro = oracle.d2d_match(oracle_fixture_maps[k], oracle_fixture_maps[k + 1], golden[f"Todom{k}"])
        T = ro.pose()
        assert (int(a), int(b), int(conv)) == (k, k + 1, ro.converged)
        assert abs(float(x) - T[0, 3]) < 2e-4 and abs(float(y) - T[1, 3]) < 2e-4  # printed with 4 decimals
        assert abs(float(yaw) - np.arctan2(T[1, 0], T[0, 0])) < 2e-4
        so = oracle.overlap_occupancy_score(oracle_fixture_maps[k], oracle_fixture_maps[k + 1], T)
        assert abs(float(score) - so) < 2e-4


def test_reference_call_sites_compile():
    """tests/harness/callsites.cpp holds the reference's own call statements of the hot path (fuser initialize / update,
    updateLinkUsingNDTRegistration, the optimiser's derivativesNDT / lineSearchMT / MoreThuente::cstep uses): they must
    compile against the facade with -Wall -Wextra clean."""
    import __graft_entry__ as g

    exe = g.build_callsites_test()
    assert os.path.exists(exe)


@pytest.mark.gpu
def test_reference_call_sites_run(oracle):
    import __graft_entry__ as g

    out = subprocess.run([g.build_callsites_test()], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    rows = {l.split()[0]: l.split()[1:] for l in out.stdout.strip().splitlines()}
    assert rows["matchFusion"][0] == "1"
    # the second corridor scan is the first one shifted by 0.15 m along -x: the registration must recover it
    assert abs(float(rows["matchFusion"][1]) - 0.15) < 0.02 and abs(float(rows["matchFusion"][2])) < 0.02
    assert float(rows["covariance"][0]) > 0 and float(rows["covariance"][1]) > 0
    assert int(rows["cells"][0]) > 20
    # updateLinkUsingNDTRegistration with a default-constructed matcher (DELTA_SCORE 1e-3): converged, pose changed, so the
    # covariance branch (:297-298) ran
    assert rows["link"][0] == "1" and rows["link"][1] == "0" and float(rows["link"][2]) > 0 and float(rows["link"][4]) > 0
    assert float(rows["blocks"][0]) < 0 and 0.0 < float(rows["blocks"][1]) <= 4.0 and int(rows["blocks"][2]) in (1, 2, 3, 4)
