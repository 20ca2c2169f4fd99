"""CPU suite, part 1: the oracle (oracle/ndt_oracle.cpp, a restatement of the reference's CPU algorithm) against the
pins the reference offers for this path: the shipped LazyGrid fixtures
(ndt_feature/data/FULL GRAPH/mapping{k}.jff -> tests/golden/full_graph.npz via tests/golden/make_golden.py), the real
node-pair envelope, and analytic checks.  "parity unpinned" for the perception_oru arithmetic itself."""
import numpy as np
import pytest

from ndt_feature_graph_b200 import synth


# ---- fixture invariants pin kernel (i)'s finalisation and the index rule -------------------------------
def test_fixture_archive_consistency(golden):
    for k in range(6):  # T_{k+1} = T_k * Tlocal_fuse_k (node 7 is the still-open last node: its pose was saved earlier)
        np.testing.assert_allclose(golden[f"T{k}"] @ golden[f"Tfuse{k}"], golden[f"T{k + 1}"], atol=1e-9)


def test_fixture_invariants(golden):
    ncap = 0
    for k in range(8):
        cov, n, mean, ctr = golden[f"cov{k}"], golden[f"n{k}"], golden[f"mean{k}"], golden[f"ctr{k}"]
        assert n.min() >= 3
        full = np.stack([cov[:, [0, 1, 2]], cov[:, [1, 3, 4]], cov[:, [2, 4, 5]]], axis=1)
        ev = np.linalg.eigvalsh(full)
        assert (ev > 0).all()
        ratio = ev[:, 2] / ev[:, 0]
        assert ratio.max() <= 1000.0 * (1 + 1e-9)  # upstream EVAL_FACTOR
        ncap += int((ratio > 999.999).sum())
        assert (np.abs(mean - ctr) <= 0.25 + 1e-6).all()  # mean inside its 0.5 m cell
        # occupancy: cells that were only ever hit (never ray-traced through) hold exactly N * ln(0.6/0.4)
        r = golden[f"occ{k}"] / np.log(1.5)
        hit_only = np.abs(r - golden[f"occn{k}"]) < 1e-3
        assert hit_only.mean() > 0.1 and (golden[f"occn{k}"][hit_only] >= 1).all()
    assert ncap > 10


def test_oracle_index_rule_matches_fixture_lattice(golden, oracle):
    hdr = golden["hdr0"]
    size = np.abs(np.ceil(hdr[:3] / hdr[3:6])).astype(int)
    m = oracle.OracleMap(0.5)
    m.initialize(*hdr[6:9], *hdr[:3])
    mean, gi = golden["mean0"], golden["gidx0"]
    idx = m.point_indices(mean.astype(np.float32))
    lin = (idx[:, 0].astype(np.int64) * size[1] + idx[:, 1]) * size[2] + idx[:, 2]
    assert np.array_equal(lin, gi)  # the voxel of every stored mean is the record it was stored in


def test_oracle_rescale_reproduces_fixture_clamp(oracle):
    """cells rebuilt from points obey the same invariants as the shipped maps"""
    rng = np.random.default_rng(0)
    pts = np.zeros((4000, 4), np.float32)
    pts[:, 0] = rng.uniform(-5, 5, 4000)
    pts[:, 1] = rng.normal(0, 0.002, 4000)  # a thin wall: eigenvalue ratio far above 1000 before the clamp
    pts[:, 2] = rng.uniform(0, 0.02, 4000)
    m = oracle.OracleMap(0.5)
    m.initialize(0, 0, 0, 20, 20, 2)
    assert m.add_points(pts) == 4000
    m.compute_cells()
    c = m.export_cells(True)
    assert c.shape[0] >= 18
    full = np.stack([c["cov"][:, [0, 1, 2]], c["cov"][:, [1, 3, 4]], c["cov"][:, [2, 4, 5]]], axis=1)
    ev = np.linalg.eigvalsh(full)
    np.testing.assert_allclose(ev[:, 2] / ev[:, 0], 1000.0, rtol=1e-6)
    alln = m.export_cells(False)
    np.testing.assert_allclose(alln["occ"], np.minimum(alln["n"] * np.log(1.5), 255.0), rtol=2e-6)


# ---- real node pairs: envelope of SURVEY.md §4 ----------------------------------------------------------
def test_oracle_real_pairs_envelope(golden, oracle, oracle_fixture_maps):
    p = oracle.default_params(delta_score=1e-6)
    for k in range(7):
        r = oracle.d2d_match(oracle_fixture_maps[k], oracle_fixture_maps[k + 1], golden[f"Todom{k}"], p)
        T, F = r.pose(), golden[f"Tfuse{k}"]
        assert r.converged
        assert np.hypot(T[0, 3] - F[0, 3], T[1, 3] - F[1, 3]) < 0.12
        assert abs(synth.robust_yaw(T) - synth.robust_yaw(F)) < 0.012


# ---- analytic: finite differences of the scalar score pin gradient and Hessian --------------------------
def test_oracle_derivatives_finite_differences(golden, oracle, oracle_fixture_maps):
    tgt, src = oracle_fixture_maps[2], oracle_fixture_maps[3]
    T = golden["Tfuse2"]
    s0, g, H, npairs = oracle.d2d_derivatives(tgt, src, T)
    assert npairs > 100

    def score(dp):
        return oracle.d2d_derivatives(tgt, src, oracle.pose_from_vec(dp) @ T, want_hessian=False)[0]

    def grad(dp):
        # the gradient is taken at p=0 of the pose it is evaluated at
        return oracle.d2d_derivatives(tgt, src, oracle.pose_from_vec(dp) @ T, want_hessian=False)[1]

    h = 1e-6
    for i in range(6):
        e = np.zeros(6)
        e[i] = h
        fd = (score(e) - score(-e)) / (2 * h)
        assert abs(fd - g[i]) <= 2e-5 * max(1.0, abs(g[i]))
    # Hessian columns (translations commute with everything at p=0: exact; rotation block to first order)
    for i in range(3):
        e = np.zeros(6)
        e[i] = h
        fdH = (grad(e) - grad(-e)) / (2 * h)
        np.testing.assert_allclose(fdH[:3], H[:3, i], rtol=5e-4, atol=1e-4 * np.abs(H).max())
    assert np.allclose(H, H.T)


def test_oracle_hessian_full_6x6_finite_differences(golden, oracle, oracle_fixture_maps):
    """All 36 entries of the D2D Hessian — rotation-rotation block with its Z / ZH terms included (the "coefficient 1 vs 2"
    trap of SURVEY.md §8a) — against second central differences of the scalar score under the reference's own
    parametrisation p -> score(Trans(p0..2) Rx(p3) Ry(p4) Rz(p5) * T) (ndt_matcher_d2d_fusion.h:1035-1043)."""
    for k in (2, 5):
        tgt, src = oracle_fixture_maps[k], oracle_fixture_maps[k + 1]
        T = golden[f"Tfuse{k}"]
        _, g, H, _ = oracle.d2d_derivatives(tgt, src, T)

        def score(dp):
            return oracle.d2d_derivatives(tgt, src, oracle.pose_from_vec(dp) @ T, want_hessian=False)[0]

        # steps far below the length scale of the score (millimetre-thin Gaussians in z: 2e-6 rad is 20 um at 10 m)
        hs = [2e-5, 2e-5, 2e-5, 2e-6, 2e-6, 2e-6]
        fd = np.zeros((6, 6))
        for i in range(6):
            for j in range(i, 6):
                ei, ej = np.zeros(6), np.zeros(6)
                ei[i], ej[j] = hs[i], hs[j]
                fd[i, j] = fd[j, i] = (score(ei + ej) - score(ei - ej) - score(ej - ei) + score(-ei - ej)) / (4 * hs[i] * hs[j])
        scale = np.abs(H).max()
        assert np.abs(fd - H).max() < 2e-4 * scale, (k, np.abs(fd - H).max() / scale)
        # the rotation block on its own (it is an order of magnitude smaller than the largest entry in these planar maps)
        rr = np.abs(H[3:, 3:]).max()
        assert np.abs(fd[3:, 3:] - H[3:, 3:]).max() < 2e-3 * rr, (k, np.abs(fd[3:, 3:] - H[3:, 3:]).max() / rr)


def test_incremental_cell_moves_agree_with_cumulative_pose(golden, oracle, oracle_fixture_maps):
    """The reference moves its copy of the source cells by every accepted increment (ndt_matcher_d2d_fusion.h:840,
    1047-1056); oracle and engine evaluate T_cumulative * cell_original instead (nothing is written back on the GPU).  The
    two data flows differ by rounding only: same iteration / pass counts and poses within 1e-9 on the real node pairs."""
    for k in range(1, 7):
        T0 = golden[f"Todom{k}"]
        a = oracle.d2d_match(oracle_fixture_maps[k], oracle_fixture_maps[k + 1], T0, oracle.default_params(delta_score=1e-6))
        b = oracle.d2d_match(oracle_fixture_maps[k], oracle_fixture_maps[k + 1], T0,
                             oracle.default_params(delta_score=1e-6, incremental_cells=1))
        assert (a.converged, a.iterations, a.n_hess_passes) == (b.converged, b.iterations, b.n_hess_passes)
        assert abs(a.n_grad_passes - b.n_grad_passes) <= 2
        assert np.abs(a.pose() - b.pose()).max() < 1e-9, (k, np.abs(a.pose() - b.pose()).max())


def test_oracle_cstep_and_linalg(oracle):
    rng = np.random.default_rng(1)
    for n in (3, 6):
        A = rng.normal(size=(n, n))
        A = A + A.T
        ev, V = oracle.eig_sym(A)
        np.testing.assert_allclose(ev, np.linalg.eigvalsh(A), atol=1e-12)
        np.testing.assert_allclose(V @ np.diag(ev) @ V.T, A, atol=1e-12)
    # dcstep case 1 (higher function value brackets the minimum): cubic/quadratic mix on f(x) = (x-1)^2
    f = lambda x: (x - 1) ** 2
    d = lambda x: 2 * (x - 1)
    info, v, br = oracle.cstep(0.0, f(0), d(0), 0.0, f(0), d(0), 3.0, f(3.0), d(3.0), False, 0.0, 15.0)
    assert info == 1 and br and abs(v[6] - 1.0) < 1e-12  # exact for a quadratic


def test_soft_constraint_pulls_to_prior(golden, oracle, oracle_fixture_maps):
    """ndt_matcher_d2d_fusion.h:11-32 / odom_hessian_test.cpp:82-108: the soft prior is x^T C x with gradient (C + C^T) x
    and Hessian C + C^T.  With a very tight Tcov the prior dominates and the pose must stay at the initial guess."""
    T0 = golden["Todom1"]
    p = oracle.default_params(delta_score=1e-6, use_soft_constraints=1)
    tight = oracle.fusion_match(oracle_fixture_maps[1], oracle_fixture_maps[2], T0, np.eye(6) * 1e-10, p)
    loose = oracle.fusion_match(oracle_fixture_maps[1], oracle_fixture_maps[2], T0, np.eye(6) * 1e+2, p)
    plain = oracle.d2d_match(oracle_fixture_maps[1], oracle_fixture_maps[2], T0, oracle.default_params(delta_score=1e-6))
    assert synth.pose_error(tight.pose(), T0) < 1e-4
    assert synth.pose_error(loose.pose(), plain.pose()) < 5e-3


def test_robust_yaw_cases(oracle):
    """utils_affine_test.cpp:29-59 inputs: yaw recovered also near +-pi roll/pitch flips"""
    for yaw in (-3.1, -1.5, -0.3, 0.0, 0.4, 2.9):
        T = synth.pose2d(1.0, -2.0, yaw)
        assert abs(oracle.robust_yaw(T) - yaw) < 1e-12


def test_oracle_p2d_finite_differences_and_known_answer(oracle):
    """NDTMatcherP2D restatement (no call site in the reference, parity unpinned): its gradient / Hessian are checked by
    central differences of the scalar score, and on a dense 2-D room it registers to the ground truth."""
    from ndt_feature_graph_b200 import synth

    ca, cb, D = synth.laser2d_pair(1, n_rays=3000)
    m = oracle.OracleMap(0.5)
    m.load_point_cloud(ca, -1.0)
    m.compute_cells()
    T = synth.perturb_pose(D, 2, dt=0.05, dr=0.01, planar=True)
    s0, g, H, npairs = oracle.p2d_derivatives(m, cb, T)
    assert npairs > 1000
    eps = 1e-6
    gn, Hn = np.zeros(6), np.zeros((6, 6))
    for i in range(6):
        d = np.zeros(6)
        d[i] = eps
        sp, gp, _, _ = oracle.p2d_derivatives(m, cb, oracle.pose_from_vec(d) @ T, want_hessian=False)
        sm, gm_, _, _ = oracle.p2d_derivatives(m, cb, oracle.pose_from_vec(-d) @ T, want_hessian=False)
        gn[i] = (sp - sm) / (2 * eps)
        Hn[:, i] = (gp - gm_) / (2 * eps)
    assert np.abs(gn - g).max() <= 1e-6 * np.abs(g).max()
    # the increment is applied on the left (T <- TR(p) T), so gradients at p != 0 differ from the p = 0 Hessian only at O(eps)
    assert np.abs(Hn - H).max() <= 2e-4 * np.abs(H).max()
    r = oracle.p2d_match(m, cb, synth.perturb_pose(D, 5, dt=0.1, dr=0.02, planar=True))
    assert r.converged and synth.pose_error(r.pose(), D) < 0.01
