"""Debug helper (GPU box): replay the optimiser state machine on the host with the ORACLE derivatives as evaluator and
compare, at every evaluated pose, the GPU derivativesNDT.  Finds the first pose where the two disagree."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle_py as O  # noqa: E402
import ndt_feature_graph_b200 as N  # noqa: E402
from ndt_feature_graph_b200 import synth  # noqa: E402

so = "/tmp/libhh.so"
subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", so,
                       os.path.join(ROOT, "tests", "harness", "host_harness.cpp")])
hh = C.CDLL(so)


class R(C.Structure):
    _fields_ = [("T", C.c_double * 16), ("score", C.c_double), ("score_best", C.c_double), ("converged", C.c_int),
                ("iterations", C.c_int), ("n_hess", C.c_int), ("n_grad", C.c_int), ("exit_code", C.c_int), ("nonfinite", C.c_int)]


CB = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double))


def trace(om, gm, T0, e):
    m = N.NDTMatcherD2D(e)
    p = O.default_params()
    step = [0]

    def cb(Tp, hess, sums):
        T = np.array([Tp[i] for i in range(16)]).reshape(4, 4).T
        s, g, H, npairs = O.d2d_derivatives(om[0], om[1], T, p, bool(hess))
        sg, gg, Hg, ng = m.derivativesNDT(gm[0], gm[1], T, bool(hess))
        dg = np.abs(gg - g).max() / max(np.abs(g).max(), 1e-300)
        dH = np.abs(Hg - H).max() / max(np.abs(H).max(), 1e-300) if hess else 0
        flag = "" if (npairs == ng and dg < 1e-9 and dH < 1e-9) else "   <<<<<< MISMATCH"
        print(f"  step {step[0]:3d} hess={hess} pairs {npairs}/{ng} score {s:.12f}/{sg:.12f} dg {dg:.2e} dH {dH:.2e}{flag}")
        step[0] += 1
        sums[0] = s
        for i in range(6):
            sums[1 + i] = g[i]
        k = 7
        for a in range(6):
            for b in range(a, 6):
                sums[k] = H[a, b]
                k += 1
        return 0

    r = R()
    T0c = np.ascontiguousarray(T0.T).ravel().copy()
    tc = np.eye(6)
    rc = hh.hh_match(T0c.ctypes.data_as(C.c_void_p), p.itr_max, p.step_control, p.regularize, C.c_double(p.delta_score), 0, 0, 0,
                     tc.ctypes.data_as(C.c_void_p), CB(cb), C.byref(r))
    assert rc == 0
    return r


def build(e, ca, cb, mode):
    om, gm = [], []
    for c in (ca, cb):
        o = O.OracleMap(0.5)
        g = N.NDTMap(e, 0.5)
        if mode == "fixed":
            o.guess_size(0, 0, 0, 100, 100, 4)
            g.guessSize(0, 0, 0, 100, 100, 4)
        o.load_point_cloud(c, 60.0)
        o.compute_cells()
        g.loadPointCloud(c, 60.0)
        g.computeNDTCells()
        om.append(o)
        gm.append(g)
    return om, gm


def main():
    e = N.Engine(0)
    cases = [(synth.laser2d_pair(0), "fixed", True), (synth.velodyne_pair(1, n_rings=32, n_az=600), "guess", False)]
    m = N.NDTMatcherD2D(e)
    for (ca, cb, D), mode, planar in cases:
        om, gm = build(e, ca, cb, mode)
        for seed in range(3):
            T0 = synth.perturb_pose(D, 10 + seed, planar=planar)
            ro = O.d2d_match(om[0], om[1], T0)
            rg = m.match(gm[0], gm[1], T0)
            err = synth.pose_error(ro.pose(), rg.pose())
            print(mode, seed, "err", err, "oracle", (ro.iterations, ro.n_hess_passes, ro.n_grad_passes, ro.exit_code, ro.score),
                  "gpu", (rg.iterations, rg.n_hess_passes, rg.n_grad_passes, rg.exit_code, rg.score))
            if err > 1e-7:
                trace(om, gm, T0, e)


if __name__ == "__main__" and len(sys.argv) == 1:
    main()


def run_sm(evalf, T0, p):
    log = []

    def cb(Tp, hess, sums):
        T = np.array([Tp[i] for i in range(16)]).reshape(4, 4).T
        s, g, H = evalf(T, bool(hess))
        log.append((T.copy(), hess, s, g.copy()))
        sums[0] = s
        for i in range(6):
            sums[1 + i] = g[i]
        k = 7
        for a in range(6):
            for b in range(a, 6):
                sums[k] = H[a, b]
                k += 1
        return 0

    r = R()
    T0c = np.ascontiguousarray(T0.T).ravel().copy()
    tc = np.eye(6)
    hh.hh_match(T0c.ctypes.data_as(C.c_void_p), p.itr_max, p.step_control, p.regularize, C.c_double(p.delta_score), 0, 0, 0,
                tc.ctypes.data_as(C.c_void_p), CB(cb), C.byref(r))
    return r, log


def diverge():
    e = N.Engine(0)
    ca, cb, D = synth.velodyne_pair(1, n_rings=32, n_az=600)
    om, gm = build(e, ca, cb, "guess")
    T0 = synth.perturb_pose(D, 11, planar=False)
    p = O.default_params()
    m = N.NDTMatcherD2D(e)
    ro, lo = run_sm(lambda T, h: O.d2d_derivatives(om[0], om[1], T, p, h)[:3], T0, p)
    rg, lg = run_sm(lambda T, h: m.derivativesNDT(gm[0], gm[1], T, h)[:3], T0, p)
    print("host-sm oracle-eval", ro.iterations, ro.n_hess, ro.n_grad, "host-sm gpu-eval", rg.iterations, rg.n_hess, rg.n_grad)
    for i, (a, b) in enumerate(zip(lo, lg)):
        dT = np.abs(a[0] - b[0]).max()
        print(i, "hess", a[1], b[1], "dT %.3e" % dT, "score %.15g %.15g" % (a[2], b[2]), "dscore %.3e" % (a[2] - b[2]),
              "dg %.3e" % np.abs(a[3] - b[3]).max())
        if dT > 1e-9:
            break


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "diverge":
    diverge()
