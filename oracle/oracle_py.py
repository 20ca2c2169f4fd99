"""ctypes binding of the CPU ORACLE (oracle/ndt_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libndt_oracle.so")


class Grid(C.Structure):
    _fields_ = [("center", C.c_double * 3), ("cell", C.c_double * 3), ("size", C.c_int32 * 3)]


CELL_DTYPE = np.dtype(
    [("mean", "<f8", 3), ("cov", "<f8", 6), ("n", "<i4"), ("has_gaussian", "<i4"), ("idx", "<i4", 3), ("occ", "<f4")],
    align=True,
)
assert CELL_DTYPE.itemsize == 96, CELL_DTYPE.itemsize


class Params(C.Structure):
    _fields_ = [
        ("n_neighbours", C.c_int32),
        ("itr_max", C.c_int32),
        ("step_control", C.c_int32),
        ("regularize", C.c_int32),
        ("delta_score", C.c_double),
        ("lfd1", C.c_double),
        ("lfd2", C.c_double),
        ("use_soft_constraints", C.c_int32),
        ("use_tikhonov", C.c_int32),
        ("n_threads", C.c_int32),
        ("planar", C.c_int32),
        ("incremental_cells", C.c_int32),
        ("reserved_", C.c_int32),
    ]


class Result(C.Structure):
    _fields_ = [
        ("T", C.c_double * 16),
        ("score", C.c_double),
        ("score_best", C.c_double),
        ("converged", C.c_int32),
        ("iterations", C.c_int32),
        ("n_hess_passes", C.c_int32),
        ("n_grad_passes", C.c_int32),
        ("pose_changed", C.c_int32),
        ("exit_code", C.c_int32),
    ]

    def pose(self):
        return np.array(self.T, dtype=np.float64).reshape(4, 4).T.copy()


def build(force=False):
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(
        os.path.join(_HERE, "ndt_oracle.cpp")
    ):
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    vp, dp, i64 = C.c_void_p, C.POINTER(C.c_double), C.c_int64
    L.orc_map_create.restype = vp
    L.orc_map_create.argtypes = [C.c_double] * 3
    L.orc_map_destroy.argtypes = [vp]
    L.orc_map_guess_size.argtypes = [vp] + [C.c_double] * 6
    L.orc_map_set_map_size.argtypes = [vp] + [C.c_double] * 3
    L.orc_map_initialize.argtypes = [vp] + [C.c_double] * 6
    L.orc_map_load_point_cloud.restype = i64
    L.orc_map_load_point_cloud.argtypes = [vp, vp, i64, C.c_double]
    L.orc_map_add_points.restype = i64
    L.orc_map_add_points.argtypes = [vp, vp, i64]
    L.orc_map_add_point_cloud.restype = i64
    L.orc_map_add_point_cloud.argtypes = [vp, vp, vp, i64, C.c_double, C.c_double, C.c_double, C.c_double]
    L.orc_map_compute_cells.argtypes = [vp, C.c_uint32, C.c_float]
    L.orc_map_from_cells.argtypes = [vp, C.POINTER(Grid), vp, i64, C.c_int]
    L.orc_map_grid.argtypes = [vp, C.POINTER(Grid)]
    L.orc_map_num_cells.restype = i64
    L.orc_map_num_cells.argtypes = [vp, C.c_int]
    L.orc_map_export_cells.restype = i64
    L.orc_map_export_cells.argtypes = [vp, vp, i64, C.c_int]
    L.orc_map_point_indices.restype = i64
    L.orc_map_point_indices.argtypes = [vp, vp, i64, vp]
    L.orc_d2d_derivatives.argtypes = [vp, vp, vp, C.POINTER(Params), C.c_int, vp, C.POINTER(i64)]
    L.orc_d2d_match.argtypes = [vp, vp, vp, C.POINTER(Params), C.POINTER(Result)]
    L.orc_d2d_line_search.restype = C.c_double
    L.orc_d2d_line_search.argtypes = [vp, vp, vp, vp, C.POINTER(Params)]
    L.orc_map_load_point_cloud_centroid.restype = i64
    L.orc_map_load_point_cloud_centroid.argtypes = [vp, vp, i64, vp, vp, vp, C.c_double]
    L.orc_p2d_derivatives.argtypes = [vp, vp, i64, vp, C.POINTER(Params), C.c_int, vp, C.POINTER(i64)]
    L.orc_p2d_match.argtypes = [vp, vp, i64, vp, C.POINTER(Params), C.POINTER(Result)]
    L.orc_fusion_match.argtypes = [vp, vp, vp, vp, C.POINTER(Params), C.POINTER(Result)]
    L.orc_d2d_covariance.argtypes = [vp, vp, vp, C.POINTER(Params), vp]
    L.orc_d2d_match_batch.argtypes = [i64, vp, vp, vp, C.POINTER(Params), C.c_int, C.c_int, vp, vp]
    L.orc_overlap_occupancy_score.restype = C.c_double
    L.orc_overlap_occupancy_score.argtypes = [vp, vp, vp]
    L.orc_cstep.argtypes = [dp] * 7 + [C.c_double, C.c_double, C.POINTER(C.c_int), C.c_double, C.c_double]
    L.orc_eig_sym.argtypes = [C.c_int, vp, vp, vp]
    L.orc_pose_from_vec.argtypes = [vp, vp]
    L.orc_robust_yaw.restype = C.c_double
    L.orc_robust_yaw.argtypes = [vp]
    L.orc_default_params.argtypes = [C.POINTER(Params)]
    _lib = L
    return L


def default_params(**kw):
    p = Params()
    lib().orc_default_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def _cm(T):
    """4x4 numpy pose -> column-major contiguous 16 doubles"""
    return np.ascontiguousarray(np.asarray(T, dtype=np.float64).T).ravel().copy()


def _pts4(pts):
    pts = np.asarray(pts, dtype=np.float32)
    if pts.shape[1] == 3:
        pts = np.concatenate([pts, np.zeros((pts.shape[0], 1), np.float32)], axis=1)
    return np.ascontiguousarray(pts)


class OracleMap:
    """lslgeneric::NDTMap(new LazyGrid(cell)) restated on the CPU."""

    def __init__(self, cell=0.5):
        c = (cell, cell, cell) if np.isscalar(cell) else tuple(cell)
        self.h = lib().orc_map_create(*c)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_map_destroy(self.h)
            self.h = None

    def guess_size(self, cx, cy, cz, sx, sy, sz):
        lib().orc_map_guess_size(self.h, cx, cy, cz, sx, sy, sz)

    def set_map_size(self, sx, sy, sz):
        lib().orc_map_set_map_size(self.h, sx, sy, sz)

    def initialize(self, cx, cy, cz, sx, sy, sz):
        lib().orc_map_initialize(self.h, cx, cy, cz, sx, sy, sz)

    def load_point_cloud(self, pts, range_limit=-1.0):
        pts = _pts4(pts)
        return lib().orc_map_load_point_cloud(self.h, pts.ctypes.data, pts.shape[0], range_limit)

    def load_point_cloud_centroid(self, pts, origin, old_centroid, map_size, range_limit):
        pts = _pts4(pts)
        o, c, ms = (np.ascontiguousarray(v, dtype=np.float64) for v in (origin, old_centroid, map_size))
        return lib().orc_map_load_point_cloud_centroid(self.h, pts.ctypes.data, pts.shape[0], o.ctypes.data, c.ctypes.data,
                                                       ms.ctypes.data, range_limit)

    def add_points(self, pts):
        pts = _pts4(pts)
        return lib().orc_map_add_points(self.h, pts.ctypes.data, pts.shape[0])

    def add_point_cloud(self, origin, pts, classifier_th=0.06, maxz=100.0, sensor_noise=0.25, occupancy_limit=255.0):
        """NDTMap::addPointCloud with the free-space ray trace (LazyGrid::traceLine) and occupancy update."""
        pts = _pts4(pts)
        o = np.ascontiguousarray(origin, dtype=np.float64)
        return lib().orc_map_add_point_cloud(self.h, o.ctypes.data, pts.ctypes.data, pts.shape[0], classifier_th, maxz,
                                             sensor_noise, occupancy_limit)

    def compute_cells(self, maxnumpoints=int(1e9), occupancy_limit=255.0):
        lib().orc_map_compute_cells(self.h, maxnumpoints, occupancy_limit)

    def from_cells(self, center, cell, size, cells, use_idx=False):
        g = Grid((C.c_double * 3)(*center), (C.c_double * 3)(*cell), (C.c_int32 * 3)(*size))
        cells = np.ascontiguousarray(cells, dtype=CELL_DTYPE)
        rc = lib().orc_map_from_cells(self.h, C.byref(g), cells.ctypes.data, cells.shape[0], int(use_idx))
        if rc != 0:
            raise ValueError("orc_map_from_cells failed")
        return self

    def grid(self):
        g = Grid()
        lib().orc_map_grid(self.h, C.byref(g))
        return np.array(g.center), np.array(g.cell), np.array(g.size)

    def num_cells(self, gaussian_only=True):
        return lib().orc_map_num_cells(self.h, int(gaussian_only))

    def export_cells(self, gaussian_only=True):
        n = lib().orc_map_num_cells(self.h, int(gaussian_only))
        out = np.zeros(n, dtype=CELL_DTYPE)
        lib().orc_map_export_cells(self.h, out.ctypes.data, n, int(gaussian_only))
        return out

    def point_indices(self, pts):
        pts = _pts4(pts)
        out = np.zeros((pts.shape[0], 3), np.int32)
        lib().orc_map_point_indices(self.h, pts.ctypes.data, pts.shape[0], out.ctypes.data)
        return out


def d2d_derivatives(tgt, src, T, params=None, want_hessian=True):
    p = params or default_params()
    out = np.zeros(43)
    Tc = _cm(T)
    npairs = C.c_int64(0)
    lib().orc_d2d_derivatives(tgt.h, src.h, Tc.ctypes.data, C.byref(p), int(want_hessian), out.ctypes.data, C.byref(npairs))
    return out[0], out[1:7].copy(), out[7:].reshape(6, 6).copy(), npairs.value


def d2d_match(tgt, src, T0, params=None):
    p = params or default_params()
    r = Result()
    Tc = _cm(T0)
    rc = lib().orc_d2d_match(tgt.h, src.h, Tc.ctypes.data, C.byref(p), C.byref(r))
    assert rc == 0
    return r


def p2d_derivatives(tgt, pts, T, params=None, want_hessian=True):
    """NDTMatcherP2D derivatives of the cloud `pts` moved by T against the map (D2D with zero source covariance)."""
    p = params or default_params()
    pts = _pts4(pts)
    out = np.zeros(43)
    Tc = _cm(T)
    npairs = C.c_int64(0)
    lib().orc_p2d_derivatives(tgt.h, pts.ctypes.data, pts.shape[0], Tc.ctypes.data, C.byref(p), int(want_hessian),
                              out.ctypes.data, C.byref(npairs))
    return out[0], out[1:7].copy(), out[7:].reshape(6, 6).copy(), npairs.value


def p2d_match(tgt, pts, T0, params=None):
    p = params or default_params()
    pts = _pts4(pts)
    r = Result()
    Tc = _cm(T0)
    rc = lib().orc_p2d_match(tgt.h, pts.ctypes.data, pts.shape[0], Tc.ctypes.data, C.byref(p), C.byref(r))
    assert rc == 0
    return r


def d2d_line_search(tgt, src, T, increment, params=None):
    """NDTMatcherD2D::lineSearchMT on the source cells moved by T: returns (step, increment possibly negated)."""
    p = params or default_params()
    inc = np.ascontiguousarray(increment, dtype=np.float64).copy()
    Tc = _cm(T)
    step = lib().orc_d2d_line_search(tgt.h, src.h, Tc.ctypes.data, inc.ctypes.data, C.byref(p))
    return step, inc


def d2d_is_stable(tgt, src, T0, base=None, tol=1e-9, **kw):
    """Does the reference algorithm reproduce ITSELF on this registration?  The optimiser is a chain of discontinuous
    decisions (More-Thuente branches, neighbourhoods that change with the pose, best-pose fallback); on ill-conditioned
    starts a 1-ulp change anywhere is amplified to O(1) in the result.  Upstream itself sums per OpenMP thread, so such a
    registration has no single "reference answer" and cannot pin parity.  Stable = same pose (within tol) with
    (a) 3 OpenMP partial sums inside derivativesNDT and (b) the initial translation nudged by one ulp, either way."""
    from numpy import nextafter, inf

    def err(ra, rb):
        A, B = ra.pose(), rb.pose()
        return max(abs(A - B).max(), 0.0)

    r0 = base if base is not None else d2d_match(tgt, src, T0, default_params(**kw))
    if err(r0, d2d_match(tgt, src, T0, default_params(n_threads=3, **kw))) > tol:
        return False
    for axis, direction in ((0, inf), (1, -inf), (2, inf)):
        T = np.array(T0, dtype=np.float64).copy()
        T[axis, 3] = nextafter(T[axis, 3], direction)
        if err(r0, d2d_match(tgt, src, T, default_params(**kw))) > max(tol, 1e-9):
            return False
    return True


def fusion_match(tgt, src, T0, Tcov, params=None):
    p = params or default_params(use_soft_constraints=1)
    r = Result()
    Tc = _cm(T0)
    cov = np.ascontiguousarray(Tcov, dtype=np.float64)
    rc = lib().orc_fusion_match(tgt.h, src.h, Tc.ctypes.data, cov.ctypes.data, C.byref(p), C.byref(r))
    assert rc == 0
    return r


def d2d_covariance(tgt, src, T, params=None):
    p = params or default_params()
    out = np.zeros(36)
    Tc = _cm(T)
    rc = lib().orc_d2d_covariance(tgt.h, src.h, Tc.ctypes.data, C.byref(p), out.ctypes.data)
    return rc, out.reshape(6, 6)


def d2d_match_batch(tgts, srcs, T0s, params=None, with_covariance=False, n_threads=1):
    p = params or default_params()
    n = len(tgts)
    ta = (C.c_void_p * n)(*[m.h for m in tgts])
    sa = (C.c_void_p * n)(*[m.h for m in srcs])
    Tc = np.concatenate([_cm(T) for T in T0s])
    res = (Result * n)()
    cov = np.zeros((n, 36))
    rc = lib().orc_d2d_match_batch(n, ta, sa, Tc.ctypes.data, C.byref(p), int(with_covariance), n_threads, res,
                                   cov.ctypes.data)
    assert rc == 0 or with_covariance
    return list(res), cov.reshape(n, 6, 6)


def overlap_occupancy_score(ref, mov, T):
    Tc = _cm(T)
    return lib().orc_overlap_occupancy_score(ref.h, mov.h, Tc.ctypes.data)


def pose_from_vec(p6):
    p = np.ascontiguousarray(p6, dtype=np.float64)
    out = np.zeros(16)
    lib().orc_pose_from_vec(p.ctypes.data, out.ctypes.data)
    return out.reshape(4, 4).T.copy()


def robust_yaw(T):
    Tc = _cm(T)
    return lib().orc_robust_yaw(Tc.ctypes.data)


def eig_sym(A):
    A = np.ascontiguousarray(A, dtype=np.float64)
    n = A.shape[0]
    ev = np.zeros(n)
    V = np.zeros((n, n))
    lib().orc_eig_sym(n, A.ctypes.data, ev.ctypes.data, V.ctypes.data)
    return ev, V


def cstep(stx, fx, dx, sty, fy, dy, stp, fp, dp, brackt, stmin, stmax):
    v = [C.c_double(x) for x in (stx, fx, dx, sty, fy, dy, stp)]
    b = C.c_int(int(brackt))
    info = lib().orc_cstep(*[C.byref(x) for x in v], fp, dp, C.byref(b), stmin, stmax)
    return info, [x.value for x in v], bool(b.value)
