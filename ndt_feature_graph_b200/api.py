"""ctypes host mirror of the reference's NDT interface on top of the C ABI (include/ndtb.h).

Names and argument meaning follow the reference call sites:
  NDTMap(LazyGrid(res))                 ndt_feature_fuser_hmt.cpp:87,196
  .initialize / .guessSize / .setMapSize / .loadPointCloud / .addPointCloud / .computeNDTCells
                                        ndt_feature_fuser_hmt.cpp:89-94,201-227,485-486
  NDTMatcherD2D().match / .covariance   ndt_feature_graph.cpp:261-298
  .derivativesNDT                       ndt_matcher_d2d_fusion.h:856
  matchFusion                           ndt_matcher_d2d_fusion.h:797-1155
Poses are 4x4 numpy arrays (Eigen::Affine3d).  Everything computes on the GPU; a missing library or
device raises NdtbError (never a silent CPU path).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
HOST, DEVICE = 0, 1


def lib_path():
    # NDTB_LIB: alternative build of the same library (kernel-variant A/B runs); default = the in-tree build
    return os.environ.get("NDTB_LIB") or os.path.join(_HERE, "lib", "libndtb.so")


class NdtbError(RuntimeError):
    pass


class Grid(C.Structure):
    _fields_ = [("center", C.c_double * 3), ("cell", C.c_double * 3), ("size", C.c_int32 * 3)]


CELL_DTYPE = np.dtype(
    [("mean", "<f8", 3), ("cov", "<f8", 6), ("n", "<i4"), ("has_gaussian", "<i4"), ("idx", "<i4", 3), ("occ", "<f4")],
    align=True,
)
assert CELL_DTYPE.itemsize == 96


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
        ("ctas_per_match", C.c_int32),
        ("pass_budget", C.c_int32),
        ("planar", C.c_int32),
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
        ("status", C.c_int32),
        ("n_src_cells", C.c_int32),
        ("n_tgt_cells", C.c_int32),
        ("tgt_table_entries", C.c_int32),
        ("kernel_ms", C.c_float),
        ("n_exec_passes", C.c_int32),
    ]

    def pose(self):
        return np.array(self.T, dtype=np.float64).reshape(4, 4).T.copy()


class FuserParams(C.Structure):
    """ndtb_fuser_params: NDTFeatureFuserHMT::Params (ndt_feature_fuser_hmt.h:58-207) + sensor pose + motion model."""

    _fields_ = [
        ("resolution", C.c_double), ("map_size_x", C.c_double), ("map_size_y", C.c_double), ("map_size_z", C.c_double),
        ("sensor_range", C.c_double), ("max_translation_norm", C.c_double), ("max_rotation_norm", C.c_double),
        ("delta_score", C.c_double), ("neighbours", C.c_int32), ("itr_max", C.c_int32), ("step_control", C.c_int32),
        ("global_transf", C.c_int32), ("use_soft_constraints", C.c_int32), ("use_tikhonov", C.c_int32),
        ("compute_cov", C.c_int32), ("fusion2d", C.c_int32), ("all_matches_valid", C.c_int32),
        ("fuse_incomplete", C.c_int32), ("check_consistency", C.c_int32), ("force_odom_as_est", C.c_int32),
        ("sensor_pose", C.c_double * 16), ("motion", C.c_double * 6),
    ]


RESULT_DTYPE = np.dtype(
    [("T", "<f8", 16), ("score", "<f8"), ("score_best", "<f8"), ("converged", "<i4"), ("iterations", "<i4"),
     ("n_hess_passes", "<i4"), ("n_grad_passes", "<i4"), ("pose_changed", "<i4"), ("exit_code", "<i4"),
     ("status", "<i4"), ("n_src_cells", "<i4"), ("n_tgt_cells", "<i4"), ("tgt_table_entries", "<i4"), ("kernel_ms", "<f4"), ("n_exec_passes", "<i4")]
)
assert RESULT_DTYPE.itemsize == C.sizeof(Result) == 192

_lib = None


def load_library():
    """dlopen lib/libndtb.so and declare the ABI.  Raises NdtbError when the library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise NdtbError(f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(nvcc, sm_100a). There is no CPU fallback.")
    L = C.CDLL(path)
    vp, i64, dbl = C.c_void_p, C.c_int64, C.c_double
    PP, PR = C.POINTER(Params), C.POINTER(Result)
    sig = {
        "ndtb_version": (C.c_int, []),
        "ndtb_strerror": (C.c_char_p, [C.c_int]),
        "ndtb_last_error": (C.c_char_p, [vp]),
        "ndtb_ctx_create": (C.c_int, [C.c_int, vp, C.POINTER(vp)]),
        "ndtb_ctx_destroy": (None, [vp]),
        "ndtb_ctx_synchronize": (C.c_int, [vp]),
        "ndtb_ctx_launch_count": (i64, [vp]),
        "ndtb_ctx_sm_count": (C.c_int, [vp]),
        "ndtb_ctx_enable_timing": (C.c_int, [vp, C.c_int]),
        "ndtb_ctx_match_time": (C.c_int, [vp, C.POINTER(dbl), C.POINTER(i64)]),
        "ndtb_ctx_build_time": (C.c_int, [vp, C.POINTER(dbl), C.POINTER(i64)]),
        "ndtb_default_params": (None, [PP]),
        "ndtb_map_create": (C.c_int, [vp, dbl, dbl, dbl, C.POINTER(vp)]),
        "ndtb_map_destroy": (None, [vp]),
        "ndtb_map_guess_size": (C.c_int, [vp] + [dbl] * 6),
        "ndtb_map_set_map_size": (C.c_int, [vp] + [dbl] * 3),
        "ndtb_map_initialize": (C.c_int, [vp] + [dbl] * 6),
        "ndtb_map_load_point_cloud": (C.c_int, [vp, vp, i64, dbl, C.c_int, C.POINTER(i64)]),
        "ndtb_map_add_points": (C.c_int, [vp, vp, i64, C.c_int, C.POINTER(i64)]),
        "ndtb_map_compute_cells": (C.c_int, [vp, C.c_uint32, C.c_float]),
        "ndtb_map_build_batch": (C.c_int, [vp, i64, vp, vp, vp, dbl, C.c_int, C.c_uint32, C.c_float]),
        "ndtb_map_from_cells": (C.c_int, [vp, C.POINTER(Grid), vp, i64, C.c_int]),
        "ndtb_map_grid": (C.c_int, [vp, C.POINTER(Grid)]),
        "ndtb_map_num_cells": (i64, [vp, C.c_int]),
        "ndtb_map_export_cells": (i64, [vp, vp, i64, C.c_int]),
        "ndtb_map_point_indices": (i64, [vp, vp, i64, C.c_int, vp]),
        "ndtb_d2d_derivatives": (C.c_int, [vp, vp, vp, vp, PP, C.c_int, vp, C.POINTER(i64)]),
        "ndtb_d2d_match": (C.c_int, [vp, vp, vp, vp, PP, PR]),
        "ndtb_fusion_match": (C.c_int, [vp, vp, vp, vp, vp, PP, PR]),
        "ndtb_d2d_covariance": (C.c_int, [vp, vp, vp, vp, PP, vp]),
        "ndtb_p2d_derivatives": (C.c_int, [vp, vp, vp, i64, C.c_int, vp, PP, C.c_int, vp, C.POINTER(i64)]),
        "ndtb_p2d_match": (C.c_int, [vp, vp, vp, i64, C.c_int, vp, PP, PR]),
        "ndtb_d2d_match_batch": (C.c_int, [vp, i64, vp, vp, vp, PP, C.c_int, C.c_int, vp, vp]),
        "ndtb_register_scans": (C.c_int, [vp, i64, vp, vp, vp, vp, vp, dbl, vp, dbl, PP, C.c_int, C.c_int, C.c_int, vp, vp]),
        "ndtb_jff_write_cells": (C.c_int, [C.c_char_p, C.POINTER(Grid), vp, i64]),
        "ndtb_jff_read_cells": (C.c_int, [C.c_char_p, C.POINTER(Grid), vp, i64, C.POINTER(i64)]),
        "ndtb_map_write_jff": (C.c_int, [vp, C.c_char_p]),
        "ndtb_map_load_jff": (C.c_int, [vp, C.c_char_p]),
        "ndtb_overlap_score": (C.c_int, [vp, vp, vp, vp, C.POINTER(dbl)]),
        "ndtb_d2d_derivatives_cells": (C.c_int, [vp, vp, vp, i64, vp, PP, C.c_int, vp, C.POINTER(i64)]),
        "ndtb_d2d_line_search_cells": (C.c_int, [vp, vp, vp, i64, vp, PP, C.POINTER(dbl)]),
        "ndtb_mt_cstep": (C.c_int, [C.POINTER(dbl)] * 7 + [dbl, dbl, C.POINTER(C.c_int), dbl, dbl]),
        "ndtb_eig_sym3": (C.c_int, [vp, C.c_int, vp, vp, C.POINTER(C.c_int32)]),
        "ndtb_map_load_point_cloud_centroid": (C.c_int, [vp, vp, i64, C.c_int, vp, vp, vp, dbl]),
        "ndtb_overlap_score_batch": (C.c_int, [vp, i64, vp, vp, vp, i64, C.c_int, C.c_int, vp]),
        "ndtb_edge_msg_pack": (i64, [C.c_uint32, C.c_uint32, vp, vp, vp, dbl, vp, i64]),
        "ndtb_edge_msg_unpack": (C.c_int, [vp, i64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), vp, vp, vp, C.POINTER(C.c_int32), C.POINTER(dbl)]),
        "ndtb_map_msg_pack": (i64, [C.c_uint32, C.c_uint32, C.c_uint32, C.c_char_p, vp, vp, i64, vp, i64]),
        "ndtb_map_msg_unpack": (C.c_int, [vp, i64, vp, C.c_char_p, C.c_int32, vp, vp, i64, C.POINTER(i64), C.POINTER(i64)]),
        "ndtb_node_msg_pack": (i64, [vp, vp, i64, vp, i64]),
        "ndtb_node_msg_unpack": (C.c_int, [vp, i64, vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)]),
        "ndtb_graph_msg_pack": (i64, [C.c_uint32, C.c_uint32, C.c_uint32, C.c_char_p, vp, vp, dbl, i64, vp, vp, i64, vp, vp, vp, i64]),
        "ndtb_graph_msg_unpack": (C.c_int, [vp, i64, vp, C.c_char_p, C.c_int32, vp, vp, C.POINTER(dbl), C.POINTER(i64), vp, vp, i64,
                                            C.POINTER(i64), vp, vp, i64]),
        "ndtb_pose_archive_write": (C.c_int, [C.c_char_p, vp]),
        "ndtb_pose_archive_read": (C.c_int, [C.c_char_p, vp]),
        "ndtb_eval_string": (C.c_int, [vp, C.c_int, C.c_char_p, C.c_int32]),
        "ndtb_comm_unique_id": (C.c_int, [vp]),
        "ndtb_comm_create": (C.c_int, [vp, vp, C.c_int, C.c_int, C.POINTER(vp)]),
        "ndtb_comm_destroy": (None, [vp]),
        "ndtb_gather_results": (C.c_int, [vp, vp, i64, vp]),
        "ndtb_map_add_point_cloud": (C.c_int, [vp, vp, vp, i64, C.c_int, dbl, dbl, dbl, dbl]),
        "ndtb_transform_point_cloud": (C.c_int, [vp, vp, vp, i64, C.c_int, vp, C.c_int]),
        "ndtb_fuser_default_params": (None, [C.POINTER(FuserParams)]),
        "ndtb_fuser_create": (C.c_int, [vp, C.POINTER(FuserParams), C.POINTER(vp)]),
        "ndtb_fuser_destroy": (None, [vp]),
        "ndtb_fuser_initialize": (C.c_int, [vp, vp, vp, i64, C.c_int]),
        "ndtb_fuser_update": (C.c_int, [vp, vp, vp, i64, C.c_int, C.c_int, vp, PR, vp]),
        "ndtb_fuser_map": (vp, [vp]),
        "ndtb_fuser_pose": (C.c_int, [vp, vp]),
        "ndtb_fuser_set_pose": (C.c_int, [vp, vp]),
        "ndtb_graph_create": (C.c_int, [vp, C.POINTER(FuserParams), dbl, C.POINTER(vp)]),
        "ndtb_graph_destroy": (None, [vp]),
        "ndtb_graph_set_new_node_dist": (C.c_int, [vp, dbl]),
        "ndtb_graph_initialize": (C.c_int, [vp, vp, vp, i64, C.c_int]),
        "ndtb_graph_update": (C.c_int, [vp, vp, vp, i64, C.c_int, vp]),
        "ndtb_graph_num_nodes": (i64, [vp]),
        "ndtb_graph_node": (C.c_int, [vp, i64, vp, vp, vp, C.POINTER(vp), C.POINTER(C.c_int32)]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    L._abi = sorted(sig)
    _lib = L
    return L


def abi_symbols():
    return load_library()._abi


def _cm(T):
    return np.ascontiguousarray(np.asarray(T, dtype=np.float64).T).ravel().copy()


def _pts4(pts):
    pts = np.asarray(pts, dtype=np.float32)
    if pts.ndim != 2 or pts.shape[1] not in (3, 4):
        raise ValueError("points must be [n,3] or [n,4]")
    if pts.shape[1] == 3:
        pts = np.concatenate([pts, np.zeros((pts.shape[0], 1), np.float32)], axis=1)
    return np.ascontiguousarray(pts)


def jff_write_cells(path, center, cell, size, cells):
    """NDTMap::writeToJFF format from host arrays (no GPU): cells = structured array CELL_DTYPE with voxel indices."""
    L = load_library()
    g = Grid((C.c_double * 3)(*center), (C.c_double * 3)(*cell), (C.c_int32 * 3)(*[int(s) for s in size]))
    cells = np.ascontiguousarray(cells, dtype=CELL_DTYPE)
    rc = L.ndtb_jff_write_cells(str(path).encode(), C.byref(g), cells.ctypes.data, cells.shape[0])
    if rc != 0:
        raise NdtbError(f"jff write failed: {L.ndtb_strerror(rc).decode()}")


def jff_read_cells(path):
    """NDTMap::loadFromJFF format into host arrays (no GPU): returns (center, cell, size, cells)."""
    L = load_library()
    g = Grid()
    n = C.c_int64(0)
    rc = L.ndtb_jff_read_cells(str(path).encode(), C.byref(g), None, 0, C.byref(n))
    if rc != 0:
        raise NdtbError(f"jff read failed: {L.ndtb_strerror(rc).decode()}")
    cells = np.zeros(max(n.value, 1), CELL_DTYPE)
    rc = L.ndtb_jff_read_cells(str(path).encode(), C.byref(g), cells.ctypes.data, n.value, C.byref(n))
    if rc != 0:
        raise NdtbError(f"jff read failed: {L.ndtb_strerror(rc).decode()}")
    return np.array(g.center), np.array(g.cell), np.array(g.size), cells[: n.value].copy()


class Engine:
    """One ndtb_ctx: a GPU + a stream.  Single owner (one per host thread / per GPU)."""

    def __init__(self, device=0, stream=None):
        self.L = load_library()
        h = C.c_void_p()
        rc = self.L.ndtb_ctx_create(int(device), C.c_void_p(stream) if stream else None, C.byref(h))
        if rc != 0:
            raise NdtbError(f"ndtb_ctx_create(device={device}) failed: {self.L.ndtb_strerror(rc).decode()}")
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.L.ndtb_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc != 0:
            raise NdtbError(f"{self.L.ndtb_strerror(rc).decode()} [{self.L.ndtb_last_error(self.h).decode()}]")

    def synchronize(self):
        self.check(self.L.ndtb_ctx_synchronize(self.h))

    @property
    def launch_count(self):
        return int(self.L.ndtb_ctx_launch_count(self.h))

    @property
    def sm_count(self):
        return int(self.L.ndtb_ctx_sm_count(self.h))

    def enable_timing(self, on=True):
        self.check(self.L.ndtb_ctx_enable_timing(self.h, int(on)))

    def match_time(self):
        """(milliseconds, launches) of the registration kernel since the last call (synchronises)."""
        ms, n = C.c_double(0), C.c_int64(0)
        self.check(self.L.ndtb_ctx_match_time(self.h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def build_time(self):
        """(milliseconds, calls) of the batched map builds since the last call (synchronises)."""
        ms, n = C.c_double(0), C.c_int64(0)
        self.check(self.L.ndtb_ctx_build_time(self.h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def register_scans_raw(self, n, tgt_ptrs, n_tgt, src_ptrs, n_src, T0s_cm, cell, range_limit, params, with_covariance,
                           in_mem, out_mem, res_ptr, cov_ptr):
        """Pointer-level ndtb_register_scans (bench.py: device-resident inputs / outputs)."""
        self.check(self.L.ndtb_register_scans(self.h, n, tgt_ptrs, n_tgt, src_ptrs, n_src, T0s_cm, cell, None, range_limit,
                                              C.byref(params), int(with_covariance), in_mem, out_mem, res_ptr, cov_ptr))

    def default_params(self, **kw):
        p = Params()
        self.L.ndtb_default_params(C.byref(p))
        for k, v in kw.items():
            setattr(p, k, v)
        return p

    # ---- batched entry points -------------------------------------------------------------
    def build_maps(self, maps, clouds, range_limit=-1.0, maxnumpoints=0xFFFFFFFF, occupancy_limit=255.0):
        """loadPointCloud + computeNDTCells for many maps in a handful of launches (host clouds)."""
        n = len(maps)
        cl = [_pts4(c) for c in clouds]
        mh = (C.c_void_p * n)(*[m.h for m in maps])
        ph = (C.c_void_p * n)(*[c.ctypes.data for c in cl])
        nn = (C.c_int64 * n)(*[c.shape[0] for c in cl])
        self.check(self.L.ndtb_map_build_batch(self.h, n, mh, ph, nn, range_limit, HOST, maxnumpoints, occupancy_limit))

    def match_batch(self, tgts, srcs, T0s, params=None, with_covariance=False):
        """NDTFeatureGraph::updateLinksUsingNDTRegistration (ndt_feature_graph.cpp:347-353) as one launch."""
        p = params or self.default_params()
        n = len(tgts)
        ta = (C.c_void_p * n)(*[m.h for m in tgts])
        sa = (C.c_void_p * n)(*[m.h for m in srcs])
        Tc = np.concatenate([_cm(T) for T in T0s]) if n else np.zeros(0)
        res = np.zeros(n, RESULT_DTYPE)
        cov = np.zeros((n, 36))
        self.check(self.L.ndtb_d2d_match_batch(self.h, n, ta, sa, Tc.ctypes.data, C.byref(p), int(with_covariance), HOST,
                                               res.ctypes.data, cov.ctypes.data if with_covariance else None))
        return res, cov.reshape(n, 6, 6)

    def overlap_scores(self, refs, movs, Ts):
        """overlapNDTOccupancyScore of n links in one launch (ndt_feature_graph.cpp:335-342)."""
        n = len(refs)
        ra = (C.c_void_p * n)(*[m.h for m in refs])
        ma = (C.c_void_p * n)(*[m.h for m in movs])
        Tc = np.concatenate([_cm(T) for T in Ts]) if n else np.zeros(0)
        out = np.zeros(n)
        self.check(self.L.ndtb_overlap_score_batch(self.h, n, ra, ma, Tc.ctypes.data, 128, HOST, HOST, out.ctypes.data))
        return out

    def register_scans(self, tgt_clouds, src_clouds, T0s, cell=0.5, map_size=None, range_limit=-1.0, params=None,
                       with_covariance=False):
        """Front-end step for a batch of scan pairs from HOST clouds: build both local maps, match, covariance."""
        p = params or self.default_params()
        n = len(tgt_clouds)
        tc = [_pts4(c) for c in tgt_clouds]
        sc = [_pts4(c) for c in src_clouds]
        tp = (C.c_void_p * n)(*[c.ctypes.data for c in tc])
        sp = (C.c_void_p * n)(*[c.ctypes.data for c in sc])
        tn = (C.c_int64 * n)(*[c.shape[0] for c in tc])
        sn = (C.c_int64 * n)(*[c.shape[0] for c in sc])
        Tc = np.concatenate([_cm(T) for T in T0s])
        ms = np.asarray(map_size, dtype=np.float64) if map_size is not None else None
        res = np.zeros(n, RESULT_DTYPE)
        cov = np.zeros((n, 36))
        self.check(self.L.ndtb_register_scans(self.h, n, tp, tn, sp, sn, Tc.ctypes.data, cell,
                                              ms.ctypes.data if ms is not None else None, range_limit, C.byref(p),
                                              int(with_covariance), HOST, HOST, res.ctypes.data,
                                              cov.ctypes.data if with_covariance else None))
        return res, cov.reshape(n, 6, 6)


def edge_msg_pack(ref_idx, mov_idx, T, cov3, cov6, score):
    """ROS1 wire bytes of ndt_feature/NDTEdgeMsg (edgeToMsg, ndtgraph_conversion.h:17-34)."""
    L = load_library()
    Tc = _cm(T)
    c3 = np.ascontiguousarray(cov3, dtype=np.float64)
    c6 = np.ascontiguousarray(cov6, dtype=np.float64) if cov6 is not None else None
    n = L.ndtb_edge_msg_pack(ref_idx, mov_idx, Tc.ctypes.data, c3.ctypes.data, c6.ctypes.data if c6 is not None else None, score, None, 0)
    buf = np.zeros(n, np.uint8)
    L.ndtb_edge_msg_pack(ref_idx, mov_idx, Tc.ctypes.data, c3.ctypes.data, c6.ctypes.data if c6 is not None else None, score, buf.ctypes.data, n)
    return buf.tobytes()


def edge_msg_unpack(data):
    L = load_library()
    buf = np.frombuffer(data, np.uint8)
    a, b, has, s = C.c_uint32(), C.c_uint32(), C.c_int32(), C.c_double()
    T, c3, c6 = np.zeros(16), np.zeros(9), np.zeros(36)
    rc = L.ndtb_edge_msg_unpack(buf.ctypes.data, len(buf), C.byref(a), C.byref(b), T.ctypes.data, c3.ctypes.data, c6.ctypes.data,
                                C.byref(has), C.byref(s))
    if rc != 0:
        raise NdtbError("malformed NDTEdgeMsg")
    return a.value, b.value, T.reshape(4, 4).T.copy(), c3.reshape(3, 3), (c6.reshape(6, 6) if has.value else None), s.value


class NodeFields(C.Structure):
    """ndtb_node_fields: NDTNodeMsg + NDTFeatureFuserHMTMsg without the map (poses column-major)."""
    _fields_ = [("Tnow", C.c_double * 16), ("Tlast_fuse", C.c_double * 16), ("Todom", C.c_double * 16), ("ctr", C.c_uint32),
                ("nb_updates", C.c_uint32), ("T", C.c_double * 16), ("cov9", C.c_double * 9), ("Tlocal_odom", C.c_double * 16),
                ("Tlocal_fuse", C.c_double * 16), ("time_last_update", C.c_double)]

    POSES = ("Tnow", "Tlast_fuse", "Todom", "T", "Tlocal_odom", "Tlocal_fuse")

    @classmethod
    def make(cls, cov=None, ctr=0, nb_updates=0, time_last_update=0.0, **poses):
        f = cls()
        for k in cls.POSES:
            getattr(f, k)[:] = _cm(poses.get(k, np.eye(4))).tolist()
        f.cov9[:] = np.ascontiguousarray(np.eye(3) if cov is None else cov, dtype=np.float64).ravel().tolist()
        f.ctr, f.nb_updates, f.time_last_update = ctr, nb_updates, time_last_update
        return f

    def pose(self, k):
        return np.array(getattr(self, k)[:]).reshape(4, 4).T.copy()


def _sized(call):
    n = call(None, 0)
    if n < 0:
        raise NdtbError("cannot pack message")
    buf = np.zeros(n, np.uint8)
    call(buf.ctypes.data, n)
    return buf.tobytes()


def map_msg_pack(grid, cells, frame_id="/world", stamp=(0, 0, 0)):
    """ndt_map/NDTMapMsg of lslgeneric::toMessage [upstream] from a map's grid and ALL its cells (Gaussian cells are written)."""
    L = load_library()
    cells = np.ascontiguousarray(cells, dtype=CELL_DTYPE)
    return _sized(lambda p, n: L.ndtb_map_msg_pack(stamp[0], stamp[1], stamp[2], frame_id.encode(), C.byref(grid), cells.ctypes.data,
                                                   len(cells), p, n))


def map_msg_unpack(data):
    """-> (grid, Gaussian cells without voxel indices: insert with from_cells(use_idx=False), frame_id, stamp, bytes consumed)"""
    L = load_library()
    buf = np.frombuffer(data, np.uint8)
    g, n, used = Grid(), C.c_int64(), C.c_int64()
    st, frame = (C.c_uint32 * 3)(), C.create_string_buffer(256)
    if L.ndtb_map_msg_unpack(buf.ctypes.data, len(buf), st, frame, 256, C.byref(g), None, 0, C.byref(n), C.byref(used)) != 0:
        raise NdtbError("malformed NDTMapMsg")
    cells = np.zeros(max(n.value, 1), CELL_DTYPE)
    L.ndtb_map_msg_unpack(buf.ctypes.data, len(buf), st, frame, 256, C.byref(g), cells.ctypes.data, len(cells), C.byref(n), C.byref(used))
    return g, cells[: n.value], frame.value.decode(), tuple(st), used.value


def node_msg_pack(fields, map_msg):
    """ndt_feature/NDTNodeMsg (nodeToMsg, ndtgraph_conversion.h:47-57)."""
    L = load_library()
    m = np.frombuffer(map_msg, np.uint8)
    return _sized(lambda p, n: L.ndtb_node_msg_pack(C.byref(fields), m.ctypes.data, len(m), p, n))


def node_msg_unpack(data):
    """-> (NodeFields, NDTMapMsg bytes, bytes consumed) (msgToNode, ndtgraph_conversion.h:147-187)"""
    L = load_library()
    buf = np.frombuffer(data, np.uint8)
    f, mo, ml, used = NodeFields(), C.c_int64(), C.c_int64(), C.c_int64()
    if L.ndtb_node_msg_unpack(buf.ctypes.data, len(buf), C.byref(f), C.byref(mo), C.byref(ml), C.byref(used)) != 0:
        raise NdtbError("malformed NDTNodeMsg")
    return f, bytes(data[mo.value: mo.value + ml.value]), used.value


def graph_msg_pack(sensor_pose, Tnow, distance_moved, node_msgs, edge_msgs, frame_id="/world", stamp=(0, 0, 0)):
    """ndt_feature/NDTGraphMsg (NDTGraphToMsg, ndtgraph_conversion.h:59-83)."""
    L = load_library()
    sp, tn = _cm(sensor_pose), _cm(Tnow)
    nb = [np.frombuffer(m, np.uint8) for m in node_msgs]
    eb = [np.frombuffer(m, np.uint8) for m in edge_msgs]
    npp = (C.c_void_p * max(len(nb), 1))(*[b.ctypes.data for b in nb])
    epp = (C.c_void_p * max(len(eb), 1))(*[b.ctypes.data for b in eb])
    nl = np.array([len(b) for b in nb] or [0], np.int64)
    el = np.array([len(b) for b in eb] or [0], np.int64)
    return _sized(lambda p, n: L.ndtb_graph_msg_pack(stamp[0], stamp[1], stamp[2], frame_id.encode(), sp.ctypes.data, tn.ctypes.data,
                                                     float(distance_moved), len(nb), npp, nl.ctypes.data, len(eb), epp, el.ctypes.data, p, n))


def graph_msg_unpack(data):
    """-> dict(sensor_pose, Tnow, distance_moved, nodes=[NDTNodeMsg bytes], edges=[NDTEdgeMsg bytes], frame_id, stamp)
    (msgToNDTGraph, ndtgraph_conversion.h:189-216)"""
    L = load_library()
    buf = np.frombuffer(data, np.uint8)
    st, frame = (C.c_uint32 * 3)(), C.create_string_buffer(256)
    sp, tn, dist, nn, ne = np.zeros(16), np.zeros(16), C.c_double(), C.c_int64(), C.c_int64()
    if L.ndtb_graph_msg_unpack(buf.ctypes.data, len(buf), st, frame, 256, sp.ctypes.data, tn.ctypes.data, C.byref(dist), C.byref(nn), None,
                               None, 0, C.byref(ne), None, None, 0) != 0:
        raise NdtbError("malformed NDTGraphMsg")
    no, nl = np.zeros(max(nn.value, 1), np.int64), np.zeros(max(nn.value, 1), np.int64)
    eo, el = np.zeros(max(ne.value, 1), np.int64), np.zeros(max(ne.value, 1), np.int64)
    L.ndtb_graph_msg_unpack(buf.ctypes.data, len(buf), st, frame, 256, sp.ctypes.data, tn.ctypes.data, C.byref(dist), C.byref(nn),
                            no.ctypes.data, nl.ctypes.data, len(no), C.byref(ne), eo.ctypes.data, el.ctypes.data, len(eo))
    return {"sensor_pose": sp.reshape(4, 4).T.copy(), "Tnow": tn.reshape(4, 4).T.copy(), "distance_moved": dist.value,
            "nodes": [bytes(data[no[i]: no[i] + nl[i]]) for i in range(nn.value)],
            "edges": [bytes(data[eo[i]: eo[i] + el[i]]) for i in range(ne.value)], "frame_id": frame.value.decode(), "stamp": tuple(st)}


def pose_archive_write(path, T):
    Tc = _cm(T)
    if load_library().ndtb_pose_archive_write(str(path).encode(), Tc.ctypes.data) != 0:
        raise NdtbError(f"cannot write {path}")


def pose_archive_read(path):
    T = np.zeros(16)
    if load_library().ndtb_pose_archive_read(str(path).encode(), T.ctypes.data) != 0:
        raise NdtbError(f"cannot read {path}")
    return T.reshape(4, 4).T.copy()


def eval_string(T, planar=False):
    buf = C.create_string_buffer(512)
    Tc = _cm(T)
    n = load_library().ndtb_eval_string(Tc.ctypes.data, int(planar), buf, 512)
    if n < 0:
        raise NdtbError("eval string")
    return buf.value.decode()


class Comm:
    """ndtb_comm: the NCCL communicator of a sharded batch (one rank per GPU); gather() = ndtb_gather_results."""

    @staticmethod
    def unique_id():
        L = load_library()
        buf = C.create_string_buffer(128)
        rc = L.ndtb_comm_unique_id(buf)
        if rc != 0:
            raise NdtbError("ndtb_comm_unique_id failed (libnccl.so.2 not loadable?)")
        return buf.raw

    def __init__(self, engine, unique_id, rank, world):
        self.e = engine
        h = C.c_void_p()
        idb = C.create_string_buffer(bytes(unique_id), 128)
        engine.check(engine.L.ndtb_comm_create(engine.h, idb, int(rank), int(world), C.byref(h)))
        self.h, self.rank, self.world = h, rank, world

    def gather(self, local_dev_ptr, n_local, all_dev_ptr):
        self.e.check(self.e.L.ndtb_gather_results(self.h, local_dev_ptr, int(n_local), all_dev_ptr))

    def close(self):
        if getattr(self, "h", None):
            self.e.L.ndtb_comm_destroy(self.h)
            self.h = None


class LazyGrid:
    """lslgeneric::LazyGrid(cellSize): only carries the resolution, the grid lives in the NDTMap."""

    def __init__(self, cell_size):
        self.cell = (cell_size,) * 3 if np.isscalar(cell_size) else tuple(cell_size)


class NDTMap:
    """lslgeneric::NDTMap(new LazyGrid(res)) resident in HBM."""

    def __init__(self, engine, index=0.5, borrowed_handle=None):
        self.e = engine
        self.owned = borrowed_handle is None
        if borrowed_handle is not None:  # a map owned by a fuser / graph node
            self.h = C.c_void_p(borrowed_handle)
            return
        cell = index.cell if isinstance(index, LazyGrid) else LazyGrid(index).cell
        h = C.c_void_p()
        engine.check(engine.L.ndtb_map_create(engine.h, *cell, C.byref(h)))
        self.h = h

    def __del__(self):
        try:
            if self.owned and getattr(self, "h", None) and getattr(self.e, "h", None):
                self.e.L.ndtb_map_destroy(self.h)
            self.h = None
        except Exception:
            pass

    def addPointCloudTraced(self, origin, pts, classifierTh=0.06, maxz=100.0, sensor_noise=0.25, occupancy_limit=255.0):
        """NDTMap::addPointCloud(origin, pc, classifierTh, maxz, sensor_noise, occupancy_limit) with the free-space ray
        trace (ndt_feature_fuser_hmt.cpp:92,485); takes effect at the next computeNDTCells."""
        pts = _pts4(pts)
        o = np.ascontiguousarray(origin, dtype=np.float64)
        self.e.check(self.e.L.ndtb_map_add_point_cloud(self.h, o.ctypes.data, pts.ctypes.data, pts.shape[0], HOST, classifierTh,
                                                       maxz, sensor_noise, occupancy_limit))

    def guessSize(self, cx, cy, cz, sx, sy, sz):
        self.e.check(self.e.L.ndtb_map_guess_size(self.h, cx, cy, cz, sx, sy, sz))

    def setMapSize(self, sx, sy, sz):
        self.e.check(self.e.L.ndtb_map_set_map_size(self.h, sx, sy, sz))

    def initialize(self, cx, cy, cz, sx, sy, sz):
        self.e.check(self.e.L.ndtb_map_initialize(self.h, cx, cy, cz, sx, sy, sz))

    def loadPointCloud(self, pts, range_limit=-1.0, want_count=True):
        pts = _pts4(pts)
        nb = C.c_int64(0)
        self.e.check(self.e.L.ndtb_map_load_point_cloud(self.h, pts.ctypes.data, pts.shape[0], range_limit, HOST,
                                                        C.byref(nb) if want_count else None))
        return nb.value

    def loadPointCloudCentroid(self, pts, origin, old_centroid, map_size, range_limit):
        """NDTMap::loadPointCloudCentroid (ndt_feature_fuser_hmt.cpp:199-217)."""
        pts = _pts4(pts)
        o, c, ms = (np.ascontiguousarray(v, dtype=np.float64) for v in (origin, old_centroid, map_size))
        self.e.check(self.e.L.ndtb_map_load_point_cloud_centroid(self.h, pts.ctypes.data, pts.shape[0], HOST, o.ctypes.data,
                                                                 c.ctypes.data, ms.ctypes.data, float(range_limit)))

    def addPointCloud(self, pts, want_count=True):
        """End-point binning part of NDTMap::addPointCloud (no free-space ray tracing)."""
        pts = _pts4(pts)
        nb = C.c_int64(0)
        self.e.check(self.e.L.ndtb_map_add_points(self.h, pts.ctypes.data, pts.shape[0], HOST,
                                                  C.byref(nb) if want_count else None))
        return nb.value

    def computeNDTCells(self, maxnumpoints=int(1e9), occupancy_limit=255.0):
        self.e.check(self.e.L.ndtb_map_compute_cells(self.h, int(maxnumpoints), float(occupancy_limit)))

    def from_cells(self, center, cell, size, cells, use_idx=False):
        g = Grid((C.c_double * 3)(*center), (C.c_double * 3)(*cell), (C.c_int32 * 3)(*[int(s) for s in size]))
        cells = np.ascontiguousarray(cells, dtype=CELL_DTYPE)
        self.e.check(self.e.L.ndtb_map_from_cells(self.h, C.byref(g), cells.ctypes.data, cells.shape[0], int(use_idx)))
        return self

    def grid(self):
        g = Grid()
        self.e.check(self.e.L.ndtb_map_grid(self.h, C.byref(g)))
        return np.array(g.center), np.array(g.cell), np.array(g.size)

    def numberOfActiveCells(self):
        return int(self.e.L.ndtb_map_num_cells(self.h, 1))

    def num_cells(self, gaussian_only=True):
        return int(self.e.L.ndtb_map_num_cells(self.h, int(gaussian_only)))

    def export_cells(self, gaussian_only=True):
        n = self.num_cells(False)
        out = np.zeros(max(n, 1), dtype=CELL_DTYPE)
        k = self.e.L.ndtb_map_export_cells(self.h, out.ctypes.data, n, int(gaussian_only))
        if k < 0:
            self.e.check(int(k))
        return out[:k].copy()

    def point_indices(self, pts):
        pts = _pts4(pts)
        out = np.zeros((pts.shape[0], 3), np.int32)
        k = self.e.L.ndtb_map_point_indices(self.h, pts.ctypes.data, pts.shape[0], HOST, out.ctypes.data)
        if k < 0:
            self.e.check(int(k))
        return out, int(k)

    def writeToJFF(self, path):
        """NDTMap::writeToJFF (ndt_feature_fuser_hmt.cpp:15): 0 on success."""
        return int(self.e.L.ndtb_map_write_jff(self.h, str(path).encode()))

    def loadFromJFF(self, path):
        """NDTMap::loadFromJFF (ndt_feature_fuser_hmt.cpp:24,39): 0 on success."""
        return int(self.e.L.ndtb_map_load_jff(self.h, str(path).encode()))

    def overlapNDTOccupancyScore(self, mov, T):
        """ndt_feature::overlapNDTOccupancyScore(ref=self, mov, T) (ndt_feature_node.h:213-252)."""
        s = C.c_double(0)
        Tc = _cm(T)
        self.e.check(self.e.L.ndtb_overlap_score(self.e.h, self.h, mov.h, Tc.ctypes.data, C.byref(s)))
        return s.value


class NDTMatcherD2D:
    """lslgeneric::NDTMatcherD2D with its public knobs (ndt_feature_graph.cpp:261-262)."""

    def __init__(self, engine, **knobs):
        self.e = engine
        self.params = engine.default_params(**knobs)

    @property
    def n_neighbours(self):
        return self.params.n_neighbours

    @n_neighbours.setter
    def n_neighbours(self, v):
        self.params.n_neighbours = int(v)

    def derivativesNDT(self, target, source, T, computeHessian=True):
        """score, gradient[6], Hessian[6,6], n_pairs of the source cells moved by T against the target map."""
        out = np.zeros(43)
        Tc = _cm(T)
        npairs = C.c_int64(0)
        self.e.check(self.e.L.ndtb_d2d_derivatives(self.e.h, target.h, source.h, Tc.ctypes.data, C.byref(self.params),
                                                   int(computeHessian), out.ctypes.data, C.byref(npairs)))
        return out[0], out[1:7].copy(), out[7:].reshape(6, 6).copy(), npairs.value

    def derivativesNDTCells(self, source_cells, target, computeHessian=True):
        """The cell-vector overload (ndt_matcher_d2d_fusion.h:856): source_cells = CELL_DTYPE records already moved."""
        cells = np.ascontiguousarray(source_cells, dtype=CELL_DTYPE)
        out = np.zeros(43)
        npairs = C.c_int64(0)
        self.e.check(self.e.L.ndtb_d2d_derivatives_cells(self.e.h, target.h, cells.ctypes.data, cells.shape[0], None,
                                                         C.byref(self.params), int(computeHessian), out.ctypes.data, C.byref(npairs)))
        return out[0], out[1:7].copy(), out[7:].reshape(6, 6).copy(), npairs.value

    def lineSearchMT(self, increment, source_cells, target):
        """NDTMatcherD2D::lineSearchMT (ndt_matcher_d2d_fusion.h:1013): returns (step, increment possibly negated)."""
        cells = np.ascontiguousarray(source_cells, dtype=CELL_DTYPE)
        inc = np.ascontiguousarray(increment, dtype=np.float64).copy()
        step = C.c_double(0)
        self.e.check(self.e.L.ndtb_d2d_line_search_cells(self.e.h, target.h, cells.ctypes.data, cells.shape[0], inc.ctypes.data,
                                                         C.byref(self.params), C.byref(step)))
        return step.value, inc

    def match(self, target, source, T, useInitialGuess=True):
        """Returns the Result (res.pose() is the refined T, res.converged the bool upstream returns)."""
        T0 = np.asarray(T, dtype=np.float64) if useInitialGuess else np.eye(4)
        r = Result()
        Tc = _cm(T0)
        self.e.check(self.e.L.ndtb_d2d_match(self.e.h, target.h, source.h, Tc.ctypes.data, C.byref(self.params), C.byref(r)))
        return r

    def matchFusion(self, target, source, T, Tcov):
        """ndt_feature::matchFusion with useNDT only (soft constraint / Tikhonov from self.params)."""
        r = Result()
        Tc = _cm(T)
        cov = np.ascontiguousarray(Tcov, dtype=np.float64)
        self.e.check(self.e.L.ndtb_fusion_match(self.e.h, target.h, source.h, Tc.ctypes.data, cov.ctypes.data,
                                                C.byref(self.params), C.byref(r)))
        return r

    def covariance(self, target, source, T):
        out = np.zeros(36)
        Tc = _cm(T)
        rc = self.e.L.ndtb_d2d_covariance(self.e.h, target.h, source.h, Tc.ctypes.data, C.byref(self.params), out.ctypes.data)
        if rc not in (0, -5):
            self.e.check(rc)
        return rc == 0, out.reshape(6, 6)


class NDTMatcherP2D:
    """lslgeneric::NDTMatcherP2D (no call site in the reference; BASELINE config C3): point cloud against an NDT map."""

    def __init__(self, engine, **knobs):
        self.e = engine
        self.params = engine.default_params(**knobs)

    def derivativesPointCloud(self, target, cloud, T, computeHessian=True):
        pts = _pts4(cloud)
        out = np.zeros(43)
        Tc = _cm(T)
        npairs = C.c_int64(0)
        self.e.check(self.e.L.ndtb_p2d_derivatives(self.e.h, target.h, pts.ctypes.data, pts.shape[0], HOST, Tc.ctypes.data,
                                                   C.byref(self.params), int(computeHessian), out.ctypes.data, C.byref(npairs)))
        return out[0], out[1:7].copy(), out[7:].reshape(6, 6).copy(), npairs.value

    def match(self, target, cloud, T, useInitialGuess=True):
        pts = _pts4(cloud)
        T0 = np.asarray(T, dtype=np.float64) if useInitialGuess else np.eye(4)
        r = Result()
        Tc = _cm(T0)
        self.e.check(self.e.L.ndtb_p2d_match(self.e.h, target.h, pts.ctypes.data, pts.shape[0], HOST, Tc.ctypes.data,
                                             C.byref(self.params), C.byref(r)))
        return r


class NDTMatcherD2D_2D(NDTMatcherD2D):
    """lslgeneric::NDTMatcherD2D_2D (matchFusion2d, ndt_matcher_d2d_fusion.h:1159-1176): D2D estimating (x, y, yaw) only."""

    def __init__(self, engine, **knobs):
        super().__init__(engine, planar=1, **knobs)
