"""Host mirror of the reference's front end over the C ABI: NDTFeatureFuserHMT and the NDTFeatureGraph node chain.

  NDTFeatureFuserHMT::initialize / update   ndt_feature/src/ndt_feature_src/ndt_feature_fuser_hmt.cpp:65-102, :108-512
  NDTFeatureGraph::initialize / update      ndt_feature/src/ndt_feature_src/ndt_feature_graph.cpp:24-55, :60-144

Everything computes on the GPU (ndtb_fuser_* / ndtb_graph_* in csrc/fuser.cu); this module only marshals arguments.
"""
import ctypes as C

import numpy as np

from . import api
from .api import HOST, FuserParams, NDTMap, Result, _cm, _pts4


def fuser_params(engine, sensor_pose=None, motion=None, **kw):
    """ndtb_fuser_params with the reference's defaults (NDTFeatureFuserHMT::Params()), overridden by keyword."""
    p = FuserParams()
    engine.L.ndtb_fuser_default_params(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    if sensor_pose is not None:
        p.sensor_pose = (C.c_double * 16)(*_cm(sensor_pose))
    if motion is not None:  # Cd, Ct, Dd, Dt, Td, Tt
        p.motion = (C.c_double * 6)(*[float(x) for x in motion])
    return p


def _pose(buf):
    return np.array(buf, dtype=np.float64).reshape(4, 4).T.copy()


class NDTFeatureFuserHMT:
    def __init__(self, engine, params):
        self.e, self.params = engine, params
        h = C.c_void_p()
        engine.check(engine.L.ndtb_fuser_create(engine.h, C.byref(params), C.byref(h)))
        self.h = h
        self.last_result = None
        self.last_cov = None

    def __del__(self):
        try:
            if getattr(self, "h", None) and getattr(self.e, "h", None):
                self.e.L.ndtb_fuser_destroy(self.h)
            self.h = None
        except Exception:
            pass

    def initialize(self, init_pose, cloud):
        pts = _pts4(cloud)
        Tc = _cm(init_pose)
        self.e.check(self.e.L.ndtb_fuser_initialize(self.h, Tc.ctypes.data, pts.ctypes.data, pts.shape[0], HOST))

    def update(self, Tmotion, cloud, update_ndt_map=True):
        pts = _pts4(cloud)
        Tc = _cm(Tmotion)
        out = np.zeros(16)
        cov = np.zeros(36)
        r = Result()
        self.e.check(self.e.L.ndtb_fuser_update(self.h, Tc.ctypes.data, pts.ctypes.data, pts.shape[0], HOST, int(update_ndt_map),
                                                out.ctypes.data, C.byref(r), cov.ctypes.data))
        self.last_result, self.last_cov = r, cov.reshape(6, 6)
        return _pose(out)

    @property
    def map(self):
        return NDTMap(self.e, borrowed_handle=self.e.L.ndtb_fuser_map(self.h))

    @property
    def Tnow(self):
        out = np.zeros(16)
        self.e.check(self.e.L.ndtb_fuser_pose(self.h, out.ctypes.data))
        return _pose(out)

    @Tnow.setter
    def Tnow(self, T):
        Tc = _cm(T)
        self.e.check(self.e.L.ndtb_fuser_set_pose(self.h, Tc.ctypes.data))


class _NodeView:
    """One node of the graph: .T, .Tlocal_odom, .Tlocal_fuse, .nbUpdates, .map.map (the node's NDT map)."""

    class _Fuser:
        def __init__(self, m):
            self.map = m

    def __init__(self, graph, k):
        e = graph.e
        T, To, Tf = np.zeros(16), np.zeros(16), np.zeros(16)
        mh, nb = C.c_void_p(), C.c_int32(0)
        e.check(e.L.ndtb_graph_node(graph.h, k, T.ctypes.data, To.ctypes.data, Tf.ctypes.data, C.byref(mh), C.byref(nb)))
        self.T, self.Tlocal_odom, self.Tlocal_fuse, self.nbUpdates = _pose(T), _pose(To), _pose(Tf), nb.value
        self.map = self._Fuser(NDTMap(e, borrowed_handle=mh.value))


class NDTFeatureGraph:
    def __init__(self, engine, params, new_node_transl_dist=1.0):
        self.e = engine
        h = C.c_void_p()
        engine.check(engine.L.ndtb_graph_create(engine.h, C.byref(params), float(new_node_transl_dist), C.byref(h)))
        self.h = h
        self._dist = float(new_node_transl_dist)

    def __del__(self):
        try:
            if getattr(self, "h", None) and getattr(self.e, "h", None):
                self.e.L.ndtb_graph_destroy(self.h)
            self.h = None
        except Exception:
            pass

    @property
    def new_node_transl_dist(self):
        return self._dist

    @new_node_transl_dist.setter
    def new_node_transl_dist(self, d):
        self._dist = float(d)
        self.e.check(self.e.L.ndtb_graph_set_new_node_dist(self.h, self._dist))

    def initialize(self, init_pose, cloud):
        pts = _pts4(cloud)
        Tc = _cm(init_pose)
        self.e.check(self.e.L.ndtb_graph_initialize(self.h, Tc.ctypes.data, pts.ctypes.data, pts.shape[0], HOST))

    def update(self, Tmotion, cloud):
        pts = _pts4(cloud)
        Tc = _cm(Tmotion)
        out = np.zeros(16)
        self.e.check(self.e.L.ndtb_graph_update(self.h, Tc.ctypes.data, pts.ctypes.data, pts.shape[0], HOST, out.ctypes.data))
        return _pose(out)

    @property
    def nodes(self):
        return [_NodeView(self, k) for k in range(int(self.e.L.ndtb_graph_num_nodes(self.h)))]


def make_graph(resolution, map_size, sensor_range, neighbours, itr_max, delta_score, soft, tikhonov, sensor_pose, motion,
               new_node_transl_dist, engine=None, device=0):
    """The configuration of scripts/replay_mapping.py (launch/henrik_replay_mapperbag_fuser.launch)."""
    e = engine or api.Engine(device)
    p = fuser_params(e, sensor_pose=sensor_pose, motion=motion, resolution=resolution, map_size_x=map_size[0],
                     map_size_y=map_size[1], map_size_z=map_size[2], sensor_range=sensor_range, neighbours=neighbours,
                     itr_max=itr_max, delta_score=delta_score, global_transf=0, use_soft_constraints=int(soft),
                     use_tikhonov=int(tikhonov), all_matches_valid=1)
    g = NDTFeatureGraph(e, p, new_node_transl_dist)
    g.engine = e
    return g
