"""ndt_feature_graph_b200 — B200-native NDT registration engine.

The product is the CUDA library ``lib/libndtb.so`` (sources in ``csrc/``, C ABI in ``include/ndtb.h``).
This package is the thin Python host mirror of the reference's C++ interface for the hot path
(``lslgeneric::NDTMap`` / ``LazyGrid`` / ``NDTMatcherD2D``, reached from
``ndt_feature/src/ndt_feature_src/ndt_feature_graph.cpp:260-345`` and ``ndt_feature_fuser_hmt.cpp:108-512``)
used by the parity tests and bench.py.  There is no CPU fallback: without the built library or
without a CUDA device every compute call raises.
"""
from .api import (  # noqa: F401
    Engine,
    LazyGrid,
    NDTMap,
    NDTMatcherD2D,
    NDTMatcherD2D_2D,
    NDTMatcherP2D,
    NdtbError,
    Params,
    Result,
    CELL_DTYPE,
    lib_path,
)
