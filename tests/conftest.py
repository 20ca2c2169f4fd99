import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "full_graph.npz"))


@pytest.fixture(scope="session")
def oracle():
    import oracle_py

    oracle_py.lib()
    return oracle_py


@pytest.fixture(scope="session")
def engine():
    """The CUDA engine through the C ABI.  No fallback: a missing library / device is a hard error."""
    import ndt_feature_graph_b200 as N

    e = N.Engine(0)
    yield e
    e.close()


def fixture_cells(golden, k, dtype):
    """Gaussian cells of shipped map k (ndt_feature/data/FULL GRAPH/mapping{k}.jff) as a structured array + grid."""
    hdr = golden[f"hdr{k}"]
    size_m, cell, center = hdr[:3], hdr[3:6], hdr[6:9]
    size = np.abs(np.ceil(size_m / cell)).astype(int)
    n = golden[f"mean{k}"].shape[0]
    cells = np.zeros(n, dtype)
    cells["mean"] = golden[f"mean{k}"]
    cells["cov"] = golden[f"cov{k}"]
    cells["n"] = golden[f"n{k}"]
    cells["has_gaussian"] = 1
    gi = golden[f"gidx{k}"]
    cells["idx"] = np.stack([gi // (size[1] * size[2]), (gi // size[2]) % size[1], gi % size[2]], 1)
    return center, cell, size, cells


@pytest.fixture(scope="session")
def oracle_fixture_maps(golden, oracle):
    maps = []
    for k in range(8):
        center, cell, size, cells = fixture_cells(golden, k, oracle.CELL_DTYPE)
        maps.append(oracle.OracleMap(0.5).from_cells(center, cell, size, cells, use_idx=True))
    return maps


@pytest.fixture(scope="session")
def gpu_fixture_maps(golden, engine):
    import ndt_feature_graph_b200 as N

    maps = []
    for k in range(8):
        center, cell, size, cells = fixture_cells(golden, k, N.CELL_DTYPE)
        maps.append(N.NDTMap(engine, 0.5).from_cells(center, cell, size, cells, use_idx=True))
    return maps
