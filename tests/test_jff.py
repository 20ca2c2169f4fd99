"""JFF map files (NDTMap::writeToJFF / loadFromJFF [upstream], ndt_feature_fuser_hmt.cpp:15,24,39): host-side format
code checked against the golden extraction of the maps the reference ships, byte for byte where the reference checkout
is available (authoring container only), and round trips elsewhere."""
import os

import numpy as np
import pytest

from ndt_feature_graph_b200 import api

from conftest import fixture_cells

REF = "/root/reference/ndt_feature/data/FULL GRAPH"


def _golden_cells(golden, k):
    center, cell, size, cells = fixture_cells(golden, k, api.CELL_DTYPE)
    # occupancy of every informative cell (the golden file keeps it separately)
    lin = (cells["idx"][:, 0].astype(np.int64) * size[1] + cells["idx"][:, 1]) * size[2] + cells["idx"][:, 2]
    occ = dict(zip(golden[f"occidx{k}"].tolist(), golden[f"occ{k}"].tolist()))
    cells["occ"] = [occ.get(int(i), 0.0) for i in lin]
    return center, cell, size, cells


def test_round_trip_host_only(golden, tmp_path):
    for k in (0, 5):
        center, cell, size, cells = _golden_cells(golden, k)
        path = tmp_path / f"m{k}.jff"
        api.jff_write_cells(path, center, cell, size, cells)
        assert os.path.getsize(path) == 10 + 4 + 72 + 480 + 181 * int(np.prod(size))
        c2, s2, n2, back = api.jff_read_cells(path)
        assert np.array_equal(c2, center) and np.array_equal(s2, cell) and np.array_equal(n2, size)
        order = np.lexsort((cells["idx"][:, 2], cells["idx"][:, 1], cells["idx"][:, 0]))
        for f in ("idx", "mean", "cov", "n", "has_gaussian", "occ"):
            assert np.array_equal(back[f], cells[f][order]), f
    with pytest.raises(api.NdtbError):
        api.jff_read_cells(tmp_path / "missing.jff")


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout (shipped .jff maps) only exists in the authoring container")
def test_reader_and_writer_against_the_shipped_maps(golden, tmp_path):
    for k in range(8):
        src = os.path.join(REF, f"mapping{k}.jff")
        center, cell, size, cells = api.jff_read_cells(src)
        hdr = golden[f"hdr{k}"]
        assert np.array_equal(center, hdr[6:9]) and np.array_equal(cell, hdr[3:6])
        assert np.array_equal(size, np.abs(np.ceil(hdr[:3] / hdr[3:6])).astype(int))
        g = cells[cells["has_gaussian"] == 1]
        lin = (g["idx"][:, 0].astype(np.int64) * size[1] + g["idx"][:, 1]) * size[2] + g["idx"][:, 2]
        assert np.array_equal(lin, golden[f"gidx{k}"])
        assert np.array_equal(g["mean"], golden[f"mean{k}"]) and np.array_equal(g["cov"], golden[f"cov{k}"])
        assert np.array_equal(g["n"], golden[f"n{k}"])
        # re-write and compare with the original file: every byte the format code models, in every record
        out = tmp_path / f"rw{k}.jff"
        api.jff_write_cells(out, center, cell, size, cells)
        a, b = open(src, "rb").read(), open(out, "rb").read()
        assert len(a) == len(b) and a[:86] == b[:86]
        ra = np.frombuffer(a[566:], np.uint8).reshape(-1, 181)
        rb = np.frombuffer(b[566:], np.uint8).reshape(-1, 181)
        has = ra[:, 136:140].copy().view("<i4").ravel() == 1
        modelled = np.r_[0:40, 128:181]          # centre, cell size, N, flags, occupancy, colour: all 80 000 records
        assert np.array_equal(ra[:, modelled], rb[:, modelled])
        # covariance, mean and the two doubles after them: defined where hasGaussian_ (uninitialised memory elsewhere)
        assert has.sum() == golden[f"gidx{k}"].size and np.array_equal(ra[has][:, 40:128], rb[has][:, 40:128])


@pytest.mark.gpu
def test_map_write_and_load_jff(engine, golden, gpu_fixture_maps, tmp_path):
    import ndt_feature_graph_b200 as N

    m = gpu_fixture_maps[2]
    path = tmp_path / "node2.jff"
    assert m.writeToJFF(path) == 0
    back = N.NDTMap(engine, 0.5)
    assert back.loadFromJFF(path) == 0
    a, b = m.export_cells(False), back.export_cells(False)
    for f in ("idx", "mean", "cov", "n", "has_gaussian", "occ"):
        assert np.array_equal(a[f], b[f]), f
    T = golden["Todom2"]
    d = N.NDTMatcherD2D(engine)
    r1 = d.match(m, gpu_fixture_maps[3], T)
    r2 = d.match(back, gpu_fixture_maps[3], T)
    assert list(r1.T) == list(r2.T)
    assert back.loadFromJFF(tmp_path / "missing.jff") != 0
