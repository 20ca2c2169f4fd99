"""CPU suite, part 3: the C-ABI library loads without a GPU and exports every symbol include/ndtb.h declares;
compute entry points fail loudly (no CPU fallback); struct layouts of the Python mirror match the header."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "ndtb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ndtb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from ndt_feature_graph_b200 import api

    L = api.load_library()
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/ndtb.h but not exported by libndtb.so"
    assert set(api.abi_symbols()) == set(names)  # the Python mirror binds exactly the declared ABI
    assert L.ndtb_version() == 200


def test_struct_layouts():
    from ndt_feature_graph_b200 import api

    assert C.sizeof(api.Params) == 64
    assert C.sizeof(api.Result) == 192 == api.RESULT_DTYPE.itemsize
    assert api.CELL_DTYPE.itemsize == 96
    for f, _ in api.Result._fields_:
        if f != "T":
            assert getattr(api.Result, f).offset == api.RESULT_DTYPE.fields[f][1]


def test_no_cpu_fallback():
    """Without a CUDA device the engine must refuse to compute (and it must never import the oracle)."""
    import torch

    import ndt_feature_graph_b200 as N

    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the -m gpu suite")
    with pytest.raises(N.NdtbError):
        N.Engine(0)
    src = ""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "ndt_feature_graph_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src += open(os.path.join(dirpath, f)).read()
    assert "oracle_py" not in src and "ndt_oracle" not in src and "libndt_oracle" not in src


def test_graft_entry_build_products():
    assert os.path.exists(os.path.join(ROOT, "ndt_feature_graph_b200", "lib", "libndtb.so"))
    import __graft_entry__ as g

    assert callable(g.build) and callable(g.smoke)
