"""CPU suite: the C-ABI library loads and exports every symbol include/gbrl_b200.h declares; host-side
argument handling of the GBRL mirror; no compute calls (there is no GPU here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "gbrl_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(gbrl_b200_[a-z_0-9]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from gbrl_b200 import _capi
    assert os.path.exists(_capi.LIB_PATH), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = C.CDLL(_capi.LIB_PATH)
    declared = _header_symbols()
    assert len(declared) >= 25
    for s in declared:
        assert hasattr(lib, s), "missing export " + s
    assert sorted(_capi.EXPORTS) == declared, "gbrl_b200/_capi.py EXPORTS out of sync with the header"


def test_ctypes_struct_layout_matches_header():
    from gbrl_b200 import _capi
    assert C.sizeof(_capi.Config) == 19 * 4
    assert C.sizeof(_capi.Metadata) == 17 * 4 + 4 + 5 * 8 + 8 + 6 * 8   # 4 bytes padding before the int64 block, float + padding, 6 x int64


def test_no_cpu_fallback():
    from gbrl_b200 import GBRL
    with pytest.raises(RuntimeError):
        GBRL(input_dim=4, output_dim=1, policy_dim=1, device="cpu")


def test_create_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from gbrl_b200 import GBRL
    with pytest.raises(RuntimeError, match="CUDA|fallback"):
        GBRL(input_dim=4, output_dim=1, policy_dim=1, device="cuda")


def test_argument_forms():
    from gbrl_b200.gbrl_cpp import _Arg
    a = np.zeros((5, 3), np.float32)
    x = _Arg(a, "obs", "step", True)
    assert x.shape == (5, 3) and x.dev == 0 and x.ptr == a.ctypes.data
    t = _Arg((1234, (7, 2), "torch.float32", "cuda"), "obs", "step", True)
    assert t.ptr == 1234 and t.shape == (7, 2) and t.dev == 1
    assert _Arg(None, "obs", "step", True).ptr is None
    with pytest.raises(RuntimeError):
        _Arg(None, "grads", "step", False)
    with pytest.raises(RuntimeError):
        _Arg(np.zeros((5, 3), np.float64), "obs", "step", True)
    with pytest.raises(RuntimeError):
        _Arg(np.zeros((5, 6), np.float32)[:, ::2], "obs", "step", True)
    with pytest.raises(RuntimeError):
        _Arg((1, (2, 2), "torch.float64", "cpu"), "obs", "step", True)


def test_product_path_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under gbrl_b200/ may reference it."""
    pkg = os.path.join(ROOT, "gbrl_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, fn), errors="replace").read()
                assert "oracle" not in txt.lower(), "%s mentions the oracle" % fn


def test_reference_side_cpp_backend_compiles_and_links():
    """INTEGRATION.md section 2 is code: integration/b200_backend.h (written against the reference's types.h) was compiled and
    linked against libgbrl_b200.so by `make -C oracle backend_check`; the resulting object loads and exports its probe."""
    path = os.path.join(ROOT, "oracle", "_ref", "libb200_backend_check.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libb200_backend_check.so not built (reference headers not mounted)")
    from gbrl_b200 import _capi
    _capi.lib()
    L = C.CDLL(path)
    assert hasattr(L, "gbrl_b200_backend_check")
