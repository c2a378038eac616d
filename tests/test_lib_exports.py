"""The C-ABI library loads on a CPU-only box and exports every symbol declared in include/wedetect_b200.h."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "wedetect_b200.h")


def _declared():
    src = open(os.path.join(ROOT, "include", "wedetect_b200.h")).read()
    body = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(wd_[a-z0-9_]+)\s*\(", body)))


def test_header_symbols_are_exported():
    from wedetect_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert "wd_op_run" in names and "wd_program_create" in names and len(names) >= 12
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(_lib.EXPORTS) == names


def test_struct_layout_matches_header():
    from wedetect_b200 import _lib
    hdr = open(HEADER).read()
    ni, nf, np_ = (int(re.search(rf"#define {k} (\d+)", hdr).group(1)) for k in ("WD_OP_NI", "WD_OP_NF", "WD_OP_NP"))
    assert (ni, nf, np_) == (_lib.WD_OP_NI, _lib.WD_OP_NF, _lib.WD_OP_NP)
    assert ctypes.sizeof(_lib.WdOp) == (4 + 4 * ni + 4 * nf + 7) // 8 * 8 + 8 * np_  # kind, i[], f[], pad, p[]
    lib = _lib.load(require_gpu=False)
    assert lib.wd_version() == 200
    # the power of two activations are stored at (fp16 hi/lo planes) is compiled into the library; the host mirrors it
    assert float(re.search(r"#define WD_ACT_PLANE_SCALE ([0-9.]+)f", hdr).group(1)) == _lib.ACT_PLANE_SCALE == lib.wd_act_plane_scale()
    assert lib.wd_pp_workspace_bytes(2, 8400, 80, 30000) > 2 * 8400 * 80 * 16


def test_product_fails_loudly_without_gpu():
    import torch
    from wedetect_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.WdError):
        _lib.load(require_gpu=True)
    from wedetect_b200.detector import YOLOWorldDetector
    with pytest.raises(_lib.WdError):
        YOLOWorldDetector(size="tiny")


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under wedetect_b200/ may import it."""
    pkg = os.path.join(ROOT, "wedetect_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn
