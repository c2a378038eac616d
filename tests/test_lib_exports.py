"""The C-ABI library loads on a CPU-only box and exports every symbol declared in include/wedetect_b200.h."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


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
    assert ctypes.sizeof(_lib.WdOp) == 4 + 4 * 40 + 4 * 8 + 4 + 8 * 16  # kind, i[40], f[8], pad, p[16]
    lib = _lib.load(require_gpu=False)
    assert lib.wd_version() == 100
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
