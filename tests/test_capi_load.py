"""CPU: the C-ABI library loads, exports every symbol include/xm_b200.h declares, and refuses to run without a GPU
(no CPU fallback).  No compute calls here."""
import ctypes
import os
import re

import pytest

from xm_code_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "xm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(xm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = capi.load()
    names = declared_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"libxm_b200.so does not export {n}"
    for n in capi.EXPORTS:
        assert hasattr(lib, n)


def test_struct_layouts_match_header():
    assert ctypes.sizeof(capi.XmOptions) == 11 * 4
    assert ctypes.sizeof(capi.XmLogRec) == 4 * 4 + 3 * 8
    assert ctypes.sizeof(capi.XmCertInfo) == 4 * 4 + 5 * 8 and ctypes.sizeof(capi.XmSolveResult) == 8 * 4 + 7 * 8
    assert ctypes.sizeof(capi.XmStats) == 5 * 4 + 4 + 5 * 8 + 4 * 4 + 4 * 8   # 5 ints + pad, 5 doubles, 4 ints, 4 doubles


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.XmError):
        capi.Handle(device=0)


def test_product_does_not_import_oracle():
    # the shipped package must never route through the oracle
    for dirpath, _, files in os.walk(os.path.join(ROOT, "xm_code_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "xm_oracle" not in txt, f
    txt = open(os.path.join(ROOT, "XM", "src", "XM_main.cpp")).read()
    assert "oracle" not in txt
