"""The C-ABI library loads on a CPU box and exports every symbol include/t21ps.h declares
(no compute calls without a GPU), and the Python side fails loudly without CUDA."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from tools21cm_b200 import _lib


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "t21ps.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(t21_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    for s in ("t21_plan_create", "t21_plan_destroy", "t21_workspace_bytes", "t21_ps_binned", "t21_ps_nd",
              "t21_bin_only", "t21_last_error", "t21_launch_count"):
        assert s in syms


@pytest.mark.skipif(not os.path.exists(_lib.LIB_PATH), reason="libt21ps.so not built (run __graft_entry__.build())")
def test_library_exports_every_declared_symbol():
    lib = C.CDLL(_lib.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(lib, s), "missing export %s" % s
    assert _lib.load().t21_version() >= 100


def test_binspec_struct_matches_header_layout():
    # field order of the ctypes mirror must follow the C struct
    text = open(os.path.join(ROOT, "include", "t21ps.h")).read()
    body = text[text.index("typedef struct t21_binspec {"):text.index("} t21_binspec;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = re.findall(r"\b([a-z_0-9]+)(?:\[\d+\])?;", body)
    assert names == [f[0] for f in _lib.BinSpecC._fields_]


def test_no_cpu_fallback_without_cuda():
    import torch
    import tools21cm_b200 as t2c
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError):
        t2c.power_spectrum_1d(np.zeros((16, 16, 16), dtype=np.float32), kbins=4, box_dims=10.0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "tools21cm_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src, (dirpath, f)
