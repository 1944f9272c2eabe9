"""The C-ABI library loads on a CPU-only box and exports every symbol include/uoc.h declares."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "uoc.h")).read()
    return sorted(set(re.findall(r"UOC_API\s+[^;(]*?\b(uoc_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    syms = _declared_symbols()
    for must in ("uoc_meanshift_cluster", "uoc_backbone_forward", "uoc_backbone_create", "uoc_select_seeds",
                 "uoc_hill_climb", "uoc_label_seeds", "uoc_assign_labels", "uoc_last_error"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    from unseenobjectclustering_b200 import _lib, build
    build.build_cuda()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in _declared_symbols():
        assert hasattr(lib, s), "symbol %s declared in include/uoc.h is not exported" % s
    # the Python binding table covers the header exactly
    assert sorted(_lib.SIGNATURES.keys()) == _declared_symbols()


def test_no_compute_without_gpu_fails_loudly():
    """Without a CUDA device the product path must raise, never fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from unseenobjectclustering_b200 import _lib, mean_shift
    lib = _lib.load()
    assert lib.uoc_version() >= 100
    assert lib.uoc_device_info(None, None, None) != 0
    assert _lib.last_error() != ""
    with pytest.raises(_lib.UocError):
        mean_shift.cluster_fields(torch.zeros(1, 64, 8, 8))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "unseenobjectclustering_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                if f == "build.py":
                    continue   # build() compiles the checker; it never calls it
                bad = re.findall(r"^\s*(?:import|from)\s+\S*(?:uoc_oracle|ref_harness|oracle)\b|#include[^\n]*oracle|"
                                 r"CDLL\([^\n]*oracle", src, flags=re.M)
                assert not bad, (f, bad)
