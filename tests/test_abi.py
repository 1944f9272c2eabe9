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


C_CALLER = r'''
#include <stdio.h>
#include <string.h>
#include "uoc.h"
/* a plain C99 caller of the boundary: sizes, one compute call, the error text */
int main(void) {
  float x[2 * 8];
  int64_t first = 0, selected[2];
  int32_t labels[8];
  size_t ws = uoc_meanshift_workspace_bytes(1, 8, 2, 2);
  int rc;
  memset(x, 0, sizeof(x));
  if (ws == 0 || uoc_meanshift_workspace_bytes(0, 8, 2, 2) != 0) return 10;
  if (uoc_version() < 100) return 11;
  rc = uoc_meanshift_cluster(x, 16, 8, NULL, 1, 8, 2, 2, 20.0f, 10, 0.04f, &first, labels, selected, NULL, NULL,
                             (void*)x, ws, 0, NULL);
  printf("rc=%d err=%s\n", rc, uoc_last_error());
  return rc == UOC_OK ? 12 : 0;          /* host pointers / no device: the call must refuse, not compute */
}
'''


def test_header_is_plain_c_and_a_c_program_links_against_the_library(tmp_path):
    """include/uoc.h is the drop-in boundary: it must compile as C99 (no C++-isms, no torch / CUDA headers) and a C program
    must link against the shared library alone.  On a box without a GPU the compute call returns an error code with a text."""
    import subprocess
    import torch
    from unseenobjectclustering_b200 import _lib, build
    build.build_cuda()
    src = tmp_path / "caller.c"
    src.write_text(C_CALLER)
    exe = str(tmp_path / "caller")
    libdir = os.path.dirname(_lib.LIB_PATH)
    libname = os.path.basename(_lib.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), str(src),
                    "-o", exe, "-L", libdir, "-l:" + libname, "-Wl,-rpath," + libdir], check=True, capture_output=True, text=True)
    if torch.cuda.is_available():
        return                                            # the compute half of this test is for the CPU-only box
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "rc=" in r.stdout and "err=" in r.stdout and not r.stdout.strip().endswith("err=")
