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


def test_integration_stub_matches_the_header():
    """The reference-side ctypes stub printed in INTEGRATION.md section 2 is executable as written (against the built
    library) and its argtypes have exactly the parameters include/uoc.h declares for the entry point it binds."""
    from unseenobjectclustering_b200 import _lib, build
    build.build_cuda()
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = re.search(r"```python\n(# lib/utils/mean_shift_b200\.py.*?)```", doc, flags=re.S).group(1)
    assert 'ctypes.CDLL("libuoc_b200.so")' in block
    ns = {}
    exec(compile(block.replace('ctypes.CDLL("libuoc_b200.so")', "ctypes.CDLL(%r)" % _lib.LIB_PATH), "INTEGRATION.md", "exec"), ns)
    header = open(os.path.join(ROOT, "include", "uoc.h")).read()
    decl = re.search(r"UOC_API\s+int\s+uoc_meanshift_cluster\s*\((.*?)\)\s*;", header, flags=re.S).group(1)
    n_params = len([p for p in decl.split(",") if p.strip()])
    assert len(ns["_lib"].uoc_meanshift_cluster.argtypes) == n_params == len(_lib.SIGNATURES["uoc_meanshift_cluster"][1])
    assert callable(ns["clustering_features"])


def test_binding_table_has_the_headers_parameter_counts_and_kinds():
    """Every entry of the ctypes table (_lib.SIGNATURES) has as many parameters as its declaration in include/uoc.h, and
    pointer / integer / float kinds agree position by position (a mismatch would corrupt the call frame silently)."""
    from unseenobjectclustering_b200 import _lib
    header = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "uoc.h")).read(), flags=re.S)
    decls = dict(re.findall(r"UOC_API\s+[^;(]*?\b(uoc_[a-z0-9_]+)\s*\((.*?)\)\s*;", header, flags=re.S))
    assert sorted(decls) == sorted(_lib.SIGNATURES)

    def kind_of_c(param):
        p = " ".join(param.split())
        if "*" in p or "uoc_stream_t" in p:
            return "ptr"
        if re.search(r"\b(float|double)\b", p):
            return "float"
        if re.search(r"\b(int64_t|uint64_t|size_t|long long)\b", p):
            return "int64"
        return "int"

    def kind_of_ctypes(t):
        if t in (ctypes.c_float, ctypes.c_double):
            return "float"
        if t in (ctypes.c_void_p, ctypes.c_char_p) or hasattr(t, "contents") or getattr(t, "_type_", None) == "P":
            return "ptr"
        if ctypes.sizeof(t) == 8:
            return "int64"
        return "int"

    for name, params in decls.items():
        plist = [p for p in params.split(",") if p.strip() and p.strip() != "void"]
        args = _lib.SIGNATURES[name][1]
        assert len(plist) == len(args), (name, len(plist), len(args))
        for i, (p, a) in enumerate(zip(plist, args)):
            assert kind_of_c(p) == kind_of_ctypes(a), (name, i, p.strip(), a)
