"""In-tree build of the CUDA C-ABI library (sm_100a only) and of the C oracle.

`python -m unseenobjectclustering_b200.build` or __graft_entry__.build().  nvcc cross-compiles
without a GPU; the resulting libuoc_b200.so sits next to this file so that it travels with the repo
snapshot to the GPU box (it is git-ignored, not gpurun-ignored).
"""
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB_NAME = "libuoc_b200.so"
LIB_PATH = os.path.join(HERE, LIB_NAME)
SOURCES = ["uoc_runtime.cu", "cluster_kernels.cu", "fps_tc.cu", "meanshift_tc.cu", "assign_tc.cu", "cluster_api.cu", "conv_tc.cu", "conv_pair.cu", "conv_wres.cu",
           "backbone_kernels.cu", "backbone.cu", "input_prep.cu", "refine.cu", "metrics.cu"]
HEADERS = ["uoc_common.cuh", "cluster.cuh", "conv.cuh", os.path.join(ROOT, "include", "uoc.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo", "-Xcompiler",
              "-fPIC,-fvisibility=hidden", "--use_fast_math=false"]
NVCC_FLAGS = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")]

ORACLE_SRC = os.path.join(ROOT, "oracle", "uoc_oracle_c.c")
ORACLE_LIB = os.path.join(ROOT, "oracle", "libuoc_oracle.so")


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _digest(paths, extra=""):
    h = hashlib.sha256(extra.encode())
    for p in paths:
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def build_cuda(force=False, verbose=False):
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    hdrs = [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    stamp = os.path.join(HERE, ".libuoc_b200.sha256")
    digest = _digest(srcs + hdrs, " ".join(NVCC_FLAGS))
    if not force and os.path.exists(LIB_PATH) and os.path.exists(stamp) and open(stamp).read().strip() == digest:
        return LIB_PATH
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    objs = []
    for s in srcs:
        o = os.path.join(objdir, os.path.basename(s) + ".o")
        objs.append(o)
        cmd = [nvcc] + NVCC_FLAGS + ["-c", s, "-o", o]
        if verbose:
            print(" ".join(cmd))
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s" % (" ".join(cmd), out.decode(errors="replace")))
        elif verbose and out:
            print(out.decode(errors="replace"))
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + objs
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if r.returncode != 0:
        raise RuntimeError("link failed: %s\n%s" % (" ".join(cmd), r.stdout.decode(errors="replace")))
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB_PATH


def build_oracle(force=False):
    """gcc build of the C restatement used by the tests as the bit-exact checker."""
    if not os.path.exists(ORACLE_SRC):
        return None
    stamp = ORACLE_LIB + ".sha256"
    digest = _digest([ORACLE_SRC])
    if not force and os.path.exists(ORACLE_LIB) and os.path.exists(stamp) and open(stamp).read().strip() == digest:
        return ORACLE_LIB
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        raise RuntimeError("gcc not found")
    cmd = [cc, "-O2", "-fPIC", "-shared", "-std=c11", "-ffp-contract=off", "-fopenmp", "-o", ORACLE_LIB, ORACLE_SRC, "-lm"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed: %s\n%s" % (" ".join(cmd), r.stdout.decode(errors="replace")))
    with open(stamp, "w") as f:
        f.write(digest)
    return ORACLE_LIB


if __name__ == "__main__":
    force = "--force" in sys.argv
    print(build_cuda(force=force, verbose="-v" in sys.argv))
    print(build_oracle(force=force))
