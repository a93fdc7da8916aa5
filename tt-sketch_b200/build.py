#!/usr/bin/env python
"""Build libttsk.so (hand-written sm_100a CUDA + C ABI) in-tree with nvcc.

    python tt-sketch_b200/build.py [--force]

The shared library has no Python/torch dependency: it links only the CUDA runtime
(statically) and is loaded with ctypes by tt_sketch/_backend.py.
"""
import concurrent.futures
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libttsk.so")
OBJ = os.path.join(HERE, "build")
SOURCES = ["ttsk_sparse_pass_x0.cu", "ttsk_sparse_pass_x1.cu", "ttsk_sparse_gen.cu", "ttsk_sparse_gather.cu", "ttsk_api.cu", "ttsk_gauss.cu", "ttsk_gemm.cu", "ttsk_sparse.cu",
           "ttsk_linalg.cu", "ttsk_tt.cu", "ttsk_dense.cu", "ttsk_tns.cu"]
NVCC = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v"]
if os.environ.get("TTSK_PROFILE"):  # profiling build: enables the TTSK_ABLATE / TTSK_L2_FETCH switches (never shipped)
    FLAGS.append("-DTTSK_PROFILE")


def _newer(target, deps):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".inc", ".h"))]
    hs.append(os.path.join(os.path.dirname(HERE), "include", "ttsk.h"))
    return hs


def _compile(src):
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    deps = [os.path.join(CSRC, src)] + _headers()
    if _newer(obj, deps):
        return obj, ""
    cmd = [NVCC] + FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, p.stdout, p.stderr))
    with open(obj + ".ptxas.log", "w") as f:
        f.write(p.stderr)
    return obj, p.stderr


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    with concurrent.futures.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        results = list(ex.map(_compile, SOURCES))
    objs = [r[0] for r in results]
    if not _newer(OUT, objs):
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT] + objs + ["-cudart", "static"]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (p.stdout, p.stderr))
    if verbose:
        for _, log in results:
            sys.stderr.write(log)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
