"""Builds libspfresh_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m spfresh_b200.build [--force] [--verbose]

The library has no torch dependency: it is a plain C-ABI shared object (include/spfresh_b200.h)
with the CUDA runtime linked statically.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libspfresh_b200.so")
SOURCES = ["api.cu", "support.cu", "assign_exact.cu", "assign_tc.cu", "resolve.cu", "assign_api.cu",
           "ops.cu", "search.cu", "scan_tc.cu", "comm.cu", "kmeans.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
         "-Xcudafe", "--diag_suppress=177", "-Wno-deprecated-declarations"]


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(os.path.dirname(HERE), "include", "spfresh_b200.h"))
    return hs


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src, verbose):
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return src, r.returncode, r.stdout + r.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    hdrs = _headers()
    todo = [s for s in SOURCES
            if force or _stale(os.path.join(OBJ, s.replace(".cu", ".o")), [os.path.join(CSRC, s)] + hdrs)]
    if todo:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(todo))) as ex:
            for src, rc, out in ex.map(lambda s: _compile(s, verbose), todo):
                if verbose or rc != 0:
                    sys.stderr.write(f"--- {src}\n{out}\n")
                if rc != 0:
                    raise RuntimeError(f"nvcc failed on {src}")
    objs = [os.path.join(OBJ, s.replace(".cu", ".o")) for s in SOURCES]
    if todo or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-cudart", "static", "-lpthread", "-ldl", "-lrt"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
