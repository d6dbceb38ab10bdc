"""Build ``libgnngp_b200.so`` in-tree with nvcc for sm_100a (no other architecture, no JIT cache).

    python -m gnngp_b200.build            # or: from gnngp_b200.build import build; build()

Each ``csrc/*.cu`` is compiled to an object under ``build/`` (in parallel) and linked, with the CUDA
runtime linked statically, into ``gnn-gp_b200/lib/libgnngp_b200.so``.  The shared object is
git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import concurrent.futures
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libgnngp_b200.so")
OBJ_DIR = os.path.join(ROOT, "build", "obj")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _nvcc() -> str:
    path = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(path):
        raise RuntimeError("nvcc not found; the CUDA path cannot be built")
    return path


def _newest_header() -> float:
    deps = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(ROOT, "include", "*.h"))
    return max(os.path.getmtime(p) for p in deps)


def _compile(src: str, force: bool) -> str:
    obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
    if (not force and os.path.exists(obj) and os.path.getmtime(obj) >= os.path.getmtime(src)
            and os.path.getmtime(obj) >= _newest_header()):
        return obj
    cmd = [_nvcc()] + NVCC_FLAGS + ["-c", src, "-o", obj]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    log = obj[:-2] + ".ptxas.log"
    with open(log, "w") as fh:
        fh.write(proc.stdout + proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s" % (src, proc.stdout + proc.stderr))
    return obj


def build(force: bool = False, verbose: bool = True) -> str:
    os.makedirs(OBJ_DIR, exist_ok=True)
    os.makedirs(LIB_DIR, exist_ok=True)
    sources = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(sources))) as pool:
        objs = list(pool.map(lambda s: _compile(s, force), sources))
    if (force or not os.path.exists(LIB_PATH)
            or any(os.path.getmtime(o) > os.path.getmtime(LIB_PATH) for o in objs)):
        cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + objs
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError("link failed:\n" + proc.stdout + proc.stderr)
    if verbose:
        print("built %s (%d bytes)" % (LIB_PATH, os.path.getsize(LIB_PATH)))
    return LIB_PATH


if __name__ == "__main__":
    build(force="--force" in sys.argv)
