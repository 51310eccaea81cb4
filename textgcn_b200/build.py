"""Build libtgcn_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
INCLUDE = os.path.join(ROOT, "include")
LIB_PATH = os.path.join(PKG_DIR, "libtgcn_b200.so")
OBJ_DIR = os.path.join(PKG_DIR, "csrc", "_obj")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-I", INCLUDE, "-I", CSRC,
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


DEBUG_LIB_PATH = os.path.join(PKG_DIR, "libtgcn_b200_dbg.so")


def build(force: bool = False, verbose: bool = False, debug: bool = False) -> str:
    """Compile every csrc/*.cu to an object and link the shared library.  Returns its path.
    ``debug``: the debug-assert flavour (-DTGCN_DEBUG_BOUNDS: every index a kernel dereferences is range-checked with a device
    assert) as libtgcn_b200_dbg.so; load it with TGCN_B200_LIB=<path> (tests / tools only)."""
    sources = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    headers = sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + sorted(glob.glob(os.path.join(INCLUDE, "*.h")))
    obj_dir = OBJ_DIR + ("_dbg" if debug else "")
    lib_path = DEBUG_LIB_PATH if debug else LIB_PATH
    flags = NVCC_FLAGS + (["-DTGCN_DEBUG_BOUNDS"] if debug else [])
    os.makedirs(obj_dir, exist_ok=True)
    nvcc = _nvcc()
    objs, jobs = [], []
    for src in sources:
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + headers):
            jobs.append([nvcc, *flags, "-c", src, "-o", obj])

    def run(cmd):
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed: {' '.join(cmd)}\n{res.stdout}\n{res.stderr}")

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as pool:
        list(pool.map(run, jobs))
    if force or jobs or _stale(lib_path, objs):
        run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", lib_path, *objs, "-ldl"])
    return lib_path


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True, debug="--debug" in sys.argv))
