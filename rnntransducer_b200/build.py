"""In-tree build of librnnt_b200.so (hand-written sm_100a kernels + the C ABI).

``python -m rnntransducer_b200.build`` or ``__graft_entry__.build()``.  nvcc cross-compiles
without a GPU; the resulting ``.so`` sits next to this file so it travels with the tree.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "librnnt_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",  # plain -arch=sm_100a drops the 'a' features here
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-shared",
    "-Xptxas", "-v",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    mt = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + [
        os.path.join(os.path.dirname(HERE), "include", "rnnt_b200.h"), __file__]
    return any(os.path.getmtime(d) > mt for d in deps)


def _compile_one(nvcc, src, obj, hdr_mtime):
    """One translation unit -> object file (skipped when the object is newer than source + headers)."""
    if os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), hdr_mtime):
        return src, 0, ""
    cmd = [nvcc] + [f for f in NVCC_FLAGS if f != "-shared"] + ["-c", "-o", obj, src]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    return src, proc.returncode, " ".join(cmd) + "\n" + proc.stdout + proc.stderr


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: librnnt_b200.so must be built where the CUDA toolkit is")
    from concurrent.futures import ThreadPoolExecutor
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    hdrs = glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(os.path.dirname(HERE), "include", "rnnt_b200.h"), __file__]
    hdr_mtime = float("inf") if force else max(os.path.getmtime(h) for h in hdrs)
    objs = [os.path.join(objdir, os.path.basename(s)[:-3] + ".o") for s in sources()]
    with ThreadPoolExecutor(max_workers=min(len(objs), os.cpu_count() or 1)) as ex:
        results = list(ex.map(lambda so: _compile_one(nvcc, so[0], so[1], hdr_mtime), zip(sources(), objs)))
    log = "".join(r[2] for r in results)
    tmp = LIB + ".tmp"
    failed = [r[0] for r in results if r[1] != 0]
    if not failed:
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC", "-o", tmp] + objs
        proc = subprocess.run(cmd, capture_output=True, text=True)
        log += " ".join(cmd) + "\n" + proc.stdout + proc.stderr
        if proc.returncode != 0:
            failed = ["link"]
    with open(os.path.join(HERE, "build.log"), "a" if not force else "w") as f:
        f.write(log)
    if failed:
        sys.stderr.write(log)
        raise RuntimeError(f"nvcc failed building librnnt_b200.so ({failed})")
    os.replace(tmp, LIB)
    if verbose:
        print(log)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
