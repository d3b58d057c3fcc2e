"""In-tree build of libmvster_b200.so (nvcc, sm_100a only).

``build_library()`` compiles every ``csrc/*.cu`` translation unit with
``nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo`` (in parallel) and links them into
``libmvster_b200.so`` next to this file.  The CUDA runtime is linked statically so the library does not depend on
which libcudart the host process (torch) has loaded.  Nothing here needs a GPU: nvcc cross-compiles.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
BUILD_DIR = os.path.join(PKG_DIR, "build")
LIB_PATH = os.path.join(PKG_DIR, "libmvster_b200.so")
INCLUDE = os.path.join(os.path.dirname(PKG_DIR), "include")

# no --use_fast_math: parity with the reference needs IEEE division / sqrt / expf
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: cannot build libmvster_b200.so")
    return exe


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest(path: str) -> str:
    h = hashlib.sha1()
    headers = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh"))
    for dep in [path] + headers + [os.path.join(INCLUDE, "mvster_b200.h")]:
        with open(dep, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile_one(src: str, verbose: bool) -> str:
    obj = os.path.join(BUILD_DIR, src[:-3] + ".o")
    stamp = obj + ".sha1"
    digest = _digest(os.path.join(CSRC, src))
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == digest:
        return obj
    cmd = [_nvcc()] + NVCC_FLAGS + ["-I", INCLUDE, "-c", os.path.join(CSRC, src), "-o", obj]
    if verbose:
        print("[mvster build]", " ".join(cmd), flush=True)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, res.stdout, res.stderr))
    with open(stamp, "w") as f:
        f.write(digest)
    return obj


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile (if stale) and link ``libmvster_b200.so``; returns its path."""
    os.makedirs(BUILD_DIR, exist_ok=True)
    if force:
        for f in os.listdir(BUILD_DIR):
            os.remove(os.path.join(BUILD_DIR, f))
    srcs = _sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile_one(s, verbose), srcs))
    newest = max(os.path.getmtime(o) for o in objs)
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < newest:
        cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + objs + \
              ["-cudart", "static"]
        if verbose:
            print("[mvster build]", " ".join(cmd), flush=True)
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (res.stdout, res.stderr))
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
