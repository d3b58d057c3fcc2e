#!/usr/bin/env python
"""Build an experimental variant of libmvster_b200.so with extra -D macros (same ABI), for A/B timing on the GPU box.

    python scripts/build_variant.py NAME -DMVSTER_TMA_MINB=5 ...   ->  variants/libvar_NAME.so
    MVSTER_B200_LIB=variants/libvar_NAME.so python scripts/bench_extra.py --which stages
"""
import os, subprocess, sys
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deep_reconstruction_with_epipolar_lines_mvster_b200 import _build as B

name, defs = sys.argv[1], sys.argv[2:]
out_dir = os.path.join(ROOT, "variants", "var_" + name)
os.makedirs(out_dir, exist_ok=True)
ONLY = tuple(os.environ.get("MVSTER_VARIANT_UNITS", "epi_fwd.cu").split(","))  # units the macros touch; the rest is reused


def cc(src):
    obj = os.path.join(out_dir, src[:-3] + ".o")
    cmd = [B._nvcc()] + B.NVCC_FLAGS + defs + ["-I", B.INCLUDE, "-c", os.path.join(B.CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        raise SystemExit(r.stderr)
    return obj


if not os.environ.get('MVSTER_VARIANT_NO_MAIN'):
    B.build_library()  # parallel invocations: build the main library once first and set MVSTER_VARIANT_NO_MAIN=1
with ThreadPoolExecutor(4) as ex:
    objs = list(ex.map(cc, ONLY))
objs += [os.path.join(B.BUILD_DIR, s[:-3] + ".o") for s in B._sources() if s not in ONLY]
lib = os.path.join(ROOT, "variants", "libvar_%s.so" % name)
r = subprocess.run([B._nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs +
                   ["-cudart", "static"], capture_output=True, text=True)
if r.returncode:
    raise SystemExit(r.stderr)
print(lib)
