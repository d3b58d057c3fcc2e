#!/usr/bin/env python
"""Golden fixture for the fusion tail from the UNMODIFIED reference:
``depth2pts_np`` / ``get_pixel_grids_np`` lifted out of ``test_mvs4.py`` with ``ast`` (the file is not importable), and
``save_pfm`` / ``read_pfm`` imported from ``datasets/data_io.py``.  Inputs come from the committed filter fixture.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_fusion.py      ->  tests/golden/fusion.npz
"""
import ast
import importlib.util
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("MVSTER_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True


def main():
    tree = ast.parse(open(os.path.join(REF, "test_mvs4.py")).read())
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("depth2pts_np", "get_pixel_grids_np")]
    ns = {"np": np}
    exec(compile(ast.Module(body=keep, type_ignores=[]), "test_mvs4.py", "exec"), ns)
    spec = importlib.util.spec_from_file_location("ref_data_io", os.path.join(REF, "datasets", "data_io.py"))
    dio = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(dio)

    f = np.load(os.path.join(HERE, "filter.npz"))
    avg, final, ks, es, pairs = f["depth_avg"], f["final"], f["ks"], f["es"], f["pairs"]
    rec = {}
    with np.errstate(all="ignore"):
        pts = [ns["depth2pts_np"](avg[i], ks[int(pairs[i, 0])], es[int(pairs[i, 0])]) for i in range(len(pairs))]
    rec["points"] = np.stack(pts)                                          # [R, H*W, 3] float64
    rec["vertices"] = np.concatenate([p[final[i].flatten()] for i, p in enumerate(pts)], 0)
    rng = np.random.RandomState(5)
    imgs = rng.uniform(0, 1, size=(len(ks),) + avg.shape[1:] + (3,)).astype(np.float32)
    rec["images"] = imgs
    rec["colors"] = np.concatenate([(imgs[int(pairs[i, 0])][final[i]] * 255).astype(np.uint8) for i in range(len(pairs))], 0)
    # PFM: the reference writer's exact bytes for a depth map and a 3-channel image, and its reader's output
    depth = np.nan_to_num(avg[1].astype(np.float32))
    with tempfile.TemporaryDirectory() as d:
        dio.save_pfm(os.path.join(d, "a.pfm"), depth)
        rec["pfm_gray_bytes"] = np.frombuffer(open(os.path.join(d, "a.pfm"), "rb").read(), dtype=np.uint8)
        dio.save_pfm(os.path.join(d, "b.pfm"), imgs[0][:6, :5])
        rec["pfm_color_bytes"] = np.frombuffer(open(os.path.join(d, "b.pfm"), "rb").read(), dtype=np.uint8)
        back, scale = dio.read_pfm(os.path.join(d, "a.pfm"))
        assert np.array_equal(back, depth) and scale == 1.0
    rec["pfm_gray"] = depth
    rec["pfm_color"] = imgs[0][:6, :5]
    np.savez_compressed(os.path.join(HERE, "fusion.npz"), **rec)
    print("fusion: %d views, %d kept points, %.2f MB" % (len(pts), len(rec["vertices"]),
                                                      os.path.getsize(os.path.join(HERE, "fusion.npz")) / 1e6))


if __name__ == "__main__":
    main()
