"""The reference arm of bench.py (oracle/ref_arm.py: the UNMODIFIED reference through MVS4net.forward with stand-in
FPN4 / reg2d) against the oracle's port of the same cascade on the same scene: a second pin of the oracle, and a check
that the baseline the driver times really is the reference's own code path.  CPU only."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_arm  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_arm.available(), reason="oracle/_ref absent (python oracle/make_ref.py needs /root/reference)")


def test_manifest_matches_the_copied_files():
    import hashlib, json
    with open(os.path.join(ref_arm.REF_DIR, "MANIFEST.json")) as f:
        man = json.load(f)
    assert set(man["sha256"]) == {"models/__init__.py", "models/MVS4Net.py", "models/mvs4net_utils.py", "test_mvs4.py"}
    for rel, digest in man["sha256"].items():
        with open(os.path.join(ref_arm.REF_DIR, rel), "rb") as f:
            assert hashlib.sha256(f.read()).hexdigest() == digest, rel


def test_reference_cascade_equals_the_oracle_port():
    import bench
    from deep_reconstruction_with_epipolar_lines_mvster_b200 import synthetic as syn
    from oracle import mvster_oracle as O
    sc = bench.make_scene(bench.scene_seed(0, 1), 5, 128, 192, jitter_index=1)
    mk = [lambda hypo, gt=sc["gt"][s], nz=sc["noise"][s]: bench.tracking_logits(hypo, gt, nz) for s in range(4)]
    rc = ref_arm.ReferenceCascade(torch, sc["features"], sc["projs"], sc["depth_values"], mk, syn.STAGE_GROUPS,
                                  syn.STAGE_NDEPTHS, syn.STAGE_SPLIT_ITV, 2.0)
    d_ref, c_ref = rc.run()
    d_ref2, _ = rc.run()                      # fixed logits: the timed passes repeat the set-up pass exactly
    d_port, c_port = O.cascade_port(sc["features"], sc["projs"], sc["depth_values"], mk, syn.STAGE_GROUPS,
                                    syn.STAGE_NDEPTHS, syn.STAGE_SPLIT_ITV, 2.0)
    assert torch.equal(d_ref, d_ref2)
    assert d_ref.shape == d_port.shape == (1, 128, 192)
    # same op sequence on the same inputs: the port is bit-identical to the reference's own forward
    assert torch.equal(d_ref, d_port)
    assert (c_ref - c_port).abs().max().item() == 0.0


def test_both_bench_arms_share_one_config_dictionary():
    import argparse, bench
    args = argparse.Namespace(views=5, height=864, width=1152, scenes=8, dtype="fp32")
    assert bench.workload_config(args, 1) == bench.workload_config(args, 1)
    assert "host_affinity" not in bench.workload_config(args, 8)
