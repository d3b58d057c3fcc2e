"""CPU-side tests: the C-ABI library loads and exports everything include/mvster_b200.h declares, error paths fail
loudly without a GPU, and the host-side sharding logic works under a world_size-2 gloo group."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

import deep_reconstruction_with_epipolar_lines_mvster_b200 as mv
from deep_reconstruction_with_epipolar_lines_mvster_b200 import synthetic as syn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "mvster_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mvster_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = mv.load(build_if_missing=True)
    declared = _header_symbols()
    assert len(declared) >= 15
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert sorted(mv.EXPORTED_SYMBOLS) == declared
    assert lib.mvster_version() == 1


def test_library_is_sm100a_only_and_has_256bit_loads():
    out = subprocess.run(["cuobjdump", "-lelf", mv.library_path()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback():
    feats = [syn.smooth_features(1, 8, 8, 8, s) for s in range(2)]
    proj = torch.from_numpy(syn.proj_matrices(1, 2, 8, 8, 3))
    hypo = torch.full((1, 4, 8, 8), 600.0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        mv.epipolar_aggregate(feats, proj, hypo, 4, 2.0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        mv.schedule_inverse_range(torch.ones(1, 4, 4), torch.ones(1, 4, 4), 4, 8, 8)
    with pytest.raises(RuntimeError):
        mv.check_geometric_consistency(np.ones((4, 4), np.float32), np.eye(3), np.eye(4), np.ones((4, 4), np.float32),
                                       np.eye(3), np.eye(4), device="cpu")


@pytest.mark.skipif(torch.cuda.is_available(), reason="exercises the no-device error path")
def test_c_abi_reports_missing_device():
    import ctypes
    lib = mv.load()
    buf = (ctypes.c_float * 64)()
    st = lib.mvster_compose_homographies(ctypes.cast(buf, ctypes.c_void_p), ctypes.cast(buf, ctypes.c_void_p), 1, 2, None)
    assert st == 5  # MVSTER_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.mvster_last_error()
    st = lib.mvster_compose_homographies(None, None, 1, 2, None)
    assert st == 1 and b"null" in lib.mvster_last_error()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "deep_reconstruction_with_epipolar_lines_mvster_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            text = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in text and "from oracle" not in text, fn


def test_sharding_round_robin():
    assert mv.shard_round_robin(8, 0, 1) == list(range(8))
    parts = [mv.shard_round_robin(49, r, 8) for r in range(8)]
    assert sorted(sum(parts, [])) == list(range(49))
    assert max(map(len, parts)) - min(map(len, parts)) <= 1
    with pytest.raises(ValueError):
        mv.shard_round_robin(8, 2, 2)


def test_numa_cpu_list_parsing_and_graceful_absence():
    from deep_reconstruction_with_epipolar_lines_mvster_b200 import sharding
    assert sharding._parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    if not torch.cuda.is_available():   # no GPU: the topology lookup must return None, not raise
        assert sharding.gpu_local_cpus(0) is None and sharding.bind_to_gpu_numa_node(0) is None


def test_synthetic_rig_layout():
    p = syn.proj_matrices(2, 5, 512, 640, 0)
    assert p.shape == (2, 5, 2, 4, 4) and p.dtype == np.float32
    assert np.allclose(p[0, 0, 0], np.eye(4)) and p[0, 0, 1, 3, 3] == 0
    assert abs(p[0, 0, 1, 0, 0] - 2892.33 * 640 / 1600 / 8) < 1e-3
    assert syn.stage_shape(864, 1152, 0) == (108, 144)


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
import deep_reconstruction_with_epipolar_lines_mvster_b200 as mv
dist.init_process_group("gloo", init_method="env://")
rank, _, world = mv.rank_world()
mine = mv.shard_round_robin(8, rank, world)           # 8 scenes over the ranks, no data-path collective
t = torch.tensor([float(10 + rank)], dtype=torch.float64)   # stand-in for this rank's device time
dist.barrier()
dist.all_reduce(t, op=dist.ReduceOp.MAX)               # the only exchange: max-over-ranks timing
cnt = torch.tensor([len(mine)]); dist.all_reduce(cnt)
assert t.item() == 10 + world - 1 and cnt.item() == 8, (t, cnt)
if rank == 0: print("OK", world, mine)
dist.destroy_process_group()
"""


def test_world_size_2_gloo_sharding(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT="29517")
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=120) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "OK 2 [0, 2, 4, 6]" in outs[0][0]


def test_bench_reference_arm_prints_contract_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--height", "64", "--width", "64", "--views", "3"],
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr
    import json
    line = json.loads(res.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "depth maps/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["e2e"]["h2d_bytes_per_step"] == 0
