"""CPU-side tests of the fusion tail: PFM / PLY writers against the reference's own bytes, and
the float64 oracle of depth2pts against the reference's points (tests/golden/fusion.npz, make_golden_fusion.py)."""
import numpy as np

import deep_reconstruction_with_epipolar_lines_mvster_b200 as mv
from oracle import mvster_oracle as O


def test_save_pfm_is_byte_identical_to_the_reference_writer(golden, tmp_path):
    g = golden("fusion")
    for key in ("gray", "color"):
        path = tmp_path / (key + ".pfm")
        mv.save_pfm(str(path), g["pfm_" + key])
        assert path.read_bytes() == g["pfm_%s_bytes" % key].tobytes(), key
        back, scale = mv.read_pfm(str(path))
        assert scale == 1.0 and np.array_equal(back, g["pfm_" + key])


def test_read_pfm_reads_the_reference_writers_file(golden, tmp_path):
    g = golden("fusion")
    path = tmp_path / "ref.pfm"
    path.write_bytes(g["pfm_gray_bytes"].tobytes())
    data, scale = mv.read_pfm(str(path))
    assert scale == 1.0 and data.dtype == np.float32 and np.array_equal(data, g["pfm_gray"])


def test_depth2pts_oracle_matches_reference(golden):
    g, f = golden("fusion"), golden("filter")
    with np.errstate(all="ignore"):
        for i in range(len(f["pairs"])):
            r = int(f["pairs"][i, 0])
            pts = O.depth2pts_np(f["depth_avg"][i], f["ks"][r], f["es"][r])
            ok = np.isfinite(g["points"][i]).all(1)
            assert np.abs(pts[ok] - g["points"][i][ok]).max() < 1e-9
            assert np.array_equal(np.isfinite(pts).all(1), ok)


def test_write_ply_round_trip(tmp_path):
    rng = np.random.RandomState(0)
    xyz = rng.normal(0, 100, (17, 3))
    rgb = rng.randint(0, 256, (17, 3)).astype(np.uint8)
    path = tmp_path / "cloud.ply"
    mv.write_ply(str(path), xyz, rgb)
    raw = path.read_bytes()
    head, _, body = raw.partition(b"end_header\n")
    assert b"format binary_little_endian 1.0" in head and b"element vertex 17" in head
    assert [l.split()[-1] for l in head.split(b"\n") if l.startswith(b"property")] == [b"x", b"y", b"z", b"red", b"green", b"blue"]
    vert = np.frombuffer(body, dtype=[("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("r", "u1"), ("g", "u1"), ("b", "u1")])
    assert len(vert) == 17 and np.allclose(vert["y"], xyz[:, 1].astype(np.float32)) and np.array_equal(vert["b"], rgb[:, 2])
