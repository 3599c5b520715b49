"""On-disk formats either side of the path: parquet columns of the reference scene files and the
official-3DGS binary PLY layout (quaternion wxyz <-> xyzw, SH DC + channel-major rest)."""
import struct

import numpy as np
import torch

from taichi_3d_gaussian_splatting_b200 import GaussianPointCloudScene as Scene
from taichi_3d_gaussian_splatting_b200.scene_io import FEATURE_COLUMNS, PLY_PROPERTIES, read_ply_vertices


def _random_scene(n=37, ratio=None, seed=0):
    g = torch.Generator().manual_seed(seed)
    xyz = torch.randn((n, 3), generator=g)
    feat = torch.randn((n, 56), generator=g)
    feat[:, :4] = feat[:, :4] / feat[:, :4].norm(dim=1, keepdim=True)
    return Scene(xyz, Scene.PointCloudSceneConfig(max_num_points_ratio=ratio), point_cloud_features=feat), xyz, feat


def test_capacity_padding_and_masks():
    scene, xyz, feat = _random_scene(ratio=2.5)
    assert scene.point_cloud.shape == (92, 3) and scene.point_cloud_features.shape == (92, 56)
    assert scene.point_invalid_mask.dtype == torch.int8 and scene.point_object_id.dtype == torch.int32
    assert scene.point_invalid_mask[:37].sum() == 0 and scene.point_invalid_mask[37:].all()
    assert torch.equal(scene.point_cloud[:37].detach(), xyz) and float(scene.point_cloud[37:].abs().max()) == 0.0


def test_parquet_round_trip_and_columns(tmp_path):
    import pandas as pd
    scene, xyz, feat = _random_scene(ratio=2.0)
    path = str(tmp_path / "scene.parquet")
    scene.to_parquet(path)  # only valid rows are written (GaussianPointCloudScene.py:133-134)
    frame = pd.read_parquet(path)
    assert list(frame.columns) == ["x", "y", "z"] + FEATURE_COLUMNS and len(frame) == 37
    back = Scene.from_parquet(path)
    assert torch.equal(back.point_cloud.detach(), xyz) and torch.equal(back.point_cloud_features.detach(), feat)


def test_parquet_without_features_is_initialised(tmp_path):
    import pandas as pd
    rng = np.random.default_rng(0)
    pts = rng.normal(size=(200, 3))
    rgb = rng.integers(0, 256, size=(200, 3))
    path = str(tmp_path / "sparse.parquet")
    pd.DataFrame(np.concatenate([pts, rgb], axis=1), columns=["x", "y", "z", "r", "g", "b"]).to_parquet(path)
    cfg = Scene.PointCloudSceneConfig(max_num_points_ratio=3.0, add_sphere=True, num_points_sphere=50, initial_alpha=0.05,
                                      initial_covariance_ratio=0.1, max_initial_covariance=3000.0)
    scene = Scene.from_parquet(path, cfg, generator=np.random.default_rng(1))
    n_valid = int((scene.point_invalid_mask == 0).sum())
    assert n_valid == 250 and scene.point_cloud.shape[0] == 750  # 200 + 50 sphere points, x3 capacity (Truck YAML: x10)
    f = scene.point_cloud_features.detach()
    assert torch.allclose(f[:n_valid, :4].norm(dim=1), torch.ones(n_valid), atol=1e-5)
    assert torch.allclose(f[:, 7], torch.full((750,), 0.05))
    assert torch.allclose(f[:n_valid, 4], f[:n_valid, 5]) and bool((f[:n_valid, 4] < 5).all())  # isotropic log-scales
    # colour: sigmoid(DC * c0) reproduces rgb/255 (clamped to 0.99)
    expect = np.clip(rgb[:5] / 255.0, 0, 0.99)
    got = torch.sigmoid(f[:5, [8, 24, 40]] * 0.28209479177387814).numpy()
    assert np.allclose(got, expect, atol=1e-5)
    r = scene.point_cloud[200:250].detach().norm(dim=1)  # the shell: radius = half extent * 4
    extent = max(pts[:, i].max() - pts[:, i].min() for i in range(3)) / 2
    assert torch.allclose(r, torch.full((50,), float(extent * 4.0)), rtol=1e-4)


def test_ply_layout_and_round_trip(tmp_path):
    scene, xyz, feat = _random_scene()
    path = str(tmp_path / "point_cloud.ply")
    scene.to_ply(path)
    raw = open(path, "rb").read()
    header, body = raw.split(b"end_header\n", 1)
    lines = header.decode().split("\n")
    assert lines[:3] == ["ply", "format binary_little_endian 1.0", "element vertex 37"]
    assert [ln.split()[-1] for ln in lines[3:] if ln.startswith("property float")] == PLY_PROPERTIES
    assert len(body) == 37 * 62 * 4
    first = struct.unpack("<62f", body[:62 * 4])
    assert np.allclose(first[0:3], xyz[0].numpy()) and first[3:6] == (0.0, 0.0, 0.0)
    assert np.allclose(first[6:9], feat[0, [8, 24, 40]].numpy())                    # f_dc = DC of r, g, b
    assert np.allclose(first[9:24], feat[0, 9:24].numpy())                          # f_rest: r1..r15 first
    assert np.isclose(first[54], feat[0, 7]) and np.allclose(first[55:58], feat[0, 4:7].numpy())
    assert np.allclose(first[58:62], feat[0, [3, 0, 1, 2]].numpy())                  # rot = w x y z
    back = Scene.from_ply(path)
    assert torch.allclose(back.point_cloud.detach(), xyz)
    assert torch.allclose(back.point_cloud_features.detach(), feat, atol=1e-6)
    v = read_ply_vertices(path)
    assert set(v) == set(PLY_PROPERTIES) and v["x"].shape == (37,)


# ---------------------------------------------------------------- against the REFERENCE's own scene class
# tests/golden/make_scene_golden.py ran /root/reference's GaussianPointCloudScene on sparse_points.parquet and stored
# what it built (scene_vectors.json) and wrote (reference_scene.parquet); it also verified that the reference loads a
# parquet written by our class with identical tensors.
def _scene_golden():
    import json
    import os
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    with open(os.path.join(here, "scene_vectors.json")) as f:
        return here, json.load(f)


def test_reads_the_parquet_the_reference_wrote():
    import os
    here, ref = _scene_golden()
    assert ref["reference_reads_our_parquet"] is True
    scene = Scene.from_parquet(os.path.join(here, "reference_scene.parquet"))
    valid = np.array(ref["point_invalid_mask"]) == 0
    assert scene.point_cloud.shape == (int(valid.sum()), 3)
    assert np.array_equal(scene.point_cloud.detach().numpy(), np.array(ref["point_cloud"], dtype=np.float32)[valid])
    assert np.array_equal(scene.point_cloud_features.detach().numpy(),
                          np.array(ref["point_cloud_features"], dtype=np.float32)[valid])


def test_initialisation_from_a_sparse_cloud_matches_the_reference():
    import os
    here, ref = _scene_golden()
    cfg = Scene.PointCloudSceneConfig(**ref["config"])
    scene = Scene.from_parquet(os.path.join(here, "sparse_points.parquet"), config=cfg)
    exp_xyz = np.array(ref["point_cloud"], dtype=np.float32)
    exp_f = np.array(ref["point_cloud_features"], dtype=np.float32)
    assert scene.point_cloud.shape == exp_xyz.shape == (100, 3)  # 40 points x max_num_points_ratio 2.5
    assert scene.point_invalid_mask.tolist() == ref["point_invalid_mask"]
    assert scene.point_object_id.tolist() == ref["point_object_id"]
    assert np.array_equal(scene.point_cloud.detach().numpy(), exp_xyz)
    f = scene.point_cloud_features.detach().numpy()
    valid = np.array(ref["point_invalid_mask"]) == 0
    # everything except the random unit quaternion is deterministic: kNN log-scales (ratio 0.7, clipped at 0.4),
    # opacity logit, SH DC from the 0..255 colours (clamped to 0.99; a zero channel gives -inf like the reference)
    assert np.allclose(f[:, 4:7], exp_f[:, 4:7], rtol=0, atol=1e-6)
    assert np.array_equal(f[:, 7], exp_f[:, 7]) and float(f[0, 7]) == -1.5
    with np.errstate(invalid="ignore"):
        same = (f[:, 8:] == exp_f[:, 8:]) | (np.abs(f[:, 8:] - exp_f[:, 8:]) <= 1e-5 * np.abs(exp_f[:, 8:]))
    assert same.all()
    assert np.isneginf(f[0, 24]) and np.isneginf(exp_f[0, 24])  # g = 0 on point 0
    assert np.allclose(np.linalg.norm(f[:, 0:4], axis=1), 1.0, atol=1e-6)
    assert np.allclose(np.linalg.norm(exp_f[:, 0:4], axis=1), 1.0, atol=1e-6)
    assert (f[~valid, 4:7] == 0).all() and (exp_f[~valid, 4:7] == 0).all()
