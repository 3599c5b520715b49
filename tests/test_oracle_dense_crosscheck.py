"""Cross-check the C oracle against the independent dense float64 evaluator + torch autograd
(tests/torch_reference.py) on small scenes -- pins what the reference's own tests do not
(multi-splat ordering, early termination, clamp, w-recursion of the backward, SH colour gradients,
hook statistics)."""
import numpy as np
import pytest
import torch

from taichi_3d_gaussian_splatting_b200.synthetic import make_scene
from taichi_3d_gaussian_splatting_b200.utils import inverse_SE3_qt_torch

from helpers import oracle_backward, oracle_forward, rel_err_global
from torch_reference import dense_render, postprocess_feature_grads


def _scene(seed, n=400, h=32, w=48, sigma=0.12, yaw=4.0, sh_degree=3):
    sc = make_scene(n, h, w, sigma, seed, sh_degree=sh_degree, yaw_degrees=yaw)
    sc.point_cloud[:, 2] = sc.point_cloud[:, 2] * 0.5  # denser coverage, some points behind near
    sc.point_cloud_features[:, 7] += 1.5  # more opaque -> exercises saturation / early stop
    sc.point_invalid_mask[::7] = 1
    return sc


@pytest.mark.parametrize("seed,band", [(11, 3), (12, 1), (13, 0)])
def test_oracle_matches_dense_autograd(seed, band):
    sc = _scene(seed)
    o, fwd, feats_n = oracle_forward(sc)
    H, W = sc.camera_info.camera_height, sc.camera_info.camera_width
    q_cp, t_cp = inverse_SE3_qt_torch(sc.q_pointcloud_camera, sc.t_pointcloud_camera)
    xyz = sc.point_cloud.clone().double().requires_grad_(True)
    feats = torch.from_numpy(feats_n).double().requires_grad_(True)  # q already normalised by the oracle
    image, aux = dense_render(xyz, feats, sc.point_invalid_mask, sc.camera_info.camera_intrinsics,
                              q_cp, t_cp, H, W)
    # forward
    assert aux["ids"].tolist() == fwd.point_id_in_camera_list.tolist()
    assert np.allclose(aux["uv"].detach().numpy(), fwd.point_uv, atol=2e-4)
    assert np.allclose(aux["conic"].detach().numpy(), fwd.point_uv_conic_and_rescale, rtol=2e-4, atol=1e-6)
    assert np.allclose(aux["color"].detach().numpy(), fwd.point_color, atol=1e-5)
    assert aux["ntiles"].tolist() == fwd.num_overlap_tiles.tolist()
    assert np.abs(image.detach().numpy() - fwd.image).max() < 1e-4
    assert np.abs(aux["acc_alpha"].detach().numpy() - fwd.pixel_accumulated_alpha).max() < 1e-4
    assert (aux["count"].numpy() == fwd.pixel_valid_point_count).all()
    assert np.abs(aux["depth"].detach().numpy() - fwd.depth).max() < 1e-3
    assert fwd.pixel_valid_point_count.max() >= 5  # the scene really exercises multi-splat blending
    # backward
    g = torch.Generator().manual_seed(seed + 100)
    grad_image = torch.randn((H, W, 3), generator=g, dtype=torch.float32)
    (image * grad_image.double()).sum().backward()
    bwd = oracle_backward(o, fwd, sc, feats_n, grad_image.numpy(), band)
    exp_feat = postprocess_feature_grads(feats.grad, band).numpy()
    assert rel_err_global(bwd.grad_pointcloud, xyz.grad.numpy()) < 1e-4
    for sl in (slice(0, 4), slice(4, 7), slice(7, 8), slice(8, 56)):
        assert rel_err_global(bwd.grad_pointcloud_features[:, sl], exp_feat[:, sl]) < 1e-4
    ids = fwd.point_id_in_camera_list
    assert np.allclose(bwd.grad_point_in_camera, bwd.grad_pointcloud[ids])
    mask = np.ones(sc.point_cloud.shape[0], bool)
    mask[ids] = False
    assert (bwd.grad_pointcloud[mask] == 0).all() and (bwd.grad_pointcloud_features[mask] == 0).all()


def test_oracle_quaternion_normalised_in_place():
    sc = _scene(5)
    sc.point_cloud_features[:, :4] *= 3.0
    before = sc.point_cloud_features.numpy().copy()
    _, fwd, feats = oracle_forward(sc)
    ids = fwd.point_id_in_camera_list
    assert np.allclose(np.linalg.norm(feats[ids, :4], axis=1), 1.0, atol=1e-6)
    out = np.ones(before.shape[0], bool)
    out[ids] = False
    assert (feats[out] == before[out]).all()  # rows outside the frustum untouched (GPCR:264-266)
    assert (feats[:, 4:] == before[:, 4:]).all()


def test_oracle_sort_is_stable_and_ranges_consistent():
    sc = _scene(6, n=1500, sigma=0.2)
    _, fwd, _ = oracle_forward(sc, depth_to_sort_key_scale=2.0)  # coarse keys -> many ties
    keys, vals = fwd.point_in_camera_sort_key, fwd.point_offset_with_sort_key
    assert (np.diff(keys) >= 0).all()
    same = np.diff(keys) == 0
    assert same.sum() > 50
    assert (np.diff(vals)[same] > 0).all()  # ties keep ascending in-camera offset
    tiles = (keys >> 32).astype(np.int64)
    for t in np.unique(tiles):
        s, e = fwd.tile_points_start[t], fwd.tile_points_end[t]
        assert (tiles[s:e] == t).all() and (s == 0 or tiles[s - 1] != t) and (e == len(tiles) or tiles[e] != t)
    empty = np.setdiff1d(np.arange(fwd.tile_points_start.shape[0]), np.unique(tiles))
    assert (fwd.tile_points_start[empty] == 0).all() and (fwd.tile_points_end[empty] == 0).all()


def test_oracle_offscreen_bbox_quirk():
    """GPCR:81-103: a splat entirely left/top of the image still gets tile column/row 0; one
    entirely right/bottom gets none."""
    import ctypes
    from oracle import lib
    uv = np.array([[-40.0, 20.0], [20.0, -40.0], [200.0, 20.0], [20.0, 200.0], [20.0, 20.0]], np.float32)
    r = np.array([2.0, 2.0, 2.0, 2.0, 0.1], np.float32)
    n = np.zeros(5, np.int32)
    lib().gso_generate_num_overlap_tiles(n.ctypes.data_as(ctypes.c_void_p), uv.ctypes.data_as(ctypes.c_void_p),
                                         r.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(5),
                                         ctypes.c_int(64), ctypes.c_int(64))
    assert n.tolist() == [1, 1, 0, 0, 1]


def test_oracle_empty_and_degenerate_inputs():
    """Edge cases the reference handles implicitly (GPCR:915-918, 934, 959, 980): no points, no point in the
    frustum, K == 0 -> zero image (the reference returns uninitialised memory there; we define zeros)."""
    from oracle import OracleRasterisation
    o = OracleRasterisation()
    K = np.array([[50.0, 0, 16], [0, 50.0, 16], [0, 0, 1]], np.float32)
    q, t = np.array([[0, 0, 0, 1]], np.float32), np.zeros((1, 3), np.float32)
    f = o.forward(np.zeros((0, 3), np.float32), np.zeros((0, 56), np.float32), np.zeros(0, np.int8),
                  np.zeros(0, np.int32), K, 32, 32, q, t)
    assert f.image.shape == (32, 32, 3) and not f.image.any() and f.point_id_in_camera_list.shape == (0,)
    xyz = np.array([[0, 0, -5.0], [100.0, 0, 5.0]], np.float32)  # behind the camera / far outside the margin
    feat = np.zeros((2, 56), np.float32)
    feat[:, 3] = 1
    f = o.forward(xyz, feat, np.zeros(2, np.int8), np.zeros(2, np.int32), K, 32, 32, q, t)
    assert f.point_id_in_camera_list.shape == (0,) and not f.image.any() and not f.tile_points_end.any()
    b = o.backward(f, np.ones((32, 32, 3), np.float32), xyz, feat, np.zeros(2, np.int32), K, t, 3)
    assert not b.grad_pointcloud.any() and not b.grad_pointcloud_features.any()


def test_oracle_splat_covering_every_tile_and_frustum_borders():
    from oracle import OracleRasterisation
    o = OracleRasterisation(near_plane=0.8, far_plane=10.0)
    K = np.array([[40.0, 0, 32], [0, 40.0, 32], [0, 0, 1]], np.float32)
    q, t = np.array([[0, 0, 0, 1]], np.float32), np.zeros((1, 3), np.float32)
    # point 0: huge splat at the centre -> all 16 tiles; points 1-4 sit exactly on the strict depth limits
    # (near < z < far) and on the +-48 px margin (-48 <= u < W + 48)
    xyz = np.array([[0, 0, 2.0], [0, 0, 0.8], [0, 0, 10.0], [-4.0, 0.4, 2.0], [4.0, 0, 2.0]], np.float32)
    feat = np.zeros((5, 56), np.float32)
    feat[:, 3] = 1
    feat[0, 4:7] = 2.0
    feat[1:, 4:7] = -3.0
    f = o.forward(xyz, feat, np.zeros(5, np.int8), np.zeros(5, np.int32), K, 64, 64, q, t)
    # u = 32 + 40 * x / z: x = -4 -> u = -48 (inside, >=), x = +4 -> u = 112 = W + 48 (outside, strict <)
    assert f.point_id_in_camera_list.tolist() == [0, 3]
    assert f.num_overlap_tiles[0] == 16 and (f.tile_points_end - f.tile_points_start >= 1).all()
    assert f.num_overlap_tiles[1] == 1  # left of the image: still gets tile column 0 (GPCR:81-103 quirk)
    assert f.pixel_valid_point_count.min() >= 1
