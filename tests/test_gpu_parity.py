"""-m gpu parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical inputs.

Tolerances (BASELINE.json north_star): forward RGB within 1e-4 abs, gradients within 1e-3 rel
(denominator floor 1e-6 * max|g|).  Integer / index work (compaction order, tile counts,
keys, sort order, tile ranges, last-effective offsets, pixel counts) must be bit-exact; the per-point
float attributes are bit-exact too because the preprocess kernel evaluates the oracle's op order
(-fmad=false, exp rounded once from double).
"""
import numpy as np
import pytest
import torch

from taichi_3d_gaussian_splatting_b200.synthetic import CONFIGS, make_scene

from helpers import grad_close, oracle_backward, oracle_forward, rel_err
from gpu_helpers import count_above, cuda_scene, make_op, n, run_forward

pytestmark = pytest.mark.gpu


def _small_scene(seed, npts=3000, h=64, w=96, sigma=0.06, yaw=7.0, sh_degree=3):
    sc = make_scene(npts, h, w, sigma, seed, sh_degree=sh_degree, yaw_degrees=yaw)
    sc.point_cloud[:, 2] = sc.point_cloud[:, 2] * 0.6
    sc.point_cloud_features[:, 7] += 1.0
    sc.point_cloud_features[:, :4] *= 1.7  # un-normalised quaternions on input
    sc.point_invalid_mask[::11] = 1
    return sc


def _check_stages(frame, fwd, exact_floats=True, max_tiles=None):
    assert frame.num_points_in_camera == fwd.point_id_in_camera_list.shape[0]
    assert (n(frame.point_id_in_camera_list) == fwd.point_id_in_camera_list).all()
    assert (n(frame.num_overlap_tiles) == fwd.num_overlap_tiles).all()
    pairs = [(frame.point_uv, fwd.point_uv), (frame.point_in_camera, fwd.point_in_camera),
             (frame.point_uv_conic_and_rescale, fwd.point_uv_conic_and_rescale),
             (frame.point_alpha_after_activation, fwd.point_alpha_after_activation),
             (frame.point_color, fwd.point_color), (frame.point_radii, fwd.point_radii)]
    for got, exp in pairs:
        got = n(got)
        if exact_floats:
            assert np.array_equal(got, exp), f"max abs diff {np.abs(got - exp).max()}"
        else:
            assert np.allclose(got, exp, rtol=1e-5, atol=1e-5)
    L = frame.layout
    okeys = fwd.point_in_camera_sort_key
    packed = ((okeys >> 32) << L.depth_bits) | (okeys & 0xFFFFFFFF)
    if frame.flags & 8:  # GSB_FLAG_KEEP_ALL_TILE_PAIRS: the sorted list IS the reference's list
        assert frame.num_keys == fwd.point_offset_with_sort_key.shape[0]
        assert (n(frame.sorted_keys) == packed).all()
        assert (n(frame.point_offset_with_sort_key) == fwd.point_offset_with_sort_key).all()
        assert (n(frame.tile_points_start) == fwd.tile_points_start).all()
        assert (n(frame.tile_points_end) == fwd.tile_points_end).all()
    else:
        _check_filtered_lists(frame, fwd, packed, max_tiles)


def _check_filtered_lists(frame, fwd, packed, max_tiles=None):
    """Default mode: only (tile, splat) pairs that can reach alpha >= 1/255 on some pixel get a key.  Per tile the
    emitted list must be a SUBSEQUENCE of the reference's list (same relative order, same keys), and every dropped
    pair must be one that contributes to no pixel of the tile in the oracle's arithmetic."""
    mk, mv = n(frame.sorted_keys), n(frame.point_offset_with_sort_key)
    ms, me = n(frame.tile_points_start), n(frame.tile_points_end)
    W, H = frame.width, frame.height
    tiles_x = W // 16
    assert frame.num_keys == mk.shape[0] <= packed.shape[0]
    assert (np.diff(mk) >= 0).all()
    tiles = np.arange(ms.shape[0])
    if max_tiles is not None and tiles.shape[0] > max_tiles:
        tiles = np.random.default_rng(0).choice(tiles, max_tiles, replace=False)
    ys, xs = np.meshgrid(np.arange(16), np.arange(16), indexing="ij")
    dropped_total = 0
    for t in tiles:
        os_, oe = fwd.tile_points_start[t], fwd.tile_points_end[t]
        ref_vals, ref_keys = fwd.point_offset_with_sort_key[os_:oe], packed[os_:oe]
        got_vals, got_keys = mv[ms[t]:me[t]], mk[ms[t]:me[t]]
        # subsequence: walk the reference list once
        keep = np.zeros(ref_vals.shape[0], bool)
        j = 0
        for i in range(ref_vals.shape[0]):
            if j < got_vals.shape[0] and ref_vals[i] == got_vals[j] and ref_keys[i] == got_keys[j]:
                keep[i] = True
                j += 1
        assert j == got_vals.shape[0], f"tile {t}: emitted list is not a subsequence of the reference list"
        drop = ref_vals[~keep]
        if drop.shape[0] == 0:
            continue
        dropped_total += drop.shape[0]
        px = (t % tiles_x) * 16 + xs.reshape(-1) + 0.5
        py = (t // tiles_x) * 16 + ys.reshape(-1) + 0.5
        uv, cr = fwd.point_uv[drop], fwd.point_uv_conic_and_rescale[drop]
        dx, dy = px[None, :] - uv[:, 0:1], py[None, :] - uv[:, 1:2]
        power = -0.5 * (dx * dx * cr[:, 0:1] + dy * dy * cr[:, 2:3]) - dx * dy * cr[:, 1:2]
        alpha = np.exp(power) * cr[:, 3:4] * fwd.point_alpha_after_activation[drop][:, None]
        assert alpha.max() < 1.0 / 255.0, f"tile {t}: dropped a pair with alpha {alpha.max()}"
    return dropped_total


@pytest.mark.parametrize("keep_all", [False, True])
@pytest.mark.parametrize("exact_exp", [True, False])
@pytest.mark.parametrize("force_key64", [False, True])
def test_c1_forward_backward_vs_oracle(exact_exp, force_key64, keep_all):
    """BASELINE config 1 (correctness gate): 1e4 Gaussians, 256x256, SH deg 0."""
    scene = make_scene(**CONFIGS["C1"])
    o, fwd, feats_n = oracle_forward(scene)
    sc = cuda_scene(scene, requires_grad=True)
    captured = {}
    op = make_op(hook=lambda h: captured.setdefault("hook", h), exact_exp=exact_exp, force_key64=force_key64,
                 keep_all_tile_pairs=keep_all)
    image, depth, count = run_forward(op, sc, band=0)
    frame = op.last_frame
    assert frame.layout.key_bytes == (8 if force_key64 else 4)
    _check_stages(frame, fwd)
    # in-place quaternion normalisation (GPCR:264-266)
    assert np.array_equal(n(sc.point_cloud_features), feats_n)
    # forward image: 1e-4 abs
    assert np.abs(n(image) - fwd.image).max() <= 1e-4
    assert np.abs(n(depth) - fwd.depth).max() <= 1e-3
    assert count_above(n(count), fwd.pixel_valid_point_count, 0) == 0
    g = torch.Generator().manual_seed(1)
    grad_image = torch.randn(image.shape, generator=g, dtype=torch.float32)
    image.backward(grad_image.cuda())
    bwd = oracle_backward(o, fwd, scene, feats_n, grad_image.numpy(), 0)
    assert grad_close(n(sc.point_cloud.grad), bwd.grad_pointcloud)[0], grad_close(n(sc.point_cloud.grad), bwd.grad_pointcloud)
    gf = n(sc.point_cloud_features.grad)
    for sl in (slice(0, 4), slice(4, 7), slice(7, 8), slice(8, 56)):
        assert grad_close(gf[:, sl], bwd.grad_pointcloud_features[:, sl])[0], grad_close(gf[:, sl], bwd.grad_pointcloud_features[:, sl])
    h = captured["hook"]
    assert (n(h.point_id_in_camera_list) == bwd.point_id_in_camera_list).all()
    assert grad_close(n(h.grad_point_in_camera), bwd.grad_point_in_camera)[0], grad_close(n(h.grad_point_in_camera), bwd.grad_point_in_camera)
    assert grad_close(n(h.grad_pointfeatures_in_camera), bwd.grad_pointfeatures_in_camera)[0], grad_close(n(h.grad_pointfeatures_in_camera), bwd.grad_pointfeatures_in_camera)
    assert grad_close(n(h.grad_viewspace), bwd.grad_viewspace)[0], grad_close(n(h.grad_viewspace), bwd.grad_viewspace)
    assert grad_close(n(h.magnitude_grad_viewspace), bwd.magnitude_grad_viewspace)[0], grad_close(n(h.magnitude_grad_viewspace), bwd.magnitude_grad_viewspace)
    assert np.abs(n(h.magnitude_grad_viewspace_on_image) - bwd.magnitude_grad_viewspace_on_image).max() <= \
        1e-3 * max(1.0, np.abs(bwd.magnitude_grad_viewspace_on_image).max())
    assert (n(h.num_overlap_tiles) == bwd.num_overlap_tiles).all()
    assert (n(h.num_affected_pixels) == bwd.num_affected_pixels).all()
    assert np.array_equal(n(h.point_depth), bwd.point_depth)
    assert np.array_equal(n(h.point_uv_in_camera), bwd.point_uv_in_camera)
    assert h.grad_point_in_camera.shape[1] == 3 and h.grad_pointfeatures_in_camera.shape[1] == 56 \
        and h.grad_viewspace.shape[1] == 2  # reference test_backward_hook widths


@pytest.mark.parametrize("seed,band,scale", [(21, 3, 100.0), (22, 1, 10.0), (23, 2, 2.0)])
def test_small_scene_sh3_vs_oracle(seed, band, scale):
    """SH deg 3, rotated camera, un-normalised q, invalid slots, saturation, coarse depth keys (ties)."""
    scene = _small_scene(seed)
    o, fwd, feats_n = oracle_forward(scene, depth_to_sort_key_scale=scale, near_plane=0.4)
    sc = cuda_scene(scene, requires_grad=True)
    op = make_op(exact_exp=True, depth_to_sort_key_scale=scale, near_plane=0.4)
    image, depth, count = run_forward(op, sc, band=band)
    _check_stages(op.last_frame, fwd)
    assert fwd.pixel_valid_point_count.max() >= 8
    assert np.abs(n(image) - fwd.image).max() <= 1e-4
    assert (n(count) == fwd.pixel_valid_point_count).all()
    g = torch.Generator().manual_seed(seed)
    grad_image = torch.randn(image.shape, generator=g, dtype=torch.float32)
    image.backward(grad_image.cuda())
    bwd = oracle_backward(o, fwd, scene, feats_n, grad_image.numpy(), band)
    assert grad_close(n(sc.point_cloud.grad), bwd.grad_pointcloud)[0], grad_close(n(sc.point_cloud.grad), bwd.grad_pointcloud)
    gf = n(sc.point_cloud_features.grad)
    for sl in (slice(0, 4), slice(4, 7), slice(7, 8), slice(8, 56)):
        assert grad_close(gf[:, sl], bwd.grad_pointcloud_features[:, sl])[0], grad_close(gf[:, sl], bwd.grad_pointcloud_features[:, sl])


def test_two_points_scene_golden():
    """Scene of reference tests/GaussianPointCloudRasterisation_test.py:152-205, hand-derived values
     through the CUDA path."""
    dev = "cuda"
    xyz = torch.tensor([[0.0, 0.0, 1.0], [0.0, 0.0, 2.0]], device=dev)
    feat = torch.zeros((2, 56), device=dev)
    feat[:, 3] = 1.0
    feat[0, 4:7] = 1.0
    feat[1, 4:7] = 4.0
    feat[:, 8] = 5.0
    feat[:, 24] = 1.0
    feat[:, 40] = 1.0
    from taichi_3d_gaussian_splatting_b200 import CameraInfo
    from gpu_helpers import Input
    mask = torch.tensor([1, 0], dtype=torch.int8, device=dev)
    op = make_op(near_plane=0.0, far_plane=10.0, exact_exp=True)
    image, depth, count = op(Input(
        point_cloud=xyz, point_cloud_features=feat, point_object_id=torch.zeros(2, dtype=torch.int32, device=dev),
        point_invalid_mask=mask,
        camera_info=CameraInfo(torch.tensor([[1., 0, 8], [0, 1, 8], [0, 0, 1]], device=dev), 16, 16, 0),
        q_pointcloud_camera=torch.tensor([[0., 0, 0, 1]], device=dev),
        t_pointcloud_camera=torch.zeros((1, 3), device=dev), color_max_sh_band=0))
    img = n(image)
    assert np.allclose(img[7, 7], [0.40162393, 0.28481966, 0.28481966], atol=2e-6)
    for (v, u) in [(0, 0), (0, 15), (15, 15)]:
        assert np.allclose(img[v, u], [0.37256175, 0.26420963, 0.26420963], atol=2e-6)
    assert np.allclose(n(depth), 2.0, atol=1e-5)
    assert (n(count) == 1).all()
    assert op.last_frame.num_points_in_camera == 1 and op.last_frame.num_keys == 1


def test_multi_object_poses_vs_oracle():
    scene = _small_scene(31, npts=2000)
    scene.point_object_id[1::2] = 1
    q = torch.tensor([[0.0, 0.03, 0.0, 1.0], [0.02, -0.05, 0.01, 0.99]])
    scene.q_pointcloud_camera = q / q.norm(dim=-1, keepdim=True)
    scene.t_pointcloud_camera = torch.tensor([[0.0, 0.0, 0.0], [0.3, -0.1, -0.5]])
    o, fwd, feats_n = oracle_forward(scene)
    sc = cuda_scene(scene, requires_grad=True)
    op = make_op(exact_exp=True)
    image, _, _ = run_forward(op, sc, band=3)
    _check_stages(op.last_frame, fwd)
    assert np.abs(n(image) - fwd.image).max() <= 1e-4
    grad_image = torch.ones(image.shape)
    image.backward(grad_image.cuda())
    bwd = oracle_backward(o, fwd, scene, feats_n, grad_image.numpy(), 3)
    assert grad_close(n(sc.point_cloud.grad), bwd.grad_pointcloud)[0], grad_close(n(sc.point_cloud.grad), bwd.grad_pointcloud)
    assert grad_close(n(sc.point_cloud_features.grad), bwd.grad_pointcloud_features)[0], grad_close(n(sc.point_cloud_features.grad), bwd.grad_pointcloud_features)


def test_empty_and_degenerate_frames():
    """K == 0 (the reference returns uninitialised memory, GPCR:967-997; we define zeros), all-invalid
    scene, and no gradient when nothing requires grad (GPCR:1028)."""
    scene = _small_scene(41, npts=500)
    scene.point_invalid_mask[:] = 1
    sc = cuda_scene(scene, requires_grad=True)
    op = make_op()
    image, depth, count = run_forward(op, sc)
    assert op.last_frame.num_points_in_camera == 0 and op.last_frame.num_keys == 0
    assert float(image.abs().max()) == 0.0 and float(depth.abs().max()) == 0.0 and int(count.max()) == 0
    image.sum().backward()
    assert float(sc.point_cloud.grad.abs().max()) == 0.0
    assert float(sc.point_cloud_features.grad.abs().max()) == 0.0
    # everything behind the camera
    scene2 = _small_scene(42, npts=500)
    scene2.point_cloud[:, 2] = -scene2.point_cloud[:, 2].abs() - 1.0
    image2, _, _ = run_forward(make_op(), cuda_scene(scene2))
    assert float(image2.abs().max()) == 0.0
    # no requires_grad -> autograd never calls backward; forward under no_grad works
    with torch.no_grad():
        image3, _, _ = run_forward(make_op(), cuda_scene(_small_scene(43)))
    assert not image3.requires_grad


def test_key_capacity_overflow_regrows():
    scene = _small_scene(51)
    _, fwd, _ = oracle_forward(scene)
    op = make_op(initial_key_capacity=64, exact_exp=True)
    image, _, _ = run_forward(op, cuda_scene(scene))
    assert op.last_frame.key_capacity >= op.last_frame.num_keys > 64
    _check_stages(op.last_frame, fwd)
    assert np.abs(n(image) - fwd.image).max() <= 1e-4


def test_rgb_only_matches_full():
    scene = _small_scene(61)
    full, _, _ = run_forward(make_op(exact_exp=True), cuda_scene(scene))
    rgb, depth, count = run_forward(make_op(exact_exp=True, rgb_only=True), cuda_scene(scene))
    assert torch.equal(full, rgb)
    assert float(depth.abs().max()) == 0.0 and int(count.max()) == 0


def test_requires_cuda_tensors_and_alignment():
    scene = _small_scene(71, npts=100)
    op = make_op()
    with pytest.raises(RuntimeError, match="no CPU path"):
        run_forward(op, scene)  # CPU tensors: there is no fallback
    from taichi_3d_gaussian_splatting_b200 import CameraInfo
    sc = cuda_scene(scene)
    sc.camera_info = CameraInfo(sc.camera_info.camera_intrinsics, 60, 96, 0)
    with pytest.raises(AssertionError):
        run_forward(op, sc)  # H % 16 != 0 (GPCR:1193-1194)


def test_find_tile_start_and_end_known_answer(golden):
    """reference tests/GaussianPointCloudRasterisation_test.py:18-51 through the CUDA kernel."""
    from taichi_3d_gaussian_splatting_b200 import find_tile_start_and_end
    g = golden["tile_ranges"]
    keys = torch.tensor(g["keys"], dtype=torch.int64, device="cuda")
    start = torch.zeros(g["num_tiles"], dtype=torch.int32, device="cuda")
    end = torch.zeros(g["num_tiles"], dtype=torch.int32, device="cuda")
    find_tile_start_and_end(keys, start, end)
    assert start.tolist() == g["start"] and end.tolist() == g["end"]


def test_rgb_only_refuses_backward():
    sc = cuda_scene(_small_scene(62, npts=300), requires_grad=True)
    image, _, _ = run_forward(make_op(rgb_only=True), sc)
    with pytest.raises(RuntimeError, match="inference-only"):
        image.sum().backward()


def test_two_frames_in_flight_before_backward():
    """Two forwards (different cameras) before either backward: each frame owns its workspace."""
    scene = _small_scene(63, npts=1500)
    sc = cuda_scene(scene, requires_grad=True)
    op = make_op(exact_exp=True)
    img_a, _, _ = run_forward(op, sc)
    sc2 = cuda_scene(_small_scene(63, npts=1500, yaw=-5.0))
    sc2.point_cloud, sc2.point_cloud_features = sc.point_cloud, sc.point_cloud_features
    img_b, _, _ = run_forward(op, sc2)
    (img_a.sum() + 2.0 * img_b.sum()).backward()
    g_both = sc.point_cloud.grad.clone()
    sc.point_cloud.grad = None
    sc.point_cloud_features.grad = None
    ia, _, _ = run_forward(op, sc)
    ia.sum().backward()
    ga = sc.point_cloud.grad.clone()
    sc.point_cloud.grad = None
    ib, _, _ = run_forward(op, sc2)
    (2.0 * ib.sum()).backward()
    gb = sc.point_cloud.grad.clone()
    assert torch.allclose(g_both, ga + gb, rtol=1e-4, atol=1e-4 * float(g_both.abs().max()))


def test_c2_full_size_statistical_parity_and_properties():
    """BASELINE config 2 at FULL size (4.3e5 Gaussians, 976x544, SH deg 3, fwd+bwd), default fast-exp path.
    Bit-exact integer stages; image within 1e-4 except the handful of (pixel, splat) pairs that sit within
    an ulp of the alpha = 1/255 cut-off (each moves a pixel by <= 4e-3); gradients within tolerance on all
    but a tiny fraction of entries (the same flips), never off by more than 1e-3 of the largest gradient."""
    scene = make_scene(**CONFIGS["C2"])
    o, fwd, feats_n = oracle_forward(scene)
    sc = cuda_scene(scene, requires_grad=True)
    op = make_op()
    image, depth, count = run_forward(op, sc, band=3)
    frame = op.last_frame
    _check_stages(frame, fwd, max_tiles=150)
    # size-independent properties
    keys = frame.sorted_keys
    assert bool((keys[1:] >= keys[:-1]).all())
    same = keys[1:] == keys[:-1]
    vals = frame.point_offset_with_sort_key
    assert bool((vals[1:][same] > vals[:-1][same]).all())  # stability: ties keep ascending offset
    assert frame.num_keys <= int(frame.num_overlap_tiles.sum()) == fwd.point_offset_with_sort_key.shape[0]
    assert frame.num_keys < 0.8 * fwd.point_offset_with_sort_key.shape[0]  # about a third of the pairs can never contribute
    assert bool(torch.isfinite(image).all()) and float(image.min()) >= 0.0
    d = np.abs(n(image) - fwd.image)
    assert (d > 1e-4).sum() <= 40 and d.max() <= 5e-3
    assert count_above(n(count), fwd.pixel_valid_point_count, 0) <= 40
    g = torch.Generator().manual_seed(5)
    grad_image = torch.randn(image.shape, generator=g, dtype=torch.float32)
    image.backward(grad_image.cuda())
    bwd = oracle_backward(o, fwd, scene, feats_n, grad_image.numpy(), 3)
    gx, gf = n(sc.point_cloud.grad), n(sc.point_cloud_features.grad)
    for got, exp in ((gx, bwd.grad_pointcloud), (gf[:, :4], bwd.grad_pointcloud_features[:, :4]),
                     (gf[:, 4:7], bwd.grad_pointcloud_features[:, 4:7]), (gf[:, 7:8], bwd.grad_pointcloud_features[:, 7:8]),
                     (gf[:, 8:], bwd.grad_pointcloud_features[:, 8:])):
        ok, worst, nviol = grad_close(got, exp)
        assert nviol <= 2e-3 * exp.size, (nviol, exp.size)
        assert np.abs(got - exp).max() <= 1e-3 * np.abs(exp).max()
    # linearity of the backward in dL/dimage
    sc.point_cloud.grad = None
    sc.point_cloud_features.grad = None
    image2, _, _ = run_forward(op, sc, band=3)
    image2.backward(2.0 * grad_image.cuda())
    assert np.allclose(n(sc.point_cloud.grad), 2.0 * gx, rtol=1e-3, atol=1e-4 * np.abs(gx).max())


def test_render_host_c_abi_entry_point():
    """gsb200_render_host: scene resident on the device, pose + intrinsics from HOST memory, image back in HOST
    memory (the render-loop body of gaussian_point_render.py:106-121) -- must equal the operator's image."""
    import ctypes
    from taichi_3d_gaussian_splatting_b200 import _lib
    scene = _small_scene(81)
    sc = cuda_scene(scene)
    op = make_op(exact_exp=True)
    image, depth, count = run_forward(op, sc)
    frame = op.last_frame
    H, W, N = frame.height, frame.width, sc.point_cloud.shape[0]
    lib = _lib.load()
    ws = torch.empty(frame.layout.total_bytes, dtype=torch.uint8, device="cuda")
    out = [torch.empty((H, W, 3), device="cuda"), torch.empty((H, W), device="cuda"), torch.empty((H, W), device="cuda"),
           torch.empty((H, W), dtype=torch.int32, device="cuda"), torch.empty((H, W), dtype=torch.int32, device="cuda")]
    cfg = op.config
    args = _lib.GsbForwardArgs(
        num_points=N, pointcloud=sc.point_cloud.data_ptr(), pointcloud_features=sc.point_cloud_features.data_ptr(),
        point_invalid_mask=sc.point_invalid_mask.data_ptr(), point_object_id=sc.point_object_id.data_ptr(), num_objects=1,
        camera_height=H, camera_width=W, near_plane=cfg.near_plane, far_plane=cfg.far_plane,
        depth_to_sort_key_scale=cfg.depth_to_sort_key_scale, rgb_only=0, flags=frame.flags, workspace=ws.data_ptr(),
        workspace_bytes=frame.layout.total_bytes, key_capacity=frame.key_capacity, rasterized_image=out[0].data_ptr(),
        rasterized_depth=out[1].data_ptr(), pixel_accumulated_alpha=out[2].data_ptr(),
        pixel_offset_of_last_effective_point=out[3].data_ptr(), pixel_valid_point_count=out[4].data_ptr(),
        stream=torch.cuda.current_stream().cuda_stream)
    q_host = scene.q_pointcloud_camera.clone().pin_memory()
    t_host = scene.t_pointcloud_camera.clone().pin_memory()
    K_host = scene.camera_info.camera_intrinsics.clone().pin_memory()
    staging = torch.empty(32, device="cuda")
    image_host = torch.empty((H, W, 3)).pin_memory()
    counters = torch.zeros(4, dtype=torch.int64).pin_memory()
    rc = lib.gsb200_render_host(ctypes.byref(args), q_host.data_ptr(), t_host.data_ptr(), K_host.data_ptr(),
                                staging.data_ptr(), image_host.data_ptr(), counters.data_ptr())
    _lib.check(rc, "gsb200_render_host")
    # the first forward already normalised q in place; normalising again may move q by an ulp
    assert torch.allclose(image_host, image.cpu(), atol=1e-5)
    assert counters[0].item() == frame.num_points_in_camera and counters[1].item() == frame.num_keys and counters[2].item() == 0
    # error convention: bad arguments return a negative code and a message, never throw
    assert lib.gsb200_render_host(None, None, None, None, None, None, None) < 0
    assert b"null" in lib.gsb200_last_error()


def test_filtered_and_unfiltered_key_lists_render_identically():
    """Dropping the (tile, splat) pairs that cannot reach alpha >= 1/255 must not change a single bit of the image,
    depth, accumulated alpha or pixel counts; gradients agree up to the order of the atomic sums."""
    scene = _small_scene(91, npts=4000, sigma=0.08)
    out = {}
    for keep_all in (False, True):
        sc = cuda_scene(scene, requires_grad=True)
        op = make_op(keep_all_tile_pairs=keep_all)
        image, depth, count = run_forward(op, sc)
        image.backward(torch.ones_like(image))
        out[keep_all] = (image.detach().clone(), depth.clone(), count.clone(), sc.point_cloud.grad.clone(),
                         sc.point_cloud_features.grad.clone(), op.last_frame.num_keys)
    a, b = out[False], out[True]
    assert a[5] < b[5]
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])
    assert torch.allclose(a[3], b[3], rtol=1e-4, atol=1e-5 * float(b[3].abs().max()))
    assert torch.allclose(a[4], b[4], rtol=1e-4, atol=1e-5 * float(b[4].abs().max()))


def test_render_views_with_two_frames_in_flight_equals_sequential_rendering():
    """parallel.render_views(..., streams=[s0, s1]): consecutive frames on alternating streams (each frame owns its
    workspace and outputs) must give exactly the images of one-after-the-other rendering."""
    from taichi_3d_gaussian_splatting_b200.parallel import render_views
    scenes = [cuda_scene(_small_scene(95, npts=3000, yaw=float(y))) for y in (-6, -2, 2, 6, 9)]
    base = scenes[0]
    for sc in scenes[1:]:
        sc.point_cloud, sc.point_cloud_features = base.point_cloud, base.point_cloud_features
    op = make_op()

    def make_input(i):
        sc = scenes[i]
        from gpu_helpers import Input
        return Input(point_cloud=sc.point_cloud, point_cloud_features=sc.point_cloud_features, point_object_id=sc.point_object_id,
                     point_invalid_mask=sc.point_invalid_mask, camera_info=sc.camera_info, q_pointcloud_camera=sc.q_pointcloud_camera,
                     t_pointcloud_camera=sc.t_pointcloud_camera, color_max_sh_band=3)
    seq = render_views(op, make_input, range(len(scenes)))
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    for st in streams:
        st.wait_stream(torch.cuda.current_stream())
    par = render_views(op, make_input, range(len(scenes)), streams=streams)
    torch.cuda.synchronize()
    # (every forward re-normalises the q of the rows it sees in place, GPCR:264-266; normalising a unit vector again may move
    #  it by an ulp, so the two passes are compared to float rounding rather than bit for bit)
    for i in range(len(scenes)):
        assert torch.allclose(seq[i][0], par[i][0], atol=1e-5), i
        assert torch.allclose(seq[i][1], par[i][1], atol=1e-3), i
        assert int((seq[i][2] != par[i][2]).sum()) <= 2, i
