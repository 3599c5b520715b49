"""The whole CUDA path on the CPU: preprocess -> (host stable sort + tile ranges) -> forward blend -> loop A of the
backward (default and experimental kernel) -> per-point chain rule, every kernel the UNMODIFIED CUDA source running under
the lock-step SIMT emulator of ``tests/simt``, chained through the same buffers the library uses (packed records,
in-camera offsets, accumulator rows) and compared with the oracle's image and final dense gradients under the path's
criteria (forward <= 1e-4, gradients 1e-3 relative).  Only the radix sort (TMA bulk copies, ``match.any``) is replaced by
``numpy.argsort(kind="stable")``; it has its own bit-exact GPU tests.  Test infrastructure, not a product path."""
import ctypes

import numpy as np
import pytest

from helpers import grad_close, oracle_backward, oracle_forward
from taichi_3d_gaussian_splatting_b200.synthetic import make_scene
from test_simt_preprocess_cpu import _large_splats, _run as run_preprocess, build_emulator


@pytest.fixture(scope="module")
def emu():
    L = build_emulator()
    L.emu_backward_points.restype = ctypes.c_longlong
    return L


def _pipeline(emu, scene, transposed, exact, band):
    c = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    pre = run_preprocess(emu, scene, {}, key64=False, filter_tiles=True)
    M, Kk = int(pre.counters[0]), int(pre.counters[1])
    order = np.argsort(pre.keys[:Kk], kind="stable")
    sk, sv = pre.keys[:Kk][order].astype(np.int64), np.ascontiguousarray(pre.vals[:Kk][order])
    tile = sk >> pre.depth_bits
    start = np.searchsorted(tile, np.arange(pre.T), side="left").astype(np.int32)
    end = np.searchsorted(tile, np.arange(pre.T), side="right").astype(np.int32)
    H, W = pre.H, pre.W
    image, depth, acc = np.zeros((H, W, 3), np.float32), np.zeros((H, W), np.float32), np.zeros((H, W), np.float32)
    last, cnt = np.zeros((H, W), np.int32), np.zeros((H, W), np.int32)
    emu.emu_blend_forward(0, int(exact), H, W, c(start), c(end), c(sv), c(pre.records), c(image), c(depth), c(acc), c(last), c(cnt))
    g = np.random.default_rng(77).standard_normal((H, W, 3)).astype(np.float32)
    accum, mag = np.zeros((max(M, 1), 12), np.float32), np.zeros((H, W, 2), np.float32)
    emu.emu_blend_backward(int(transposed), int(exact), 1, H, W, c(start), c(end), c(sv), c(pre.records), c(g), c(acc), c(last),
                           c(accum), c(mag))
    N = pre.point_offset.shape[0]
    q = scene.q_pointcloud_camera.numpy().astype(np.float32).copy()
    t = scene.t_pointcloud_camera.numpy().astype(np.float32).copy()
    poses = np.zeros((q.shape[0], 20), np.float32)
    emu.emu_pose(q.shape[0], c(q), c(t), c(poses))
    xyz = scene.point_cloud.numpy().astype(np.float32).copy()
    K = scene.camera_info.camera_intrinsics.numpy().astype(np.float32).copy()
    obj = scene.point_object_id.numpy().astype(np.int32).copy()
    gx, gf = np.full((N, 3), 7.0, np.float32), np.full((N, 56), 7.0, np.float32)  # every row must be overwritten
    f = ctypes.c_float
    emu.emu_backward_points(ctypes.c_longlong(N), c(pre.point_offset), c(pre.records), c(pre.pic), c(accum), c(poses), c(xyz),
                            c(pre.feats), c(obj), c(t), c(K), band, f(1.0), f(0.5), f(20.0), f(5.0), f(1.0), c(gx), c(gf))
    return image, cnt, g, gx, gf


@pytest.mark.parametrize("transposed", [False, True])
@pytest.mark.parametrize("which,band", [("small", 3), ("small", 1), ("large", 3)])
def test_emulated_cuda_path_reproduces_the_oracle_end_to_end(emu, which, band, transposed):
    if which == "small":
        scene = make_scene(700, 48, 64, 0.06, 41, sh_degree=3, yaw_degrees=3.0)
        scene.point_cloud_features[:, 0:4] *= 0.6
    else:
        scene = _large_splats(0.3)
    o, fwd, feats_n = oracle_forward(scene)
    image, cnt, g, gx, gf = _pipeline(emu, scene, transposed, True, band)
    assert np.abs(image - fwd.image).max() <= 1e-4
    assert int((cnt != fwd.pixel_valid_point_count).sum()) == 0
    bwd = oracle_backward(o, fwd, scene, feats_n, g, band)
    loose = which == "large"  # deep lists: the criterion of tests/test_gpu_zz_large_splats.py
    kw = dict(rtol=2e-3, floor_frac=5e-5) if loose else {}
    assert grad_close(gx, bwd.grad_pointcloud, **kw)[0], grad_close(gx, bwd.grad_pointcloud, **kw)
    for sl in (slice(0, 4), slice(4, 7), slice(7, 8), slice(8, 56)):
        assert grad_close(gf[:, sl], bwd.grad_pointcloud_features[:, sl], **kw)[0], (sl, grad_close(gf[:, sl], bwd.grad_pointcloud_features[:, sl], **kw))
