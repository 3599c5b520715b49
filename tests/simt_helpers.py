"""The whole CUDA path executed on the CPU from the unmodified kernel sources (tests/simt, see simt_emu.h): the stages are
chained exactly as ``csrc/api.cu`` chains them -- preprocess -> radix sort -> tile ranges -> forward blend -> loop A of
the backward -> per-point chain rule -- through the library's own buffers.  Test infrastructure."""
import ctypes
from types import SimpleNamespace

import numpy as np

from test_simt_preprocess_cpu import _run as run_preprocess


def c(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def emu_sort(emu, keys, vals, end_bit):
    keys, vals = np.ascontiguousarray(keys), np.ascontiguousarray(vals, dtype=np.int32)
    ko, vo = np.empty_like(keys), np.empty_like(vals)
    if keys.shape[0]:
        assert emu.emu_sort_pairs(c(keys), c(vals), c(ko), c(vo), ctypes.c_longlong(keys.shape[0]), keys.dtype.itemsize, end_bit) > 0
    return ko, vo


def emulated_operator(emu, scene, grad_image, band=3, transposed=False, exact=True, cfg=None, filter_tiles=True,
                      factors=(1.0, 0.5, 20.0, 5.0, 1.0)):
    """Forward + backward of the CUDA path under the emulator.  ``scene``: CPU tensors with the operator's input fields;
    ``cfg``: near_plane / far_plane / depth_to_sort_key_scale; ``factors``: grad q, s, alpha, colour, high-order colour
    (GPCR:782-786).  Returns the outputs, the per-point stage tensors, the dense gradients and the hook tensors."""
    cfg = cfg or {}
    pre = run_preprocess(emu, scene, cfg, key64=False, filter_tiles=filter_tiles)
    M, Kk = int(pre.counters[0]), int(pre.counters[1])
    sk, sv = emu_sort(emu, pre.keys[:Kk], pre.vals[:Kk], pre.tile_bits + pre.depth_bits)
    order = np.argsort(pre.keys[:Kk], kind="stable")
    assert np.array_equal(sk, pre.keys[:Kk][order]) and np.array_equal(sv, pre.vals[:Kk][order])
    start, end = np.zeros(pre.T, np.int32), np.zeros(pre.T, np.int32)
    emu.emu_tile_ranges(c(sk), ctypes.c_longlong(Kk), sk.dtype.itemsize, pre.depth_bits, pre.T, c(start), c(end))
    tile = sk.astype(np.int64) >> pre.depth_bits
    assert np.array_equal(end - start, np.bincount(tile, minlength=pre.T)[:pre.T])
    H, W = pre.H, pre.W
    image, depth, acc = np.zeros((H, W, 3), np.float32), np.zeros((H, W), np.float32), np.zeros((H, W), np.float32)
    last, cnt = np.zeros((H, W), np.int32), np.zeros((H, W), np.int32)
    emu.emu_blend_forward(0, int(exact), H, W, c(start), c(end), c(sv), c(pre.records), c(image), c(depth), c(acc), c(last), c(cnt))
    g = np.ascontiguousarray(grad_image, dtype=np.float32)
    accum, mag = np.zeros((max(M, 1), 12), np.float32), np.zeros((H, W, 2), np.float32)
    emu.emu_blend_backward(int(transposed), int(exact), 1, H, W, c(start), c(end), c(sv), c(pre.records), c(g), c(acc), c(last),
                           c(accum), c(mag))
    N = pre.point_offset.shape[0]
    q = scene.q_pointcloud_camera.numpy().astype(np.float32).copy()
    t = scene.t_pointcloud_camera.numpy().astype(np.float32).copy()
    poses = np.zeros((q.shape[0], 20), np.float32)
    emu.emu_pose(q.shape[0], c(q), c(t), c(poses))
    xyz = scene.point_cloud.detach().numpy().astype(np.float32).copy()
    K = scene.camera_info.camera_intrinsics.numpy().astype(np.float32).copy()
    obj = scene.point_object_id.numpy().astype(np.int32).copy()
    gx, gf = np.full((N, 3), 7.0, np.float32), np.full((N, 56), 7.0, np.float32)  # every row must be overwritten
    f = ctypes.c_float
    emu.emu_backward_points(ctypes.c_longlong(N), c(pre.point_offset), c(pre.records), c(pre.pic), c(accum), c(poses), c(xyz),
                            c(pre.feats), c(obj), c(t), c(K), int(band) if band in (0, 1, 2) else 3, *(f(v) for v in factors),
                            c(gx), c(gf))
    ids = pre.point_id[:M]
    r = pre.records[:M]
    return SimpleNamespace(
        image=image, depth=depth, count=cnt, acc_alpha=acc, last_effective=last, features_after_forward=pre.feats,
        point_id_in_camera_list=ids, num_overlap_tiles=pre.num_tiles[:M], point_uv=r[:, 0:2].copy(),
        point_uv_conic_and_rescale=r[:, 2:6].copy(), point_alpha_after_activation=r[:, 6].copy(), point_color=r[:, 8:11].copy(),
        point_radii=r[:, 11].copy(), point_in_camera=pre.pic[:M], grad_pointcloud=gx, grad_pointcloud_features=gf,
        hook=SimpleNamespace(grad_point_in_camera=gx[ids], grad_pointfeatures_in_camera=gf[ids], grad_viewspace=accum[:M, 0:2],
                             magnitude_grad_viewspace=accum[:M, 9], magnitude_grad_viewspace_on_image=mag,
                             num_affected_pixels=np.round(accum[:M, 10]).astype(np.int32)))
