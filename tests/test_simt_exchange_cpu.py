"""The view-parallel exchange kernels on the CPU (kernel sources under the SIMT emulator): the COMPACT per-point kernel
(GSB_FLAG_COMPACT_GRADS) of R views, their columns summed / stacked as the collectives of parallel.py do, and
gsb200_expand_view_gradients must reproduce the SUM over the views of the dense gradients (GPCR:1102-1125, 1167-1182)."""
import ctypes

import numpy as np
import pytest
import torch

from taichi_3d_gaussian_splatting_b200.synthetic import make_scene
from simt_helpers import build_emulator, c, emulated_backward, emulated_forward


@pytest.fixture(scope="module")
def emu():
    return build_emulator()


def _views(num_views):
    """One scene (two objects with poses of their own), `num_views` cameras."""
    base = make_scene(500, 48, 64, 0.07, 17, sh_degree=3)
    base.point_object_id[1::3] = 1
    out = []
    for v in range(num_views):
        sc = make_scene(500, 48, 64, 0.07, 17, sh_degree=3, yaw_degrees=4.0 * v - 3.0)
        sc.point_object_id = base.point_object_id.clone()
        sc.point_invalid_mask[::7] = 1  # rows outside every frustum: zeros in every buffer
        q0 = sc.q_pointcloud_camera[0]
        q1 = torch.tensor([0.02 * v, float(q0[1]) - 0.03, 0.01, float(q0[3])])
        sc.q_pointcloud_camera = torch.stack([q0, q1 / q1.norm()])
        sc.t_pointcloud_camera = torch.tensor([[0.1 * v, 0.0, -0.2], [0.3, -0.1 * v, -0.5]])
        out.append(sc)
    return out


@pytest.mark.parametrize("band", [3, 1, 0])
@pytest.mark.parametrize("num_views", [1, 3])
def test_compact_rows_plus_expansion_equal_the_sum_of_the_dense_gradients(emu, num_views, band):
    views = _views(num_views)
    N = views[0].point_cloud.shape[0]
    n_obj = 2
    stride = 3 * N + 3 * n_obj + 5  # any stride >= 3N + 3 n_obj
    dense_x, dense_f = np.zeros((N, 3), np.float64), np.zeros((N, 56), np.float64)
    gsum = np.zeros((N, 12), np.float32)
    blocks = np.zeros((num_views, stride), np.float32)
    feats = None
    for v, sc in enumerate(views):
        if feats is not None:  # the forward normalises q in place: later views see the normalised rows, like a real trainer
            sc.point_cloud_features = feats
        st = emulated_forward(emu, sc, exact=True)
        feats = torch.from_numpy(st.pre.feats.copy())
        g = np.random.default_rng(5 + v).standard_normal(st.image.shape).astype(np.float32)
        gx, gf, _ = emulated_backward(emu, st, g, band=band, transposed=True)
        dense_x += gx
        dense_f += gf
        s12, c3, _ = emulated_backward(emu, st, g, band=band, transposed=True, compact=True)
        assert (s12[:, 11] == 0).all()
        outside = st.pre.point_offset < 0
        assert outside.any() and (s12[outside] == 0).all() and (c3[outside] == 0).all()
        # single view: the compact columns ARE the dense ones
        assert np.array_equal(s12[:, 0:3], gx) and np.array_equal(s12[:, 3:11], gf[:, 0:8])
        gsum += s12                                   # all-reduce(sum)
        blocks[v, :3 * N] = c3.reshape(-1)            # all-gather: colour-argument gradients ...
        blocks[v, 3 * N:3 * N + 3 * n_obj] = sc.t_pointcloud_camera.numpy().reshape(-1)  # ... and the camera centres
    xyz = views[0].point_cloud.numpy().astype(np.float32).copy()
    obj = views[0].point_object_id.numpy().astype(np.int32).copy()
    out_x, out_f = np.full((N, 3), 7.0, np.float32), np.full((N, 56), 7.0, np.float32)
    f = ctypes.c_float
    assert emu.emu_expand_view_gradients(ctypes.c_longlong(N), num_views, c(gsum), c(blocks), ctypes.c_longlong(stride), c(xyz),
                                         c(obj), band, f(5.0), f(1.0), c(out_x), c(out_f), 0) > 0
    # the two halves of the split expansion (SH columns from the blocks alone / summed columns from grad_sum alone) together
    # write exactly what the one-pass expansion writes, each leaving the other's piece untouched
    part_x, part_f = np.full((N, 3), 7.0, np.float32), np.full((N, 56), 7.0, np.float32)
    nan_sum, nan_blocks = np.full_like(gsum, np.nan), blocks.copy()
    assert emu.emu_expand_view_gradients(ctypes.c_longlong(N), num_views, c(nan_sum), c(blocks), ctypes.c_longlong(stride), c(xyz),
                                         c(obj), band, f(5.0), f(1.0), c(part_x), c(part_f), 1) > 0
    assert (part_x == 7.0).all() and (part_f[:, :8] == 7.0).all() and np.array_equal(part_f[:, 8:], out_f[:, 8:])
    nan_blocks[:, :3 * N] = np.nan
    assert emu.emu_expand_view_gradients(ctypes.c_longlong(N), num_views, c(gsum), c(nan_blocks), ctypes.c_longlong(stride), c(xyz),
                                         c(obj), band, f(5.0), f(1.0), c(part_x), c(part_f), 2) > 0
    assert np.array_equal(part_x, out_x) and np.array_equal(part_f, out_f)
    if num_views == 1:  # same operations in the same order as the dense kernel
        assert np.array_equal(out_x, dense_x.astype(np.float32)) and np.array_equal(out_f, dense_f.astype(np.float32))
    scale = np.abs(dense_f).max()
    assert np.abs(out_x - dense_x).max() <= 1e-6 * np.abs(dense_x).max()
    assert np.abs(out_f - dense_f).max() <= 1e-6 * scale
    cleared = {3: 16, 1: 4, 0: 1}[band]
    sh = out_f[:, 8:].reshape(N, 3, 16)
    assert (sh[:, :, cleared:] == 0).all() and np.abs(sh[:, :, :cleared]).max() > 0
