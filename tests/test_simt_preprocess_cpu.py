"""The fused per-point stage of the CUDA path (``csrc/preprocess.cu``: pose kernel, frustum filter, compaction by a
single-pass decoupled look-back scan, projection / conic / SH colour, tile counts, warp-cooperative reach filter and key
emission) executed on the CPU under the lock-step SIMT emulator of ``tests/simt`` -- the UNMODIFIED kernel source compiled
as host C++ -- and held to the oracle exactly like the GPU stage checks (``test_gpu_parity._check_stages``): ids, tile
counts, every float of the packed records and the in-place normalised quaternions bit for bit; with the reach filter off
the stably sorted keys / offsets equal the reference's lists, with it on every tile's list is a subsequence of the
reference's and every dropped pair is dead on all 256 pixels.  The large-splat scenes (17..96 tiles per splat: the
more-than-64-tiles branch of the cooperative filter) are the ones whose GPU test was written after the round's GPU minutes
were spent.  Test infrastructure, not a CPU path of the product."""
import ctypes
import os
import subprocess
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from helpers import oracle_forward
from taichi_3d_gaussian_splatting_b200.synthetic import make_scene
from test_gpu_parity import _check_filtered_lists

HERE = os.path.dirname(os.path.abspath(__file__))
SIMT = os.path.join(HERE, "simt")
CSRC = os.path.join(os.path.dirname(HERE), "taichi_3d_gaussian_splatting_b200", "csrc")


def build_emulator():
    out = os.path.join(SIMT, "libsimt_emu.so")
    tus = [os.path.join(SIMT, f) for f in ("emu_blend.cpp", "emu_preprocess.cpp", "emu_sort.cpp", "emu_image_loss.cpp", "emu_adam.cpp", "emu_controller.cpp")]
    deps = tus + [os.path.join(SIMT, "simt_emu.h")] + [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(s) for s in deps):
        cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
        subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-I", cuda_inc, "-o", out, *tus],
                       check=True)
    L = ctypes.CDLL(out)
    L.emu_blend_backward.restype = ctypes.c_longlong
    L.emu_blend_forward.restype = ctypes.c_longlong
    L.emu_preprocess.restype = ctypes.c_longlong
    L.emu_backward_points.restype = ctypes.c_longlong
    L.emu_sort_pairs.restype = ctypes.c_longlong
    L.emu_image_loss.restype = ctypes.c_longlong
    L.emu_image_loss_temp_bytes.restype = ctypes.c_longlong
    return L


@pytest.fixture(scope="module")
def emu():
    return build_emulator()


def _bit_width(v):
    return int(v).bit_length()


def _run(emu, scene, fwd_cfg, key64, filter_tiles):
    xyz = scene.point_cloud.numpy().astype(np.float32).copy()
    feats = scene.point_cloud_features.detach().numpy().astype(np.float32).copy()
    N = xyz.shape[0]
    ci = scene.camera_info
    H, W = ci.camera_height, ci.camera_width
    far, scale, near = fwd_cfg.get("far_plane", 1000.0), fwd_cfg.get("depth_to_sort_key_scale", 100.0), fwd_cfg.get("near_plane", 0.8)
    T = (H // 16) * (W // 16)
    tile_bits = _bit_width(max(T - 1, 0))
    mk = np.float32(far) * np.float32(scale)
    depth_bits = max(_bit_width(int(mk)), 1)  # csrc/api.cu compute_layout
    if key64 or tile_bits + depth_bits > 32:
        key_bytes, depth_bits = 8, 32
    else:
        key_bytes = 4
    cap = 64 * N + 4096
    counters = np.zeros(8, np.int64)
    point_id, point_offset, num_tiles = (np.full(N, -9, np.int32) for _ in range(3))
    records, pic = np.zeros((N, 12), np.float32), np.zeros((N, 3), np.float32)
    keys = np.zeros(cap, np.uint32 if key_bytes == 4 else np.uint64)
    vals = np.zeros(cap, np.int32)
    c = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    q = scene.q_pointcloud_camera.numpy().astype(np.float32).copy()
    t = scene.t_pointcloud_camera.numpy().astype(np.float32).copy()
    K = ci.camera_intrinsics.numpy().astype(np.float32).copy()
    inv = scene.point_invalid_mask.numpy().astype(np.int8).copy()
    obj = scene.point_object_id.numpy().astype(np.int32).copy()
    sw = emu.emu_preprocess(
        ctypes.c_longlong(N), c(xyz), c(feats), c(inv), c(obj), q.shape[0], c(q), c(t), c(K), W, H, ctypes.c_float(near),
        ctypes.c_float(far), ctypes.c_float(scale), depth_bits, key_bytes, int(filter_tiles), 0, ctypes.c_longlong(cap),
        c(counters), c(point_id), c(point_offset), c(num_tiles), c(records), c(pic), c(keys), c(vals))
    assert sw > 0
    return SimpleNamespace(feats=feats, counters=counters, point_id=point_id, point_offset=point_offset, num_tiles=num_tiles,
                           records=records, pic=pic, keys=keys, vals=vals, depth_bits=depth_bits, tile_bits=tile_bits, H=H, W=W, T=T)


def _check(out, fwd, feats_n, filter_tiles):
    M = int(out.counters[0])
    ids = fwd.point_id_in_camera_list
    assert M == ids.shape[0] and np.array_equal(out.point_id[:M], ids)
    assert np.array_equal(out.num_tiles[:M], fwd.num_overlap_tiles)
    inverse = np.full(out.point_offset.shape[0], -1, np.int32)
    inverse[ids] = np.arange(M, dtype=np.int32)
    assert np.array_equal(out.point_offset, inverse)
    r = out.records[:M]
    for got, exp in ((r[:, 0:2], fwd.point_uv), (r[:, 2:6], fwd.point_uv_conic_and_rescale),
                     (r[:, 6], fwd.point_alpha_after_activation), (r[:, 7], fwd.point_in_camera[:, 2]),
                     (r[:, 8:11], fwd.point_color), (r[:, 11], fwd.point_radii), (out.pic[:M], fwd.point_in_camera)):
        assert np.array_equal(got, exp), float(np.abs(got - exp).max())
    assert np.array_equal(out.feats, feats_n)  # in-place quaternion normalisation, GPCR:264-266
    Kk = int(out.counters[1])
    assert out.counters[2] == 0
    order = np.argsort(out.keys[:Kk], kind="stable")  # the device sort is a stable LSD radix sort
    sk, sv = out.keys[:Kk][order].astype(np.int64), out.vals[:Kk][order]
    okeys = fwd.point_in_camera_sort_key
    packed = ((okeys >> 32) << out.depth_bits) | (okeys & 0xFFFFFFFF)
    if not filter_tiles:
        assert np.array_equal(sk, packed) and np.array_equal(sv, fwd.point_offset_with_sort_key)
        return 0
    tile = sk >> out.depth_bits
    start = np.searchsorted(tile, np.arange(out.T), side="left").astype(np.int32)
    end = np.searchsorted(tile, np.arange(out.T), side="right").astype(np.int32)
    frame = SimpleNamespace(sorted_keys=torch.from_numpy(sk), point_offset_with_sort_key=torch.from_numpy(sv),
                            tile_points_start=torch.from_numpy(start), tile_points_end=torch.from_numpy(end),
                            width=out.W, height=out.H, num_keys=Kk)
    return _check_filtered_lists(frame, fwd, packed)


def _large_splats(sigma):
    scene = make_scene(1500, 128, 192, sigma, 31, sh_degree=3, yaw_degrees=4.0)  # tests/test_gpu_zz_large_splats.py
    scene.point_cloud_features[:, 7] -= 2.0
    return scene


@pytest.mark.parametrize("filter_tiles", [True, False])
@pytest.mark.parametrize("sigma", [0.3, 0.6])
def test_preprocess_source_on_large_splats(emu, sigma, filter_tiles):
    scene = _large_splats(sigma)
    o, fwd, feats_n = oracle_forward(scene)
    assert int((fwd.num_overlap_tiles > 64).sum()) >= 20 and int(fwd.num_overlap_tiles.max()) == 96
    out = _run(emu, scene, {}, key64=False, filter_tiles=filter_tiles)
    dropped = _check(out, fwd, feats_n, filter_tiles)
    if filter_tiles:
        assert dropped > 0


@pytest.mark.parametrize("key64", [False, True])
@pytest.mark.parametrize("filter_tiles", [True, False])
def test_preprocess_source_on_a_small_scene(emu, key64, filter_tiles):
    """Rotated camera, un-normalised quaternions, invalid slots, coarse depth keys (ties), three scan CTAs."""
    scene = make_scene(300, 64, 96, 0.08, 21, sh_degree=3, yaw_degrees=-7.0)
    scene.point_cloud_features[:, 0:4] *= 1.7
    scene.point_invalid_mask[::7] = 1
    cfg = dict(depth_to_sort_key_scale=10.0, near_plane=0.4)
    o, fwd, feats_n = oracle_forward(scene, **cfg)
    out = _run(emu, scene, cfg, key64=key64, filter_tiles=filter_tiles)
    _check(out, fwd, feats_n, filter_tiles)
