"""The fused per-point stage of the CUDA path (``csrc/preprocess.cu``: pose kernel, frustum filter, compaction by a
single-pass decoupled look-back scan, projection / conic / SH colour, tile counts, warp-cooperative reach filter and key
emission) executed on the CPU under the lock-step SIMT emulator of ``tests/simt`` -- the UNMODIFIED kernel source compiled
as host C++ -- and held to the oracle exactly like the GPU stage checks (``test_gpu_parity._check_stages``): ids, tile
counts, every float of the packed records and the in-place normalised quaternions bit for bit; with the reach filter off
the stably sorted keys / offsets equal the reference's lists, with it on every tile's list is a subsequence of the
reference's and every dropped pair is dead on all 256 pixels.  The large-splat scenes (17..96 tiles per splat: the
more-than-64-tiles branch of the cooperative filter) are the ones whose GPU test was written after the round's GPU minutes
were spent.  Test infrastructure, not a CPU path of the product."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from helpers import oracle_forward
from taichi_3d_gaussian_splatting_b200.synthetic import make_scene
from simt_helpers import build_emulator, run_preprocess as _run
from test_gpu_parity import _check_filtered_lists



@pytest.fixture(scope="module")
def emu():
    return build_emulator()


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
