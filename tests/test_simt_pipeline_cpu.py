"""The whole CUDA path on the CPU: preprocess -> (host stable sort + tile ranges) -> forward blend -> loop A of the
backward (default and experimental kernel) -> per-point chain rule, every kernel the UNMODIFIED CUDA source running under
the lock-step SIMT emulator of ``tests/simt``, chained through the same buffers the library uses (packed records,
in-camera offsets, accumulator rows) and compared with the oracle's image and final dense gradients under the path's
criteria (forward <= 1e-4, gradients 1e-3 relative).  No stage is replaced: the one-sweep radix sort and the tile-range
kernel run from ``csrc/sort.cu`` as well (its TMA bulk copy becomes a memcpy under the emulator) and are also checked
against ``numpy.argsort(kind="stable")`` on their own.  Test infrastructure, not a product path."""
import ctypes

import numpy as np
import pytest

from helpers import grad_close, oracle_backward, oracle_forward
from taichi_3d_gaussian_splatting_b200.synthetic import make_scene
from simt_helpers import c, emu_sort, emulated_operator
from simt_helpers import build_emulator
from test_simt_preprocess_cpu import _large_splats


@pytest.fixture(scope="module")
def emu():
    return build_emulator()


@pytest.mark.parametrize("dtype,end_bit", [(np.uint32, 30), (np.uint32, 13), (np.uint64, 45), (np.uint64, 64)])
@pytest.mark.parametrize("n", [1, 37, 3071, 3072, 3073, 10000])
def test_emulated_radix_sort_is_the_stable_sort(emu, dtype, end_bit, n):
    """csrc/sort.cu: tiles of 3072 keys (exact multiple, one short, one over), many ties, partial last TMA granule."""
    rng = np.random.default_rng(n + end_bit)
    live = min(end_bit, 12 if n > 100 else 3)  # few distinct keys -> long runs of equal keys: stability matters
    keys = (rng.integers(0, 1 << live, n, dtype=np.uint64) << np.uint64(end_bit - live)).astype(dtype)
    if end_bit < 8 * keys.dtype.itemsize:  # bits above end_bit are not sorted: leave them zero like the frame keys
        assert int(keys.max()) < (1 << end_bit)
    vals = rng.permutation(n).astype(np.int32)
    ko, vo = emu_sort(emu, keys, vals, end_bit)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(ko, keys[order]) and np.array_equal(vo, vals[order])


@pytest.mark.parametrize("dtype,depth_bits,tile_bits", [(np.uint32, 17, 13), (np.uint32, 15, 12), (np.uint32, 10, 3), (np.uint64, 32, 13)])
@pytest.mark.parametrize("live", [0, 1, 3, 7, 8, 9, 10, 16, "all"])
def test_emulated_radix_sort_compacts_the_dead_depth_bits(emu, dtype, depth_bits, tile_bits, live):
    """csrc/sort.cu live-bit compaction: keys tile << depth_bits | depth where only the low `live` depth bits are used (the
    frame's largest depth key is handed over like the per-point kernel leaves it).  The digits of the compacted key are cut
    out of the stored key -- including digits straddling the depth / tile boundary -- the launches beyond the last needed pass
    exit, and the result is the stable sort of the stored keys whatever the number of passes that ran."""
    live = min(depth_bits, 31) if live == "all" else live  # depth keys are non-negative int32 (31 bits in the 64-bit packing)
    if live > depth_bits:
        pytest.skip("more live bits than the depth field has")
    n = 7000
    rng = np.random.default_rng(depth_bits * 100 + tile_bits * 10 + live)
    depth = rng.integers(0, 1 << live, n, dtype=np.uint64) if live else np.zeros(n, np.uint64)
    if live:
        depth[rng.integers(0, n)] = (1 << live) - 1  # the maximum really has `live` bits
    tile = rng.integers(0, min(1 << tile_bits, 40), n, dtype=np.uint64) * np.uint64(max(((1 << tile_bits) - 1) // 39, 1))
    keys = ((tile << np.uint64(depth_bits)) | depth).astype(dtype)
    vals = rng.permutation(n).astype(np.int32)
    ko, vo = np.empty_like(keys), np.empty_like(vals)
    for max_key in ([int(depth.max())] if live < min(depth_bits, 31) else [int(depth.max()), None]):
        mk = np.array([max_key], np.int32) if max_key is not None else None
        sw = emu.emu_sort_pairs_compacted(c(keys), c(vals), c(ko), c(vo), ctypes.c_longlong(n), keys.dtype.itemsize, depth_bits,
                                          depth_bits + tile_bits, c(mk) if mk is not None else None)
        assert sw > 0  # -1: the input buffer was written
        order = np.argsort(keys, kind="stable")
        assert np.array_equal(ko, keys[order]) and np.array_equal(vo, vals[order])


def test_emulated_radix_sort_many_ctas_with_compaction(emu):
    """40 000 keys = 14 CTAs per pass: the per-digit look-back walks several rounds of four predecessors, the last block of
    the histogram kernel turns the counts of 14 blocks into prefixes, and the three-buffer rotation ends in the output after
    the 3 active passes of a C3-like layout (13 tile bits, 17 depth bits of which 10 are live; 4 launches)."""
    n = 40000
    rng = np.random.default_rng(77)
    depth = rng.integers(0, 1 << 10, n, dtype=np.uint64)
    depth[123] = (1 << 10) - 1
    tile = rng.integers(0, 8040, n, dtype=np.uint64)
    keys = ((tile << np.uint64(17)) | depth).astype(np.uint32)
    vals = np.arange(n, dtype=np.int32)
    ko, vo = np.empty_like(keys), np.empty_like(vals)
    mk = np.array([int(depth.max())], np.int32)
    assert emu.emu_sort_pairs_compacted(c(keys), c(vals), c(ko), c(vo), ctypes.c_longlong(n), 4, 17, 30, c(mk)) > 0
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(ko, keys[order]) and np.array_equal(vo, vals[order])


def test_digit_selector_of_the_compacted_key_for_every_layout(emu):
    """csrc/sort.cu make_digit_sel: for EVERY (depth_bits, tile_bits, live depth bits, pass) the two shift-and-mask pairs cut
    exactly byte `pass` of the compacted key  tile << live | depth  out of the stored key  tile << depth_bits | depth, the
    number of active passes is ceil((tile_bits + live) / 8), and a pass beyond it is reported as inactive."""
    rng = np.random.default_rng(5)
    emu.emu_sort_digit.restype = ctypes.c_int
    checked = 0
    for key_bytes, layouts in ((4, [(d, t) for d in range(1, 25) for t in (1, 5, 8, 12, 13) if d + t <= 32]),
                               (8, [(32, t) for t in (1, 7, 13, 20, 32)] + [(20, 20), (31, 33)])):
        for depth_bits, tile_bits in layouts:
            for live in range(0, min(depth_bits, 31) + 1):
                end_bit = depth_bits + tile_bits
                passes = max((tile_bits + live + 7) // 8, 1)
                for _ in range(3):
                    depth = int(rng.integers(0, 1 << live)) if live else 0
                    tile = int(rng.integers(0, 1 << min(tile_bits, 62)))
                    stored = (tile << depth_bits) | depth
                    compact = (tile << live) | depth
                    for p in range(9):
                        got = emu.emu_sort_digit(ctypes.c_ulonglong(stored), key_bytes, p, depth_bits, end_bit, live)
                        want = (compact >> (8 * p)) & 255 if p < passes else -1
                        assert got == want, (key_bytes, depth_bits, tile_bits, live, p, hex(stored), got, want)
                        checked += 1
    assert checked > 20000


def test_emulated_tile_ranges_known_answer(emu):
    """The reference's own known answer (tests/GaussianPointCloudRasterisation_test.py:18-51 shape): keys tile << 32 | depth."""
    tiles = np.array([0, 0, 0, 2, 2, 5, 5, 5, 5, 7], np.uint64)
    keys = (tiles << np.uint64(32)) | np.arange(10, dtype=np.uint64)
    start, end = np.zeros(9, np.int32), np.zeros(9, np.int32)
    emu.emu_tile_ranges(c(keys), ctypes.c_longlong(10), 8, 32, 9, c(start), c(end))
    assert start.tolist() == [0, 0, 3, 0, 0, 5, 0, 9, 0] and end.tolist() == [3, 0, 5, 0, 0, 9, 0, 10, 0]


@pytest.mark.parametrize("transposed", [False, True])
@pytest.mark.parametrize("which,band", [("small", 3), ("small", 1), ("large", 3)])
def test_emulated_cuda_path_reproduces_the_oracle_end_to_end(emu, which, band, transposed):
    if which == "small":
        scene = make_scene(700, 48, 64, 0.06, 41, sh_degree=3, yaw_degrees=3.0)
        scene.point_cloud_features[:, 0:4] *= 0.6
    else:
        scene = _large_splats(0.3)
    o, fwd, feats_n = oracle_forward(scene)
    g = np.random.default_rng(77).standard_normal(fwd.image.shape).astype(np.float32)
    out = emulated_operator(emu, scene, g, band=band, transposed=transposed)
    gx, gf = out.grad_pointcloud, out.grad_pointcloud_features
    assert np.abs(out.image - fwd.image).max() <= 1e-4
    assert int((out.count != fwd.pixel_valid_point_count).sum()) == 0
    bwd = oracle_backward(o, fwd, scene, feats_n, g, band)
    kw = {}  # the path's criterion (1e-3 relative + 1e-5 of the largest entry), also on the deep lists of the large scene
    assert grad_close(gx, bwd.grad_pointcloud, **kw)[0], grad_close(gx, bwd.grad_pointcloud, **kw)
    for sl in (slice(0, 4), slice(4, 7), slice(7, 8), slice(8, 56)):
        assert grad_close(gf[:, sl], bwd.grad_pointcloud_features[:, sl], **kw)[0], (sl, grad_close(gf[:, sl], bwd.grad_pointcloud_features[:, sl], **kw))
