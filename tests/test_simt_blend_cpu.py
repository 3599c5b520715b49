"""The blend kernels -- forward, and both CUDA implementations of loop A of the backward -- executed on the CPU under a
lock-step SIMT emulator.

``tests/simt/emu_blend.cpp`` compiles ``csrc/blend_bwd.cu`` (butterfly reduction per (warp, splat): the default,
verified on the GPU) and ``csrc/blend_bwd_transposed.cu`` (experimental: splat-per-lane accumulation after a
shared-memory transposition, GSB_FLAG_BACKWARD_TRANSPOSED) UNMODIFIED as host C++; ``tests/simt/simt_emu.h`` runs the
256 threads of a CTA as fibres that meet in the warp / block collectives.  The emulated kernels are fed the oracle's
forward state and their per-splat accumulator rows are compared with the oracle's backward (``gso_rasterisation_
backward_pixels``, GPCR:531-705) and with each other.  This checks the kernels' LOGIC (list construction, culling,
recursions, the 16-splat chunk transposition, reduction slots, row addressing); it says nothing about speed and it is
test infrastructure, not a CPU path of the product."""
import ctypes
import os

import numpy as np
import pytest

from helpers import oracle_forward
from oracle.gs_oracle import lib as oracle_lib, _p
from taichi_3d_gaussian_splatting_b200.synthetic import make_scene
from simt_helpers import build_emulator

HERE = os.path.dirname(os.path.abspath(__file__))
SIMT = os.path.join(HERE, "simt")
CSRC = os.path.join(os.path.dirname(HERE), "taichi_3d_gaussian_splatting_b200", "csrc")


@pytest.fixture(scope="module")
def emu():
    return build_emulator()


def _state(num_points, H, W, sigma, seed, **cfg):
    scene = make_scene(num_points, H, W, sigma, seed, sh_degree=3)
    o, fwd, feats = oracle_forward(scene, **cfg)
    M = fwd.point_id_in_camera_list.shape[0]
    rec = np.zeros((M, 12), np.float32)  # u v a b | c rescale opacity depth | r g b radius  
    rec[:, 0:2] = fwd.point_uv
    rec[:, 2:6] = fwd.point_uv_conic_and_rescale
    rec[:, 6] = fwd.point_alpha_after_activation
    rec[:, 7] = fwd.point_in_camera[:, 2]
    rec[:, 8:11] = fwd.point_color
    rec[:, 11] = fwd.point_radii
    g = np.random.default_rng(seed + 1).standard_normal((H, W, 3)).astype(np.float32)
    return fwd, rec, g


def _oracle_rows(fwd, g):
    """The 11 accumulator columns from the oracle's loop A (double accumulation, GPCR:531-705)."""
    H, W = g.shape[:2]
    ids = fwd.point_id_in_camera_list
    M, N = ids.shape[0], int(ids.max()) + 1
    grad_uv, cov, col = np.zeros((N, 2)), np.zeros((M, 3)), np.zeros((M, 3))
    logit, mag, npix = np.zeros(N), np.zeros(N), np.zeros(M, np.int32)
    mag_img = np.zeros((H, W, 2), np.float32)
    oracle_lib().gso_rasterisation_backward_pixels(
        ctypes.c_int(H), ctypes.c_int(W), _p(fwd.tile_points_start), _p(fwd.tile_points_end),
        _p(fwd.point_offset_with_sort_key), _p(ids), _p(g), _p(fwd.pixel_accumulated_alpha),
        _p(fwd.pixel_offset_of_last_effective_point), _p(fwd.point_uv), _p(fwd.point_uv_conic_and_rescale),
        _p(fwd.point_alpha_after_activation), _p(fwd.point_color), ctypes.c_int64(N), ctypes.c_int64(M),
        _p(grad_uv), _p(cov), _p(col), _p(logit), _p(mag), _p(npix), _p(mag_img))
    rows = np.zeros((M, 11))
    rows[:, 0:2] = grad_uv[ids]
    rows[:, 2:5] = 2.0 * cov  # the kernels defer the 1/2 of UT:345 to the per-point epilogue
    rows[:, 5:8] = col
    rows[:, 8] = logit[ids]
    rows[:, 9] = mag[ids]
    rows[:, 10] = npix
    return rows, mag_img


def _run(emu, fwd, rec, g, transposed, exact, stats):
    H, W = g.shape[:2]
    M = rec.shape[0]
    accum = np.zeros((M, 12), np.float32)
    mag_img = np.full((H, W, 2), -1.0, np.float32)
    c = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    vals = np.ascontiguousarray(fwd.point_offset_with_sort_key, dtype=np.int32)
    switches = emu.emu_blend_backward(
        int(transposed), int(exact), int(stats), H, W, c(np.ascontiguousarray(fwd.tile_points_start, dtype=np.int32)),
        c(np.ascontiguousarray(fwd.tile_points_end, dtype=np.int32)), c(vals), c(rec), c(g),
        c(np.ascontiguousarray(fwd.pixel_accumulated_alpha, dtype=np.float32)),
        c(np.ascontiguousarray(fwd.pixel_offset_of_last_effective_point, dtype=np.int32)), c(accum), c(mag_img))
    assert switches > 0
    return accum, mag_img


def _close(got, exp, rtol, floor):
    exp = np.asarray(exp, np.float64)
    tol = rtol * np.abs(exp) + floor * max(np.abs(exp).max(), 1e-30)
    bad = np.abs(got - exp) > tol
    return not bad.any(), int(bad.sum()), float((np.abs(got - exp) / tol).max())


# (points, H, W, sigma, seed): a sparse frame, a frame with several 256-splat batches per tile (lists > 256 entries,
# saturation, skip masks), and a frame of large splats
SCENES = [(600, 32, 48, 0.05, 3), (9000, 32, 32, 0.06, 4), (300, 48, 32, 0.4, 5)]


@pytest.mark.parametrize("scene", SCENES)
@pytest.mark.parametrize("transposed", [False, True])
def test_emulated_kernels_reproduce_the_oracle_rows(emu, scene, transposed):
    fwd, rec, g = _state(*scene)
    rows, mag_img = _oracle_rows(fwd, g)
    assert rows[:, 10].max() > 0
    accum, got_img = _run(emu, fwd, rec, g, transposed, exact=True, stats=True)
    assert np.array_equal(accum[:, 10], rows[:, 10])  # affected-pixel counts: exact
    assert (accum[:, 11] == 0).all()
    for cols in (slice(0, 2), slice(2, 5), slice(5, 8), slice(8, 9), slice(9, 10)):
        ok, nbad, worst = _close(accum[:, cols], rows[:, cols], 1e-3, 1e-5)  # the path's gradient criterion
        assert ok, (cols, nbad, worst)
    assert np.abs(got_img - mag_img).max() <= 1e-3 * max(1.0, float(np.abs(mag_img).max()))


@pytest.mark.parametrize("scene", SCENES)
@pytest.mark.parametrize("exact", [True, False])
def test_transposed_kernel_matches_the_butterfly_kernel(emu, scene, exact):
    fwd, rec, g = _state(*scene)
    ref, ref_img = _run(emu, fwd, rec, g, False, exact, True)
    got, got_img = _run(emu, fwd, rec, g, True, exact, True)
    assert np.array_equal(got[:, 10], ref[:, 10])
    for cols in (slice(0, 2), slice(2, 5), slice(5, 8), slice(8, 9), slice(9, 10)):
        ok, nbad, worst = _close(got[:, cols], ref[:, cols], 1e-4, 2e-6)  # same arithmetic per pixel, other summation order
        assert ok, (cols, nbad, worst)
    # the per-pixel recursion is the same code in both kernels (exact path); on the fast path both take alpha from
    # fast_alpha (common.cuh) but the transposed kernel re-derives conic * d for the magnitude image from the pre-scaled conic
    if exact:
        assert np.array_equal(got_img, ref_img)
    else:
        assert np.allclose(got_img, ref_img, rtol=2e-6, atol=1e-7 * float(np.abs(ref_img).max()))
    # without the hook statistics: columns 0..8 as in the run with them (the warps interleave differently, so the float
    # atomics of splats shared by several patches land in another order), 9..10 and the magnitude image untouched
    lean, lean_img = _run(emu, fwd, rec, g, True, exact, False)
    ok, nbad, worst = _close(lean[:, :9], got[:, :9], 1e-6, 1e-7)
    assert ok, (nbad, worst)
    assert (lean[:, 9:] == 0).all() and (lean_img == -1.0).all()
    # the same switch in the butterfly kernel
    lean_ref, lean_ref_img = _run(emu, fwd, rec, g, False, exact, False)
    ok, nbad, worst = _close(lean_ref[:, :9], ref[:, :9], 1e-6, 1e-7)
    assert ok, (nbad, worst)
    assert (lean_ref[:, 9:] == 0).all() and (lean_ref_img == -1.0).all()


@pytest.mark.parametrize("scene", SCENES)
@pytest.mark.parametrize("exact", [True, False])
def test_emulated_forward_blend_reproduces_the_oracle(emu, scene, exact):
    """csrc/blend_fwd.cu (K6, GPCR:318-485) under the emulator: staging, per-patch culling lists, saturation exits."""
    fwd, rec, g = _state(*scene)
    H, W = g.shape[:2]
    image, depth = np.full((H, W, 3), -1.0, np.float32), np.full((H, W), -1.0, np.float32)
    acc, last, cnt = np.full((H, W), -1.0, np.float32), np.full((H, W), -7, np.int32), np.full((H, W), -7, np.int32)
    c = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    args = (H, W, c(np.ascontiguousarray(fwd.tile_points_start, dtype=np.int32)),
            c(np.ascontiguousarray(fwd.tile_points_end, dtype=np.int32)),
            c(np.ascontiguousarray(fwd.point_offset_with_sort_key, dtype=np.int32)), c(rec))
    assert emu.emu_blend_forward(0, int(exact), *args, c(image), c(depth), c(acc), c(last), c(cnt)) > 0
    tol = 2e-6 if exact else 1e-4  # exact: the oracle's operation order with libm expf; fast: exp2 of a folded exponent
    assert np.abs(image - fwd.image).max() <= tol
    assert np.abs(acc - fwd.pixel_accumulated_alpha).max() <= tol
    assert np.abs(depth - fwd.depth).max() <= 1e-3 * max(1.0, float(np.abs(fwd.depth).max()))
    flips = int((cnt != fwd.pixel_valid_point_count).sum())  # a pair within an ulp of the alpha = 1/255 cut-off may flip
    assert flips <= (0 if exact else 2)
    if flips == 0:
        assert np.array_equal(last, fwd.pixel_offset_of_last_effective_point)
    rgb = np.full((H, W, 3), -1.0, np.float32)
    assert emu.emu_blend_forward(1, int(exact), *args, c(rgb), c(depth), c(acc), c(last), c(cnt)) > 0
    assert np.array_equal(rgb, image)
