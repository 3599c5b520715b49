"""EXPERIMENTAL blend backward (``backward_impl="transposed"``, csrc/blend_bwd_transposed.cu) on the GPU.

The kernel's logic is verified on the CPU (tests/test_simt_blend_cpu.py: the unmodified kernel source under a
lock-step SIMT emulator, against the oracle and against the default kernel).  It was written after this round's GPU
minutes were spent, it is NOT the default path and no measured number depends on it; these tests are therefore marked
``xfail(strict=False)`` until their first B200 run has been seen (they are expected to XPASS), and the module sorts
after every other GPU module."""
import numpy as np
import pytest
import torch

from gpu_helpers import Config, cuda_scene, n, run_forward
from helpers import grad_close, oracle_backward, oracle_forward
from taichi_3d_gaussian_splatting_b200 import GaussianPointCloudRasterisation as GPCR
from taichi_3d_gaussian_splatting_b200.synthetic import make_scene

pytestmark = [pytest.mark.gpu, pytest.mark.xfail(strict=False, reason="experimental kernel, first GPU run pending")]


def _run(scene, impl, exact, hook_store=None, band=3):
    sc = cuda_scene(scene, requires_grad=True)
    hook = (lambda h: hook_store.__setitem__("h", h)) if hook_store is not None else None
    op = GPCR(Config(), backward_valid_point_hook=hook, exact_exp=exact, backward_impl=impl)
    image, _, _ = run_forward(op, sc, band=band)
    g = torch.Generator().manual_seed(5)
    grad_image = torch.randn(image.shape, generator=g, dtype=torch.float32)
    image.backward(grad_image.cuda())
    return n(image), n(sc.point_cloud.grad), n(sc.point_cloud_features.grad), grad_image.numpy()


SCENES = [dict(num_points=4000, height=64, width=96, sigma_med=0.05, seed=11),
          dict(num_points=60000, height=128, width=128, sigma_med=0.04, seed=12),   # > 256-entry tile lists, saturation
          dict(num_points=1500, height=128, width=192, sigma_med=0.4, seed=13)]     # large splats


@pytest.mark.parametrize("scene_args", SCENES)
@pytest.mark.parametrize("exact", [True, False])
def test_transposed_backward_matches_the_default_backward(scene_args, exact):
    scene = make_scene(sh_degree=3, **scene_args)
    ref_store, got_store = {}, {}
    img_a, gx_a, gf_a, _ = _run(scene, "butterfly", exact, ref_store)
    img_b, gx_b, gf_b, _ = _run(scene, "transposed", exact, got_store)
    assert np.array_equal(img_a, img_b)
    assert grad_close(gx_b, gx_a, rtol=1e-4, floor_frac=2e-6)[0], grad_close(gx_b, gx_a, rtol=1e-4, floor_frac=2e-6)
    for sl in (slice(0, 4), slice(4, 7), slice(7, 8), slice(8, 56)):
        assert grad_close(gf_b[:, sl], gf_a[:, sl], rtol=1e-4, floor_frac=2e-6)[0], sl
    a, b = ref_store["h"], got_store["h"]
    assert torch.equal(a.num_affected_pixels, b.num_affected_pixels)
    assert torch.equal(a.magnitude_grad_viewspace_on_image, b.magnitude_grad_viewspace_on_image)
    assert grad_close(n(b.magnitude_grad_viewspace), n(a.magnitude_grad_viewspace), rtol=1e-4, floor_frac=2e-6)[0]
    assert grad_close(n(b.grad_viewspace), n(a.grad_viewspace), rtol=1e-4, floor_frac=2e-6)[0]
    # without a hook the statistics are skipped; the gradients do not change
    _, gx_c, gf_c, _ = _run(scene, "transposed", exact, None)
    assert grad_close(gx_c, gx_b, rtol=1e-5, floor_frac=1e-6)[0] and grad_close(gf_c, gf_b, rtol=1e-5, floor_frac=1e-6)[0]


def test_transposed_backward_vs_oracle():
    scene = make_scene(num_points=4000, height=64, width=96, sigma_med=0.05, seed=11, sh_degree=3)
    o, fwd, feats_n = oracle_forward(scene)
    _, gx, gf, grad_image = _run(scene, "transposed", True, None)
    bwd = oracle_backward(o, fwd, scene, feats_n, grad_image, 3)
    assert grad_close(gx, bwd.grad_pointcloud)[0], grad_close(gx, bwd.grad_pointcloud)
    for sl in (slice(0, 4), slice(4, 7), slice(7, 8), slice(8, 56)):
        assert grad_close(gf[:, sl], bwd.grad_pointcloud_features[:, sl])[0], sl
