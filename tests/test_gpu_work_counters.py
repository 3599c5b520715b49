"""-m gpu: the device-side work counters of the blend kernels (gsb200_forward_blend_work / gsb200_backward_blend_work) and the
hardware self-test of the default arithmetic path."""
import pytest
import torch

from taichi_3d_gaussian_splatting_b200 import _lib, profiling
from taichi_3d_gaussian_splatting_b200.synthetic import CONFIGS, make_scene

from gpu_helpers import Input, cuda_scene, make_op, run_forward

pytestmark = pytest.mark.gpu


def test_device_selftest():
    """rcp.approx(1) == 1 and ex2.approx(0) == 1 exactly: the branch-free backward relies on the first (a pair that does not
    contribute multiplies the transmittance by rcp(1 - 0))."""
    lib = _lib.load()
    _lib.check(lib.gsb200_device_selftest(torch.cuda.current_stream().cuda_stream), "gsb200_device_selftest")


@pytest.mark.parametrize("name", ["C1", "C2"])
def test_work_counters_agree_with_the_rendered_frame(name):
    cfg = CONFIGS[name]
    scene = make_scene(**cfg)
    sc = cuda_scene(scene)
    op = make_op()
    image, depth, count = run_forward(op, sc, band=3)
    inp = Input(point_cloud=sc.point_cloud, point_cloud_features=sc.point_cloud_features, point_object_id=sc.point_object_id,
                point_invalid_mask=sc.point_invalid_mask, camera_info=sc.camera_info, q_pointcloud_camera=sc.q_pointcloud_camera,
                t_pointcloud_camera=sc.t_pointcloud_camera, color_max_sh_band=3)
    g = torch.randn(image.shape, device="cuda")
    w = profiling.blend_work(op, inp, g)
    blended = int(count.sum())
    pixels = count.numel()
    # forward: every blended pair, plus at most one saturating pair per pixel
    assert blended <= w["forward_contributing_evaluations"] <= blended + pixels
    assert w["forward_warp_splat_visits"] * 32 >= w["forward_contributing_evaluations"]
    # backward: exactly the blended pairs -- both passes take the alpha >= 1/255 decision on identical bits (fast_alpha)
    assert w["backward_contributing_evaluations"] == blended
    assert w["backward_warp_splat_visits"] <= w["forward_warp_splat_visits"]
    assert w["forward_warp_splat_visits"] <= 8 * op.last_frame.num_keys
    # what-if counters at staging time: every visit was staged; merging two patches never needs more (pair, splat) visits than
    # the 8x4 patches and never fewer than half of them
    p84, p88, p164 = w["staged_patch_pairs_8x4"], w["staged_patch_pairs_8x8"], w["staged_patch_pairs_16x4"]
    assert w["forward_warp_splat_visits"] <= p84 <= 8 * op.last_frame.num_keys
    assert p84 / 2 <= p88 <= p84 and p84 / 2 <= p164 <= p84
