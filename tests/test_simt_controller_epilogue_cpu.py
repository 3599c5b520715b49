"""The densification controller's accumulator update fused into the epilogue of the per-point backward kernel
(``GsbBackwardArgs.ctl_*``, csrc/blend_bwd.cu) under the SIMT emulator, against ``GaussianPointAdaptiveController.update`` fed
with the hook tensors of the same backward (GaussianPointAdaptiveController.py:130-143)."""
import numpy as np
import pytest
import torch

from taichi_3d_gaussian_splatting_b200 import GaussianPointCloudRasterisation as GPCR
from taichi_3d_gaussian_splatting_b200.densification import GaussianPointAdaptiveController as Controller
from taichi_3d_gaussian_splatting_b200.synthetic import make_scene
from simt_helpers import build_emulator, emulated_backward, emulated_forward


@pytest.fixture(scope="module")
def emu():
    return build_emulator()


def test_fused_controller_epilogue_equals_the_hook_update(emu):
    scene = make_scene(600, 48, 64, 0.07, 23, sh_degree=3, yaw_degrees=2.0)
    scene.point_invalid_mask[::9] = 1
    N = scene.point_cloud.shape[0]
    ctl = Controller(Controller.GaussianPointAdaptiveControllerConfig(num_iterations_warm_up=10 ** 6),
                     Controller.GaussianPointAdaptiveControllerMaintainedParameters(
                         pointcloud=scene.point_cloud, pointcloud_features=scene.point_cloud_features,
                         point_invalid_mask=scene.point_invalid_mask, point_object_id=scene.point_object_id))
    fused = [np.zeros(N, np.int32), np.zeros(N, np.int32), np.zeros(N, np.float32), np.zeros(N, np.float32),
             np.zeros((N, 3), np.float32), np.zeros(N, np.float32)]
    for it in range(3):  # three frames accumulate
        sc = make_scene(600, 48, 64, 0.07, 23, sh_degree=3, yaw_degrees=2.0 + 3.0 * it)
        sc.point_invalid_mask[::9] = 1
        st = emulated_forward(emu, sc, exact=True)
        g = np.random.default_rng(it).standard_normal(st.image.shape).astype(np.float32)
        gx, gf, h = emulated_backward(emu, st, g, band=3, transposed=True, controller=fused)
        M = st.M
        ctl.update(GPCR.BackwardValidPointHookInput(
            point_id_in_camera_list=torch.from_numpy(st.pre.point_id[:M].copy()), grad_point_in_camera=torch.from_numpy(h.grad_point_in_camera),
            grad_pointfeatures_in_camera=torch.from_numpy(h.grad_pointfeatures_in_camera), grad_viewspace=torch.from_numpy(h.grad_viewspace),
            magnitude_grad_viewspace=torch.from_numpy(h.magnitude_grad_viewspace),
            magnitude_grad_viewspace_on_image=torch.from_numpy(h.magnitude_grad_viewspace_on_image),
            num_overlap_tiles=torch.from_numpy(st.pre.num_tiles[:M].copy()), num_affected_pixels=torch.from_numpy(h.num_affected_pixels),
            point_depth=torch.from_numpy(st.pre.pic[:M, 2].copy()), point_uv_in_camera=torch.from_numpy(st.pre.records[:M, 0:2].copy())))
    assert fused[0].max() == 3 and fused[1].max() > 0
    assert np.array_equal(fused[0], ctl.accumulated_num_in_camera.numpy())
    assert np.array_equal(fused[1], ctl.accumulated_num_pixels.numpy())
    for got, exp in ((fused[2], ctl.accumulated_view_space_position_gradients), (fused[3], ctl.accumulated_view_space_position_gradients_avg),
                     (fused[4], ctl.accumulated_position_gradients), (fused[5], ctl.accumulated_position_gradients_norm)):
        exp = exp.numpy()
        assert np.allclose(got, exp, rtol=2e-6, atol=1e-7 * float(np.abs(exp).max())), float(np.abs(got - exp).max())
