"""Golden optimisation trajectory: the loop of the reference's tests/GaussianPointAdaptiveController_test.py::test_basic
(reference operator + reference controller as backward hook + torch Adam on xyz and features, squared error to a fake
32 x 32 image, SH band schedule) run with the reference's kernels under taichi_shim.py, shortened to 30 iterations so that
no densification happens (warm-up is 500): a deterministic trajectory of losses and parameters that depends on every
gradient the operator returns -- including the hard-wired gradient factors and the SH-band masking schedule.

    python tests/golden/make_training_golden.py        # build container only; writes training_vectors.json
"""
import contextlib
import io
import json
import os
import sys
import types
from unittest import mock

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import taichi_shim  # noqa: E402

NUM_POINTS, NUM_SLOTS, ITERATIONS, BAND_INTERVAL, LR = 300, 400, 30, 8, 0.001


def fake_image():
    img = torch.zeros((32, 32, 3))
    img[:5, :2, 0], img[:5, :2, 1] = 1.0, 0.7
    img[8:24, 8:24, 0], img[8:24, 8:24, 1] = 0.5, 0.7
    img[20:28, 20:28, 0], img[20:28, 20:28, 1] = 0.8, 0.1
    return img


def initial_parameters():
    g = torch.Generator().manual_seed(21)
    xyz = (torch.rand((NUM_SLOTS, 3), generator=g) - 0.5) * 3
    feat = torch.rand((NUM_SLOTS, 56), generator=g)
    feat[:, 4:7] = -3.0 + 0.5 * torch.rand((NUM_SLOTS, 3), generator=g)
    feat[:, 7] = 0.5
    mask = torch.zeros(NUM_SLOTS, dtype=torch.int8)
    mask[NUM_POINTS:] = 1
    return xyz, feat, mask, torch.zeros(NUM_SLOTS, dtype=torch.int32)


CAMERA = dict(intrinsics=[[32.0, 0.0, 16.0], [0.0, 32.0, 16.0], [0.0, 0.0, 1.0]], height=32, width=32,
              q=[[0.0, 0.0, 0.0, 1.0]], t=[[0.0, 0.0, -2.0]], near_plane=1.0, far_plane=10.0)


def optimise(make_rasteriser, make_controller, camera_info, device="cpu"):
    """The loop itself, shared with the tests (they pass their own operator / controller)."""
    xyz, feat, mask, obj = (t.to(device) for t in initial_parameters())
    xyz, feat = torch.nn.Parameter(xyz), torch.nn.Parameter(feat)
    controller = make_controller(xyz, feat, mask, obj)
    rasteriser, make_input = make_rasteriser(controller.update)
    optimizer = torch.optim.Adam([xyz, feat], lr=LR)
    target = fake_image().to(device)
    q, t = torch.tensor(CAMERA["q"], device=device), torch.tensor(CAMERA["t"], device=device)
    losses = []
    for idx in range(ITERATIONS):
        optimizer.zero_grad()
        image, _, _ = rasteriser(make_input(point_cloud=xyz, point_cloud_features=feat, point_object_id=obj,
                                            point_invalid_mask=mask, camera_info=camera_info, q_pointcloud_camera=q,
                                            t_pointcloud_camera=t, color_max_sh_band=idx // BAND_INTERVAL))
        loss = ((image - target) ** 2).sum()
        loss.backward()
        optimizer.step()
        controller.refinement()
        losses.append(float(loss.detach()))
    return losses, xyz.detach().cpu(), feat.detach().cpu(), controller


def main():
    taichi_shim.install()
    plt = mock.MagicMock()
    plt.subplots.return_value = (mock.MagicMock(), mock.MagicMock())
    sys.modules["matplotlib"] = mock.MagicMock(pyplot=plt)
    sys.modules["matplotlib.pyplot"] = plt
    sys.modules.setdefault("dataclass_wizard", types.SimpleNamespace(YAMLWizard=object))
    sys.path.insert(0, "/root/reference")
    plain_sort = torch.Tensor.sort
    torch.Tensor.sort = lambda self, *a, **k: plain_sort(self, *a, **{"stable": True, **k})
    from taichi_3d_gaussian_splatting.Camera import CameraInfo
    from taichi_3d_gaussian_splatting.GaussianPointAdaptiveController import GaussianPointAdaptiveController as C
    from taichi_3d_gaussian_splatting.GaussianPointCloudRasterisation import GaussianPointCloudRasterisation as G

    def make_controller(xyz, feat, mask, obj):
        return C(config=C.GaussianPointAdaptiveControllerConfig(),
                 maintained_parameters=C.GaussianPointAdaptiveControllerMaintainedParameters(
                     pointcloud=xyz, pointcloud_features=feat, point_invalid_mask=mask, point_object_id=obj))

    def make_rasteriser(hook):
        module = G(config=G.GaussianPointCloudRasterisationConfig(near_plane=CAMERA["near_plane"], far_plane=CAMERA["far_plane"]),
                   backward_valid_point_hook=hook)
        return module, G.GaussianPointCloudRasterisationInput

    info = CameraInfo(camera_intrinsics=torch.tensor(CAMERA["intrinsics"]), camera_height=32, camera_width=32, camera_id=0)
    with contextlib.redirect_stdout(io.StringIO()):
        losses, xyz, feat, controller = optimise(make_rasteriser, make_controller, info)
    out = dict(losses=losses, xyz=xyz.tolist(), features=feat.tolist(),
               accumulated_num_in_camera=controller.accumulated_num_in_camera.tolist(),
               accumulated_num_pixels=controller.accumulated_num_pixels.tolist(),
               accumulated_view_space_position_gradients=controller.accumulated_view_space_position_gradients.tolist())
    with open(os.path.join(HERE, "training_vectors.json"), "w") as f:
        json.dump(out, f)
    print("losses", [round(x, 4) for x in losses[:3]], "...", [round(x, 4) for x in losses[-3:]])


if __name__ == "__main__":
    main()
