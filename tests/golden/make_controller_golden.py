"""Golden trajectory of the REFERENCE's GaussianPointAdaptiveController (GaussianPointAdaptiveController.py:46-353,
imported from /root/reference with taichi / matplotlib / dataclass_wizard stubbed; its two Taichi kernels are not
reached because the fixture config disables sample_from_point and the ellipsoid offset) on the scenario of
controller_fixture.py: after every iteration the maintained tensors and the six accumulators are stored.

    python tests/golden/make_controller_golden.py        # build container only; writes controller_vectors.json
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
from controller_fixture import CONFIG, ITERATIONS, hook_fields, initial_state  # noqa: E402


def main():
    ti = mock.MagicMock()
    ti.func = lambda f: f
    ti.kernel = lambda f: f
    ti.dataclass = lambda c: c
    sys.modules["taichi"] = ti
    sys.modules["taichi.math"] = ti.math
    plt = mock.MagicMock()
    plt.subplots.return_value = (mock.MagicMock(), mock.MagicMock())
    sys.modules["matplotlib"] = mock.MagicMock(pyplot=plt)
    sys.modules["matplotlib.pyplot"] = plt
    sys.modules.setdefault("dataclass_wizard", types.SimpleNamespace(YAMLWizard=object))
    sys.path.insert(0, "/root/reference")
    from taichi_3d_gaussian_splatting.GaussianPointAdaptiveController import GaussianPointAdaptiveController as C
    from taichi_3d_gaussian_splatting.GaussianPointCloudRasterisation import GaussianPointCloudRasterisation as R

    xyz, feat, mask, obj = initial_state()
    ctl = C(config=C.GaussianPointAdaptiveControllerConfig(**CONFIG),
            maintained_parameters=C.GaussianPointAdaptiveControllerMaintainedParameters(
                pointcloud=xyz, pointcloud_features=feat, point_invalid_mask=mask, point_object_id=obj))
    steps = []
    for it in range(ITERATIONS):
        with contextlib.redirect_stdout(io.StringIO()):
            ctl.update(R.BackwardValidPointHookInput(**hook_fields(it, mask)))
            info = ctl.densify_point_info
            found = None if info is None else dict(
                floater_point_id=info.floater_point_id.tolist(), transparent_point_id=info.transparent_point_id.tolist(),
                densify_point_id=info.densify_point_id.tolist(),
                densify_size_reduction_factor=info.densify_size_reduction_factor.flatten().tolist(),
                densify_point_grad_position=info.densify_point_grad_position.tolist())
            ctl.refinement()
        steps.append(dict(found=found, mask=mask.tolist(), obj=obj.tolist(), xyz=xyz.tolist(),
                          scale_alpha=feat[:, 4:8].tolist(), feat_checksum=float(torch.nan_to_num(feat).double().sum()),
                          acc_pixels=ctl.accumulated_num_pixels.tolist(), acc_in_camera=ctl.accumulated_num_in_camera.tolist(),
                          acc_view=ctl.accumulated_view_space_position_gradients.tolist(),
                          acc_view_avg=ctl.accumulated_view_space_position_gradients_avg.tolist(),
                          acc_pos=ctl.accumulated_position_gradients.tolist(),
                          acc_pos_norm=ctl.accumulated_position_gradients_norm.tolist()))
    with open(os.path.join(HERE, "controller_vectors.json"), "w") as f:
        json.dump(steps, f)
    for it, s in enumerate(steps):
        fd = s["found"]
        print(it, "valid", s["mask"].count(0), None if fd is None else
              {k: len(v) for k, v in fd.items() if k.endswith("_id")})


if __name__ == "__main__":
    main()
