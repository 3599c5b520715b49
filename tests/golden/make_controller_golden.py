"""Golden trajectory of the REFERENCE's GaussianPointAdaptiveController (GaussianPointAdaptiveController.py:46-353,
imported from /root/reference with matplotlib / dataclass_wizard stubbed and Taichi replaced by taichi_shim.py; the random
sample_from_point kernel is disabled by the fixture config, the deterministic ellipsoid-offset kernel runs in a second scenario) on the scenario of
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
    import taichi_shim  # executes the controller's Taichi kernel (compute_ellipsoid_offset) in the second scenario
    taichi_shim.install()
    plt = mock.MagicMock()
    plt.subplots.return_value = (mock.MagicMock(), mock.MagicMock())
    sys.modules["matplotlib"] = mock.MagicMock(pyplot=plt)
    sys.modules["matplotlib.pyplot"] = plt
    sys.modules.setdefault("dataclass_wizard", types.SimpleNamespace(YAMLWizard=object))
    sys.path.insert(0, "/root/reference")
    from taichi_3d_gaussian_splatting.GaussianPointAdaptiveController import GaussianPointAdaptiveController as C
    from taichi_3d_gaussian_splatting.GaussianPointCloudRasterisation import GaussianPointCloudRasterisation as R

    def run(config):
        xyz, feat, mask, obj = initial_state()
        ctl = C(config=C.GaussianPointAdaptiveControllerConfig(**config),
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
        return steps

    steps = run(CONFIG)
    # second scenario: split points are moved to the foci of their ellipsoid (compute_ellipsoid_offset, a Taichi kernel
    # of the reference: GaussianPointAdaptiveController.py:10-26 with GaussianPoint3D.get_ellipsoid_foci_vector)
    steps_foci = run({**CONFIG, "enable_ellipsoid_offset": True})
    with open(os.path.join(HERE, "controller_vectors.json"), "w") as f:
        json.dump(dict(default=steps, ellipsoid_offset=steps_foci), f)
    for it, s in enumerate(steps):
        fd = s["found"]
        print(it, "valid", s["mask"].count(0), None if fd is None else
              {k: len(v) for k, v in fd.items() if k.endswith("_id")})


if __name__ == "__main__":
    main()
