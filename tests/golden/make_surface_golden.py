"""The reference's Python surface for the path and its widened rows, read off the reference modules themselves (imported
from /root/reference under taichi_shim.py): dataclass fields with defaults, constructor / forward parameters, module
constants, public method names.  tests/test_abi_cpu.py compares our classes with it field by field.

    python tests/golden/make_surface_golden.py        # build container only; writes surface.json
"""
import dataclasses
import inspect
import json
import os
import sys
import types
from unittest import mock

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import taichi_shim  # noqa: E402


def fields_of(cls):
    out = []
    for f in dataclasses.fields(cls):
        default = None if f.default is dataclasses.MISSING else f.default
        out.append([f.name, default if isinstance(default, (int, float, bool, str, type(None))) else repr(default),
                    f.default is not dataclasses.MISSING])
    return out


def params_of(fn):
    return [p for p in inspect.signature(fn).parameters if p != "self"]


def main():
    taichi_shim.install()
    plt = mock.MagicMock()
    plt.subplots.return_value = (mock.MagicMock(), mock.MagicMock())
    sys.modules["matplotlib"] = mock.MagicMock(pyplot=plt)
    sys.modules["matplotlib.pyplot"] = plt
    sys.modules.setdefault("dataclass_wizard", types.SimpleNamespace(YAMLWizard=object))
    sys.modules.setdefault("plyfile", types.SimpleNamespace(PlyData=object, PlyElement=object))
    sys.path.insert(0, "/root/reference")
    import taichi_3d_gaussian_splatting.GaussianPointCloudRasterisation as rast
    from taichi_3d_gaussian_splatting.Camera import CameraInfo
    from taichi_3d_gaussian_splatting.GaussianPointAdaptiveController import GaussianPointAdaptiveController as C
    from taichi_3d_gaussian_splatting.GaussianPointCloudScene import GaussianPointCloudScene as S
    from taichi_3d_gaussian_splatting.ImagePoseDataset import ImagePoseDataset as D
    G = rast.GaussianPointCloudRasterisation
    out = dict(
        constants=dict(TILE_WIDTH=rast.TILE_WIDTH, TILE_HEIGHT=rast.TILE_HEIGHT, BOUNDARY_TILES=rast.BOUNDARY_TILES),
        rasterisation=dict(
            config=fields_of(G.GaussianPointCloudRasterisationConfig), input=fields_of(G.GaussianPointCloudRasterisationInput),
            hook_input=fields_of(G.BackwardValidPointHookInput), init=params_of(G.__init__), forward=params_of(G.forward)),
        camera_info=fields_of(CameraInfo),
        controller=dict(config=fields_of(C.GaussianPointAdaptiveControllerConfig),
                        maintained=fields_of(C.GaussianPointAdaptiveControllerMaintainedParameters),
                        densify_info=fields_of(C.GaussianPointAdaptiveControllerDensifyPointInfo),
                        init=params_of(C.__init__), methods=["update", "refinement", "reset_alpha"]),
        scene=dict(config=fields_of(S.PointCloudSceneConfig), init=params_of(S.__init__),
                   methods=["forward", "initialize", "to_parquet", "to_ply", "from_parquet"]),
        dataset=dict(init=params_of(D.__init__), methods=["__len__", "__getitem__", "_autoscale_image_and_camera_info"]))
    for name in out["controller"]["methods"]:
        assert callable(getattr(C, name))
    for name in out["scene"]["methods"]:
        assert callable(getattr(S, name))
    with open(os.path.join(HERE, "surface.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out["rasterisation"], indent=0)[:600])


if __name__ == "__main__":
    main()
