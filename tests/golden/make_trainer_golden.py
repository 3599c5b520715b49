"""Golden vectors for the trainer helpers: the REFERENCE's ``GaussianPointCloudTrainer._downsample_image_and_camera_info``
(GaussianPointTrainer.py:97-116) and the PSNR half of ``_compute_pnsr_and_ssim`` (:278-285), imported from
/root/reference with taichi / matplotlib / tensorboard / pytorch_msssim / plyfile / dataclass_wizard stubbed.

    python tests/golden/make_trainer_golden.py        # build container only; writes trainer_vectors.json
"""
import json
import os
import sys
import types
from unittest import mock

import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def fixture_image(h, w, seed):
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.linspace(0, 1, h), torch.linspace(0, 1, w), indexing="ij")
    base = torch.stack([0.5 + 0.5 * torch.sin(9 * xx + seed), yy * xx, 0.5 + 0.5 * torch.cos(7 * yy - seed)])
    return (base + 0.05 * torch.randn(base.shape, generator=g)).clamp(0, 1)


CASES = [(544, 976, 4), (544, 976, 2), (250, 330, 4), (97, 131, 2), (64, 64, 1)]


def main():
    # The trainer module itself does not import on Python >= 3.11 (its TrainConfig dataclass has dataclass-instance
    # defaults), so the two static helpers are lifted out of the reference source with ``ast`` and executed as they are.
    import ast
    import textwrap
    import torchvision.transforms as transforms
    sys.modules.setdefault("dataclass_wizard", types.SimpleNamespace(YAMLWizard=object))
    sys.path.insert(0, "/root/reference")
    ti = mock.MagicMock()
    sys.modules["taichi"] = ti
    sys.modules["taichi.math"] = ti.math
    from taichi_3d_gaussian_splatting.Camera import CameraInfo
    path = "/root/reference/taichi_3d_gaussian_splatting/GaussianPointTrainer.py"
    source = open(path).read()
    wanted = {"_downsample_image_and_camera_info", "_compute_pnsr_and_ssim"}
    namespace = dict(torch=torch, transforms=transforms, CameraInfo=CameraInfo, ssim=lambda *a, **k: torch.tensor(0.0))
    for node in ast.walk(ast.parse(source)):
        if isinstance(node, ast.FunctionDef) and node.name in wanted:
            node.decorator_list = []
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), namespace)
    T = types.SimpleNamespace(**{name: namespace[name] for name in wanted})
    out = dict(downsample=[], psnr=[])
    for k, (h, w, factor) in enumerate(CASES):
        image = fixture_image(h, w, k)
        info = CameraInfo(camera_intrinsics=torch.tensor([[0.9 * w, 0.0, w / 2 + 3.0], [0.0, 0.8 * h, h / 2 - 1.0], [0.0, 0.0, 1.0]]),
                          camera_height=h, camera_width=w, camera_id=k)
        small, small_info = T._downsample_image_and_camera_info(image, info, factor)
        hh, ww = small.shape[1:]
        probes = [(0, 0), (hh // 2, ww // 3), (hh - 1, ww - 1)]
        out["downsample"].append(dict(h=h, w=w, factor=factor, shape=list(small.shape), mean=float(small.double().mean()),
                                      probes=[[y, x] + small[:, y, x].tolist() for y, x in probes],
                                      K=small_info.camera_intrinsics.tolist(), camera_height=int(small_info.camera_height),
                                      camera_width=int(small_info.camera_width), camera_id=int(small_info.camera_id)))
    for k in range(3):
        a, b = fixture_image(48, 64, 10 + k), fixture_image(48, 64, 20 + k)
        psnr, _ = T._compute_pnsr_and_ssim(image_pred=a, image_gt=b)
        out["psnr"].append(float(psnr))
    with open(os.path.join(HERE, "trainer_vectors.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", len(out["downsample"]), "downsample cases,", out["psnr"])


if __name__ == "__main__":
    main()
