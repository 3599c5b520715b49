"""D-SSIM pinned to the PUBLISHED definition (tests/golden/make_ssim_known_answers.py: hand-derived analytic values and an
independent float64 textbook evaluation), for both implementations: ``loss.ssim`` (torch, CPU) and ``gsb200_image_loss``
(csrc/image_loss.cu, -m gpu).  Parity against the ``pytorch_msssim`` package the reference imports (LossFunction.py:4) stays
UNPINNED: the package is not installed in this image and no reference test touches the loss."""
import os

import numpy as np
import pytest
import torch

from taichi_3d_gaussian_splatting_b200.loss import LossFunction, ssim

HERE = os.path.dirname(os.path.abspath(__file__))
KA = np.load(os.path.join(HERE, "golden", "ssim_known_answers.npz"))
F32_TOL = 1e-4
IMAGE_CASES = sorted({k.split("/")[0] for k in KA.files if k.endswith("/x")})
CONSTANT_CASES = sorted({k.split("/")[0] for k in KA.files if k.endswith("/a")})


def _constant_pair(name, H=24, W=28):
    a, b = float(KA[name + "/a"]), float(KA[name + "/b"])
    return torch.full((3, H, W), a), torch.full((3, H, W), b)


@pytest.mark.parametrize("name", IMAGE_CASES)
def test_torch_ssim_matches_the_textbook_evaluation(name):
    x, y = torch.from_numpy(KA[name + "/x"]), torch.from_numpy(KA[name + "/y"])
    # float64: the algorithm itself, to rounding; float32 (what the trainer runs): sigma^2 = E[x^2] - mu^2 cancels against
    # C2 = 9e-4, worth up to ~1e-4 on smooth images -- inherent to a float32 evaluation of the published formula
    assert abs(float(ssim(x[None].double(), y[None].double())) - float(KA[name + "/ssim"])) <= 1e-12
    got = float(ssim(x[None], y[None]))
    assert abs(got - float(KA[name + "/ssim"])) <= F32_TOL, (got, float(KA[name + "/ssim"]))
    assert abs(float(ssim(x[None], x[None])) - 1.0) <= 1e-6  # identical images


@pytest.mark.parametrize("name", CONSTANT_CASES)
def test_torch_ssim_matches_the_hand_derived_constant_image_value(name):
    x, y = _constant_pair(name)
    a, b = float(KA[name + "/a"]), float(KA[name + "/b"])
    x64, y64 = torch.full((1, 3, 24, 28), a, dtype=torch.float64), torch.full((1, 3, 24, 28), b, dtype=torch.float64)
    assert abs(float(ssim(x64, y64)) - float(KA[name + "/ssim"])) <= 1e-12
    assert abs(float(ssim(x[None], y[None])) - float(KA[name + "/ssim"])) <= F32_TOL


def test_loss_function_mixes_l1_and_dssim_as_the_reference_does():
    """LossFunction.py:29-33: (1 - lambda) L1 + lambda (1 - SSIM) with lambda = 0.2, on a known-answer pair."""
    name = "ramp_vs_noisy_ramp"
    x, y = torch.from_numpy(KA[name + "/x"]), torch.from_numpy(KA[name + "/y"])
    cfg = LossFunction.LossFunctionConfig()
    cfg.enable_regularization = False
    loss, l1, ssim_loss = LossFunction(cfg)(x, y)
    l1_expected = float(np.abs(KA[name + "/x"].astype(np.float64) - KA[name + "/y"].astype(np.float64)).mean())
    assert abs(float(l1) - l1_expected) <= 1e-6 and abs(float(ssim_loss) - (1.0 - float(KA[name + "/ssim"]))) <= F32_TOL
    assert abs(float(loss) - (0.8 * l1_expected + 0.2 * (1.0 - float(KA[name + "/ssim"])))) <= F32_TOL


@pytest.mark.gpu
@pytest.mark.parametrize("name", IMAGE_CASES + CONSTANT_CASES)
def test_cuda_image_loss_matches_the_known_answers(name):
    from taichi_3d_gaussian_splatting_b200 import fused_image_loss_with_grad
    if name in IMAGE_CASES:
        x, y = torch.from_numpy(KA[name + "/x"]), torch.from_numpy(KA[name + "/y"])
    else:
        x, y = _constant_pair(name)
    pred = x.permute(1, 2, 0).contiguous().cuda()  # the rasteriser's (H, W, 3) layout; values already inside [0, 1]
    losses, _ = fused_image_loss_with_grad(pred, y.cuda(), lambda_value=0.2, want_grad=False)
    total, l1, dssim = (float(v) for v in losses.cpu())
    expected = float(KA[name + "/ssim"])
    assert abs(dssim - (1.0 - expected)) <= F32_TOL, (dssim, 1.0 - expected)
    l1_expected = float((x.double() - y.double()).abs().mean())
    assert abs(l1 - l1_expected) <= 1e-6 and abs(total - (0.8 * l1_expected + 0.2 * (1.0 - expected))) <= F32_TOL
