"""Training loss of the reference trainer: ``L = (1 - lambda) * L1 + lambda * (1 - SSIM)``
(+ optional exp(s) regulariser), same class / config names as
``taichi_3d_gaussian_splatting/LossFunction.py:8-54``.

The reference takes SSIM from the third-party ``pytorch_msssim`` package (``requirements.txt:4``, unpinned,
not installed in this image); its published algorithm is restated here: 11x11 Gaussian window with
sigma = 1.5 applied separably with VALID padding, K1 = 0.01, K2 = 0.03, ``data_range = 1``,
``size_average = True`` (mean over channels and batch), no non-negative clamp.  No reference test touches
the loss, so its parity is unpinned.
"""
from dataclasses import dataclass

import torch
import torch.nn as nn
import torch.nn.functional as F


def _gaussian_window(size: int, sigma: float, device, dtype) -> torch.Tensor:
    coords = torch.arange(size, device=device, dtype=dtype) - size // 2
    g = torch.exp(-(coords ** 2) / (2 * sigma ** 2))
    return g / g.sum()


def _filter(x: torch.Tensor, win: torch.Tensor) -> torch.Tensor:
    c = x.shape[1]
    k = win.numel()
    out = F.conv2d(x, win.view(1, 1, k, 1).expand(c, 1, k, 1), groups=c)
    return F.conv2d(out, win.view(1, 1, 1, k).expand(c, 1, 1, k), groups=c)


def ssim(x: torch.Tensor, y: torch.Tensor, data_range: float = 1.0, size_average: bool = True,
         win_size: int = 11, win_sigma: float = 1.5, K=(0.01, 0.03)) -> torch.Tensor:
    """Structural similarity of (B, C, H, W) images (H, W > win_size - 1)."""
    win = _gaussian_window(win_size, win_sigma, x.device, x.dtype)
    c1, c2 = (K[0] * data_range) ** 2, (K[1] * data_range) ** 2
    mu1, mu2 = _filter(x, win), _filter(y, win)
    mu1_sq, mu2_sq, mu12 = mu1 * mu1, mu2 * mu2, mu1 * mu2
    s1 = _filter(x * x, win) - mu1_sq
    s2 = _filter(y * y, win) - mu2_sq
    s12 = _filter(x * y, win) - mu12
    cs_map = (2 * s12 + c2) / (s1 + s2 + c2)
    ssim_map = ((2 * mu12 + c1) / (mu1_sq + mu2_sq + c1)) * cs_map
    per_channel = ssim_map.flatten(2).mean(-1)
    return per_channel.mean() if size_average else per_channel.mean(1)


class LossFunction(nn.Module):
    @dataclass
    class LossFunctionConfig:
        lambda_value: float = 0.2
        enable_regularization: bool = True
        regularization_weight: float = 2

    def __init__(self, config: "LossFunction.LossFunctionConfig"):
        super().__init__()
        self.config = config

    def forward(self, predicted_image, ground_truth_image, point_invalid_mask=None, pointcloud_features=None):
        """predicted / ground truth: (B, C, H, W) or (C, H, W).  Returns (L, L1, 1 - SSIM)."""
        if predicted_image.dim() == 3:
            predicted_image = predicted_image.unsqueeze(0)
        if ground_truth_image.dim() == 3:
            ground_truth_image = ground_truth_image.unsqueeze(0)
        l1 = torch.abs(predicted_image - ground_truth_image).mean()
        ld_ssim = 1 - ssim(predicted_image, ground_truth_image, data_range=1, size_average=True)
        loss = (1 - self.config.lambda_value) * l1 + self.config.lambda_value * ld_ssim
        if pointcloud_features is not None and self.config.enable_regularization:
            s = pointcloud_features[point_invalid_mask == 0, 4:7]
            loss = loss + self.config.regularization_weight * torch.norm(torch.exp(s), dim=1).mean()
        return loss, l1, ld_ssim
