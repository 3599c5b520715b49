"""-m gpu: BASELINE config 5 in miniature -- the trainer harness (two Adams, LR decay, down-sampling and SH-band
schedules, L1 + D-SSIM, densification hook) run with the CUDA operator and, on identical data and initial
state, with the CPU oracle behind the same interface: PSNR / loss trajectories must agree."""
import numpy as np
import pytest
import torch

from taichi_3d_gaussian_splatting_b200 import GaussianPointCloudRasterisation as GPCR
from taichi_3d_gaussian_splatting_b200.trainer import GaussianPointCloudTrainer

from oracle_module import OracleRasterisationModule
from trainer_helpers import hidden_scene, initial_scene, render_views, train_config

pytestmark = pytest.mark.gpu


def _views_to(views, device):
    from taichi_3d_gaussian_splatting_b200 import CameraInfo
    return [(img.to(device), q.to(device), t.to(device),
             CameraInfo(cam.camera_intrinsics.to(device), cam.camera_height, cam.camera_width, cam.camera_id))
            for img, q, t, cam in views]


def test_trainer_psnr_parity_cuda_vs_oracle():
    hidden = hidden_scene(n=400)
    views_cpu = render_views(OracleRasterisationModule(GPCR.GaussianPointCloudRasterisationConfig()), hidden)
    iters = 80
    t_cpu = GaussianPointCloudTrainer(train_config(iters), initial_scene(hidden), views_cpu,
                                      rasterisation_factory=OracleRasterisationModule)
    h_cpu = t_cpu.train(log_interval=1)
    t_gpu = GaussianPointCloudTrainer(train_config(iters), initial_scene(hidden, device="cuda"), _views_to(views_cpu, "cuda"))
    h_gpu = t_gpu.train(log_interval=1)
    psnr_cpu, psnr_gpu = t_cpu.validation(), t_gpu.validation()
    assert psnr_cpu > 26.0 and abs(psnr_cpu - psnr_gpu) < 0.15, (psnr_cpu, psnr_gpu)
    l_cpu = np.array([h["loss"] for h in h_cpu])
    l_gpu = np.array([h["loss"] for h in h_gpu])
    assert np.abs(l_cpu - l_gpu).max() < 2e-3 * l_cpu.max(), np.abs(l_cpu - l_gpu).max()
    # the target views rendered by the CUDA operator equal the oracle's
    views_gpu = render_views(GPCR(GPCR.GaussianPointCloudRasterisationConfig()), hidden, device="cuda")
    for (a, *_), (b, *_) in zip(views_cpu, views_gpu):
        assert float((a - b.cpu()).abs().max()) <= 1e-4


def test_trainer_with_densification_cuda():
    """reference tests/GaussianPointAdaptiveController_test.py:14-95: loss decreases with densification enabled."""
    hidden = hidden_scene(n=400)
    views = render_views(GPCR(GPCR.GaussianPointCloudRasterisationConfig()), hidden, device="cuda")
    trainer = GaussianPointCloudTrainer(train_config(120, densify=True), initial_scene(hidden, device="cuda"), views)
    psnr0 = trainer.validation()
    hist = trainer.train(log_interval=1)
    psnr1 = trainer.validation()
    assert psnr1 > psnr0 + 2.0, (psnr0, psnr1)
    assert np.mean([h["loss"] for h in hist[-8:]]) < 0.8 * np.mean([h["loss"] for h in hist[:8]])
    assert hist[-1]["num_valid_points"] > hist[0]["num_valid_points"]
