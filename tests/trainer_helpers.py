"""Shared set-up of the tiny training problem used by the CPU (oracle) and GPU (CUDA) trainer tests."""
import math

import torch

from taichi_3d_gaussian_splatting_b200 import CameraInfo
from taichi_3d_gaussian_splatting_b200.synthetic import make_scene
from taichi_3d_gaussian_splatting_b200.trainer import GaussianPointCloudTrainer, Scene

H, W = 64, 96
YAWS = (-6.0, -2.0, 2.0, 6.0)


def hidden_scene(n=600, seed=3):
    sc = make_scene(n, H, W, 0.10, seed, sh_degree=1)
    sc.point_cloud[:, 2] = sc.point_cloud[:, 2] * 0.5 + 1.0
    sc.point_cloud_features[:, 7] += 1.0
    return sc


def poses():
    out = []
    for yaw in YAWS:
        half = math.radians(yaw) / 2
        out.append((torch.tensor([[0.0, math.sin(half), 0.0, math.cos(half)]]), torch.zeros((1, 3))))
    return out


def initial_scene(hidden, capacity_ratio=2.0, seed=9, device="cpu"):
    """Perturbed copy of the hidden scene + spare invalid slots (the reference's fixed-capacity layout)."""
    g = torch.Generator().manual_seed(seed)
    n = hidden.point_cloud.shape[0]
    cap = int(n * capacity_ratio)
    pc = torch.zeros((cap, 3))
    feat = torch.zeros((cap, 56))
    feat[:, 3] = 1.0
    pc[:n] = hidden.point_cloud + 0.05 * torch.randn((n, 3), generator=g)
    feat[:n] = hidden.point_cloud_features
    feat[:n, 4:7] += 0.3 * torch.randn((n, 3), generator=g)
    feat[:n, 8:] = 0.5 * feat[:n, 8:]
    feat[:n, 7] = 0.5
    mask = torch.ones(cap, dtype=torch.int8)
    mask[:n] = 0
    return Scene(point_cloud=pc.to(device).requires_grad_(True), point_cloud_features=feat.to(device).requires_grad_(True),
                 point_invalid_mask=mask.to(device), point_object_id=torch.zeros(cap, dtype=torch.int32, device=device))


def render_views(module, hidden, device="cpu"):
    """Target images = renders of the hidden scene with the given rasteriser module."""
    from taichi_3d_gaussian_splatting_b200 import GaussianPointCloudRasterisation as GPCR
    views = []
    K = hidden.camera_info.camera_intrinsics.to(device)
    pc, feat = hidden.point_cloud.to(device), hidden.point_cloud_features.clone().to(device)
    mask, obj = hidden.point_invalid_mask.to(device), hidden.point_object_id.to(device)
    for q, t in poses():
        q, t = q.to(device), t.to(device)
        cam = CameraInfo(K, H, W, 0)
        with torch.no_grad():
            img, _, _ = module(GPCR.GaussianPointCloudRasterisationInput(
                point_cloud=pc, point_cloud_features=feat, point_object_id=obj, point_invalid_mask=mask,
                camera_info=cam, q_pointcloud_camera=q, t_pointcloud_camera=t, color_max_sh_band=3))
        views.append((img.clamp(0, 1).permute(2, 0, 1).contiguous(), q, t, cam))
    return views


def train_config(num_iterations, densify=False):
    C = GaussianPointCloudTrainer.TrainConfig
    cfg = C(num_iterations=num_iterations, feature_learning_rate=5e-3, position_learning_rate=2e-4,
            initial_downsample_factor=2, half_downsample_factor_interval=20,
            increase_color_max_sh_band_interval=30.0)
    ac = cfg.adaptive_controller_config
    if densify:
        ac.num_iterations_warm_up = 20
        ac.num_iterations_densify = 20
        ac.densification_view_space_position_gradients_threshold = 1.2e-3  # ~ top 15 % of the splats per frame
        ac.transparent_alpha_threshold = -3.0
        ac.num_iterations_reset_alpha = 10_000
        ac.under_reconstructed_num_pixels_threshold = 64
    else:
        ac.num_iterations_warm_up = 10 ** 9
    cfg.loss_function_config.enable_regularization = False
    return cfg
