"""Densification / pruning controller -- consumer of the rasteriser's backward hook.

Host-side policy code with the reference's names and semantics
(``taichi_3d_gaussian_splatting/GaussianPointAdaptiveController.py:46-393``): fixed-capacity point
cloud with an invalid mask, ``update`` is called from inside the operator's backward with
``BackwardValidPointHookInput`` (GPCR:1127-1142), ``refinement`` after the optimiser step.
The reference implements it with torch tensor ops plus two tiny Taichi kernels
(``compute_ellipsoid_offset`` :10-25, ``sample_from_point`` :27-42); those two are restated with
torch ops here (GaussianPoint3D.py:375-406).  The matplotlib debug plot (:272-288) is not reproduced.
"""
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from .GaussianPointCloudRasterisation import GaussianPointCloudRasterisation
from .utils import quaternion_to_rotation_matrix_torch


def compute_ellipsoid_offset(pointcloud: torch.Tensor, pointcloud_features: torch.Tensor) -> torch.Tensor:
    """Vector from the centre to a focus of each ellipsoid (GaussianPoint3D.py:375-388)."""
    s = pointcloud_features[:, 4:7]
    base = torch.zeros_like(pointcloud)
    sx, sy, sz = s[:, 0], s[:, 1], s[:, 2]
    use_y = (sx < sy) & (sy > sz)
    use_z = (sx < sz) & (sy < sz) & ~use_y
    use_x = ~(use_y | use_z)
    base[use_x, 0] = 1.0
    base[use_y, 1] = 1.0
    base[use_z, 2] = 1.0
    R = quaternion_to_rotation_matrix_torch(pointcloud_features[:, 0:4])
    base = torch.einsum("nij,nj->ni", R, base)
    es = torch.exp(s)
    r_c = es.max(dim=1).values
    r_a = es.min(dim=1).values
    return torch.sqrt(torch.clamp(r_c ** 2 - r_a ** 2, min=0.0))[:, None] * base


def sample_from_point(pointcloud: torch.Tensor, pointcloud_features: torch.Tensor,
                      generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """One sample of N(mean, R S S^T R^T) per point (GaussianPoint3D.py:390-406; the reference draws its
    normals with Box-Muller on ``ti.random``, here with ``torch.randn``)."""
    R = quaternion_to_rotation_matrix_torch(pointcloud_features[:, 0:4])
    es = torch.exp(pointcloud_features[:, 4:7])
    z = torch.randn(pointcloud.shape, device=pointcloud.device, dtype=pointcloud.dtype, generator=generator)
    return pointcloud + torch.einsum("nij,nj->ni", R, es * z)


class GaussianPointAdaptiveController:
    @dataclass
    class GaussianPointAdaptiveControllerConfig:
        # reference: GaussianPointAdaptiveController.py:53-84
        num_iterations_warm_up: int = 500
        num_iterations_densify: int = 100
        transparent_alpha_threshold: float = -0.5
        densification_view_space_position_gradients_threshold: float = 6e-6
        densification_view_avg_space_position_gradients_threshold: float = 1e3
        densification_multi_frame_view_space_position_gradients_threshold: float = 1e3
        densification_multi_frame_view_pixel_avg_space_position_gradients_threshold: float = 1e3
        densification_multi_frame_position_gradients_threshold: float = 1e3
        gaussian_split_factor_phi: float = 1.6
        num_iterations_reset_alpha: int = 3000
        reset_alpha_value: float = 0.1
        floater_num_pixels_threshold: int = 10000
        floater_near_camrea_num_pixels_threshold: int = 10000
        floater_depth_threshold: float = 100
        iteration_start_remove_floater: int = 2000
        plot_densify_interval: int = 200
        under_reconstructed_num_pixels_threshold: int = 512
        under_reconstructed_move_factor: float = 100.0
        enable_ellipsoid_offset: bool = False
        enable_sample_from_point: bool = True

    @dataclass
    class GaussianPointAdaptiveControllerMaintainedParameters:
        pointcloud: torch.Tensor  # [num_points, 3]
        pointcloud_features: torch.Tensor  # [num_points, 56]
        point_invalid_mask: torch.Tensor  # [num_points] int8
        point_object_id: torch.Tensor  # [num_points] int32

    @dataclass
    class GaussianPointAdaptiveControllerDensifyPointInfo:
        floater_point_id: torch.Tensor
        transparent_point_id: torch.Tensor
        densify_point_id: torch.Tensor
        densify_point_position_before_optimization: torch.Tensor
        densify_size_reduction_factor: torch.Tensor
        densify_point_grad_position: torch.Tensor

    def __init__(self, config: "GaussianPointAdaptiveController.GaussianPointAdaptiveControllerConfig",
                 maintained_parameters: "GaussianPointAdaptiveController.GaussianPointAdaptiveControllerMaintainedParameters",
                 generator: Optional[torch.Generator] = None, verbose: bool = False):
        self.iteration_counter = -1
        self.config = config
        self.maintained_parameters = maintained_parameters
        self.input_data = None
        self.densify_point_info = None
        self.generator = generator
        self.verbose = verbose
        self.has_plot = False
        self._reset_accumulators()

    def _reset_accumulators(self):
        col = self.maintained_parameters.pointcloud[:, 0]
        self.accumulated_num_pixels = torch.zeros_like(col, dtype=torch.int32)
        self.accumulated_num_in_camera = torch.zeros_like(col, dtype=torch.int32)
        self.accumulated_view_space_position_gradients = torch.zeros_like(col, dtype=torch.float32)
        self.accumulated_view_space_position_gradients_avg = torch.zeros_like(col, dtype=torch.float32)
        self.accumulated_position_gradients = torch.zeros_like(self.maintained_parameters.pointcloud,
                                                               dtype=torch.float32)
        self.accumulated_position_gradients_norm = torch.zeros_like(col, dtype=torch.float32)

    # GaussianPointAdaptiveController.py:130-146
    def update(self, input_data: GaussianPointCloudRasterisation.BackwardValidPointHookInput):
        self.iteration_counter += 1
        with torch.no_grad():
            ids = input_data.point_id_in_camera_list.long()
            self.accumulated_num_in_camera[ids] += 1
            self.accumulated_num_pixels[ids] += input_data.num_affected_pixels
            grad_viewspace_norm = input_data.magnitude_grad_viewspace
            self.accumulated_view_space_position_gradients[ids] += grad_viewspace_norm
            avg = grad_viewspace_norm / input_data.num_affected_pixels
            avg[torch.isnan(avg)] = 0
            self.accumulated_view_space_position_gradients_avg[ids] += avg
            self.accumulated_position_gradients[ids] += input_data.grad_point_in_camera
            self.accumulated_position_gradients_norm[ids] += input_data.grad_point_in_camera.norm(dim=1)
            if self.iteration_counter < self.config.num_iterations_warm_up:
                pass
            elif self.iteration_counter % self.config.num_iterations_densify == 0:
                self._find_densify_points(input_data)
                self.input_data = input_data

    # GaussianPointAdaptiveController.py:148-168
    def refinement(self):
        with torch.no_grad():
            if self.iteration_counter < self.config.num_iterations_warm_up:
                return
            if self.iteration_counter % self.config.num_iterations_densify == 0:
                self._add_densify_points()
                self._reset_accumulators()
            if self.iteration_counter % self.config.num_iterations_reset_alpha == 0:
                self.reset_alpha()
            self.input_data = None

    # GaussianPointAdaptiveController.py:170-270
    def _find_densify_points(self, input_data):
        cfg = self.config
        mp = self.maintained_parameters
        pointcloud, features = mp.pointcloud, mp.pointcloud_features
        point_id_list = torch.arange(pointcloud.shape[0], device=pointcloud.device)
        ids_in_camera = input_data.point_id_in_camera_list.long()
        num_affected_pixels = input_data.num_affected_pixels
        point_depth = input_data.point_depth
        average_num_affect_pixels = self.accumulated_num_pixels / self.accumulated_num_in_camera
        average_num_affect_pixels[torch.isnan(average_num_affect_pixels)] = 0

        floater_mask = torch.zeros_like(point_id_list, dtype=torch.bool)
        floater_mask_in_camera = torch.zeros_like(ids_in_camera, dtype=torch.bool)
        floater_point_id = torch.empty(0, dtype=torch.int64, device=pointcloud.device)
        if self.iteration_counter > cfg.iteration_start_remove_floater:
            floater_mask_in_camera = (num_affected_pixels > cfg.floater_near_camrea_num_pixels_threshold) & \
                (point_depth < cfg.floater_depth_threshold)
            floater_point_id = ids_in_camera[floater_mask_in_camera]
            floater_mask[floater_point_id] = True
            floater_mask = floater_mask & (mp.point_invalid_mask == 0)

        point_alpha = features[:, 7]
        nan_mask = torch.isnan(features).any(dim=1)
        transparent_point_mask = ((point_alpha < cfg.transparent_alpha_threshold) | nan_mask) & \
            (mp.point_invalid_mask == 0) & (~floater_mask)
        transparent_point_id = point_id_list[transparent_point_mask]
        will_be_remove_mask = floater_mask | transparent_point_mask

        in_camera_will_be_remove_mask = floater_mask_in_camera | transparent_point_mask[ids_in_camera]
        grad_viewspace_norm = input_data.magnitude_grad_viewspace
        in_camera_to_densify_mask = grad_viewspace_norm > cfg.densification_view_space_position_gradients_threshold
        in_camera_to_densify_mask &= ~in_camera_will_be_remove_mask
        in_camera_to_densify_mask |= (grad_viewspace_norm / num_affected_pixels >
                                      cfg.densification_view_avg_space_position_gradients_threshold)
        in_camera_to_densify_mask &= ~in_camera_will_be_remove_mask

        single_frame_mask = torch.zeros_like(point_id_list, dtype=torch.bool)
        single_frame_mask[ids_in_camera[in_camera_to_densify_mask]] = True

        mf_view = self.accumulated_view_space_position_gradients / self.accumulated_num_in_camera
        mf_view[torch.isnan(mf_view)] = 0
        multi_frame_mask = mf_view > cfg.densification_multi_frame_view_space_position_gradients_threshold
        mf_avg = self.accumulated_view_space_position_gradients_avg / self.accumulated_num_in_camera
        mf_avg[torch.isnan(mf_avg)] = 0
        multi_frame_mask |= (mf_avg / average_num_affect_pixels >
                             cfg.densification_multi_frame_view_pixel_avg_space_position_gradients_threshold)
        mf_pos = self.accumulated_position_gradients_norm / self.accumulated_num_in_camera
        multi_frame_mask |= mf_pos > cfg.densification_multi_frame_position_gradients_threshold
        to_densify_mask = (single_frame_mask | multi_frame_mask) & (~will_be_remove_mask)
        densify_point_id = point_id_list[to_densify_mask]

        position_before = pointcloud[densify_point_id].detach().clone()
        grad_position = self.accumulated_position_gradients[densify_point_id] / \
            self.accumulated_num_in_camera[densify_point_id].unsqueeze(-1)
        grad_position[torch.isnan(grad_position)] = 0
        reduction = torch.zeros_like(densify_point_id, dtype=torch.float32)
        over_reconstructed = self.accumulated_num_pixels[to_densify_mask] > cfg.under_reconstructed_num_pixels_threshold
        reduction[over_reconstructed] = float(np.log(cfg.gaussian_split_factor_phi))
        self.densify_point_info = GaussianPointAdaptiveController.GaussianPointAdaptiveControllerDensifyPointInfo(
            floater_point_id=floater_point_id, transparent_point_id=transparent_point_id,
            densify_point_id=densify_point_id, densify_point_position_before_optimization=position_before,
            densify_size_reduction_factor=reduction.unsqueeze(-1), densify_point_grad_position=grad_position)

    # GaussianPointAdaptiveController.py:290-353
    def _add_densify_points(self):
        assert self.densify_point_info is not None
        cfg, mp, info = self.config, self.maintained_parameters, self.densify_point_info
        valid_before = int(mp.point_invalid_mask.shape[0] - mp.point_invalid_mask.sum())
        num_transparent = info.transparent_point_id.shape[0]
        mp.point_invalid_mask[info.transparent_point_id] = 1
        num_floaters = info.floater_point_id.shape[0]
        mp.point_invalid_mask[info.floater_point_id] = 1
        num_densify = info.densify_point_id.shape[0]
        to_fill = torch.where(mp.point_invalid_mask == 1)[0][:num_densify]
        n_fill = 0
        if num_densify > 0:
            n_fill = min(num_densify, to_fill.shape[0])
            src = info.densify_point_id[:n_fill]
            mp.pointcloud[to_fill] = info.densify_point_position_before_optimization[:n_fill]
            mp.pointcloud_features[to_fill] = mp.pointcloud_features[src]
            mp.point_object_id[to_fill] = mp.point_object_id[src]
            mp.pointcloud_features[to_fill, 4:7] -= info.densify_size_reduction_factor[:n_fill]
            over = (info.densify_size_reduction_factor[:n_fill] > 1e-6).reshape(-1)
            under = ~over
            mp.pointcloud_features[src, 4:7] -= info.densify_size_reduction_factor[:n_fill]
            if cfg.enable_ellipsoid_offset:
                offset = compute_ellipsoid_offset(mp.pointcloud[src], mp.pointcloud_features[src])
                mp.pointcloud[to_fill] += offset
                mp.pointcloud[src] -= offset
            if cfg.enable_sample_from_point:
                over_src, over_dst = src[over], to_fill[over]
                mp.pointcloud[over_dst] = sample_from_point(mp.pointcloud[over_src], mp.pointcloud_features[over_src],
                                                            self.generator)
                mp.pointcloud[over_src] = sample_from_point(mp.pointcloud[over_src], mp.pointcloud_features[over_src],
                                                            self.generator)
                mp.pointcloud[to_fill[under]] += info.densify_point_grad_position[:n_fill][under] * \
                    cfg.under_reconstructed_move_factor
            mp.point_invalid_mask[to_fill] = 0
        valid_after = int(mp.point_invalid_mask.shape[0] - mp.point_invalid_mask.sum())
        assert valid_after == valid_before - num_transparent - num_floaters + n_fill
        if self.verbose:
            print(f"total valid points: {valid_before} -> {valid_after}, densify {num_densify} (filled {n_fill}), "
                  f"transparent {num_transparent}, floaters {num_floaters}")
        self.densify_point_info = None

    # GaussianPointAdaptiveController.py:355-358
    def reset_alpha(self):
        f = self.maintained_parameters.pointcloud_features
        f[:, 7] = torch.clamp(f[:, 7], max=self.config.reset_alpha_value)
