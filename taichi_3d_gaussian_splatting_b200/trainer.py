"""Training step around the operator (BASELINE config 5 harness).

A compact mirror of the reference loop ``GaussianPointCloudTrainer.train``
(``taichi_3d_gaussian_splatting/GaussianPointTrainer.py:118-267``) for in-memory datasets: two Adam
optimisers (features / positions, :126-129), exponential decay of the position LR every
``position_learning_rate_decay_interval`` iterations (:131-132, 182-183), image down-sampling schedule
4 -> 2 -> 1 with the crop-to-16 rule (:98-116, 139-148), SH band schedule ``it // interval`` (:164),
clamp + HWC->CHW + ``LossFunction`` (:168-175), controller ``refinement`` after the optimiser step (:194).
TensorBoard logging, the parquet/JSON dataset and validation image dumps are out of scope.
The rasteriser is injected (default: the CUDA operator) so that tests can run the identical loop with the
CPU oracle behind the same interface and compare PSNR trajectories.
"""
import math
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Tuple

import torch
import torch.nn.functional as F

from .Camera import CameraInfo
from .densification import GaussianPointAdaptiveController
from .GaussianPointCloudRasterisation import GaussianPointCloudRasterisation
from .loss import LossFunction

View = Tuple[torch.Tensor, torch.Tensor, torch.Tensor, CameraInfo]  # image (3,H,W) in [0,1], q (1,4), t (1,3), camera


def psnr(pred: torch.Tensor, target: torch.Tensor) -> float:
    mse = torch.mean((pred.clamp(0, 1) - target) ** 2).item()
    return float("inf") if mse == 0 else -10.0 * math.log10(mse)


def downsample_image_and_camera_info(image: torch.Tensor, camera_info: CameraInfo, downsample_factor: int):
    """GaussianPointTrainer.py:98-116: antialiased resize, crop to multiples of 16, scale fx fy cx cy."""
    h = camera_info.camera_height // downsample_factor
    w = camera_info.camera_width // downsample_factor
    image = F.interpolate(image[None], size=(h, w), mode="bilinear", antialias=True, align_corners=False)[0]
    w -= w % 16
    h -= h % 16
    image = image[:3, :h, :w].contiguous()
    K = camera_info.camera_intrinsics.clone()
    K[0, 0] /= downsample_factor
    K[1, 1] /= downsample_factor
    K[0, 2] /= downsample_factor
    K[1, 2] /= downsample_factor
    return image, CameraInfo(camera_intrinsics=K, camera_height=h, camera_width=w, camera_id=camera_info.camera_id)


@dataclass
class Scene:
    """The trainable tensors of ``GaussianPointCloudScene`` (GaussianPointCloudScene.py:25-60), no IO."""
    point_cloud: torch.Tensor  # (N,3) leaf, requires_grad
    point_cloud_features: torch.Tensor  # (N,56) leaf, requires_grad
    point_invalid_mask: torch.Tensor  # (N,) int8
    point_object_id: torch.Tensor  # (N,) int32


class GaussianPointCloudTrainer:
    @dataclass
    class TrainConfig:
        # defaults of GaussianPointTrainer.py:32-58
        num_iterations: int = 300000
        feature_learning_rate: float = 1e-3
        position_learning_rate: float = 1e-5
        position_learning_rate_decay_rate: float = 0.97
        position_learning_rate_decay_interval: int = 100
        increase_color_max_sh_band_interval: float = 1000.
        initial_downsample_factor: int = 4
        half_downsample_factor_interval: int = 250
        rasterisation_config: GaussianPointCloudRasterisation.GaussianPointCloudRasterisationConfig = field(
            default_factory=GaussianPointCloudRasterisation.GaussianPointCloudRasterisationConfig)
        adaptive_controller_config: GaussianPointAdaptiveController.GaussianPointAdaptiveControllerConfig = field(
            default_factory=GaussianPointAdaptiveController.GaussianPointAdaptiveControllerConfig)
        loss_function_config: LossFunction.LossFunctionConfig = field(
            default_factory=LossFunction.LossFunctionConfig)

    def __init__(self, config: "GaussianPointCloudTrainer.TrainConfig", scene: Scene, train_views: List[View],
                 rasterisation_factory: Optional[Callable] = None, generator: Optional[torch.Generator] = None,
                 fused_image_loss: bool = False, fused_adam: bool = False, fused_controller_update: bool = False,
                 fused_step: bool = False, shuffle_generator: Optional[torch.Generator] = None):
        """``fused_image_loss``: clamp + L1 + D-SSIM and their gradient in two CUDA kernels (``gsb200_image_loss``)
        instead of ~60 autograd kernels per step; same loss values (CUDA only).  ``fused_adam``: the two Adam updates as
        one kernel each (``optim.FusedAdam`` / ``gsb200_adam_step``) instead of torch's foreach path (CUDA only).
        ``fused_controller_update``: the controller's per-iteration accumulator update as one kernel
        (``gsb200_controller_update``) instead of ~15 torch launches (CUDA only).
        ``fused_step``: the WHOLE iteration as one library call (``gsb200_train_step``: forward, image loss, backward with the
        controller accumulators fused into its epilogue, both Adam updates; no autograd, no host wait; CUDA only).
        ``shuffle_generator``: a CPU ``torch.Generator``; the views are then visited in a fresh random permutation per
        epoch like the reference's shuffling DataLoader (GaussianPointTrainer.py:120-124) instead of in fixed order."""
        self.config = config
        self.fused_step = fused_step
        if fused_step and config.loss_function_config.enable_regularization:
            raise ValueError("fused_step does not implement the optional scale regulariser (LossFunction.py:33-37)")
        self.fused_image_loss = fused_image_loss
        self.fused_adam = fused_adam
        self.scene = scene
        self.train_views = train_views
        self.adaptive_controller = GaussianPointAdaptiveController(
            config=config.adaptive_controller_config,
            maintained_parameters=GaussianPointAdaptiveController.GaussianPointAdaptiveControllerMaintainedParameters(
                pointcloud=scene.point_cloud, pointcloud_features=scene.point_cloud_features,
                point_invalid_mask=scene.point_invalid_mask, point_object_id=scene.point_object_id),
            generator=generator, fused_update=fused_controller_update)
        factory = rasterisation_factory or GaussianPointCloudRasterisation
        self.rasterisation = factory(config=config.rasterisation_config,
                                     backward_valid_point_hook=self.adaptive_controller.update)
        self.loss_function = LossFunction(config=config.loss_function_config)
        self.history: List[dict] = []
        self._downsampled = {}
        self._view_generator = shuffle_generator
        self._view_order = None

    def _input(self, q, t, camera_info, band):
        s = self.scene
        return GaussianPointCloudRasterisation.GaussianPointCloudRasterisationInput(
            point_cloud=s.point_cloud, point_cloud_features=s.point_cloud_features,
            point_object_id=s.point_object_id, point_invalid_mask=s.point_invalid_mask,
            camera_info=camera_info, q_pointcloud_camera=q, t_pointcloud_camera=t, color_max_sh_band=band)

    def _train_fused(self, log_interval: int = 0):
        """The loop of ``train`` with the whole iteration enqueued by ONE library call (``fused_step.FusedTrainStep``):
        no autograd graph, no host wait, the controller's accumulators updated inside the backward kernel."""
        from .fused_step import FusedTrainStep
        cfg = self.config
        step = FusedTrainStep(self.scene, cfg.rasterisation_config, cfg.loss_function_config.lambda_value,
                              controller=self.adaptive_controller)
        self.fused_train_step = step
        position_lr = cfg.position_learning_rate
        downsample_factor = cfg.initial_downsample_factor
        for iteration in range(cfg.num_iterations):
            if iteration % cfg.half_downsample_factor_interval == 0 and iteration > 0 and downsample_factor > 1:
                downsample_factor //= 2
            view_index = self._next_view_index(iteration)
            image_gt, q, t, camera_info = self.train_views[view_index]
            if downsample_factor > 1:
                key = (view_index, downsample_factor)
                if key not in self._downsampled:
                    self._downsampled[key] = downsample_image_and_camera_info(image_gt, camera_info, downsample_factor)
                image_gt, camera_info = self._downsampled[key]
            band = iteration // cfg.increase_color_max_sh_band_interval
            step.run(image_gt, q, t, camera_info, band, cfg.feature_learning_rate, position_lr)
            if iteration % cfg.position_learning_rate_decay_interval == 0:  # ExponentialLR.step() after the optimiser step
                position_lr *= cfg.position_learning_rate_decay_rate
            self.adaptive_controller.after_fused_update(step.hook_input)
            self.adaptive_controller.refinement()
            if log_interval and iteration % log_interval == 0:
                losses = step.loss.tolist()
                self.history.append(dict(iteration=iteration, loss=losses[0], l1=losses[1],
                                         psnr=psnr(step.image.detach().clamp(0, 1).permute(2, 0, 1), image_gt),
                                         num_valid_points=int((self.scene.point_invalid_mask == 0).sum())))
        return self.history

    def _next_view_index(self, iteration: int) -> int:
        """A fresh random permutation of the views every epoch (the reference draws them from a DataLoader with shuffle=True,
        GaussianPointTrainer.py:120-124), seeded from the injected generator; without a generator: the fixed order
        ``iteration % len(views)`` (deterministic tests and golden trajectories)."""
        n = len(self.train_views)
        if self._view_generator is None:
            return iteration % n
        k = iteration % n
        if k == 0 or self._view_order is None:
            self._view_order = torch.randperm(n, generator=self._view_generator).tolist()
        return self._view_order[k]

    def train(self, log_interval: int = 0):
        if self.fused_step:
            return self._train_fused(log_interval)
        cfg = self.config
        if self.fused_adam:
            from .optim import FusedAdam as Adam
        else:
            Adam = torch.optim.Adam
        optimizer = Adam([self.scene.point_cloud_features], lr=cfg.feature_learning_rate, betas=(0.9, 0.999))
        position_optimizer = Adam([self.scene.point_cloud], lr=cfg.position_learning_rate, betas=(0.9, 0.999))
        scheduler = torch.optim.lr_scheduler.ExponentialLR(position_optimizer, gamma=cfg.position_learning_rate_decay_rate)
        downsample_factor = cfg.initial_downsample_factor
        for iteration in range(cfg.num_iterations):
            if iteration % cfg.half_downsample_factor_interval == 0 and iteration > 0 and downsample_factor > 1:
                downsample_factor //= 2
            optimizer.zero_grad()
            position_optimizer.zero_grad()
            view_index = self._next_view_index(iteration)
            image_gt, q, t, camera_info = self.train_views[view_index]
            if downsample_factor > 1:
                # the reference resizes the full-resolution frame in every iteration (GaussianPointTrainer.py:146-148); the
                # result depends only on (view, factor), so it is computed once per pair (same tensors, ~10 launches less)
                key = (view_index, downsample_factor)
                if key not in self._downsampled:
                    self._downsampled[key] = downsample_image_and_camera_info(image_gt, camera_info, downsample_factor)
                image_gt, camera_info = self._downsampled[key]
            band = iteration // cfg.increase_color_max_sh_band_interval
            image_pred, _, _ = self.rasterisation(self._input(q, t, camera_info, band))
            if self.fused_image_loss:
                loss, l1_loss, ssim_loss = self.loss_function.forward_rasterized(
                    image_pred, image_gt, point_invalid_mask=self.scene.point_invalid_mask,
                    pointcloud_features=self.scene.point_cloud_features)
                image_pred = image_pred.detach().clamp(0, 1).permute(2, 0, 1) if log_interval else image_pred
            else:
                image_pred = torch.clamp(image_pred, min=0, max=1).permute(2, 0, 1)
                loss, l1_loss, ssim_loss = self.loss_function(
                    image_pred, image_gt, point_invalid_mask=self.scene.point_invalid_mask,
                    pointcloud_features=self.scene.point_cloud_features)
            loss.backward()
            optimizer.step()
            position_optimizer.step()
            if iteration % cfg.position_learning_rate_decay_interval == 0:
                scheduler.step()
            self.adaptive_controller.refinement()
            if log_interval and iteration % log_interval == 0:
                self.history.append(dict(iteration=iteration, loss=float(loss.detach()), l1=float(l1_loss.detach()),
                                         psnr=psnr(image_pred.detach(), image_gt),
                                         num_valid_points=int((self.scene.point_invalid_mask == 0).sum())))
        return self.history

    @torch.no_grad()
    def validation(self, views: Optional[List[View]] = None) -> float:
        """Mean PSNR over the views at full resolution (GaussianPointTrainer.py:334-415 without the logging)."""
        views = views if views is not None else self.train_views
        total = 0.0
        for image_gt, q, t, camera_info in views:
            image_pred, _, _ = self.rasterisation(self._input(q, t, camera_info, 3))
            total += psnr(image_pred.permute(2, 0, 1), image_gt)
        return total / max(len(views), 1)
