"""``FusedTrainStep``: one training iteration of the reference loop (``GaussianPointTrainer.py:138-180``) as ONE call into
``libgsb200.so`` (``gsb200_train_step``): forward, clamp + L1 + D-SSIM loss and its gradient, backward with the densification
controller's accumulators updated in the epilogue of the per-point kernel, Adam on the features and on the positions --
about 15 kernels enqueued back to back with no host wait, no autograd graph and no per-iteration allocation.  The Python
trainer around it (``trainer.GaussianPointCloudTrainer(..., fused_step=True)``) only picks the view, computes the two
learning rates and runs the controller's ``refinement`` every ``num_iterations_densify`` iterations.

What the host does NOT know any more: the number of in-camera points M and of (tile, splat) pairs K of a frame (the operator
waits for them once per frame; here nothing waits).  The accumulator buffer therefore has N rows, and a frame that needs more
pairs than the key capacity turns itself into a no-op ON THE DEVICE (overflow counter checked by the accumulator update and
both Adam kernels); the host reads the counters of iteration i while it prepares iteration i+1, grows the capacity and counts
the skipped iteration in ``num_skipped_steps``.  CUDA only; there is no CPU path.
"""
import ctypes
import warnings
from types import SimpleNamespace
from typing import Optional

import torch

from . import _lib
from .GaussianPointCloudRasterisation import Frame, GaussianPointCloudRasterisation, _ptr


class FusedTrainStep:
    def __init__(self, scene, rasterisation_config, lambda_value: float = 0.2, controller=None, betas=(0.9, 0.999),
                 eps: float = 1e-8, key_capacity: Optional[int] = None):
        """``scene``: object with ``point_cloud`` (N,3), ``point_cloud_features`` (N,56), ``point_invalid_mask``,
        ``point_object_id`` (CUDA, contiguous; updated in place).  ``controller``: a ``GaussianPointAdaptiveController`` whose
        six accumulators are updated by the backward epilogue (or ``None``)."""
        self.scene = scene
        self.config = rasterisation_config
        self.lambda_value = float(lambda_value)
        self.controller = controller
        self.betas, self.eps = betas, float(eps)
        pc = scene.point_cloud
        if not pc.is_cuda:
            raise RuntimeError("FusedTrainStep needs CUDA tensors: there is no CPU path")
        self.device = pc.device
        self.N = pc.shape[0]
        self.key_capacity = int(key_capacity) if key_capacity else max(1 << 20, 8 * self.N)
        self.step_count = 0
        self.num_skipped_steps = 0
        self._lib = _lib.load()
        dev, N = self.device, self.N
        z = lambda *shape: torch.zeros(shape, dtype=torch.float32, device=dev)  # noqa: E731
        self.feature_exp_avg, self.feature_exp_avg_sq = z(N, 56), z(N, 56)
        self.position_exp_avg, self.position_exp_avg_sq = z(N, 3), z(N, 3)
        self.accum = torch.empty((max(N, 1), 12), dtype=torch.float32, device=dev)
        off = (3 * N + 3) // 4 * 4
        self._flat = z(off + 56 * N)
        self.grad_pointcloud = self._flat[:3 * N].view(N, 3)
        self.grad_pointcloud_features = self._flat[off:off + 56 * N].view(N, 56)
        self.loss = z(3)  # {loss, L1, 1 - SSIM} of the latest iteration (device)
        self._res = {}
        self._pinned = [torch.zeros(4, dtype=torch.int64).pin_memory() for _ in range(2)]
        self._events = [torch.cuda.Event() for _ in range(2)]
        for e in self._events:
            e.record()
        self._pending = [False, False]
        self._last = None

    # ------------------------------------------------------------------ per-resolution buffers
    def _buffers(self, H, W, n_obj):
        key = (H, W, n_obj, self.key_capacity)
        b = self._res.get(key)
        if b is None:
            cfg, dev = self.config, self.device
            layout = _lib.workspace_layout(self.N, n_obj, self.key_capacity, H, W, cfg.far_plane, cfg.depth_to_sort_key_scale, 0)
            e = lambda shape, dt=torch.float32: torch.empty(shape, dtype=dt, device=dev)  # noqa: E731
            temp_bytes = int(self._lib.gsb200_image_loss_temp_bytes(H, W))
            b = SimpleNamespace(layout=layout, ws=e((layout.total_bytes,), torch.uint8), image=e((H, W, 3)), depth=e((H, W)),
                                acc_alpha=e((H, W)), last_effective=e((H, W), torch.int32), count=e((H, W), torch.int32),
                                grad_image=e((H, W, 3)), mag_image=e((H, W, 2)),
                                loss_temp=torch.zeros((temp_bytes + 15) // 16 * 16, dtype=torch.uint8, device=dev),
                                temp_bytes=temp_bytes)
            self._res[key] = b
        return b

    def _check_previous(self):
        """Counters of the iteration before the latest one (their copy finished long ago): overflow -> grow and count."""
        slot = self.step_count % 2
        if not self._pending[slot]:
            return
        self._events[slot].synchronize()
        self._pending[slot] = False
        if int(self._pinned[slot][2]) != 0:
            needed = int(self._pinned[slot][1])
            self.key_capacity = int(needed * 1.25) + 4096
            self.num_skipped_steps += 1
            warnings.warn(f"FusedTrainStep: a frame needed {needed} (tile, splat) pairs; that iteration was a no-op on the device, "
                          f"the key capacity is now {self.key_capacity}")

    # ------------------------------------------------------------------ one iteration
    def run(self, image_gt: torch.Tensor, q_pointcloud_camera: torch.Tensor, t_pointcloud_camera: torch.Tensor, camera_info,
            color_max_sh_band: int, feature_learning_rate: float, position_learning_rate: float) -> None:
        sc, cfg = self.scene, self.config
        H, W = int(camera_info.camera_height), int(camera_info.camera_width)
        if image_gt.shape != (3, H, W) or not image_gt.is_contiguous() or image_gt.dtype != torch.float32:
            raise ValueError(f"image_gt must be a contiguous float32 (3, {H}, {W}) tensor")
        self._check_previous()
        q, t = q_pointcloud_camera.contiguous(), t_pointcloud_camera.contiguous()
        K = camera_info.camera_intrinsics.contiguous()
        n_obj = q.shape[0]
        b = self._buffers(H, W, n_obj)
        ctl = self.controller
        band = int(color_max_sh_band) if color_max_sh_band in (0, 1, 2) else 3
        self.step_count += 1
        slot = self.step_count % 2
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            fwd = _lib.GsbForwardArgs(
                num_points=self.N, pointcloud=_ptr(sc.point_cloud), pointcloud_features=_ptr(sc.point_cloud_features),
                point_invalid_mask=_ptr(sc.point_invalid_mask), point_object_id=_ptr(sc.point_object_id), num_objects=n_obj,
                q_pointcloud_camera=_ptr(q), t_pointcloud_camera=_ptr(t), camera_intrinsics=_ptr(K), camera_height=H,
                camera_width=W, near_plane=cfg.near_plane, far_plane=cfg.far_plane,
                depth_to_sort_key_scale=cfg.depth_to_sort_key_scale, rgb_only=0, flags=0, workspace=_ptr(b.ws),
                workspace_bytes=b.layout.total_bytes, key_capacity=self.key_capacity, rasterized_image=_ptr(b.image),
                rasterized_depth=_ptr(b.depth), pixel_accumulated_alpha=_ptr(b.acc_alpha),
                pixel_offset_of_last_effective_point=_ptr(b.last_effective), pixel_valid_point_count=_ptr(b.count), stream=stream,
                host_counters=self._pinned[slot].data_ptr(), host_counters_event=self._events[slot].cuda_event)
            flags = _lib.GSB_FLAG_BACKWARD_TRANSPOSED | (0 if ctl is not None else _lib.GSB_FLAG_NO_HOOK_STATS)
            bwd = _lib.GsbBackwardArgs(
                num_points=self.N, pointcloud=_ptr(sc.point_cloud), pointcloud_features=_ptr(sc.point_cloud_features),
                point_object_id=_ptr(sc.point_object_id), num_objects=n_obj, t_pointcloud_camera=_ptr(t),
                camera_intrinsics=_ptr(K), camera_height=H, camera_width=W, far_plane=cfg.far_plane,
                depth_to_sort_key_scale=cfg.depth_to_sort_key_scale, color_max_sh_band=band, grad_q_factor=cfg.grad_q_factor,
                grad_s_factor=cfg.grad_s_factor, grad_alpha_factor=cfg.grad_alpha_factor, grad_color_factor=cfg.grad_color_factor,
                grad_high_order_color_factor=cfg.grad_high_order_color_factor, flags=flags, workspace=_ptr(b.ws),
                workspace_bytes=b.layout.total_bytes, key_capacity=self.key_capacity, grad_rasterized_image=_ptr(b.grad_image),
                pixel_accumulated_alpha=_ptr(b.acc_alpha), pixel_offset_of_last_effective_point=_ptr(b.last_effective),
                accum=_ptr(self.accum), accum_rows=self.N, grad_pointcloud=_ptr(self.grad_pointcloud),
                grad_pointcloud_features=_ptr(self.grad_pointcloud_features), magnitude_grad_viewspace_on_image=_ptr(b.mag_image),
                stream=stream)
            if ctl is not None:
                bwd.ctl_accumulated_num_in_camera = _ptr(ctl.accumulated_num_in_camera)
                bwd.ctl_accumulated_num_pixels = _ptr(ctl.accumulated_num_pixels)
                bwd.ctl_accumulated_view_space_position_gradients = _ptr(ctl.accumulated_view_space_position_gradients)
                bwd.ctl_accumulated_view_space_position_gradients_avg = _ptr(ctl.accumulated_view_space_position_gradients_avg)
                bwd.ctl_accumulated_position_gradients = _ptr(ctl.accumulated_position_gradients)
                bwd.ctl_accumulated_position_gradients_norm = _ptr(ctl.accumulated_position_gradients_norm)
            args = _lib.GsbTrainStepArgs(
                forward=fwd, backward=bwd, ground_truth_image=_ptr(image_gt), lambda_value=self.lambda_value,
                loss_out3=_ptr(self.loss), loss_temp=_ptr(b.loss_temp), loss_temp_bytes=b.temp_bytes,
                feature_exp_avg=_ptr(self.feature_exp_avg), feature_exp_avg_sq=_ptr(self.feature_exp_avg_sq),
                position_exp_avg=_ptr(self.position_exp_avg), position_exp_avg_sq=_ptr(self.position_exp_avg_sq),
                feature_learning_rate=float(feature_learning_rate), position_learning_rate=float(position_learning_rate),
                beta1=float(self.betas[0]), beta2=float(self.betas[1]), eps=self.eps, step=self.step_count)
            _lib.check(self._lib.gsb200_train_step(ctypes.byref(args)), "gsb200_train_step")
        self._pending[slot] = True
        self._last = SimpleNamespace(buffers=b, H=H, W=W, slot=slot, keep=(q, t, K, image_gt))

    # ------------------------------------------------------------------ the latest frame, on demand (these calls wait)
    @property
    def image(self) -> torch.Tensor:
        """(H,W,3) rasterised image of the latest iteration (before the clamp)."""
        return self._last.buffers.image

    def hook_input(self) -> "GaussianPointCloudRasterisation.BackwardValidPointHookInput":
        """The latest iteration's ``BackwardValidPointHookInput`` (GPCR:806-817): built only when the controller looks for
        densification candidates (every ``num_iterations_densify`` iterations), never on the per-iteration path."""
        last = self._last
        self._events[last.slot].synchronize()
        pinned = self._pinned[last.slot]
        b = last.buffers
        frame = Frame(b.ws, b.layout, self.N, self.key_capacity, last.H, last.W, 0)
        frame.num_points_in_camera, frame.num_keys = int(pinned[0]), int(pinned[1])
        M = frame.num_points_in_camera
        ids = frame.point_id_in_camera_list
        ids64 = ids.long()
        acc = self.accum[:M]
        return GaussianPointCloudRasterisation.BackwardValidPointHookInput(
            point_id_in_camera_list=ids, grad_point_in_camera=self.grad_pointcloud[ids64],
            grad_pointfeatures_in_camera=self.grad_pointcloud_features[ids64], grad_viewspace=acc[:, 0:2].contiguous(),
            magnitude_grad_viewspace=acc[:, 9].contiguous(), magnitude_grad_viewspace_on_image=b.mag_image,
            num_overlap_tiles=frame.num_overlap_tiles, num_affected_pixels=acc[:, 10].round().to(torch.int32),
            point_uv_in_camera=frame.point_uv.contiguous(), point_depth=frame.point_in_camera[:, 2])
