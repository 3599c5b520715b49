"""TEST HELPER: the CPU oracle behind the operator's nn.Module / autograd surface (CPU tensors), so that the
trainer harness can run the identical loop with the oracle and with the CUDA operator."""
import numpy as np
import torch

from oracle import OracleRasterisation
from taichi_3d_gaussian_splatting_b200 import GaussianPointCloudRasterisation as GPCR


class OracleRasterisationModule(torch.nn.Module):
    def __init__(self, config, backward_valid_point_hook=None):
        super().__init__()
        self.config = config
        self.hook = backward_valid_point_hook
        self.oracle = OracleRasterisation(near_plane=config.near_plane, far_plane=config.far_plane,
                                          depth_to_sort_key_scale=config.depth_to_sort_key_scale,
                                          rgb_only=config.rgb_only)
        outer = self

        class _Fn(torch.autograd.Function):
            @staticmethod
            def forward(ctx, pc, feat, mask, obj, q, t, camera_info, band):
                feat_np = feat.detach().numpy()  # shares memory: the oracle normalises q in place
                fwd = outer.oracle.forward(pc.detach().numpy(), feat_np, mask.numpy(), obj.numpy(),
                                           camera_info.camera_intrinsics.numpy(), camera_info.camera_height,
                                           camera_info.camera_width, q.numpy(), t.numpy())
                ctx.fwd, ctx.band, ctx.K = fwd, band, camera_info.camera_intrinsics
                ctx.save_for_backward(pc, feat, obj, t)
                image = torch.from_numpy(fwd.image)
                depth = torch.from_numpy(fwd.depth)
                count = torch.from_numpy(fwd.pixel_valid_point_count)
                ctx.mark_non_differentiable(depth, count)
                return image, depth, count

            @staticmethod
            def backward(ctx, g_image, g_depth, g_count):
                pc, feat, obj, t = ctx.saved_tensors
                band = ctx.band if ctx.band in (0, 1, 2) else 3
                b = outer.oracle.backward(ctx.fwd, g_image.contiguous().numpy(), pc.detach().numpy(),
                                          feat.detach().numpy(), obj.numpy(), ctx.K.numpy(), t.numpy(), int(band))
                if outer.hook is not None:
                    outer.hook(GPCR.BackwardValidPointHookInput(
                        point_id_in_camera_list=torch.from_numpy(b.point_id_in_camera_list),
                        grad_point_in_camera=torch.from_numpy(b.grad_point_in_camera),
                        grad_pointfeatures_in_camera=torch.from_numpy(b.grad_pointfeatures_in_camera),
                        grad_viewspace=torch.from_numpy(b.grad_viewspace),
                        magnitude_grad_viewspace=torch.from_numpy(b.magnitude_grad_viewspace),
                        magnitude_grad_viewspace_on_image=torch.from_numpy(b.magnitude_grad_viewspace_on_image),
                        num_overlap_tiles=torch.from_numpy(b.num_overlap_tiles),
                        num_affected_pixels=torch.from_numpy(b.num_affected_pixels),
                        point_depth=torch.from_numpy(b.point_depth),
                        point_uv_in_camera=torch.from_numpy(b.point_uv_in_camera)))
                return (torch.from_numpy(b.grad_pointcloud), torch.from_numpy(b.grad_pointcloud_features),
                        None, None, None, None, None, None)

        self._fn = _Fn

    def forward(self, inp):
        return self._fn.apply(inp.point_cloud, inp.point_cloud_features, inp.point_invalid_mask,
                              inp.point_object_id, inp.q_pointcloud_camera, inp.t_pointcloud_camera,
                              inp.camera_info, inp.color_max_sh_band)
