"""Per-kernel device times of one frame, measured with CUDA events that the library records on the
launching stream between its own kernel launches (``gsb200_forward_timed`` / ``gsb200_backward_timed``).
Used by ``bench.py`` for the roofline object; not part of the operator's hot path.
"""
import ctypes
from typing import Dict

import torch

from . import _lib
from .GaussianPointCloudRasterisation import GaussianPointCloudRasterisation, _ptr

FORWARD_STAGES = ("memset", "preprocess", "sort", "tile_ranges", "blend_forward")
BACKWARD_STAGES = ("memset_grads", "blend_backward", "backward_points")


def KERNELS_PER_FORWARD(sort_passes: int) -> int:
    # pose + preprocess + histogram + one kernel per radix pass + tile ranges + blend
    return 2 + 1 + sort_passes + 1 + 1


KERNELS_PER_BACKWARD = 2  # blend backward + per-point chain rule (memsets are driver fills, not counted)


def _frame_args(op: GaussianPointCloudRasterisation, input_data, grad_image: torch.Tensor):
    """Argument blocks of one frame of ``input_data`` on buffers of their own (sized by a first run of the operator)."""
    cfg = op.config
    pc, feat = input_data.point_cloud.detach(), input_data.point_cloud_features.detach()
    ci = input_data.camera_info
    H, W, N = ci.camera_height, ci.camera_width, pc.shape[0]
    device = pc.device
    with torch.no_grad():
        op(input_data)  # sizes the key capacity
    frame = op.last_frame
    M = frame.num_points_in_camera
    layout = frame.layout
    n_obj = input_data.q_pointcloud_camera.shape[0]
    with torch.cuda.device(device):
        stream = torch.cuda.current_stream(device).cuda_stream
        ws = torch.empty((layout.total_bytes,), dtype=torch.uint8, device=device)
        image = torch.empty((H, W, 3), device=device)
        depth = torch.empty((H, W), device=device)
        acc = torch.empty((H, W), device=device)
        last = torch.empty((H, W), dtype=torch.int32, device=device)
        cnt = torch.empty((H, W), dtype=torch.int32, device=device)
        gx, gf = torch.empty_like(pc), torch.empty_like(feat)
        accum = torch.empty((max(M, 1), 12), device=device)
        mag = torch.empty((H, W, 2), device=device)
        K = ci.camera_intrinsics.contiguous()
        q = input_data.q_pointcloud_camera.contiguous()
        t = input_data.t_pointcloud_camera.contiguous()
        g = grad_image.contiguous()
        fa = _lib.GsbForwardArgs(
            num_points=N, pointcloud=_ptr(pc), pointcloud_features=_ptr(feat),
            point_invalid_mask=_ptr(input_data.point_invalid_mask), point_object_id=_ptr(input_data.point_object_id),
            num_objects=n_obj, q_pointcloud_camera=_ptr(q), t_pointcloud_camera=_ptr(t), camera_intrinsics=_ptr(K),
            camera_height=H, camera_width=W, near_plane=cfg.near_plane, far_plane=cfg.far_plane,
            depth_to_sort_key_scale=cfg.depth_to_sort_key_scale, rgb_only=1 if cfg.rgb_only else 0,
            flags=frame.flags, workspace=_ptr(ws), workspace_bytes=layout.total_bytes,
            key_capacity=frame.key_capacity, rasterized_image=_ptr(image), rasterized_depth=_ptr(depth),
            pixel_accumulated_alpha=_ptr(acc), pixel_offset_of_last_effective_point=_ptr(last),
            pixel_valid_point_count=_ptr(cnt), stream=stream)
        ba = _lib.GsbBackwardArgs(
            num_points=N, pointcloud=_ptr(pc), pointcloud_features=_ptr(feat),
            point_object_id=_ptr(input_data.point_object_id), num_objects=n_obj, t_pointcloud_camera=_ptr(t),
            camera_intrinsics=_ptr(K), camera_height=H, camera_width=W, far_plane=cfg.far_plane,
            depth_to_sort_key_scale=cfg.depth_to_sort_key_scale, color_max_sh_band=3,
            grad_q_factor=cfg.grad_q_factor, grad_s_factor=cfg.grad_s_factor,
            grad_alpha_factor=cfg.grad_alpha_factor, grad_color_factor=cfg.grad_color_factor,
            grad_high_order_color_factor=cfg.grad_high_order_color_factor, flags=op.backward_flags(frame.flags),
            workspace=_ptr(ws), workspace_bytes=layout.total_bytes, key_capacity=frame.key_capacity,
            grad_rasterized_image=_ptr(g), pixel_accumulated_alpha=_ptr(acc),
            pixel_offset_of_last_effective_point=_ptr(last), accum=_ptr(accum), accum_rows=M,
            grad_pointcloud=_ptr(gx), grad_pointcloud_features=_ptr(gf),
            magnitude_grad_viewspace_on_image=_ptr(mag), stream=stream)
    keepalive = (ws, image, depth, acc, last, cnt, gx, gf, accum, mag, K, q, t, g, pc, feat)
    return fa, ba, keepalive


def stage_times(op: GaussianPointCloudRasterisation, input_data, grad_image: torch.Tensor,
                iters: int = 5) -> Dict[str, float]:
    """Average device milliseconds per stage over ``iters`` frames of ``input_data``."""
    lib = _lib.load()
    lib.gsb200_forward_timed.argtypes = [ctypes.POINTER(_lib.GsbForwardArgs), ctypes.POINTER(ctypes.c_float)]
    lib.gsb200_forward_timed.restype = ctypes.c_int
    lib.gsb200_backward_timed.argtypes = [ctypes.POINTER(_lib.GsbBackwardArgs), ctypes.POINTER(ctypes.c_float)]
    lib.gsb200_backward_timed.restype = ctypes.c_int
    fa, ba, keepalive = _frame_args(op, input_data, grad_image)
    totals = {k: 0.0 for k in FORWARD_STAGES + BACKWARD_STAGES}
    with torch.cuda.device(input_data.point_cloud.device):
        fms = (ctypes.c_float * 8)()
        bms = (ctypes.c_float * 8)()
        for it in range(iters + 1):
            _lib.check(lib.gsb200_forward_timed(ctypes.byref(fa), fms), "gsb200_forward_timed")
            _lib.check(lib.gsb200_backward_timed(ctypes.byref(ba), bms), "gsb200_backward_timed")
            if it == 0:
                continue  # warm-up
            for i, name in enumerate(FORWARD_STAGES):
                totals[name] += fms[i]
            for i, name in enumerate(BACKWARD_STAGES):
                totals[name] += bms[i]
    del keepalive
    return {k: v / iters for k, v in totals.items()}


def blend_work(op: GaussianPointCloudRasterisation, input_data, grad_image: torch.Tensor) -> Dict[str, int]:
    """The blend kernels' real work for one frame of ``input_data``, counted on the device by the COUNT instantiations of
    the two kernels (``gsb200_forward_blend_work`` / ``gsb200_backward_blend_work``): (warp, splat) visits (32 pixel x splat
    evaluations each) and contributing evaluations -- SURVEY 8(d)'s "E" as a measurement."""
    lib = _lib.load()
    fa, ba, keepalive = _frame_args(op, input_data, grad_image)
    out = {}
    with torch.cuda.device(input_data.point_cloud.device):
        _lib.check(lib.gsb200_forward(ctypes.byref(fa)), "gsb200_forward")
        f = (ctypes.c_uint64 * 8)()
        _lib.check(lib.gsb200_forward_blend_work(ctypes.byref(fa), f), "gsb200_forward_blend_work")
        out["forward_warp_splat_visits"], out["forward_contributing_evaluations"] = int(f[0]), int(f[1])
        # what-if at staging time: (patch, splat) pairs with 8x4 patches (the kernel's), 8x8 (two pixels per thread), 16x4
        out["staged_patch_pairs_8x4"], out["staged_patch_pairs_8x8"], out["staged_patch_pairs_16x4"] = int(f[2]), int(f[3]), int(f[4])
        out["staged_patch_pairs_4x4"] = int(f[5])
        ba.flags = (ba.flags | _lib.GSB_FLAG_BACKWARD_TRANSPOSED | _lib.GSB_FLAG_NO_HOOK_STATS) & ~_lib.GSB_FLAG_EXACT_EXP
        keepalive[8].zero_()  # accum
        b = (ctypes.c_uint64 * 2)()
        _lib.check(lib.gsb200_backward_blend_work(ctypes.byref(ba), b), "gsb200_backward_blend_work")
        out["backward_warp_splat_visits"], out["backward_contributing_evaluations"] = int(b[0]), int(b[1])
    del keepalive
    return out
