"""The whole CUDA path executed on the CPU from the unmodified kernel sources (tests/simt, see simt_emu.h): the stages are
chained exactly as ``csrc/api.cu`` chains them -- preprocess -> radix sort -> tile ranges -> forward blend -> loop A of
the backward -> per-point chain rule -- through the library's own buffers.  Test infrastructure."""
import ctypes
import os
import subprocess
from types import SimpleNamespace

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SIMT = os.path.join(HERE, "simt")
CSRC = os.path.join(os.path.dirname(HERE), "taichi_3d_gaussian_splatting_b200", "csrc")


def build_emulator():
    out = os.path.join(SIMT, "libsimt_emu.so")
    tus = [os.path.join(SIMT, f) for f in ("emu_blend.cpp", "emu_preprocess.cpp", "emu_sort.cpp", "emu_image_loss.cpp", "emu_adam.cpp", "emu_controller.cpp")]
    deps = tus + [os.path.join(SIMT, "simt_emu.h")] + [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(s) for s in deps):
        cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
        subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-I", cuda_inc, "-o", out, *tus],
                       check=True)
    L = ctypes.CDLL(out)
    L.emu_blend_backward.restype = ctypes.c_longlong
    L.emu_blend_forward.restype = ctypes.c_longlong
    L.emu_preprocess.restype = ctypes.c_longlong
    L.emu_backward_points.restype = ctypes.c_longlong
    L.emu_expand_view_gradients.restype = ctypes.c_longlong
    L.emu_sort_pairs.restype = ctypes.c_longlong
    L.emu_sort_pairs_compacted.restype = ctypes.c_longlong
    L.emu_image_loss.restype = ctypes.c_longlong
    L.emu_image_loss_temp_bytes.restype = ctypes.c_longlong
    return L



def _bit_width(v):
    return int(v).bit_length()


def run_preprocess(emu, scene, fwd_cfg, key64, filter_tiles):
    """The fused per-point stage (csrc/preprocess.cu) under the emulator on a scene of CPU tensors; returns its raw buffers."""
    xyz = scene.point_cloud.numpy().astype(np.float32).copy()
    feats = scene.point_cloud_features.detach().numpy().astype(np.float32).copy()
    N = xyz.shape[0]
    ci = scene.camera_info
    H, W = ci.camera_height, ci.camera_width
    far, scale, near = fwd_cfg.get("far_plane", 1000.0), fwd_cfg.get("depth_to_sort_key_scale", 100.0), fwd_cfg.get("near_plane", 0.8)
    T = (H // 16) * (W // 16)
    tile_bits = _bit_width(max(T - 1, 0))
    mk = np.float32(far) * np.float32(scale)
    depth_bits = max(_bit_width(int(mk)), 1)  # csrc/api.cu compute_layout
    if key64 or tile_bits + depth_bits > 32:
        key_bytes, depth_bits = 8, 32
    else:
        key_bytes = 4
    cap = 64 * N + 4096
    counters = np.zeros(8, np.int64)
    point_id, point_offset, num_tiles = (np.full(N, -9, np.int32) for _ in range(3))
    records, pic = np.zeros((N, 12), np.float32), np.zeros((N, 3), np.float32)
    keys = np.zeros(cap, np.uint32 if key_bytes == 4 else np.uint64)
    vals = np.zeros(cap, np.int32)
    c = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    q = scene.q_pointcloud_camera.numpy().astype(np.float32).copy()
    t = scene.t_pointcloud_camera.numpy().astype(np.float32).copy()
    K = ci.camera_intrinsics.numpy().astype(np.float32).copy()
    inv = scene.point_invalid_mask.numpy().astype(np.int8).copy()
    obj = scene.point_object_id.numpy().astype(np.int32).copy()
    sw = emu.emu_preprocess(
        ctypes.c_longlong(N), c(xyz), c(feats), c(inv), c(obj), q.shape[0], c(q), c(t), c(K), W, H, ctypes.c_float(near),
        ctypes.c_float(far), ctypes.c_float(scale), depth_bits, key_bytes, int(filter_tiles), 0, ctypes.c_longlong(cap),
        c(counters), c(point_id), c(point_offset), c(num_tiles), c(records), c(pic), c(keys), c(vals))
    assert sw > 0
    return SimpleNamespace(feats=feats, counters=counters, point_id=point_id, point_offset=point_offset, num_tiles=num_tiles,
                           records=records, pic=pic, keys=keys, vals=vals, depth_bits=depth_bits, tile_bits=tile_bits, H=H, W=W, T=T)




def c(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def emu_sort(emu, keys, vals, end_bit):
    keys, vals = np.ascontiguousarray(keys), np.ascontiguousarray(vals, dtype=np.int32)
    ko, vo = np.empty_like(keys), np.empty_like(vals)
    if keys.shape[0]:
        assert emu.emu_sort_pairs(c(keys), c(vals), c(ko), c(vo), ctypes.c_longlong(keys.shape[0]), keys.dtype.itemsize, end_bit) > 0
    return ko, vo


def emu_sort_frame(emu, pre, Kk):
    """The frame pipeline's sort (csrc/sort.cu launch_sort): compacted digits from the frame's largest depth key, which the
    per-point kernel left in counters[4]."""
    keys, vals = np.ascontiguousarray(pre.keys[:Kk]), np.ascontiguousarray(pre.vals[:Kk], dtype=np.int32)
    ko, vo = np.empty_like(keys), np.empty_like(vals)
    max_key = np.array([int(pre.counters[4]) & 0xFFFFFFFF], np.int32)
    if Kk:
        assert emu.emu_sort_pairs_compacted(c(keys), c(vals), c(ko), c(vo), ctypes.c_longlong(Kk), keys.dtype.itemsize,
                                            pre.depth_bits, pre.tile_bits + pre.depth_bits, c(max_key)) > 0
    return ko, vo


def emulated_forward(emu, scene, cfg=None, exact=True, filter_tiles=True):
    """Forward of the CUDA path under the emulator; returns the saved-for-backward state with the outputs."""
    cfg = cfg or {}
    pre = run_preprocess(emu, scene, cfg, key64=False, filter_tiles=filter_tiles)
    M, Kk = int(pre.counters[0]), int(pre.counters[1])
    sk, sv = emu_sort_frame(emu, pre, Kk)
    order = np.argsort(pre.keys[:Kk], kind="stable")
    assert np.array_equal(sk, pre.keys[:Kk][order]) and np.array_equal(sv, pre.vals[:Kk][order])
    start, end = np.zeros(pre.T, np.int32), np.zeros(pre.T, np.int32)
    emu.emu_tile_ranges(c(sk), ctypes.c_longlong(Kk), sk.dtype.itemsize, pre.depth_bits, pre.T, c(start), c(end))
    tile = sk.astype(np.int64) >> pre.depth_bits
    assert np.array_equal(end - start, np.bincount(tile, minlength=pre.T)[:pre.T])
    H, W = pre.H, pre.W
    image, depth, acc = np.zeros((H, W, 3), np.float32), np.zeros((H, W), np.float32), np.zeros((H, W), np.float32)
    last, cnt = np.zeros((H, W), np.int32), np.zeros((H, W), np.int32)
    if Kk:
        emu.emu_blend_forward(0, int(exact), H, W, c(start), c(end), c(sv), c(pre.records), c(image), c(depth), c(acc), c(last),
                              c(cnt))
    return SimpleNamespace(pre=pre, M=M, K=Kk, start=start, end=end, sorted_vals=sv, image=image, depth=depth, acc_alpha=acc,
                           last_effective=last, count=cnt, exact=exact, scene=scene)


def emulated_backward(emu, st, grad_image, band=3, transposed=False, factors=(1.0, 0.5, 20.0, 5.0, 1.0), stats=True,
                      compact=False, controller=None):
    """Backward for a state of :func:`emulated_forward`: dense gradients + the hook tensors; ``compact``: the COMPACT
    per-point kernel instead (GSB_FLAG_COMPACT_GRADS): returns (grad_sum (N,12), grad_colour (N,3), None).
    ``controller``: six numpy accumulators (num_in_camera i32, num_pixels i32, vs_grad, vs_grad_avg, pos_grad (N,3),
    pos_grad_norm) updated by the kernel's fused controller epilogue."""
    pre, scene, M = st.pre, st.scene, st.M
    H, W = pre.H, pre.W
    g = np.ascontiguousarray(grad_image, dtype=np.float32)
    accum, mag = np.zeros((max(M, 1), 12), np.float32), np.zeros((H, W, 2), np.float32)
    if st.K:
        emu.emu_blend_backward(int(transposed), int(st.exact), int(stats), H, W, c(st.start), c(st.end), c(st.sorted_vals),
                               c(pre.records), c(g), c(st.acc_alpha), c(st.last_effective), c(accum), c(mag))
    N = pre.point_offset.shape[0]
    q = scene.q_pointcloud_camera.numpy().astype(np.float32).copy()
    t = scene.t_pointcloud_camera.numpy().astype(np.float32).copy()
    poses = np.zeros((q.shape[0], 20), np.float32)
    emu.emu_pose(q.shape[0], c(q), c(t), c(poses))
    xyz = scene.point_cloud.detach().numpy().astype(np.float32).copy()
    K = scene.camera_info.camera_intrinsics.numpy().astype(np.float32).copy()
    obj = scene.point_object_id.numpy().astype(np.int32).copy()
    gx, gf = np.full((N, 3), 7.0, np.float32), np.full((N, 56), 7.0, np.float32)  # every row must be overwritten
    f = ctypes.c_float
    if compact:
        gsum, gcol = np.full((N, 12), 7.0, np.float32), np.full((N, 3), 7.0, np.float32)
        emu.emu_backward_points(ctypes.c_longlong(N), c(pre.point_offset), c(pre.records), c(pre.pic), c(accum), c(poses), c(xyz),
                                c(pre.feats), c(obj), c(t), c(K), int(band) if band in (0, 1, 2) else 3, *(f(v) for v in factors),
                                None, None, c(gsum), c(gcol), *([None] * 6))
        return gsum, gcol, None
    emu.emu_backward_points(ctypes.c_longlong(N), c(pre.point_offset), c(pre.records), c(pre.pic), c(accum), c(poses), c(xyz),
                            c(pre.feats), c(obj), c(t), c(K), int(band) if band in (0, 1, 2) else 3, *(f(v) for v in factors),
                            c(gx), c(gf), None, None, *([c(a) for a in controller] if controller is not None else [None] * 6))
    ids = pre.point_id[:M]
    hook = SimpleNamespace(grad_point_in_camera=gx[ids], grad_pointfeatures_in_camera=gf[ids], grad_viewspace=accum[:M, 0:2].copy(),
                           magnitude_grad_viewspace=accum[:M, 9].copy(), magnitude_grad_viewspace_on_image=mag,
                           num_affected_pixels=np.round(accum[:M, 10]).astype(np.int32))
    return gx, gf, hook


def emulated_operator(emu, scene, grad_image, band=3, transposed=False, exact=True, cfg=None, filter_tiles=True,
                      factors=(1.0, 0.5, 20.0, 5.0, 1.0)):
    """Forward + backward of the CUDA path under the emulator.  ``scene``: CPU tensors with the operator's input fields;
    ``cfg``: near_plane / far_plane / depth_to_sort_key_scale; ``factors``: grad q, s, alpha, colour, high-order colour
    (GPCR:782-786).  Returns the outputs, the per-point stage tensors, the dense gradients and the hook tensors."""
    st = emulated_forward(emu, scene, cfg, exact, filter_tiles)
    gx, gf, hook = emulated_backward(emu, st, grad_image, band, transposed, factors)
    pre, M = st.pre, st.M
    ids = pre.point_id[:M]
    r = pre.records[:M]
    return SimpleNamespace(
        image=st.image, depth=st.depth, count=st.count, acc_alpha=st.acc_alpha, last_effective=st.last_effective,
        features_after_forward=pre.feats, point_id_in_camera_list=ids, num_overlap_tiles=pre.num_tiles[:M],
        point_uv=r[:, 0:2].copy(), point_uv_conic_and_rescale=r[:, 2:6].copy(), point_alpha_after_activation=r[:, 6].copy(),
        point_color=r[:, 8:11].copy(), point_radii=r[:, 11].copy(), point_in_camera=pre.pic[:M], grad_pointcloud=gx,
        grad_pointcloud_features=gf, hook=hook)


class EmulatedCudaRasterisationModule:
    """TEST HELPER: the emulated CUDA path behind the operator's module / autograd surface (CPU tensors), so that the
    behavioural tests and the optimisation-trajectory golden can run on the CUDA kernel sources without a GPU."""

    def __init__(self, config, backward_valid_point_hook=None, backward_impl="butterfly", exact_exp=True):
        import torch

        from taichi_3d_gaussian_splatting_b200 import GaussianPointCloudRasterisation as GPCR
        self.config, self.hook = config, backward_valid_point_hook
        emu = build_emulator()
        outer = self

        class _Fn(torch.autograd.Function):
            @staticmethod
            def forward(ctx, pc, feat, mask, obj, q, t, camera_info, band):
                scene = SimpleNamespace(point_cloud=pc.detach(), point_cloud_features=feat.detach(), point_invalid_mask=mask,
                                        point_object_id=obj, camera_info=camera_info, q_pointcloud_camera=q,
                                        t_pointcloud_camera=t)
                cfg = dict(near_plane=config.near_plane, far_plane=config.far_plane,
                           depth_to_sort_key_scale=config.depth_to_sort_key_scale)
                st = emulated_forward(emu, scene, cfg, exact=exact_exp)
                with torch.no_grad():  # in-place quaternion normalisation of the in-frustum rows, GPCR:264-266
                    feat.copy_(torch.from_numpy(st.pre.feats))
                st.scene.point_cloud_features = feat.detach()
                ctx.st, ctx.band = st, band
                image, depth, count = torch.from_numpy(st.image), torch.from_numpy(st.depth), torch.from_numpy(st.count)
                ctx.mark_non_differentiable(depth, count)
                return image, depth, count

            @staticmethod
            def backward(ctx, g_image, g_depth, g_count):
                st = ctx.st
                gx, gf, h = emulated_backward(
                    emu, st, g_image.contiguous().numpy(), ctx.band, transposed=backward_impl == "transposed",
                    factors=(config.grad_q_factor, config.grad_s_factor, config.grad_alpha_factor, config.grad_color_factor,
                             config.grad_high_order_color_factor))
                if outer.hook is not None:
                    pre, M = st.pre, st.M
                    outer.hook(GPCR.BackwardValidPointHookInput(
                        point_id_in_camera_list=torch.from_numpy(pre.point_id[:M].copy()),
                        grad_point_in_camera=torch.from_numpy(h.grad_point_in_camera),
                        grad_pointfeatures_in_camera=torch.from_numpy(h.grad_pointfeatures_in_camera),
                        grad_viewspace=torch.from_numpy(h.grad_viewspace),
                        magnitude_grad_viewspace=torch.from_numpy(h.magnitude_grad_viewspace),
                        magnitude_grad_viewspace_on_image=torch.from_numpy(h.magnitude_grad_viewspace_on_image),
                        num_overlap_tiles=torch.from_numpy(pre.num_tiles[:M].copy()),
                        num_affected_pixels=torch.from_numpy(h.num_affected_pixels),
                        point_depth=torch.from_numpy(pre.pic[:M, 2].copy()),
                        point_uv_in_camera=torch.from_numpy(pre.records[:M, 0:2].copy())))
                return torch.from_numpy(gx), torch.from_numpy(gf), None, None, None, None, None, None

        self._fn = _Fn

    def __call__(self, inp):
        return self._fn.apply(inp.point_cloud, inp.point_cloud_features, inp.point_invalid_mask, inp.point_object_id,
                              inp.q_pointcloud_camera, inp.t_pointcloud_camera, inp.camera_info, inp.color_max_sh_band)
