"""ctypes/numpy front-end of the CPU oracle (``gs_oracle.c``).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  The orchestration below follows the
reference operator step by step:

* forward  = GaussianPointCloudRasterisation.py:830-1023 (K1, P1, K2, K3, P2, K4, P3, K5, K6)
* backward = GaussianPointCloudRasterisation.py:1025-1163 (K7 loop A + loop B, P4, P5 hook tensors)

The device primitives the reference borrows from torch (mask compaction GPCR:861-864, cumsum
GPCR:913-922, sort GPCR:947-950) are restated with numpy / a stable C radix sort.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libgs_oracle.so")
_lib = None

TILE_WIDTH = 16
TILE_HEIGHT = 16


def build_oracle(force: bool = False) -> str:
    """Compile ``gs_oracle.c`` with the committed Makefile (gcc, strict IEEE f32)."""
    src = os.path.join(_HERE, "gs_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s", "-B", "libgs_oracle.so"], check=True)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build_oracle()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.gso_density_2d_normalized.restype = ctypes.c_float
        _lib.gso_density_from_conic.restype = ctypes.c_float
        _lib.gso_num_threads.restype = ctypes.c_int
    return _lib


def _p(a: np.ndarray):
    assert a.flags["C_CONTIGUOUS"], "oracle needs C-contiguous arrays"
    return a.ctypes.data_as(ctypes.c_void_p)


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


# --------------------------------------------------------------------------- host helpers
def _quat_mul(q0, q1):
    x0, y0, z0, w0 = q0[..., 0], q0[..., 1], q0[..., 2], q0[..., 3]
    x1, y1, z1, w1 = q1[..., 0], q1[..., 1], q1[..., 2], q1[..., 3]
    x = w0 * x1 + x0 * w1 + y0 * z1 - z0 * y1
    y = w0 * y1 - x0 * z1 + y0 * w1 + z0 * x1
    z = w0 * z1 + x0 * y1 - y0 * x1 + z0 * w1
    w = w0 * w1 - x0 * x1 - y0 * y1 - z0 * z1
    return np.stack([x, y, z, w], axis=-1)


def inverse_SE3_qt(q: np.ndarray, t: np.ndarray):
    """utils.py:396-432 ``inverse_SE3_qt_torch`` in float32 numpy (q = xyzw, batch first)."""
    q = _f32(q)
    t = _f32(t)
    q_inv = np.concatenate([-q[..., 0:3], q[..., 3:4]], axis=-1)
    qn = q_inv / np.sqrt((q_inv * q_inv).sum(-1, keepdims=True), dtype=np.float32)
    v = np.concatenate([t, np.zeros_like(t[..., :1])], axis=-1)
    q_conj = np.concatenate([-qn[..., 0:3], qn[..., 3:4]], axis=-1)
    rotated = _quat_mul(_quat_mul(qn, v), q_conj)[..., :3]
    return q_inv.astype(np.float32), (-rotated).astype(np.float32)


# --------------------------------------------------------------------------- results
@dataclass
class OracleForwardResult:
    image: np.ndarray  # (H, W, 3) f32
    depth: np.ndarray  # (H, W) f32
    pixel_valid_point_count: np.ndarray  # (H, W) i32
    pixel_accumulated_alpha: np.ndarray  # (H, W) f32
    pixel_offset_of_last_effective_point: np.ndarray  # (H, W) i32
    point_id_in_camera_list: np.ndarray  # (M,) i32
    point_uv: np.ndarray  # (M, 2)
    point_in_camera: np.ndarray  # (M, 3)
    point_uv_conic_and_rescale: np.ndarray  # (M, 4)
    point_alpha_after_activation: np.ndarray  # (M,)
    point_color: np.ndarray  # (M, 3)
    point_radii: np.ndarray  # (M,)
    num_overlap_tiles: np.ndarray  # (M,) i32
    point_in_camera_sort_key: np.ndarray  # (K,) i64 sorted
    point_offset_with_sort_key: np.ndarray  # (K,) i32 sorted
    tile_points_start: np.ndarray  # (T,) i32
    tile_points_end: np.ndarray  # (T,) i32
    q_camera_pointcloud: np.ndarray
    t_camera_pointcloud: np.ndarray


@dataclass
class OracleBackwardResult:
    grad_pointcloud: np.ndarray  # (N, 3)
    grad_pointcloud_features: np.ndarray  # (N, 56), after band clearing + factor scaling
    # BackwardValidPointHookInput fields (GPCR:806-817)
    point_id_in_camera_list: np.ndarray
    grad_point_in_camera: np.ndarray  # (M, 3)
    grad_pointfeatures_in_camera: np.ndarray  # (M, 56)
    grad_viewspace: np.ndarray  # (M, 2)
    magnitude_grad_viewspace: np.ndarray  # (M,)
    magnitude_grad_viewspace_on_image: np.ndarray  # (H, W, 2)
    num_overlap_tiles: np.ndarray  # (M,)
    num_affected_pixels: np.ndarray  # (M,)
    point_depth: np.ndarray  # (M,)
    point_uv_in_camera: np.ndarray  # (M, 2)


class OracleRasterisation:
    """CPU restatement of ``GaussianPointCloudRasterisation`` (config defaults GPCR:776-786)."""

    def __init__(self, near_plane: float = 0.8, far_plane: float = 1000.0,
                 depth_to_sort_key_scale: float = 100.0, rgb_only: bool = False,
                 grad_color_factor: float = 5.0, grad_high_order_color_factor: float = 1.0,
                 grad_s_factor: float = 0.5, grad_q_factor: float = 1.0,
                 grad_alpha_factor: float = 20.0):
        self.near_plane = near_plane
        self.far_plane = far_plane
        self.depth_to_sort_key_scale = depth_to_sort_key_scale
        self.rgb_only = rgb_only
        self.grad_color_factor = grad_color_factor
        self.grad_high_order_color_factor = grad_high_order_color_factor
        self.grad_s_factor = grad_s_factor
        self.grad_q_factor = grad_q_factor
        self.grad_alpha_factor = grad_alpha_factor

    # -- forward: GPCR:830-1023
    def forward(self, point_cloud: np.ndarray, point_cloud_features: np.ndarray,
                point_invalid_mask: np.ndarray, point_object_id: np.ndarray,
                camera_intrinsics: np.ndarray, camera_height: int, camera_width: int,
                q_pointcloud_camera: np.ndarray, t_pointcloud_camera: np.ndarray,
                ) -> OracleForwardResult:
        """NB: like the reference, normalises ``point_cloud_features[:, :4]`` IN PLACE for
        in-frustum rows (pass a float32 C-contiguous array to observe it)."""
        L = lib()
        assert camera_width % TILE_WIDTH == 0 and camera_height % TILE_HEIGHT == 0  # GPCR:1193-1194
        xyz = _f32(point_cloud)
        feat = point_cloud_features
        assert feat.dtype == np.float32 and feat.flags["C_CONTIGUOUS"]
        invalid = np.ascontiguousarray(point_invalid_mask, dtype=np.int8)
        obj = np.ascontiguousarray(point_object_id, dtype=np.int32)
        K = _f32(camera_intrinsics)
        N = xyz.shape[0]
        H, W = int(camera_height), int(camera_width)
        q_cp, t_cp = inverse_SE3_qt(q_pointcloud_camera, t_pointcloud_camera)  # GPCR:845
        q_cp = _f32(q_cp)
        t_cp = _f32(t_cp)

        # K1 (GPCR:848-860)
        mask = np.zeros(N, dtype=np.int8)
        L.gso_filter_point_in_camera(_p(xyz), _p(invalid), _p(K), _p(obj), _p(q_cp), _p(t_cp),
                                     _p(mask), ctypes.c_int64(N), ctypes.c_float(self.near_plane),
                                     ctypes.c_float(self.far_plane), ctypes.c_int(W), ctypes.c_int(H))
        # P1 (GPCR:861-870): ascending ids
        ids = np.ascontiguousarray(np.nonzero(mask)[0].astype(np.int32))
        M = ids.shape[0]
        uv = np.empty((M, 2), np.float32)
        pc = np.empty((M, 3), np.float32)
        conic = np.empty((M, 4), np.float32)
        opa = np.empty((M,), np.float32)
        color = np.zeros((M, 3), np.float32)
        radii = np.empty((M,), np.float32)
        # K2 (GPCR:887-901)
        L.gso_generate_point_attributes(_p(xyz), _p(feat), _p(K), _p(obj), _p(q_cp), _p(t_cp),
                                        _p(ids), ctypes.c_int64(M), _p(uv), _p(pc), _p(conic),
                                        _p(opa), _p(color), _p(radii))
        # K3 (GPCR:904-911)
        ntiles = np.empty((M,), np.int32)
        L.gso_generate_num_overlap_tiles(_p(ntiles), _p(uv), _p(radii), ctypes.c_int64(M),
                                         ctypes.c_int(W), ctypes.c_int(H))
        # P2 (GPCR:913-922)
        cs = np.cumsum(ntiles.astype(np.int64))
        total = int(cs[-1]) if M > 0 else 0
        acc = np.ascontiguousarray(np.concatenate([np.zeros(1, np.int64), cs[:-1]]) if M > 0
                                   else np.zeros(0, np.int64))
        keys = np.empty((total,), np.int64)
        vals = np.empty((total,), np.int32)
        # K4 (GPCR:934-945)
        if total > 0:
            L.gso_generate_point_sort_key(_p(uv), _p(pc), _p(radii), _p(acc), ctypes.c_int64(M),
                                          _p(vals), _p(keys), ctypes.c_int(W), ctypes.c_int(H),
                                          ctypes.c_float(self.depth_to_sort_key_scale))
            # P3 (GPCR:947-950): stable
            L.gso_stable_sort_pairs(_p(keys), _p(vals), ctypes.c_int64(total))
        T = (W // TILE_WIDTH) * (H // TILE_HEIGHT)
        tstart = np.zeros((T,), np.int32)
        tend = np.zeros((T,), np.int32)
        # K5 (GPCR:959-964)
        if total > 0:
            L.gso_find_tile_start_and_end(_p(keys), ctypes.c_int64(total), _p(tstart), _p(tend))
        # K6 (GPCR:967-997). K == 0: the reference returns torch.empty garbage; we define zeros.
        image = np.zeros((H, W, 3), np.float32)
        depth = np.zeros((H, W), np.float32)
        acc_alpha = np.zeros((H, W), np.float32)
        last = np.zeros((H, W), np.int32)
        count = np.zeros((H, W), np.int32)
        if total > 0:
            L.gso_rasterisation_forward(ctypes.c_int(H), ctypes.c_int(W), _p(tstart), _p(tend),
                                        _p(vals), _p(uv), _p(pc), _p(conic), _p(opa), _p(color),
                                        _p(image), _p(depth), _p(acc_alpha), _p(last), _p(count),
                                        ctypes.c_int(1 if self.rgb_only else 0))
        return OracleForwardResult(
            image=image, depth=depth, pixel_valid_point_count=count,
            pixel_accumulated_alpha=acc_alpha, pixel_offset_of_last_effective_point=last,
            point_id_in_camera_list=ids, point_uv=uv, point_in_camera=pc,
            point_uv_conic_and_rescale=conic, point_alpha_after_activation=opa, point_color=color,
            point_radii=radii, num_overlap_tiles=ntiles, point_in_camera_sort_key=keys,
            point_offset_with_sort_key=vals, tile_points_start=tstart, tile_points_end=tend,
            q_camera_pointcloud=q_cp, t_camera_pointcloud=t_cp)

    # -- backward: GPCR:1025-1163
    def backward(self, fwd: OracleForwardResult, grad_rasterized_image: np.ndarray,
                 point_cloud: np.ndarray, point_cloud_features: np.ndarray,
                 point_object_id: np.ndarray, camera_intrinsics: np.ndarray,
                 t_pointcloud_camera: np.ndarray, color_max_sh_band: int = 3,
                 ) -> OracleBackwardResult:
        L = lib()
        xyz = _f32(point_cloud)
        feat = _f32(point_cloud_features)
        obj = np.ascontiguousarray(point_object_id, dtype=np.int32)
        K = _f32(camera_intrinsics)
        t_pc = _f32(t_pointcloud_camera)
        g_img = _f32(grad_rasterized_image)
        H, W = g_img.shape[0], g_img.shape[1]
        N = xyz.shape[0]
        ids = fwd.point_id_in_camera_list
        M = ids.shape[0]
        grad_uv = np.zeros((N, 2), np.float64)
        cov_buf = np.zeros((M, 3), np.float64)
        color_buf = np.zeros((M, 3), np.float64)
        grad_logit = np.zeros((N,), np.float64)
        magnitude = np.zeros((N,), np.float64)
        n_pixels = np.zeros((M,), np.int32)
        mag_img = np.zeros((H, W, 2), np.float32)
        if fwd.point_offset_with_sort_key.shape[0] > 0:
            L.gso_rasterisation_backward_pixels(
                ctypes.c_int(H), ctypes.c_int(W), _p(fwd.tile_points_start), _p(fwd.tile_points_end),
                _p(fwd.point_offset_with_sort_key), _p(ids), _p(g_img),
                _p(fwd.pixel_accumulated_alpha), _p(fwd.pixel_offset_of_last_effective_point),
                _p(fwd.point_uv), _p(fwd.point_uv_conic_and_rescale),
                _p(fwd.point_alpha_after_activation), _p(fwd.point_color), ctypes.c_int64(N),
                ctypes.c_int64(M), _p(grad_uv), _p(cov_buf), _p(color_buf), _p(grad_logit),
                _p(magnitude), _p(n_pixels), _p(mag_img))
        grad_uv32 = grad_uv.astype(np.float32)
        cov32 = cov_buf.astype(np.float32)
        col32 = color_buf.astype(np.float32)
        grad_xyz = np.zeros((N, 3), np.float32)
        grad_feat = np.zeros((N, 56), np.float32)
        grad_feat[:, 7] = grad_logit.astype(np.float32)
        L.gso_rasterisation_backward_points(
            _p(K), _p(xyz), _p(feat), _p(obj), _p(fwd.q_camera_pointcloud),
            _p(fwd.t_camera_pointcloud), _p(t_pc), _p(ids), ctypes.c_int64(M),
            _p(fwd.point_in_camera), _p(grad_uv32), _p(cov32), _p(col32), _p(grad_xyz),
            _p(grad_feat))
        band = int(color_max_sh_band)
        L.gso_grad_postprocess(_p(grad_feat), ctypes.c_int64(N), ctypes.c_int(band),
                               ctypes.c_float(self.grad_q_factor), ctypes.c_float(self.grad_s_factor),
                               ctypes.c_float(self.grad_alpha_factor),
                               ctypes.c_float(self.grad_color_factor),
                               ctypes.c_float(self.grad_high_order_color_factor))
        return OracleBackwardResult(
            grad_pointcloud=grad_xyz, grad_pointcloud_features=grad_feat,
            point_id_in_camera_list=ids, grad_point_in_camera=grad_xyz[ids],
            grad_pointfeatures_in_camera=grad_feat[ids], grad_viewspace=grad_uv32[ids],
            magnitude_grad_viewspace=magnitude.astype(np.float32)[ids],
            magnitude_grad_viewspace_on_image=mag_img, num_overlap_tiles=fwd.num_overlap_tiles,
            num_affected_pixels=n_pixels, point_depth=fwd.point_in_camera[:, 2].copy(),
            point_uv_in_camera=fwd.point_uv)


def num_threads() -> int:
    return int(lib().gso_num_threads())


def set_num_threads(n: Optional[int]) -> None:
    lib().gso_set_num_threads(ctypes.c_int(int(n or os.cpu_count() or 1)))
