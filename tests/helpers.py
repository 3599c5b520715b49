"""Shared test helpers: run the oracle on a SyntheticScene, compare arrays."""
import numpy as np
import torch

from oracle import OracleRasterisation


def oracle_forward(scene, **cfg):
    o = OracleRasterisation(**cfg)
    feats = scene.point_cloud_features.detach().cpu().numpy().astype(np.float32).copy()
    ci = scene.camera_info
    fwd = o.forward(scene.point_cloud.detach().cpu().numpy(), feats,
                    scene.point_invalid_mask.cpu().numpy(), scene.point_object_id.cpu().numpy(),
                    ci.camera_intrinsics.cpu().numpy(), ci.camera_height, ci.camera_width,
                    scene.q_pointcloud_camera.cpu().numpy(), scene.t_pointcloud_camera.cpu().numpy())
    return o, fwd, feats


def oracle_backward(o, fwd, scene, feats_normalised, grad_image, band):
    ci = scene.camera_info
    return o.backward(fwd, np.asarray(grad_image, dtype=np.float32),
                      scene.point_cloud.detach().cpu().numpy(), feats_normalised,
                      scene.point_object_id.cpu().numpy(), ci.camera_intrinsics.cpu().numpy(),
                      scene.t_pointcloud_camera.cpu().numpy(), band)


def rel_err(a, b, floor_frac=1e-6):
    """max |a-b| / max(|b|, floor) with floor = floor_frac * max|b| (SURVEY §8(d) C1 definition)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = np.abs(b).max() if b.size else 0.0
    if scale == 0.0:
        return float(np.abs(a).max()) if a.size else 0.0
    return float((np.abs(a - b) / np.maximum(np.abs(b), floor_frac * scale)).max())


def rel_err_global(a, b):
    """max |a-b| / max|b| -- tolerance-friendly when tiny entries are dominated by rounding noise."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = np.abs(b).max() if b.size else 0.0
    if scale == 0.0:
        return float(np.abs(a).max()) if a.size else 0.0
    return float(np.abs(a - b).max() / scale)


def t2n(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


def grad_close(a, b, rtol=1e-3, floor_frac=1e-5):
    """Gradient parity: |a-b| <= rtol*|b| + floor_frac*max|b|.

    north_star asks for 1e-3 relative; the absolute floor (1e-5 of the largest gradient of the group)
    covers entries that are sums of cancelling f32 terms: the reference accumulates them with f32
    atomics in a run-to-run varying order, so their low bits are not defined by the reference either.
    Returns (ok, worst_excess_ratio, n_violations)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = np.abs(b).max() if b.size else 0.0
    tol = rtol * np.abs(b) + floor_frac * scale
    d = np.abs(a - b)
    if scale == 0.0:
        return bool((d == 0).all()), float(d.max()) if d.size else 0.0, int((d > 0).sum())
    viol = d > tol
    return bool(not viol.any()), float((d / np.maximum(tol, 1e-300)).max()), int(viol.sum())
