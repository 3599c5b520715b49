"""Host-side SE(3)/quaternion helpers that sit on the operator's call path.

Same names / argument meaning as the reference's ``taichi_3d_gaussian_splatting/utils.py``:
``inverse_SE3_qt_torch`` (utils.py:426-432) is called by the operator forward
(GaussianPointCloudRasterisation.py:845); the others are what callers and tests use to build
``q_pointcloud_camera`` / ``t_pointcloud_camera`` (utils.py:386-423, 435-492, 596-632).
Quaternions are (x, y, z, w).
"""
from typing import Tuple

import torch


def quaternion_conjugate_torch(q: torch.Tensor) -> torch.Tensor:
    return torch.cat([-q[..., 0:3], q[..., 3:4]], dim=-1)


def quaternion_multiply_torch(q0: torch.Tensor, q1: torch.Tensor) -> torch.Tensor:
    x0, y0, z0, w0 = q0.unbind(-1)
    x1, y1, z1, w1 = q1.unbind(-1)
    return torch.stack([
        w0 * x1 + x0 * w1 + y0 * z1 - z0 * y1,
        w0 * y1 - x0 * z1 + y0 * w1 + z0 * x1,
        w0 * z1 + x0 * y1 - y0 * x1 + z0 * w1,
        w0 * w1 - x0 * x1 - y0 * y1 - z0 * z1,
    ], dim=-1)


def quaternion_rotate_torch(q: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """Rotate v (.., 3) by q (.., 4); q is normalised first (utils.py:414-423)."""
    q = q / torch.norm(q, dim=-1, keepdim=True)
    v4 = torch.cat([v, torch.zeros_like(v[..., :1])], dim=-1)
    return quaternion_multiply_torch(quaternion_multiply_torch(q, v4),
                                     quaternion_conjugate_torch(q))[..., :3]


def inverse_SE3_qt_torch(q: torch.Tensor, t: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """(q, t) of T  ->  (q, t) of T^-1:  q' = conj(q), t' = -R(q')·t  (utils.py:426-432)."""
    q_inv = quaternion_conjugate_torch(q)
    t_inv = -quaternion_rotate_torch(q_inv, t)
    return q_inv, t_inv


def quaternion_to_rotation_matrix_torch(q: torch.Tensor) -> torch.Tensor:
    """(.., 4) xyzw -> (.., 3, 3).  Like the reference (utils.py:596-632) the polynomial is
    evaluated on the components as given (callers pass unit quaternions)."""
    x, y, z, w = q.unbind(-1)
    R = torch.empty((*q.shape[:-1], 3, 3), dtype=q.dtype, device=q.device)
    R[..., 0, 0] = 1 - 2 * (y * y + z * z)
    R[..., 0, 1] = 2 * (x * y - z * w)
    R[..., 0, 2] = 2 * (x * z + y * w)
    R[..., 1, 0] = 2 * (x * y + z * w)
    R[..., 1, 1] = 1 - 2 * (x * x + z * z)
    R[..., 1, 2] = 2 * (y * z - x * w)
    R[..., 2, 0] = 2 * (x * z - y * w)
    R[..., 2, 1] = 2 * (y * z + x * w)
    R[..., 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def rotation_matrix_to_quaternion_torch(R: torch.Tensor) -> torch.Tensor:
    """(B, 3, 3) -> (B, 4) xyzw, branch on the largest diagonal term (utils.py:435-483)."""
    B = R.shape[0]
    m00, m11, m22 = R[:, 0, 0], R[:, 1, 1], R[:, 2, 2]
    trace = m00 + m11 + m22
    q = torch.zeros(B, 4, dtype=R.dtype, device=R.device)
    c0 = trace > 0
    c1 = (m00 > m11) & (m00 > m22) & ~c0
    c2 = (m11 > m22) & ~c0 & ~c1
    c3 = ~c0 & ~c1 & ~c2
    eps = torch.finfo(R.dtype).tiny
    s0 = 0.5 / torch.sqrt(torch.clamp(1 + trace, min=eps))
    cand0 = torch.stack([(R[:, 2, 1] - R[:, 1, 2]) * s0, (R[:, 0, 2] - R[:, 2, 0]) * s0,
                         (R[:, 1, 0] - R[:, 0, 1]) * s0, 0.25 / s0], dim=-1)
    s1 = 2.0 * torch.sqrt(torch.clamp(1 + m00 - m11 - m22, min=eps))
    cand1 = torch.stack([0.25 * s1, (R[:, 0, 1] + R[:, 1, 0]) / s1, (R[:, 0, 2] + R[:, 2, 0]) / s1,
                         (R[:, 2, 1] - R[:, 1, 2]) / s1], dim=-1)
    s2 = 2.0 * torch.sqrt(torch.clamp(1 + m11 - m00 - m22, min=eps))
    cand2 = torch.stack([(R[:, 0, 1] + R[:, 1, 0]) / s2, 0.25 * s2, (R[:, 1, 2] + R[:, 2, 1]) / s2,
                         (R[:, 0, 2] - R[:, 2, 0]) / s2], dim=-1)
    s3 = 2.0 * torch.sqrt(torch.clamp(1 + m22 - m00 - m11, min=eps))
    cand3 = torch.stack([(R[:, 0, 2] + R[:, 2, 0]) / s3, (R[:, 1, 2] + R[:, 2, 1]) / s3, 0.25 * s3,
                         (R[:, 1, 0] - R[:, 0, 1]) / s3], dim=-1)
    for c, cand in ((c0, cand0), (c1, cand1), (c2, cand2), (c3, cand3)):
        q = torch.where(c[:, None], cand, q)
    return q


def SE3_to_quaternion_and_translation_torch(transform: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """(B, 4, 4) -> q (B, 4) xyzw, t (B, 3)  (utils.py:486-492)."""
    return rotation_matrix_to_quaternion_torch(transform[..., :3, :3]), transform[..., :3, 3]
