"""Camera dataclasses -- the argument types of the rasteriser operator.

Mirrors the reference's ``taichi_3d_gaussian_splatting/Camera.py:7-21`` (``CameraInfo`` is an
argument of ``GaussianPointCloudRasterisationInput``; ``CameraView`` is imported beside it at
GaussianPointCloudRasterisation.py:4).
"""
from dataclasses import dataclass
from typing import Optional

import torch


@dataclass
class CameraInfo:
    camera_intrinsics: torch.Tensor  # 3x3 f32 pinhole matrix (device tensor in the reference)
    camera_height: int
    camera_width: int
    camera_id: int


@dataclass
class CameraView:
    camera_view_id: int
    T_pointcloud_camera: torch.Tensor  # 4x4 SE(3), camera frame -> pointcloud frame
    camera_id: int
    image_id: int
    timestamp: Optional[int] = None
