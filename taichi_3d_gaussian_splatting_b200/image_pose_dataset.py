"""Posed-image dataset feeding the operator (the data format on the caller side of the path).

Same record format and item contract as the reference's ``ImagePoseDataset``
(``taichi_3d_gaussian_splatting/ImagePoseDataset.py:16-103``; format described in ``docs/RawDataFormat.md``):
a JSON list of records with ``image_path``, ``T_pointcloud_camera`` (4x4, camera -> point cloud),
``camera_intrinsics`` (3x3), ``camera_height``, ``camera_width``, ``camera_id``.  ``dataset[i]`` returns
``(image (3,H,W) float32 in [0,1], q_pointcloud_camera (1,4) xyzw, t_pointcloud_camera (1,3), CameraInfo)`` where

* the intrinsics are rescaled from the recorded size to the size of the image actually on disk (:78-83),
* H and W are cropped to multiples of the 16-pixel tile, which the rasteriser requires (:84-88; GPCR:1193-1194),
* frames with a side above ``MAX_RESOLUTION_TRAIN`` are resized (shorter side 1024, longer side capped at 1600,
  antialiased), cropped again and their fx, fy, cx, cy scaled (:41-66).
Records are parsed with ``json`` (no pandas needed); relative image paths are resolved against the JSON file.
Pinned against the reference class itself: ``tests/golden/make_dataset_golden.py`` imports it (Taichi stubbed) and
stores its outputs for a small generated dataset; ``tests/test_dataset_cpu.py`` compares.
"""
import json
import os
from typing import List, Tuple

import numpy as np
import torch
import torch.utils.data

from .Camera import CameraInfo
from .GaussianPointCloudRasterisation import TILE_HEIGHT, TILE_WIDTH
from .utils import SE3_to_quaternion_and_translation_torch

MAX_RESOLUTION_TRAIN = 1600
_REQUIRED = ("image_path", "T_pointcloud_camera", "camera_intrinsics", "camera_height", "camera_width", "camera_id")


def _crop_to_tiles(image: torch.Tensor) -> torch.Tensor:
    h = image.shape[1] - image.shape[1] % TILE_HEIGHT
    w = image.shape[2] - image.shape[2] % TILE_WIDTH
    return image[:3, :h, :w].contiguous()


def _load_image(path: str) -> torch.Tensor:
    """(C,H,W) float32 in [0,1] -- what ``torchvision.transforms.functional.to_tensor`` yields for 8-bit images."""
    import PIL.Image
    with PIL.Image.open(path) as im:
        arr = np.array(im.convert("RGB") if im.mode not in ("RGB", "RGBA", "L") else im)  # writable copy
    if arr.ndim == 2:
        arr = arr[:, :, None]
    return torch.from_numpy(np.ascontiguousarray(arr)).permute(2, 0, 1).to(torch.float32).div(255.0)


class ImagePoseDataset(torch.utils.data.Dataset):
    def __init__(self, dataset_json_path: str):
        super().__init__()
        with open(dataset_json_path) as f:
            self.records: List[dict] = json.load(f)
        self.root = os.path.dirname(os.path.abspath(dataset_json_path))
        for i, rec in enumerate(self.records):
            missing = [k for k in _REQUIRED if k not in rec]
            assert not missing, f"record {i} of {dataset_json_path} lacks {missing}"

    def __len__(self) -> int:
        return len(self.records)

    @staticmethod
    def _autoscale_image_and_camera_info(image: torch.Tensor, camera_info: CameraInfo) -> Tuple[torch.Tensor, CameraInfo]:
        if max(camera_info.camera_height, camera_info.camera_width) <= MAX_RESOLUTION_TRAIN:
            return image, camera_info
        import torchvision.transforms.functional as TF
        resized = TF.resize(image, size=1024, max_size=MAX_RESOLUTION_TRAIN, antialias=True)
        sy = resized.shape[1] / camera_info.camera_height
        sx = resized.shape[2] / camera_info.camera_width
        resized = _crop_to_tiles(resized)
        K = camera_info.camera_intrinsics.clone()
        K[0, 0] *= sx
        K[0, 2] *= sx
        K[1, 1] *= sy
        K[1, 2] *= sy
        return resized, CameraInfo(camera_intrinsics=K, camera_height=resized.shape[1], camera_width=resized.shape[2],
                                   camera_id=camera_info.camera_id)

    def __getitem__(self, idx: int):
        rec = self.records[idx]
        path = rec["image_path"]
        if not os.path.isabs(path) and not os.path.exists(path):
            path = os.path.join(self.root, path)
        image = _load_image(path)
        T = torch.tensor(rec["T_pointcloud_camera"], dtype=torch.float32).reshape(4, 4)
        q, t = SE3_to_quaternion_and_translation_torch(T.unsqueeze(0))
        K = torch.tensor(rec["camera_intrinsics"], dtype=torch.float32).reshape(3, 3)
        # the image on disk decides the size, not the recorded (COLMAP) one
        K[0, :] = K[0, :] * image.shape[2] / rec["camera_width"]
        K[1, :] = K[1, :] * image.shape[1] / rec["camera_height"]
        image = _crop_to_tiles(image)
        info = CameraInfo(camera_intrinsics=K, camera_height=image.shape[1], camera_width=image.shape[2],
                          camera_id=rec["camera_id"])
        image, info = self._autoscale_image_and_camera_info(image, info)
        return image, q, t, info
