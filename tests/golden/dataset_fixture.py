"""Deterministic little posed-image dataset used by make_dataset_golden.py (reference side) and by
tests/test_dataset_cpu.py (our side): three PNG frames -- one odd-sized, one whose recorded size differs from the
file's, one above the 1600-pixel limit -- plus the JSON records of docs/RawDataFormat.md."""
import json
import os

import numpy as np


def _frame(h, w, seed):
    rng = np.random.default_rng(seed)
    yy, xx = np.meshgrid(np.linspace(0, 1, h), np.linspace(0, 1, w), indexing="ij")
    img = np.stack([0.5 + 0.5 * np.sin(6 * xx + seed), yy, 0.5 + 0.5 * np.cos(5 * xx * yy + seed)], axis=-1)
    img += 0.02 * rng.standard_normal(img.shape)
    return (np.clip(img, 0, 1) * 255).astype(np.uint8)


def _pose(k):
    c, s = np.cos(0.2 * k), np.sin(0.2 * k)
    T = np.eye(4)
    T[:3, :3] = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]]) @ np.array([[1, 0, 0], [0, np.cos(0.1 * k), -np.sin(0.1 * k)],
                                                                          [0, np.sin(0.1 * k), np.cos(0.1 * k)]])
    T[:3, 3] = [0.3 * k, -0.1 * k, 1.0 + k]
    return T.tolist()


def write_dataset(folder):
    import PIL.Image
    os.makedirs(folder, exist_ok=True)
    spec = [  # file size (h, w), recorded size (h, w)
        ((70, 50), (70, 50)),
        ((100, 131), (200, 262)),      # file is half the recorded resolution
        ((1100, 1700), (1100, 1700)),  # above MAX_RESOLUTION_TRAIN: autoscaled
    ]
    records = []
    for k, ((h, w), (rh, rw)) in enumerate(spec):
        path = os.path.join(folder, f"frame_{k}.png")
        PIL.Image.fromarray(_frame(h, w, k)).save(path)
        records.append(dict(image_path=path, T_pointcloud_camera=_pose(k),
                            camera_intrinsics=[[0.9 * rw, 0.0, rw / 2 + 1.5], [0.0, 0.8 * rh, rh / 2 - 2.0], [0.0, 0.0, 1.0]],
                            camera_height=rh, camera_width=rw, camera_id=k))
    json_path = os.path.join(folder, "dataset.json")
    with open(json_path, "w") as f:
        json.dump(records, f)
    return json_path
