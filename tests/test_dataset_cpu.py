"""ImagePoseDataset against golden vectors produced by the REFERENCE's own class
(tests/golden/make_dataset_golden.py imports /root/reference's ImagePoseDataset with taichi stubbed and stores its
outputs for the fixture dataset of tests/golden/dataset_fixture.py)."""
import json
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from dataset_fixture import write_dataset  # noqa: E402

from taichi_3d_gaussian_splatting_b200.image_pose_dataset import MAX_RESOLUTION_TRAIN, ImagePoseDataset  # noqa: E402


@pytest.fixture(scope="module")
def dataset(tmp_path_factory):
    return ImagePoseDataset(write_dataset(str(tmp_path_factory.mktemp("posed_images"))))


def test_items_match_the_reference_class(dataset):
    with open(os.path.join(HERE, "golden", "dataset_vectors.json")) as f:
        golden = json.load(f)
    assert len(dataset) == len(golden) == 3
    for i, ref in enumerate(golden):
        image, q, t, info = dataset[i]
        assert list(image.shape) == ref["shape"] and image.dtype == torch.float32 and image.is_contiguous()
        assert (info.camera_height, info.camera_width, info.camera_id) == (ref["camera_height"], ref["camera_width"], ref["camera_id"])
        assert info.camera_height % 16 == 0 and info.camera_width % 16 == 0  # GPCR:1193-1194
        assert max(info.camera_height, info.camera_width) <= MAX_RESOLUTION_TRAIN
        # item 2 goes through the antialiased resize: allow for a different torchvision build on the test machine
        tol = 1e-6 if i < 2 else 2e-3
        assert abs(float(image.double().mean()) - ref["mean"]) <= tol
        for y, x, r, g, b in ref["probes"]:
            assert np.allclose(image[:, y, x].numpy(), [r, g, b], atol=tol)
        assert np.allclose(info.camera_intrinsics.numpy(), np.array(ref["K"]), rtol=1e-6, atol=1e-5)
        assert np.allclose(q.numpy(), np.array(ref["q"]), atol=1e-6) and q.shape == (1, 4)
        assert np.allclose(t.numpy(), np.array(ref["t"]), atol=1e-6) and t.shape == (1, 3)


def test_items_feed_the_operator_input(dataset):
    from taichi_3d_gaussian_splatting_b200 import GaussianPointCloudRasterisation as GPCR
    image, q, t, info = dataset[1]
    n = 8
    inp = GPCR.GaussianPointCloudRasterisationInput(
        point_cloud=torch.zeros(n, 3), point_cloud_features=torch.zeros(n, 56), point_object_id=torch.zeros(n, dtype=torch.int32),
        point_invalid_mask=torch.zeros(n, dtype=torch.int8), camera_info=info, q_pointcloud_camera=q, t_pointcloud_camera=t)
    assert inp.camera_info.camera_intrinsics.shape == (3, 3) and image.shape[1:] == (info.camera_height, info.camera_width)


def test_missing_column_is_reported(tmp_path):
    p = tmp_path / "bad.json"
    p.write_text(json.dumps([{"image_path": "x.png", "camera_id": 0}]))
    with pytest.raises(AssertionError, match="lacks"):
        ImagePoseDataset(str(p))
