"""Golden vectors for the dataset row: run the REFERENCE's ImagePoseDataset (ImagePoseDataset.py:16-103, imported from
/root/reference with taichi stubbed) on the generated fixture dataset and store what it returns.

    python tests/golden/make_dataset_golden.py        # build container only; writes dataset_vectors.json
"""
import json
import os
import sys
import tempfile
import types
from unittest import mock

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from dataset_fixture import write_dataset  # noqa: E402


def main():
    ti = mock.MagicMock()
    ti.func = lambda f: f
    ti.kernel = lambda f: f
    ti.dataclass = lambda c: c
    sys.modules["taichi"] = ti
    sys.modules["taichi.math"] = ti.math
    sys.modules.setdefault("dataclass_wizard", types.SimpleNamespace(YAMLWizard=object))
    sys.path.insert(0, "/root/reference")
    from taichi_3d_gaussian_splatting.ImagePoseDataset import ImagePoseDataset
    out = []
    with tempfile.TemporaryDirectory() as d:
        ds = ImagePoseDataset(write_dataset(d))
        for i in range(len(ds)):
            image, q, t, info = ds[i]
            h, w = image.shape[1:]
            probes = [(0, 0), (h // 2, w // 3), (h - 1, w - 1), (h // 3, w - 1)]
            out.append(dict(shape=list(image.shape), mean=float(image.double().mean()),
                            probes=[[int(y), int(x)] + image[:, y, x].tolist() for y, x in probes],
                            q=q.tolist(), t=t.tolist(), K=info.camera_intrinsics.tolist(),
                            camera_height=int(info.camera_height), camera_width=int(info.camera_width),
                            camera_id=int(info.camera_id)))
    with open(os.path.join(HERE, "dataset_vectors.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", len(out), "items")


if __name__ == "__main__":
    main()
