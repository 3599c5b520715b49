"""Golden files for the scene row: run the REFERENCE's GaussianPointCloudScene (GaussianPointCloudScene.py:12-239,
imported from /root/reference; its unused-here imports dataclass_wizard / plyfile are stubbed) and store

* ``sparse_points.parquet``   -- a 40-point x,y,z,r,g,b cloud (what a COLMAP import produces),
* ``reference_scene.parquet`` -- the scene the reference builds from it (spare capacity, kNN scales, rgb -> SH DC)
  written by the reference's own ``to_parquet``,
* ``scene_vectors.json``      -- the reference's in-memory tensors for that scene, and the result of the reference
  reading a parquet written by OUR ``to_parquet`` (cross-compatibility, checked here because /root/reference does not
  travel).

    python tests/golden/make_scene_golden.py        # build container only
"""
import json
import os
import sys
import tempfile
import types

import numpy as np
import pandas as pd
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def main():
    sys.modules.setdefault("dataclass_wizard", types.SimpleNamespace(YAMLWizard=object))
    sys.modules.setdefault("plyfile", types.SimpleNamespace(PlyData=object, PlyElement=object))
    sys.path.insert(0, "/root/reference")
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "ref_scene", "/root/reference/taichi_3d_gaussian_splatting/GaussianPointCloudScene.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    RefScene = ref.GaussianPointCloudScene

    rng = np.random.default_rng(5)
    sparse = pd.DataFrame(dict(x=rng.normal(0, 1.0, 40), y=rng.normal(0, 0.6, 40), z=rng.normal(3, 0.8, 40),
                               r=rng.integers(0, 256, 40).astype(np.float64), g=rng.integers(0, 256, 40).astype(np.float64),
                               b=rng.integers(0, 256, 40).astype(np.float64)))
    sparse.loc[0, ["r", "g", "b"]] = [255.0, 0.0, 128.0]  # exercises the clamp to 0.99 and logit(0) = -inf
    sparse_path = os.path.join(HERE, "sparse_points.parquet")
    sparse.to_parquet(sparse_path)

    cfg = RefScene.PointCloudSceneConfig(max_num_points_ratio=2.5, initial_alpha=-1.5, initial_covariance_ratio=0.7,
                                         max_initial_covariance=0.4)
    torch.manual_seed(0)
    scene = RefScene.from_parquet(sparse_path, config=cfg)
    scene.to_parquet(os.path.join(HERE, "reference_scene.parquet"))
    out = dict(config=dict(max_num_points_ratio=2.5, initial_alpha=-1.5, initial_covariance_ratio=0.7, max_initial_covariance=0.4),
               point_cloud=scene.point_cloud.detach().tolist(), point_cloud_features=scene.point_cloud_features.detach().tolist(),
               point_invalid_mask=scene.point_invalid_mask.tolist(), point_object_id=scene.point_object_id.tolist())

    # cross read: a parquet written by OUR class must load in the reference with identical tensors
    sys.path.insert(0, ROOT)
    from taichi_3d_gaussian_splatting_b200 import GaussianPointCloudScene as OurScene
    ours = OurScene.from_parquet(os.path.join(HERE, "reference_scene.parquet"))
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "ours.parquet")
        ours.to_parquet(p)
        back = RefScene.from_parquet(p, config=RefScene.PointCloudSceneConfig())
        valid = scene.point_invalid_mask == 0
        assert torch.equal(back.point_cloud.detach(), scene.point_cloud.detach()[valid])
        assert torch.equal(back.point_cloud_features.detach(), scene.point_cloud_features.detach()[valid])
    out["reference_reads_our_parquet"] = True
    with open(os.path.join(HERE, "scene_vectors.json"), "w") as f:
        json.dump(out, f)
    print("wrote scene golden:", scene.point_cloud.shape, "capacity,", int((scene.point_invalid_mask == 0).sum()), "valid")


if __name__ == "__main__":
    main()
