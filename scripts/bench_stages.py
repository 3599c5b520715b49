"""Per-stage device times at the bench workload for the library selected by GSB200_LIB_PATH."""
import os, sys, json, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from taichi_3d_gaussian_splatting_b200 import GaussianPointCloudRasterisation as GPCR, profiling
from taichi_3d_gaussian_splatting_b200.synthetic import CONFIGS, make_scene
name = sys.argv[1] if len(sys.argv) > 1 else "C3"
cfg = CONFIGS[name]
scene = make_scene(**cfg).to("cuda")
op = GPCR(GPCR.GaussianPointCloudRasterisationConfig(), keep_all_tile_pairs=bool(int(os.environ.get('GSB_KEEP_ALL', '0'))),
          backward_impl=os.environ.get('GSB_BENCH_BACKWARD_IMPL') or None)
inp = GPCR.GaussianPointCloudRasterisationInput(
    point_cloud=scene.point_cloud, point_cloud_features=scene.point_cloud_features,
    point_object_id=scene.point_object_id, point_invalid_mask=scene.point_invalid_mask,
    camera_info=scene.camera_info, q_pointcloud_camera=scene.q_pointcloud_camera,
    t_pointcloud_camera=scene.t_pointcloud_camera, color_max_sh_band=3)
g = torch.randn((cfg["height"], cfg["width"], 3), generator=torch.Generator().manual_seed(1)).cuda()
t = profiling.stage_times(op, inp, g, iters=10)
print(os.environ.get("GSB200_LIB_PATH", "default"), name, "K", op.last_frame.num_keys, {k: round(v * 1e3, 1) for k, v in t.items()}, "sum_us", round(sum(t.values()) * 1e3, 1))
