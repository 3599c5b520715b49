"""fwd+bwd step time of the bench workload for operator variants (CUDA events, median of regions).

    python scripts/bench_step.py C3            # the library selected by GSB200_LIB_PATH (default: the in-tree build)
"""
import os, sys, statistics, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from taichi_3d_gaussian_splatting_b200 import GaussianPointCloudRasterisation as GPCR
from taichi_3d_gaussian_splatting_b200.synthetic import CONFIGS, make_scene

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
cfg = CONFIGS[name]
scene = make_scene(**cfg).to("cuda")
scene.point_cloud.requires_grad_(True)
scene.point_cloud_features.requires_grad_(True)
g = torch.randn((cfg["height"], cfg["width"], 3), generator=torch.Generator().manual_seed(1234)).cuda()


def make_input():
    return GPCR.GaussianPointCloudRasterisationInput(
        point_cloud=scene.point_cloud, point_cloud_features=scene.point_cloud_features,
        point_object_id=scene.point_object_id, point_invalid_mask=scene.point_invalid_mask,
        camera_info=scene.camera_info, q_pointcloud_camera=scene.q_pointcloud_camera,
        t_pointcloud_camera=scene.t_pointcloud_camera, color_max_sh_band=3)


def timed(fn, k):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k


for label, kw in (("default", {}), ("butterfly_backward", {"backward_impl": "butterfly"})):
    op = GPCR(GPCR.GaussianPointCloudRasterisationConfig(), **kw)
    inp = make_input()

    def step():
        scene.point_cloud.grad = None
        scene.point_cloud_features.grad = None
        image, _, _ = op(inp)
        image.backward(g)

    def fwd():
        with torch.no_grad():
            op(inp)
    for _ in range(5):
        step()
    steps = [timed(step, 20) for _ in range(7)]
    fwds = [timed(fwd, 20) for _ in range(5)]
    print(name, label, "fwd+bwd ms", round(statistics.median(steps), 4), "min", round(min(steps), 4), "| fwd ms", round(statistics.median(fwds), 4))
