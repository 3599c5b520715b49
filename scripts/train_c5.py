"""BASELINE config 5 harness: 1k trainer iterations with the tat_truck_every_8_test.yaml shapes and hyper-parameters
(config/tat_truck_every_8_test.yaml: 976x544 frames, down-sampling 4 -> 2 -> 1 every 250 iterations, near 0.4 / far 2000 /
depth-key scale 10, feature LR 5e-3, position LR 1e-5 (the YAML's misspelt key leaves the default), densification warm-up
1000, capacity = 10 x (sparse points + 10 000 sphere points)), on SYNTHETIC targets: renders of a hidden scene.
Prints one JSON line: iterations/s, PSNR before/after, point counts."""
import json, math, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from taichi_3d_gaussian_splatting_b200 import CameraInfo, GaussianPointCloudRasterisation as GPCR, GaussianPointCloudScene
from taichi_3d_gaussian_splatting_b200.synthetic import make_scene
from taichi_3d_gaussian_splatting_b200.trainer import GaussianPointCloudTrainer

H, W = 544, 976
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
dev = "cuda"
torch.manual_seed(0)
hidden = make_scene(140_000, H, W, 0.03, seed=11, sh_degree=3)  # Truck 7k-iteration scale (README.md:63: ~2.3e5 points)
rast_cfg = GPCR.GaussianPointCloudRasterisationConfig(near_plane=0.4, far_plane=2000.0, depth_to_sort_key_scale=10.0)
renderer = GPCR(rast_cfg)
K = hidden.camera_info.camera_intrinsics.to(dev)
hp, hf = hidden.point_cloud.to(dev), hidden.point_cloud_features.to(dev)
hm, ho = hidden.point_invalid_mask.to(dev), hidden.point_object_id.to(dev)
views = []
for i in range(31):  # "every 8th" of a 251-frame capture
    yaw = math.radians(-12 + 24 * i / 30)
    q = torch.tensor([[0.0, math.sin(yaw / 2), 0.0, math.cos(yaw / 2)]], device=dev)
    t = torch.tensor([[0.6 * math.sin(yaw), 0.0, 0.0]], device=dev)
    cam = CameraInfo(K, H, W, 0)
    with torch.no_grad():
        img, _, _ = renderer(GPCR.GaussianPointCloudRasterisationInput(
            point_cloud=hp, point_cloud_features=hf, point_object_id=ho, point_invalid_mask=hm, camera_info=cam,
            q_pointcloud_camera=q, t_pointcloud_camera=t, color_max_sh_band=3))
    views.append((img.clamp(0, 1).permute(2, 0, 1).contiguous(), q, t, cam))

# sparse initial cloud (every 10th hidden point, jittered, with its colour) + background sphere, x10 capacity
import pandas as pd, tempfile
g = np.random.default_rng(5)
sparse = hidden.point_cloud[::10].numpy() + g.normal(scale=0.02, size=(hidden.point_cloud[::10].shape[0], 3))
rgb = (torch.sigmoid(hidden.point_cloud_features[::10][:, [8, 24, 40]] * 0.28209479177387814).numpy() * 255).round()
tmp = os.path.join(tempfile.mkdtemp(), "point_cloud.parquet")
pd.DataFrame(np.concatenate([sparse, rgb], axis=1), columns=["x", "y", "z", "r", "g", "b"]).to_parquet(tmp)
scene = GaussianPointCloudScene.from_parquet(tmp, GaussianPointCloudScene.PointCloudSceneConfig(
    max_num_points_ratio=10.0, add_sphere=True, initial_alpha=0.05, max_initial_covariance=3000.0,
    initial_covariance_ratio=0.1), generator=g).to(dev)
cfg = GaussianPointCloudTrainer.TrainConfig(num_iterations=iters, feature_learning_rate=5e-3, position_learning_rate=1e-5,
                                            position_learning_rate_decay_rate=0.9947, rasterisation_config=rast_cfg)
ac = cfg.adaptive_controller_config
ac.num_iterations_warm_up, ac.num_iterations_densify = 1000, 100
ac.densification_view_space_position_gradients_threshold = 3e-6
ac.transparent_alpha_threshold, ac.reset_alpha_value, ac.num_iterations_reset_alpha = -2.0, -1.9, 4000
cfg.loss_function_config.enable_regularization = False
fused = "--fused" in sys.argv  # fused image loss (gsb200_image_loss), Adam (gsb200_adam_step) and controller update (gsb200_controller_update)
trainer = GaussianPointCloudTrainer(cfg, scene, views, fused_image_loss=fused, fused_adam=fused, fused_controller_update=fused)
psnr0 = trainer.validation(views[::5])
torch.cuda.synchronize(); t0 = time.perf_counter()
hist = trainer.train(log_interval=50)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
psnr1 = trainer.validation(views[::5])
print(json.dumps({"config": "C5 tat_truck_every_8_test.yaml shapes, synthetic targets", "iterations": iters, "fused_step": fused,
                  "seconds": round(dt, 2), "iterations_per_s": round(iters / dt, 1), "psnr_before": round(psnr0, 2),
                  "psnr_after": round(psnr1, 2), "points_allocated": int(scene.point_cloud.shape[0]),
                  "points_valid_start": hist[0]["num_valid_points"], "points_valid_end": hist[-1]["num_valid_points"],
                  "loss_first": round(hist[0]["loss"], 4), "loss_last": round(hist[-1]["loss"], 4)}))
