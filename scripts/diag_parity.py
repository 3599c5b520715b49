"""GPU diagnostic: error statistics of the CUDA path vs the oracle (not a test)."""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import oracle_backward, oracle_forward
from gpu_helpers import cuda_scene, make_op, n, run_forward
from taichi_3d_gaussian_splatting_b200.synthetic import CONFIGS, make_scene


def stats(name, a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    scale = np.abs(b).max()
    d = np.abs(a - b)
    rel = d / np.maximum(np.abs(b), 1e-6 * scale)
    i = np.unravel_index(np.argmax(rel), rel.shape)
    print(f"{name:28s} max|b|={scale:.3e} max abs err={d.max():.3e} ({d.max()/scale:.2e} of max) "
          f"max rel={rel.max():.3e} at {i}: got {a[i]:.6e} exp {b[i]:.6e}; "
          f"p99 rel={np.percentile(rel, 99):.2e} p99.9={np.percentile(rel, 99.9):.2e} n(rel>1e-3)={(rel > 1e-3).sum()}")


def run(scene, band, exact, label, **cfg):
    o, fwd, feats_n = oracle_forward(scene, **cfg)
    sc = cuda_scene(scene, requires_grad=True)
    cap = {}
    op = make_op(hook=lambda h: cap.setdefault("h", h), exact_exp=exact, **cfg)
    image, depth, count = run_forward(op, sc, band=band)
    print(f"== {label} exact_exp={exact} M={op.last_frame.num_points_in_camera} K={op.last_frame.num_keys}")
    di = np.abs(n(image) - fwd.image)
    print(f"image max abs err {di.max():.3e}  n(>1e-4)={(di > 1e-4).sum()}  n(>1e-5)={(di > 1e-5).sum()}  "
          f"count mismatches {(n(count) != fwd.pixel_valid_point_count).sum()}  depth max err {np.abs(n(depth) - fwd.depth).max():.3e}")
    g = torch.Generator().manual_seed(1)
    grad_image = torch.randn(image.shape, generator=g, dtype=torch.float32)
    image.backward(grad_image.cuda())
    bwd = oracle_backward(o, fwd, scene, feats_n, grad_image.numpy(), band)
    stats("grad_xyz", n(sc.point_cloud.grad), bwd.grad_pointcloud)
    gf = n(sc.point_cloud_features.grad)
    for nm, sl in (("grad_q", slice(0, 4)), ("grad_s", slice(4, 7)), ("grad_logit", slice(7, 8)), ("grad_sh", slice(8, 56))):
        stats(nm, gf[:, sl], bwd.grad_pointcloud_features[:, sl])
    h = cap["h"]
    stats("hook grad_viewspace", n(h.grad_viewspace), bwd.grad_viewspace)
    stats("hook magnitude", n(h.magnitude_grad_viewspace), bwd.magnitude_grad_viewspace)
    stats("hook mag_on_image", n(h.magnitude_grad_viewspace_on_image), bwd.magnitude_grad_viewspace_on_image)
    print("n_affected mismatch", (n(h.num_affected_pixels) != bwd.num_affected_pixels).sum())


if __name__ == "__main__":
    for exact in (True, False):
        run(make_scene(**CONFIGS["C1"]), 0, exact, "C1")
    run(make_scene(4000, 128, 128, 0.05, seed=7, sh_degree=3, yaw_degrees=3.0), 3, False, "smoke")
    if len(sys.argv) > 1:
        run(make_scene(**CONFIGS["C2"]), 3, False, "C2")
