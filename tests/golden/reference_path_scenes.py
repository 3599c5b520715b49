"""Small scenes for the full-path golden vectors (shared by make_reference_path_golden.py, which runs the REFERENCE's
kernels on them, and by the tests, which run the oracle / the CUDA operator on them).  Every scene is a dict of torch
CPU tensors + config; generated from seeds, nothing is read from disk."""
import math

import torch


def _base(num_points, height, width, sigma, seed, sh_degree=3, yaw=0.0):
    g = torch.Generator().manual_seed(seed)
    u = torch.rand((num_points, 3), generator=g)
    xyz = torch.stack([u[:, 0] * 8 - 4, u[:, 1] * 5 - 2.5, u[:, 2] * 8 + 2], dim=-1)
    q = torch.randn((num_points, 4), generator=g)
    q = q / q.norm(dim=-1, keepdim=True)
    s = torch.randn((num_points, 3), generator=g) * 0.5 + math.log(sigma)
    logit = torch.rand((num_points, 1), generator=g) * 6 - 3
    sh = torch.zeros((num_points, 3, 16))
    sh[:, :, 0] = torch.randn((num_points, 3), generator=g) * 1.5
    n_rest = (sh_degree + 1) ** 2 - 1
    if n_rest:
        sh[:, :, 1:1 + n_rest] = torch.randn((num_points, 3, n_rest), generator=g) * 0.2
    half = math.radians(yaw) / 2.0
    return dict(
        point_cloud=xyz.contiguous(), point_cloud_features=torch.cat([q, s, logit, sh.reshape(num_points, 48)], dim=-1).contiguous(),
        point_invalid_mask=torch.zeros(num_points, dtype=torch.int8), point_object_id=torch.zeros(num_points, dtype=torch.int32),
        camera_intrinsics=torch.tensor([[0.6 * width, 0.0, width / 2.0], [0.0, 0.6 * width, height / 2.0], [0.0, 0.0, 1.0]]),
        camera_height=height, camera_width=width,
        q_pointcloud_camera=torch.tensor([[0.0, math.sin(half), 0.0, math.cos(half)]]), t_pointcloud_camera=torch.zeros((1, 3)),
        color_max_sh_band=3, near_plane=0.8, far_plane=1000.0, depth_to_sort_key_scale=100.0, grad_seed=seed + 1000)


def baseline_config_1():
    """BASELINE config 1 (the correctness gate): 1e4 Gaussians, 256 x 256, sigma_med 0.03, seed 0, SH deg 0.
    Same generator as taichi_3d_gaussian_splatting_b200.synthetic.make_scene(**CONFIGS["C1"])."""
    c1 = _base(10_000, 256, 256, 0.03, 0, sh_degree=0)
    c1.update(color_max_sh_band=0)
    return c1


def reduced_config_2():
    """BASELINE config 2 at a tenth of its Gaussians (SURVEY 8(d) names "a reduced C2 (N = 4.3e4)" as the CPU-feasible
    stand-in): 976 x 544 = 61 x 34 tiles, seed 1, SH deg 3; sigma_med 0.02 * 10^(1/3) keeps the number of splats per tile
    in C2's range (mean ~240, max ~400: several 256-splat groups per tile)."""
    return _base(43_000, 544, 976, 0.02 * 10 ** (1.0 / 3.0), 1, sh_degree=3)


def scenes():
    out = {}
    # A: rotated camera, un-normalised quaternions (normalised in place by the forward), unused slots, SH deg 3
    a = _base(60, 32, 48, 0.08, 3, yaw=5.0)
    a["point_cloud_features"][:, :4] *= 1.3
    a["point_invalid_mask"][::9] = 1
    out["A_basic"] = a
    # B: dense and opaque: every tile list is longer than one 256-splat group, pixels saturate (early termination,
    # the splat that would saturate a pixel is not blended), the 0.99 clamp is hit
    b = _base(520, 32, 32, 0.6, 4)
    b["point_cloud"][:, 2] = b["point_cloud"][:, 2] * 0.5 + 1.5
    b["point_cloud_features"][:, 7] += 3.0
    out["B_dense_saturating"] = b
    # C: two objects with their own camera->pointcloud poses, SH band 1 (gradient masking), coarse depth keys (ties in
    # the sort key resolved by the stable order), closer near plane
    c = _base(90, 48, 32, 0.1, 5, yaw=-8.0)
    c["point_object_id"][::2] = 1
    c["q_pointcloud_camera"] = torch.tensor([[0.0, math.sin(-0.07), 0.0, math.cos(-0.07)], [0.05, 0.02, -0.03, 0.99]])
    c["t_pointcloud_camera"] = torch.tensor([[0.0, 0.0, 0.0], [0.3, -0.2, 0.5]])
    c.update(color_max_sh_band=1, depth_to_sort_key_scale=2.0, near_plane=0.4)
    out["C_two_objects_band1_ties"] = c
    # D: frustum borders and the bounding-box quirk: splats left / above the image still get tile column / row 0,
    # splats right / below get none; points behind the camera and beyond the far plane; SH band 0
    d = _base(80, 32, 32, 0.12, 6, sh_degree=0)
    d["point_cloud"][:20, 0] = -3.4 - 0.05 * torch.arange(20)   # off to the left, inside the 48-pixel margin for some
    d["point_cloud"][20:35, 1] = -2.2 - 0.03 * torch.arange(15)  # off the top
    d["point_cloud"][35:45, 0] = 3.5 + 0.05 * torch.arange(10)   # off to the right
    d["point_cloud"][45:50, 2] = -1.0                             # behind the camera
    d["point_cloud"][50:55, 2] = 20.0                             # beyond the far plane below
    d.update(color_max_sh_band=0, far_plane=15.0)
    out["D_borders_band0"] = d
    # E: 3 x 4 tiles, splats from sub-pixel (the +0.3 low-pass / rescale path dominates) to screen-filling, opacities from
    # nearly 0 (never reaches 1/255) to nearly 1, SH band 2
    e = _base(150, 48, 64, 0.05, 7, yaw=12.0)
    e["point_cloud_features"][:30, 4:7] -= 3.0     # tiny: projected sigma well below a pixel
    e["point_cloud_features"][30:45, 4:7] += 2.5   # huge: cover every tile
    e["point_cloud_features"][45:60, 7] = -7.0     # almost transparent
    e["point_cloud_features"][60:75, 7] = 7.0      # almost opaque
    e.update(color_max_sh_band=2, depth_to_sort_key_scale=10.0)
    out["E_extreme_scales_band2"] = e
    return out
