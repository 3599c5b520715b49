"""Deterministic scenario shared by make_controller_golden.py (reference side) and tests/test_trainer_cpu.py (our side):
60 slots (40 hold Gaussians), five backward-hook calls with seeded statistics, a config that makes every branch of the
controller fire within five iterations (transparent / floater removal, clone, split, opacity reset).  The two stochastic
options (sample_from_point, ellipsoid offset -- Taichi kernels in the reference) are switched off."""
import torch

CONFIG = dict(num_iterations_warm_up=1, num_iterations_densify=2, transparent_alpha_threshold=-0.5,
              densification_view_space_position_gradients_threshold=0.5,
              densification_view_avg_space_position_gradients_threshold=0.2,
              densification_multi_frame_view_space_position_gradients_threshold=0.9,
              densification_multi_frame_view_pixel_avg_space_position_gradients_threshold=1e3,
              densification_multi_frame_position_gradients_threshold=1.6,
              gaussian_split_factor_phi=1.6, num_iterations_reset_alpha=4, reset_alpha_value=0.1,
              floater_near_camrea_num_pixels_threshold=80, floater_depth_threshold=1.5,
              iteration_start_remove_floater=0, plot_densify_interval=10 ** 9,
              under_reconstructed_num_pixels_threshold=60, under_reconstructed_move_factor=100.0,
              enable_ellipsoid_offset=False, enable_sample_from_point=False)
N, N_VALID, ITERATIONS = 60, 40, 5


def initial_state():
    g = torch.Generator().manual_seed(11)
    xyz = torch.randn((N, 3), generator=g)
    feat = torch.randn((N, 56), generator=g)
    feat[:, 7] = torch.rand((N,), generator=g) * 3 - 1.0  # some below the transparent threshold
    feat[5, 20] = float("nan")                             # a NaN row is treated as transparent
    mask = torch.zeros(N, dtype=torch.int8)
    mask[N_VALID:] = 1
    obj = (torch.arange(N) % 3).to(torch.int32)
    return xyz, feat, mask, obj


def hook_fields(iteration, mask):
    """Synthetic BackwardValidPointHookInput fields for the points that are valid at this iteration."""
    g = torch.Generator().manual_seed(100 + iteration)
    ids = torch.nonzero(mask == 0).flatten()
    keep = torch.rand(ids.shape, generator=g) < 0.8  # ~80 % of the valid points are in the frustum
    ids = ids[keep].to(torch.int32)
    m = ids.shape[0]
    pixels = torch.randint(0, 120, (m,), generator=g, dtype=torch.int32)
    pixels[::7] = 0  # never touched a pixel: 0/0 in the averages
    return dict(
        point_id_in_camera_list=ids,
        grad_point_in_camera=torch.randn((m, 3), generator=g),
        grad_pointfeatures_in_camera=torch.randn((m, 56), generator=g),
        grad_viewspace=torch.randn((m, 2), generator=g),
        magnitude_grad_viewspace=torch.rand((m,), generator=g),
        magnitude_grad_viewspace_on_image=torch.rand((32, 32, 2), generator=g),
        num_overlap_tiles=torch.randint(1, 5, (m,), generator=g, dtype=torch.int32),
        num_affected_pixels=pixels,
        point_depth=torch.rand((m,), generator=g) * 3,
        point_uv_in_camera=torch.rand((m, 2), generator=g) * 32)
