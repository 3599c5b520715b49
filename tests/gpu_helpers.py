"""Helpers for the -m gpu parity tests: run the CUDA operator on a SyntheticScene."""
import numpy as np
import torch

from taichi_3d_gaussian_splatting_b200 import GaussianPointCloudRasterisation as GPCR

Config = GPCR.GaussianPointCloudRasterisationConfig
Input = GPCR.GaussianPointCloudRasterisationInput


def make_op(hook=None, exact_exp=False, force_key64=False, initial_key_capacity=None, keep_all_tile_pairs=False, **cfg):
    return GPCR(Config(**cfg), backward_valid_point_hook=hook, exact_exp=exact_exp, force_key64=force_key64,
                initial_key_capacity=initial_key_capacity, keep_all_tile_pairs=keep_all_tile_pairs)


def cuda_scene(scene, requires_grad=False):
    sc = scene.to("cuda")
    if requires_grad:
        sc.point_cloud.requires_grad_(True)
        sc.point_cloud_features.requires_grad_(True)
    return sc


def run_forward(op, sc, band=3):
    return op(Input(point_cloud=sc.point_cloud, point_cloud_features=sc.point_cloud_features,
                    point_object_id=sc.point_object_id, point_invalid_mask=sc.point_invalid_mask,
                    camera_info=sc.camera_info, q_pointcloud_camera=sc.q_pointcloud_camera,
                    t_pointcloud_camera=sc.t_pointcloud_camera, color_max_sh_band=band))


def n(t):
    return t.detach().cpu().numpy()


def count_above(a, b, tol):
    return int((np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)) > tol).sum())
