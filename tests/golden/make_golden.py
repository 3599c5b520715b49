"""Generate golden vectors from the REFERENCE's own importable code.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference's Taichi kernels cannot run here (no ``taichi`` wheel), but its pure-torch helpers in
``taichi_3d_gaussian_splatting/utils.py`` import fine once ``taichi`` / ``dataclass_wizard`` are
stubbed.  We call them on the inputs of the reference's known-answer tests and store the outputs:

* ``single_point``  -- ``torch_single_point_alpha_forward`` (utils.py:513-558) + torch autograd on the
  inputs of tests/GaussianPointCloudRasterisation_test.py:353-548 (the reference test's own oracle).
* ``inverse_se3``   -- ``inverse_SE3_qt_torch`` (utils.py:426-432), cf. tests/utils_test.py:141-157.
* ``rot_to_quat``   -- ``SE3_to_quaternion_and_translation_torch`` (utils.py:486-492).
* ``quat_to_rot``   -- ``quaternion_to_rotation_matrix_torch`` (utils.py:596-632).
* ``tile_ranges``   -- literal known answer of tests/GaussianPointCloudRasterisation_test.py:18-51.
* ``cov_projection``-- inputs of tests/GaussianPoint3D_test.py:12-54 with the numpy formula that test uses.
* ``gaussian_2d``   -- inputs of tests/utils_test.py:286-348 evaluated with float64 closed forms
  (the test compares against scipy.stats.multivariate_normal; the normalised variant drops the
  1/(2*pi*sqrt(det)) factor).

The output file ``reference_vectors.json`` is committed; tests read only the JSON.
"""
import contextlib
import io
import json
import os
import sys
import types
from unittest import mock

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors.json")


def import_reference_utils():
    ti = mock.MagicMock()
    # decorators must return the function unchanged so module import succeeds
    ti.func = lambda f: f
    ti.kernel = lambda f: f
    ti.dataclass = lambda c: c
    sys.modules["taichi"] = ti
    sys.modules["taichi.math"] = ti.math
    sys.modules.setdefault("dataclass_wizard", types.SimpleNamespace(YAMLWizard=object))
    sys.path.insert(0, REF)
    from taichi_3d_gaussian_splatting import utils  # noqa: E402
    return utils


def main():
    utils = import_reference_utils()
    torch.manual_seed(0)
    golden = {}

    # ---- single point alpha (+ autograd), GaussianPointCloudRasterisation_test.py:353-380
    T_camera_pointcloud = torch.tensor([[1., 0., 0., 0.], [0., 1., 0., 0.], [0., 0., 1., 2.],
                                        [0., 0., 0., 1.]])
    K = torch.tensor([[32., 0., 16.], [0., 32., 16.], [0., 0., 1.]])
    xyz = torch.tensor([-0.4325, -0.7224, -0.4733], dtype=torch.float32, requires_grad=True)
    features = torch.tensor([
        0.0115, 0.5507, 0.6920, 0.4666, np.log(0.6306), np.log(0.0871), np.log(0.0112), 1.7667,
        2.2963, 0.1560, 0.8710, 0.3418, 0.3658, 0.1913, 0.8727, 0.3608,
        0.6874, 0.7516, 0.9281, 0.5649, 0.9469, 0.9090, 0.7356, 0.5436,
        1.7886, 0.7542, 0.9568, 0.2868, 0.3552, 0.3872, 0.0827, 0.4101,
        0.7783, 0.6266, 0.9601, 0.8252, 0.7846, 0.0183, 0.6635, 0.4688,
        -1.4012, 0.1584, 0.3252, 0.5403, 0.4992, 0.2780, 0.7412, 0.5056,
        0.8236, 0.9722, 0.5467, 0.6644, 0.2583, 0.0953, 0.3986, 0.2265],
        dtype=torch.float32, requires_grad=True)
    pixel_uv = torch.tensor([3, 3])
    with contextlib.redirect_stdout(io.StringIO()):
        alpha = utils.torch_single_point_alpha_forward(
            point_xyz=xyz, point_q=features[:4], point_s=features[4:7],
            T_camera_pointcloud=T_camera_pointcloud, camera_intrinsics=K,
            point_alpha=features[7], pixel_uv=pixel_uv)
        alpha.backward()
    golden["single_point"] = dict(
        T_camera_pointcloud=T_camera_pointcloud.tolist(), camera_intrinsics=K.tolist(),
        xyz=xyz.detach().tolist(), features=features.detach().tolist(), pixel_uv=[3, 3],
        alpha=float(alpha), grad_xyz=xyz.grad.tolist(), grad_features_0_8=features.grad[:8].tolist(),
        atol_alpha=1e-4, atol_grad_xyz=1e-4, atol_grad_features=1e-2)

    # ---- inverse_SE3_qt_torch on random poses
    q = torch.randn(6, 4)
    q = q / q.norm(dim=-1, keepdim=True)
    t = torch.randn(6, 3)
    q_inv, t_inv = utils.inverse_SE3_qt_torch(q, t)
    golden["inverse_se3"] = dict(q=q.tolist(), t=t.tolist(), q_inv=q_inv.tolist(), t_inv=t_inv.tolist())

    # ---- rotation matrix <-> quaternion
    R = utils.quaternion_to_rotation_matrix_torch(q)
    Tm = torch.eye(4).repeat(6, 1, 1)
    Tm[:, :3, :3] = R
    Tm[:, :3, 3] = t
    q_back, t_back = utils.SE3_to_quaternion_and_translation_torch(Tm)
    golden["quat_to_rot"] = dict(q=q.tolist(), R=R.tolist())
    golden["rot_to_quat"] = dict(T=Tm.tolist(), q=q_back.tolist(), t=t_back.tolist())

    # ---- tile ranges known answer (GaussianPointCloudRasterisation_test.py:18-51)
    golden["tile_ranges"] = dict(
        keys=[0x100000000, 0x100000001, 0x200000000, 0x200000001, 0x200000002, 0x300000000,
              0x300000001], num_tiles=4, start=[0, 0, 2, 5], end=[0, 2, 5, 7])

    # ---- covariance projection (GaussianPoint3D_test.py:12-54), numpy formula of that test in f64
    xyz_c = np.array([-0.1316, -0.2471, 1.0090])
    exp_s = np.array([0.7606, 0.9650, 0.1946])
    qc = np.array([0.0229, 0.9774, 0.1204, 0.1725])
    Rq = utils.quaternion_to_rotation_matrix_torch(torch.tensor(qc)).numpy()  # polynomial on raw q
    S = np.diag(exp_s)
    Sigma = Rq @ S @ S.T @ Rq.T
    fx = fy = 32.0
    J = np.array([[fx / xyz_c[2], 0, -fx * xyz_c[0] / xyz_c[2] ** 2],
                  [0, fy / xyz_c[2], -fy * xyz_c[1] / xyz_c[2] ** 2]])
    cov = J @ Sigma @ J.T
    golden["cov_projection"] = dict(xyz=xyz_c.tolist(), exp_s=exp_s.tolist(), q=qc.tolist(),
                                    camera_intrinsics=[[32, 0, 16], [0, 32, 16], [0, 0, 1]],
                                    cov=cov.tolist(), R=Rq.tolist(), rtol=1e-2)

    # ---- 2-D Gaussian density and gradients (utils_test.py:286-348), f64 closed forms
    mean = np.array([2.0, 3.0])
    cov2 = np.array([[1.0, 0.3], [0.3, 1.5]])
    x = np.array([3.0, 4.0])
    d = x - mean
    inv = np.linalg.inv(cov2)
    p = float(np.exp(-0.5 * d @ inv @ d))
    golden["gaussian_2d"] = dict(mean=mean.tolist(), cov=cov2.tolist(), xy=x.tolist(), p_normalized=p,
                                 d_p_d_mean=(p * inv @ d).tolist(),
                                 d_p_d_cov=(0.5 * p * (inv @ np.outer(d, d) @ inv)).tolist())

    with open(OUT, "w") as f:
        json.dump(golden, f, indent=1)
    print("wrote", OUT)
    print(json.dumps(golden["single_point"], indent=1)[:600])


if __name__ == "__main__":
    main()
