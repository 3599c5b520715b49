"""The fused image loss of the trainer step (``csrc/image_loss.cu``: clamp + L1 + D-SSIM and the gradient w.r.t. the
rasterised image, two kernels) executed on the CPU under the SIMT emulator -- the unmodified kernel source -- and compared
with the package's torch restatement of the reference's loss (``loss.py``: GaussianPointTrainer.py:168-175 clamp / permute,
LossFunction.py:20-38 with the published pytorch_msssim algorithm) and torch autograd.  Test infrastructure."""
import ctypes

import numpy as np
import pytest
import torch

from taichi_3d_gaussian_splatting_b200.loss import LossFunction
from simt_helpers import build_emulator


@pytest.fixture(scope="module")
def emu():
    return build_emulator()


def _torch_reference(pred_hwc, gt_chw, lam, upstream):
    pred = torch.tensor(pred_hwc, requires_grad=True)
    loss_fn = LossFunction(LossFunction.LossFunctionConfig(lambda_value=lam, enable_regularization=False))
    image = torch.clamp(pred, min=0, max=1).permute(2, 0, 1)  # GaussianPointTrainer.py:168-171
    loss, l1, dssim = loss_fn(image, torch.tensor(gt_chw))
    (loss * upstream).backward()
    return float(loss), float(l1), float(dssim), pred.grad.numpy()


@pytest.mark.parametrize("H,W,lam,upstream", [(32, 48, 0.2, 1.0), (37, 29, 0.2, 1.0), (16, 16, 1.0, 0.5), (64, 21, 0.0, 2.0)])
def test_fused_image_loss_source_matches_the_torch_loss(emu, H, W, lam, upstream):
    rng = np.random.default_rng(H * 100 + W)
    # smooth-ish images with values outside [0, 1] (the clamp must stop their gradient) and exact ties with the target
    gt = rng.random((3, H, W)).astype(np.float32)
    pred = (gt.transpose(1, 2, 0) + 0.3 * rng.standard_normal((H, W, 3))).astype(np.float32)
    pred[::5, ::3] = gt.transpose(1, 2, 0)[::5, ::3]
    pred = np.ascontiguousarray(pred)
    temp = np.zeros(int(emu.emu_image_loss_temp_bytes(H, W)) + 16, np.uint8)
    out, grad = np.zeros(3, np.float32), np.full((H, W, 3), 9.0, np.float32)
    c = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    f = ctypes.c_float
    for _ in range(2):  # second call on the same temp buffer: the ticket must have been re-armed
        assert emu.emu_image_loss(c(pred), c(gt), H, W, f(lam), f(upstream), c(out), c(grad), c(temp)) > 0
    loss, l1, dssim, ref_grad = _torch_reference(pred, gt, lam, upstream)
    assert abs(out[0] - loss) <= 2e-6 * max(1.0, abs(loss)) and abs(out[1] - l1) <= 2e-6 and abs(out[2] - dssim) <= 5e-6
    scale = np.abs(ref_grad).max()
    assert np.abs(grad - ref_grad).max() <= 2e-5 * scale
    assert (grad[(pred < 0) | (pred > 1)] == 0).all()
    # loss only (validation): no gradient buffer
    out2 = np.zeros(3, np.float32)
    assert emu.emu_image_loss(c(pred), c(gt), H, W, f(lam), f(upstream), c(out2), None, c(temp)) > 0
    assert np.array_equal(out2, out)


@pytest.mark.parametrize("n", [7, 4096, 59 * 37])
def test_fused_adam_source_follows_torch_adam(emu, n):
    """csrc/adam.cu under the emulator vs torch.optim.Adam as the reference trainer configures it
    (GaussianPointTrainer.py:126-129), five steps with a decaying learning rate, grid-stride and tail paths."""
    rng = np.random.default_rng(n)
    w0 = rng.standard_normal(n).astype(np.float32)
    ref = torch.tensor(w0, requires_grad=True)
    opt = torch.optim.Adam([ref], lr=5e-3, betas=(0.9, 0.999))
    sched = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=0.97)
    w, m, v = w0.copy(), np.zeros(n, np.float32), np.zeros(n, np.float32)
    c = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    d = ctypes.c_double
    for step in range(1, 6):
        g = (rng.standard_normal(n) * (10.0 ** rng.integers(-3, 2))).astype(np.float32)
        g[::9] = 0.0
        ref.grad = torch.tensor(g)
        lr = opt.param_groups[0]["lr"]
        opt.step()
        sched.step()
        emu.emu_adam_step(c(w), c(g), c(m), c(v), ctypes.c_longlong(n), d(lr), d(0.9), d(0.999), d(1e-8), step, 2)
        assert np.abs(w - ref.detach().numpy()).max() <= 1e-6 * max(1.0, float(np.abs(w0).max()))
        st = opt.state[ref]
        # same formulas, float32: torch's CPU lerp / addcmul may fuse multiply-adds, so allow a few ulps
        em, ev = st["exp_avg"].numpy(), st["exp_avg_sq"].numpy()
        assert np.allclose(m, em, rtol=2e-6, atol=1e-6 * np.abs(em).max())
        assert np.allclose(v, ev, rtol=2e-6, atol=1e-6 * np.abs(ev).max())


@pytest.mark.parametrize("n,clamp", [(5, True), (3 * 64 * 96, True), (300001, False)])
def test_fused_l1_source_matches_torch(emu, n, clamp):
    """csrc/loss.cu (gsb200_l1_loss: clamp + mean |pred - gt| + gradient, deterministic two-level sum) under the emulator."""
    emu.emu_l1_loss_temp_bytes.restype = ctypes.c_longlong
    rng = np.random.default_rng(n)
    gt = rng.random(n).astype(np.float32)
    pred = (gt + 0.4 * rng.standard_normal(n)).astype(np.float32)
    pred[::7] = gt[::7]
    temp = np.zeros(int(emu.emu_l1_loss_temp_bytes()), np.uint8)
    loss, grad = np.zeros(1, np.float32), np.full(n, 9.0, np.float32)
    c = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    for _ in range(2):  # the ticket is re-armed by the call
        emu.emu_l1_loss(c(pred), c(gt), ctypes.c_longlong(n), int(clamp), ctypes.c_float(2.0), c(loss), c(grad), c(temp))
    p = torch.tensor(pred, requires_grad=True)
    ref = ((torch.clamp(p, 0, 1) if clamp else p) - torch.tensor(gt)).abs().mean()
    (2.0 * ref).backward()
    assert abs(float(loss[0]) - float(ref)) <= 1e-6 * max(1.0, float(ref))
    assert np.allclose(grad, p.grad.numpy(), rtol=1e-6, atol=0)


def test_fused_controller_update_source_matches_the_torch_update(emu):
    """csrc/controller.cu under the emulator vs GaussianPointAdaptiveController.update's torch ops
    (GaussianPointAdaptiveController.py:130-143), incl. the 0/0 -> 0 rule for splats without affected pixels."""
    from taichi_3d_gaussian_splatting_b200 import GaussianPointAdaptiveController as C
    from taichi_3d_gaussian_splatting_b200 import GaussianPointCloudRasterisation as GPCR
    g = torch.Generator().manual_seed(2)
    N, M = 3000, 1100
    ctl = C(C.GaussianPointAdaptiveControllerConfig(num_iterations_warm_up=10 ** 9),
            C.GaussianPointAdaptiveControllerMaintainedParameters(
                pointcloud=torch.zeros((N, 3)), pointcloud_features=torch.zeros((N, 56)),
                point_invalid_mask=torch.zeros(N, dtype=torch.int8), point_object_id=torch.zeros(N, dtype=torch.int32)))
    acc = [np.zeros(N, np.int32), np.zeros(N, np.int32), np.zeros(N, np.float32), np.zeros(N, np.float32),
           np.zeros((N, 3), np.float32), np.zeros(N, np.float32)]
    c = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    for it in range(3):
        ids = torch.randperm(N, generator=g)[:M].sort().values.to(torch.int32)
        npix = torch.randint(0, 50, (M,), generator=g, dtype=torch.int32)
        mag = torch.rand(M, generator=g) * (npix > 0)
        gxyz = torch.randn((M, 3), generator=g)
        ctl.update(GPCR.BackwardValidPointHookInput(
            point_id_in_camera_list=ids, grad_point_in_camera=gxyz, grad_pointfeatures_in_camera=torch.zeros((M, 56)),
            grad_viewspace=torch.zeros((M, 2)), magnitude_grad_viewspace=mag,
            magnitude_grad_viewspace_on_image=torch.zeros((16, 16, 2)), num_overlap_tiles=torch.ones(M, dtype=torch.int32),
            num_affected_pixels=npix, point_depth=torch.ones(M), point_uv_in_camera=torch.zeros((M, 2))))
        emu.emu_controller_update(c(ids.numpy()), ctypes.c_longlong(M), c(npix.numpy()), c(mag.numpy()),
                                  c(np.ascontiguousarray(gxyz.numpy())), *(c(a) for a in acc), 3)
    assert np.array_equal(acc[0], ctl.accumulated_num_in_camera.numpy()) and np.array_equal(acc[1], ctl.accumulated_num_pixels.numpy())
    for got, exp in ((acc[2], ctl.accumulated_view_space_position_gradients), (acc[3], ctl.accumulated_view_space_position_gradients_avg),
                     (acc[4], ctl.accumulated_position_gradients), (acc[5], ctl.accumulated_position_gradients_norm)):
        assert np.allclose(got, exp.numpy(), rtol=1e-6, atol=1e-7)
    assert acc[3].max() > 0 and np.isfinite(acc[3]).all()
