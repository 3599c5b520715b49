"""Fused clamp + L1 loss + gradient (gsb200_l1_loss) against torch autograd on the same tensors."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _torch_ref(pred, gt, clamp01, weight=1.0):
    p = pred.clone().requires_grad_(True)
    x = torch.clamp(p, min=0, max=1) if clamp01 else p  # GaussianPointTrainer.py:168-170
    loss = torch.abs(x - gt).mean()  # LossFunction.py:29
    (weight * loss).backward()
    return loss.detach(), p.grad


@pytest.mark.parametrize("shape", [(64, 96, 3), (1, 7), (1072, 1920, 3), (5,)])
@pytest.mark.parametrize("clamp01", [False, True])
def test_fused_l1_matches_torch(shape, clamp01):
    from taichi_3d_gaussian_splatting_b200 import fused_l1_loss_with_grad
    g = torch.Generator().manual_seed(3)
    pred = (torch.rand(shape, generator=g) * 1.6 - 0.3).cuda()
    gt = torch.rand(shape, generator=g).cuda()
    flat = pred.view(-1)
    flat[0] = 0.0      # clamp boundaries pass the gradient (torch.clamp semantics)
    flat[1] = 1.0
    gt.view(-1)[2] = float(flat[2])  # exact tie: sign(0) = 0
    exp_loss, exp_grad = _torch_ref(pred, gt, clamp01, weight=0.8)
    loss, grad = fused_l1_loss_with_grad(pred, gt, clamp01=clamp01, weight=0.8)
    assert abs(float(loss) - float(exp_loss)) <= 2e-6 * max(1.0, abs(float(exp_loss)))
    # torch scales by (1/n) as a product, the kernel divides: equal up to 1 ulp, identical sign pattern
    assert torch.equal(torch.sign(grad), torch.sign(exp_grad)) and torch.allclose(grad, exp_grad, rtol=1e-6, atol=0)
    # second call on the same temp buffer, deterministic
    loss2, grad2 = fused_l1_loss_with_grad(pred, gt, clamp01=clamp01, weight=0.8)
    assert float(loss2) == float(loss) and torch.equal(grad2, grad)


def test_fused_l1_autograd_wrapper_and_errors():
    from taichi_3d_gaussian_splatting_b200 import fused_l1_loss
    g = torch.Generator().manual_seed(4)
    pred = torch.rand((32, 48, 3), generator=g).cuda().requires_grad_(True)
    gt = torch.rand((32, 48, 3), generator=g).cuda()
    loss = fused_l1_loss(pred, gt, clamp01=True)
    (3.0 * loss).backward()
    exp_loss, exp_grad = _torch_ref(pred.detach(), gt, True, weight=3.0)
    assert abs(float(loss) - float(exp_loss)) <= 2e-6
    assert torch.allclose(pred.grad, exp_grad, rtol=1e-6, atol=0)
    with pytest.raises(RuntimeError, match="no CPU path"):
        fused_l1_loss(pred.detach().cpu(), gt.cpu())
