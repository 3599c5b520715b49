"""``FusedAdam``: torch.optim.Adam as the reference trainer uses it (``GaussianPointTrainer.py:126-129``: one instance
per parameter tensor, betas (0.9, 0.999), eps 1e-8, no weight decay, no amsgrad) with the update done by ONE CUDA kernel
per tensor (``gsb200_adam_step``, csrc/adam.cu) instead of torch's multi-kernel foreach path.  It is a
``torch.optim.Optimizer``, so ``ExponentialLR`` (position-LR decay, :131-132) and ``zero_grad`` work unchanged.
CUDA float32 contiguous parameters only: there is no CPU path."""
import torch

from . import _lib


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1 and 0 <= betas[1] < 1):
            raise ValueError("invalid Adam hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        for group in self.param_groups:
            beta1, beta2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("FusedAdam needs contiguous CUDA float32 parameters (there is no CPU path)")
                grad = p.grad.contiguous()
                state = self.state[p]
                if not state:
                    state["step"] = 0
                    state["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    state["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                state["step"] += 1
                with torch.cuda.device(p.device):
                    stream = torch.cuda.current_stream(p.device).cuda_stream
                    _lib.check(lib.gsb200_adam_step(p.data_ptr(), grad.data_ptr(), state["exp_avg"].data_ptr(),
                                                    state["exp_avg_sq"].data_ptr(), p.numel(), float(group["lr"]),
                                                    float(beta1), float(beta2), float(group["eps"]), int(state["step"]),
                                                    stream), "gsb200_adam_step")
        return loss
