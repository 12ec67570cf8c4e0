"""Adam with torch.optim.Adam's semantics on one hand-written kernel per parameter (rl_adam_step).

Drop-in for ``torch.optim.Adam(params, lr, betas, eps, weight_decay)`` (what run_predictorplus.py:51
builds); the trainer works with either.  Parameters without a gradient are skipped, like torch."""
import torch

from . import _lib
from .engine import _stream


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0):
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError("invalid Adam hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        L = _lib.lib()
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                _lib.require_cuda(p, "parameter")
                if p.dtype != torch.float32 or not p.is_contiguous():
                    raise _lib.RlError("rnnlogic_b200.optim.Adam handles contiguous fp32 parameters")
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p)
                    st["exp_avg_sq"] = torch.zeros_like(p)
                st["step"] += 1
                g = p.grad.contiguous().float()
                _lib.check(L.rl_adam_step(p.numel(), p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(),
                                          st["exp_avg_sq"].data_ptr(), group["lr"], b1, b2, group["eps"],
                                          group["weight_decay"], st["step"], _stream()), "rl_adam_step")
        return loss
