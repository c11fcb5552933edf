"""Dense Adam over the embedding tables with torch.optim.Adam's arithmetic
(/root/reference/LightGCN_SPEX/code/main_rec.py:23,37), one fused kernel per table
(spex_adam_f32: streams p, g, m, v once — 28 bytes per element).

Because propagation touches every row, gradients of both tables are dense, so the
optimiser is pure HBM streaming.  When both embedding weights are views of one fused table
(spex_b200.model.LightGCN) and their gradients are views of one buffer, a single launch updates
the whole table.
"""
from __future__ import annotations

import torch

from . import ops


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda:
                    raise RuntimeError("FusedAdam runs on sm_100a only (no CPU fallback)")
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["step"] += 1
                if not p.is_contiguous():
                    raise RuntimeError("FusedAdam needs contiguous parameters")
                ops.adam_step(p, p.grad, st["exp_avg"], st["exp_avg_sq"], group["lr"], b1, b2,
                              group["eps"], st["step"])
        return loss
