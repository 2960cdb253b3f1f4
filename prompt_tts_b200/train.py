"""The reference's training step (train.py:86-120) on the B200 path, without the autograd bridge:

    noise ~ N(0,1); t ~ U{0..999}; x_t = sqrt(acp_t) x0 + sqrt(1-acp_t) noise      (DDPMScheduler.add_noise, :96-98)
    pred = model(x_t, t, ids, mask).sample                                          (:100-105)
    loss = mse(pred, noise); loss.backward()                                        (:107,115)

`DenoiserTrainStep` runs add_noise -> tape forward -> MSE (+ its gradient) -> tape backward with static buffers so the
whole step can be captured in one CUDA graph; parameter gradients land in `param.grad` (fp32, reference layout), so the
reference's `clip_grad_norm_` / `AdamW` lines keep working on top of it.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import engine as E
from . import ops


def ddpm_tables(n: int = 1000, beta_start: float = 1e-4, beta_end: float = 0.02, device="cuda"):
    """sqrt(alphas_cumprod), sqrt(1 - alphas_cumprod) of the linear-beta DDPM schedule (train.py:32-36)."""
    betas = torch.linspace(beta_start, beta_end, n, dtype=torch.float32)
    acp = torch.cumprod(1.0 - betas, 0)
    return (acp ** 0.5).to(device), ((1 - acp) ** 0.5).to(device)


class DenoiserTrainStep:
    def __init__(self, model, grad_sync=None):
        """grad_sync: optional object with `on_backward_start(tape)`, `on_grads_ready(list_of_grads)` and `finish()`
        (see prompt_tts_b200.dp.GradSync) used to overlap the gradient all-reduce with the backward sweep."""
        self.model = model
        self.params = [p for p in model.parameters()]
        self.cache = E.get_cache(model)
        self.sa: Optional[torch.Tensor] = None
        self.sb: Optional[torch.Tensor] = None
        if grad_sync is None:      # single GPU: still use the flat gradient buffer (one memset per step, no per-tensor zero fills)
            from .dp import GradSync
            grad_sync = GradSync(model, world_size=1)
        self.grad_sync = grad_sync
        self.loss = None

    def __call__(self, x0, noise, t, ids, mask=None, loss_out: Optional[torch.Tensor] = None, gscale: float = 1.0):
        """x0, noise fp32 [B, C, T]; t int64 [B]; ids int32 [B, Lt].  Returns the loss (0-d fp32 CUDA tensor)."""
        if not x0.is_cuda:
            raise ops._lib.PtError("DenoiserTrainStep: inputs must be CUDA tensors; there is no CPU fallback")
        if self.sa is None:
            self.sa, self.sb = ddpm_tables(device=x0.device)
        B = x0.shape[0]
        xt = torch.empty_like(x0)
        ops.call("add_noise", ops._p(x0), ops._p(noise), ops._p(t), ops._p(self.sa), ops._p(self.sb), ops._p(xt), B, x0[0].numel(), ops._stream())
        tape = E.Tape(self.cache, recording=True)
        if self.grad_sync is not None:
            self.grad_sync.attach(tape)
        enc = self.model.text_encoder._fwd(tape, ids)
        pred, seed = self.model.unet._fwd(tape, xt, t, enc)
        loss = torch.zeros((), dtype=torch.float32, device=x0.device) if loss_out is None else loss_out.zero_()
        dpred = torch.empty_like(pred)
        ops.call("mse_fwd_bwd", ops._p(pred), ops._p(noise), ops._p(loss), ops._p(dpred), pred.numel(), gscale, ops._stream())
        seed([dpred])
        tape.backward()
        if self.grad_sync is not None:
            self.grad_sync.finish()
        for p in self.params:
            g = tape.pgrads.get(id(p))
            if g is None:
                continue
            if p.grad is None:
                p.grad = g
            elif p.grad.data_ptr() != g.data_ptr():
                p.grad.add_(g)
        self.loss = loss
        return loss
