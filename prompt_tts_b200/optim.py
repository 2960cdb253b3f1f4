"""The reference's optimiser lines (train.py:41-47 AdamW(lr 1e-5, betas (0.95, 0.999), weight_decay 1e-6, eps 1e-8);
train.py:116-120 clip_grad_norm_(1.0) -> step -> zero_grad) as two kernels over flat buffers.

`DenoiserTrainStep` already leaves all gradients of a step in one flat fp32 buffer (dp.GradSync, laid out in the order
the backward sweep completes them, all-reduced in place).  `FusedClipAdamW` lays the parameters out in the same order in
one flat master buffer (every `nn.Parameter` becomes a view into it, so `state_dict()`, checkpoints and the module tree
are unchanged), and a step is: one sum-of-squares over the gradient buffer, one AdamW kernel over (p, g, m, v) that
applies the clip factor from the device-side norm -- no host synchronisation, no per-tensor launches (707 tensors).
Parameters that never receive a gradient (the 32 dead `proj_out` tensors) are left untouched, as torch.optim.AdamW does
for `grad is None`.
"""
from __future__ import annotations

import torch

from . import ops


class FusedClipAdamW:
    def __init__(self, stepper, lr: float = 1e-5, betas=(0.95, 0.999), weight_decay: float = 1e-6, eps: float = 1e-8,
                 max_norm: float = 1.0):
        self.stepper = stepper
        self.lr, self.betas, self.wd, self.eps, self.max_norm = lr, betas, weight_decay, eps, max_norm
        self.t = 0
        self.pflat = self.m = self.v = self.gnorm_sq = None

    def _build(self) -> None:
        gs = self.stepper.grad_sync
        if gs is None or gs.layout is None:
            raise ops._lib.PtError("FusedClipAdamW.step: run one DenoiserTrainStep first (it defines the flat gradient layout)")
        dev = gs.flat.device
        self.pflat = torch.empty(gs.total, dtype=torch.float32, device=dev)
        for params, off, n in gs.groups:
            o = off
            for p in params:
                k = p.numel()
                self.pflat[o:o + k].copy_(p.data.reshape(-1))
                p.data = self.pflat[o:o + k].view(p.shape)       # the module tree now reads / checkpoints the master buffer
                o += k
            assert o == off + n
        self.m = torch.zeros_like(self.pflat)
        self.v = torch.zeros_like(self.pflat)
        self.gnorm_sq = torch.zeros((), dtype=torch.float32, device=dev)

    @torch.no_grad()
    def step(self) -> torch.Tensor:
        """Clip to max_norm (global L2 norm over all gradients, like clip_grad_norm_) and apply AdamW in place.
        Returns the squared gradient norm as a 0-d device tensor (no synchronisation)."""
        if self.pflat is None:
            self._build()
        gs = self.stepper.grad_sync
        self.t += 1
        self.gnorm_sq.zero_()
        ops.call("sumsq_f32", ops._p(gs.flat), gs.total, ops._p(self.gnorm_sq), ops._stream())
        ops.call("adamw_step", ops._p(self.pflat), ops._p(gs.flat), ops._p(self.m), ops._p(self.v), gs.total, self.lr, self.betas[0],
                 self.betas[1], self.eps, self.wd, self.t, ops._p(self.gnorm_sq), self.max_norm, 1.0, ops._stream())
        self.stepper.cache.epoch += 1        # packed bf16 weight copies are stale now
        return self.gnorm_sq
