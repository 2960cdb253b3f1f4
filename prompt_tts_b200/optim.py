"""The reference's optimiser lines (train.py:41-47 AdamW(lr 1e-5, betas (0.95, 0.999), weight_decay 1e-6, eps 1e-8);
train.py:116-120 clip_grad_norm_(1.0) -> step -> zero_grad) as three kernels over flat buffers.

`DenoiserTrainStep` leaves all gradients of a step in one flat fp32 buffer (dp.GradSync, laid out in the order the backward
sweep completes them, all-reduced in place).  `FusedClipAdamW` keeps, in that same layout,
    pflat  fp32 master parameters -- every `nn.Parameter` becomes a view of it (k=3 conv weights a permuted view of their tap-major
           storage), so `state_dict()`, checkpoints and the module tree are unchanged;
    m, v   AdamW moments;
    wflat  the bf16 shadow of pflat -- the GEMM weight operands (`engine.PackCache` static entries) are views of it, written by
           the AdamW kernel itself, so there is no fp32 -> bf16 re-pack pass after an update (round 1: 292 launches, 3.2 GB).
A step is: sum of squares of the gradient buffer -> `pt_adamw_prepare` (one thread: step += 1, bias corrections, clip factor,
all in DEVICE memory) -> one elementwise AdamW kernel.  No host value is baked into a launch: the step counter, the learning
rate and the clip factor are read from device memory, so a CUDA graph that captured `step()` replays the reference's optimiser
exactly (`tests/test_optim_gpu.py::test_fused_adamw_graph_replays`).  Parameters that never receive a gradient (the 32 dead
`proj_out` tensors) are left untouched, as torch.optim.AdamW does for `grad is None`.
"""
from __future__ import annotations

from typing import Dict

import torch

from . import ops

# slots of the device-side state (include/prompt_tts_b200.h)
LR, BETA1, BETA2, EPS, WD, MAX_NORM, GSCALE, STEP, CLIP, STEP_SIZE, INV_SQRT_BC2, DECAY, GNORM, GNORM_SQ, NSTATE = 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 16
NPARTS = 1184      # blocks of the deterministic norm reduction (8 per SM)


class FusedClipAdamW:
    def __init__(self, stepper, lr: float = 1e-5, betas=(0.95, 0.999), weight_decay: float = 1e-6, eps: float = 1e-8,
                 max_norm: float = 1.0):
        self.stepper = stepper
        self.lr, self.betas, self.wd, self.eps, self.max_norm = lr, betas, weight_decay, eps, max_norm
        self.pflat = self.m = self.v = self.wflat = self.partials = self.state = None

    # ------------------------------------------------------------------------------------------ flat state
    def attach(self) -> None:
        """Build the flat master / moment / shadow buffers (needs the gradient layout: run one DenoiserTrainStep first) and
        re-home every parameter into the master buffer.  Called by the first `step()`; call it earlier to capture a step whose
        GEMMs already read the shadow."""
        if self.pflat is not None:
            return
        gs = self.stepper.grad_sync
        if gs is None or gs.layout is None:
            raise ops._lib.PtError("FusedClipAdamW: run one DenoiserTrainStep first (it defines the flat gradient layout)")
        dev = gs.flat.device
        self.pflat = torch.zeros(gs.total, dtype=torch.float32, device=dev)
        self.wflat = torch.zeros(gs.total, dtype=torch.bfloat16, device=dev)
        cache = self.stepper.cache
        for params, off, n, kind in gs.groups:
            sl = self.pflat[off:off + n]
            if kind == "conv":
                (p,) = params
                Co, Ci, k = p.shape
                packed = sl.view(Co, k, Ci)
                packed.copy_(p.data.permute(0, 2, 1))
                p.data = packed.permute(0, 2, 1)           # logical [Co, Ci, k]; storage tap-major = the GEMM layout
                cache.add_static("conv", params, self.wflat[off:off + n].view(Co, k * Ci), self._refresher(off, n))
                continue
            o = off
            for p in params:
                kk = p.numel()
                self.pflat[o:o + kk].copy_(p.data.reshape(-1))
                p.data = self.pflat[o:o + kk].view(p.shape)       # the module tree now reads / checkpoints the master buffer
                o += kk
            assert o == off + n
            p0 = params[0]
            if p0.dim() >= 2:
                rows = sum(p.shape[0] for p in params)
                cache.add_static("lin", params, self.wflat[off:off + n].view(rows, n // rows), self._refresher(off, n))
            elif p0.dim() == 1 and len(params) > 1:
                cache.add_static("bias", params, self.pflat[off:off + n], None)
        ops.cast_bf16(self.pflat, self.wflat)
        self.m = torch.zeros_like(self.pflat)
        self.v = torch.zeros_like(self.pflat)
        self.partials = torch.zeros(NPARTS, dtype=torch.float32, device=dev)
        host = [0.0] * NSTATE
        host[LR], host[BETA1], host[BETA2], host[EPS], host[WD], host[MAX_NORM], host[GSCALE] = (
            self.lr, self.betas[0], self.betas[1], self.eps, self.wd, self.max_norm if self.max_norm else 0.0, 1.0)
        self.state = torch.tensor(host, dtype=torch.float32, device=dev)       # slot STEP = int32 0
        if gs.comm is not None:
            gs.cast_back = False          # the AdamW kernel reads the averaged bf16 gradients directly

    def _refresher(self, off: int, n: int):
        def refresh():      # a parameter of this group was written from outside (load_state_dict, p.data.copy_)
            ops.cast_bf16(self.pflat[off:off + n], self.wflat[off:off + n])
        return refresh

    def sync_shadow(self) -> None:
        """Re-derive the whole bf16 shadow from the fp32 masters."""
        ops.cast_bf16(self.pflat, self.wflat)

    # ------------------------------------------------------------------------------------------ hyper-parameters (device side)
    def set_lr(self, lr: float) -> None:
        """What an LR scheduler calls (train.py:60-65,119): a device write, valid between replays of a captured step."""
        self.lr = lr
        if self.state is not None:
            self.state[LR:LR + 1].fill_(lr)

    @property
    def step_count(self) -> int:
        return 0 if self.state is None else int(self.state[STEP:STEP + 1].view(torch.int32).item())

    # ------------------------------------------------------------------------------------------ step
    @torch.no_grad()
    def step(self) -> torch.Tensor:
        """Clip to max_norm (global L2 norm over all gradients, like clip_grad_norm_) and apply AdamW in place.
        Returns the squared gradient norm as a 0-d device tensor (no synchronisation).  A no-op on micro-steps that do not
        synchronise gradients (accelerate's optimizer wrapper under `accumulate`, train.py:80,118)."""
        self.attach()
        if not getattr(self.stepper, "sync_gradients", True):
            return self.gnorm_sq
        gs = self.stepper.grad_sync
        g_bf16 = gs.comm is not None and not gs.cast_back
        # global norm, deterministically (identical gradients -> bit-identical clip factor on every rank: replicas never drift apart)
        ops.call("sumsq_partials", ops._p(gs.comm if g_bf16 else gs.flat), 1 if g_bf16 else 0, gs.total, ops._p(self.partials), NPARTS, ops._stream())
        ops.call("adamw_prepare_det", ops._p(self.state), ops._p(self.partials), NPARTS, ops._stream())
        ops.call("adamw_step_dev", ops._p(self.pflat), ops._p(gs.comm if g_bf16 else gs.flat), 1 if g_bf16 else 0, ops._p(self.m),
                 ops._p(self.v), ops._p(self.wflat), gs.total, ops._p(self.state), ops._stream())
        return self.gnorm_sq

    @property
    def gnorm_sq(self) -> torch.Tensor:
        """Squared global gradient norm of the last step (0-d device tensor, no synchronisation)."""
        return self.state[GNORM_SQ]

    # ------------------------------------------------------------------------------------------ checkpoints (train.py:141-143)
    def state_dict(self) -> Dict[str, torch.Tensor]:
        self.attach()
        return {"state": self.state.clone(), "m": self.m.clone(), "v": self.v.clone(),
                "layout": [(off, n, kind) for _, off, n, kind in self.stepper.grad_sync.groups]}

    def load_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        self.attach()
        want = [(off, n, kind) for _, off, n, kind in self.stepper.grad_sync.groups]
        if list(map(tuple, sd["layout"])) != want:
            raise ops._lib.PtError("FusedClipAdamW.load_state_dict: the checkpoint was written for a different flat layout")
        self.state.copy_(sd["state"])
        self.m.copy_(sd["m"])
        self.v.copy_(sd["v"])
        self.lr = float(self.state[LR].item())
        self.sync_shadow()
