"""Diffusion sampling on the B200 path (SURVEY section 8 row N1 -- not in the reference tree, which ships no sampler).

The loop is what a user of the reference would write with diffusers 0.15 on top of `TTSSingleSpeaker`:

    scheduler = DDPMScheduler(num_train_timesteps=1000)           # train.py:32-36 (linear betas 1e-4 .. 0.02, epsilon prediction)
    scheduler.set_timesteps(100)
    text_emb = model.text_encoder(ids, mask)                      # once
    for t in scheduler.timesteps:                                 # 990, 980, ..., 0
        eps = model.unet(x, t, encoder_hidden_states=text_emb).sample
        x = scheduler.step(eps, t, x).prev_sample

Here the text encoder runs once, the cross-attention K/V projections of its output are computed once and reused by all
steps (they do not depend on x or t), one denoiser forward is captured in a CUDA graph and replayed per step, and the
scheduler step is one fused kernel (`pt_ddpm_step`), optionally in-painting the first `prompt` frames from a noised copy
of the speech prompt (the builder-defined prompt protocol of SURVEY 8d config 4).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import engine as E
from . import ops


def ddpm_alphas_cumprod(n: int = 1000, beta_start: float = 1e-4, beta_end: float = 0.02) -> torch.Tensor:
    betas = torch.linspace(beta_start, beta_end, n, dtype=torch.float32)
    return torch.cumprod(1.0 - betas, 0)


class DDPMSampler:
    def __init__(self, model, n_infer: int = 100, n_train: int = 1000, use_graph: bool = True):
        self.model = model
        self.n_infer, self.n_train = n_infer, n_train
        self.stride = n_train // n_infer
        self.acp = ddpm_alphas_cumprod(n_train)               # host copy: the step coefficients are scalars per step
        self.timesteps = [i * self.stride for i in range(n_infer)][::-1]
        self.use_graph = use_graph
        self.cache = E.get_cache(model)
        self._graph = None
        self._static: Dict[str, torch.Tensor] = {}

    # ---- one denoiser forward on static buffers (x, t) -> eps, with the text encoding and its K/V projections held fixed
    def _denoise(self) -> None:
        tape = E.Tape(self.cache, recording=False)
        tape.kv_cache = self._kv
        y, _ = self.model.unet._fwd(tape, self._static["x"], self._static["t"], self._enc)
        self._static["eps"].copy_(y)

    @torch.no_grad()
    def sample(self, ids: torch.Tensor, T: int, x_T: Optional[torch.Tensor] = None, noises: Optional[torch.Tensor] = None,
               prompt: Optional[torch.Tensor] = None, prompt_noise: Optional[torch.Tensor] = None, seed: int = 0,
               return_codes: bool = False, prompt_tokens: Optional[torch.Tensor] = None) -> torch.Tensor:
        """ids int32 [B, Lt]; returns x_0 fp32 [B, C, T] (or int64 codes in [0, 1023] with return_codes).
        x_T / noises[n_infer, B, C, T] / prompt_noise may be given for reproducibility against an oracle; otherwise they are
        drawn from a generator seeded with `seed`.  prompt fp32 [B, C, P]: clean prompt frames in-painted at every step.
        prompt_tokens float [B, P, cross_attention_dim] (optional, default off): a speech-prompt encoding appended to the text
        encoding as extra cross-attention tokens -- `Unet1DConditionModel.forward` takes `encoder_hidden_states` of any length
        (unet_1d_condition.py:557); with None the conditioning is exactly the reference's."""
        if not ids.is_cuda:
            raise ops._lib.PtError("DDPMSampler: inputs must be CUDA tensors; there is no CPU fallback")
        dev = ids.device
        B = ids.shape[0]
        Cin = self.model.unet.conv_in.weight.shape[1]
        gen = torch.Generator(device=dev).manual_seed(seed)
        x = x_T.clone().float() if x_T is not None else torch.randn(B, Cin, T, device=dev, generator=gen)
        # text encoder once; K/V of every cross-attention layer are filled in by the first denoiser call and then reused
        tape = E.Tape(self.cache, recording=False)
        self._enc = self.model.text_encoder._fwd(tape, ids.to(torch.int32).contiguous())
        if prompt_tokens is not None:
            if prompt_tokens.shape[0] != B or prompt_tokens.shape[2] != self._enc.data.shape[2]:
                raise ops._lib.PtError(f"prompt_tokens must be [B, P, {self._enc.data.shape[2]}], got {tuple(prompt_tokens.shape)}")
            self._enc = E.Var(torch.cat([self._enc.data, ops.cast_bf16(prompt_tokens.float().contiguous())], dim=1), needs_grad=False)
        self._kv: Dict[int, E.Var] = {}
        self._static = {"x": x, "t": torch.zeros(B, dtype=torch.int64, device=dev), "eps": torch.empty_like(x)}
        self._graph = None
        keep = 0
        known = None
        if prompt is not None:
            keep = prompt.shape[-1]
            known = torch.zeros_like(x)
            if prompt_noise is None:
                prompt_noise = torch.randn(self.n_infer, B, Cin, keep, device=dev, generator=gen)
        noise = torch.empty_like(x)
        for i, t in enumerate(self.timesteps):
            self._static["t"].fill_(t)
            if self.use_graph and i == 1:        # call 0 ran eagerly (fills the K/V cache, warms the weight packs); capture call 1
                torch.cuda.synchronize()
                self._graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self._graph):
                    self._denoise()
                self._graph.replay()             # capture only records; this executes step 1
            elif self._graph is not None:
                self._graph.replay()
            else:
                self._denoise()
            t_prev = t - self.stride
            acp_t = float(self.acp[t])
            acp_prev = float(self.acp[t_prev]) if t_prev >= 0 else 1.0
            if t_prev >= 0:
                if noises is not None:
                    noise.copy_(noises[i])
                else:
                    noise.normal_(generator=gen)
            if known is not None:
                # in-paint: the prompt frames of x_{t_prev} are the clean prompt noised to level t_prev (clean at the last step)
                sa, sb = (acp_prev ** 0.5, (1.0 - acp_prev) ** 0.5) if t_prev >= 0 else (1.0, 0.0)
                known[..., :keep] = sa * prompt + sb * prompt_noise[i]
            ops.call("ddpm_step", ops._p(self._static["eps"]), ops._p(x), ops._p(noise if t_prev >= 0 else None), ops._p(known), ops._p(x),
                     x.numel(), T, keep, acp_t, acp_prev, ops._stream())
        self._graph = None
        if return_codes:
            codes = torch.empty(x.shape, dtype=torch.int64, device=dev)
            ops.call("codes_affine_inv", ops._p(x), ops._p(codes), x.numel(), ops._stream())
            return codes
        return x
