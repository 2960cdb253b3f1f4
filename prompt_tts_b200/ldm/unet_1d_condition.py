"""B200 drop-in for the reference's tts/ldm/unet_1d_condition.py: Unet1DConditionModel (:37-739) and
UNet1DConditionOutput (:28-35).  conv_in -> time embedding -> down blocks -> mid -> up blocks (skip concat) ->
GroupNorm + SiLU + conv_out, executed as one tape over channels-last bf16 activations."""
from __future__ import annotations

import os
from typing import Any, Dict, Optional, Tuple, Union

import torch
from torch import nn

from .. import engine as E
from .. import ops
from ..utils import BaseOutput, Config
from .attention import Timesteps, TimestepEmbedding
from .resnet import time_projection
from .unet_blocks import UNetMidBlock1DCrossAttn, get_down_block, get_up_block


class UNet1DConditionOutput(BaseOutput):
    """`.sample`: [B, out_channels, L] (unet_1d_condition.py:28-35)."""
    sample: torch.FloatTensor


# measurement switch: the direct (CUDA-core) conv_in / conv_out kernels instead of the implicit-GEMM path
_DIRECT_CONV_IO = os.environ.get("PT_DIRECT_CONV_IO", "0") == "1"


class Unet1DConditionModel(nn.Module):
    _supports_gradient_checkpointing = True

    def __init__(self, sample_size: Optional[int] = None, in_channels: int = 4, out_channels: int = 4,
                 center_input_sample: bool = False, flip_sin_to_cos: bool = True, freq_shift: int = 0,
                 down_block_types: Tuple[str] = ("CrossAttnDownBlock1D", "CrossAttnDownBlock1D", "CrossAttnDownBlock1D", "DownBlock1D"),
                 mid_block_type: Optional[str] = "UNetMidBlock1DCrossAttn",
                 up_block_types: Tuple[str] = ("UpBlock1D", "CrossAttnUpBlock1D", "CrossAttnUpBlock1D", "CrossAttnUpBlock1D"),
                 only_cross_attention: Union[bool, Tuple[bool]] = False, block_out_channels: Tuple[int] = (320, 640, 1280, 1280),
                 layers_per_block: Union[int, Tuple[int]] = 2, downsample_padding: int = 1, mid_block_scale_factor: float = 1,
                 act_fn: str = "silu", norm_num_groups: Optional[int] = 32, norm_eps: float = 1e-5,
                 cross_attention_dim: Union[int, Tuple[int]] = 1280, encoder_hid_dim: Optional[int] = None,
                 attention_head_dim: Union[int, Tuple[int]] = 8, use_linear_projection: bool = False,
                 class_embed_type: Optional[str] = None, num_class_embeds: Optional[int] = None, upcast_attention: bool = False,
                 resnet_time_scale_shift: str = "default", resnet_skip_time_act: bool = False, resnet_out_scale_factor: int = 1.0,
                 time_embedding_type: str = "positional", time_embedding_act_fn: Optional[str] = None,
                 timestep_post_act: Optional[str] = None, time_cond_proj_dim: Optional[int] = None, conv_in_kernel: int = 3,
                 conv_out_kernel: int = 3, projection_class_embeddings_input_dim: Optional[int] = None,
                 class_embeddings_concat: bool = False, mid_block_only_cross_attention: Optional[bool] = None,
                 cross_attention_norm: Optional[str] = None, use_timestep_embedding: bool = True):
        super().__init__()
        cfg = dict(locals())
        cfg.pop("self"); cfg.pop("__class__", None)
        self.config = Config(cfg)
        if len(down_block_types) != len(up_block_types):
            raise ValueError("Must provide the same number of `down_block_types` as `up_block_types`.")
        if len(block_out_channels) != len(down_block_types):
            raise ValueError("Must provide the same number of `block_out_channels` as `down_block_types`.")
        unsupported = dict(center_input_sample=center_input_sample, encoder_hid_dim=encoder_hid_dim, class_embed_type=class_embed_type,
                           num_class_embeds=num_class_embeds, time_embedding_act_fn=time_embedding_act_fn,
                           time_cond_proj_dim=time_cond_proj_dim, class_embeddings_concat=class_embeddings_concat)
        bad = {k: v for k, v in unsupported.items() if v}
        if bad or time_embedding_type != "positional" or conv_in_kernel != 3 or conv_out_kernel != 3 or norm_num_groups is None:
            raise ValueError(f"Unet1DConditionModel (B200 path): options outside the reference's train path are not built: {bad}")
        n = len(down_block_types)
        self.conv_in = nn.Conv1d(in_channels, block_out_channels[0], kernel_size=3, padding=1)
        time_embed_dim = block_out_channels[0] * 4
        self.time_proj = Timesteps(block_out_channels[0], flip_sin_to_cos, freq_shift)
        self.time_embedding = TimestepEmbedding(block_out_channels[0], time_embed_dim, act_fn=act_fn, post_act_fn=timestep_post_act)
        self.encoder_hid_proj = None
        self.class_embedding = None
        self.time_embed_act = None
        self.down_blocks = nn.ModuleList([])
        self.up_blocks = nn.ModuleList([])
        if isinstance(only_cross_attention, bool):
            only_cross_attention = [only_cross_attention] * n
        if isinstance(attention_head_dim, int):
            attention_head_dim = (attention_head_dim,) * n
        if isinstance(cross_attention_dim, int):
            cross_attention_dim = (cross_attention_dim,) * n
        if isinstance(layers_per_block, int):
            layers_per_block = [layers_per_block] * n
        output_channel = block_out_channels[0]
        for i, typ in enumerate(down_block_types):
            input_channel, output_channel = output_channel, block_out_channels[i]
            self.down_blocks.append(get_down_block(
                typ, num_layers=layers_per_block[i], in_channels=input_channel, out_channels=output_channel,
                temb_channels=time_embed_dim, add_downsample=i != n - 1, resnet_eps=norm_eps, resnet_act_fn=act_fn,
                resnet_groups=norm_num_groups, cross_attention_dim=cross_attention_dim[i],
                attn_num_head_channels=attention_head_dim[i], downsample_padding=downsample_padding,
                use_linear_projection=use_linear_projection, only_cross_attention=only_cross_attention[i],
                upcast_attention=upcast_attention, resnet_time_scale_shift=resnet_time_scale_shift))
        if mid_block_type == "UNetMidBlock1DCrossAttn":
            self.mid_block = UNetMidBlock1DCrossAttn(
                in_channels=block_out_channels[-1], temb_channels=time_embed_dim, resnet_eps=norm_eps, resnet_act_fn=act_fn,
                output_scale_factor=mid_block_scale_factor, resnet_time_scale_shift=resnet_time_scale_shift,
                cross_attention_dim=cross_attention_dim[-1], attn_num_head_channels=attention_head_dim[-1],
                resnet_groups=norm_num_groups, use_linear_projection=use_linear_projection, upcast_attention=upcast_attention)
        elif mid_block_type is None:
            self.mid_block = None
        else:
            raise ValueError(f"unknown mid_block_type : {mid_block_type}")
        self.num_upsamplers = 0
        rev_ch = list(reversed(block_out_channels))
        rev_heads = list(reversed(attention_head_dim))
        rev_layers = list(reversed(layers_per_block))
        rev_cross = list(reversed(cross_attention_dim))
        rev_only = list(reversed(only_cross_attention))
        output_channel = rev_ch[0]
        for i, typ in enumerate(up_block_types):
            is_final = i == n - 1
            prev_output_channel, output_channel = output_channel, rev_ch[i]
            input_channel = rev_ch[min(i + 1, n - 1)]
            if not is_final:
                self.num_upsamplers += 1
            self.up_blocks.append(get_up_block(
                typ, num_layers=rev_layers[i] + 1, in_channels=input_channel, out_channels=output_channel,
                prev_output_channel=prev_output_channel, temb_channels=time_embed_dim, add_upsample=not is_final,
                resnet_eps=norm_eps, resnet_act_fn=act_fn, resnet_groups=norm_num_groups, cross_attention_dim=rev_cross[i],
                attn_num_head_channels=rev_heads[i], use_linear_projection=use_linear_projection,
                only_cross_attention=rev_only[i], upcast_attention=upcast_attention,
                resnet_time_scale_shift=resnet_time_scale_shift))
        self.conv_norm_out = nn.GroupNorm(num_channels=block_out_channels[0], num_groups=norm_num_groups, eps=norm_eps)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv1d(block_out_channels[0], out_channels, kernel_size=3, padding=1)
        self._norm_eps, self._groups = norm_eps, norm_num_groups

    @property
    def dtype(self):
        return self.conv_in.weight.dtype

    @property
    def device(self):
        return self.conv_in.weight.device

    # -------------------------------------------------------------------------------------------- tape forward
    def _all_resnets(self):
        rs = []
        for b in self.down_blocks:
            rs += list(b.resnets)
        if self.mid_block is not None:
            rs += list(self.mid_block.resnets)
        for b in self.up_blocks:
            rs += list(b.resnets)
        return rs

    def _cross_attentions(self):
        """Every cross-attention module (attn2 of each BasicTransformerBlock) in forward order."""
        out = []
        blocks = list(self.down_blocks) + ([self.mid_block] if self.mid_block is not None else []) + list(self.up_blocks)
        for b in blocks:
            for t1d in getattr(b, "attentions", None) or []:
                for tb in t1d.transformer_blocks:
                    if tb.attn2 is not None:
                        out.append(tb.attn2)
        return out

    def _group_cross_kv(self, tape, enc: E.Var) -> None:
        """All cross-attention layers project the SAME text encoding (transformer_1d.py:258-265 -> diffusers Attention.to_k / to_v): one
        GEMM [B*Lk, 768] x [768, sum 2C] instead of 16, one data-gradient GEMM with K = sum 2C instead of 16 accumulating passes over
        the encoder gradient, one weight-gradient launch instead of 16.  When sampling (tape.kv_cache) the projection is step-invariant
        and computed once."""
        atts = self._cross_attentions()
        if not atts or len({a.to_k.weight.shape[1] for a in atts}) != 1 or os.environ.get("PT_GROUP_KV", "1") == "0":
            return
        cached = tape.kv_cache is not None and not tape.recording
        kv_all = tape.kv_cache.get("all") if cached else None
        if kv_all is None:
            ws = []
            for a in atts:
                ws += [a.to_k.weight, a.to_v.weight]
            kv_all = E.linear(tape, enc, ws)
            if cached:
                tape.kv_cache["all"] = kv_all
        off = 0
        for a in atts:
            tape.kv_group[id(a)] = (kv_all, off, off + a.inner)
            off += 2 * a.inner

    def _fwd(self, tape, sample_ncl: torch.Tensor, t_i64: torch.Tensor, enc: E.Var, want_dx: bool = False):
        """sample_ncl fp32 [B, Cin, L] -> (fp32 [B, Cout, L], seed(grad_out) closure)."""
        B, Cin, L = sample_ncl.shape
        C0 = self.conv_in.weight.shape[0]
        up = 2 ** self.num_upsamplers
        if any(s % up != 0 for s in sample_ncl.shape[-2:]):
            # the reference forwards `upsample_size` only to UpBlock1D and fails in CrossAttnUpBlock1D
            raise ops._lib.PtError(f"sample shape {tuple(sample_ncl.shape[-2:])} must be a multiple of {up} in both trailing dims")
        # time embedding (unet_1d_condition.py:606-629): sinusoid (fp32) -> Linear -> SiLU -> Linear
        sin = E.Var(ops.cast_bf16(ops.time_sinusoid(t_i64, C0)), needs_grad=False)
        e1 = E.linear(tape, sin, [self.time_embedding.linear_1.weight], [self.time_embedding.linear_1.bias], out_f32=True)
        emb = E.linear(tape, E.silu_f32(tape, e1), [self.time_embedding.linear_2.weight], [self.time_embedding.linear_2.bias], out_f32=True)
        shifts = time_projection(tape, emb, self._all_resnets())
        self._group_cross_kv(tape, enc)
        si = 0

        def take(k):
            nonlocal si
            r = shifts[si:si + k]
            si += k
            return r

        # conv_in (:654).  With a channel count that is a multiple of 8 (16-byte rows: a valid TMA operand) it runs on the same
        # implicit-GEMM path as every other k=3 convolution -- the 8-wide contraction is zero-filled up to one 64-deep stage, which
        # wastes tensor-core work that is not missed (8 us against 220 us for the direct kernel, `profiles/r01_bw_probe_v9.txt`).
        w, bia = self.conv_in.weight, self.conv_in.bias
        if Cin % 8 == 0 and not _DIRECT_CONV_IO:
            h_in = E.conv3(tape, E.Var(ops.ncl_to_nlc(sample_ncl), needs_grad=False), w, bia)
        else:
            h0 = torch.empty(B, L, C0, dtype=E.BF16, device=sample_ncl.device)
            ops.call("conv_in_fwd", ops._p(sample_ncl), ops._p(w.detach()), ops._p(bia.detach()), ops._p(h0), B, Cin, L, C0, ops._stream())
            h_in = E.Var(h0)      # NB: a distinct name -- the closure below must not see later rebinding of `h`

            def conv_in_bwd():
                if h_in.grad is not None:
                    ops.call("conv_in_bwd", ops._p(h_in.grad), ops._p(sample_ncl), ops._p(tape.pgrad(w)), ops._p(tape.pgrad(bia)), B, Cin, L, C0, ops._stream())
            tape.record(conv_in_bwd)
        h = h_in

        skips = [h]
        for blk in self.down_blocks:
            h, outs = blk._fwd(tape, h, take(len(blk.resnets)), enc)
            skips += outs
        if self.mid_block is not None:
            h = self.mid_block._fwd(tape, h, take(len(self.mid_block.resnets)), enc)
        for blk in self.up_blocks:
            k = len(blk.resnets)
            res, skips = skips[-k:], skips[:-k]
            if getattr(blk, "has_cross_attention", False):
                h = blk._fwd(tape, h, res, take(k), enc)
            else:
                h = blk._fwd(tape, h, res, take(k))
        # GroupNorm + SiLU + conv_out (:731-734)
        hn = E.groupnorm(tape, h, self.conv_norm_out.weight, self.conv_norm_out.bias, self._norm_eps, True, self._groups)
        wo, bo = self.conv_out.weight, self.conv_out.bias
        Cout = wo.shape[0]
        if Cout % 8 == 0 and not _DIRECT_CONV_IO:
            # same implicit-GEMM path (N = 8 of a 64-wide tile); the layout change back to fp32 [B, C, L] is a separate small kernel
            y_cl = E.conv3(tape, hn, wo, bo)

            def seed_gemm(gouts):
                y_cl.grad, y_cl.owned = ops.ncl_to_nlc(gouts[0].float().contiguous()), True

            return ops.nlc_to_ncl(y_cl.data), seed_gemm
        y = torch.empty(B, Cout, L, dtype=E.F32, device=sample_ncl.device)
        ops.call("conv_out_fwd", ops._p(hn.data), ops._p(wo.detach()), ops._p(bo.detach()), ops._p(y), B, C0, L, Cout, ops._stream())

        def seed(gouts):
            g = gouts[0].float().contiguous()
            dh = torch.empty_like(hn.data)
            ops.call("conv_out_bwd", ops._p(g), ops._p(hn.data), ops._p(wo.detach()), ops._p(dh), ops._p(tape.pgrad(wo)), ops._p(tape.pgrad(bo)),
                     B, C0, L, Cout, ops._stream())
            hn.grad, hn.owned = dh, True

        return y, seed

    @staticmethod
    def _timesteps(timestep, B, device):
        """tensor [B] | 0-d tensor | int | float -> int64 [B] on device (unet_1d_condition.py:606-620)."""
        if not torch.is_tensor(timestep):
            timestep = torch.tensor([timestep], dtype=torch.int64, device=device)
        elif timestep.dim() == 0:
            timestep = timestep[None]
        return timestep.to(device=device, dtype=torch.int64).expand(B).contiguous()

    def forward(self, sample: torch.FloatTensor, timestep: Union[torch.Tensor, float, int], encoder_hidden_states: torch.Tensor,
                class_labels: Optional[torch.Tensor] = None, timestep_cond: Optional[torch.Tensor] = None,
                attention_mask: Optional[torch.Tensor] = None, cross_attention_kwargs: Optional[Dict[str, Any]] = None,
                down_block_additional_residuals: Optional[Tuple[torch.Tensor]] = None,
                mid_block_additional_residual: Optional[torch.Tensor] = None, return_dict: bool = True):
        if class_labels is not None or timestep_cond is not None or down_block_additional_residuals is not None or mid_block_additional_residual is not None:
            raise ValueError("Unet1DConditionModel (B200 path): class / ControlNet-style extra inputs are not on the reference's path")
        # attention_mask is accepted and -- exactly as in the reference -- has no effect.
        if not sample.is_cuda:
            raise ops._lib.PtError("Unet1DConditionModel: inputs must be CUDA tensors; there is no CPU fallback")
        t = self._timesteps(timestep, sample.shape[0], sample.device)
        params = [p for p in self.parameters()]

        def runner(tape, s, e):
            enc = E.Var(ops.cast_bf16(e.detach().float().contiguous()))
            y, seed = self._fwd(tape, s.detach().float().contiguous(), t, enc)
            return (y,), seed, lambda: [None, ops.cast_f32(enc.grad) if enc.grad is not None else None]

        out = E.TapeFunction.apply(runner, E.get_cache(self), 2, sample, encoder_hidden_states, *params)
        if not return_dict:
            return (out,)
        return UNet1DConditionOutput(sample=out)
