"""B200 drop-ins for the reference's tts/ldm/unet_blocks.py: factories (:10-128), UpBlock1D (:131-202),
DownBlock1D (:205-281), CrossAttnDownBlock1D (:284-408), CrossAttnUpBlock1D (:411-529),
UNetMidBlock1DCrossAttn (:532-620).  Same ctor args / forward signatures / state_dict keys."""
from __future__ import annotations

import torch
from torch import nn

from .. import engine as E
from .resnet import Downsample1D, ResnetBlock1D, Upsample1D, time_projection
from .transformer_1d import Transformer1DModel


def get_down_block(down_block_type, num_layers, in_channels, out_channels, temb_channels, add_downsample, resnet_eps,
                   resnet_act_fn, attn_num_head_channels, resnet_groups=None, cross_attention_dim=None,
                   downsample_padding=None, use_linear_projection=False, only_cross_attention=False, upcast_attention=False,
                   resnet_time_scale_shift="default", resnet_skip_time_act=False, resnet_out_scale_factor=1.0,
                   cross_attention_norm=None):
    down_block_type = down_block_type[7:] if down_block_type.startswith("UNetRes") else down_block_type
    if down_block_type == "CrossAttnDownBlock1D":
        if cross_attention_dim is None:
            raise ValueError("cross_attention_dim must be specified for CrossAttnDownBlock1D")
        return CrossAttnDownBlock1D(num_layers=num_layers, in_channels=in_channels, out_channels=out_channels,
                                    temb_channels=temb_channels, add_downsample=add_downsample, resnet_eps=resnet_eps,
                                    resnet_act_fn=resnet_act_fn, resnet_groups=resnet_groups,
                                    downsample_padding=downsample_padding, cross_attention_dim=cross_attention_dim,
                                    attn_num_head_channels=attn_num_head_channels,
                                    use_linear_projection=use_linear_projection, only_cross_attention=only_cross_attention,
                                    upcast_attention=upcast_attention, resnet_time_scale_shift=resnet_time_scale_shift)
    elif down_block_type == "DownBlock1D":
        return DownBlock1D(num_layers=num_layers, in_channels=in_channels, out_channels=out_channels,
                           temb_channels=temb_channels, add_downsample=add_downsample, resnet_eps=resnet_eps,
                           resnet_act_fn=resnet_act_fn, resnet_groups=resnet_groups, downsample_padding=downsample_padding,
                           resnet_time_scale_shift=resnet_time_scale_shift)
    raise ValueError(f"{down_block_type} does not exist.")


def get_up_block(up_block_type, num_layers, in_channels, out_channels, prev_output_channel, temb_channels, add_upsample,
                 resnet_eps, resnet_act_fn, attn_num_head_channels, resnet_groups=None, cross_attention_dim=None,
                 use_linear_projection=False, only_cross_attention=False, upcast_attention=False,
                 resnet_time_scale_shift="default", resnet_skip_time_act=False, resnet_out_scale_factor=1.0,
                 cross_attention_norm=None):
    up_block_type = up_block_type[7:] if up_block_type.startswith("UNetRes") else up_block_type
    if up_block_type == "CrossAttnUpBlock1D":
        if cross_attention_dim is None:
            raise ValueError("cross_attention_dim must be specified for CrossAttnUpBlock1D")
        return CrossAttnUpBlock1D(num_layers=num_layers, in_channels=in_channels, out_channels=out_channels,
                                  prev_output_channel=prev_output_channel, temb_channels=temb_channels,
                                  add_upsample=add_upsample, resnet_eps=resnet_eps, resnet_act_fn=resnet_act_fn,
                                  resnet_groups=resnet_groups, cross_attention_dim=cross_attention_dim,
                                  attn_num_head_channels=attn_num_head_channels, use_linear_projection=use_linear_projection,
                                  only_cross_attention=only_cross_attention, upcast_attention=upcast_attention,
                                  resnet_time_scale_shift=resnet_time_scale_shift)
    elif up_block_type == "UpBlock1D":
        return UpBlock1D(num_layers=num_layers, in_channels=in_channels, out_channels=out_channels,
                         prev_output_channel=prev_output_channel, temb_channels=temb_channels, add_upsample=add_upsample,
                         resnet_eps=resnet_eps, resnet_act_fn=resnet_act_fn, resnet_groups=resnet_groups,
                         resnet_time_scale_shift=resnet_time_scale_shift)
    raise ValueError(f"{up_block_type} does not exist.")


def _resnet(cin, cout, temb, eps, groups, dropout, tss, act, osf, pre_norm):
    return ResnetBlock1D(in_channels=cin, out_channels=cout, temb_channels=temb, eps=eps, groups=groups, dropout=dropout,
                         time_embedding_norm=tss, non_linearity=act, output_scale_factor=osf, pre_norm=pre_norm)


def _xf(heads, channels, cross, groups, ulp=False, oca=False, uca=False):
    return Transformer1DModel(heads, channels // heads, in_channels=channels, num_layers=1, cross_attention_dim=cross,
                              norm_num_groups=groups, use_linear_projection=ulp, only_cross_attention=oca, upcast_attention=uca)


class _BlockBase(nn.Module):
    """Shared plumbing for the public `forward`s: build the time shifts for this block's resnets when called standalone."""

    def _shifts(self, tape, temb_var):
        return time_projection(tape, temb_var, list(self.resnets))

    def _run(self, body, tensors, kinds):
        return E.run_module(self, body, tensors, kinds)


class UpBlock1D(_BlockBase):
    def __init__(self, in_channels: int, prev_output_channel: int, out_channels: int, temb_channels: int, dropout: float = 0.0,
                 num_layers: int = 1, resnet_eps: float = 1e-6, resnet_time_scale_shift: str = "default",
                 resnet_act_fn: str = "swish", resnet_groups: int = 32, resnet_pre_norm: bool = True,
                 output_scale_factor=1.0, add_upsample=True):
        super().__init__()
        resnets = []
        for i in range(num_layers):
            res_skip_channels = in_channels if (i == num_layers - 1) else out_channels
            resnet_in_channels = prev_output_channel if i == 0 else out_channels
            resnets.append(_resnet(resnet_in_channels + res_skip_channels, out_channels, temb_channels, resnet_eps, resnet_groups,
                                   dropout, resnet_time_scale_shift, resnet_act_fn, output_scale_factor, resnet_pre_norm))
        self.resnets = nn.ModuleList(resnets)
        self.upsamplers = nn.ModuleList([Upsample1D(out_channels, use_conv=True, out_channels=out_channels)]) if add_upsample else None
        self.gradient_checkpointing = False

    def _fwd(self, tape, h, skips, shifts, upsample_size=None):
        skips = list(skips)
        for resnet, ts in zip(self.resnets, shifts):
            h = E.concat_channels(tape, h, skips.pop())
            h = resnet._fwd(tape, h, ts)
        if self.upsamplers is not None:
            for up in self.upsamplers:
                h = up._fwd(tape, h, upsample_size)
        return h

    def forward(self, hidden_states, res_hidden_states_tuple, temb=None, upsample_size=None):
        n = len(res_hidden_states_tuple)

        def body(tape, h, t, *skips):
            return self._fwd(tape, h, skips, self._shifts(tape, t), upsample_size)
        return self._run(body, [hidden_states, temb, *res_hidden_states_tuple], ["ncl", "f32"] + ["ncl"] * n)


class DownBlock1D(_BlockBase):
    def __init__(self, in_channels: int, out_channels: int, temb_channels: int, dropout: float = 0.0, num_layers: int = 1,
                 resnet_eps: float = 1e-6, resnet_time_scale_shift: str = "default", resnet_act_fn: str = "swish",
                 resnet_groups: int = 32, resnet_pre_norm: bool = True, output_scale_factor=1.0, add_downsample=True,
                 downsample_padding=1):
        super().__init__()
        resnets = []
        for i in range(num_layers):
            in_channels = in_channels if i == 0 else out_channels
            resnets.append(_resnet(in_channels, out_channels, temb_channels, resnet_eps, resnet_groups, dropout,
                                   resnet_time_scale_shift, resnet_act_fn, output_scale_factor, resnet_pre_norm))
        self.resnets = nn.ModuleList(resnets)
        self.downsamplers = nn.ModuleList([Downsample1D(out_channels, use_conv=True, out_channels=out_channels,
                                                        padding=downsample_padding, name="op")]) if add_downsample else None
        self.gradient_checkpointing = False

    def _fwd(self, tape, h, shifts, enc=None):
        outs = []
        for resnet, ts in zip(self.resnets, shifts):
            h = resnet._fwd(tape, h, ts)
            outs.append(h)
        if self.downsamplers is not None:
            for d in self.downsamplers:
                h = d._fwd(tape, h)
            outs.append(h)
        return h, outs

    def forward(self, hidden_states, temb=None):
        """-> (hidden_states, output_states) as unet_blocks.py:257-281: every resnet output (and the down-sampled state) is a skip."""
        def body(tape, h, t):
            hh, outs = self._fwd(tape, h, self._shifts(tape, t))
            return [hh] + list(outs)
        res = E.run_module(self, body, [hidden_states, temb], ["ncl", "f32"], multi=True)
        return res[0], tuple(res[1:])


class CrossAttnDownBlock1D(_BlockBase):
    def __init__(self, in_channels: int, out_channels: int, temb_channels: int, dropout: float = 0.0, num_layers: int = 1,
                 resnet_eps: float = 1e-6, resnet_time_scale_shift: str = "default", resnet_act_fn: str = "swish",
                 resnet_groups: int = 32, resnet_pre_norm: bool = True, attn_num_head_channels=1, cross_attention_dim=1280,
                 output_scale_factor=1.0, downsample_padding=1, add_downsample=True, has_cross_attention=True,
                 use_linear_projection=False, only_cross_attention=False, upcast_attention=False):
        super().__init__()
        self.has_cross_attention = has_cross_attention
        self.attn_num_head_channels = attn_num_head_channels
        resnets, attentions = [], []
        for i in range(num_layers):
            in_channels = in_channels if i == 0 else out_channels
            resnets.append(_resnet(in_channels, out_channels, temb_channels, resnet_eps, resnet_groups, dropout,
                                   resnet_time_scale_shift, resnet_act_fn, output_scale_factor, resnet_pre_norm))
            attentions.append(_xf(attn_num_head_channels, out_channels, cross_attention_dim, resnet_groups,
                                  use_linear_projection, only_cross_attention, upcast_attention))
        self.attentions = nn.ModuleList(attentions)
        self.resnets = nn.ModuleList(resnets)
        self.downsamplers = nn.ModuleList([Downsample1D(out_channels, use_conv=True, out_channels=out_channels,
                                                        padding=downsample_padding, name="op")]) if add_downsample else None
        self.gradient_checkpointing = False

    def _fwd(self, tape, h, shifts, enc):
        outs = []
        for resnet, attn, ts in zip(self.resnets, self.attentions, shifts):
            h = resnet._fwd(tape, h, ts)
            h = attn._fwd(tape, h, enc)
            outs.append(h)
        if self.downsamplers is not None:
            for d in self.downsamplers:
                h = d._fwd(tape, h)
            outs.append(h)
        return h, outs

    def forward(self, hidden_states, temb=None, encoder_hidden_states=None, attention_mask=None, cross_attention_kwargs=None):
        """-> (hidden_states, output_states) as unet_blocks.py:359-408 (attention_mask accepted and unused, as there)."""
        def body(tape, h, t, e):
            hh, outs = self._fwd(tape, h, self._shifts(tape, t), e)
            return [hh] + list(outs)
        res = E.run_module(self, body, [hidden_states, temb, encoder_hidden_states], ["ncl", "f32", "blc"], multi=True)
        return res[0], tuple(res[1:])


class CrossAttnUpBlock1D(_BlockBase):
    def __init__(self, in_channels: int, out_channels: int, prev_output_channel: int, temb_channels: int, dropout: float = 0.0,
                 num_layers: int = 1, resnet_eps: float = 1e-6, resnet_time_scale_shift: str = "default",
                 resnet_act_fn: str = "swish", resnet_groups: int = 32, resnet_pre_norm: bool = True,
                 attn_num_head_channels=1, cross_attention_dim=1280, output_scale_factor=1.0, add_upsample=True,
                 has_cross_attention=True, use_linear_projection=False, only_cross_attention=False, upcast_attention=False):
        super().__init__()
        self.has_cross_attention = has_cross_attention
        self.attn_num_head_channels = attn_num_head_channels
        resnets, attentions = [], []
        for i in range(num_layers):
            res_skip_channels = in_channels if (i == num_layers - 1) else out_channels
            resnet_in_channels = prev_output_channel if i == 0 else out_channels
            resnets.append(_resnet(resnet_in_channels + res_skip_channels, out_channels, temb_channels, resnet_eps, resnet_groups,
                                   dropout, resnet_time_scale_shift, resnet_act_fn, output_scale_factor, resnet_pre_norm))
            attentions.append(_xf(attn_num_head_channels, out_channels, cross_attention_dim, resnet_groups,
                                  use_linear_projection, only_cross_attention, upcast_attention))
        self.attentions = nn.ModuleList(attentions)
        self.resnets = nn.ModuleList(resnets)
        self.upsamplers = nn.ModuleList([Upsample1D(out_channels, use_conv=True, out_channels=out_channels)]) if add_upsample else None
        self.gradient_checkpointing = False

    def _fwd(self, tape, h, skips, shifts, enc, upsample_size=None):
        skips = list(skips)
        for resnet, attn, ts in zip(self.resnets, self.attentions, shifts):
            h = E.concat_channels(tape, h, skips.pop())
            h = resnet._fwd(tape, h, ts)
            h = attn._fwd(tape, h, enc)
        if self.upsamplers is not None:
            for up in self.upsamplers:
                h = up._fwd(tape, h)       # the reference ignores upsample_size here (unet_blocks.py:525-527)
        return h

    def forward(self, hidden_states, res_hidden_states_tuple, temb=None, encoder_hidden_states=None, cross_attention_kwargs=None,
                upsample_size=None, attention_mask=None):
        n = len(res_hidden_states_tuple)

        def body(tape, h, t, e, *skips):
            return self._fwd(tape, h, skips, self._shifts(tape, t), e, upsample_size)
        return self._run(body, [hidden_states, temb, encoder_hidden_states, *res_hidden_states_tuple], ["ncl", "f32", "blc"] + ["ncl"] * n)


class UNetMidBlock1DCrossAttn(_BlockBase):
    def __init__(self, in_channels: int, temb_channels: int, dropout: float = 0.0, num_layers: int = 1, resnet_eps: float = 1e-6,
                 resnet_time_scale_shift: str = "default", resnet_act_fn: str = "swish", resnet_groups: int = 32,
                 resnet_pre_norm: bool = True, attn_num_head_channels=1, output_scale_factor=1.0, cross_attention_dim=1280,
                 use_linear_projection=False, upcast_attention=False):
        super().__init__()
        self.has_cross_attention = True
        self.attn_num_head_channels = attn_num_head_channels
        resnet_groups = resnet_groups if resnet_groups is not None else min(in_channels // 4, 32)
        resnets = [_resnet(in_channels, in_channels, temb_channels, resnet_eps, resnet_groups, dropout,
                           resnet_time_scale_shift, resnet_act_fn, output_scale_factor, resnet_pre_norm)]
        attentions = []
        for _ in range(num_layers):
            attentions.append(_xf(attn_num_head_channels, in_channels, cross_attention_dim, resnet_groups,
                                  use_linear_projection, False, upcast_attention))
            resnets.append(_resnet(in_channels, in_channels, temb_channels, resnet_eps, resnet_groups, dropout,
                                   resnet_time_scale_shift, resnet_act_fn, output_scale_factor, resnet_pre_norm))
        self.attentions = nn.ModuleList(attentions)
        self.resnets = nn.ModuleList(resnets)

    def _fwd(self, tape, h, shifts, enc):
        h = self.resnets[0]._fwd(tape, h, shifts[0])
        for attn, resnet, ts in zip(self.attentions, self.resnets[1:], shifts[1:]):
            h = attn._fwd(tape, h, enc)
            h = resnet._fwd(tape, h, ts)
        return h

    def forward(self, hidden_states, temb=None, encoder_hidden_states=None, attention_mask=None, cross_attention_kwargs=None):
        def body(tape, h, t, e):
            return self._fwd(tape, h, self._shifts(tape, t), e)
        return self._run(body, [hidden_states, temb, encoder_hidden_states], ["ncl", "f32", "blc"])
