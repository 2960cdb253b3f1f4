"""B200 drop-ins for the reference's tts/ldm/resnet.py: ResnetBlock1D (:99-283), Upsample1D (:11-49),
Downsample1D (:52-96).  Same constructor arguments, forward signatures and state_dict keys; the children
(`nn.Conv1d`, `nn.GroupNorm`, `nn.Linear`) only own parameters -- arithmetic runs in libpt_b200.so."""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from .. import engine as E
from .. import ops


class Upsample1D(nn.Module):
    """nearest x2 (or to `output_size`) then Conv1d k3 p1 (resnet.py:36-49)."""

    def __init__(self, channels, use_conv=False, use_conv_transpose=False, out_channels=None, name="conv"):
        super().__init__()
        assert not use_conv_transpose, "conv-transpose up-sampling is not on the reference path"
        self.channels = channels
        self.out_channels = out_channels or channels
        self.use_conv = use_conv
        self.use_conv_transpose = use_conv_transpose
        self.name = name
        self.conv = nn.Conv1d(self.channels, self.out_channels, 3, padding=1) if use_conv else None

    def _fwd(self, tape, x: E.Var, output_size=None) -> E.Var:
        L = x.data.shape[1]
        if output_size is not None:
            size = output_size[-1] if isinstance(output_size, (tuple, list, torch.Size)) else int(output_size)
            if size != 2 * L:
                raise ops._lib.PtError(f"Upsample1D: output_size={size} != 2*L={2 * L} is not supported by the B200 path")
        h = E.upsample2(tape, x)
        if self.use_conv:
            h = E.conv3(tape, h, self.conv.weight, self.conv.bias)
        return h

    def forward(self, x, output_size=None):
        assert x.shape[1] == self.channels
        return E.run_module(self, lambda tape, v: self._fwd(tape, v, output_size), [x], ["ncl"])


class Downsample1D(nn.Module):
    """Conv1d k3 stride 2 padding 1 (resnet.py:87-96; the blocks build it with use_conv=True, name='op')."""

    def __init__(self, channels, use_conv=False, out_channels=None, padding=1, name="conv"):
        super().__init__()
        assert use_conv and padding == 1, "only the strided-conv down-sampler of the reference path is implemented"
        self.channels = channels
        self.out_channels = out_channels or channels
        self.use_conv = use_conv
        self.padding = padding
        self.name = name
        conv = nn.Conv1d(self.channels, self.out_channels, 3, stride=2, padding=padding)
        if name == "conv":
            self.Conv1d_0 = conv
        self.conv = conv

    def _fwd(self, tape, x: E.Var) -> E.Var:
        return E.conv3(tape, x, self.conv.weight, self.conv.bias, stride=2)

    def forward(self, hidden_states):
        assert hidden_states.shape[1] == self.channels
        return E.run_module(self, lambda tape, v: self._fwd(tape, v), [hidden_states], ["ncl"])


class ResnetBlock1D(nn.Module):
    """h = conv1(silu(GN1 x)) + Linear(silu(temb))[:, :, None]; h = conv2(silu(GN2 h)); out = (shortcut(x) | x) + h
    (resnet.py:231-283, time_embedding_norm='default', dropout 0, output_scale_factor 1)."""

    def __init__(self, *, in_channels, out_channels=None, conv_shortcut=False, dropout=0.0, temb_channels=512, groups=32,
                 groups_out=None, pre_norm=True, eps=1e-6, non_linearity="swish", skip_time_act=False,
                 time_embedding_norm="default", kernel=None, output_scale_factor=1.0, use_in_shortcut=None, up=False,
                 down=False, conv_shortcut_bias: bool = True, conv_1d_out_channels: Optional[int] = None):
        super().__init__()
        assert time_embedding_norm == "default" and not up and not down and dropout == 0.0 and output_scale_factor == 1.0
        assert non_linearity in ("swish", "silu") and not skip_time_act
        self.in_channels = in_channels
        out_channels = in_channels if out_channels is None else out_channels
        self.out_channels = out_channels
        self.up, self.down = up, down
        self.output_scale_factor = output_scale_factor
        self.time_embedding_norm = time_embedding_norm
        self.skip_time_act = skip_time_act
        self.groups = groups
        self.groups_out = groups if groups_out is None else groups_out
        self.eps = eps
        self.norm1 = nn.GroupNorm(num_groups=groups, num_channels=in_channels, eps=eps, affine=True)
        self.conv1 = nn.Conv1d(in_channels, out_channels, kernel_size=3, stride=1, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, out_channels) if temb_channels is not None else None
        self.norm2 = nn.GroupNorm(num_groups=self.groups_out, num_channels=out_channels, eps=eps, affine=True)
        self.dropout = nn.Dropout(dropout)
        conv_1d_out_channels = conv_1d_out_channels or out_channels
        self.conv2 = nn.Conv1d(out_channels, conv_1d_out_channels, kernel_size=3, stride=1, padding=1)
        self.upsample = self.downsample = None
        self.use_in_shortcut = self.in_channels != self.out_channels if use_in_shortcut is None else use_in_shortcut
        self.conv_shortcut = nn.Conv1d(in_channels, out_channels, kernel_size=1, stride=1, padding=0) if self.use_in_shortcut else None

    def _fwd(self, tape, x: E.Var, tshift: Optional[E.TimeShift]) -> E.Var:
        h = E.groupnorm(tape, x, self.norm1.weight, self.norm1.bias, self.eps, True, self.groups)
        h = E.conv3(tape, h, self.conv1.weight, self.conv1.bias, tshift=tshift)
        h = E.groupnorm(tape, h, self.norm2.weight, self.norm2.bias, self.eps, True, self.groups_out)
        sc = x
        if self.conv_shortcut is not None:
            sc = E.linear(tape, x, [self.conv_shortcut.weight], [self.conv_shortcut.bias])
        return E.conv3(tape, h, self.conv2.weight, self.conv2.bias, residual=sc)

    def forward(self, input_tensor, temb):
        def body(tape, x, t):
            ts = None
            if temb is not None and self.time_emb_proj is not None:
                ts = time_projection(tape, t, [self])[0]
            return self._fwd(tape, x, ts)
        if temb is None:
            return E.run_module(self, lambda tape, x: self._fwd(tape, x, None), [input_tensor], ["ncl"])
        return E.run_module(self, body, [input_tensor, temb], ["ncl", "f32"])


def time_projection(tape, emb: E.Var, resnets):
    """One batched GEMM for the time_emb_proj of every resnet in `resnets` (resnet.py:175,255-261):
    proj[B, sum Co] = Linear_cat(silu(emb)); each resnet consumes a column slice as a per-(b, c) shift in conv1's
    epilogue and writes its slice of the gradient."""
    se = E.silu_f32(tape, emb)
    rs = [r for r in resnets if r.time_emb_proj is not None]
    proj = E.linear(tape, se, [r.time_emb_proj.weight for r in rs], [r.time_emb_proj.bias for r in rs], out_f32=True)
    dproj = None
    if tape.recording:
        dproj = torch.zeros_like(proj.data)
        proj.grad, proj.owned = dproj, True
    out, off = [], 0
    for r in resnets:
        if r.time_emb_proj is None:
            out.append(None)
            continue
        out.append(E.TimeShift(proj.data, dproj, off))
        off += r.out_channels
    return out
