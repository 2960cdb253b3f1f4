"""B200 drop-in for the reference's tts/ldm/transformer_1d.py (Transformer1DModel :26-310)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch
from torch import nn

from .. import engine as E
from ..utils import BaseOutput, Config
from .attention import BasicTransformerBlock


@dataclass
class Transformer1DModelOutput(BaseOutput):
    sample: torch.FloatTensor = None

    def __post_init__(self):
        self["sample"] = self.sample


class Transformer1DModel(nn.Module):
    """r = x; h = conv1x1(GN(x, eps 1e-6)); [B,C,L]->[B,L,C]; BasicTransformerBlock(h, enc); back; out = h + r.
    `proj_out` is constructed (state_dict parity, transformer_1d.py:190) and never applied (:275-279)."""

    def __init__(self, num_attention_heads: int = 16, attention_head_dim: int = 88, in_channels: Optional[int] = None,
                 out_channels: Optional[int] = None, num_layers: int = 1, dropout: float = 0.0, norm_num_groups: int = 32,
                 cross_attention_dim: Optional[int] = None, attention_bias: bool = False, sample_size: Optional[int] = None,
                 num_vector_embeds: Optional[int] = None, patch_size: Optional[int] = None, activation_fn: str = "geglu",
                 num_embeds_ada_norm: Optional[int] = None, use_linear_projection: bool = False,
                 only_cross_attention: bool = False, upcast_attention: bool = False, norm_type: str = "layer_norm",
                 norm_elementwise_affine: bool = True):
        super().__init__()
        assert in_channels is not None and patch_size is None and not use_linear_projection
        self.config = Config(num_attention_heads=num_attention_heads, attention_head_dim=attention_head_dim,
                             in_channels=in_channels, out_channels=out_channels, num_layers=num_layers, dropout=dropout,
                             norm_num_groups=norm_num_groups, cross_attention_dim=cross_attention_dim,
                             attention_bias=attention_bias, activation_fn=activation_fn,
                             use_linear_projection=use_linear_projection, only_cross_attention=only_cross_attention,
                             upcast_attention=upcast_attention, norm_type=norm_type)
        self.use_linear_projection = use_linear_projection
        self.num_attention_heads = num_attention_heads
        self.attention_head_dim = attention_head_dim
        inner_dim = num_attention_heads * attention_head_dim
        self.is_input_continuous = True
        self.in_channels = in_channels
        self.norm_num_groups = norm_num_groups
        self.norm = nn.GroupNorm(num_groups=norm_num_groups, num_channels=in_channels, eps=1e-6, affine=True)
        self.proj_in = nn.Conv1d(in_channels, inner_dim, kernel_size=1, stride=1, padding=0)
        self.transformer_blocks = nn.ModuleList([
            BasicTransformerBlock(inner_dim, num_attention_heads, attention_head_dim, dropout=dropout,
                                  cross_attention_dim=cross_attention_dim, activation_fn=activation_fn,
                                  num_embeds_ada_norm=num_embeds_ada_norm, attention_bias=attention_bias,
                                  only_cross_attention=only_cross_attention, upcast_attention=upcast_attention,
                                  norm_type=norm_type, norm_elementwise_affine=norm_elementwise_affine)
            for _ in range(num_layers)])
        self.out_channels = in_channels if out_channels is None else out_channels
        self.proj_out = nn.Conv1d(inner_dim, in_channels, kernel_size=1, stride=1, padding=0)   # dead weight, kept for checkpoints

    def _fwd(self, tape, x: E.Var, enc) -> E.Var:
        h = E.groupnorm(tape, x, self.norm.weight, self.norm.bias, 1e-6, False, self.norm_num_groups)
        h = E.linear(tape, h, [self.proj_in.weight], [self.proj_in.bias])
        for blk in self.transformer_blocks:
            h = blk._fwd(tape, h, enc)
        return E.add(tape, h, x)

    def forward(self, hidden_states, encoder_hidden_states=None, timestep=None, class_labels=None,
                cross_attention_kwargs=None, return_dict: bool = True):
        if encoder_hidden_states is None:
            raise ValueError("Transformer1DModel on the B200 path needs encoder_hidden_states (the blocks own a cross-attention)")
        out = E.run_module(self, lambda tape, x, e: self._fwd(tape, x, e), [hidden_states, encoder_hidden_states], ["ncl", "blc"])
        if not return_dict:
            return (out,)
        return Transformer1DModelOutput(sample=out)
