"""BasicTransformerBlock / FeedForward / GEGLU, diffusers 0.15 semantics
(import sites: transformer_1d.py:11,165-178; models.py:8,95-100).

forward signature is the 0.15 one: (hidden_states, encoder_hidden_states=None, timestep=None,
attention_mask=None, cross_attention_kwargs=None, class_labels=None) -- the reference's
TextEncoder passes its mask as the 2nd positional (models.py:118), i.e. as
encoder_hidden_states, which is ignored because those blocks have no attn2.
"""
import torch.nn.functional as F
from torch import nn

from .attention_processor import Attention


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        x, gate = self.proj(x).chunk(2, dim=-1)
        return x * F.gelu(gate)  # exact erf GELU


class FeedForward(nn.Module):
    def __init__(self, dim, dim_out=None, mult=4, dropout=0.0, activation_fn="geglu", final_dropout=False):
        super().__init__()
        inner = int(dim * mult)
        dim_out = dim_out if dim_out is not None else dim
        assert activation_fn == "geglu", "only the path the reference builds is restated"
        self.net = nn.ModuleList([GEGLU(dim, inner), nn.Dropout(dropout), nn.Linear(inner, dim_out)])
        if final_dropout:
            self.net.append(nn.Dropout(dropout))

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, num_attention_heads, attention_head_dim, dropout=0.0, cross_attention_dim=None,
                 activation_fn="geglu", num_embeds_ada_norm=None, attention_bias=False, only_cross_attention=False,
                 upcast_attention=False, norm_elementwise_affine=True, norm_type="layer_norm", final_dropout=False):
        super().__init__()
        assert norm_type == "layer_norm" and num_embeds_ada_norm is None
        self.only_cross_attention = only_cross_attention
        self.attn1 = Attention(query_dim=dim, heads=num_attention_heads, dim_head=attention_head_dim,
                               dropout=dropout, bias=attention_bias,
                               cross_attention_dim=cross_attention_dim if only_cross_attention else None,
                               upcast_attention=upcast_attention)
        self.ff = FeedForward(dim, dropout=dropout, activation_fn=activation_fn, final_dropout=final_dropout)
        if cross_attention_dim is not None:
            self.attn2 = Attention(query_dim=dim, cross_attention_dim=cross_attention_dim,
                                   heads=num_attention_heads, dim_head=attention_head_dim, dropout=dropout,
                                   bias=attention_bias, upcast_attention=upcast_attention)
            self.norm2 = nn.LayerNorm(dim, elementwise_affine=norm_elementwise_affine)
        else:
            self.attn2 = None
            self.norm2 = None
        self.norm1 = nn.LayerNorm(dim, elementwise_affine=norm_elementwise_affine)
        self.norm3 = nn.LayerNorm(dim, elementwise_affine=norm_elementwise_affine)

    def forward(self, hidden_states, encoder_hidden_states=None, timestep=None, attention_mask=None,
                cross_attention_kwargs=None, class_labels=None):
        cross_attention_kwargs = cross_attention_kwargs if cross_attention_kwargs is not None else {}
        n = self.norm1(hidden_states)
        a = self.attn1(n, encoder_hidden_states=encoder_hidden_states if self.only_cross_attention else None,
                       attention_mask=attention_mask, **cross_attention_kwargs)
        hidden_states = a + hidden_states
        if self.attn2 is not None:
            n = self.norm2(hidden_states)
            a = self.attn2(n, encoder_hidden_states=encoder_hidden_states, attention_mask=attention_mask,
                           **cross_attention_kwargs)
            hidden_states = a + hidden_states
        n = self.norm3(hidden_states)
        hidden_states = self.ff(n) + hidden_states
        return hidden_states
