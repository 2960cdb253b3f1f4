"""Attention + processors, diffusers 0.15 semantics (import sites: unet_1d_condition.py:20,482).

to_q/to_k/to_v: Linear WITHOUT bias; to_out = [Linear(inner, query_dim, bias=True), Dropout];
scale = dim_head ** -0.5; softmax over keys; no mask on any path the reference exercises
(SURVEY.md §3.4); non-causal.  AttnProcessor and AttnProcessor2_0 are the same mathematics.
"""
import torch
import torch.nn.functional as F
from torch import nn


class AttnProcessor:
    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None):
        b, lq, _ = hidden_states.shape
        ctx = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        q = attn.to_q(hidden_states)
        k = attn.to_k(ctx)
        v = attn.to_v(ctx)
        h = attn.heads
        d = q.shape[-1] // h
        q = q.view(b, lq, h, d).transpose(1, 2)
        k = k.view(b, -1, h, d).transpose(1, 2)
        v = v.view(b, -1, h, d).transpose(1, 2)
        if attention_mask is not None:
            # [B, 1, Lk] additive -> broadcast over heads and queries
            attention_mask = attention_mask.view(b, 1, 1, -1) if attention_mask.dim() == 3 else attention_mask
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=attention_mask, dropout_p=0.0, is_causal=False)
        o = o.transpose(1, 2).reshape(b, lq, h * d).to(q.dtype)
        o = attn.to_out[0](o)
        o = attn.to_out[1](o)
        return o


AttnProcessor2_0 = AttnProcessor
AttentionProcessor = AttnProcessor


class Attention(nn.Module):
    def __init__(self, query_dim, cross_attention_dim=None, heads=8, dim_head=64, dropout=0.0, bias=False,
                 upcast_attention=False, processor=None):
        super().__init__()
        inner = dim_head * heads
        cross_attention_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.upcast_attention = upcast_attention
        self.to_q = nn.Linear(query_dim, inner, bias=bias)
        self.to_k = nn.Linear(cross_attention_dim, inner, bias=bias)
        self.to_v = nn.Linear(cross_attention_dim, inner, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(dropout)])
        self.processor = processor if processor is not None else AttnProcessor()

    def set_processor(self, processor):
        self.processor = processor

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **kw):
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                              attention_mask=attention_mask)
