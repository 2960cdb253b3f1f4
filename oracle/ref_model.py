"""CPU restatement of the reference denoiser hot path.  TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import this file; the product (`prompt_tts_b200`) never does.

It restates, as plain functions over a reference-format ``state_dict`` (fp32 torch tensors, any
device), what these reference functions compute:

  * TTSSingleSpeaker.forward           /root/reference/tts/models.py:150-172
  * TextEncoder.forward + PE           /root/reference/tts/models.py:11-52,106-120
  * Unet1DConditionModel.forward       /root/reference/tts/ldm/unet_1d_condition.py:553-739
  * {CrossAttn,}{Down,Up}Block1D, Mid  /root/reference/tts/ldm/unet_blocks.py:179-202,257-281,359-408,482-529,603-620
  * ResnetBlock1D / Up / Downsample1D  /root/reference/tts/ldm/resnet.py:36-49,87-96,231-283
  * Transformer1DModel.forward         /root/reference/tts/ldm/transformer_1d.py:247-279
  * BasicTransformerBlock / Attention / GEGLU / Timesteps / TimestepEmbedding / DDPMScheduler:
    third-party diffusers ^0.15.1 (reference pyproject.toml:14), NOT vendored and NOT installable
    here -> published algorithm restated (see oracle/diffusers_shim).  **Parity unpinned** at
    that boundary.

Pinning: `oracle/make_golden.py` runs the UNMODIFIED reference modules (imported from
/root/reference with the shim on sys.path) and stores inputs/outputs/gradients under
`tests/golden/`; `tests/test_oracle.py` checks this restatement against those vectors.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------
# small pieces
# ----------------------------------------------------------------------------------------------
def timestep_sinusoid(t: torch.Tensor, dim: int, flip_sin_to_cos: bool = True, freq_shift: float = 0.0):
    """diffusers `Timesteps` (used at unet_1d_condition.py:209,622): [cos|sin] when flipped."""
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / (half - freq_shift))
    ang = t[:, None].float() * freqs[None, :]
    s, c = torch.sin(ang), torch.cos(ang)
    return torch.cat([c, s], -1) if flip_sin_to_cos else torch.cat([s, c], -1)


def text_positional_encoding(L: int, D: int, seq_len: int, device) -> torch.Tensor:
    """models.py:19-70.  PositionalEncodingPermute1D(seq_len) applied to [B, L, D]: the tensor is
    permuted to [B, D, L], so the *position* index runs over the D feature axis and the
    *frequency* index over L:  pe[l, d] = sin(d * w[l//2]) (l even) / cos(d * w[l//2]) (l odd),
    w[i] = 10000^(-2i / ceil_even(seq_len)).  Returns [L, D]."""
    ch = int(math.ceil(seq_len / 2) * 2)
    inv_freq = 1.0 / (10000 ** (torch.arange(0, ch, 2, device=device).float() / ch))
    pos = torch.arange(D, device=device).float()
    ang = pos[:, None] * inv_freq[None, :]                       # [D, ch/2]
    emb = torch.stack((ang.sin(), ang.cos()), -1).flatten(-2)    # [D, ch] interleaved sin,cos
    return emb[:, :L].transpose(0, 1).contiguous()               # [L, D]


def _lin(sd, p, x):
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


def _gn(sd, p, x, eps, groups=32):
    return F.group_norm(x, groups, sd[p + ".weight"], sd[p + ".bias"], eps)


def _ln(sd, p, x):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-5)


def _attn(sd, p, x, ctx, heads):
    """diffusers Attention + AttnProcessor2_0: no q/k/v bias, scale 1/sqrt(d), no mask."""
    B, Lq, _ = x.shape
    q, k, v = _lin(sd, p + ".to_q", x), _lin(sd, p + ".to_k", ctx), _lin(sd, p + ".to_v", ctx)
    d = q.shape[-1] // heads
    q = q.view(B, Lq, heads, d).transpose(1, 2)
    k = k.view(B, -1, heads, d).transpose(1, 2)
    v = v.view(B, -1, heads, d).transpose(1, 2)
    w = torch.softmax((q @ k.transpose(-1, -2)) * (d ** -0.5), dim=-1)
    o = (w @ v).transpose(1, 2).reshape(B, Lq, heads * d)
    return _lin(sd, p + ".to_out.0", o)


def basic_transformer_block(sd, p, h, enc, heads):
    """diffusers 0.15 BasicTransformerBlock: pre-LN; attn1 (self) -> attn2 (cross, only if the
    block owns one) -> GEGLU FFN; each with a residual add."""
    h = _attn(sd, p + ".attn1", _ln(sd, p + ".norm1", h), _ln(sd, p + ".norm1", h), heads) + h
    if (p + ".attn2.to_q.weight") in sd:
        h = _attn(sd, p + ".attn2", _ln(sd, p + ".norm2", h), enc, heads) + h
    a, g = _lin(sd, p + ".ff.net.0.proj", _ln(sd, p + ".norm3", h)).chunk(2, dim=-1)
    return _lin(sd, p + ".ff.net.2", a * F.gelu(g)) + h


def resnet_block(sd, p, x, emb, eps=1e-5):
    """resnet.py:231-283 with time_embedding_norm='default', output_scale_factor 1, dropout 0."""
    h = F.conv1d(F.silu(_gn(sd, p + ".norm1", x, eps)), sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], padding=1)
    h = h + _lin(sd, p + ".time_emb_proj", F.silu(emb))[:, :, None]
    h = F.conv1d(F.silu(_gn(sd, p + ".norm2", h, eps)), sd[p + ".conv2.weight"], sd[p + ".conv2.bias"], padding=1)
    if (p + ".conv_shortcut.weight") in sd:
        x = F.conv1d(x, sd[p + ".conv_shortcut.weight"], sd[p + ".conv_shortcut.bias"])
    return x + h


def transformer_1d(sd, p, x, enc, heads):
    """transformer_1d.py:247-279: GN(eps 1e-6) -> 1x1 conv -> [B,L,C] -> block -> [B,C,L] -> +res.
    proj_out exists in the state_dict but is never applied (:190 vs :275-279)."""
    h = F.conv1d(_gn(sd, p + ".norm", x, 1e-6), sd[p + ".proj_in.weight"], sd[p + ".proj_in.bias"])
    h = basic_transformer_block(sd, p + ".transformer_blocks.0", h.permute(0, 2, 1), enc, heads)
    return h.permute(0, 2, 1) + x


def _count(sd, prefix):
    n = 0
    while any(k.startswith(f"{prefix}.{n}.") for k in sd):
        n += 1
    return n


# ----------------------------------------------------------------------------------------------
# model
# ----------------------------------------------------------------------------------------------
def text_encoder(sd, cfg, ids):
    p = "text_encoder"
    x = F.embedding(ids.long(), sd[p + ".word_embedding.weight"])
    B, L, D = x.shape
    x = x + text_positional_encoding(L, D, cfg["cmu_seq_len"], x.device)[None]
    heads = D // cfg["attention_head_dim"]
    for i in range(_count(sd, p + ".transformer_blocks")):
        x = basic_transformer_block(sd, f"{p}.transformer_blocks.{i}", x, None, heads)   # mask is a no-op (§3.4)
    return x


def unet(sd, cfg, sample, timestep, enc, p="unet", heads=8):
    """unet_1d_condition.py:553-739.  The UNet's `attention_head_dim` is never forwarded by
    TTSSingleSpeaker (models.py:138-148) so it is always the class default 8, which the blocks
    interpret as the NUMBER of heads (unet_blocks.py:331-333)."""
    B = sample.shape[0]
    if not torch.is_tensor(timestep):
        timestep = torch.tensor([timestep], device=sample.device)
    elif timestep.dim() == 0:
        timestep = timestep[None]
    timestep = timestep.to(sample.device).expand(B)
    c0 = sd[p + ".conv_in.weight"].shape[0]
    emb = timestep_sinusoid(timestep, c0, True, 0.0)
    emb = _lin(sd, p + ".time_embedding.linear_2", F.silu(_lin(sd, p + ".time_embedding.linear_1", emb)))

    h = F.conv1d(sample, sd[p + ".conv_in.weight"], sd[p + ".conv_in.bias"], padding=1)
    skips = [h]
    for i in range(_count(sd, p + ".down_blocks")):
        bp = f"{p}.down_blocks.{i}"
        has_attn = _count(sd, bp + ".attentions") > 0
        for j in range(_count(sd, bp + ".resnets")):
            h = resnet_block(sd, f"{bp}.resnets.{j}", h, emb)
            if has_attn:
                h = transformer_1d(sd, f"{bp}.attentions.{j}", h, enc, heads)
            skips.append(h)
        if (bp + ".downsamplers.0.conv.weight") in sd:
            h = F.conv1d(h, sd[bp + ".downsamplers.0.conv.weight"], sd[bp + ".downsamplers.0.conv.bias"], stride=2, padding=1)
            skips.append(h)

    mp = p + ".mid_block"
    if (mp + ".resnets.0.conv1.weight") in sd:
        h = resnet_block(sd, mp + ".resnets.0", h, emb)
        for j in range(_count(sd, mp + ".attentions")):
            h = transformer_1d(sd, f"{mp}.attentions.{j}", h, enc, heads)
            h = resnet_block(sd, f"{mp}.resnets.{j + 1}", h, emb)

    for i in range(_count(sd, p + ".up_blocks")):
        bp = f"{p}.up_blocks.{i}"
        has_attn = _count(sd, bp + ".attentions") > 0
        for j in range(_count(sd, bp + ".resnets")):
            h = torch.cat([h, skips.pop()], dim=1)
            h = resnet_block(sd, f"{bp}.resnets.{j}", h, emb)
            if has_attn:
                h = transformer_1d(sd, f"{bp}.attentions.{j}", h, enc, heads)
        if (bp + ".upsamplers.0.conv.weight") in sd:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = F.conv1d(h, sd[bp + ".upsamplers.0.conv.weight"], sd[bp + ".upsamplers.0.conv.bias"], padding=1)

    h = F.silu(_gn(sd, p + ".conv_norm_out", h, 1e-5))
    return F.conv1d(h, sd[p + ".conv_out.weight"], sd[p + ".conv_out.bias"], padding=1)


def tts_forward(sd, cfg, sample, timestep, text_seq_ids, attention_mask=None):
    """TTSSingleSpeaker.forward (models.py:150-172); attention_mask accepted and unused (§3.4)."""
    enc = text_encoder(sd, cfg, text_seq_ids)
    return unet(sd, cfg, sample, timestep, enc)


# ----------------------------------------------------------------------------------------------
# training-step glue (train.py:86-120) and DDPM scheduler arithmetic (diffusers 0.15)
# ----------------------------------------------------------------------------------------------
def ddpm_alphas_cumprod(n=1000, beta_start=1e-4, beta_end=0.02):
    return torch.cumprod(1.0 - torch.linspace(beta_start, beta_end, n, dtype=torch.float32), 0)


def add_noise(x0, noise, t, acp=None):
    acp = ddpm_alphas_cumprod() if acp is None else acp
    acp = acp.to(x0.device)
    sa = (acp[t] ** 0.5).view(-1, *([1] * (x0.dim() - 1)))
    sb = ((1 - acp[t]) ** 0.5).view(-1, *([1] * (x0.dim() - 1)))
    return sa * x0 + sb * noise


def ddpm_step(eps_hat, t: int, x_t, noise, acp=None, n_infer=100, n_train=1000):
    """DDPMScheduler.step, epsilon prediction, clip_sample=True, fixed_small variance."""
    acp = ddpm_alphas_cumprod(n_train) if acp is None else acp
    prev_t = t - n_train // n_infer
    a_t = acp[t]
    a_prev = acp[prev_t] if prev_t >= 0 else torch.tensor(1.0)
    b_t, b_prev = 1 - a_t, 1 - a_prev
    cur_alpha = a_t / a_prev
    cur_beta = 1 - cur_alpha
    x0 = ((x_t - b_t ** 0.5 * eps_hat) / a_t ** 0.5).clamp(-1, 1)
    mean = (a_prev ** 0.5 * cur_beta / b_t) * x0 + (cur_alpha ** 0.5 * b_prev / b_t) * x_t
    if t > 0:
        var = torch.clamp(b_prev / b_t * cur_beta, min=1e-20)
        mean = mean + var ** 0.5 * noise
    return mean


def train_step_loss(sd, cfg, codes, noise, t, ids, mask=None):
    """train.py:96-107: x_t = add_noise(x0, eps, t); loss = mse(model(x_t, t, ids, mask), eps)."""
    pred = tts_forward(sd, cfg, add_noise(codes, noise, t), t, ids, mask)
    return F.mse_loss(pred.float(), noise.float()), pred


def codes_to_x0(codes_int):
    """dataloader.py:64,77,143,168-170: Normalize(0.5,0.5)(codes/1023) = 2*codes/1023 - 1."""
    return (codes_int.float() / 1023 - 0.5) / 0.5


def collate(codes_list, cmu_sequences, max_seq_length):
    """TTS_SingleSpkr_Collate_Fn.__call__ (tts/dataloader.py:145-188) + SingleSpeakerDataset's `code = npy / 1023` (:64,77) on
    plain arrays: returns (code fp32 [B, 8, T], cmu_sequence_id int32 [B, max_len], attention_mask int32 [B, max_len])."""
    import numpy as np
    code = torch.FloatTensor(np.array([np.asarray(c) / 1023 for c in codes_list]))
    code = (code - 0.5) / 0.5                                   # torchvision Normalize([0.5], [0.5])
    ids = torch.zeros(len(cmu_sequences), max_seq_length, dtype=torch.int64).tolist()
    mask = torch.zeros(len(cmu_sequences), max_seq_length, dtype=torch.int64).tolist()
    for i, ex in enumerate(cmu_sequences):                      # _collate_batch_helpler, dataloader.py:123-137 (pad token 0)
        k = min(len(ex), max_seq_length)
        ids[i][:k] = list(ex[:k])
        mask[i][:k] = [1] * k
    return code, torch.IntTensor(ids), torch.IntTensor(mask)


# ----------------------------------------------------------------------------------------------
# parameter table: reference-format names and shapes, built without the reference
# ----------------------------------------------------------------------------------------------
def param_shapes(cfg):
    """Names/shapes of the reference `state_dict()` for `TTSSingleSpeaker(cfg)`,
    derived from the constructors cited above.  Used to make random weights on the GPU box
    where /root/reference does not exist."""
    out = {}
    D = cfg["cross_attention_dim"]

    def lin(p, o, i, bias=True):
        out[p + ".weight"] = (o, i)
        if bias:
            out[p + ".bias"] = (o,)

    def norm(p, c):
        out[p + ".weight"] = (c,)
        out[p + ".bias"] = (c,)

    def conv(p, o, i, k):
        out[p + ".weight"] = (o, i, k)
        out[p + ".bias"] = (o,)

    def tblock(p, dim, cross):
        lin(p + ".attn1.to_q", dim, dim, False); lin(p + ".attn1.to_k", dim, dim, False)
        lin(p + ".attn1.to_v", dim, dim, False); lin(p + ".attn1.to_out.0", dim, dim)
        lin(p + ".ff.net.0.proj", 8 * dim, dim); lin(p + ".ff.net.2", dim, 4 * dim)
        if cross is not None:
            lin(p + ".attn2.to_q", dim, dim, False); lin(p + ".attn2.to_k", dim, cross, False)
            lin(p + ".attn2.to_v", dim, cross, False); lin(p + ".attn2.to_out.0", dim, dim)
            norm(p + ".norm2", dim)
        norm(p + ".norm1", dim); norm(p + ".norm3", dim)

    def resnet(p, ci, co, temb):
        norm(p + ".norm1", ci); conv(p + ".conv1", co, ci, 3); lin(p + ".time_emb_proj", co, temb)
        norm(p + ".norm2", co); conv(p + ".conv2", co, co, 3)
        if ci != co:
            conv(p + ".conv_shortcut", co, ci, 1)

    def xf(p, c):
        norm(p + ".norm", c); conv(p + ".proj_in", c, c, 1)
        tblock(p + ".transformer_blocks.0", c, D); conv(p + ".proj_out", c, c, 1)

    out["text_encoder.word_embedding.weight"] = (cfg["cmu_vocab_len"], D)
    out["text_encoder.pos_embedding.penc.inv_freq"] = (int(math.ceil(cfg["cmu_seq_len"] / 2) * 2) // 2,)
    for i in range(cfg["text_encoder_layers"]):
        tblock(f"text_encoder.transformer_blocks.{i}", D, None)

    boc = list(cfg["block_out_channels"]); n = len(boc); lpb = cfg["layers_per_block"]; temb = boc[0] * 4
    conv("unet.conv_in", boc[0], cfg["in_channels"], 3)
    lin("unet.time_embedding.linear_1", temb, boc[0]); lin("unet.time_embedding.linear_2", temb, temb)
    oc = boc[0]
    for i, typ in enumerate(cfg["down_block_types"]):
        ic, oc = oc, boc[i]
        for j in range(lpb):
            resnet(f"unet.down_blocks.{i}.resnets.{j}", ic if j == 0 else oc, oc, temb)
            if typ.startswith("CrossAttn"):
                xf(f"unet.down_blocks.{i}.attentions.{j}", oc)
        if i != n - 1:
            conv(f"unet.down_blocks.{i}.downsamplers.0.conv", oc, oc, 3)
    xf("unet.mid_block.attentions.0", boc[-1])
    resnet("unet.mid_block.resnets.0", boc[-1], boc[-1], temb); resnet("unet.mid_block.resnets.1", boc[-1], boc[-1], temb)
    rev = boc[::-1]; oc = rev[0]
    for i, typ in enumerate(cfg["up_block_types"]):
        prev, oc, ic = oc, rev[i], rev[min(i + 1, n - 1)]
        for j in range(lpb + 1):
            skip = ic if j == lpb else oc
            rin = prev if j == 0 else oc
            resnet(f"unet.up_blocks.{i}.resnets.{j}", rin + skip, oc, temb)
            if typ.startswith("CrossAttn"):
                xf(f"unet.up_blocks.{i}.attentions.{j}", oc)
        if i != n - 1:
            conv(f"unet.up_blocks.{i}.upsamplers.0.conv", oc, oc, 3)
    norm("unet.conv_norm_out", boc[0]); conv("unet.conv_out", cfg["out_channels"], boc[0], 3)
    return out


def random_state_dict(cfg, seed=0, device="cpu", scale=1.0):
    """Seeded random weights with torch-default-like fan-in scaling (NOT bit-identical to the
    reference's init; tests that need reference-initialised weights use tests/golden)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in param_shapes(cfg).items():
        if k.endswith("inv_freq"):
            ch = shp[0] * 2
            sd[k] = 1.0 / (10000 ** (torch.arange(0, ch, 2).float() / ch))
        elif len(shp) == 1 and (".norm" in k or "conv_norm_out" in k) and k.endswith(".weight"):
            sd[k] = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif len(shp) == 1:
            sd[k] = 0.05 * torch.randn(shp, generator=g)
        elif "word_embedding" in k:
            sd[k] = torch.randn(shp, generator=g)
        else:
            fan_in = 1
            for s in shp[1:]:
                fan_in *= s
            sd[k] = scale * torch.randn(shp, generator=g) / math.sqrt(fan_in)
    return {k: v.to(device) for k, v in sd.items()}
