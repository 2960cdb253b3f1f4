"""TEST INFRASTRUCTURE ONLY -- CPU restatement of EnCodec's SEANet encoder / decoder.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.

What it restates.  The reference reaches this arithmetic through the third-party `encodec` package (pinned `^0.1.1`,
pyproject.toml:11; absent from /root/reference and from this image):
    data_preparation/generate_code.py:13-15,48   EncodecModel.encodec_model_24khz(); set_target_bandwidth(6.0); model.encode(wav)
    decode_codec.py:8-9,16                        model.decode([(codes, None)])
encodec 0.1.1's modules (`modules/seanet.py` SEANetEncoder / SEANetDecoder / SEANetResnetBlock, `modules/conv.py`
SConv1d / SConvTranspose1d / pad1d / get_extra_padding_for_conv1d, `modules/lstm.py` SLSTM) are restated line for line by
`transformers.models.encodec.modeling_encodec` (transformers 5.5.0, installed in this image); the citations below are to that
file (`ME:<line>`).  Every function here is written with explicit loops / index arithmetic (no nn.Module, no F.conv1d in the
decisive path) so that it is an independent statement of the algorithm; `tests/test_oracle.py` pins it against transformers'
modules on seeded weights and against `tests/golden/seanet_golden.npz` (made by `oracle/make_golden.py`).

Parity status: pinned to transformers' restatement; **unpinned against encodec 0.1.1 itself** (not installable offline) and
against the pretrained 24 kHz weights (not downloadable) -- all vectors use seeded random weights.
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np

# EncodecConfig() defaults = the 24 kHz model (probed in this image): causal, reflect padding, weight norm, conv shortcut
CFG_24KHZ = dict(audio_channels=1, num_filters=32, kernel_size=7, last_kernel_size=7, residual_kernel_size=3,
                 dilation_growth_rate=2, compress=2, num_lstm_layers=2, num_residual_layers=1,
                 upsampling_ratios=(8, 5, 4, 2), use_conv_shortcut=True, use_causal_conv=True, pad_mode="reflect",
                 trim_right_ratio=1.0, hidden_size=128, sampling_rate=24000, codebook_size=1024, codebook_dim=128)

# a narrow configuration for fast tests (same topology, 4 stages, 2 LSTM layers)
CFG_TINY = dict(CFG_24KHZ, num_filters=4, hidden_size=16)


# ------------------------------------------------------------------------------------------------ parameters
def layer_plan(cfg) -> Dict[str, list]:
    """The module lists of EncodecEncoder.__init__ (ME:281-303) and EncodecDecoder.__init__ (ME:313-341) as
    (index, kind, spec) tuples; indices are the positions in `encoder.layers` / `decoder.layers` (ELUs occupy indices too)."""
    nf, ratios = cfg["num_filters"], tuple(cfg["upsampling_ratios"])
    enc, i, scale = [], 0, 1
    enc.append((i, "conv", dict(ci=cfg["audio_channels"], co=nf, k=cfg["kernel_size"], stride=1, dil=1))); i += 1
    for r in reversed(ratios):
        dim = scale * nf
        for j in range(cfg["num_residual_layers"]):
            enc.append((i, "res", dict(dim=dim, dil=cfg["dilation_growth_rate"] ** j))); i += 1
        i += 1                                                     # nn.ELU
        enc.append((i, "conv", dict(ci=dim, co=2 * dim, k=2 * r, stride=r, dil=1))); i += 1
        scale *= 2
    enc.append((i, "lstm", dict(dim=scale * nf))); i += 1
    i += 1                                                         # nn.ELU
    enc.append((i, "conv", dict(ci=scale * nf, co=cfg["hidden_size"], k=cfg["last_kernel_size"], stride=1, dil=1)))
    dec, i = [], 0
    scale = 2 ** len(ratios)
    dec.append((i, "conv", dict(ci=cfg["hidden_size"], co=scale * nf, k=cfg["kernel_size"], stride=1, dil=1))); i += 1
    dec.append((i, "lstm", dict(dim=scale * nf))); i += 1
    for r in ratios:
        dim = scale * nf
        i += 1                                                     # nn.ELU
        dec.append((i, "convtr", dict(ci=dim, co=dim // 2, k=2 * r, stride=r))); i += 1
        for j in range(cfg["num_residual_layers"]):
            dec.append((i, "res", dict(dim=dim // 2, dil=cfg["dilation_growth_rate"] ** j))); i += 1
        scale //= 2
    i += 1                                                         # nn.ELU
    dec.append((i, "conv", dict(ci=nf, co=cfg["audio_channels"], k=cfg["last_kernel_size"], stride=1, dil=1)))
    return {"encoder": enc, "decoder": dec}


def param_shapes(cfg) -> Dict[str, tuple]:
    """state_dict keys (transformers naming, torch >= 2.1 parametrised weight norm: original0 = g, original1 = v) -> shapes."""
    out = {}

    def conv(prefix, ci, co, k, transposed=False):
        out[prefix + ".conv.bias"] = (co,)
        lead = ci if transposed else co
        out[prefix + ".conv.parametrizations.weight.original0"] = (lead, 1, 1)
        out[prefix + ".conv.parametrizations.weight.original1"] = (ci, co, k) if transposed else (co, ci, k)

    for side, plan in layer_plan(cfg).items():
        for idx, kind, s in plan:
            p = f"{side}.layers.{idx}"
            if kind == "conv":
                conv(p, s["ci"], s["co"], s["k"])
            elif kind == "convtr":
                conv(p, s["ci"], s["co"], s["k"], transposed=True)
            elif kind == "res":
                hid = s["dim"] // cfg["compress"]
                conv(p + ".block.1", s["dim"], hid, cfg["residual_kernel_size"])
                conv(p + ".block.3", hid, s["dim"], 1)
                if cfg["use_conv_shortcut"]:
                    conv(p + ".shortcut", s["dim"], s["dim"], 1)
            else:
                H = s["dim"]
                for l in range(cfg["num_lstm_layers"]):
                    out[f"{p}.lstm.weight_ih_l{l}"] = (4 * H, H)
                    out[f"{p}.lstm.weight_hh_l{l}"] = (4 * H, H)
                    out[f"{p}.lstm.bias_ih_l{l}"] = (4 * H,)
                    out[f"{p}.lstm.bias_hh_l{l}"] = (4 * H,)
    return out


def make_weights(cfg, seed: int) -> Dict[str, np.ndarray]:
    """Seeded synthetic weights (no pretrained checkpoint is reachable offline).  Drawn with numpy's PCG64 in sorted-key order so
    the same dictionary can be rebuilt on any machine; scaled so activations stay O(1) through both stacks."""
    rng = np.random.Generator(np.random.PCG64(seed))
    transposed = {f"{side}.layers.{idx}" for side, plan in layer_plan(cfg).items() for idx, kind, _ in plan if kind == "convtr"}
    P = {}
    for key, shape in sorted(param_shapes(cfg).items()):
        if key.endswith("original1"):
            # taps that meet in one output: Ci * K for a convolution, Ci * (K / stride = 2) for a transposed one
            fan_in = shape[0] * 2 if key.split(".conv.")[0] in transposed else shape[1] * shape[2]
            P[key] = (rng.standard_normal(shape) / math.sqrt(fan_in)).astype(np.float32)
        elif key.endswith("original0"):
            P[key] = (1.0 + 0.25 * rng.standard_normal(shape)).astype(np.float32)       # set relative to |v| below
        elif "lstm.weight" in key:
            P[key] = (rng.uniform(-1.0, 1.0, shape) / math.sqrt(shape[1])).astype(np.float32)
        else:
            P[key] = (0.05 * rng.standard_normal(shape)).astype(np.float32)
    for key in list(P):
        if key.endswith("original0"):
            v = P[key[:-1] + "1"]
            n = np.sqrt((v.astype(np.float64) ** 2).sum(axis=(1, 2), keepdims=True))
            # g = 1.3 |v| (1 + 0.25 n): an effective weight of He-like gain so ELU stacks neither die nor blow up
            P[key] = (1.3 * n * P[key]).astype(np.float32)
    return P


# ------------------------------------------------------------------------------------------------ primitives
def fold_weight_norm(g: np.ndarray, v: np.ndarray) -> np.ndarray:
    """torch.nn.utils.parametrizations.weight_norm, dim = 0 (ME:103-108, 171-176): w = g * v / |v|, the norm over every axis but the
    first (for ConvTranspose1d the first axis is the INPUT channel)."""
    n = np.sqrt((v.astype(np.float32) ** 2).sum(axis=tuple(range(1, v.ndim)), keepdims=True, dtype=np.float32))
    return (g * (v / n)).astype(np.float32)


def elu(x: np.ndarray) -> np.ndarray:
    """nn.ELU(alpha = 1)."""
    return np.where(x > 0, x, np.expm1(np.minimum(x, 0))).astype(np.float32)


def conv_out_len(L: int, stride: int) -> int:
    """Length after EncodecConv1d.forward: the extra right padding of ME:125-133 makes it ceil(L / stride)."""
    return -(-L // stride)


def pad_amounts(k: int, stride: int, dil: int, causal: bool):
    """(left, right-before-extra) padding of EncodecConv1d.forward, ME:153-163."""
    total = (k - 1) * dil + 1 - stride
    if causal:
        return total, 0
    right = total // 2
    return total - right, right


def src_index(q: int, L: int, reflect: bool) -> int:
    """Index into the unpadded signal for padded position q (left pad removed): F.pad(..., 'reflect') mirrors WITHOUT repeating
    the edge sample; -1 = zero."""
    if q < 0:
        return -q if reflect else -1
    if q >= L:
        return 2 * (L - 1) - q if reflect else -1
    return q


def conv1d(x: np.ndarray, w: np.ndarray, b: np.ndarray, stride: int, dil: int, causal: bool, pad_mode: str) -> np.ndarray:
    """EncodecConv1d.forward, ME:150-170, by index arithmetic: x [B, Ci, L] -> [B, Co, ceil(L / stride)].
    The small-input branch of `_pad1d` (ME:143-148: length <= pad) is not restated: such inputs are rejected."""
    B, Ci, L = x.shape
    Co, _, K = w.shape
    left, right = pad_amounts(K, stride, dil, causal)
    Lo = conv_out_len(L, stride)
    reflect = pad_mode == "reflect"
    over = (Lo - 1) * stride + (K - 1) * dil - left - (L - 1)           # how far the last window reaches past the end
    if reflect and (left > L - 1 or over > L - 1):
        raise ValueError(f"input of length {L} is shorter than the reflect padding ({left}, {over})")
    y = np.zeros((B, Co, Lo), np.float32)
    t = np.arange(Lo)
    for k in range(K):
        q = t * stride + k * dil - left
        idx = np.array([src_index(int(v), L, reflect) for v in q])
        xs = np.where(idx[None, None, :] >= 0, x[:, :, np.maximum(idx, 0)], 0.0).astype(np.float32)     # [B, Ci, Lo]
        y += np.einsum("oc,bct->bot", w[:, :, k], xs, dtype=np.float32)
    return (y + b[None, :, None]).astype(np.float32)


def conv_transpose1d(x: np.ndarray, w: np.ndarray, b: np.ndarray, stride: int, causal: bool, trim_right_ratio: float = 1.0):
    """EncodecConvTranspose1d.forward, ME:183-208: full transposed convolution (length (L-1)*stride + K), then the fixed padding
    K - stride is trimmed (all of it on the right when causal).  x [B, Ci, L], w [Ci, Co, K] -> [B, Co, L * stride]."""
    B, Ci, L = x.shape
    _, Co, K = w.shape
    full = np.zeros((B, Co, (L - 1) * stride + K), np.float32)
    for k in range(K):
        full[:, :, k:k + (L - 1) * stride + 1:stride] += np.einsum("co,bcl->bol", w[:, :, k], x, dtype=np.float32)
    full += b[None, :, None]
    total = K - stride
    right = math.ceil(total * trim_right_ratio) if causal else total // 2
    left = total - right
    return np.ascontiguousarray(full[:, :, left:full.shape[-1] - right])


def lstm(x: np.ndarray, P, prefix: str, layers: int) -> np.ndarray:
    """EncodecLSTM.forward, ME:219-223: nn.LSTM over time (gate order i, f, g, o; zero initial state) plus the skip connection.
    x [B, C, T] -> [B, C, T]."""
    B, H, T = x.shape
    seq = np.transpose(x, (2, 0, 1)).astype(np.float32)                 # [T, B, H]
    inp = seq
    for l in range(layers):
        wih, whh = P[f"{prefix}.lstm.weight_ih_l{l}"], P[f"{prefix}.lstm.weight_hh_l{l}"]
        bias = P[f"{prefix}.lstm.bias_ih_l{l}"] + P[f"{prefix}.lstm.bias_hh_l{l}"]
        h = np.zeros((B, H), np.float32)
        c = np.zeros((B, H), np.float32)
        out = np.empty((T, B, H), np.float32)
        xg = inp @ wih.T + bias                                          # [T, B, 4H]
        for t in range(T):
            g = xg[t] + h @ whh.T
            i_, f_, g_, o_ = g[:, :H], g[:, H:2 * H], g[:, 2 * H:3 * H], g[:, 3 * H:]
            sig = lambda a: (1.0 / (1.0 + np.exp(-np.maximum(a, -80.0)))).astype(np.float32)     # exp(80) is finite in fp32
            c = sig(f_) * c + sig(i_) * np.tanh(g_)
            h = (sig(o_) * np.tanh(c)).astype(np.float32)
            out[t] = h
        inp = out
    return np.ascontiguousarray(np.transpose(inp + seq, (1, 2, 0)))


def _conv_params(P, prefix):
    return fold_weight_norm(P[prefix + ".conv.parametrizations.weight.original0"],
                            P[prefix + ".conv.parametrizations.weight.original1"]), P[prefix + ".conv.bias"]


def resblock(x, P, prefix, cfg, dil):
    """EncodecResnetBlock.forward, ME:256-261: shortcut(x) + conv1(ELU(conv3(ELU(x))))."""
    causal, pm = cfg["use_causal_conv"], cfg["pad_mode"]
    w, b = _conv_params(P, prefix + ".block.1")
    h = conv1d(elu(x), w, b, 1, dil, causal, pm)
    w, b = _conv_params(P, prefix + ".block.3")
    h = conv1d(elu(h), w, b, 1, 1, causal, pm)
    if cfg["use_conv_shortcut"]:
        w, b = _conv_params(P, prefix + ".shortcut")
        x = conv1d(x, w, b, 1, 1, causal, pm)
    return (x + h).astype(np.float32)


def _run(x, P, cfg, side):
    causal, pm = cfg["use_causal_conv"], cfg["pad_mode"]
    plan = layer_plan(cfg)[side]
    prev_idx = -1
    for idx, kind, s in plan:
        if idx - prev_idx == 2:                                     # an nn.ELU sits between the two parametrised layers
            x = elu(x)
        prev_idx = idx
        p = f"{side}.layers.{idx}"
        if kind == "conv":
            w, b = _conv_params(P, p)
            x = conv1d(x, w, b, s["stride"], s["dil"], causal, pm)
        elif kind == "convtr":
            w, b = _conv_params(P, p)
            x = conv_transpose1d(x, w, b, s["stride"], causal, cfg["trim_right_ratio"])
        elif kind == "res":
            x = resblock(x, P, p, cfg, s["dil"])
        else:
            x = lstm(x, P, p, cfg["num_lstm_layers"])
    return x


def encoder(wav: np.ndarray, P, cfg=CFG_24KHZ) -> np.ndarray:
    """EncodecEncoder.forward, ME:305-308: wav [B, 1, S] -> latents [B, hidden, ceil-chain(S / 320)]."""
    return _run(np.asarray(wav, np.float32), P, cfg, "encoder")


def decoder(lat: np.ndarray, P, cfg=CFG_24KHZ) -> np.ndarray:
    """EncodecDecoder.forward, ME:343-346: latents [B, hidden, T] -> wav [B, 1, 320 T]."""
    return _run(np.asarray(lat, np.float32), P, cfg, "decoder")


def num_quantizers(bandwidth_kbps: float, cfg=CFG_24KHZ) -> int:
    """encodec `ResidualVectorQuantizer.get_num_quantizers_for_bandwidth` / ME:416-422: floor(bw * 1000 / (log2(bins) * frame rate))."""
    frame_rate = math.ceil(cfg["sampling_rate"] / int(np.prod(cfg["upsampling_ratios"])))
    per_q = math.log2(cfg["codebook_size"]) * frame_rate
    return max(1, int(math.floor(bandwidth_kbps * 1000 / per_q)))


def to_transformers_model(P, cfg):
    """Pinning helper: a `transformers.EncodecModel` carrying exactly the weights `P` (used by make_golden.py and tests)."""
    import torch
    from transformers import EncodecConfig, EncodecModel
    keys = ("audio_channels", "num_filters", "kernel_size", "last_kernel_size", "residual_kernel_size", "dilation_growth_rate",
            "compress", "num_lstm_layers", "num_residual_layers", "use_conv_shortcut", "use_causal_conv", "pad_mode",
            "trim_right_ratio", "hidden_size", "sampling_rate", "codebook_size")
    kw = {k: cfg[k] for k in keys}
    kw["upsampling_ratios"] = list(cfg["upsampling_ratios"])
    kw["codebook_dim"] = cfg["hidden_size"]
    m = EncodecModel(EncodecConfig(**kw)).eval()
    sd = m.state_dict()
    for k, v in P.items():
        assert tuple(sd[k].shape) == tuple(v.shape), (k, sd[k].shape, v.shape)
        sd[k] = torch.from_numpy(v.copy())
    m.load_state_dict(sd)
    return m
