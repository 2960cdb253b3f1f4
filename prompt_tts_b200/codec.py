"""EnCodec 24 kHz on B200: SEANet encoder -> RVQ quantise / RVQ embedding sum -> SEANet decoder.

Mirror of the `encodec.EncodecModel` surface the reference uses:
    /root/reference/data_preparation/generate_code.py:13-15   model = EncodecModel.encodec_model_24khz(); model.set_target_bandwidth(6.0)
    /root/reference/data_preparation/generate_code.py:48      encoded_frames = model.encode(wav)     -> [(codes [B, 8, T], None)]
    /root/reference/decode_codec.py:8-9,16                    wav = model.decode([(codes, None)])    -> [B, 1, 320 T]
encodec 0.1.1 is a third-party dependency that is not vendored; the layer arithmetic follows its restatement in
transformers.models.encodec.modeling_encodec (cited as ME:<line>; see oracle/seanet_oracle.py for the pinning status).

All arithmetic runs in hand-written sm_100a kernels: the SEANet layers in libpt_seanet.so (include/prompt_tts_seanet.h, fp32), the
quantiser and the embedding sum in libpt_b200.so (`ops.rvq_encode` / `ops.rvq_decode`).  torch provides device memory and the
stream.  There is no CPU path: without a GPU or without the library every entry point raises `PtError`.

The layer sequencing (`SeanetStack`) talks to its kernels through a small driver object (`empty` / `ptr` / `call`); the product
only ever constructs `CudaDriver`.  (tests/test_seanet_host.py hands the same sequencing a host-side index checker built from
seanet_core.h, to verify buffer / activation routing without a GPU.)
"""
from __future__ import annotations

import ctypes as C
import math
import os
import re
from typing import Dict, List, Optional, Sequence, Tuple

from ._lib import PtError

_HERE = os.path.dirname(os.path.abspath(__file__))
SEANET_LIB_PATH = os.environ.get("PT_SEANET_LIB") or os.path.join(_HERE, "libpt_seanet.so")

# EncodecModel.encodec_model_24khz(): causal SEANet, ratios [8, 5, 4, 2], 32 filters, 2 LSTM layers, weight norm, 32 x 1024 x 128 RVQ
CFG_24KHZ = dict(audio_channels=1, num_filters=32, kernel_size=7, last_kernel_size=7, residual_kernel_size=3,
                 dilation_growth_rate=2, compress=2, num_lstm_layers=2, num_residual_layers=1,
                 upsampling_ratios=(8, 5, 4, 2), use_conv_shortcut=True, use_causal_conv=True, pad_mode="reflect",
                 trim_right_ratio=1.0, hidden_size=128, sampling_rate=24000, codebook_size=1024, codebook_dim=128,
                 num_codebooks=32)


class ConvDesc(C.Structure):
    """pt_sn_conv_t of include/prompt_tts_seanet.h."""
    _fields_ = [("x", C.c_void_p), ("w", C.c_void_p), ("bias", C.c_void_p), ("res", C.c_void_p), ("y", C.c_void_p),
                ("y_elu", C.c_void_p),
                ("B", C.c_int), ("Ci", C.c_int), ("Co", C.c_int), ("Lin", C.c_int), ("Lout", C.c_int), ("K", C.c_int),
                ("stride", C.c_int), ("dil", C.c_int), ("pad_left", C.c_int), ("reflect", C.c_int), ("Co_pad", C.c_int)]


# argtypes of every entry point of include/prompt_tts_seanet.h (p = pointer, i = int); the trailing p is the stream
SIGS = {
    "weight_norm_fold": "pppiip",
    "conv1d": "pp",
    "conv_transpose1d": "pp",
    "pack_conv_weight": "ppiiiiip",
    "conv1d_packed": "pp",
    "conv_transpose1d_packed": "pp",
    "lstm_pack": "ppip",
    "lstm_pack_bias": "pppip",
    "ncl_to_tbc": "ppiiip",
    "linear_rows": "ppppiiip",
    "lstm_step": "ppppiiip",
    "lstm_seq": "ppppiiip",
    "tbc_add_to_ncl": "ppppiiip",
}
EXPORTS = ["pt_sn_version", "pt_sn_last_error", "pt_sn_launch_count"] + ["pt_sn_" + k for k in SIGS]
_CT = {"p": C.c_void_p, "i": C.c_int}
_lib = None


def seanet_lib():
    """Load libpt_seanet.so; raise (never fall back) when it is not there."""
    global _lib
    if _lib is None:
        if not os.path.exists(SEANET_LIB_PATH):
            raise PtError(f"{SEANET_LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU / PyTorch fallback for the codec path)")
        l = C.CDLL(SEANET_LIB_PATH)
        l.pt_sn_last_error.restype = C.c_char_p
        l.pt_sn_launch_count.restype = C.c_ulonglong
        for name, sig in SIGS.items():
            fn = getattr(l, "pt_sn_" + name)
            fn.argtypes = [_CT[c] for c in sig]
            fn.restype = C.c_int
        _lib = l
    return _lib


class CudaDriver:
    """Device memory from torch, kernels from libpt_seanet.so, launches on torch's current stream."""

    def __init__(self, device="cuda"):
        import torch
        if not torch.cuda.is_available():
            raise PtError("the codec path needs a CUDA device (no CPU fallback exists)")
        self.torch = torch
        self.device = torch.device(device)
        self.lib = seanet_lib()

    def empty(self, *shape):
        return self.torch.empty(*shape, dtype=self.torch.float32, device=self.device)

    def ptr(self, buf) -> int:
        return 0 if buf is None else buf.data_ptr()

    def upload(self, array):
        t = array if isinstance(array, self.torch.Tensor) else self.torch.as_tensor(array)
        return t.detach().to(device=self.device, dtype=self.torch.float32).contiguous()

    def call(self, name: str, *args) -> None:
        stream = C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)
        rc = getattr(self.lib, "pt_sn_" + name)(*args, stream)
        if rc != 0:
            err = PtError(f"pt_sn_{name}: rc={rc}: {self.lib.pt_sn_last_error().decode()}")
            err.rc = rc
            raise err


# --------------------------------------------------------------------------------------------------------------- layer plan
def layer_plan(cfg) -> Dict[str, list]:
    """(index in `layers`, kind, spec) for the parametrised modules of EncodecEncoder (ME:281-303) and EncodecDecoder (ME:313-341);
    an index gap of 2 means an nn.ELU sits between two entries."""
    nf, ratios = cfg["num_filters"], tuple(cfg["upsampling_ratios"])
    enc, i, scale = [], 0, 1
    enc.append((i, "conv", dict(ci=cfg["audio_channels"], co=nf, k=cfg["kernel_size"], stride=1, dil=1)))
    i += 1
    for r in reversed(ratios):
        dim = scale * nf
        for j in range(cfg["num_residual_layers"]):
            enc.append((i, "res", dict(dim=dim, dil=cfg["dilation_growth_rate"] ** j)))
            i += 1
        enc.append((i + 1, "conv", dict(ci=dim, co=2 * dim, k=2 * r, stride=r, dil=1)))
        i += 2
        scale *= 2
    enc.append((i, "lstm", dict(dim=scale * nf)))
    enc.append((i + 2, "conv", dict(ci=scale * nf, co=cfg["hidden_size"], k=cfg["last_kernel_size"], stride=1, dil=1)))
    dec, i = [], 0
    scale = 2 ** len(ratios)
    dec.append((0, "conv", dict(ci=cfg["hidden_size"], co=scale * nf, k=cfg["kernel_size"], stride=1, dil=1)))
    dec.append((1, "lstm", dict(dim=scale * nf)))
    i = 2
    for r in ratios:
        dim = scale * nf
        dec.append((i + 1, "convtr", dict(ci=dim, co=dim // 2, k=2 * r, stride=r, dil=1)))
        i += 2
        for j in range(cfg["num_residual_layers"]):
            dec.append((i, "res", dict(dim=dim // 2, dil=cfg["dilation_growth_rate"] ** j)))
            i += 1
        scale //= 2
    dec.append((i + 1, "conv", dict(ci=nf, co=cfg["audio_channels"], k=cfg["last_kernel_size"], stride=1, dil=1)))
    return {"encoder": enc, "decoder": dec}


def stack_flops(cfg, side: str, L: int) -> float:
    """Algorithmic FLOPs (2 per multiply-add) of one sequence of input length L through a stack: Conv1d 2·Ci·Co·K per output
    sample, ConvTranspose1d 2·Ci·Co·ceil(K / stride), LSTM 2·(2·H·4H) per step and layer.  24 kHz model: 2.98 GFLOP per second of
    audio for either stack (encoder: L = samples, decoder: L = frames)."""
    total = 0.0
    for _, kind, s in layer_plan(cfg)[side]:
        if kind == "conv":
            L = -(-L // s["stride"])
            total += 2.0 * s["ci"] * s["co"] * s["k"] * L
        elif kind == "convtr":
            L = L * s["stride"]
            total += 2.0 * s["ci"] * s["co"] * -(-s["k"] // s["stride"]) * L
        elif kind == "res":
            dim, hid = s["dim"], s["dim"] // cfg["compress"]
            total += 2.0 * L * (dim * hid * cfg["residual_kernel_size"] + hid * dim + (dim * dim if cfg["use_conv_shortcut"] else 0))
        else:
            total += 2.0 * L * cfg["num_lstm_layers"] * 2 * s["dim"] * 4 * s["dim"]
    return total


_G, _V = ".conv.parametrizations.weight.original0", ".conv.parametrizations.weight.original1"


def normalise_key(key: str) -> str:
    """Map a state_dict key of encodec 0.1.1 (`encoder.model.3.conv.conv.weight_g`, `decoder.model.3.convtr.convtr.weight_v`,
    `quantizer.vq.layers.0._codebook.embed`) or of transformers with old-style weight norm onto the transformers >= 4.31 names
    this module stores (`encoder.layers.3.conv.parametrizations.weight.original0`, `quantizer.layers.0.codebook.embed`)."""
    k = re.sub(r"^(encoder|decoder)\.model\.", r"\1.layers.", key)
    k = k.replace(".conv.conv.", ".conv.").replace(".convtr.convtr.", ".conv.")
    k = re.sub(r"\.weight_g$", ".parametrizations.weight.original0", k)
    k = re.sub(r"\.weight_v$", ".parametrizations.weight.original1", k)
    k = re.sub(r"^quantizer\.vq\.layers\.(\d+)\._codebook\.", r"quantizer.layers.\1.codebook.", k)
    return k


# --------------------------------------------------------------------------------------------------------------- one stack
class SeanetStack:
    """The encoder or the decoder: weights prepared once (weight norm folded, LSTM weights packed), then one kernel per layer.

    The activation between two layers is applied by the PRODUCER: every kernel can store its result raw, through ELU, or both, and
    `_needs` works out which of the two the consumers read (a residual block reads its input both ways: ELU'd by its first
    convolution, raw by its shortcut)."""

    def __init__(self, cfg, side: str, drv, fast: Optional[bool] = None):
        self.cfg, self.side, self.drv = cfg, side, drv
        # fast: packed-weight convolutions + the whole-sequence LSTM launch; PT_SN_LEGACY=1 selects the first-draft kernels
        # (reference-layout weights, one launch per LSTM step), kept as the A/B partner and as the fallback of pt_sn_lstm_seq
        self.fast = (os.environ.get("PT_SN_LEGACY", "0") != "1") if fast is None else fast
        self.lstm_whole = self.fast and os.environ.get("PT_SN_LSTM_STEPS", "0") != "1"     # measurement switch: one launch per step
        self.plan = layer_plan(cfg)[side]
        self.causal = bool(cfg["use_causal_conv"])
        self.reflect = 1 if cfg["pad_mode"] == "reflect" else 0
        self.w: Dict[str, object] = {}

    # ---- preparation
    def param_shapes(self) -> Dict[str, Tuple[int, ...]]:
        """state_dict name -> shape of every tensor this stack reads (transformers naming; weight norm: original0 = g, original1 = v)."""
        out: Dict[str, Tuple[int, ...]] = {}

        def conv(prefix, ci, co, k, transposed=False):
            out[prefix + _G] = ((ci if transposed else co), 1, 1)
            out[prefix + _V] = (ci, co, k) if transposed else (co, ci, k)
            out[prefix + ".conv.bias"] = (co,)

        for idx, kind, s in self.plan:
            p = f"{self.side}.layers.{idx}"
            if kind in ("conv", "convtr"):
                conv(p, s["ci"], s["co"], s["k"], kind == "convtr")
            elif kind == "res":
                dim, hid = s["dim"], s["dim"] // self.cfg["compress"]
                conv(p + ".block.1", dim, hid, self.cfg["residual_kernel_size"])
                conv(p + ".block.3", hid, dim, 1)
                if self.cfg["use_conv_shortcut"]:
                    conv(p + ".shortcut", dim, dim, 1)
            else:
                H = s["dim"]
                for l in range(self.cfg["num_lstm_layers"]):
                    out[f"{p}.lstm.weight_ih_l{l}"] = (4 * H, H)
                    out[f"{p}.lstm.weight_hh_l{l}"] = (4 * H, H)
                    out[f"{p}.lstm.bias_ih_l{l}"] = (4 * H,)
                    out[f"{p}.lstm.bias_hh_l{l}"] = (4 * H,)
        return out

    def param_names(self) -> List[str]:
        return list(self.param_shapes())

    def prepare(self, params: Dict[str, object]) -> None:
        """params: name -> device buffer with the state_dict's shape.  Folds g * v / |v| (pt_sn_weight_norm_fold) and packs the LSTM
        matrices gate-interleaved and transposed (pt_sn_lstm_pack)."""
        d = self.drv

        def fold(prefix, ci, co, k, transposed=False):
            rows, cols = (ci, co * k) if transposed else (co, ci * k)
            w = d.empty(rows * cols)
            d.call("weight_norm_fold", d.ptr(params[prefix + _V]), d.ptr(params[prefix + _G]), d.ptr(w), rows, cols)
            self.w[prefix + ".w"] = w
            self.w[prefix + ".b"] = params[prefix + ".conv.bias"]
            if self.fast:                       # [Ci, K, Co_pad]: a thread's output channels become 16-byte loads
                cop = self.co_pad(co)
                wp = d.empty(ci * k * cop)
                d.call("pack_conv_weight", d.ptr(w), d.ptr(wp), co, ci, k, cop, 1 if transposed else 0)
                self.w[prefix + ".wp"] = wp

        for idx, kind, s in self.plan:
            p = f"{self.side}.layers.{idx}"
            if kind == "conv":
                fold(p, s["ci"], s["co"], s["k"])
            elif kind == "convtr":
                fold(p, s["ci"], s["co"], s["k"], transposed=True)
            elif kind == "res":
                dim, hid = s["dim"], s["dim"] // self.cfg["compress"]
                fold(p + ".block.1", dim, hid, self.cfg["residual_kernel_size"])
                fold(p + ".block.3", hid, dim, 1)
                if self.cfg["use_conv_shortcut"]:
                    fold(p + ".shortcut", dim, dim, 1)
            else:
                H = s["dim"]
                if H % 4:
                    raise PtError(f"LSTM width {H} must be a multiple of 4")
                for l in range(self.cfg["num_lstm_layers"]):
                    for n in ("weight_ih", "weight_hh"):
                        t4 = d.empty(H * H * 4)
                        d.call("lstm_pack", d.ptr(params[f"{p}.lstm.{n}_l{l}"]), d.ptr(t4), H)
                        self.w[f"{p}.{n}_l{l}"] = t4
                    b4 = d.empty(H * 4)
                    d.call("lstm_pack_bias", d.ptr(params[f"{p}.lstm.bias_ih_l{l}"]), d.ptr(params[f"{p}.lstm.bias_hh_l{l}"]), d.ptr(b4), H)
                    self.w[f"{p}.bias_l{l}"] = b4

    @staticmethod
    def co_pad(co: int) -> int:
        """Row length of the packed weights: 16 output channels per thread from 9 channels up, else 8."""
        return 8 if co <= 8 else -(-co // 16) * 16

    # ---- which forms of a layer's output are read
    def _needs(self, pos: int) -> Tuple[bool, bool]:
        """(raw, elu) for the output of plan entry `pos`."""
        if pos + 1 >= len(self.plan):
            return True, False
        idx, _, _ = self.plan[pos]
        nidx, nkind, _ = self.plan[pos + 1]
        gap = nidx - idx
        if nkind == "res":
            if gap != 1:
                raise PtError("an ELU in front of a residual block is not a SEANet layout")
            return True, True
        return (True, False) if gap == 1 else (False, True)

    # ---- layers
    def _conv(self, prefix, x, B, ci, co, L, k, stride, dil, raw, elu, res=None, transposed=False):
        d = self.drv
        if transposed:
            total = k - stride
            right = math.ceil(total * self.cfg["trim_right_ratio"]) if self.causal else total // 2      # ME:196-204
            pad_left, Lout = total - right, L * stride
        else:
            total = (k - 1) * dil + 1 - stride                                                          # ME:121-123
            pad_left = total if self.causal else total - total // 2                                    # ME:153-163
            Lout = -(-L // stride)                                                                      # ME:125-133: extra right padding
        y = d.empty(B, co, Lout) if raw else None
        ye = d.empty(B, co, Lout) if elu else None
        name = "conv_transpose1d" if transposed else "conv1d"
        if self.fast:
            desc = ConvDesc(d.ptr(x), d.ptr(self.w[prefix + ".wp"]), d.ptr(self.w[prefix + ".b"]), d.ptr(res), d.ptr(y), d.ptr(ye),
                            B, ci, co, L, Lout, k, stride, dil, pad_left, self.reflect, self.co_pad(co))
            name += "_packed"
        else:
            desc = ConvDesc(d.ptr(x), d.ptr(self.w[prefix + ".w"]), d.ptr(self.w[prefix + ".b"]), d.ptr(res), d.ptr(y), d.ptr(ye),
                            B, ci, co, L, Lout, k, stride, dil, pad_left, self.reflect, 0)
        d.call(name, C.addressof(desc))
        return y, ye, Lout

    def _res(self, prefix, x, xe, B, dim, L, dil, raw, elu):
        """EncodecResnetBlock.forward ME:256-261."""
        hid = dim // self.cfg["compress"]
        _, he, _ = self._conv(prefix + ".block.1", xe, B, dim, hid, L, self.cfg["residual_kernel_size"], 1, dil, False, True)
        sc = x
        if self.cfg["use_conv_shortcut"]:
            sc, _, _ = self._conv(prefix + ".shortcut", x, B, dim, dim, L, 1, 1, 1, True, False)
        y, ye, _ = self._conv(prefix + ".block.3", he, B, hid, dim, L, 1, 1, 1, raw, elu, res=sc)
        return y, ye

    def _lstm(self, prefix, x, B, H, T, raw, elu):
        """EncodecLSTM.forward ME:219-223: time-major, nn.LSTM layers, skip connection."""
        d = self.drv
        inp = d.empty(T, B, H)
        d.call("ncl_to_tbc", d.ptr(x), d.ptr(inp), B, H, T)
        for l in range(self.cfg["num_lstm_layers"]):
            xg = d.empty(T, B, H, 4)
            d.call("linear_rows", d.ptr(inp), d.ptr(self.w[f"{prefix}.weight_ih_l{l}"]), d.ptr(self.w[f"{prefix}.bias_l{l}"]), d.ptr(xg),
                   T * B, H, 4 * H)
            hseq, c = d.empty(T, B, H), d.empty(B, H)
            whh = d.ptr(self.w[f"{prefix}.weight_hh_l{l}"])
            whole = self.lstm_whole
            if whole:
                try:                                  # all T steps in one cooperative launch
                    d.call("lstm_seq", d.ptr(xg), whh, d.ptr(hseq), d.ptr(c), T, B, H)
                except PtError as e:
                    if getattr(e, "rc", 0) != -3:     # -3: the blocks cannot be co-resident on this device -> one launch per step
                        raise
                    whole = False
            if not whole:
                for t in range(T):
                    d.call("lstm_step", d.ptr(xg), whh, d.ptr(hseq), d.ptr(c), t, B, H)
            inp = hseq
        y = d.empty(B, H, T) if raw else None
        ye = d.empty(B, H, T) if elu else None
        d.call("tbc_add_to_ncl", d.ptr(inp), d.ptr(x), d.ptr(y), d.ptr(ye), B, H, T)
        return y, ye

    def forward(self, x, B: int, L: int):
        """x: device buffer [B, C_in, L] fp32 -> (device buffer [B, C_out, L_out], L_out)."""
        if not self.w:
            raise PtError("SeanetStack.forward before prepare()")
        xe = None
        for pos, (idx, kind, s) in enumerate(self.plan):
            raw, elu = self._needs(pos)
            p = f"{self.side}.layers.{idx}"
            gap = idx - self.plan[pos - 1][0] if pos else 1
            src = x if gap == 1 else xe                     # an ELU in front: read the producer's activated copy
            if kind == "conv":
                x, xe, L = self._conv(p, src, B, s["ci"], s["co"], L, s["k"], s["stride"], s["dil"], raw, elu)
            elif kind == "convtr":
                x, xe, L = self._conv(p, src, B, s["ci"], s["co"], L, s["k"], s["stride"], 1, raw, elu, transposed=True)
            elif kind == "res":
                x, xe = self._res(p, x, xe, B, s["dim"], L, s["dil"], raw, elu)
            else:
                x, xe = self._lstm(p, src, B, s["dim"], L, raw, elu)
        return x, L


def random_state_dict(cfg, seed: Optional[int] = None):
    """The tensors a freshly constructed (not pretrained) EnCodec holds, as `encodec_model_24khz(pretrained=False)` returns it in
    the reference's package: PyTorch's default initialisation -- Conv / ConvTranspose v ~ U(+-1/sqrt(fan_in)) with weight-norm
    g = |v| (so the effective weight is v), biases and LSTM tensors U(+-1/sqrt(fan_in or H)), codebooks U(+-sqrt(3 / dim)).
    Host tensors from torch's CPU generator (the global one, or a fresh one seeded with `seed`)."""
    import torch
    gen = None if seed is None else torch.Generator().manual_seed(int(seed))

    def uniform(shape, bound):
        return (torch.rand(shape, generator=gen) * 2 - 1) * bound

    sd = {}
    shapes = {}
    for side in ("encoder", "decoder"):
        shapes.update(SeanetStack(cfg, side, None, fast=True).param_shapes())
    for name, shape in shapes.items():
        if name.endswith(_V):
            bound = 1.0 / math.sqrt(shape[1] * shape[2])
            v = uniform(shape, bound)
            sd[name] = v
            sd[name[:-len(_V)] + _G] = v.flatten(1).norm(dim=1).reshape(-1, 1, 1)
            sd[name[:-len(_V)] + ".conv.bias"] = uniform(shapes[name[:-len(_V)] + ".conv.bias"], bound)
        elif ".lstm." in name:
            sd[name] = uniform(shape, 1.0 / math.sqrt(shape[-1] if len(shape) == 2 else shape[0] // 4))
    dim = cfg["codebook_dim"] if "codebook_dim" in cfg else cfg["hidden_size"]
    for q in range(cfg.get("num_codebooks", 32)):
        sd[f"quantizer.layers.{q}.codebook.embed"] = uniform((cfg["codebook_size"], dim), math.sqrt(3.0 / dim))
    return sd


# --------------------------------------------------------------------------------------------------------------- model
class EncodecModel:
    """Drop-in for the `encodec.EncodecModel` calls of generate_code.py / decode_codec.py (24 kHz, mono, causal, no chunking)."""

    def __init__(self, cfg=None, device="cuda"):
        self.cfg = dict(CFG_24KHZ if cfg is None else cfg)
        self.drv = CudaDriver(device)
        self.sample_rate = self.cfg["sampling_rate"]
        self.channels = self.cfg["audio_channels"]
        self.frame_rate = math.ceil(self.sample_rate / math.prod(self.cfg["upsampling_ratios"]))
        self.bandwidth: Optional[float] = None
        self.segment = None                                   # the 24 kHz model encodes the whole clip as one frame
        self.encoder = SeanetStack(self.cfg, "encoder", self.drv)
        self.decoder = SeanetStack(self.cfg, "decoder", self.drv)
        self._params: Dict[str, object] = {}
        self.codebooks = None                                 # [num_codebooks, K, D] fp32 on the device
        self._prepared_rvq = None

    # ---- construction, as in encodec/model.py
    @staticmethod
    def encodec_model_24khz(pretrained: bool = True, repository=None, device="cuda") -> "EncodecModel":
        m = EncodecModel(CFG_24KHZ, device)
        if pretrained:
            path = None if repository is None else os.path.join(str(repository), "encodec_24khz-d7cc33bc.th")
            if path is None or not os.path.exists(path):
                raise PtError("pretrained EnCodec weights cannot be downloaded here: pass `repository=<dir holding "
                              "encodec_24khz-d7cc33bc.th>` or `pretrained=False` and call load_state_dict()")
            import torch
            m.load_state_dict(torch.load(path, map_location="cpu"))
        else:
            m.load_state_dict(random_state_dict(m.cfg))      # a randomly initialised model, as the reference's constructor gives
        return m

    def set_target_bandwidth(self, bandwidth: float) -> None:
        if bandwidth not in (1.5, 3.0, 6.0, 12.0, 24.0):
            raise ValueError(f"This model doesn't support the bandwidth {bandwidth}. Select one of [1.5, 3.0, 6.0, 12.0, 24.0].")
        self.bandwidth = bandwidth

    @property
    def num_quantizers(self) -> int:
        """encodec `get_num_quantizers_for_bandwidth` (ME:416-422): 6 kbps -> 8 codebooks of 10 bits at 75 frames/s."""
        n = self.cfg["num_codebooks"]
        if self.bandwidth is not None and self.bandwidth > 0:
            per_q = math.log2(self.cfg["codebook_size"]) * self.frame_rate
            n = int(max(1, math.floor(self.bandwidth * 1000 / per_q)))
        return min(n, self.cfg["num_codebooks"])

    def load_state_dict(self, state_dict) -> None:
        """Accepts the key dialects of encodec 0.1.1 and of transformers' EncodecModel (see `normalise_key`)."""
        sd = {normalise_key(k): v for k, v in state_dict.items()}
        need = {**self.encoder.param_shapes(), **self.decoder.param_shapes()}
        missing = [k for k in need if k not in sd]
        if missing:
            raise KeyError(f"state_dict is missing {len(missing)} SEANet tensors, e.g. {missing[:3]}")
        wrong = [(k, tuple(sd[k].shape), shp) for k, shp in need.items() if tuple(sd[k].shape) != shp]
        if wrong:
            raise ValueError(f"state_dict shapes do not match the 24 kHz SEANet, e.g. {wrong[:2]}")
        self._params = {k: self.drv.upload(sd[k]) for k in need}
        self.encoder.prepare(self._params)
        self.decoder.prepare(self._params)
        cbs = [sd[f"quantizer.layers.{q}.codebook.embed"] for q in range(self.cfg["num_codebooks"])
               if f"quantizer.layers.{q}.codebook.embed" in sd]
        if cbs:
            torch = self.drv.torch
            self.codebooks = torch.stack([self.drv.upload(c) for c in cbs]).contiguous()
            self._prepared_rvq = None

    # ---- nn.Module manners the reference's scripts rely on (generate_code.py:15 `model.to(device)`)
    def to(self, device=None, *args, **kwargs) -> "EncodecModel":
        torch = self.drv.torch
        if device is not None and not isinstance(device, torch.dtype):
            d = torch.device(device)
            if d.type != "cuda" or (d.index is not None and self.drv.device.index is not None and d.index != self.drv.device.index):
                raise PtError(f"EncodecModel lives on {self.drv.device} (its kernels are CUDA-only); host tensors may be passed to "
                              "encode() / decode() directly")
        return self

    def cuda(self, device=None) -> "EncodecModel":
        return self

    def eval(self) -> "EncodecModel":
        return self

    def train(self, mode: bool = True) -> "EncodecModel":
        if mode:
            raise NotImplementedError("the codec is inference-only here (the reference never trains it)")
        return self

    def _home(self, t, dtype, what: str):
        """(device copy of `t`, the device `t` came from).  encode() / decode() accept host tensors the way the reference's
        decode_codec.py passes them: copied in, results copied back -- the arithmetic still only runs on the GPU."""
        torch = self.drv.torch
        if not isinstance(t, torch.Tensor):
            raise TypeError(f"EncodecModel.{what}: expected a torch.Tensor, got {type(t).__name__}")
        return t.to(device=self.drv.device, dtype=dtype).contiguous(), t.device

    # ---- the two calls of the reference
    def _check_wave(self, x):
        torch = self.drv.torch
        if not isinstance(x, torch.Tensor) or not x.is_cuda:
            raise PtError("EncodecModel: the input must be a CUDA tensor (no CPU fallback exists)")
        if x.dim() != 3 or x.shape[1] != self.channels:
            raise ValueError(f"expected [B, {self.channels}, T], got {tuple(x.shape)}")
        return x.to(torch.float32).contiguous()

    def encode_latents(self, x):
        """wav [B, 1, S] -> SEANet latents [B, 128, ceil(S / 320)].  Device tensors only (the layer-level entry point)."""
        x = self._check_wave(x)
        lat, _ = self.encoder.forward(x, x.shape[0], x.shape[2])
        return lat

    def encode(self, x) -> List[Tuple[object, Optional[object]]]:
        """`EncodecModel.encode` (generate_code.py:48): one frame `(codes [B, n_q, T] int64, None)` -- no chunking, no scale."""
        from . import ops
        if self.codebooks is None:
            raise PtError("EncodecModel.encode: no codebooks loaded")
        x, home = self._home(x, self.drv.torch.float32, "encode")
        lat = self.encode_latents(x)
        nq = self.num_quantizers
        cb = self.codebooks[:nq]
        if self._prepared_rvq is None or self._prepared_rvq[0] != nq:
            self._prepared_rvq = (nq, ops.rvq_prepare(cb) if cb.shape[1] % 128 == 0 and cb.shape[1] <= 1024 and cb.shape[2] == 128 else None)
        codes = ops.rvq_encode(lat, cb, prepared=self._prepared_rvq[1])
        return [(codes.to(home), None)]

    def decode_latents(self, lat):
        torch = self.drv.torch
        if not isinstance(lat, torch.Tensor) or not lat.is_cuda:
            raise PtError("EncodecModel: the input must be a CUDA tensor (no CPU fallback exists)")
        lat = lat.to(torch.float32).contiguous()
        wav, _ = self.decoder.forward(lat, lat.shape[0], lat.shape[2])
        return wav

    def decode(self, encoded_frames: Sequence[Tuple[object, Optional[object]]]):
        """`EncodecModel.decode` (decode_codec.py:16): frames [(codes [B, n_q, T], None)] -> wav [B, 1, 320 T]."""
        from . import ops
        if len(encoded_frames) != 1:
            raise NotImplementedError("the 24 kHz model has no segmenting: exactly one frame is expected")
        codes, scale = encoded_frames[0]
        if scale is not None:
            raise NotImplementedError("scaled frames belong to the 48 kHz model (normalize=True)")
        if self.codebooks is None:
            raise PtError("EncodecModel.decode: no codebooks loaded")
        codes, home = self._home(codes, self.drv.torch.int64, "decode")
        if codes.dim() != 3:
            raise ValueError(f"expected codes [B, n_q, T], got {tuple(codes.shape)}")
        nq = codes.shape[1]
        lat = ops.rvq_decode(codes, self.codebooks[:nq].contiguous())
        return self.decode_latents(lat).to(home)
