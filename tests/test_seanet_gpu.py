"""GPU parity of the EnCodec SEANet row, through the C ABI of libpt_seanet.so, against oracle/seanet_oracle.py.

Tolerance: fp32 kernels against an fp32 oracle that sums in a different order -- relative L2 error <= 1e-4 on latents and waveform
(measured ~1e-6 per layer); the RVQ codes of OUR latents are bit-exact against the C oracle on the same latents, and agree with the
codes of the oracle's latents except where a rounding-level difference of the latents crosses a decision boundary."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = 1e-4


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


@pytest.fixture(scope="module")
def drv(cuda):
    from prompt_tts_b200 import codec
    return codec.CudaDriver(cuda)


def _dev(drv, a):
    return drv.upload(np.ascontiguousarray(a, np.float32))


CONV_CASES = [
    # B, Ci, Co, L, K, stride, dil, causal, reflect
    (2, 1, 32, 5000, 7, 1, 1, True, True),
    (1, 3, 13, 1100, 7, 1, 1, False, True),
    (2, 32, 64, 2523, 4, 2, 1, True, True),
    (1, 64, 128, 601, 8, 4, 1, True, True),
    (1, 16, 24, 333, 10, 5, 1, False, True),
    (2, 24, 40, 129, 16, 8, 1, True, True),
    (1, 32, 16, 97, 3, 1, 2, True, True),
    (1, 6, 3, 97, 3, 1, 4, False, False),
    (2, 32, 1, 4000, 7, 1, 1, True, True),
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv1d_kernel(drv, case):
    import seanet_oracle as so
    from prompt_tts_b200 import codec
    B, Ci, Co, L, K, stride, dil, causal, reflect = case
    rng = np.random.default_rng(hash(case) % 2**32)
    x = rng.standard_normal((B, Ci, L)).astype(np.float32)
    w = rng.standard_normal((Co, Ci, K)).astype(np.float32) / np.float32(np.sqrt(Ci * K))
    b = rng.standard_normal(Co).astype(np.float32)
    ref = so.conv1d(x, w, b, stride, dil, causal, "reflect" if reflect else "constant")
    res = rng.standard_normal(ref.shape).astype(np.float32)
    Lout = ref.shape[-1]
    left, _ = so.pad_amounts(K, stride, dil, causal)
    xd, wd, bd, rd = _dev(drv, x), _dev(drv, w), _dev(drv, b), _dev(drv, res)
    y, ye = drv.empty(B, Co, Lout).fill_(float("nan")), drv.empty(B, Co, Lout).fill_(float("nan"))
    d = codec.ConvDesc(xd.data_ptr(), wd.data_ptr(), bd.data_ptr(), rd.data_ptr(), y.data_ptr(), ye.data_ptr(),
                       B, Ci, Co, L, Lout, K, stride, dil, left, 1 if reflect else 0)
    drv.call("conv1d", C.addressof(d))
    assert rel(y.cpu().numpy(), ref + res) < 1e-5
    assert rel(ye.cpu().numpy(), so.elu(ref + res)) < 1e-5
    # packed weights: 8 / 16 output channels per thread; same products in the same order -> identical bits
    for cop in sorted({codec.SeanetStack.co_pad(Co), -(-Co // 8) * 8}):
        wp = drv.empty(Ci, K, cop).fill_(float("nan"))
        drv.call("pack_conv_weight", wd.data_ptr(), wp.data_ptr(), Co, Ci, K, cop, 0)
        y3, ye3 = drv.empty(B, Co, Lout).fill_(float("nan")), drv.empty(B, Co, Lout).fill_(float("nan"))
        d = codec.ConvDesc(xd.data_ptr(), wp.data_ptr(), bd.data_ptr(), rd.data_ptr(), y3.data_ptr(), ye3.data_ptr(),
                           B, Ci, Co, L, Lout, K, stride, dil, left, 1 if reflect else 0, cop)
        drv.call("conv1d_packed", C.addressof(d))
        assert rel(y3.cpu().numpy(), ref + res) < 1e-5
        assert torch.equal(y3, y) and torch.equal(ye3, ye), cop


CONVTR_CASES = [
    # B, Ci, Co, L, K, stride, causal
    (2, 64, 32, 75, 16, 8, True),
    (1, 48, 24, 130, 10, 5, True),
    (1, 16, 8, 300, 8, 4, False),
    (2, 8, 4, 3000, 4, 2, True),
    (1, 3, 7, 40, 7, 3, True),
]


@pytest.mark.parametrize("case", CONVTR_CASES)
def test_conv_transpose_kernel(drv, case):
    import seanet_oracle as so
    from prompt_tts_b200 import codec
    B, Ci, Co, L, K, stride, causal = case
    rng = np.random.default_rng(hash(case) % 2**32)
    x = rng.standard_normal((B, Ci, L)).astype(np.float32)
    w = rng.standard_normal((Ci, Co, K)).astype(np.float32) / np.float32(np.sqrt(Ci * 2))
    b = rng.standard_normal(Co).astype(np.float32)
    ref = so.conv_transpose1d(x, w, b, stride, causal)
    total = K - stride
    left = 0 if causal else total - total // 2
    xd, wd, bd = _dev(drv, x), _dev(drv, w), _dev(drv, b)
    y, ye = drv.empty(*ref.shape).fill_(float("nan")), drv.empty(*ref.shape).fill_(float("nan"))
    d = codec.ConvDesc(xd.data_ptr(), wd.data_ptr(), bd.data_ptr(), 0, y.data_ptr(), ye.data_ptr(),
                       B, Ci, Co, L, ref.shape[-1], K, stride, 1, left, 0)
    drv.call("conv_transpose1d", C.addressof(d))
    assert rel(y.cpu().numpy(), ref) < 1e-5
    assert rel(ye.cpu().numpy(), so.elu(ref)) < 1e-5
    cop = codec.SeanetStack.co_pad(Co)
    wp = drv.empty(Ci, K, cop).fill_(float("nan"))
    drv.call("pack_conv_weight", wd.data_ptr(), wp.data_ptr(), Co, Ci, K, cop, 1)
    y3, ye3 = drv.empty(*ref.shape).fill_(float("nan")), drv.empty(*ref.shape).fill_(float("nan"))
    d = codec.ConvDesc(xd.data_ptr(), wp.data_ptr(), bd.data_ptr(), 0, y3.data_ptr(), ye3.data_ptr(),
                       B, Ci, Co, L, ref.shape[-1], K, stride, 1, left, 0, cop)
    drv.call("conv_transpose1d_packed", C.addressof(d))
    assert rel(y3.cpu().numpy(), ref) < 1e-5
    assert torch.equal(y3, y) and torch.equal(ye3, ye)


def test_bad_arguments_raise(drv):
    from prompt_tts_b200 import codec
    from prompt_tts_b200._lib import PtError
    x = drv.empty(1, 1, 4)
    w = drv.empty(1, 1, 7)
    y = drv.empty(1, 1, 4)
    d = codec.ConvDesc(x.data_ptr(), w.data_ptr(), 0, 0, y.data_ptr(), 0, 1, 1, 1, 4, 4, 7, 1, 1, 6, 1)     # reflect pad 6 > length - 1
    with pytest.raises(PtError, match="shorter than its reflect padding"):
        drv.call("conv1d", C.addressof(d))
    with pytest.raises(PtError, match="multiples of 4"):
        drv.call("linear_rows", x.data_ptr(), w.data_ptr(), 0, y.data_ptr(), 1, 3, 4)


@pytest.mark.parametrize("B,H,T", [(1, 16, 9), (3, 64, 21), (37, 32, 5), (32, 512, 12), (2, 1024, 3)])
def test_lstm_whole_sequence_equals_steps_and_oracle(drv, B, H, T):
    """pt_sn_lstm_seq (one cooperative launch, grid barrier per step) against T launches of pt_sn_lstm_step and the oracle's
    explicit loop; H = 1024 needs 256 co-resident blocks, which a B200 does not have: the -3 fallback is exercised."""
    import seanet_oracle as so
    from prompt_tts_b200 import codec
    cfg = dict(so.CFG_TINY)
    rng = np.random.default_rng(B * 100 + H)
    prefix = "encoder.layers.13"
    P = {}
    for l in range(2):
        for n, shape in (("weight_ih", (4 * H, H)), ("weight_hh", (4 * H, H)), ("bias_ih", (4 * H,)), ("bias_hh", (4 * H,))):
            P[f"{prefix}.lstm.{n}_l{l}"] = (rng.uniform(-1, 1, shape) / np.sqrt(H) * 2).astype(np.float32)
    x = rng.standard_normal((B, H, T)).astype(np.float32)
    ref = so.lstm(x, P, prefix, 2)
    st = codec.SeanetStack(cfg, "encoder", drv)
    keep = []           # uploaded operands must outlive the asynchronous pack kernels (a freed block is handed to the next upload)

    def up(a):
        keep.append(_dev(drv, a))
        return keep[-1].data_ptr()

    for l in range(2):
        for n in ("weight_ih", "weight_hh"):
            t4 = drv.empty(H, H, 4)
            drv.call("lstm_pack", up(P[f"{prefix}.lstm.{n}_l{l}"]), t4.data_ptr(), H)
            st.w[f"{prefix}.{n}_l{l}"] = t4
        b4 = drv.empty(H, 4)
        drv.call("lstm_pack_bias", up(P[f"{prefix}.lstm.bias_ih_l{l}"]), up(P[f"{prefix}.lstm.bias_hh_l{l}"]), b4.data_ptr(), H)
        st.w[f"{prefix}.bias_l{l}"] = b4
    xd = _dev(drv, x)
    lib = codec.seanet_lib()
    outs = {}
    for whole in (True, False):
        st.lstm_whole = whole
        n0 = lib.pt_sn_launch_count()
        y, ye = st._lstm(prefix, xd, B, H, T, True, True)
        torch.cuda.synchronize()
        launches = lib.pt_sn_launch_count() - n0
        # transposes (2) + input projections (2) + either one launch per layer or one per step and layer
        assert launches == 4 + (2 if whole and H <= 512 else 2 * T), (whole, launches)
        outs[whole] = (y, ye)
    # the whole-sequence kernel does the step kernel's arithmetic in the step kernel's order: identical bits
    assert torch.equal(outs[True][0], outs[False][0]) and torch.equal(outs[True][1], outs[False][1])
    # against the oracle: the recurrence amplifies the last-bit differences between CUDA's and the host's expf / tanhf
    assert rel(outs[True][0].cpu().numpy(), ref) < TOL
    assert rel(outs[True][1].cpu().numpy(), so.elu(ref)) < TOL


@pytest.mark.parametrize("fast", [True, False])
@pytest.mark.parametrize("name,S,B", [("tiny", 3203, 3), ("tiny_noshortcut", 2900, 1), ("tiny_noncausal", 3333, 2), ("k24", 3040, 2)])
def test_stacks_match_oracle(drv, name, S, B, fast):
    import seanet_oracle as so
    from prompt_tts_b200 import codec
    cfg = {"tiny": so.CFG_TINY, "k24": so.CFG_24KHZ, "tiny_noshortcut": dict(so.CFG_TINY, use_conv_shortcut=False),
           "tiny_noncausal": dict(so.CFG_TINY, use_causal_conv=False)}[name]
    P = so.make_weights(cfg, 11)
    x = (np.random.default_rng(2).standard_normal((B, 1, S)) * 0.3).astype(np.float32)
    enc = codec.SeanetStack(cfg, "encoder", drv, fast=fast)
    enc.prepare({k: drv.upload(P[k]) for k in enc.param_names()})
    lat, T = enc.forward(_dev(drv, x), B, S)
    lat_ref = so.encoder(x, P, cfg)
    assert tuple(lat.shape) == lat_ref.shape and T == lat_ref.shape[-1]
    assert rel(lat.cpu().numpy(), lat_ref) < TOL
    dec = codec.SeanetStack(cfg, "decoder", drv, fast=fast)
    dec.prepare({k: drv.upload(P[k]) for k in dec.param_names()})
    wav, L = dec.forward(_dev(drv, lat_ref), B, T)
    wav_ref = so.decoder(lat_ref, P, cfg)
    assert tuple(wav.shape) == wav_ref.shape and L == 320 * T
    assert rel(wav.cpu().numpy(), wav_ref) < TOL


def test_golden_vectors_on_gpu(drv):
    import os
    import seanet_oracle as so
    from prompt_tts_b200 import codec
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "seanet_golden.npz"))
    for name, cfg in (("tiny", so.CFG_TINY), ("k24", so.CFG_24KHZ)):
        P = so.make_weights(cfg, int(g[f"{name}_seed"]))
        enc, dec = codec.SeanetStack(cfg, "encoder", drv), codec.SeanetStack(cfg, "decoder", drv)
        enc.prepare({k: drv.upload(P[k]) for k in enc.param_names()})
        dec.prepare({k: drv.upload(P[k]) for k in dec.param_names()})
        wav = g[f"{name}_wav"]
        lat, T = enc.forward(_dev(drv, wav), wav.shape[0], wav.shape[2])
        assert rel(lat.cpu().numpy(), g[f"{name}_lat"]) < TOL
        out, _ = dec.forward(_dev(drv, g[f"{name}_lat"]), wav.shape[0], T)
        assert rel(out.cpu().numpy(), g[f"{name}_out"]) < TOL


def test_encodec_model_encode_decode(cuda):
    """The reference's two calls (generate_code.py:13-15,48; decode_codec.py:16) end to end: SEANet encoder -> RVQ codes -> embedding
    sum -> SEANet decoder, 24 kHz widths, 6 kbps = 8 codebooks."""
    import torch
    import rvq_oracle
    import seanet_oracle as so
    from prompt_tts_b200 import codec
    cfg = so.CFG_24KHZ
    P = so.make_weights(cfg, 5)
    rng = np.random.default_rng(9)
    B, S = 2, 6400
    wav = (rng.standard_normal((B, 1, S)) * 0.3).astype(np.float32)
    lat_ref = so.encoder(wav, P, cfg)
    # codebooks on the scale of the latents so that all 8 stages carry signal
    cb = (rng.standard_normal((32, 1024, 128)) * lat_ref.std()).astype(np.float32)
    for q in range(1, 32):
        cb[q] *= 0.7 ** q
    sd = {k: torch.from_numpy(v) for k, v in P.items()}
    sd.update({f"quantizer.layers.{q}.codebook.embed": torch.from_numpy(cb[q]) for q in range(32)})
    model = codec.EncodecModel.encodec_model_24khz(pretrained=False, device=cuda)
    model.load_state_dict(sd)
    model.set_target_bandwidth(6.0)
    assert model.num_quantizers == 8
    frames = model.encode(torch.from_numpy(wav).to(cuda))
    assert len(frames) == 1 and frames[0][1] is None
    codes = frames[0][0]
    assert codes.dtype == torch.int64 and tuple(codes.shape) == (B, 8, S // 320)
    lat = model.encode_latents(torch.from_numpy(wav).to(cuda)).cpu().numpy()
    assert rel(lat, lat_ref) < TOL
    # the quantiser is exact on the latents it is given ...
    assert np.array_equal(codes.cpu().numpy(), rvq_oracle.encode(lat, cb[:8]))
    # ... and fp32-rounding differences of the latents move few decisions
    agree = (codes.cpu().numpy() == rvq_oracle.encode(lat_ref, cb[:8])).mean()
    assert agree > 0.98, agree
    out = model.decode([(codes, None)])
    assert tuple(out.shape) == (B, 1, S)
    want = so.decoder(rvq_oracle.decode(codes.cpu().numpy(), cb[:8]), P, cfg)
    assert rel(out.cpu().numpy(), want) < TOL
    # encodec-0.1.1 key dialect loads to the same model
    old = {}
    for k, v in sd.items():
        k2 = k.replace(".layers.", ".model.", 1) if k.startswith(("encoder.", "decoder.")) else k
        k2 = k2.replace(".conv.parametrizations.weight.original0", ".conv.conv.weight_g").replace(
            ".conv.parametrizations.weight.original1", ".conv.conv.weight_v").replace(".conv.bias", ".conv.conv.bias")
        k2 = k2.replace("quantizer.layers.", "quantizer.vq.layers.").replace(".codebook.embed", "._codebook.embed")
        old[k2] = v
    m2 = codec.EncodecModel.encodec_model_24khz(pretrained=False, device=cuda)
    m2.load_state_dict(old)
    m2.set_target_bandwidth(6.0)
    assert torch.equal(m2.encode(torch.from_numpy(wav).to(cuda))[0][0], codes)
    # decode_codec.py keeps the codes on the host: host tensors are copied in and the result comes back to the host
    out_h = model.decode([(codes.cpu(), None)])
    assert out_h.device.type == "cpu" and torch.equal(out_h, out.cpu())
    codes_h = model.encode(torch.from_numpy(wav))[0][0]
    assert codes_h.device.type == "cpu" and torch.equal(codes_h, codes.cpu())
    from prompt_tts_b200._lib import PtError
    with pytest.raises(PtError):
        model.encode_latents(torch.from_numpy(wav))      # the layer-level entry points take device tensors only
    assert model.to("cuda") is model and model.eval() is model      # generate_code.py:15


def test_full_batch_properties(cuda):
    """The reference's data-preparation batch (32 clips x 12 s, generate_code.py:94-96) is too large for the numpy oracle; at that
    size the path is checked through properties that hold bit for bit: clips are independent (a sub-batch gives the same codes and
    latents), the 24 kHz model is causal (samples from s0 on change no latent frame before s0 / 320 and, in the decoder, frames from
    f0 on change no sample before 320 f0), and decode(encode(x)) has the input's length."""
    import os
    import torch
    from prompt_tts_b200 import codec
    B, secs = (int(v) for v in os.environ.get("PT_SN_TEST_FULL", "32,12").split(","))
    model = codec.EncodecModel(codec.CFG_24KHZ, cuda)
    model.load_state_dict(codec.random_state_dict(model.cfg, seed=4))
    model.set_target_bandwidth(6.0)
    S = 24000 * secs
    g = torch.Generator().manual_seed(0)
    wav = (torch.randn(B, 1, S, generator=g) * 0.3).to(cuda)
    lat = model.encode_latents(wav)
    codes = model.encode(wav)[0][0]
    assert tuple(lat.shape) == (B, 128, 75 * secs) and tuple(codes.shape) == (B, 8, 75 * secs)
    assert bool(torch.isfinite(lat).all()) and int(codes.min()) >= 0 and int(codes.max()) < 1024
    # independence of the clips
    lo, hi = B // 3, B // 3 + max(1, B // 4)
    assert torch.equal(model.encode_latents(wav[lo:hi].contiguous()), lat[lo:hi])
    assert torch.equal(model.encode(wav[lo:hi].contiguous())[0][0], codes[lo:hi])
    # causality of the encoder
    f0 = (75 * secs) // 2
    wav2 = wav.clone()
    wav2[:, :, 320 * f0:] += 0.1 * torch.randn(B, 1, S - 320 * f0, generator=g).to(cuda)
    lat2 = model.encode_latents(wav2)
    assert torch.equal(lat2[:, :, :f0], lat[:, :, :f0]) and not torch.equal(lat2[:, :, f0:], lat[:, :, f0:])
    # decoder: length, finiteness, causality
    out = model.decode([(codes, None)])
    assert tuple(out.shape) == (B, 1, S) and bool(torch.isfinite(out).all())
    codes2 = codes.clone()
    codes2[:, :, f0:] = (codes2[:, :, f0:] + 1) % 1024
    out2 = model.decode([(codes2, None)])
    assert torch.equal(out2[:, :, :320 * f0], out[:, :, :320 * f0]) and not torch.equal(out2[:, :, 320 * f0:], out[:, :, 320 * f0:])
