"""Per-kernel time of the SEANet stacks on the GPU: CUDA events around every C-ABI call of one encoder + one decoder pass
(B x seconds of 24 kHz audio), aggregated by kernel and shape, with the algorithmic FLOPs of each.  A diagnostic (event pairs
around single launches serialise the stream), not a bench number.

    python tools/seanet_profile.py [B] [seconds] [out.json]        # default 32 12 gpurun_out/seanet/layers.json
    PT_SN_PROFILE_PLAIN=1 ... : no events, one pass of each stack (the target of an ncu launch list)
"""
import ctypes as C
import json
import os
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from prompt_tts_b200 import codec  # noqa: E402


class TimingDriver(codec.CudaDriver):
    def __init__(self, device="cuda:0"):
        super().__init__(device)
        self.records = []
        self.on = False

    def call(self, name, *args):
        if not self.on:
            return super().call(name, *args)
        key, flops = name, 0.0
        if name in ("conv1d", "conv_transpose1d", "conv1d_packed", "conv_transpose1d_packed"):
            d = C.cast(args[0], C.POINTER(codec.ConvDesc)).contents
            taps = d.K if name.startswith("conv1d") else -(-d.K // d.stride)
            flops = 2.0 * d.B * d.Ci * d.Co * taps * d.Lout
            key = f"{name} {d.Ci}->{d.Co} k{d.K} s{d.stride} d{d.dil} L{d.Lout}"
        elif name == "linear_rows":
            R, Kd, N = args[4], args[5], args[6]
            flops = 2.0 * R * Kd * N
            key = f"{name} {R}x{Kd}x{N}"
        elif name == "lstm_step":
            B, H = args[5], args[6]
            flops = 2.0 * B * H * 4 * H
            key = f"{name} B{B} H{H}"
        elif name == "lstm_seq":
            T, B, H = args[4], args[5], args[6]
            flops = 2.0 * T * B * H * 4 * H
            key = f"{name} T{T} B{B} H{H}"
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        super().call(name, *args)
        e1.record()
        self.records.append((key, flops, e0, e1))


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    secs = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    out = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "gpurun_out", "seanet", "layers.json")
    cfg = codec.CFG_24KHZ
    drv = TimingDriver()
    P = codec.random_state_dict(cfg, seed=1)
    enc, dec = codec.SeanetStack(cfg, "encoder", drv), codec.SeanetStack(cfg, "decoder", drv)
    enc.prepare({k: drv.upload(P[k]) for k in enc.param_names()})
    dec.prepare({k: drv.upload(P[k]) for k in dec.param_names()})
    S = 24000 * secs
    wav = torch.randn(B, 1, S, device="cuda") * 0.3
    lat = torch.randn(B, cfg["hidden_size"], S // 320, device="cuda")
    if os.environ.get("PT_SN_PROFILE_PLAIN"):              # one cold pass of each stack: what an ncu launch list captures
        enc.forward(wav, B, S)
        dec.forward(lat, B, S // 320)
        torch.cuda.synchronize()
        return
    enc.forward(wav, B, S)
    dec.forward(lat, B, S // 320)
    torch.cuda.synchronize()
    res = {"B": B, "seconds": secs, "device": torch.cuda.get_device_name(0)}
    for side, stack, x, L in (("encoder", enc, wav, S), ("decoder", dec, lat, S // 320)):
        drv.records, drv.on = [], True
        stack.forward(x, B, L)
        torch.cuda.synchronize()
        drv.on = False
        agg = OrderedDict()
        for key, flops, e0, e1 in drv.records:
            a = agg.setdefault(key, {"launches": 0, "ms": 0.0, "gflop": 0.0})
            a["launches"] += 1
            a["ms"] += e0.elapsed_time(e1)
            a["gflop"] += flops / 1e9
        for a in agg.values():
            a["tflops"] = a["gflop"] / a["ms"] if a["ms"] > 0 else 0.0
        res[side] = {"total_ms": sum(a["ms"] for a in agg.values()), "kernels": agg}
    os.makedirs(os.path.dirname(out), exist_ok=True)
    json.dump(res, open(out, "w"), indent=1)
    for side in ("encoder", "decoder"):
        print(side, round(res[side]["total_ms"], 2), "ms")
        for k, a in res[side]["kernels"].items():
            print(f"  {k:48s} x{a['launches']:5d} {a['ms']:9.3f} ms {a['tflops']:7.2f} TFLOP/s")


if __name__ == "__main__":
    main()
