"""One-process GPU check of the SEANet row: the parity tests of tests/test_seanet_gpu.py, then CUDA-event timings of the
encoder / decoder stacks at 24 kHz widths.  Writes gpurun_out/seanet/{pytest.log,timing.json} as it goes.

    gpurun --timeout 170 -- 'timeout 160 python tools/seanet_check.py'
"""
import io
import json
import os
import sys
import time
from contextlib import redirect_stderr, redirect_stdout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
OUT = os.path.join(ROOT, "gpurun_out", "seanet")
os.makedirs(OUT, exist_ok=True)
T0 = time.time()


def run_tests():
    import pytest
    buf = io.StringIO()
    with redirect_stdout(buf), redirect_stderr(buf):
        rc = pytest.main([os.path.join(ROOT, "tests", "test_seanet_gpu.py"), "-q", "-m", "gpu", "-p", "no:cacheprovider", "--timeout=60",
                          "-x" if "-x" in sys.argv else "--maxfail=50"])
    open(os.path.join(OUT, "pytest.log"), "w").write(buf.getvalue())
    print(buf.getvalue()[-3000:])
    return int(rc)


def timings():
    import numpy as np
    import torch
    import seanet_oracle as so
    from prompt_tts_b200 import codec
    res = {"device": torch.cuda.get_device_name(0)}
    drv = codec.CudaDriver("cuda:0")
    cfg = so.CFG_24KHZ
    P = so.make_weights(cfg, 1)
    enc, dec = codec.SeanetStack(cfg, "encoder", drv), codec.SeanetStack(cfg, "decoder", drv)
    enc.prepare({k: drv.upload(P[k]) for k in enc.param_names()})
    dec.prepare({k: drv.upload(P[k]) for k in dec.param_names()})
    lib = codec.seanet_lib()
    for B, secs in ((8, 4), (32, 12)):
        if time.time() - T0 > 95:
            res[f"B{B}x{secs}s"] = "skipped (time)"
            break
        S = 24000 * secs
        wav = torch.randn(B, 1, S, device="cuda") * 0.3
        out = {}
        for name, stack, x, L in (("encoder", enc, wav, S), ("decoder", dec, None, S // 320)):
            if x is None:
                x = torch.randn(B, cfg["hidden_size"], L, device="cuda")
            stack.forward(x, B, L)                              # warm-up
            torch.cuda.synchronize()
            n0 = lib.pt_sn_launch_count()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            y, _ = stack.forward(x, B, L)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            out[name] = {"ms": ms, "launches": int(lib.pt_sn_launch_count() - n0), "audio_s_per_s": B * secs / (ms / 1e3),
                         "finite": bool(torch.isfinite(y).all().item())}
            res[f"B{B}x{secs}s"] = out
            json.dump(res, open(os.path.join(OUT, "timing.json"), "w"), indent=1)
        del wav
        torch.cuda.empty_cache()
    # CPU comparison: transformers' EncodecModel (the restatement the oracle is pinned to) on the host cores, 1 x 4 s
    try:
        m = so.to_transformers_model(P, cfg)
        x = torch.randn(1, 1, 24000 * 4) * 0.3
        with torch.no_grad():
            t = time.time(); lat = m.encoder(x); te = time.time() - t
            t = time.time(); m.decoder(lat); td = time.time() - t
        res["cpu_transformers_1x4s"] = {"encoder_s": te, "decoder_s": td, "threads": torch.get_num_threads()}
    except Exception as e:  # noqa: BLE001
        res["cpu_transformers_1x4s"] = repr(e)
    json.dump(res, open(os.path.join(OUT, "timing.json"), "w"), indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    rc = run_tests()
    try:
        timings()
    except Exception as e:  # noqa: BLE001
        import traceback
        traceback.print_exc()
        open(os.path.join(OUT, "timing_error.txt"), "w").write(traceback.format_exc())
    print("pytest rc", rc, "elapsed", round(time.time() - T0, 1))
    sys.exit(rc)
