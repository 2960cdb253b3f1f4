"""One-process GPU check of the SEANet row: the parity tests of tests/test_seanet_gpu.py, then CUDA-event timings of the
encoder / decoder stacks at 24 kHz widths.  Writes gpurun_out/seanet/{pytest.log,timing.json} as it goes.

    gpurun --timeout 95 -- 'timeout -s KILL 85 python tools/seanet_check.py'
"""
import io
import json
import os
import sys
import time
from contextlib import redirect_stderr, redirect_stdout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out", "seanet")
os.makedirs(OUT, exist_ok=True)
T0 = time.time()


def probe_lstm_seq():
    """The whole-sequence LSTM launch is the one kernel with a grid-wide barrier: run its smallest test in a child process with a
    short timeout first, so that a hang costs 30 s and the rest of this script still runs (on the per-step LSTM)."""
    import subprocess
    cmd = [sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_seanet_gpu.py"), "-q", "-m", "gpu", "-p", "no:cacheprovider",
           "-k", "lstm_whole and 3-64-21"]
    try:
        r = subprocess.run(cmd, timeout=30, capture_output=True, text=True)
        ok, tail = r.returncode == 0, (r.stdout + r.stderr)[-1500:]
    except subprocess.TimeoutExpired:
        ok, tail = False, "TIMEOUT (30 s): pt_sn_lstm_seq hung"
    open(os.path.join(OUT, "lstm_seq_probe.log"), "w").write(f"ok={ok}\n{tail}\n")
    print("lstm_seq probe:", "ok" if ok else "FAILED", tail[-300:])
    if not ok:
        os.environ["PT_SN_LSTM_STEPS"] = "1"
    return ok


def run_tests(lstm_ok=True):
    import pytest
    buf = io.StringIO()
    with redirect_stdout(buf), redirect_stderr(buf):
        rc = pytest.main([os.path.join(ROOT, "tests", "test_seanet_gpu.py"), "-q", "-m", "gpu", "-p", "no:cacheprovider", "--timeout=60",
                          "-x" if "-x" in sys.argv else "--maxfail=50"] + ([] if lstm_ok else ["-k", "not lstm_whole"]))
    open(os.path.join(OUT, "pytest.log"), "w").write(buf.getvalue())
    print(buf.getvalue()[-3000:])
    return int(rc)


def timings():
    import torch
    from prompt_tts_b200 import codec
    res = {"device": torch.cuda.get_device_name(0)}
    drv = codec.CudaDriver("cuda:0")
    cfg = codec.CFG_24KHZ
    P = codec.random_state_dict(cfg, seed=1)
    lib = codec.seanet_lib()
    res["lstm_whole_sequence"] = os.environ.get("PT_SN_LSTM_STEPS", "0") != "1"
    for B, secs, fast in ((32, 12, True), (8, 4, True)) if "--quick" in sys.argv else ((32, 12, True), (32, 12, False), (8, 4, True)):
        if time.time() - T0 > 60:
            res[f"B{B}x{secs}s"] = "skipped (time)"
            break
        enc, dec = codec.SeanetStack(cfg, "encoder", drv, fast=fast), codec.SeanetStack(cfg, "decoder", drv, fast=fast)
        enc.prepare({k: drv.upload(P[k]) for k in enc.param_names()})
        dec.prepare({k: drv.upload(P[k]) for k in dec.param_names()})
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
            res[f"B{B}x{secs}s_{'fast' if fast else 'legacy'}"] = out
            json.dump(res, open(os.path.join(OUT, "timing.json"), "w"), indent=1)
        del wav
        torch.cuda.empty_cache()
    # CPU comparison: transformers' EncodecModel (the restatement the oracle is pinned to) on the host cores, 1 x 4 s
    try:
        if "--quick" in sys.argv:
            raise RuntimeError("skipped (--quick)")
        from transformers import EncodecConfig, EncodecModel
        m = EncodecModel(EncodecConfig()).eval()              # the 24 kHz architecture, random weights: timing only
        x = torch.randn(1, 1, 24000 * 4) * 0.3
        with torch.no_grad():
            t = time.time(); lat = m.encoder(x); te = time.time() - t
            t = time.time(); m.decoder(lat); td = time.time() - t
        res["cpu_transformers_1x4s"] = {"encoder_s": te, "decoder_s": td, "threads": torch.get_num_threads()}
    except Exception as e:  # noqa: BLE001
        res["cpu_transformers_1x4s"] = repr(e)
    json.dump(res, open(os.path.join(OUT, "timing.json"), "w"), indent=1)
    print(json.dumps(res))


def extras():
    """The per-kernel table of the default path, bench.py's codec leg and smoke()'s codec check."""
    import torch
    sys.argv = [sys.argv[0], "32", "12", os.path.join(OUT, "layers_fast_seq.json")]
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import seanet_profile
    seanet_profile.main()
    import bench
    line = bench.codec_throughput(torch.device("cuda:0"), torch)
    json.dump(line, open(os.path.join(OUT, "bench_codec.json"), "w"), indent=1)
    print(json.dumps(line))
    import __graft_entry__ as g
    print("smoke_codec", g.smoke_codec(torch.device("cuda:0")))
    open(os.path.join(OUT, "smoke_codec.txt"), "w").write("ok\n")


if __name__ == "__main__":
    lstm_ok = True if "--quick" in sys.argv else probe_lstm_seq()      # --quick: no child-process probe of pt_sn_lstm_seq
    rc = run_tests(lstm_ok)
    for stage in (timings, extras):
        try:
            stage()
        except Exception:  # noqa: BLE001
            import traceback
            traceback.print_exc()
            open(os.path.join(OUT, stage.__name__ + "_error.txt"), "w").write(traceback.format_exc())
    print("pytest rc", rc, "elapsed", round(time.time() - T0, 1))
    sys.exit(rc)
