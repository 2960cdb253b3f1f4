"""Per-shape timing of every pt_gemm call in one train step (each call synchronised and event-timed on its own)."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from prompt_tts_b200 import ops
from prompt_tts_b200.models import TTSSingleSpeaker
from prompt_tts_b200.train import DenoiserTrainStep

cfg = bench.load_cfg(bench.CFG)
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = TTSSingleSpeaker(cfg).to(dev)
st = DenoiserTrainStep(model)
inp = bench.synth(cfg, bench.BATCH, bench.T_FRAMES, 1000, dev)
for _ in range(2):
    for p in model.parameters():
        p.grad = None
    st(inp["x0"], inp["noise"], inp["t"], inp["ids"], inp["mask"])
torch.cuda.synchronize()
recs = []
seen = {}
orig = ops.gemm
def timed(a, b, segs, M, N, out, **kw):
    k_total = sum(s.nk * s.nrep for s in segs)
    nz = kw.get("nz2", 1) * kw.get("nz3", 1)
    key = (M, N, k_total, nz, kw.get("splitk", 1), len(segs), a[0].kmajor, b[0].kmajor, kw.get("out_mode", 0), kw.get("residual") is not None)
    if key not in seen:
        # pure GPU time: replay a CUDA graph holding 4 copies of this launch
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(4):
                orig(a, b, segs, M, N, out, **kw)
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        seen[key] = e0.elapsed_time(e1) / 4
    orig(a, b, segs, M, N, out, **kw)
    recs.append((key, 2.0 * M * N * k_total * nz, seen[key]))
ops.gemm = timed
for p in model.parameters():
    p.grad = None
st(inp["x0"], inp["noise"], inp["t"], inp["ids"], inp["mask"])
ops.gemm = orig
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for k, f, ms in recs:
    agg[k][0] += 1; agg[k][1] += f; agg[k][2] += ms
tot_ms = sum(v[2] for v in agg.values()); tot_f = sum(v[1] for v in agg.values())
print(f"total {tot_ms:.2f} ms, {tot_f/1e12:.2f} TFLOP, {tot_f/tot_ms/1e9:.0f} TFLOP/s, {len(recs)} calls")
print("   ms     n   TF/s  us/call (M, N, K, nz, splitk, nseg, a_kmajor, b_kmajor, out_mode, residual)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][2])[:60]:
    print(f"{v[2]:7.3f} {v[0]:4d} {v[1]/v[2]/1e9:6.0f} {v[2]/v[0]*1e3:7.1f}  {k}")
