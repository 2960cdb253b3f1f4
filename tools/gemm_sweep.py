"""Times every distinct pt_gemm call of one train step with each tile width forced (block_n) and with the library's own
choice; prints where the heuristic loses.  Used to calibrate the tile-width cost model in csrc/gemm_tcgen05.cu."""
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
seen = {}
count = collections.Counter()
orig = ops.gemm
BNS = (0, 64, 128, 160, 192, 224, 256, 257)

def time_one(a, b, segs, M, N, out, kw):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(4):
            orig(a, b, segs, M, N, out, **kw)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 4 * 1e3

def timed(a, b, segs, M, N, out, **kw):
    k_total = sum(s.nk * s.nrep for s in segs)
    nz = kw.get("nz2", 1) * kw.get("nz3", 1)
    key = (M, N, k_total, nz, len(segs), a[0].kmajor, b[0].kmajor, kw.get("out_mode", 0), kw.get("residual") is not None)
    count[key] += 1
    if key not in seen:
        torch.cuda.synchronize()
        res = {}
        for bn in BNS:
            kw2 = dict(kw); kw2["block_n"] = bn
            try:
                res[bn] = time_one(a, b, segs, M, N, out, kw2)
            except Exception as e:
                res[bn] = float("inf")
        seen[key] = res
    orig(a, b, segs, M, N, out, **kw)
ops.gemm = timed
for p in model.parameters():
    p.grad = None
st(inp["x0"], inp["noise"], inp["t"], inp["ids"], inp["mask"])
ops.gemm = orig
tot_auto = sum(seen[k][0] * n for k, n in count.items())
tot_best = sum(min(seen[k].values()) * n for k, n in count.items())
print(f"auto {tot_auto/1e3:.2f} ms   per-shape best {tot_best/1e3:.2f} ms   ({len(seen)} shapes, {sum(count.values())} calls)")
print("  n   auto    b64   b128   b160   b192   b224   b256 b256nc  best  lost_us  (M, N, K, nz, nseg, a_kmajor, b_kmajor, out_mode, residual)")
rows = []
for k, n in count.items():
    r = seen[k]
    best = min(r, key=lambda b_: r[b_] if b_ else float("inf"))
    rows.append(((r[0] - r[best]) * n, n, r, best, k))
for lost, n, r, best, k in sorted(rows, key=lambda t: -t[1] * t[2][0])[:70]:
    fl = 2.0 * k[0] * k[1] * k[2] * k[3]
    print(f"{n:3d} {r[0]:6.1f} {r[64]:6.1f} {r[128]:6.1f} {r[160]:6.1f} {r[192]:6.1f} {r[224]:6.1f} {r[256]:6.1f} {r[257]:6.1f}  {best:4d} {lost:7.1f}  {fl/min(r.values())/1e6:5.0f}TF {k}")
