"""In-graph kernel times of the bench step: torch.profiler (CUPTI activity records) over replays of the captured step, aggregated by
kernel name.  Unlike an ncu launch list (cold caches, serialised) these are the durations the kernels have INSIDE the step.
    python tools/step_profile.py [full]      # `full`: the step + optimiser"""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402
from prompt_tts_b200.models import TTSSingleSpeaker  # noqa: E402
from prompt_tts_b200.optim import FusedClipAdamW  # noqa: E402
from prompt_tts_b200.train import DenoiserTrainStep  # noqa: E402

full = len(sys.argv) > 1 and sys.argv[1] == "full"
cfg = bench.load_cfg(bench.CFG)
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = TTSSingleSpeaker(cfg).to(dev)
st = DenoiserTrainStep(model)
opt = FusedClipAdamW(st)
inp = bench.synth(cfg, bench.BATCH, bench.T_FRAMES, 1000, dev)


def step():
    st(inp["x0"], inp["noise"], inp["t"], inp["ids"], inp["mask"])
    if full:
        opt.step()


step()
opt.attach()
step()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    step()
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
N = 3
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(N):
        g.replay()
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
t0, t1 = None, None
for e in prof.events():
    if e.device_type is not None and str(e.device_type).endswith("CUDA") and e.device_time_total > 0:
        name = e.name
        for cut in ("(anonymous namespace)::", "<unnamed>::", "void "):
            name = name.replace(cut, "")
        name = name.split("(")[0]
        agg[name][0] += 1
        agg[name][1] += e.device_time_total
tot = sum(v[1] for v in agg.values()) / N / 1e3
print(f"sum of kernel time per step: {tot:.2f} ms over {sum(v[0] for v in agg.values()) // N} kernels")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1] / N / 1e3:8.3f} ms {100 * v[1] / N / 1e3 / tot:5.1f}%  n={v[0] // N:4d}  avg {v[1] / v[0]:7.1f} us  {k[:90]}")
