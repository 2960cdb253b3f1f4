"""Time one DDPMSampler.sample() call at the bench's sampling configuration (B = 64, T = 1504, 100 steps, 225-frame prompt)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from prompt_tts_b200.models import TTSSingleSpeaker  # noqa: E402
from prompt_tts_b200.sample import DDPMSampler  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
cfg = bench.load_cfg(bench.CFG)
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = TTSSingleSpeaker(cfg).to(dev).eval()
inp = bench.synth(cfg, 64, 1504, 4000, dev)
prompt = inp["x0"][..., :225].contiguous()
DDPMSampler(model, n_infer=4).sample(inp["ids"], 1504, prompt=prompt, seed=1)
smp = DDPMSampler(model, n_infer=steps)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
x = smp.sample(inp["ids"], 1504, prompt=prompt, seed=0)
e1.record()
torch.cuda.synchronize()
print(f"sample(): {e0.elapsed_time(e1) / 1e3:.3f} s for {steps} steps, finite {bool(torch.isfinite(x).all())}")
