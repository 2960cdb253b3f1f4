"""Eager fwd+bwd timing of the full 1d_config on one B200 (no CUDA graph): quick look at where time goes."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from util import load_cfg, synth_inputs
from prompt_tts_b200.models import TTSSingleSpeaker

cfg_name = sys.argv[1] if len(sys.argv) > 1 else "1d_config"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
T = int(sys.argv[3]) if len(sys.argv) > 3 else 752
cfg = load_cfg(cfg_name)
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = TTSSingleSpeaker(cfg).to(dev)
inp = synth_inputs(cfg, B, T, seed=1, device=dev)
for p in model.parameters():
    p.grad = None

def step():
    out = model(inp["x0"], inp["t"], inp["ids"], inp["mask"]).sample
    loss = torch.nn.functional.mse_loss(out, inp["noise"])
    loss.backward()
    return loss

for i in range(3):
    t0 = time.time()
    l = step()
    torch.cuda.synchronize()
    print("warm", i, time.time() - t0, float(l), flush=True)
    for p in model.parameters():
        p.grad = None
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 5
t0 = time.time()
e0.record()
for _ in range(n):
    step()
    for p in model.parameters():
        p.grad = None
e1.record()
torch.cuda.synchronize()
print(json.dumps({"cfg": cfg_name, "B": B, "T": T, "ms_per_step_gpu": e0.elapsed_time(e1) / n, "ms_per_step_wall": (time.time() - t0) / n * 1e3,
                  "frames_per_s": B * T / (e0.elapsed_time(e1) / n / 1e3), "max_mem_gb": torch.cuda.max_memory_allocated() / 2**30}))
