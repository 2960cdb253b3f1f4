"""Runs N eager train steps of the bench workload (for ncu launch lists / single-kernel captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from prompt_tts_b200.models import TTSSingleSpeaker
from prompt_tts_b200.train import DenoiserTrainStep
from prompt_tts_b200 import _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
cfg = bench.load_cfg(bench.CFG)
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = TTSSingleSpeaker(cfg).to(dev)
st = DenoiserTrainStep(model)
inp = bench.synth(cfg, bench.BATCH, bench.T_FRAMES, 1000, dev)
for i in range(n):
    for p in model.parameters():
        p.grad = None
    l0 = _lib.lib().pt_launch_count()
    loss = st(inp["x0"], inp["noise"], inp["t"], inp["ids"], inp["mask"])
    torch.cuda.synchronize()
    print("step", i, float(loss), "launches", _lib.lib().pt_launch_count() - l0, flush=True)
