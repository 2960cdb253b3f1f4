"""Two launches of the tensor-core RVQ quantiser at 256 clips x 900 frames (for ncu captures)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from prompt_tts_b200 import ops  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(0)
cb = torch.randn(8, 1024, 128, device="cuda", generator=g)
lat = torch.randn(int(sys.argv[1]) if len(sys.argv) > 1 else 256, 128, 900, device="cuda", generator=g)
prep = ops.rvq_prepare(cb)
for _ in range(2):
    c = ops.rvq_encode(lat, cb, prepared=prep)
torch.cuda.synchronize()
print("ok", int(c.sum()))
