"""One RVQ encode + decode (both decode kernels) at the reference batch shape, nothing else (for ncu captures)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from prompt_tts_b200 import ops  # noqa: E402

bs = int(sys.argv[1]) if len(sys.argv) > 1 else 32
g = torch.Generator(device="cuda").manual_seed(0)
cb = torch.randn(8, 1024, 128, device="cuda", generator=g)
lat = torch.randn(bs, 128, 900, device="cuda", generator=g)
for _ in range(2):
    codes = ops.rvq_encode(lat, cb)
    big = codes.repeat(max(1, 512 // bs), 1, 1)
    a = ops.rvq_decode(big, cb)
    b = ops.rvq_decode(big, cb, gather_l2=True)
torch.cuda.synchronize()
print("ok", bool(torch.equal(a, b)))
