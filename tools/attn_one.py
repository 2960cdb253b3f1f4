"""One fused-attention forward + backward at a bench shape, nothing else (for ncu captures).
    python tools/attn_one.py full_d40_self [repeats]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch  # noqa: E402

from attn_probe import CASES  # noqa: E402
from prompt_tts_b200 import ops  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "full_d40_self"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
B, H, Lq, Lk, d, same, _ = CASES[name]
g = torch.Generator(device="cuda").manual_seed(0)
C = H * d
if same:
    qkv = (torch.randn(B, Lq, 3 * C, device="cuda", generator=g)).to(torch.bfloat16)
    q, k, v = qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:]
    dqkv = torch.empty_like(qkv)
    dq, dk, dv = dqkv[:, :, :C], dqkv[:, :, C:2 * C], dqkv[:, :, 2 * C:]
else:
    q = torch.randn(B, Lq, C, device="cuda", generator=g).to(torch.bfloat16)
    kv = torch.randn(B, Lk, 2 * C, device="cuda", generator=g).to(torch.bfloat16)
    k, v = kv[:, :, :C], kv[:, :, C:]
    dq, dkv = torch.empty_like(q), torch.empty_like(kv)
    dk, dv = dkv[:, :, :C], dkv[:, :, C:]
do = torch.randn(B, Lq, C, device="cuda", generator=g).to(torch.bfloat16)
o = torch.empty(B, Lq, C, device="cuda", dtype=torch.bfloat16)
lse = torch.empty(B, H, Lq, device="cuda")
for _ in range(reps):
    ops.attn_fwd(q, k, v, o, lse, H, d, d ** -0.5)
    ops.attn_bwd(q, k, v, o, lse, do, dq, dk, dv, H, d, d ** -0.5)
torch.cuda.synchronize()
print("ok", name, float(o.float().abs().mean()))
