"""Per-warp event timeline of the fused attention kernels (development aid).  Needs a library built with -DPT_ATTN_TRACE
(tools/build_trace.sh builds ab/libpt_trace.so), loaded through PT_B200_LIB:
    PT_B200_LIB=ab/libpt_trace.so python tools/attn_trace.py full_d40_self fwd|dq|dkv
Prints, for CTA 0, the mean clocks between consecutive trace points per warp role (steady state) and one work item's raw timeline."""
import collections
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch  # noqa: E402

from attn_probe import CASES  # noqa: E402
from prompt_tts_b200 import _lib, ops  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "full_d40_self"
which = sys.argv[2] if len(sys.argv) > 2 else "fwd"
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
lib = _lib.lib()
lib.pt_attn_set_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
MODE = {"fwd": 0, "dq": 1, "dkv": 2}[which]
for _ in range(2):
    ops.attn_fwd(q, k, v, o, lse, H, d, d ** -0.5)
    ops.attn_bwd(q, k, v, o, lse, do, dq, dk, dv, H, d, d ** -0.5)
torch.cuda.synchronize()
NW, NE = 12, 1024
trace = torch.zeros(NW, NE, 2, dtype=torch.int64, device="cuda")


def run(fn):
    trace.zero_()
    torch.cuda.synchronize()
    lib.pt_attn_set_trace(ctypes.c_void_p(trace.data_ptr()), MODE)
    fn()
    torch.cuda.synchronize()
    lib.pt_attn_set_trace(ctypes.c_void_p(0), -1)
    return trace.cpu().numpy()


if which == "fwd":
    t = run(lambda: ops.attn_fwd(q, k, v, o, lse, H, d, d ** -0.5))
else:
    t = run(lambda: ops.attn_bwd(q, k, v, o, lse, do, dq, dk, dv, H, d, d ** -0.5))
roles = {0: "producer", 1: "stage1 g0", 11: "stage1 g1", 10: "stage2", 2: "xform g0 q2", 3: "xform g0 q3", 6: "xform g1 q2"}
for w, role in roles.items():
    ev = [(int(a), int(b)) for a, b in t[w] if b != 0]
    if not ev:
        continue
    t0 = ev[0][1]
    # mean delta for each (prev id -> id) transition, skipping the first fifth (warm-up)
    trans = collections.OrderedDict()
    for (ia, ta), (ib, tb) in zip(ev[len(ev) // 5:-1], ev[len(ev) // 5 + 1:]):
        trans.setdefault((ia, ib), []).append(tb - ta)
    print(f"== warp {w} ({role}): {len(ev)} events, span {ev[-1][1] - t0} clk")
    for (ia, ib), v_ in trans.items():
        v_ = sorted(v_)
        print(f"   {ia:2d} -> {ib:2d}: n={len(v_):4d} mean {sum(v_) / len(v_):8.0f}  median {v_[len(v_) // 2]:6d}  max {v_[-1]:6d}")
w = 2
ev = [(int(a), int(b)) for a, b in t[w] if b != 0]
print("== raw timeline warp 2, events 60..140 (id, clk since first, delta)")
for i in range(60, min(140, len(ev))):
    print(f"   {ev[i][0]:2d} {ev[i][1] - ev[0][1]:8d} {ev[i][1] - ev[i - 1][1]:6d}")
