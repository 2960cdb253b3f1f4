"""Diagnostic for csrc/rvq_tc.cu: dump the tensor-core scores of stage 0 for the first 128 frames and compare them with
2 * dot(bf16(r), bf16(e)) - |e|^2 computed by torch, and the final codes with the exhaustive kernel."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from prompt_tts_b200 import ops  # noqa: E402
from prompt_tts_b200.ops import _p, call  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(0)
Q, K = 2, 1024
cb = torch.randn(Q, K, 128, device="cuda", generator=g)
lat = torch.randn(1, 128, 200, device="cuda", generator=g)
dbg = torch.full((128, K), float("nan"), device="cuda")
call("rvq_tc_debug_scores", _p(dbg))
a = ops.rvq_encode(lat, cb)
torch.cuda.synchronize()
call("rvq_tc_debug_scores", None)
b = ops.rvq_encode(lat, cb, exhaustive=True)
r = lat[0, :, :128].t().contiguous()                     # [128 frames, 128 d]
ref = 2 * (r.to(torch.bfloat16).float() @ cb[0].to(torch.bfloat16).float().t()) - (cb[0] ** 2).sum(-1)[None]
err = (dbg - ref).abs()
print("scores: nan", int(torch.isnan(dbg).sum()), "max abs err", float(err.nan_to_num(1e9).max()), "ref abs mean", float(ref.abs().mean()))
print("per 32-col chunk max err:", [round(float(err[:, c * 32:(c + 1) * 32].nan_to_num(1e9).max()), 3) for c in range(0, 32, 1)])
print("per 32-row max err:", [round(float(err[q * 32:(q + 1) * 32].nan_to_num(1e9).max()), 3) for q in range(4)])
print("dbg[0,:8]", dbg[0, :8].tolist())
print("ref[0,:8]", ref[0, :8].tolist())
print("codes equal:", bool(torch.equal(a, b)), "mismatching frames stage0:", int((a[:, 0] != b[:, 0]).sum()), "of", a.shape[2], "overflow", ops.rvq_encode.last_prepared.overflow_frames())
