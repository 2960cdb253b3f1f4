"""One pt_gemm shape, repeated: python tools/gemm_one.py M N K a_kmajor b_kmajor out_mode block_n [reps]  (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from prompt_tts_b200 import ops
M, N, K, ak, bk, mode, bn = [int(x) for x in sys.argv[1:8]]
reps = int(sys.argv[8]) if len(sys.argv) > 8 else 3
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
A = torch.randn((M, K) if ak else (K, M), device=dev, generator=g).to(torch.bfloat16)
B = torch.randn((N, K) if bk else (K, N), device=dev, generator=g).to(torch.bfloat16)
out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16 if mode == 0 else torch.float32)
for _ in range(reps):
    ops.gemm([ops.operand(A, bool(ak))], [ops.operand(B, bool(bk))], [ops.segment(K)], M, N, out, out_mode=mode, block_n=bn)
torch.cuda.synchronize()
ref = (A.float() if ak else A.float().t()) @ (B.float().t() if bk else B.float())
print("rel err", ((out.float() / (reps if mode == 2 else 1) - ref).norm() / ref.norm()).item())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ops.gemm([ops.operand(A, bool(ak))], [ops.operand(B, bool(bk))], [ops.segment(K)], M, N, out, out_mode=mode, block_n=bn)
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 20 * 1e3
print(f"{us:.1f} us  {2.0*M*N*K/us/1e6:.0f} TFLOP/s")
