"""Achieved HBM bandwidth of the bandwidth-bound kernels at the bench shapes (algorithmic bytes / CUDA-event time)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from prompt_tts_b200 import ops
dev = "cuda"
B = 32

def t(fn, n=10):
    """GPU time per call with launch overhead removed: replay a CUDA graph holding n copies."""
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def row(name, us, nbytes):
    print(f"{name:44s} {us:8.1f} us  {nbytes/us/1e3:7.0f} GB/s  ({nbytes/1e6:.0f} MB)")

for L, C in ((752, 320), (752, 640), (752, 960), (376, 640), (376, 1280), (376, 1920), (188, 1280), (188, 2560), (94, 1280), (94, 2560)):
    x = torch.randn(B, L, C, device=dev).to(torch.bfloat16); dy = torch.randn_like(x)
    gamma = torch.randn(C, device=dev); beta = torch.randn(C, device=dev)
    dg = torch.zeros(C, device=dev); db = torch.zeros(C, device=dev)
    n = x.numel() * 2
    stats = ops.groupnorm_stats(x, 32, 1e-5)
    row(f"gn_stats   L={L} C={C}", t(lambda: ops.groupnorm_stats(x, 32, 1e-5)), n)
    row(f"gn_apply   L={L} C={C}", t(lambda: ops.groupnorm_apply(x, stats, gamma, beta, 32, True)), 2 * n)
    row(f"gn_bwd     L={L} C={C} (reduce+apply)", t(lambda: ops.groupnorm_bwd(dy, x, stats, gamma, beta, dg, db, 32, True)), 5 * n)
    add = torch.randn_like(x)
    row(f"gn_bwd+add L={L} C={C} (reduce+apply)", t(lambda: ops.groupnorm_bwd(dy, x, stats, gamma, beta, dg, db, 32, True, dx_add=add, out=add)), 6 * n)
for M, C in ((24064, 320), (12032, 640), (6016, 1280), (17600, 768)):
    x = torch.randn(M, C, device=dev).to(torch.bfloat16); dy = torch.randn_like(x); add = torch.randn_like(x)
    gamma = torch.randn(C, device=dev); beta = torch.randn(C, device=dev)
    dg = torch.zeros(C, device=dev); db = torch.zeros(C, device=dev)
    n = x.numel() * 2
    y, rs = ops.layernorm_fwd(x, gamma, beta)
    row(f"ln_fwd     M={M} C={C}", t(lambda: ops.layernorm_fwd(x, gamma, beta)), 2 * n)
    row(f"ln_bwd+add M={M} C={C}", t(lambda: ops.layernorm_bwd(dy, x, rs, gamma, dg, db, dx_add=add)), 4 * n)
for M, F in ((24064, 1280), (12032, 2560), (6016, 5120), (17600, 3072)):
    u = torch.randn(M, 2 * F, device=dev).to(torch.bfloat16); dy = torch.randn(M, F, device=dev).to(torch.bfloat16)
    row(f"geglu_fwd  M={M} F={F}", t(lambda: ops.geglu_fwd(u)), M * F * 6)
    row(f"geglu_bwd  M={M} F={F}", t(lambda: ops.geglu_bwd(dy, u)), M * F * 10)
for M, C in ((24064, 320), (24064, 2560), (6016, 1280), (6016, 10240)):
    x = torch.randn(M, C, device=dev).to(torch.bfloat16); o = torch.zeros(C, device=dev)
    row(f"colsum     M={M} C={C}", t(lambda: ops.colsum(x, o)), M * C * 2)
