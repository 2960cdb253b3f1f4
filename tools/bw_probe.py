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
    row(f"gn_fwd     L={L} C={C} (one call)", t(lambda: ops.groupnorm_fwd(x, gamma, beta, 32, 1e-5, True)), 2 * n)
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
# conv_in / conv_out (k=3 convolutions with 8 channels on one side, direct kernels doing the layout change) and the text embedding
L, C0, Cio = 752, 320, 8
xs = torch.randn(B, Cio, L, device=dev); w_in = torch.randn(C0, Cio, 3, device=dev) * 0.1; b_in = torch.randn(C0, device=dev)
h0 = torch.empty(B, L, C0, device=dev, dtype=torch.bfloat16); dh0 = torch.randn(B, L, C0, device=dev).to(torch.bfloat16)
dw_in = torch.zeros_like(w_in); db_in = torch.zeros_like(b_in)
w_out = torch.randn(Cio, C0, 3, device=dev) * 0.1; b_out = torch.randn(Cio, device=dev); yo = torch.empty(B, Cio, L, device=dev); go = torch.randn(B, Cio, L, device=dev)
dw_out = torch.zeros_like(w_out); db_out = torch.zeros_like(b_out); dh = torch.empty_like(h0)
P, S = ops._p, ops._stream
row("conv_in_fwd  32x8x752 -> 320", t(lambda: ops.call("conv_in_fwd", P(xs), P(w_in), P(b_in), P(h0), B, Cio, L, C0, S())), h0.numel() * 2 + xs.numel() * 4)
row("conv_in_bwd  (dw, dbias)", t(lambda: ops.call("conv_in_bwd", P(dh0), P(xs), P(dw_in), P(db_in), B, Cio, L, C0, S())), h0.numel() * 2 + xs.numel() * 4)
row("conv_out_fwd 320 -> 32x8x752", t(lambda: ops.call("conv_out_fwd", P(dh0), P(w_out), P(b_out), P(yo), B, C0, L, Cio, S())), h0.numel() * 2 + xs.numel() * 4)
row("conv_out_bwd (dh, dw, dbias)", t(lambda: ops.call("conv_out_bwd", P(go), P(dh0), P(w_out), P(dh), P(dw_out), P(db_out), B, C0, L, Cio, S())), 2 * h0.numel() * 2 + xs.numel() * 4)
V, D, Lt = 150, 768, 550
ids = torch.randint(1, V, (B, Lt), device=dev, dtype=torch.int32); ids[:, 400:] = 0
E_ = torch.randn(V, D, device=dev); pe = torch.randn(Lt, D, device=dev); xe = torch.empty(B, Lt, D, device=dev, dtype=torch.bfloat16)
dxe = torch.randn(B, Lt, D, device=dev).to(torch.bfloat16); dE = torch.zeros_like(E_)
row("text_embed_fwd 32x550x768", t(lambda: ops.call("text_embed_fwd", P(ids), P(E_), P(pe), P(xe), B, Lt, D, V, S())), xe.numel() * 2)
row("text_embed_bwd", t(lambda: ops.call("text_embed_bwd", P(ids), P(dxe), P(dE), B, Lt, D, V, S())), xe.numel() * 2)
for (Bc, Lc, Ca, Cb) in ((32, 752, 320, 320), (32, 376, 640, 640), (32, 188, 1280, 1280)):
    a = torch.randn(Bc, Lc, Ca, device=dev).to(torch.bfloat16); o = torch.empty(Bc, Lc, Ca + Cb, device=dev, dtype=torch.bfloat16)
    row(f"copy2d (concat half) L={Lc} C={Ca}", t(lambda: ops.copy2d(a, Ca, o, Ca + Cb, Bc * Lc, Ca)), 2 * a.numel() * 2)
    b2 = torch.randn_like(a)
    row(f"add_bf16 L={Lc} C={Ca}", t(lambda: ops.add_(a, b2)), 3 * a.numel() * 2)
