"""GroupNorm forward / backward time per call at the bench shapes with COLD caches: every call of a timed graph works on its own
tensors (the set is larger than L2), so the figure is HBM-side.  Run twice to compare the two paths:
    python tools/gn_probe.py ; PT_GN_NO_CLUSTER=1 python tools/gn_probe.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from prompt_tts_b200 import ops  # noqa: E402

dev = "cuda"
B = 32
path = "streaming" if os.environ.get("PT_GN_NO_CLUSTER") else "cluster"
for L, C in ((752, 320), (752, 640), (376, 640), (376, 960), (188, 1280), (188, 1920), (94, 1280), (94, 2560)):
    nbytes = B * L * C * 2
    K = max(4, (300 << 20) // (3 * nbytes) + 1)
    xs = [torch.randn(B, L, C, device=dev).to(torch.bfloat16) for _ in range(K)]
    dys = [torch.randn(B, L, C, device=dev).to(torch.bfloat16) for _ in range(K)]
    adds = [torch.randn(B, L, C, device=dev).to(torch.bfloat16) for _ in range(K)]
    gamma, beta = torch.randn(C, device=dev), torch.randn(C, device=dev)
    dg, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    stats = [ops.groupnorm_fwd(x, gamma, beta, 32, 1e-5, True)[1] for x in xs]

    def timed(fn):
        for i in range(K):
            fn(i)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(K):
                fn(i)
        g.replay()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / K * 1e3)
        return best

    tf = timed(lambda i: ops.groupnorm_fwd(xs[i], gamma, beta, 32, 1e-5, True))
    tb = timed(lambda i: ops.groupnorm_bwd(dys[i], xs[i], stats[i], gamma, beta, dg, db, 32, True, dx_add=adds[i], out=adds[i]))
    print(f"{path:9s} L={L:4d} C={C:5d} ({nbytes / 1e6:5.1f} MB): fwd {tf:6.1f} us = {2 * nbytes / tf / 1e3:5.0f} GB/s (2 passes)   "
          f"bwd+add {tb:6.1f} us = {4 * nbytes / tb / 1e3:5.0f} GB/s (4 passes)", flush=True)
