"""One launch of each bandwidth-bound kernel at two bench shapes (after one warm-up launch) -- the target of an
`ncu --set full -k regex:'gn_|ln_|geglu|colsum'` capture."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from prompt_tts_b200 import ops
dev = "cuda"
B = 32
for rep in range(2):
    for L, C in ((752, 320), (376, 1280)):
        x = torch.randn(B, L, C, device=dev).to(torch.bfloat16); dy = torch.randn_like(x); add = torch.randn_like(x)
        gamma = torch.randn(C, device=dev); beta = torch.randn(C, device=dev)
        dg = torch.zeros(C, device=dev); db = torch.zeros(C, device=dev)
        stats = ops.groupnorm_stats(x, 32, 1e-5)
        ops.groupnorm_apply(x, stats, gamma, beta, 32, True)
        ops.groupnorm_bwd(dy, x, stats, gamma, beta, dg, db, 32, True, dx_add=add)
        x2 = x.view(B * L, C); dy2 = dy.view(B * L, C)
        y, rs = ops.layernorm_fwd(x2, gamma, beta)
        ops.layernorm_bwd(dy2, x2, rs, gamma, dg, db, dx_add=add.view(B * L, C))
        u = torch.randn(B * L, 8 * C, device=dev).to(torch.bfloat16); dh = torch.randn(B * L, 4 * C, device=dev).to(torch.bfloat16)
        ops.geglu_fwd(u)
        ops.geglu_bwd(dh, u)
        o = torch.zeros(C, device=dev)
        ops.colsum(x2, o)
        torch.cuda.synchronize()
print("ok")
