"""GPU probe for pt_attn_fwd / pt_attn_bwd against a torch fp32 reference, one subprocess per case (a trapped
kernel kills the CUDA context).  Usage:
    python tools/attn_probe.py            # all cases -> gpurun_out/attn_probe.json
    python tools/attn_probe.py NAME       # one case in-process (prints a JSON line)
"""
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

# name: (B, H, Lq, Lk, d, self_attention, time_it)
CASES = {
    "tiny_d8": (2, 8, 24, 24, 8, True, False),
    "tiny_d16_cross": (2, 8, 47, 33, 16, False, False),
    "d40_128": (1, 2, 128, 128, 40, True, False),
    "d40_200_cross": (2, 3, 200, 150, 40, False, False),
    "d64_550": (2, 12, 550, 550, 64, True, False),
    "d80_376_cross": (2, 8, 376, 550, 80, False, False),
    "d160_188": (2, 8, 188, 188, 160, True, False),
    "d160_94_cross": (2, 8, 94, 550, 160, False, False),
    "full_d40_self": (32, 8, 752, 752, 40, True, True),
    "full_d40_cross": (32, 8, 752, 550, 40, False, True),
    "full_d80_self": (32, 8, 376, 376, 80, True, True),
    "full_d80_cross": (32, 8, 376, 550, 80, False, True),
    "full_d160_self": (32, 8, 188, 188, 160, True, True),
    "full_d160_cross": (32, 8, 188, 550, 160, False, True),
    "full_text_d64": (32, 12, 550, 550, 64, True, True),
}


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def run(name):
    import torch
    from prompt_tts_b200 import ops
    B, H, Lq, Lk, d, same, timed = CASES[name]
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    C = H * d
    scale = d ** -0.5
    if same:
        qkv = (torch.randn(B, Lq, 3 * C, device=dev, generator=g) * 1.5).to(torch.bfloat16)
        q, k, v = qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:]
        dqkv = torch.full_like(qkv, float("nan"))
        dq, dk, dv = dqkv[:, :, :C], dqkv[:, :, C:2 * C], dqkv[:, :, 2 * C:]
    else:
        qb = (torch.randn(B, Lq, C, device=dev, generator=g) * 1.5).to(torch.bfloat16)
        kvb = (torch.randn(B, Lk, 2 * C, device=dev, generator=g) * 1.5).to(torch.bfloat16)
        q, k, v = qb, kvb[:, :, :C], kvb[:, :, C:]
        dq = torch.full_like(qb, float("nan"))
        dkv = torch.full_like(kvb, float("nan"))
        dk, dv = dkv[:, :, :C], dkv[:, :, C:]
    do = torch.randn(B, Lq, C, device=dev, generator=g).to(torch.bfloat16)
    o = torch.full((B, Lq, C), float("nan"), device=dev, dtype=torch.bfloat16)
    lse = torch.full((B, H, Lq), float("nan"), device=dev, dtype=torch.float32)

    ops.attn_fwd(q, k, v, o, lse, H, d, scale)
    torch.cuda.synchronize()
    ops.attn_bwd(q, k, v, o, lse, do, dq, dk, dv, H, d, scale)
    torch.cuda.synchronize()

    def heads(t, L):
        return t.float().reshape(B, L, H, d).permute(0, 2, 1, 3)

    res = {}
    chunk = max(1, min(B, (1 << 28) // (H * Lq * Lk)))   # bound the fp32 [b, H, Lq, Lk] reference
    errs = {k_: [0.0, 0.0] for k_ in ("o", "lse", "dq", "dk", "dv")}
    for b0 in range(0, B, chunk):
        sl = slice(b0, b0 + chunk)
        bb = q[sl].shape[0]
        qf = q[sl].float().reshape(bb, Lq, H, d).permute(0, 2, 1, 3).requires_grad_(True)
        kf = k[sl].float().reshape(bb, Lk, H, d).permute(0, 2, 1, 3).requires_grad_(True)
        vf = v[sl].float().reshape(bb, Lk, H, d).permute(0, 2, 1, 3).requires_grad_(True)
        s = (qf @ kf.transpose(-1, -2)) * scale
        lse_ref = torch.logsumexp(s, dim=-1)
        of = torch.softmax(s, dim=-1) @ vf
        dof = do[sl].float().reshape(bb, Lq, H, d).permute(0, 2, 1, 3)
        of.backward(dof)
        pairs = {"o": (o[sl].float().reshape(bb, Lq, H, d).permute(0, 2, 1, 3), of.detach()), "lse": (lse[sl], lse_ref.detach()),
                 "dq": (dq[sl].float().reshape(bb, Lq, H, d).permute(0, 2, 1, 3), qf.grad),
                 "dk": (dk[sl].float().reshape(bb, Lk, H, d).permute(0, 2, 1, 3), kf.grad),
                 "dv": (dv[sl].float().reshape(bb, Lk, H, d).permute(0, 2, 1, 3), vf.grad)}
        for k_, (a, r) in pairs.items():
            errs[k_][0] += (a.float() - r.float()).pow(2).sum().item()
            errs[k_][1] += r.float().pow(2).sum().item()
        del s, of, qf, kf, vf
    for k_, (n, dsum) in errs.items():
        res["err_" + k_] = (n / (dsum + 1e-30)) ** 0.5 if n == n else float("nan")
    if timed:
        def t(fn, n=10):
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n * 1e3
        res["fwd_us"] = t(lambda: ops.attn_fwd(q, k, v, o, lse, H, d, scale))
        res["bwd_us"] = t(lambda: ops.attn_bwd(q, k, v, o, lse, do, dq, dk, dv, H, d, scale))
        fl = 4.0 * B * H * Lq * Lk * d
        res["fwd_tflops"] = fl / res["fwd_us"] / 1e6
        res["bwd_tflops"] = 2.5 * fl / res["bwd_us"] / 1e6
    return res


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] in CASES:
        print(json.dumps(run(sys.argv[1])))
        sys.exit(0)
    sel = [a for a in sys.argv[1:]] or list(CASES)
    sel = [n for n in CASES if any(n.startswith(s) for s in sel)] if sys.argv[1:] else sel
    out = {}
    for name in sel:
        try:
            r = subprocess.run([sys.executable, __file__, name], capture_output=True, text=True, timeout=240)
            line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else ""
            out[name] = json.loads(line) if r.returncode == 0 and line.startswith("{") else {"rc": r.returncode, "err": r.stderr[-600:]}
        except subprocess.TimeoutExpired:
            out[name] = {"rc": "timeout"}
        print(name, json.dumps(out[name]), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/attn_probe.json", "w") as f:
        json.dump(out, f, indent=1)
