"""Repeat the fused attention forward / backward at one shape with NaN-prefilled outputs and report WHERE non-finite or
wrong values land (batch, head, 128-row block, row, column) -- the tool that root-caused the round-1 `full_d40_self` NaN.
    python tools/attn_nan_hunt.py B H Lq Lk d self|cross input_scale iterations
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from prompt_tts_b200 import ops  # noqa: E402


def main():
    B, H, Lq, Lk, d = (int(a) for a in sys.argv[1:6])
    same = sys.argv[6] == "self"
    amp = float(sys.argv[7])
    iters = int(sys.argv[8])
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    C = H * d
    scale = d ** -0.5
    if same:
        qkv = (torch.randn(B, Lq, 3 * C, device=dev, generator=g) * amp).to(torch.bfloat16)
        q, k, v = qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:]
        dqkv = torch.empty_like(qkv)
        dq, dk, dv = dqkv[:, :, :C], dqkv[:, :, C:2 * C], dqkv[:, :, 2 * C:]
        grads = [dqkv]
    else:
        qb = (torch.randn(B, Lq, C, device=dev, generator=g) * amp).to(torch.bfloat16)
        kvb = (torch.randn(B, Lk, 2 * C, device=dev, generator=g) * amp).to(torch.bfloat16)
        q, k, v = qb, kvb[:, :, :C], kvb[:, :, C:]
        dq = torch.empty_like(qb)
        dkv = torch.empty_like(kvb)
        dk, dv = dkv[:, :, :C], dkv[:, :, C:]
        grads = [dq, dkv]
    do = torch.randn(B, Lq, C, device=dev, generator=g).to(torch.bfloat16)
    o = torch.empty(B, Lq, C, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(B, H, Lq, device=dev)

    # reference O in chunks (fp32)
    ref = torch.empty(B, Lq, C, device=dev)
    for b0 in range(B):
        qf = q[b0].float().view(Lq, H, d).transpose(0, 1)
        kf = k[b0].float().view(Lk, H, d).transpose(0, 1)
        vf = v[b0].float().view(Lk, H, d).transpose(0, 1)
        ref[b0] = (torch.softmax((qf @ kf.transpose(-1, -2)) * scale, -1) @ vf).transpose(0, 1).reshape(Lq, C)

    report = {"shape": [B, H, Lq, Lk, d, same, amp], "bad_iters": 0, "events": []}
    for it in range(iters):
        o.fill_(float("nan"))
        lse.fill_(float("nan"))
        ops.attn_fwd(q, k, v, o, lse, H, d, scale)
        torch.cuda.synchronize()
        of = o.float()
        bad = ~torch.isfinite(of)
        err = (of - ref).abs()
        wrong = (err > 0.05 * ref.abs().max()) & ~bad
        nb, nw = int(bad.sum()), int(wrong.sum())
        nl = int((~torch.isfinite(lse)).sum())
        if nb or nw or nl:
            report["bad_iters"] += 1
            if len(report["events"]) < 6:
                idx = torch.nonzero(bad | wrong)[:4000].cpu()
                # summarise: (b, head, row block) triples and the distinct in-block rows / columns hit
                trip = {}
                for bb, ll, cc in idx.tolist():
                    key = (bb, cc // d, ll // 128)
                    e = trip.setdefault(key, [set(), set()])
                    e[0].add(ll % 128)
                    e[1].add(cc % d)
                summ = [{"b": kk[0], "h": kk[1], "rblk": kk[2], "work_item": (kk[0] * H + kk[1]) * ((Lq + 127) // 128) + kk[2],
                         "rows": sorted(vv[0])[:40], "nrows": len(vv[0]), "cols": sorted(vv[1])} for kk, vv in list(trip.items())[:12]]
                report["events"].append({"iter": it, "nonfinite": nb, "wrong": nw, "lse_nonfinite": nl, "where": summ})
    # backward: finite everywhere it must write
    bad_bwd = 0
    for it in range(max(1, iters // 4)):
        for t in grads:
            t.fill_(float("nan"))
        ops.attn_bwd(q, k, v, o, lse, do, dq, dk, dv, H, d, scale)
        torch.cuda.synchronize()
        if not all(bool(torch.isfinite(t.float()).all()) for t in (dq, dk, dv)):
            bad_bwd += 1
    report["bad_bwd_iters"] = bad_bwd
    report["rel_o_last"] = ((o.float() - ref).norm() / ref.norm()).item()
    print(json.dumps(report))


if __name__ == "__main__":
    main()
