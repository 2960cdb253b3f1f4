"""GPU probe for pt_gemm: every operand-majorness / segment / epilogue combination against torch fp32 matmul.
Each case runs in its own process (a trapped kernel kills the CUDA context).  Usage:
    python tools/gemm_probe.py            # all cases, JSON summary to gpurun_out/gemm_probe.json
    python tools/gemm_probe.py CASE       # one case in-process
"""
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def rel_err(a, b):
    a = a.float()
    b = b.float()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def cases():
    import torch
    import torch.nn.functional as F
    from prompt_tts_b200 import ops
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)

    def rn(*s):
        return torch.randn(*s, device=dev, generator=g).to(torch.bfloat16)

    out = {}

    def nt(M, N, K, bn=0, mode=ops.OUT_BF16):
        def f():
            A, B = rn(M, K), rn(N, K)
            o = torch.empty(M, N, device=dev, dtype=torch.bfloat16 if mode == ops.OUT_BF16 else torch.float32)
            ops.gemm([ops.operand(A, True)], [ops.operand(B, True)], [ops.segment(K)], M, N, o, out_mode=mode, block_n=bn)
            torch.cuda.synchronize()
            return rel_err(o, A.float() @ B.float().t())
        return f

    out["nt_128x64x64_bn64"] = nt(128, 64, 64, 64)
    out["nt_128x128x64_bn128"] = nt(128, 128, 64, 128)
    out["nt_256x128x256_bn128"] = nt(256, 128, 256, 128)
    out["nt_300x320x960_bn64"] = nt(300, 320, 960, 64)
    out["nt_300x320x960_bn128"] = nt(300, 320, 960, 128)
    out["nt_300x320x960_bn160"] = nt(300, 320, 960, 160)
    out["nt_300x512x960_bn256"] = nt(300, 512, 960, 256)
    out["nt_4096x1280x1280_auto_f32"] = nt(4096, 1280, 1280, 0, ops.OUT_F32)
    out["nt_k40"] = nt(200, 64, 40, 64)

    def nn_(M, N, K, bn=0):  # B stored [K, N] (MN-major)
        def f():
            A, B = rn(M, K), rn(K, N)
            o = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
            ops.gemm([ops.operand(A, True)], [ops.operand(B, False)], [ops.segment(K)], M, N, o, block_n=bn)
            torch.cuda.synchronize()
            return rel_err(o, A.float() @ B.float())
        return f

    out["nn_128x64x64_bn64"] = nn_(128, 64, 64, 64)
    out["nn_256x128x128_bn128"] = nn_(256, 128, 128, 128)
    out["nn_300x320x960_bn160"] = nn_(300, 320, 960, 160)
    out["nn_300x512x200_bn256"] = nn_(300, 512, 200, 256)

    def tn(M, N, K, bn=0, splitk=1):  # A stored [K, M], B stored [K, N]
        def f():
            A, B = rn(K, M), rn(K, N)
            o = torch.zeros(M, N, device=dev, dtype=torch.float32)
            ops.gemm([ops.operand(A, False)], [ops.operand(B, False)], [ops.segment(K)], M, N, o,
                     out_mode=ops.OUT_F32_ATOMIC_ADD if splitk > 1 else ops.OUT_F32, splitk=splitk, block_n=bn)
            torch.cuda.synchronize()
            return rel_err(o, A.float().t() @ B.float())
        return f

    out["tn_128x64x64_bn64"] = tn(128, 64, 64, 64)
    out["tn_256x128x256_bn128"] = tn(256, 128, 256, 128)
    out["tn_320x320x3000_bn128_split5"] = tn(320, 320, 3000, 128, 5)

    def tt(M, N, K, bn=0):  # A stored [K, M], B stored [N, K]
        def f():
            A, B = rn(K, M), rn(N, K)
            o = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
            ops.gemm([ops.operand(A, False)], [ops.operand(B, True)], [ops.segment(K)], M, N, o, block_n=bn)
            torch.cuda.synchronize()
            return rel_err(o, A.float().t() @ B.float().t())
        return f

    out["tt_256x128x192_bn128"] = tt(256, 128, 192, 128)

    def conv3(Bn, L, Ci, Co, bn=0):
        def f():
            x, w = rn(Bn, L, Ci), rn(Co, Ci, 3)
            bias = torch.randn(Co, device=dev, generator=g)
            tsh = torch.randn(Bn, Co, device=dev, generator=g)
            res = rn(Bn, L, Co)
            wp = w.permute(0, 2, 1).contiguous().view(Co, 3 * Ci)
            o = torch.empty(Bn, L, Co, device=dev, dtype=torch.bfloat16)
            segs = [ops.segment(Ci, a_shift=t - 1, b_k0=t * Ci) for t in range(3)]
            ops.gemm([ops.operand(x, True, batched=True)], [ops.operand(wp, True)], segs, L, Co, o,
                     out_strides=(Co, L * Co, 0), nz2=Bn, bias=bias, bias_z2=tsh, residual=res,
                     res_strides=(Co, L * Co, 0), block_n=bn)
            torch.cuda.synchronize()
            ref = F.conv1d(x.float().transpose(1, 2), w.float(), bias, padding=1) + tsh[:, :, None] + res.float().transpose(1, 2)
            return rel_err(o.float().transpose(1, 2), ref)
        return f

    out["conv3_b3_L94_64to128"] = conv3(3, 94, 64, 128)
    out["conv3_b2_L752_320to320_bn160"] = conv3(2, 752, 320, 320, 160)

    def conv3_dx(Bn, L, Ci, Co):
        def f():
            dy, w = rn(Bn, L, Co), rn(Co, Ci, 3)
            wp = w.permute(0, 2, 1).contiguous().view(Co, 3 * Ci)   # [Co, (tap, Ci)]
            dx = torch.empty(Bn, L, Ci, device=dev, dtype=torch.bfloat16)
            # dx[l, ci] = sum_t sum_co dy[l - (t-1), co] * w[co, ci, t]; B = wp viewed MN-major: N = ci (contiguous), K = co (rows)
            segs = [ops.segment(Co, a_shift=1 - t, b_shift=t * Ci) for t in range(3)]
            ops.gemm([ops.operand(dy, True, batched=True)], [ops.operand(wp, False)], segs, L, Ci, dx,
                     out_strides=(Ci, L * Ci, 0), nz2=Bn)
            torch.cuda.synchronize()
            ref = F.conv_transpose1d(dy.float().transpose(1, 2), w.float(), padding=1)
            return rel_err(dx.float().transpose(1, 2), ref)
        return f

    out["conv3_dx_b3_L94_64from128"] = conv3_dx(3, 94, 64, 128)

    def conv3_dw(Bn, L, Ci, Co, splitk):
        def f():
            dy, x = rn(Bn, L, Co), rn(Bn, L, Ci)
            dwp = torch.zeros(Co, 3 * Ci, device=dev, dtype=torch.float32)
            for t in range(3):
                seg = ops.segment(L, b_k0=t - 1, nrep=Bn, rep_is_batch=True)
                ops.gemm([ops.operand(dy, False, batched=True)], [ops.operand(x, False, batched=True)], [seg], Co, Ci,
                         dwp[:, t * Ci:], out_strides=(3 * Ci, 0, 0), out_mode=ops.OUT_F32_ATOMIC_ADD, splitk=splitk)
            torch.cuda.synchronize()
            xx = x.float().transpose(1, 2).requires_grad_(True)
            w = torch.zeros(Co, Ci, 3, device=dev, requires_grad=True)
            F.conv1d(xx, w, padding=1).backward(dy.float().transpose(1, 2))
            ref = w.grad.permute(0, 2, 1).reshape(Co, 3 * Ci)
            return rel_err(dwp, ref)
        return f

    out["conv3_dw_b3_L94_split1"] = conv3_dw(3, 94, 64, 128, 1)
    out["conv3_dw_b4_L200_split4"] = conv3_dw(4, 200, 128, 128, 4)

    def attn_qk(Bn, H, Lq, Lk, d):
        def f():
            q, k = rn(Bn, Lq, H, d), rn(Bn, Lk, H, d)
            Lkp = (Lk + 7) // 8 * 8
            S = torch.full((Bn, H, Lq, Lkp), float("nan"), device=dev, dtype=torch.float32)
            ops.gemm([ops.operand(q.permute(0, 2, 1, 3), True, batched=True)], [ops.operand(k.permute(0, 2, 1, 3), True, batched=True)],
                     [ops.segment(d)], Lq, Lk, S, out_strides=(Lkp, Lq * Lkp, H * Lq * Lkp), out_mode=ops.OUT_F32,
                     nz2=H, nz3=Bn, alpha=d ** -0.5)
            torch.cuda.synchronize()
            ref = torch.einsum("blhd,bmhd->bhlm", q.float(), k.float()) * d ** -0.5
            return rel_err(S[..., :Lk], ref)
        return f

    out["attn_qk_d40_Lk552"] = attn_qk(2, 8, 188, 552, 40)
    out["attn_qk_d160_Lk94"] = attn_qk(2, 8, 94, 94, 160)

    def attn_pv(Bn, H, Lq, Lk, d):
        def f():
            Lkp = (Lk + 7) // 8 * 8
            P = torch.full((Bn, H, Lq, Lkp), float("nan"), device=dev, dtype=torch.bfloat16)
            P[..., :Lk] = torch.softmax(torch.randn(Bn, H, Lq, Lk, device=dev, generator=g), -1).to(torch.bfloat16)
            v = rn(Bn, Lk, H, d)
            o = torch.empty(Bn, Lq, H, d, device=dev, dtype=torch.bfloat16)
            ops.gemm([ops.operand(P[..., :Lk], True, batched=True)], [ops.operand(v.permute(0, 2, 1, 3), False, batched=True)],
                     [ops.segment(Lk)], Lq, d, o, out_strides=(H * d, d, Lq * H * d), nz2=H, nz3=Bn)
            torch.cuda.synchronize()
            ref = torch.einsum("bhlm,bmhd->blhd", P[..., :Lk].float(), v.float())
            return rel_err(o, ref)
        return f

    out["attn_pv_d40_Lk550"] = attn_pv(2, 8, 188, 550, 40)
    out["attn_pv_d80_Lk376"] = attn_pv(2, 8, 376, 376, 80)
    out["attn_pv_d160_Lk94"] = attn_pv(2, 8, 94, 94, 160)
    return out


def perf():
    """Throughput of plain NT GEMMs (TFLOP/s), CUDA-event timed."""
    import torch
    from prompt_tts_b200 import ops
    res = {}
    for (M, N, K, bn) in [(24064, 320, 960, 160), (24064, 320, 960, 64), (12032, 640, 1920, 128), (6016, 1280, 3840, 256),
                          (6016, 1280, 3840, 128), (8192, 8192, 8192, 256), (24064, 2560, 320, 256), (24064, 320, 1280, 160)]:
        A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
        B = torch.randn(N, K, device="cuda").to(torch.bfloat16)
        o = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        a, b, s = [ops.operand(A, True)], [ops.operand(B, True)], [ops.segment(K)]
        for _ in range(3):
            ops.gemm(a, b, s, M, N, o, block_n=bn)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 20
        for _ in range(n):
            ops.gemm(a, b, s, M, N, o, block_n=bn)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        e0.record()
        for _ in range(n):
            torch.matmul(A, B.t(), out=o)
        e1.record()
        torch.cuda.synchronize()
        ms_t = e0.elapsed_time(e1) / n
        res[f"{M}x{N}x{K}_bn{bn}"] = {"ms": ms, "tflops": 2.0 * M * N * K / ms / 1e9, "cublas_tflops": 2.0 * M * N * K / ms_t / 1e9}
    return res


if __name__ == "__main__":
    if len(sys.argv) > 1:   # child: run cases [start:] until the first failure
        start = int(sys.argv[1])
        cs = cases()
        names = list(cs.keys())
        for i in range(start, len(names)):
            try:
                err = cs[names[i]]()
                print("RESULT " + json.dumps({"i": i, "name": names[i], "rel_err": err}), flush=True)
            except Exception as e:  # CUDA context is likely dead: let the parent restart after this case
                print("RESULT " + json.dumps({"i": i, "name": names[i], "error": str(e)[-500:]}), flush=True)
                sys.exit(3)
        try:
            print("RESULT " + json.dumps({"i": len(names), "name": "__perf__", "perf": perf()}), flush=True)
        except Exception as e:
            print("RESULT " + json.dumps({"i": len(names), "name": "__perf__", "error": str(e)[-500:]}), flush=True)
        sys.exit(0)
    summary = {}
    start, total = 0, None
    while True:
        try:
            r = subprocess.run([sys.executable, __file__, str(start)], capture_output=True, text=True, timeout=600)
            stdout = r.stdout
        except subprocess.TimeoutExpired as e:
            stdout = (e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or "")
        last = start - 1
        for l in stdout.splitlines():
            if l.startswith("RESULT "):
                d = json.loads(l[7:])
                summary[d["name"]] = {k: v for k, v in d.items() if k not in ("i", "name")}
                print(d, flush=True)
                last = d["i"]
                if d["name"] == "__perf__":
                    total = d["i"]
        if total is not None:
            break
        if last < start:   # the case at `start` died without a RESULT line (hang / trap): record and skip it
            summary[f"case_{start}"] = {"error": "no result (timeout or crash)"}
            print("case", start, "died", flush=True)
            last = start
        start = last + 1
        if start > 64:
            break
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(summary, open("gpurun_out/gemm_probe.json", "w"), indent=1)
    bad = [n for n, v in summary.items() if n != "__perf__" and not (v.get("rel_err", 1) < 1e-2)]
    print("FAILED:", bad)
