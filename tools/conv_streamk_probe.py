"""Small-L convolutions (levels 2 / 3 of the bench model: 188 / 94 rows per sample): the data-parallel bf16 launch (160 tiles of 128 x 256 on
148 SMs: two waves, the second almost empty) against a stream-K launch into an fp32 scratch (+ memset + a finishing pass).
    python tools/conv_streamk_probe.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from prompt_tts_b200 import ops  # noqa: E402

dev = "cuda"
B = 32


def timed(fn, n=8):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n * 1e3)
    return best


for L, Ci, Co in ((94, 1280, 1280), (94, 2560, 1280), (94, 1280, 2560), (188, 1280, 1280), (188, 2560, 1280), (188, 640, 1280)):
    x = torch.randn(B, L, Ci, device=dev).to(torch.bfloat16)
    wp = (torch.randn(Co, 3 * Ci, device=dev) * 0.02).to(torch.bfloat16)
    bias = torch.randn(Co, device=dev)
    res = torch.randn(B, L, Co, device=dev).to(torch.bfloat16)
    out = torch.empty(B, L, Co, device=dev, dtype=torch.bfloat16)
    acc = torch.zeros(B, L, Co, device=dev)
    segs = [ops.segment(Ci, a_shift=t - 1, b_k0=t * Ci) for t in range(3)]
    a_ops, b_ops = [ops.operand(x, True, batched=True)], [ops.operand(wp, True)]

    def dp():
        ops.gemm(a_ops, b_ops, segs, L, Co, out, out_strides=(Co, L * Co, 0), nz2=B, bias=bias, residual=res, res_strides=(Co, L * Co, 0))

    def sk():
        acc.zero_()
        ops.gemm(a_ops, b_ops, segs, L, Co, acc, out_strides=(Co, L * Co, 0), nz2=B, out_mode=ops.OUT_F32_ATOMIC_ADD)

    def sk_only():
        ops.gemm(a_ops, b_ops, segs, L, Co, acc, out_strides=(Co, L * Co, 0), nz2=B, out_mode=ops.OUT_F32_ATOMIC_ADD)

    sw_bn = 96 if L <= 96 else 192
    segs_w = [ops.segment(Ci, a_k0=t * Ci, b_shift=t - 1) for t in range(3)]
    out_t = torch.empty_like(out)

    def swapped(bias_=bias, res_=res):
        ops.gemm(b_ops, a_ops, segs_w, Co, L, out_t, out_strides=(Co, L * Co, 0), nz2=B, bias=bias_, residual=res_, res_strides=(Co, L * Co, 0),
                 block_n=sw_bn, out_transposed=True)

    t_sw = timed(swapped)
    t_sw_plain = timed(lambda: swapped(None, None))
    t_dp_plain = timed(lambda: ops.gemm(a_ops, b_ops, segs, L, Co, out, out_strides=(Co, L * Co, 0), nz2=B))
    print(f"   weights on rows (block_n {sw_bn}): {t_sw:6.1f} us ({2.0 * B * L * Co * 3 * Ci / t_sw / 1e6:5.0f} TF), without bias/residual {t_sw_plain:6.1f} us;"
          f" rows of x on rows without bias/residual {t_dp_plain:6.1f} us")
    t_256 = timed(lambda: ops.gemm(a_ops, b_ops, segs, L, Co, out, out_strides=(Co, L * Co, 0), nz2=B, bias=bias, residual=res,
                                   res_strides=(Co, L * Co, 0), block_n=256))
    o1 = out.clone()
    print(f"   rows of x on rows, block_n 256 forced (pairs along the batch axis unless PT_GEMM_NO_PAIR_Z2): {t_256:6.1f} us "
          f"({2.0 * B * L * Co * 3 * Ci / t_256 / 1e6:5.0f} TF)")
    t_dp, t_sk, t_sk0 = timed(dp), timed(sk), timed(sk_only)
    dp()
    acc.zero_()
    sk_only()
    torch.cuda.synchronize()
    ref = acc + bias + res.float()
    err = ((out.float() - ref).norm() / ref.norm()).item()
    assert torch.equal(o1, out) or ((o1.float() - out.float()).norm() / out.float().norm()).item() < 1e-5, "block_n 256 result differs" 
    fl = 2.0 * B * L * Co * 3 * Ci
    print(f"L={L:4d} Ci={Ci:5d} Co={Co:5d}: data-parallel bf16 {t_dp:6.1f} us ({fl / t_dp / 1e6:5.0f} TF)   stream-K fp32 {t_sk0:6.1f} us "
          f"(+ memset {t_sk - t_sk0:4.1f} us; a finishing pass would add ~{(acc.numel() * 6 + res.numel() * 2) / 5e6:4.1f} us)   rel diff {err:.1e}", flush=True)
