"""GPU: pt_gemm (persistent tcgen05 GEMM family) against fp32 torch contractions, through the C ABI.

The case list is tools/gemm_probe.py's: every operand majorness (NT / NN / TN / TT), ragged M / N / K, every tile width,
Conv1d k=3 forward (bias + time shift + residual epilogue), its data gradient and weight gradient (stream-K atomic
accumulate), the batched head-strided contractions -- plus the paths added later: cluster multicast on / off must agree,
and the three conv taps of a weight gradient in ONE launch (tap = z2) must equal three launches."""
import os
import sys

import pytest
import torch
import torch.nn.functional as F

from util import ROOT, rel

sys.path.insert(0, os.path.join(ROOT, "tools"))
pytestmark = pytest.mark.gpu


def bf(x):
    return x.to(torch.bfloat16)


def test_probe_cases_match_fp32(cuda):
    import gemm_probe
    cs = gemm_probe.cases()
    worst = {}
    for name, fn in cs.items():
        e = fn()
        worst[name] = e
        tol = 1e-5 if ("f32" in name or name.startswith("tn_") or "conv3_dw" in name or "attn_qk" in name) else 4e-3
        assert e < tol, (name, e)
    assert len(worst) >= 25


@pytest.mark.parametrize("M,N,K", [(300, 512, 960), (1000, 1280, 320), (257, 256, 64)])
def test_cluster_multicast_agrees_with_single_cta(cuda, M, N, K):
    """block_n = 256 pairs CTAs into clusters that share the B tile by TMA multicast; 257 forbids it.  Odd numbers of row tiles
    leave the second CTA of the last pair with a fully out-of-bounds tile."""
    from prompt_tts_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(7)
    A, B = bf(torch.randn(M, K, device=cuda, generator=g)), bf(torch.randn(N, K, device=cuda, generator=g))
    outs = []
    for bn in (256, 257):
        o = torch.full((M, N), float("nan"), device=cuda, dtype=torch.bfloat16)
        ops.gemm([ops.operand(A, True)], [ops.operand(B, True)], [ops.segment(K)], M, N, o, block_n=bn)
        outs.append(o)
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1])
    assert rel(outs[0], A.float() @ B.float().t()) < 4e-3


@pytest.mark.parametrize("M,N,K,b_kmajor", [(1024, 512, 1088, True), (2048, 1280, 1024, True), (1152, 640, 1472, False), (4096, 320, 1032, False),
                                           (1100, 2560, 1024, True), (2048, 1280, 320, True)])
def test_cta_pair_mode(cuda, M, N, K, b_kmajor):
    """block_n = 256 with an even (or >= 8) number of row tiles and >= 16 k-iterations runs as tcgen05 CTA pairs (cta_group::2: one 256 x 256 tile per
    pair, each CTA holding half of B); block_n = 257 is the same tile width on independent CTAs.  Same k order -> identical bits.
    Ragged M (the second CTA of the last pair partly / fully out of bounds), ragged N and K, both B majornesses, fused epilogue."""
    from prompt_tts_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M + N)
    A = bf(torch.randn(M, K, device=cuda, generator=g))
    B = bf(torch.randn(N, K, device=cuda, generator=g)) if b_kmajor else bf(torch.randn(K, N, device=cuda, generator=g))
    bias = torch.randn(N, device=cuda, generator=g)
    res = bf(torch.randn(M, N, device=cuda, generator=g))
    outs = []
    for bn in (256, 257):
        o = torch.full((M, N), float("nan"), device=cuda, dtype=torch.bfloat16)
        for _ in range(3):      # several launches: barrier phases / TMEM allocation are per launch, the schedule is persistent
            ops.gemm([ops.operand(A, True)], [ops.operand(B, b_kmajor)], [ops.segment(K)], M, N, o, bias=bias, residual=res, block_n=bn)
        outs.append(o)
    torch.cuda.synchronize()
    ref = A.float() @ (B.float().t() if b_kmajor else B.float()) + bias + res.float()
    assert torch.isfinite(outs[0].float()).all()
    assert rel(outs[0], ref) < 4e-3
    assert torch.equal(outs[0], outs[1])


def test_cta_pair_mode_batched_conv(cuda):
    """The k=3 convolution as it runs at level 0 of the bench model (3 shifted segments, batch on z2, time shift + residual epilogue)
    with the 256-wide pair tiles forced, against F.conv1d."""
    from prompt_tts_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(5)
    Bn, L, Ci, Co = 4, 752, 384, 512
    x = bf(torch.randn(Bn, L, Ci, device=cuda, generator=g))
    w = bf(torch.randn(Co, Ci, 3, device=cuda, generator=g) * 0.1)
    wp = w.permute(0, 2, 1).reshape(Co, 3 * Ci).contiguous()
    bias = torch.randn(Co, device=cuda, generator=g)
    shift = torch.randn(Bn, Co, device=cuda, generator=g)
    res = bf(torch.randn(Bn, L, Co, device=cuda, generator=g))
    o = torch.full((Bn, L, Co), float("nan"), device=cuda, dtype=torch.bfloat16)
    segs = [ops.segment(Ci, a_shift=t - 1, b_k0=t * Ci) for t in range(3)]
    ops.gemm([ops.operand(x, True, batched=True)], [ops.operand(wp, True)], segs, L, Co, o, out_strides=(Co, L * Co, 0), nz2=Bn, bias=bias,
             bias_z2=shift, bias_z2_stride=Co, residual=res, res_strides=(Co, L * Co, 0), block_n=256)
    torch.cuda.synchronize()
    ref = F.conv1d(x.float().transpose(1, 2), w.float(), bias, padding=1) + shift[:, :, None] + res.float().transpose(1, 2)
    assert rel(o.float().transpose(1, 2), ref) < 4e-3


@pytest.mark.parametrize("bn", [64, 128, 160, 192, 224, 256])
def test_every_tile_width(cuda, bn):
    from prompt_tts_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(bn)
    M, N, K = 520, 960, 200          # ragged in M (4.06 tiles), N for every width, and K (3.125 k-blocks)
    A, B = bf(torch.randn(M, K, device=cuda, generator=g)), bf(torch.randn(K, N, device=cuda, generator=g))
    bias = torch.randn(N, device=cuda, generator=g)
    res = bf(torch.randn(M, N, device=cuda, generator=g))
    o = torch.full((M, N), float("nan"), device=cuda, dtype=torch.bfloat16)
    ops.gemm([ops.operand(A, True)], [ops.operand(B, False)], [ops.segment(K)], M, N, o, bias=bias, residual=res, block_n=bn)
    torch.cuda.synchronize()
    assert rel(o, A.float() @ B.float() + bias + res.float()) < 4e-3


def test_conv_weight_gradient_three_taps_one_launch(cuda):
    from prompt_tts_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    Bn, L, Ci, Co = 3, 94, 64, 128
    dy, x = bf(torch.randn(Bn, L, Co, device=cuda, generator=g)), bf(torch.randn(Bn, L, Ci, device=cuda, generator=g))
    dw = torch.zeros(Co, Ci, 3, device=cuda)
    seg = ops.segment(L, b_k0=-1, b_k0_z2=1, nrep=Bn, rep_is_batch=True)
    ops.gemm([ops.operand(dy, False, batched=True)], [ops.operand(x, False, batched=True)], [seg], Co, Ci, dw.view(Co, Ci * 3),
             out_strides=(3 * Ci, 1, 0), nz2=3, out_mode=ops.OUT_F32_ATOMIC_ADD, out_stride_n=3)
    torch.cuda.synchronize()
    xx = x.float().transpose(1, 2).requires_grad_(True)
    w = torch.zeros(Co, Ci, 3, device=cuda, requires_grad=True)
    F.conv1d(xx, w, padding=1).backward(dy.float().transpose(1, 2))
    assert rel(dw, w.grad) < 1e-5


@pytest.mark.parametrize("Ci,Co", [(8, 320), (320, 8), (16, 64)])
def test_thin_conv_on_the_gemm_path(cuda, Ci, Co):
    """conv_in (8 -> C) and conv_out (C -> 8) of the reference model (unet_1d_condition.py:193,734) run on the implicit-GEMM conv
    path with the 8-wide side zero-filled to one tile: forward, bias, weight and data gradients against F.conv1d in fp32."""
    from prompt_tts_b200 import engine as E
    g = torch.Generator(device="cuda").manual_seed(7)
    Bn, L = 3, 94
    x = bf(torch.randn(Bn, L, Ci, device=cuda, generator=g))
    w = (torch.randn(Co, Ci, 3, device=cuda, generator=g) * 0.2).requires_grad_(True)
    b = torch.randn(Co, device=cuda, generator=g).requires_grad_(True)
    dy = bf(torch.randn(Bn, L, Co, device=cuda, generator=g))
    tape = E.Tape(E.PackCache())
    xv = E.Var(x)
    y = E.conv3(tape, xv, w, b)
    y.grad, y.owned = dy, True
    tape.backward()
    torch.cuda.synchronize()
    xr = x.float().transpose(1, 2).requires_grad_(True)
    wr = bf(w.detach()).float().requires_grad_(True)      # the path contracts bf16 copies of the weights
    br = b.detach().clone().requires_grad_(True)
    yr = F.conv1d(xr, wr, br, padding=1)
    yr.backward(dy.float().transpose(1, 2))
    assert rel(y.data.float().transpose(1, 2), yr) < 5e-3
    assert rel(tape.pgrads[id(w)], wr.grad) < 1e-4 and rel(tape.pgrads[id(b)], br.grad) < 1e-4
    assert rel(xv.grad.float().transpose(1, 2), xr.grad) < 5e-3


@pytest.mark.parametrize("L,Ci,Co", [(94, 256, 384), (47, 128, 128), (64, 384, 256), (81, 128, 256)])
def test_conv_with_the_weights_on_the_tile_rows(cuda, L, Ci, Co):
    """Few rows per sample: engine.conv3 runs the convolution with M = channels, N = L and a transposed-output epilogue (tile width 96 or
    64 here) instead of 128-row tiles that are mostly padding.  Forward with bias, per-sample shift and residual, data
    gradient accumulated onto an existing gradient, weight and bias gradients -- against F.conv1d in fp32; and the two forms agree."""
    from prompt_tts_b200 import engine as E
    assert E._swap_tile(L, Co) and E._swap_tile(L, Ci)
    g = torch.Generator(device="cuda").manual_seed(L)
    Bn = 3
    x = bf(torch.randn(Bn, L, Ci, device=cuda, generator=g))
    w = (torch.randn(Co, Ci, 3, device=cuda, generator=g) * 0.1).requires_grad_(True)
    b = torch.randn(Co, device=cuda, generator=g).requires_grad_(True)
    res = bf(torch.randn(Bn, L, Co, device=cuda, generator=g))
    dy = bf(torch.randn(Bn, L, Co, device=cuda, generator=g))
    prev = bf(torch.randn(Bn, L, Ci, device=cuda, generator=g))
    outs = []
    for swap in (True, False):
        E.CONV_SWAP = swap
        try:
            tape = E.Tape(E.PackCache())
            xv, rv = E.Var(x), E.Var(res)
            xv.grad, xv.owned = prev.clone(), True
            y = E.conv3(tape, xv, w, b, residual=rv)
            y.grad, y.owned = dy, True
            tape.backward()
            torch.cuda.synchronize()
            outs.append((y.data.clone(), xv.grad.clone(), tape.pgrads[id(w)].clone(), tape.pgrads[id(b)].clone()))
        finally:
            E.CONV_SWAP = True
    xr = x.float().transpose(1, 2).requires_grad_(True)
    wr = bf(w.detach()).float().requires_grad_(True)
    br = b.detach().clone().requires_grad_(True)
    yr = F.conv1d(xr, wr, br, padding=1) + res.float().transpose(1, 2)
    yr.backward(dy.float().transpose(1, 2))
    y_sw, dx_sw, dw_sw, db_sw = outs[0]
    assert rel(y_sw.float().transpose(1, 2), yr) < 5e-3
    assert rel(dx_sw.float().transpose(1, 2), xr.grad + prev.float().transpose(1, 2)) < 6e-3
    assert rel(dw_sw, wr.grad) < 1e-4 and rel(db_sw, br.grad) < 1e-4
    # same contraction order per output element in both forms: equal up to the bf16 rounding of identical fp32 sums
    assert rel(y_sw, outs[1][0]) < 1e-5 and rel(dx_sw, outs[1][1]) < 1e-5
