"""Thin Python wrappers over the C ABI (include/prompt_tts_b200.h).  Tensors are torch CUDA tensors used
only as device-memory handles; every computation below is a hand-written sm_100a kernel."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import Gemm, Operand, Segment, OUT_BF16, OUT_F32, OUT_F32_ATOMIC_ADD, call

BF16 = torch.bfloat16
F32 = torch.float32


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def _need(t: torch.Tensor, dtype, what: str):
    if not t.is_cuda:
        raise _lib.PtError(f"{what}: tensor is not on a CUDA device (no CPU fallback exists)")
    if t.dtype != dtype:
        raise _lib.PtError(f"{what}: expected {dtype}, got {t.dtype}")


# ------------------------------------------------------------------------------------------------ GEMM
def operand(t: torch.Tensor, kmajor: bool, batched: bool = False) -> Operand:
    """`t` is a <=4-D bf16 view in torch order [d3, d2, d1, d0] with d0 contiguous."""
    _need(t, BF16, "gemm operand")
    assert t.dim() <= 4 and t.stride(-1) == 1, (t.shape, t.stride())
    op = Operand()
    op.ptr = t.data_ptr()
    nd = t.dim()
    for i in range(4):
        if i < nd:
            op.dim[i] = t.shape[nd - 1 - i]
            op.stride[i] = t.stride(nd - 1 - i)
        else:
            op.dim[i] = 1
            op.stride[i] = 0
    op.kmajor = 1 if kmajor else 0
    op.batched = 1 if batched else 0
    return op


def segment(nk, a_idx=0, b_idx=0, a_k0=0, b_k0=0, a_shift=0, b_shift=0, nrep=1, rep_is_batch=False, rep_c2_0=0) -> Segment:
    s = Segment()
    s.a_idx, s.b_idx, s.a_k0, s.b_k0 = a_idx, b_idx, a_k0, b_k0
    s.a_mn_shift, s.b_mn_shift, s.nk, s.nrep = a_shift, b_shift, nk, nrep
    s.rep_is_batch, s.rep_c2_0 = (1 if rep_is_batch else 0), rep_c2_0
    return s


def gemm(a: Sequence[Operand], b: Sequence[Operand], segs: Sequence[Segment], M: int, N: int, out: torch.Tensor,
         out_strides=(None, 0, 0), out_mode: int = OUT_BF16, nz2: int = 1, nz3: int = 1, splitk: int = 1,
         alpha: float = 1.0, bias: Optional[torch.Tensor] = None, bias_z2: Optional[torch.Tensor] = None,
         residual: Optional[torch.Tensor] = None, res_strides=(None, 0, 0), block_n: int = 0) -> None:
    """out[z2, z3, m, n] (element strides out_strides = (m, z2, z3)) = epilogue(sum over segments)."""
    g = Gemm()
    for i, o in enumerate(a):
        g.a[i] = o
    for i, o in enumerate(b):
        g.b[i] = o
    for i, s in enumerate(segs):
        g.seg[i] = s
    g.nseg = len(segs)
    g.M, g.N, g.nz2, g.nz3, g.splitk, g.block_n = M, N, nz2, nz3, splitk, block_n
    g.out = out.data_ptr()
    g.out_dtype = out_mode
    if out_mode == OUT_BF16:
        _need(out, BF16, "gemm out")
    else:
        _need(out, F32, "gemm out")
    g.out_stride_m = out_strides[0] if out_strides[0] is not None else N
    g.out_stride_z2, g.out_stride_z3 = out_strides[1], out_strides[2]
    g.alpha = alpha
    if bias is not None:
        _need(bias, F32, "gemm bias")
        g.bias = bias.data_ptr()
    if bias_z2 is not None:
        _need(bias_z2, F32, "gemm bias_z2")
        g.bias_z2 = bias_z2.data_ptr()
    if residual is not None:
        _need(residual, BF16, "gemm residual")
        g.residual = residual.data_ptr()
        g.res_stride_m = res_strides[0] if res_strides[0] is not None else N
        g.res_stride_z2, g.res_stride_z3 = res_strides[1], res_strides[2]
    call("gemm", C.byref(g), _stream())
