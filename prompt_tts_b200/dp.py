"""Data-parallel gradient exchange for the denoiser (the reference's only parallelism: accelerate -> torch DDP,
train.py:25-29,67-69,80,115).  One process per GPU; the batch is sharded, weights are replicated.

`GradSync` replaces DDP's reducer for the tape engine:
  * all parameter gradients of a step live in ONE flat fp32 buffer laid out in the order in which the backward sweep
    completes them (learned on the first step; static afterwards -- every later step asserts it), so a bucket is a contiguous
    slice -- no packing copies.  `param.grad` is a view of that buffer from the first step on;
  * as soon as the sweep has completed a bucket's worth of gradients, `all_reduce(AVG)` is issued asynchronously on that
    slice (NCCL over NVLink/NVSwitch; the NCCL stream orders itself after the compute stream), overlapping the rest of the
    backward; `finish()` makes the compute stream wait for the outstanding buckets.  With `comm_dtype=torch.bfloat16` a bucket is
    cast to a bf16 staging slice first (half the NVLink bytes and half the time NCCL's CTAs compete with the GEMMs for SMs);
    the fp32 buffer stays the accumulator and gets the averaged values back in `finish()`;
  * gradient accumulation (`accelerator.accumulate` / DDP `no_sync`, train.py:27,80): micro-steps with `sync=False` neither zero
    the buffer nor communicate; the last micro-step of a window reduces the accumulated sum;
  * parameters and buffers are broadcast from rank 0 at construction (what DDP's constructor does);
  * the 32 dead `proj_out` tensors never get a gradient and are simply absent from the buffer -- no
    `find_unused_parameters` graph walk.
The path has one exchange step per training step and no other collective.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

ALIGN = 64      # every group starts on a 64-element boundary of the flat buffer (16-byte aligned bf16 / fp32 views, whole vectors)


class GradSync:
    def __init__(self, model, world_size: Optional[int] = None, bucket_mb: float = 128.0, group=None,
                 comm_dtype: torch.dtype = torch.float32, broadcast: bool = True):
        self.group = group
        self.world = world_size if world_size is not None else (dist.get_world_size(group) if dist.is_initialized() else 1)
        assert comm_dtype in (torch.float32, torch.bfloat16)
        self.comm_dtype = comm_dtype
        self.bucket_elems = int(bucket_mb * (1 << 20) // 4) // ALIGN * ALIGN
        self.layout: Optional[Dict[int, Tuple[int, int, str]]] = None      # id(first param of a group) -> (offset, numel, kind)
        self.groups: List[Tuple[List[torch.Tensor], int, int, str]] = []   # (params, offset, numel, kind) in buffer order
        self.flat: Optional[torch.Tensor] = None
        self.comm: Optional[torch.Tensor] = None        # bf16 staging buffer (comm_dtype == bf16)
        self.cast_back = True       # copy the averaged bf16 values back into the fp32 buffer (an optimiser that reads `comm` clears it)
        self.total = 0
        self._done = 0
        self._sent = 0
        self._works: List = []
        self._ranges: List[Tuple[int, int]] = []
        self._tape = None
        self._sync = True
        self.n_buckets_last = 0
        if self.world > 1 and broadcast and model is not None:
            with torch.no_grad():
                for t in list(model.parameters()) + list(model.buffers()):
                    dist.broadcast(t.data, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)

    # ---- tape hooks
    def attach(self, tape, zero: bool = True, sync: bool = True) -> None:
        """zero: start a new accumulation window (clear the flat buffer); sync: reduce across ranks during this backward."""
        self._tape = tape
        self._done = self._sent = 0
        self._works, self._ranges = [], []
        self._sync = sync
        tape.on_ready = self._ready
        if self.layout is not None:
            if zero:
                self.flat.zero_()
            tape.grad_alloc = self._alloc

    def _alloc(self, params: Sequence[torch.Tensor], kind: str) -> Optional[torch.Tensor]:
        ent = self.layout.get(id(params[0]))
        if ent is None:
            return None
        off, n, k = ent
        if k != kind:
            raise RuntimeError(f"GradSync: gradient layout kind changed for a parameter group ({k} -> {kind})")
        return self.flat[off:off + n]

    def _ready(self, new) -> None:
        if self.layout is None:
            return
        for params, buf, kind in new:
            ent = self.layout.get(id(params[0]))
            if ent is None or ent[0] != self._done:
                raise RuntimeError("GradSync: the backward sweep completed gradients in a different order than on the first step "
                                   "(the flat layout, and with it the bucket boundaries, assume a static model)")
            self._done = ent[0] + _pad(ent[1])
        if not self._sync:
            return
        while self._done - self._sent >= self.bucket_elems:
            self._launch(self._sent, self._sent + self.bucket_elems)

    def _launch(self, lo: int, hi: int) -> None:
        self._sent = hi
        if self.world == 1:
            return
        if self._tape is not None and hasattr(self._tape, "join_side"):
            self._tape.join_side()      # bias gradients computed on the side stream belong to this bucket too
        if self.comm_dtype == torch.bfloat16:
            from . import ops
            ops.cast_bf16(self.flat[lo:hi], self.comm[lo:hi])
            w = dist.all_reduce(self.comm[lo:hi], op=dist.ReduceOp.AVG, group=self.group, async_op=True)
        else:
            w = dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.AVG, group=self.group, async_op=True)
        self._works.append(w)
        self._ranges.append((lo, hi))

    def finish(self) -> None:
        tape = self._tape
        if self.layout is None:
            # first step: learn the completion order, build the flat buffer, move this step's gradients into it (param.grad will
            # be views of it from now on) and reduce them in one go
            order = tape.pgrad_order
            self.layout, off = {}, 0
            self.groups = []
            for params, buf, kind in order:
                self.layout[id(params[0])] = (off, buf.numel(), kind)
                self.groups.append((list(params), off, buf.numel(), kind))
                off += _pad(buf.numel())
            self.total = off
            dev = order[0][1].device
            self.flat = torch.zeros(off, dtype=torch.float32, device=dev)
            if self.comm_dtype == torch.bfloat16 and self.world > 1:
                self.comm = torch.zeros(off, dtype=torch.bfloat16, device=dev)
            for params, buf, kind in order:
                o, n, _ = self.layout[id(params[0])]
                self.flat[o:o + n].copy_(buf.reshape(-1))
                tape.rebind(params, kind, self.flat[o:o + n])
            if self.world > 1 and self._sync:
                self._launch(0, self.total)
                self._wait()
            self.n_buckets_last = 1
            return
        if not self._sync:
            self.n_buckets_last = 0
            return
        if self._done > self._sent:
            self._launch(self._sent, self._done)
        self.n_buckets_last = len(self._works)
        self._wait()

    def _wait(self) -> None:
        for w in self._works:
            w.wait()
        if self.comm_dtype == torch.bfloat16 and self.cast_back and self._ranges:
            from . import ops
            lo, hi = self._ranges[0][0], self._ranges[-1][1]
            ops.call("cast_bf16_to_f32", ops._p(self.comm[lo:hi]), ops._p(self.flat[lo:hi]), hi - lo, ops._stream())
        self._works, self._ranges = [], []


def _pad(n: int) -> int:
    return (n + ALIGN - 1) // ALIGN * ALIGN
