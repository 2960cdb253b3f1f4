"""Data-parallel gradient exchange for the denoiser (the reference's only parallelism: accelerate -> torch DDP,
train.py:25-29,67-69,115).  One process per GPU; the batch is sharded, weights are replicated.

`GradSync` replaces DDP's reducer for the tape engine:
  * all parameter gradients of a step live in ONE flat fp32 buffer laid out in the order in which the backward sweep
    completes them (learned on the first step; static afterwards), so a bucket is a contiguous slice -- no packing copies;
  * as soon as the sweep has completed a bucket's worth of gradients, `all_reduce(AVG)` is issued asynchronously on that
    slice (NCCL over NVLink/NVSwitch; the NCCL stream orders itself after the compute stream), overlapping the rest of the
    backward; `finish()` makes the compute stream wait for the outstanding buckets;
  * the 32 dead `proj_out` tensors never get a gradient and are simply absent from the buffer -- no
    `find_unused_parameters` graph walk.
The path has one exchange step per training step and no other collective.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


class GradSync:
    def __init__(self, model, world_size: Optional[int] = None, bucket_mb: float = 128.0, group=None):
        self.group = group
        self.world = world_size if world_size is not None else (dist.get_world_size(group) if dist.is_initialized() else 1)
        self.bucket_elems = int(bucket_mb * (1 << 20) // 4)
        self.layout: Optional[Dict[int, Tuple[int, int]]] = None      # id(first param of a group) -> (offset, numel)
        self.flat: Optional[torch.Tensor] = None
        self.total = 0
        self._done = 0
        self._sent = 0
        self._works: List = []
        self._tape = None
        self.n_buckets_last = 0

    # ---- tape hooks
    def attach(self, tape) -> None:
        self._tape = tape
        self._done = self._sent = 0
        self._works = []
        tape.on_ready = self._ready
        if self.layout is not None:
            self.flat.zero_()
            tape.grad_alloc = self._alloc

    def _alloc(self, params: Sequence[torch.Tensor]) -> Optional[torch.Tensor]:
        ent = self.layout.get(id(params[0]))
        if ent is None:
            return None
        off, n = ent
        return self.flat[off:off + n]

    def _ready(self, new) -> None:
        if self.layout is None:
            return
        for params, buf in new:
            self._done += buf.numel()
        while self._done - self._sent >= self.bucket_elems:
            self._launch(self._sent, self._sent + self.bucket_elems)

    def _launch(self, lo: int, hi: int) -> None:
        if self.world == 1:
            self._sent = hi
            return
        w = dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.AVG, group=self.group, async_op=True)
        self._works.append(w)
        self._sent = hi

    def finish(self) -> None:
        tape = self._tape
        if self.layout is None:
            # first step: learn the completion order, build the flat buffer, and reduce this step's gradients in one go
            order = tape.pgrad_order
            self.layout, off = {}, 0
            self.groups = []      # (parameters whose gradients share one contiguous slice, offset, numel) in buffer order
            for params, buf in order:
                self.layout[id(params[0])] = (off, buf.numel())
                self.groups.append((list(params), off, buf.numel()))
                off += buf.numel()
            self.total = off
            dev = order[0][1].device
            self.flat = torch.zeros(off, dtype=torch.float32, device=dev)
            for params, buf in order:
                o, n = self.layout[id(params[0])]
                self.flat[o:o + n].copy_(buf.reshape(-1))
            if self.world > 1:
                dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.group)
                for params, buf in order:
                    o, n = self.layout[id(params[0])]
                    buf.reshape(-1).copy_(self.flat[o:o + n])
            self.n_buckets_last = 1
            return
        if self._done > self._sent:
            self._launch(self._sent, self._done)
        self.n_buckets_last = len(self._works)
        for w in self._works:
            w.wait()
        self._works = []
