"""Data-parallel gradient exchange for the transformer trainers: ONE flat fp32 bucket, ONE all-reduce.

The reference reaches the same collective through accelerate -> torch DDP (trainer.py:75, 396, 834, 1136,
1432: `accelerator.prepare`, and the reducer hooks fired by `accelerator.backward`, :580, 949, 1247, 1548)
with 25 MB buckets and `find_unused_parameters=True`.  Here every parameter's `.grad` is a view into one
contiguous buffer, so the whole exchange is a single NCCL all-reduce over NVLink / NVSwitch followed by
a scale by 1/world (mean, as DDP does).  Works on CPU with gloo for the world_size-2 tests.

Overlap with the backward (what DDP's bucket hooks do): `Transformer.grad_ready_hook(i)` fires when layer i's
gradients are complete; `reduce_range_async` then starts the all-reduce of that layer's slice of the flat buffer on
the communication stream while the earlier layers are still being differentiated, and `finish()` reduces whatever
is left (embeddings, heads), waits and applies the 1/world scale once.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class FlatGradBucket:
    def __init__(self, params, dtype=torch.float32):
        self.params = [p for p in params if p.requires_grad]
        assert self.params, "no trainable parameters"
        dev = self.params[0].device
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel, device=dev, dtype=dtype)
        off = 0
        self._span = {}
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            self._span[id(p)] = (off, off + n)
            off += n
        self._pending, self._done = [], []  # async work handles / [lo, hi) intervals already handed to the collective

    def attach(self, module):
        """let every `Transformer` stack inside `module` accumulate its parameter gradients straight into this
        bucket (no per-parameter temporaries; parameter hooks do not fire for those parameters)."""
        from .transformer import Transformer

        for m in module.modules():
            if isinstance(m, Transformer):
                m.accumulate_into_grad = True
        return self

    def zero_(self):
        """use this instead of `optimizer.zero_grad()` (whose default set_to_none=True detaches the views)"""
        self.sync_views()
        self.flat.zero_()

    def sync_views(self):
        """re-establish `p.grad is a view of self.flat`.  `optimizer.zero_grad(set_to_none=True)` (the torch default)
        or an assignment to `p.grad` detaches a parameter from the bucket; its gradient is then copied in and the
        view restored, so the collective never reduces a stale buffer."""
        for p in self.params:
            lo, hi = self._span[id(p)]
            view = self.flat[lo:hi].view_as(p)
            g = p.grad
            if g is None:
                view.zero_()
                p.grad = view
            elif g.data_ptr() != view.data_ptr() or g.dtype != self.flat.dtype or not g.is_contiguous():
                view.copy_(g)
                p.grad = view

    def all_reduce_mean(self, group=None, async_op=False):
        """sum over ranks then divide by world size (DDP semantics).  Returns the work handle if async."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return None
        world = dist.get_world_size(group)
        self.sync_views()
        if self._has_avg(group) and not async_op:
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=group)  # ncclAvg: no separate 1/world pass
            return None
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        if async_op:
            return work
        self.flat.div_(world)
        return None

    @staticmethod
    def _has_avg(group=None):
        return dist.get_backend(group) == "nccl"

    # ---- overlapped exchange ----------------------------------------------------------------------
    def range_of(self, params):
        """[lo, hi) of the flat buffer covered by `params` (they must be adjacent in the bucket)."""
        spans = sorted(self._span[id(p)] for p in params if id(p) in self._span)
        assert spans, "parameters are not part of this bucket"
        for (_, e), (s2, _) in zip(spans, spans[1:]):
            assert e == s2, "parameters are not contiguous in the bucket"
        return spans[0][0], spans[-1][1]

    def reduce_range_async(self, lo, hi, group=None):
        """start the SUM all-reduce of flat[lo:hi]; the 1/world scale is applied by finish()."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1 or hi <= lo:
            return
        op = dist.ReduceOp.AVG if self._has_avg(group) else dist.ReduceOp.SUM
        self._pending.append(dist.all_reduce(self.flat[lo:hi], op=op, group=group, async_op=True))
        self._done.append((lo, hi))

    def finish(self, group=None):
        """all-reduce every slice not yet handed over, wait for all of them, scale by 1/world (mean)."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            self._pending, self._done = [], []
            return
        if not self._done:
            self.sync_views()
        avg = self._has_avg(group)
        op = dist.ReduceOp.AVG if avg else dist.ReduceOp.SUM
        pos = 0
        for lo, hi in sorted(self._done) + [(self.numel, self.numel)]:
            if lo > pos:
                self._pending.append(dist.all_reduce(self.flat[pos:lo], op=op, group=group, async_op=True))
            pos = max(pos, hi)
        for w in self._pending:
            w.wait()
        self._pending, self._done = [], []
        if not avg:
            self.flat.div_(dist.get_world_size(group))

    def grad_norm(self):
        self.sync_views()
        return self.flat.norm(2)

    def clip_grad_norm_(self, max_norm):
        """global L2 clip over the flat bucket (accelerator.clip_grad_norm_, trainer.py:596, 954, 1252, 1553)."""
        norm = self.grad_norm()
        scale = (max_norm / (norm + 1e-6)).clamp(max=1.0)
        self.flat.mul_(scale)
        return norm
