"""Gradient synchronisation of the data-parallel training configuration (BASELINE config 5; the reference trains
single-process, SURVEY 8e: "replicas + gradient all-reduce (new)").

One process per GPU holds a full replica; after ``loss.backward()`` the gradients are averaged over the ranks.  They live
in a few flat fp32 buckets (``p.grad`` is a VIEW into its bucket, so there is no pack / unpack copy), filled in reverse
parameter order -- the order in which backward produces them -- and every bucket is one asynchronous ``all_reduce`` (NCCL
over NVLink / NVSwitch on the GPUs, gloo in the CPU tests), so the reduction of the late layers' gradients overlaps the
backward of the early ones when ``reduce_bucket`` is called from a hook, or runs as ``len(buckets)`` overlapping
collectives when called once after backward.  Parameters that never receive a gradient are left out: the backbone's
``mobile.classifier`` (1.24 M parameters, not on the forward path).
"""

from __future__ import annotations

from typing import Callable, Iterable, List, Tuple

import torch

try:
    import torch.distributed as dist
except ImportError:  # pragma: no cover
    dist = None


def _unused(name: str) -> bool:
    return name.startswith("mobile.classifier")


class GradBuckets:
    def __init__(self, named_params: Iterable[Tuple[str, torch.nn.Parameter]], bucket_bytes: int = 16 << 20,
                 skip: Callable[[str], bool] = _unused):
        named_params = list(named_params)
        items = [(n, p) for n, p in named_params if p.requires_grad and not skip(n)]
        self.skipped = [n for n, p in named_params if p.requires_grad and skip(n)]
        self.names: List[List[str]] = []
        self.buckets: List[torch.Tensor] = []
        self.params: List[List[torch.nn.Parameter]] = []
        self._checked = False
        cur, cur_names, cur_elems = [], [], 0
        limit = max(1, bucket_bytes // 4)
        for n, p in reversed(items):  # backward produces gradients back to front
            if cur and cur_elems + p.numel() > limit:
                self._close(cur, cur_names)
                cur, cur_names, cur_elems = [], [], 0
            cur.append(p)
            cur_names.append(n)
            cur_elems += p.numel()
        if cur:
            self._close(cur, cur_names)

    def _close(self, params, names):
        dev = params[0].device
        flat = torch.zeros(sum(p.numel() for p in params), dtype=torch.float32, device=dev)
        off = 0
        for p in params:
            if p.dtype != torch.float32:
                raise ValueError("GradBuckets expects fp32 master parameters")
            p.grad = flat[off:off + p.numel()].view_as(p)  # autograd accumulates into this view in place
            off += p.numel()
        self.buckets.append(flat)
        self.names.append(names)
        self.params.append(list(params))

    def zero_(self):
        """Zero the gradients IN PLACE.  Use this (or ``optimizer.zero_grad(set_to_none=False)``) instead of the default
        ``optimizer.zero_grad()``: ``set_to_none=True`` detaches every ``p.grad`` from its bucket."""
        self.attach()
        for b in self.buckets:
            b.zero_()

    def attach(self, strict: bool = False) -> int:
        """Make every ``p.grad`` a view into its bucket again.  A gradient that was replaced (``zero_grad(set_to_none=
        True)`` followed by a backward that allocated a fresh tensor) is copied into the bucket first, so nothing is
        lost; with ``strict`` a detached gradient raises instead.  -> number of gradients re-attached."""
        fixed = 0
        for flat, params in zip(self.buckets, self.params):
            off = 0
            for p in params:
                view = flat[off:off + p.numel()].view_as(p)
                off += p.numel()
                g = p.grad
                if g is not None and g.data_ptr() == view.data_ptr() and g.shape == view.shape:
                    continue
                if strict:
                    raise RuntimeError("GradBuckets: a .grad no longer aliases its bucket (optimizer.zero_grad() defaults to "
                                       "set_to_none=True; use GradBuckets.zero_() or zero_grad(set_to_none=False))")
                if g is None:
                    view.zero_()
                else:
                    view.copy_(g)
                p.grad = view
                fixed += 1
        return fixed

    @staticmethod
    def _world() -> int:
        return dist.get_world_size() if dist is not None and dist.is_available() and dist.is_initialized() else 1

    def reduce_bucket(self, i: int):
        """Start the all-reduce of bucket ``i``; -> a handle for ``finish`` (None in a single process)."""
        if i == 0 or not self._checked:
            # gradients produced after optimizer.zero_grad(set_to_none=True) live outside the buckets: fold them back in
            # before anything is reduced (otherwise stale buckets would be averaged and the replicas diverge silently)
            self.attach()
            self._checked = True
        if self._world() == 1:
            return None
        return dist.all_reduce(self.buckets[i], op=dist.ReduceOp.SUM, async_op=True)

    def finish(self, handles) -> None:
        """Wait for the collectives and turn the sums into means."""
        world = self._world()
        self._checked = False
        for h in handles:
            if h is not None:
                h.wait()
        if world > 1:
            for b in self.buckets:
                b.mul_(1.0 / world)

    def all_reduce(self) -> None:
        """All buckets, overlapping; blocks until the averaged gradients are in place."""
        self.finish([self.reduce_bucket(i) for i in range(len(self.buckets))])
