"""Host -> device prefetch for the training loop (reference: src/scripts/train.py:420-428, ``im = im.cuda(); lb =
lb.cuda()`` at the top of every iteration, i.e. the copy of batch i sits between step i - 1 and step i on one stream).

``DevicePrefetcher`` wraps any iterable of tensor tuples (a ``DataLoader`` with ``pin_memory=True``) and issues the
copies of batch i + 1 on a side stream before it hands out batch i, so they run under the kernels of step i.
"""

from __future__ import annotations

from typing import Iterable, Iterator, Optional, Tuple

import torch


class DevicePrefetcher:
    def __init__(self, loader: Iterable, device):
        self.loader, self.device = loader, torch.device(device)
        # one side stream for the life of the object: the caching allocator keeps a pool per stream, so a fresh stream per
        # epoch would start every epoch with cudaMallocs
        self._side = torch.cuda.Stream(self.device) if self.device.type == "cuda" else None

    def _load(self, it, stream) -> Optional[Tuple[torch.Tensor, ...]]:
        try:
            batch = next(it)
        except StopIteration:
            return None
        if isinstance(batch, torch.Tensor):
            batch = (batch,)
        if stream is None:
            return tuple(t.to(self.device) for t in batch)
        with torch.cuda.stream(stream):
            return tuple(t.to(self.device, non_blocking=True) for t in batch)

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, ...]]:
        it = iter(self.loader)
        if self.device.type != "cuda":
            while (b := self._load(it, None)) is not None:
                yield b
            return
        side = self._side
        nxt = self._load(it, side)
        while nxt is not None:
            main = torch.cuda.current_stream(self.device)
            main.wait_stream(side)
            cur = nxt
            for t in cur:
                t.record_stream(main)  # the side stream's allocator must not reuse it while step i reads it
            nxt = self._load(it, side)
            yield cur

    def __len__(self):
        return len(self.loader)
