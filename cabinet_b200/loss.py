"""Drop-in for the reference training loss ``OhemCELoss`` (``src/utils/loss.py:11-83``) on the device.

Same constructor, same ``forward(logits, labels) -> scalar``, same edge cases (no valid pixel -> 0 that still
requires grad; ``n_min`` clamped to the number of valid pixels; class weights as a buffer that follows ``.to()``), but the
per-pixel cross-entropy, the selection (threshold set or k largest -- an exact radix select instead of a sort of every
valid loss) and the gradient are the C-ABI kernels ``cabinet_ohem_ce_forward`` / ``_backward``; nothing is decided on
the host, so the loss does not synchronise the training step.  CUDA tensors only (no CPU path).
"""

from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib
from ._lib import BF16, F32


class _OhemCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, weight, thresh, n_min, ignore_lb):
        if not logits.is_cuda:
            raise RuntimeError("cabinet_b200.loss.OhemCELoss has no CPU path: move logits and labels to a CUDA device")
        if logits.dim() != 4 or labels.shape != (logits.shape[0],) + tuple(logits.shape[2:]):
            raise ValueError(f"expected logits (N,C,H,W) and labels (N,H,W), got {tuple(logits.shape)} / {tuple(labels.shape)}")
        if logits.dtype not in (torch.float32, torch.bfloat16):
            logits = logits.float()  # fp16 autocast outputs: cross_entropy runs in fp32 under autocast anyway
        x = logits.contiguous()
        if labels.dtype not in (torch.int64, torch.uint8):
            labels = labels.long()
        lb = labels.contiguous()
        if lb.device != x.device:
            raise RuntimeError(f"labels on {lb.device}, logits on {x.device}")
        N, C, H, W = x.shape
        lib = _lib.load()
        dev = x.device
        loss_px = torch.empty(N * H * W, dtype=torch.float32, device=dev)
        ws = torch.empty((int(lib.cabinet_ohem_workspace_bytes()) + 7) // 8, dtype=torch.int64, device=dev)
        out = torch.empty((), dtype=torch.float32, device=dev)
        w = None
        if weight is not None:
            w = weight.to(device=dev, dtype=torch.float32).contiguous()
            if w.numel() != C:
                raise ValueError(f"class weight has {w.numel()} entries for {C} classes")
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            rc = (lib.cabinet_ohem_ce_forward(x.data_ptr(), BF16 if x.dtype == torch.bfloat16 else F32, lb.data_ptr(),
                                               0 if lb.dtype == torch.int64 else 1, N, C, H * W,
                                               w.data_ptr() if w is not None else None, int(ignore_lb), float(thresh),
                                               max(1, int(n_min)), loss_px.data_ptr(), ws.data_ptr(), out.data_ptr(),
                                               stream))
        _lib.check(rc, "ohem_ce_forward")
        ctx.save_for_backward(x, lb, loss_px, ws, w if w is not None else torch.empty(0, device=dev))
        ctx.has_weight = w is not None
        ctx.in_dtype = logits.dtype
        return out

    @staticmethod
    def backward(ctx, grad_out):
        x, lb, loss_px, ws, w = ctx.saved_tensors
        N, C, H, W = x.shape
        lib = _lib.load()
        g = grad_out.to(device=x.device, dtype=torch.float32).contiguous()
        grad = torch.empty_like(x)
        with torch.cuda.device(x.device):
            stream = torch.cuda.current_stream(x.device).cuda_stream
            rc = (lib.cabinet_ohem_ce_backward(x.data_ptr(), BF16 if x.dtype == torch.bfloat16 else F32, lb.data_ptr(),
                                                0 if lb.dtype == torch.int64 else 1, N, C, H * W,
                                                w.data_ptr() if ctx.has_weight else None, loss_px.data_ptr(),
                                                ws.data_ptr(), g.data_ptr(), grad.data_ptr(), stream))
        _lib.check(rc, "ohem_ce_backward")
        return grad, None, None, None, None, None


class OhemCELoss(nn.Module):
    """Online-hard-example-mining cross-entropy (reference: src/utils/loss.py:11-83)."""

    def __init__(self, thresh, n_min, ignore_lb=255, weight=None):
        super().__init__()
        self.thresh = float(thresh)
        self.n_min = int(n_min)
        self.ignore_lb = ignore_lb
        if weight is not None and not isinstance(weight, torch.Tensor):
            weight = torch.tensor(weight, dtype=torch.float32)
        self.register_buffer("weight", weight)

    def forward(self, logits, labels):
        w = self.weight if isinstance(self.weight, torch.Tensor) else None
        return _OhemCE.apply(logits, labels, w, self.thresh, self.n_min, self.ignore_lb)

    def extra_repr(self):
        return f"thresh={self.thresh}, n_min={self.n_min}, ignore_lb={self.ignore_lb}"
