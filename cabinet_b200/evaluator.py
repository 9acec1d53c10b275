"""Device-resident mirror of the reference evaluator ``MscEvalV0`` (``src/scripts/evaluate.py:32-256``).

Same constructor and result dict.  Differences, all on the hot-path side of the boundary:
  * the confusion matrix is an int64 (C, C) tensor that lives on the device (the reference keeps float64 on the
    CPU and ships an int64 mask over PCIe per image); counts are integers < 2^53 so the values are identical;
  * the fast mode (``scales=(1.0,)``, ``flip=False``, image not larger than the crop) is ONE fused call per batch:
    forward -> bilinear x8 -> argmax -> ``hist[pred, label]`` (``CABiNet.accumulate_hist``), logits never reach HBM;
  * the general mode (multi-scale / flip / sliding window) keeps the reference's control flow (pad, window grid,
    scales) on the host and runs its arithmetic as three kernels on fp32 NCHW probability accumulators: per chip
    ``cabinet_upsample_softmax_accum`` (x8 upsample of the 1/8 class map + softmax + flip average + overlap-count
    weight + window add; full-resolution logits never exist), per scale ``cabinet_prob_resize_accum`` (un-pad +
    resize back + sum over scales; also used to rescale the input), per batch ``cabinet_argmax_hist_nchw`` (argmax +
    confusion matrix).  A model without ``class_map8`` (any ``nn.Module`` returning logits) takes the reference's
    steps as torch ops on the device and shares the last kernel;
  * under ``torch.distributed`` every rank evaluates its own shard of the loader and the histograms are
    all-reduced once (the reference: ``dist.reduce(dst=0)``, ``evaluate.py:230-235``); every rank gets the result.
"""

from __future__ import annotations

import contextlib
import math
from typing import Any, Dict, Sequence

import numpy as np
import torch
import torch.nn.functional as F

from .constants import EVAL_STRIDE_RATE

try:
    import torch.distributed as dist
except ImportError:  # pragma: no cover
    dist = None


def _device_guard(dev):
    """Make ``dev`` the current CUDA device while C-ABI calls are enqueued (streams are per device); no-op on CPU."""
    dev = torch.device(dev)
    return torch.cuda.device(dev) if dev.type == "cuda" else contextlib.nullcontext()


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous, balanced split of ``n_items`` units over ``world`` ranks (first ranks take the remainder)."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def reduce_hist(hist: torch.Tensor) -> torch.Tensor:
    """Sum the per-rank int64 confusion matrices (no-op without an initialised process group)."""
    if dist is not None and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM)
    return hist


def metrics_from_hist(hist) -> Dict[str, Any]:
    """IoU / accuracy tail (reference: evaluate.py:239-251), float64 on the host."""
    h = hist.detach().cpu().numpy().astype(np.float64) if isinstance(hist, torch.Tensor) else np.asarray(hist, np.float64)
    ious = np.diag(h) / (h.sum(axis=0) + h.sum(axis=1) - np.diag(h) + 1e-8)
    return {"mIoU": np.nanmean(ious), "accuracy": np.diag(h).sum() / h.sum(),
            "iou_per_class": {f"class_{i}": ious[i] for i in range(len(ious))}, "confusion_matrix": h}


UAVID_MEAN_STD = ((0.480, 0.499, 0.457), (0.225, 0.208, 0.228))  # reference: src/datasets/uavid.py:179-180


def normalize_u8(images_u8: torch.Tensor, out: torch.Tensor, mean, std) -> torch.Tensor:
    """uint8 (N,H,W,3) device tensor -> normalised fp32 (N,3,H,W): torchvision ToTensor + Normalize on the device."""
    from . import _lib

    N, H, W, C = images_u8.shape
    if C != 3 or images_u8.dtype != torch.uint8 or not images_u8.is_contiguous():
        raise ValueError("normalize_u8 expects a contiguous uint8 (N, H, W, 3) tensor")
    if (tuple(out.shape) != (N, 3, H, W) or out.dtype != torch.float32 or not out.is_contiguous()
            or out.device != images_u8.device):
        raise ValueError(f"normalize_u8: out must be a contiguous fp32 {(N, 3, H, W)} tensor on {images_u8.device}, got "
                         f"{tuple(out.shape)} {out.dtype} on {out.device}")
    with _device_guard(out.device):
        rc = _lib.load().cabinet_normalize_u8(images_u8.data_ptr(), out.data_ptr(), N, H, W, *mean, *std,
                                              torch.cuda.current_stream(out.device).cuda_stream)
    _lib.check(rc, "normalize_u8")
    return out


class MscEvalV0:
    def __init__(self, model, dataloader, n_classes: int, ignore_label: int = 255, scales: Sequence[float] = (1.0,),
                 flip: bool = False, cropsize: int = 1024, device: torch.device = None, u8_mean_std=UAVID_MEAN_STD):
        self.model, self.dl, self.n_classes, self.ignore_label = model, dataloader, n_classes, ignore_label
        self.scales, self.flip, self.cropsize = tuple(scales), flip, cropsize
        self.device = device or next(model.parameters()).device
        # loaders may yield raw uint8 (N,H,W,3) images instead of normalised fp32 (N,3,H,W): they are uploaded as
        # bytes (4x less PCIe traffic) and normalised on the device with these per-channel statistics
        self.u8_mean_std = u8_mean_std

    def _crop_hw(self):
        """``cropsize`` is an int (the reference: square chips) or an (h, w) pair: whole rectangular images as ONE chip
        in the fast mode (BASELINE config 3 / 4 forward shapes).  The sliding-window general mode needs the int form."""
        cs = self.cropsize
        return (int(cs[0]), int(cs[1])) if isinstance(cs, (tuple, list)) else (int(cs), int(cs))

    # ---- general mode: the reference algorithm, tensors stay on the device
    def eval_chip(self, crop):
        prob = F.softmax(self.model(crop)[0].float(), dim=1)
        if self.flip:
            fl = self.model(torch.flip(crop, dims=(3,)).contiguous())[0].float()
            prob = (prob + F.softmax(torch.flip(fl, dims=(3,)), dim=1)) * 0.5
        return prob

    def crop_eval(self, image):
        cs = self.cropsize
        N, _, H, W = image.shape
        indices = None
        if H < cs or W < cs:  # centre zero-pad (reference: evaluate.py:60-72,102-111)
            tgt = (cs, cs) if max(H, W) < cs else (cs if H < W else H, cs if W < H else W)
            ph, pw = max(tgt[0] - H, 0), max(tgt[1] - W, 0)
            padded = torch.zeros(N, 3, tgt[0], tgt[1], device=image.device)
            padded[:, :, ph // 2: ph // 2 + H, pw // 2: pw // 2 + W] = image
            indices, image = (ph // 2, ph // 2 + H, pw // 2, pw // 2 + W), padded
        fh, fw = image.shape[2:]
        prob = torch.zeros((N, self.n_classes, fh, fw), device=image.device)
        count = torch.zeros((1, 1, fh, fw), device=image.device)
        if fh < cs or fw < cs:
            prob += self.eval_chip(image)
            count += 1
        else:
            stride = int(cs * EVAL_STRIDE_RATE)
            for iy in range(math.ceil((fh - cs) / stride) + 1):
                for ix in range(math.ceil((fw - cs) / stride) + 1):
                    y1, x1 = min(fh, stride * iy + cs), min(fw, stride * ix + cs)
                    prob[:, :, y1 - cs:y1, x1 - cs:x1] += self.eval_chip(image[:, :, y1 - cs:y1, x1 - cs:x1].contiguous())
                    count[:, :, y1 - cs:y1, x1 - cs:x1] += 1
        prob = prob / count.clamp(min=1)
        if indices is not None:
            prob = prob[:, :, indices[0]:indices[1], indices[2]:indices[3]]
        return prob

    def scale_crop_eval(self, image, scale):
        H, W = image.shape[2:]
        scaled = F.interpolate(image, [int(H * scale), int(W * scale)], mode="bilinear", align_corners=False)
        return F.interpolate(self.crop_eval(scaled), (H, W), mode="bilinear", align_corners=False)

    def _hist_from_probs(self, probs, labels, hist):
        """argmax over the class planes + ``hist[pred, label]`` in one kernel (reference: evaluate.py:222-228)."""
        from . import _lib

        probs, labels = probs.contiguous(), labels.contiguous()
        N, C, H, W = probs.shape
        _lib.check(_lib.load().cabinet_argmax_hist_nchw(probs.data_ptr(), N, C, H * W, None, labels.data_ptr(),
                                                         0 if labels.dtype == torch.int64 else 1, self.ignore_label,
                                                         hist.data_ptr(), torch.cuda.current_stream(hist.device).cuda_stream),
                   "argmax_hist_nchw")

    # ---- general mode on the fused kernels (models exposing the 1/8-resolution class map)
    @staticmethod
    def window_grid(full: int, cs: int):
        """Window starts along one axis and the per-pixel 1/overlap-count (reference: evaluate.py:127-146,149)."""
        stride = int(cs * EVAL_STRIDE_RATE)
        starts = [min(full, stride * i + cs) - cs for i in range(math.ceil((full - cs) / stride) + 1)]
        count = np.zeros(full, dtype=np.float32)
        for s0 in starts:
            count[s0:s0 + cs] += 1
        return starts, (1.0 / np.maximum(count, 1)).astype(np.float32)

    def _resize_accum(self, src, crop, dst):
        """dst (N,C,H,W) += bilinear(src[:, :, y0:y0+h, x0:x0+w] -> (H, W)), align_corners=False."""
        from . import _lib

        N, C, SH, SW = src.shape
        _lib.check(_lib.load().cabinet_prob_resize_accum(src.data_ptr(), N, C, SH, SW, crop[0], crop[1], crop[2], crop[3],
                                                          dst.data_ptr(), dst.shape[2], dst.shape[3],
                                                          torch.cuda.current_stream(dst.device).cuda_stream),
                   "prob_resize_accum")

    def _crop_eval_into(self, image, dst):
        """crop_eval (reference: evaluate.py:89-159) accumulating the normalised, un-padded probability map of
        ``image`` (N,3,H,W) into ``dst`` (N,C,H,W) fp32: one class-map forward (two with flip) + one fused kernel per
        window."""
        from . import _lib

        if isinstance(self.cropsize, (tuple, list)):
            raise ValueError("the sliding-window mode needs an int cropsize (square chips, reference evaluate.py:97)")
        lib, cs = _lib.load(), self.cropsize
        N, _, H, W = image.shape
        hst = wst = 0
        if H < cs or W < cs:  # centre zero-pad (reference: evaluate.py:60-72,102-111)
            tgt = (cs, cs) if max(H, W) < cs else (cs if H < W else H, cs if W < H else W)
            ph, pw = max(tgt[0] - H, 0), max(tgt[1] - W, 0)
            hst, wst = ph // 2, pw // 2
            padded = torch.zeros(N, 3, tgt[0], tgt[1], device=image.device)
            padded[:, :, hst:hst + H, wst:wst + W] = image
            image = padded
        fh, fw = image.shape[2:]
        if fh < cs or fw < cs:  # unreachable after padding, kept for parity with evaluate.py:121-124
            ys, xs, ch, cw = [0], [0], fh, fw
            inv_y = inv_x = None
        else:
            key = (fh, fw, cs, image.device)
            cache = self.__dict__.setdefault("_grids", {})
            if key not in cache:
                (ys, iy), (xs, ix) = self.window_grid(fh, cs), self.window_grid(fw, cs)
                cache[key] = (ys, xs, torch.from_numpy(iy).to(image.device), torch.from_numpy(ix).to(image.device))
            ys, xs, inv_y, inv_x = cache[key]
            ch = cw = cs
        # chips of one image batch are independent forwards: run several windows (and their flipped copies) as ONE
        # class-map forward of up to `max_chip_batch` images -- the low-resolution layers of the network are latency
        # bound at small batches.  Static chip buffers per group size: a repeated buffer replays a captured CUDA graph.
        wins = [(y0, x0) for y0 in ys for x0 in xs]
        per_win = N * (2 if self.flip else 1)
        group = max(1, getattr(self, "max_chip_batch", 32) // per_win)
        slots = self.__dict__.setdefault("_chips", {})
        stream = torch.cuda.current_stream(image.device).cuda_stream
        C = self.n_classes
        for g0 in range(0, len(wins), group):
            gw = wins[g0:g0 + group]
            skey = (len(gw) * per_win, ch, cw, image.device)
            if skey not in slots:
                slots[skey] = torch.empty((len(gw) * per_win, 3, ch, cw), device=image.device)
            chips = slots[skey]
            for i, (y0, x0) in enumerate(gw):
                view = image[:, :, y0:y0 + ch, x0:x0 + cw]
                chips[i * per_win:i * per_win + N].copy_(view)
                if self.flip:
                    chips[i * per_win + N:(i + 1) * per_win].copy_(torch.flip(view, dims=(3,)))
            maps = self.model.class_map8(chips)  # (len(gw) * per_win, h8, w8, C) fp32
            for i, (y0, x0) in enumerate(gw):
                m = maps[i * per_win:i * per_win + N]
                mf = maps[i * per_win + N:(i + 1) * per_win] if self.flip else None
                _lib.check(lib.cabinet_upsample_softmax_accum(
                    m.data_ptr(), mf.data_ptr() if mf is not None else None, N, m.shape[1], m.shape[2], C, ch, cw,
                    dst.data_ptr(), dst.stride(0), dst.stride(1), dst.stride(2), y0 - hst, x0 - wst, H, W,
                    inv_y[y0:].data_ptr() if inv_y is not None else None,
                    inv_x[x0:].data_ptr() if inv_x is not None else None, 1.0, stream), "upsample_softmax_accum")

    def _probs_fused(self, images):
        """sum over scales of scale_crop_eval (reference: evaluate.py:149-159,216-220) -> (N,C,H,W) fp32."""
        N, _, H, W = images.shape
        probs = torch.zeros((N, self.n_classes, H, W), device=images.device)
        for s in self.scales:
            hs, ws = int(H * s), int(W * s)
            if (hs, ws) == (H, W):  # F.interpolate to the same size is the identity
                self._crop_eval_into(images, probs)
                continue
            scaled = torch.zeros((N, 3, hs, ws), device=images.device)
            self._resize_accum(images, (0, 0, H, W), scaled)
            ps = torch.zeros((N, self.n_classes, hs, ws), device=images.device)
            self._crop_eval_into(scaled, ps)
            self._resize_accum(ps, (0, 0, hs, ws), probs)
        return probs

    def _fast_pipelined(self, dev, hist, masks_out=None, hist_trace=None):
        """Fast mode over the whole loader with host->device copies ahead of the fused forward.

        A ring of device buffer sets; a copy stream uploads work item i+1.. while the compute stream runs item i
        (forward + x8 upsample + argmax + confusion matrix in one fused tail).  fp32 host batches are PCIe-bound
        (12.6 MB per 1024^2 image against ~0.17 ms of compute), so they are cut into chunks of ``self.chunk`` images
        (default 8): the forward of a chunk starts as soon as ITS images have landed, and only the last chunk's
        forward is not hidden behind an upload.  ``masks_out``: optional list that receives a pinned uint8 host tensor
        per batch (asynchronous D2H, valid after the final synchronise).  ``hist_trace``: optional pinned int64
        (n_batches, C, C) host tensor that receives the running confusion matrix after every batch (asynchronous D2H
        of the step's result on the read-back stream; 8 C^2 bytes per batch)."""
        cur = torch.cuda.current_stream(dev)
        st = self.__dict__.setdefault("_pipe", {"stream": torch.cuda.Stream(dev), "d2h": torch.cuda.Stream(dev),
                                                "bufs": {}})
        copy_stream, ring = st["stream"], st["bufs"]  # device buffers persist across evaluate() calls
        d2h_stream = st["d2h"]  # mask read-back beside the next forward (on the compute stream it would serialise)
        # the fused forward + confusion-matrix call is replayed as a CUDA graph keyed by its buffer addresses: accumulate
        # into a persistent matrix (the caller's `hist` is a fresh allocation per evaluate(), which would force a
        # re-capture whenever the allocator hands out a different block -- it does under NCCL) and add it at the end
        if st.get("hist") is None or st["hist"].shape != hist.shape:
            st["hist"] = torch.zeros_like(hist)
        user_hist, hist = hist, st["hist"]
        hist.zero_()
        copy_stream.wait_stream(cur)
        n_batches, item = 0, 0
        for i, (images, labels) in enumerate(self.dl):
            if labels.dim() == 4:
                labels = labels.squeeze(1)
            if labels.dtype not in (torch.int64, torch.uint8):
                labels = labels.long()
            u8 = images.dtype == torch.uint8
            H, W = images.shape[1:3] if u8 else images.shape[2:]
            N = images.shape[0]
            if (H, W) != self._crop_hw():
                # not one chip == one image: this batch takes the general path (pad / sliding windows), in stream order
                if u8:
                    x32 = torch.empty((N, 3, H, W), dtype=torch.float32, device=dev)
                    images = normalize_u8(images.to(dev, non_blocking=True).contiguous(), x32, *self.u8_mean_std)
                self._general_batch(images.to(dev, non_blocking=True), labels.to(dev, non_blocking=True), hist, dev)
                n_batches += 1
                continue
            chunk = getattr(self, "chunk", 8)
            if u8 or chunk <= 0 or N <= chunk:  # uint8 uploads are compute-bound: one forward per batch
                chunk = max(N, 1)
            host = None
            if masks_out is not None:
                host = torch.empty((N, H, W), dtype=torch.uint8, pin_memory=True) if len(masks_out) <= i else masks_out[i]
                if len(masks_out) <= i:
                    masks_out.append(host)
            for c0 in range(0, N, chunk):
                im, lb = images[c0:c0 + chunk], labels[c0:c0 + chunk]
                n = im.shape[0]
                # ring of 2 (whole batches) or 4 (chunks) slots per work-item shape
                rkey = (tuple(im.shape), im.dtype, lb.dtype)
                slots = ring.get(rkey)
                if slots is None:
                    nslot = 2 if chunk >= N else 4
                    slots = ring[rkey] = {"next": 0, "sets": [], "n": nslot}
                    # the caching allocator may hand out memory that kernels already queued on the compute stream
                    # still touch: order the copy stream after them, and tell the allocator about the second stream
                    copy_stream.wait_stream(cur)
                    for _ in range(nslot):
                        bi = torch.empty(im.shape, dtype=im.dtype if u8 else torch.float32, device=dev)
                        bl = torch.empty(lb.shape, dtype=lb.dtype, device=dev)
                        bi.record_stream(copy_stream)
                        bl.record_stream(copy_stream)
                        slots["sets"].append({"img": bi, "lab": bl, "ready": torch.cuda.Event(), "consumed": None})
                    # one fp32 staging buffer PER work-item shape (a short last batch must not shrink the buffer the
                    # full-size batches of the next evaluate() normalise into)
                    slots["x32"] = torch.empty((n, 3, H, W), dtype=torch.float32, device=dev) if u8 else None
                slot = slots["sets"][slots["next"]]
                slots["next"] = (slots["next"] + 1) % slots["n"]
                with torch.cuda.stream(copy_stream):
                    if slot["consumed"] is not None:
                        copy_stream.wait_event(slot["consumed"])  # the forward that last read this buffer set is done
                    slot["img"].copy_(im, non_blocking=True)
                    slot["lab"].copy_(lb, non_blocking=True)
                    slot["ready"].record(copy_stream)
                cur.wait_event(slot["ready"])
                if slot.get("d2h_done") is not None:
                    cur.wait_event(slot["d2h_done"])  # this slot's mask buffer (a graph output) has been read back
                xin = normalize_u8(slot["img"], slots["x32"], *self.u8_mean_std) if u8 else slot["img"]
                mask = self.model.accumulate_hist(xin, slot["lab"], hist, self.ignore_label)
                if host is not None:
                    done = torch.cuda.Event()
                    done.record(cur)
                    mask.record_stream(d2h_stream)
                    with torch.cuda.stream(d2h_stream):
                        d2h_stream.wait_event(done)
                        host[c0:c0 + n].copy_(mask, non_blocking=True)
                        slot["d2h_done"] = torch.cuda.Event()
                        slot["d2h_done"].record(d2h_stream)
                slot["consumed"] = torch.cuda.Event()
                slot["consumed"].record(cur)
                item += 1
            if hist_trace is not None and i < hist_trace.shape[0]:
                snap = st.setdefault("snap", {}).get(i % 4)
                if snap is None:  # device-side snapshots (the graph keeps adding to `hist` while the copy is in flight)
                    snap = st["snap"][i % 4] = (torch.empty_like(hist), torch.cuda.Event())
                    snap[0].record_stream(d2h_stream)
                else:
                    cur.wait_event(snap[1])  # its previous read-back has finished
                snap[0].copy_(hist)
                done = torch.cuda.Event()
                done.record(cur)
                with torch.cuda.stream(d2h_stream):
                    d2h_stream.wait_event(done)
                    hist_trace[i].copy_(snap[0], non_blocking=True)
                    snap[1].record(d2h_stream)
            n_batches += 1
        cur.wait_stream(d2h_stream)  # the caller's stream order (and any event it records next) covers the read-backs
        user_hist.add_(hist)
        return n_batches

    @torch.no_grad()
    def evaluate(self, masks_out=None, hist_trace=None) -> Dict[str, Any]:
        self.model.eval()
        dev = next(self.model.parameters()).device
        hist = torch.zeros((self.n_classes, self.n_classes), dtype=torch.int64, device=dev)
        fast = self.scales == (1.0,) and not self.flip and hasattr(self.model, "accumulate_hist")
        with _device_guard(dev):
            if fast and getattr(self, "pipelined", True):
                # fast vs general is decided per batch inside the loop (no peeking at the loader: a one-shot iterable
                # would lose its first batch, a DataLoader would spin up its workers twice)
                self._fast_pipelined(dev, hist, masks_out, hist_trace)
            else:
                for images, labels in self.dl:
                    images = images.to(dev, non_blocking=True)
                    labels = labels.to(dev, non_blocking=True)
                    if images.dtype == torch.uint8:
                        x32 = torch.empty((images.shape[0], 3) + tuple(images.shape[1:3]), dtype=torch.float32, device=dev)
                        images = normalize_u8(images.contiguous(), x32, *self.u8_mean_std)
                    self._general_batch(images, labels, hist, dev, fast)
        reduce_hist(hist)
        return metrics_from_hist(hist)

    def _general_batch(self, images, labels, hist, dev, fast=False):
        """One device-resident batch through the reference's control flow (evaluate.py:204-228)."""
        if labels.dim() == 4:
            labels = labels.squeeze(1)
        if labels.dtype not in (torch.int64, torch.uint8):
            labels = labels.long()
        H, W = images.shape[2:]
        if fast and (H, W) == self._crop_hw():  # one chip == the image: argmax(softmax) == argmax
            self.model.accumulate_hist(images.float().contiguous(), labels.contiguous(), hist, self.ignore_label)
            return
        if images.size(0) == 0:
            return
        if hasattr(self.model, "class_map8") and getattr(self, "fused_general", True):
            probs = self._probs_fused(images.float().contiguous())
        else:  # any other nn.Module: the reference's steps as device tensor ops
            probs = torch.zeros((images.size(0), self.n_classes, H, W), device=dev)
            for s in self.scales:
                probs += self.scale_crop_eval(images.float(), s)
        self._hist_from_probs(probs, labels, hist)

    __call__ = evaluate
