"""ORACLE (test infrastructure) — CPU restatement of the evaluation tail of the hot path.

Restates ``MscEvalV0`` of the reference (``src/scripts/evaluate.py:32-256``) around an arbitrary
``model(crop) -> (logits, aux)`` callable: zero-pad, sliding-window chips with overlap-count
normalisation, flip-TTA, multi-scale sum, argmax, ``hist[pred, label]`` confusion matrix and the
IoU / accuracy tail.  Pinned against the imported reference evaluator by
``oracle/make_golden.py`` -> ``tests/golden/evaluator_*.npz``.
"""

from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

EVAL_STRIDE_RATE = 5 / 6.0  # reference: src/models/constants.py:26


def compute_hist(pred: np.ndarray, label: np.ndarray, n_classes: int, ignore_label: int = 255) -> np.ndarray:
    """reference: evaluate.py:162-191 — valid = label != ignore; clip both; bincount(pred*C + label)."""
    pred = np.asarray(pred)
    label = np.asarray(label)
    valid = label != ignore_label
    p = np.clip(pred[valid].astype(np.int64), 0, n_classes - 1)
    l = np.clip(label[valid].astype(np.int64), 0, n_classes - 1)
    return np.bincount(p * n_classes + l, minlength=n_classes ** 2).reshape(n_classes, n_classes)


def metrics(hist: np.ndarray) -> dict:
    """reference: evaluate.py:239-251."""
    hist = hist.astype(np.float64)
    ious = np.diag(hist) / (hist.sum(axis=0) + hist.sum(axis=1) - np.diag(hist) + 1e-8)
    return {"mIoU": np.nanmean(ious), "accuracy": np.diag(hist).sum() / hist.sum(),
            "iou_per_class": {f"class_{i}": ious[i] for i in range(len(ious))}, "confusion_matrix": hist}


def pad_tensor(t: torch.Tensor, size):
    """reference: evaluate.py:60-72 — centre zero-pad to ``size``; returns (padded, [hst, hed, wst, wed])."""
    N, C, H, W = t.shape
    ph, pw = max(size[0] - H, 0), max(size[1] - W, 0)
    hst, wst = ph // 2, pw // 2
    out = torch.zeros(N, C, size[0], size[1], device=t.device)
    out[:, :, hst:hst + H, wst:wst + W] = t
    return out, [hst, hst + H, wst, wst + W]


def chip_windows(full_h: int, full_w: int, cropsize: int):
    """Sliding-window chip origins (reference: evaluate.py:124-135)."""
    stride = int(cropsize * EVAL_STRIDE_RATE)
    n_x = math.ceil((full_w - cropsize) / stride) + 1
    n_y = math.ceil((full_h - cropsize) / stride) + 1
    wins = []
    for iy in range(n_y):
        for ix in range(n_x):
            y_end = min(full_h, stride * iy + cropsize)
            x_end = min(full_w, stride * ix + cropsize)
            wins.append((y_end - cropsize, y_end, x_end - cropsize, x_end))
    return wins


@torch.no_grad()
def eval_chip(model, crop, flip):
    """reference: evaluate.py:74-87."""
    prob = F.softmax(model(crop)[0].float(), dim=1)
    if flip:
        fl = model(torch.flip(crop, dims=(3,)))[0].float()
        prob = (prob + F.softmax(torch.flip(fl, dims=(3,)), dim=1)) * 0.5
    return prob


@torch.no_grad()
def crop_eval(model, image, n_classes, cropsize, flip):
    """reference: evaluate.py:89-148."""
    N, C, H, W = image.shape
    indices = None
    if H < cropsize or W < cropsize:
        long_size = max(H, W)
        target = (cropsize, cropsize) if long_size < cropsize else (cropsize if H < W else H, cropsize if W < H else W)
        image, indices = pad_tensor(image, target)
    fh, fw = image.shape[2:]
    prob = torch.zeros((N, n_classes, fh, fw), device=image.device)
    count = torch.zeros((1, 1, fh, fw), device=image.device)
    if fh < cropsize or fw < cropsize:
        prob += eval_chip(model, image, flip)
        count += 1
    else:
        for y0, y1, x0, x1 in chip_windows(fh, fw, cropsize):
            prob[:, :, y0:y1, x0:x1] += eval_chip(model, image[:, :, y0:y1, x0:x1], flip)
            count[:, :, y0:y1, x0:x1] += 1
    prob = prob / count.clamp(min=1)
    if indices is not None:
        hst, hed, wst, wed = indices
        prob = prob[:, :, hst:hed, wst:wed]
    return prob


@torch.no_grad()
def scale_crop_eval(model, image, scale, n_classes, cropsize, flip):
    """reference: evaluate.py:150-159."""
    N, C, H, W = image.shape
    scaled = F.interpolate(image, [int(H * scale), int(W * scale)], mode="bilinear", align_corners=False)
    prob = crop_eval(model, scaled, n_classes, cropsize, flip)
    return F.interpolate(prob, (H, W), mode="bilinear", align_corners=False)


@torch.no_grad()
def evaluate(model, batches, n_classes, ignore_label=255, scales=(1.0,), flip=False, cropsize=1024):
    """reference: evaluate.py:193-253 (single process; the dist.reduce is a plain sum of per-rank hists)."""
    hist = np.zeros((n_classes, n_classes), dtype=np.float64)
    for images, labels in batches:
        labels_np = labels.cpu().numpy()
        if labels_np.ndim == 4:
            labels_np = labels_np.squeeze(1)
        probs = torch.zeros((images.size(0), n_classes, *images.shape[-2:]), device=images.device)
        for s in scales:
            probs += scale_crop_eval(model, images, s, n_classes, cropsize, flip)
        preds = torch.argmax(probs, dim=1).cpu().numpy()
        for i in range(labels_np.shape[0]):
            hist += compute_hist(preds[i], labels_np[i], n_classes, ignore_label)
    return metrics(hist)
