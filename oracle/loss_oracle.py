"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference training loss (``src/utils/loss.py:38-80``).

Pinned against the imported, unmodified reference by ``oracle/make_golden_loss.py`` -> ``tests/golden/ohem_*.npz``
(``tests/test_oracle_golden.py``).  Nothing under ``cabinet_b200/`` imports this module.
"""

from __future__ import annotations

import torch
import torch.nn.functional as F


def ohem_ce_loss(logits: torch.Tensor, labels: torch.Tensor, thresh: float, n_min: int, ignore_lb: int = 255,
                 weight: torch.Tensor | None = None) -> torch.Tensor:
    """(N,C,H,W) logits, (N,H,W) labels -> scalar (differentiable).  reference: loss.py:38-80."""
    loss = F.cross_entropy(logits.float(), labels, weight=weight, ignore_index=ignore_lb, reduction="none")  # :51-57
    valid = loss[labels != ignore_lb]                                                                      # :60-61
    if valid.numel() == 0:                                                                                 # :64-65
        return torch.zeros((), requires_grad=True)
    srt, _ = torch.sort(valid, descending=True)                                                            # :68
    k = min(int(n_min), srt.numel())                                                                       # :71
    if srt[k - 1] > thresh:                                                                                # :74-75
        return srt[srt > thresh].mean()
    return srt[:k].mean()                                                                                  # :76-80


OHEM_CASES = [
    # name, (N, C, H, W), thresh, n_min, ignore stripe, class weights, logit scale, quantise (ties)
    ("thresh_set", (2, 19, 24, 32), 0.7, 100, True, False, 1.0, 0.0),        # many losses above thresh
    ("topk", (2, 8, 24, 32), 6.0, 300, True, False, 1.0, 0.0),               # almost none above thresh: k largest
    ("topk_all", (1, 5, 16, 20), 50.0, 100000, True, False, 1.0, 0.0),       # n_min > #valid: every valid pixel
    ("weighted", (2, 8, 16, 24), 0.7, 64, True, True, 2.0, 0.0),
    ("weighted_topk", (2, 8, 16, 24), 9.0, 200, False, True, 1.0, 0.0),
    ("ties_topk", (1, 4, 16, 16), 9.0, 77, False, False, 1.0, 1.0),          # quantised logits: k-th value inside a tie group
    ("all_ignored", (1, 6, 8, 8), 0.7, 10, None, False, 1.0, 0.0),
]


def make_case(shape, ignore, weighted, scale, quant, seed=31):
    N, C, H, W = shape
    g = torch.Generator().manual_seed(seed + C + H)
    logits = torch.randn(N, C, H, W, generator=g) * scale
    if quant:
        logits = torch.round(logits / quant) * quant
    labels = torch.randint(0, C, (N, H, W), generator=g)
    if ignore is None:
        labels[:] = 255
    elif ignore:
        labels[:, H // 4: H // 4 + 2, :] = 255
        labels[0, :, ::7] = 255
    weight = (0.5 + torch.rand(C, generator=g)) if weighted else None
    return logits, labels, weight
