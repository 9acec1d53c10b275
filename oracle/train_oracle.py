"""TEST INFRASTRUCTURE ONLY -- CPU restatement of one training step's forward + backward (BASELINE config 5).

reference: ``src/scripts/train.py:430-441`` (``out, out16 = net(im)``; ``loss = criteria_p(out, lb) + criteria_16(out16, lb)``;
``loss.backward()``) with the model in ``.train()`` (batch-statistics BatchNorm, running statistics updated with
momentum 0.1 and the unbiased variance, ``torch.nn.BatchNorm2d`` defaults).  The forward is ``oracle/cabinet_oracle.py``
with train-mode BN switched on, the loss ``oracle/loss_oracle.py``, the gradients plain autograd of those.

Pinned by ``oracle/make_golden_train.py`` -> ``tests/golden/train_step_*.npz`` against the imported reference
(loss, gradient norm of every parameter, full gradients of four layers, updated running statistics of four BN layers).
This is the parity target of the training-step kernels; nothing under ``cabinet_b200/`` imports it.
"""

from __future__ import annotations

import torch

from oracle import cabinet_oracle
from oracle.loss_oracle import ohem_ce_loss

BN_MOMENTUM = 0.1

TRAIN_CASES = [  # name, mode, classes, (N, H, W), thresh, n_min
    ("small_c8_2x64x96", "small", 8, (2, 64, 96), 0.7, 2 * 64 * 96 // 16),
    ("large_c5_2x64x64", "large", 5, (2, 64, 64), 0.7, 2 * 64 * 64 // 16),
]
FULL_GRAD_KEYS = ("conv_out.conv_out.weight", "sb.conv1.conv.weight", "ab.a2block.gamma", "mobile.features.1.conv.0.weight")
STAT_KEYS = ("sb.conv1.bn", "mobile.features.0.1", "ffm.convblk.bn", "ab.b2")


def train_step(sd: dict, x: torch.Tensor, labels: torch.Tensor, cfgs, thresh: float, n_min: int, ignore_lb: int = 255):
    """-> (loss, {param: grad}, {bn prefix: (new running_mean, new running_var)}) for one forward + backward."""
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()
              if v.is_floating_point() and not k.endswith(("running_mean", "running_var"))}
    work = dict(sd)
    work.update(params)
    cabinet_oracle.BATCH_STATS = {}
    try:
        out, out16 = cabinet_oracle.cabinet_forward_graph(work, x.float(), cfgs)
        stats = cabinet_oracle.BATCH_STATS
    finally:
        cabinet_oracle.BATCH_STATS = None
    loss = ohem_ce_loss(out, labels, thresh, n_min, ignore_lb) + ohem_ce_loss(out16, labels, thresh, n_min, ignore_lb)
    loss.backward()
    grads = {k: (p.grad if p.grad is not None else None) for k, p in params.items()}
    running = {}
    for prefix, (mean, var, _) in stats.items():
        running[prefix] = ((1 - BN_MOMENTUM) * sd[prefix + ".running_mean"] + BN_MOMENTUM * mean,
                           (1 - BN_MOMENTUM) * sd[prefix + ".running_var"] + BN_MOMENTUM * var)
    return loss.detach(), grads, running
