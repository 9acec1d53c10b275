"""Generates ``tests/golden/ohem_*.npz`` from the UNMODIFIED reference ``OhemCELoss`` (/root/reference/src/utils/loss.py).

Build container only:  python oracle/make_golden_loss.py
Each fixture holds the seeded inputs' parameters, the reference loss and the reference gradient w.r.t. the logits.
"""

from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.dont_write_bytecode = True

from oracle.loss_oracle import OHEM_CASES, make_case  # noqa: E402


def main():
    sys.path.insert(0, "/root/reference")
    from src.utils.loss import OhemCELoss

    out = ROOT / "tests" / "golden"
    for name, shape, thresh, n_min, ignore, weighted, scale, quant in OHEM_CASES:
        logits, labels, weight = make_case(shape, ignore, weighted, scale, quant)
        x = logits.clone().requires_grad_(True)
        crit = OhemCELoss(thresh=thresh, n_min=n_min, ignore_lb=255, weight=weight)
        loss = crit(x, labels)
        loss.backward()
        grad = x.grad if x.grad is not None else torch.zeros_like(x)
        np.savez_compressed(out / f"ohem_{name}.npz", loss=np.float64(loss.item()), grad=grad.numpy(),
                            logits_sum=np.float64(logits.double().sum().item()), labels_sum=np.int64(labels.sum().item()))
        print(name, float(loss), float(grad.abs().sum()))


if __name__ == "__main__":
    main()
