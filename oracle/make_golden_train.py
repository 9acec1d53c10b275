"""Generates ``tests/golden/train_step_*.npz`` from the UNMODIFIED reference model + loss in ``.train()`` mode.

Build container only:  python oracle/make_golden_train.py
"""

from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.dont_write_bytecode = True

from cabinet_b200.constants import BACKBONE_CFGS  # noqa: E402
from cabinet_b200.synthetic import make_input, make_labels, perturb_state_dict, state_dict_digest  # noqa: E402
from oracle.train_oracle import FULL_GRAD_KEYS, STAT_KEYS, TRAIN_CASES  # noqa: E402


def main():
    sys.path.insert(0, "/root/reference")
    from src.models.cabinet import CABiNet
    from src.utils.loss import OhemCELoss

    torch.set_num_threads(8)
    for name, mode, C, (N, H, W), thresh, n_min in TRAIN_CASES:
        torch.manual_seed(0)
        net = CABiNet(C, mode=mode, cfgs=BACKBONE_CFGS[mode])
        sd = perturb_state_dict(net.state_dict())
        net.load_state_dict(sd)
        net.train()
        x, lb = make_input(N, H, W), make_labels(N, H, W, C)
        crit_p, crit_16 = OhemCELoss(thresh, n_min, 255), OhemCELoss(thresh, n_min, 255)
        out, out16 = net(x)
        loss = crit_p(out, lb) + crit_16(out16, lb)
        loss.backward()
        named = dict(net.named_parameters())
        norms = {k: (float(p.grad.norm()) if p.grad is not None else -1.0) for k, p in named.items()}
        after = net.state_dict()
        np.savez_compressed(
            ROOT / "tests" / "golden" / f"train_step_{name}.npz", digest=state_dict_digest(sd), loss=np.float64(loss.item()),
            grad_keys=np.array(list(norms)), grad_norms=np.array(list(norms.values()), dtype=np.float64),
            **{"grad__" + k: named[k].grad.numpy() for k in FULL_GRAD_KEYS},
            **{"mean__" + k: after[k + ".running_mean"].numpy() for k in STAT_KEYS},
            **{"var__" + k: after[k + ".running_var"].numpy() for k in STAT_KEYS})
        print(name, float(loss), sum(1 for v in norms.values() if v < 0), "params without grad")


if __name__ == "__main__":
    main()
