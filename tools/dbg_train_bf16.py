"""Per-parameter gradient error of the bf16-activation training step against the CPU train oracle (debug)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from cabinet_b200.constants import BACKBONE_CFGS
from cabinet_b200.loss import OhemCELoss
from cabinet_b200.synthetic import build_model, make_input, make_labels
from oracle.train_oracle import train_step

mode, C, N, H, W = "large", 6, 2, 96, 80
thresh, n_min = 0.7, N * H * W // 16
sd = {k: v.clone() for k, v in build_model(C, mode).state_dict().items()}
x, lb = make_input(N, H, W), make_labels(N, H, W, C)
loss_ref, grads_ref, _ = train_step(sd, x, lb, BACKBONE_CFGS[mode], thresh, n_min)
for precision in sys.argv[1:] or ["bf16"]:
    model = build_model(C, mode).cuda().train()
    model.train_precision = precision
    out, out16 = model(x.cuda())
    loss = OhemCELoss(thresh, n_min, 255)(out, lb.cuda()) + OhemCELoss(thresh, n_min, 255)(out16, lb.cuda())
    loss.backward()
    print(precision, "loss", float(loss), float(loss_ref))
    for k, p in model.named_parameters():
        if grads_ref.get(k) is None:
            continue
        g, r = p.grad.cpu(), grads_ref[k]
        print(f"{k:55s} {float((g - r).norm() / (r.norm() + 1e-20)):.3e}  |ref| {float(r.norm()):.3e}")
