"""OHEM loss forward + backward at the BASELINE config-5 shape: C-ABI kernels vs the sort-based torch restatement on the
same GPU (context only; the reference itself would run exactly these torch ops)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from cabinet_b200.loss import OhemCELoss  # noqa: E402
from oracle.loss_oracle import ohem_ce_loss  # noqa: E402

N, C, H, W = 8, 8, 1024, 1024
g = torch.Generator(device="cuda").manual_seed(3)
labels = torch.randint(0, C, (N, H, W), device="cuda", generator=g)
labels[:, 100:140, :] = 255
n_min = N * H * W // 16
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for dtype in (torch.float32, torch.bfloat16):
    for mode, scale, boost in (("thresh", 3.0, 0.0), ("topk", 0.05, 6.0)):
        logits = (torch.randn(N, C, H, W, device="cuda", generator=g) * scale).to(dtype)
        if boost:
            logits.scatter_add_(1, labels.clamp(max=C - 1).unsqueeze(1), torch.full((N, 1, H, W), boost, device="cuda", dtype=dtype))
        crit = OhemCELoss(0.7, n_min, 255)
        for name, fn in (("kernels", lambda x: crit(x, labels)), ("torch sort", lambda x: ohem_ce_loss(x, labels, 0.7, n_min, 255))):
            ts = []
            for it in range(6):
                x = logits.clone().requires_grad_(True)
                torch.cuda.synchronize()
                e0.record()
                loss = fn(x)
                loss.backward()
                e1.record()
                torch.cuda.synchronize()
                if it >= 2:
                    ts.append(e0.elapsed_time(e1))
            ms = sorted(ts)[len(ts) // 2]
            nbytes = logits.numel() * logits.element_size() * 3 + labels.numel() * 8 * 2  # fwd read + bwd read + grad write
            print(f"{str(dtype):15s} {mode:7s} {name:10s} fwd+bwd {ms:7.3f} ms  ({nbytes / ms / 1e6:6.0f} GB/s algorithmic) loss {float(loss):.6f}", flush=True)
