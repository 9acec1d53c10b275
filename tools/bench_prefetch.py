"""Config-5 step fed from pinned host memory: copies on the compute stream vs DevicePrefetcher (same process, same box)."""
import sys, time, torch
sys.path.insert(0, ".")
from cabinet_b200.loss import OhemCELoss
from cabinet_b200.prefetch import DevicePrefetcher
from cabinet_b200.synthetic import build_model, make_input, make_labels
B, S, C, K = 8, 1024, 8, 10
dev = torch.device("cuda")
model = build_model(C, "large").cuda().train()
model.train_precision = "bf16"; model.logits_dtype = torch.bfloat16
xh, lh = make_input(B, S, S).pin_memory(), make_labels(B, S, S, C).pin_memory()
crit = OhemCELoss(0.7, B * S * S // 16, 255)
def step(x, lb):
    model.zero_grad(set_to_none=True)
    out, out16 = model(x)
    loss = crit(out, lb) + crit(out16, lb)
    loss.backward()
    return loss
for _ in range(4): step(xh.cuda(), lh.cuda())
for name in ("inline", "prefetch", "inline", "prefetch"):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    if name == "inline":
        for _ in range(K):
            lv = float(step(xh.to(dev, non_blocking=True), lh.to(dev, non_blocking=True)).item())
    else:
        for xd, ld in DevicePrefetcher(((xh, lh) for _ in range(K)), dev):
            lv = float(step(xd, ld).item())
    torch.cuda.synchronize(); ms = (time.perf_counter() - t0) / K * 1e3
    print(f"{name:9s} {ms:.2f} ms/step {B / ms * 1e3:.0f} img/s")
