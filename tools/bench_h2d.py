"""Host->device copy ceiling of the box (pinned fp32 batch of the bench workload)."""
import torch
x = torch.empty(16, 3, 1024, 1024, dtype=torch.float32).pin_memory()
lab = torch.empty(16, 1024, 1024, dtype=torch.uint8).pin_memory()
dx = torch.empty_like(x, device="cuda"); dl = torch.empty_like(lab, device="cuda")
s = torch.cuda.Stream()
for name, fn in [("images 201 MB", lambda: dx.copy_(x, non_blocking=True)),
                 ("images + labels 218 MB", lambda: (dx.copy_(x, non_blocking=True), dl.copy_(lab, non_blocking=True)))]:
    with torch.cuda.stream(s):
        fn(); s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(10): fn()
        e1.record(s); s.synchronize()
    ms = e0.elapsed_time(e1) / 10
    nbytes = x.numel() * 4 + (lab.numel() if "labels" in name else 0)
    print(f"{name}: {ms:.3f} ms per batch, {nbytes / ms / 1e6:.1f} GB/s -> ceiling {16 / ms * 1e3:.0f} img/s")
