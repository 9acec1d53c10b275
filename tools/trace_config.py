"""Per-launch trace of one forward at any BASELINE shape: python tools/trace_config.py <mode> <batch> <H> <W> [classes]"""
import sys

import torch

sys.path.insert(0, ".")
from cabinet_b200.synthetic import build_model, make_input  # noqa: E402

mode, B, H, W = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
C = int(sys.argv[5]) if len(sys.argv) > 5 else 8
model = build_model(C, mode).cuda()
model.logits_dtype = torch.bfloat16
x = make_input(B, H, W).cuda()
eng = model.engine()
for _ in range(3):
    model(x)
eng.start_trace()
for _ in range(3):
    model(x)
rows = eng.stop_trace()
n = len(rows) // 3
tot = 0.0
for i in range(n):
    ms = sum(rows[i + j * n]["ms"] for j in range(3)) / 3
    r = rows[i]
    tot += ms
    floor = max(r["bytes"] / 6.5275e12, r["flops"] / 1.3695e15) * 1e3
    print(f"{r['kernel']:22s} {r['layer']:28s} {ms * 1e3:8.1f} us  floor {floor * 1e3:7.1f} us")
print(f"{mode} {B}x{H}x{W}: {tot:.3f} ms traced, {B / tot * 1e3:.1f} img/s")
