"""Per-launch CUDA-event trace of one forward (batch 16, Large, 1024x1024): python tools/trace_layers.py [substr ...]"""
import sys

import torch

sys.path.insert(0, ".")
from cabinet_b200.synthetic import build_model, make_input  # noqa: E402

pats = sys.argv[1:]
model = build_model(8, "large").cuda()
model.logits_dtype = torch.bfloat16
x = make_input(16, 1024, 1024).cuda()
eng = model.engine()
for fold in (True, False):
    eng.fold_low_up = fold
    for _ in range(2):
        model(x)
    eng.start_trace()
    for _ in range(3):
        model(x)
    rows = eng.stop_trace()
    n = len(rows) // 3
    print(f"fold_low_up={fold}: {sum(r['ms'] for r in rows) / 3:.3f} ms traced")
    for i in range(n):
        r = rows[2 * n + i]
        if not pats or any(p in r["layer"] or p in r["kernel"] for p in pats):
            ms = min(rows[j * n + i]["ms"] for j in range(3))
            print(f"  {r['kernel']:22s} {r['layer']:28s} {ms * 1e3:8.1f} us  {r['bytes'] / ms / 1e6 if ms else 0:7.0f} GB/s {r['flops'] / ms / 1e9 if ms else 0:7.0f} TF/s")
