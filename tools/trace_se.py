"""SE hard-swish blocks: scale_act + project vs the fused A-operand prologue (engine.fuse_se), per-layer CUDA-event times."""
import sys

import torch

sys.path.insert(0, ".")
from cabinet_b200.synthetic import build_model, make_input  # noqa: E402

model = build_model(8, "large").cuda()
model.logits_dtype = torch.bfloat16
x = make_input(16, 1024, 1024).cuda()
eng = model.engine()
for fuse in (False, True):
    eng.fuse_se = fuse
    for _ in range(2):
        model(x)
    eng.start_trace()
    for _ in range(3):
        model(x)
    rows = eng.stop_trace()
    n = len(rows) // 3
    tot = 0.0
    print(f"fuse_se={fuse}: {sum(r['ms'] for r in rows) / 3:.3f} ms traced")
    for i in range(n):
        r = rows[2 * n + i]
        if any(f"mobile.f{b}." in r["layer"] for b in (4, 11, 12, 13, 14, 15)) and (r["kernel"] == "scale_act" or "project" in r["layer"]):
            ms = min(rows[j * n + i]["ms"] for j in range(3))
            tot += ms
            print(f"  {r['kernel']:12s} {r['layer']:24s} {ms * 1e3:8.1f} us")
    print(f"  sum {tot * 1e3:.1f} us")
