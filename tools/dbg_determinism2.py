"""Run-to-run reproducibility of the full forward (bit equality of the bf16 logits over many runs)."""
import sys, torch
sys.path.insert(0, ".")
from cabinet_b200.synthetic import build_model, make_input
for size, B, runs in ((256, 4, 40), (1024, 8, 12)):
    model = build_model(8, "large").cuda()
    model.logits_dtype = torch.bfloat16
    model.use_cuda_graph = True
    x = make_input(B, size, size).cuda()
    ref = None; bad = 0; worst = 0.0
    for i in range(runs):
        f = model(x)[0].float().clone()
        if ref is None: ref = f
        else:
            d = float((f - ref).abs().max())
            if d > 0: bad += 1; worst = max(worst, d)
    print(f"size {size} batch {B}: {bad} of {runs - 1} repeats differ from the first run (max abs diff {worst})")
