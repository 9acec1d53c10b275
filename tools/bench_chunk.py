"""End-to-end fast-mode evaluation (fp32 pinned host batches) for several upload chunk sizes."""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from cabinet_b200.evaluator import MscEvalV0  # noqa: E402
from cabinet_b200.synthetic import build_model, make_input, make_labels  # noqa: E402

B, S, C, K = 16, 1024, 8, 10
model = build_model(C, "large").cuda()
model.logits_dtype = torch.bfloat16
model.use_cuda_graph = True
x = make_input(B, S, S, seed=7).pin_memory()
lb = make_labels(B, S, S, C, seed=11).to(torch.uint8).pin_memory()
masks = [torch.empty((B, S, S), dtype=torch.uint8).pin_memory() for _ in range(K)]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for chunk in (8, 16):
    ev = MscEvalV0(model, [(x, lb)] * 4, C, 255, (1.0,), False, cropsize=S)
    ev.chunk = chunk
    ev.evaluate(masks_out=masks)
    ev.dl = [(x, lb)] * K
    ts = []
    for _ in range(5):
        torch.cuda.synchronize()
        e0.record()
        ev.evaluate(masks_out=masks)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    best = min(ts)
    print(f"chunk {chunk:2d}: {best / K:.3f} ms/step  {B * K / best * 1e3:.0f} img/s   all runs (ms/step): "
          + " ".join(f"{t / K:.3f}" for t in ts), flush=True)
    del ev
    model.engine()._graphs.clear()
