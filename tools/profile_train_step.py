"""One bf16 training step (Large, batch 8, 1024x1024) for ncu captures: `python tools/profile_train_step.py [steps]`."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from cabinet_b200.loss import OhemCELoss
from cabinet_b200.synthetic import build_model, make_input, make_labels

B, S, C = 8, 1024, 8
model = build_model(C, "large").cuda().train()
model.train_precision = "bf16"
model.logits_dtype = torch.bfloat16
x, lb = make_input(B, S, S).cuda(), make_labels(B, S, S, C).cuda()
crit = OhemCELoss(0.7, B * S * S // 16, 255)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    model.zero_grad(set_to_none=True)
    out, out16 = model(x)
    (crit(out, lb) + crit(out16, lb)).backward()
torch.cuda.synchronize()
print("done")
