"""Two forwards of BASELINE configs[1] (Large, 16x3x1024x1024, bf16): one warm-up, one to be profiled under ncu."""
import sys

import torch

sys.path.insert(0, ".")
from cabinet_b200.synthetic import build_model, make_input  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
model = build_model(8, "large").cuda()
model.logits_dtype = torch.bfloat16
x = make_input(B, 1024, 1024).cuda()
for _ in range(2):
    model(x)
    torch.cuda.synchronize()
print("launches per forward:", model.engine().launches)
