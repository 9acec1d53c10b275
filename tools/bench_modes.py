"""Device-resident forward throughput vs (cuda graph, sub-batch): python tools/bench_modes.py [B]"""
import sys

import torch

sys.path.insert(0, ".")
from cabinet_b200.synthetic import build_model, make_input  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
model = build_model(8, "large").cuda()
model.logits_dtype = torch.bfloat16
x = make_input(B, 1024, 1024).cuda()
for graph in (True,):
    for sb in (0, 8, 4, 2, 1):
        model.use_cuda_graph, model.sub_batch = graph, sb
        for _ in range(3):
            model(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K = 10
        e0.record()
        for _ in range(K):
            model(x)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        print(f"graph={graph} sub_batch={sb or B}: {ms:.3f} ms/step  {B / ms * 1e3:.0f} img/s", flush=True)
