"""Per-launch trace of one forward (CUDA events around every kernel): python tools/trace_json.py [batch] [size] [out.json]"""
import json
import sys

import torch

sys.path.insert(0, ".")

from cabinet_b200.synthetic import build_model, make_input

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
S = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
out = sys.argv[3] if len(sys.argv) > 3 else "gpurun_out/trace.json"
model = build_model(8, "large").cuda()
model.logits_dtype = torch.bfloat16
x = make_input(B, S, S).cuda()
eng = model.engine()
for _ in range(3):
    model(x)
eng.start_trace()
for _ in range(3):
    model(x)
rows = eng.stop_trace()
n = len(rows) // 3
agg = []
for i in range(n):
    r = dict(rows[i])
    r["ms"] = sum(rows[i + j * n]["ms"] for j in range(3)) / 3
    r["GB/s"] = r["bytes"] / (r["ms"] * 1e-3) / 1e9 if r["ms"] else 0
    r["TF/s"] = r["flops"] / (r["ms"] * 1e-3) / 1e12 if r["ms"] else 0
    agg.append(r)
tot = sum(r["ms"] for r in agg)
print(f"batch {B} size {S}: {tot:.3f} ms/step traced, {B / tot * 1e3:.1f} img/s")
for r in sorted(agg, key=lambda r: -r["ms"])[:45]:
    print(f"{r['kernel']:22s} {r['layer']:24s} {r['ms']:8.3f} ms {100 * r['ms'] / tot:5.1f}%  {r['GB/s']:8.1f} GB/s {r['TF/s']:7.2f} TF/s")
json.dump(agg, open(out, "w"), indent=0)
