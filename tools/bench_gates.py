"""True cost of the SE / FFM gate kernels inside a CUDA graph (back-to-back replay): python tools/bench_gates.py"""
import sys

import torch

sys.path.insert(0, ".")
from cabinet_b200.synthetic import build_model  # noqa: E402

B = 16
model = build_model(8, "large").cuda()
eng = model.engine()
ses = [(e["se"], e["dw"].c, e["dw"].name) for e in eng.blocks if "se" in e] + [(eng.ffm_gate, 256, "ffm")]
gaps = [torch.rand(B, c, device="cuda") for _, c, _ in ses]


def run_all():
    for (g, c, name), gap in zip(ses, gaps):
        eng.gate(gap, 4096, g, name)


def timeit(fn, name):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per replay")


timeit(run_all, f"{2 * len(ses)} gate_fc launches")
for (g, c, name), gap in zip(ses, gaps):
    timeit(lambda: eng.gate(gap, 4096, g, name), f"  {name} C={c} Cmid={g.cmid} (2 launches)")
