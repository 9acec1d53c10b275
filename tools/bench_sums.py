"""expand_sums timing vs nsplit (CUDA events around 20 back-to-back launches): python tools/bench_sums.py"""
import sys

import torch

sys.path.insert(0, ".")
from cabinet_b200.engine import Map  # noqa: E402
from cabinet_b200.synthetic import build_model  # noqa: E402

model = build_model(8, "large").cuda()
eng = model.engine()
blk = {i + 1: b for i, b in enumerate(eng.blocks)}
for bi, (h, w, c), splits in ((5, (128, 128, 40), (1, 2, 4, 9, 18, 36, 64)), (12, (64, 64, 112), (1, 2, 4, 8, 16)),
                              (11, (64, 64, 80), (1, 2, 4, 8, 16)), (14, (32, 32, 160), (1, 2, 4))):
    t = torch.randn(16, h, w, c, device="cuda").to(torch.bfloat16)
    x = Map(t, 16, h, w, c, c)
    e = blk[bi]
    gap = torch.zeros(16 * e["dw"].c, dtype=torch.int64, device="cuda")
    for ns in splits:
        def run():
            eng._run("expand_sums", "x", 0, 0, eng.lib.cabinet_expand_sums, x.ptr, x.ld, 16, h, w, c, e["w1t"].data_ptr(),
                     e["auxt"].data_ptr(), e["dw"].c, e["pw1"].act, e["dw"].k, ns, gap.data_ptr(), eng.stream)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            run()
        e1.record()
        torch.cuda.synchronize()
        print(f"f{bi} nsplit {ns:3d}: {e0.elapsed_time(e1) / 20 * 1e3:7.1f} us")
