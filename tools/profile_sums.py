"""expand_sums on the Large f5 / f12 shapes (batch 16) for ncu: python tools/profile_sums.py"""
import sys

import torch

sys.path.insert(0, ".")
from cabinet_b200.engine import Map  # noqa: E402
from cabinet_b200.synthetic import build_model  # noqa: E402

model = build_model(8, "large").cuda()
eng = model.engine()
blk = {i + 1: b for i, b in enumerate(eng.blocks)}
for bi, (h, w, c) in ((5, (128, 128, 40)), (12, (64, 64, 112))):
    t = torch.randn(16, h, w, c, device="cuda").to(torch.bfloat16)
    x = Map(t, 16, h, w, c, c)
    e = blk[bi]
    for _ in range(2):
        gap = torch.zeros(16 * e["dw"].c, dtype=torch.int64, device="cuda")
        eng._run("expand_sums", "x", 0, 0, eng.lib.cabinet_expand_sums, x.ptr, x.ld, 16, h, w, c, e["w1t"].data_ptr(),
                 e["auxt"].data_ptr(), e["dw"].c, e["pw1"].act, e["dw"].k, 0, gap.data_ptr(), eng.stream)
        torch.cuda.synchronize()
    print("ran", bi)
