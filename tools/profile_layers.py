"""Runs a handful of representative layers of BASELINE configs[1] standalone (warm-up + 1 profiled launch each)."""
import sys

import torch

sys.path.insert(0, ".")
from cabinet_b200.engine import Map  # noqa: E402
from cabinet_b200.synthetic import build_model  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
which = sys.argv[2].split(",") if len(sys.argv) > 2 else None
model = build_model(8, "large").cuda()
eng = model.engine()


def rnd(n, h, w, c):
    t = torch.randn(n, h, w, c, device="cuda").to(torch.bfloat16)
    return Map(t, n, h, w, c, c)


blk = {i + 1: b for i, b in enumerate(eng.blocks)}
jobs = {
    "f2.expand": lambda: eng.conv(rnd(B, 512, 512, 16), blk[2]["pw1"]),
    "f1.project": lambda: eng.conv(rnd(B, 512, 512, 16), blk[1]["pw2"]),
    "f3.project": lambda: eng.conv(rnd(B, 256, 256, 72), blk[3]["pw2"]),
    "f7.expand": lambda: eng.conv(rnd(B, 128, 128, 40), blk[7]["pw1"]),
    "f12.expand": lambda: eng.conv(rnd(B, 64, 64, 112), blk[12]["pw1"]),
    "ffm.convblk": lambda: eng.conv(rnd(B, 128, 128, 384), eng.ffm_blk),
    "conv_out.conv": lambda: eng.conv(rnd(B, 128, 128, 256), eng.head_conv),
    "sb.conv2": lambda: eng.conv(rnd(B, 512, 512, 64), eng.sb2),
    "ab.b1": lambda: eng.conv(rnd(B, 32, 32, 1216), eng.b1),
    "f2.dw": lambda: eng.dwconv(rnd(B, 512, 512, 64), blk[2]["dw"]),
    "f5.dw": lambda: eng.dwconv(rnd(B, 128, 128, 120), blk[5]["dw"]),
    "f12.dw": lambda: eng.dwconv(rnd(B, 64, 64, 672), blk[12]["dw"]),
    "f14.dw": lambda: eng.dwconv(rnd(B, 32, 32, 960), blk[14]["dw"]),
    "f2.fused": lambda: eng.mbconv_fused(rnd(B, 512, 512, 16), blk[2], None),
    "f3.fused": lambda: eng.mbconv_fused(rnd(B, 256, 256, 24), blk[3], None),
    "f5.fused": lambda: eng.mbconv_fused(rnd(B, 128, 128, 40), blk[5], torch.zeros(B * 120, dtype=torch.int64, device="cuda")),
    "f7.fused": lambda: eng.mbconv_fused(rnd(B, 128, 128, 40), blk[7], None),
}
for name, fn in jobs.items():
    if which and name not in which:
        continue
    fn()
    torch.cuda.synchronize()
    fn()
    torch.cuda.synchronize()
    print("ran", name)
