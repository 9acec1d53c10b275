"""Micro-benchmark of single layers (CUDA events over R back-to-back launches): python tools/bench_layers.py [flags...]"""
import sys

import torch

sys.path.insert(0, ".")
from cabinet_b200 import _lib  # noqa: E402
from cabinet_b200.engine import Map  # noqa: E402
from cabinet_b200.synthetic import build_model  # noqa: E402

B = 16
flags = [int(a) for a in sys.argv[1:]] or [0]
model = build_model(8, "large").cuda()
eng = model.engine()
lib = _lib.load()


def rnd(n, h, w, c):
    t = torch.randn(n, h, w, c, device="cuda").to(torch.bfloat16)
    return Map(t, n, h, w, c, c)


blk = {i + 1: b for i, b in enumerate(eng.blocks)}
jobs_all = {
    "f2.expand 16->64@512": (lambda x: eng.conv(x, blk[2]["pw1"]), (B, 512, 512, 16), 64),
    "f1.project 16->16@512": (lambda x: eng.conv(x, blk[1]["pw2"]), (B, 512, 512, 16), 16),
    "f3.project 72->24@256": (lambda x: eng.conv(x, blk[3]["pw2"]), (B, 256, 256, 72), 24),
    "f7.expand 40->240@128": (lambda x: eng.conv(x, blk[7]["pw1"]), (B, 128, 128, 40), 240),
    "f12.expand 112->672@64": (lambda x: eng.conv(x, blk[12]["pw1"]), (B, 64, 64, 112), 672),
    "ffm.convblk 384->256@128": (lambda x: eng.conv(x, eng.ffm_blk), (B, 128, 128, 384), 256),
    "conv_out.conv 3x3 256->256@128": (lambda x: eng.conv(x, eng.head_conv), (B, 128, 128, 256), 256),
    "sb.conv2 3x3s2 64->64@512": (lambda x: eng.conv(x, eng.sb2), (B, 512, 512, 64), 64 / 4),
    "f2.dw 3x3s2 64@512": (lambda x: eng.dwconv(x, blk[2]["dw"]), (B, 512, 512, 64), 64 / 4),
    "f5.dw 5x5 120@128": (lambda x: eng.dwconv(x, blk[5]["dw"]), (B, 128, 128, 120), 120),
    "f12.dw 3x3 672@64": (lambda x: eng.dwconv(x, blk[12]["dw"]), (B, 64, 64, 672), 672),
    "mb.f2 16->64->24 s2@512": (lambda x: eng.mbconv_fused(x, blk[2], None), (B, 512, 512, 16), 24 / 4),
    "mb.f3 24->72->24@256": (lambda x: eng.mbconv_fused(x, blk[3], None), (B, 256, 256, 24), 24),
    "mb.f5 40->120 k5@128": (lambda x: eng.mbconv_fused(x, blk[5], torch.zeros(B, 120, device="cuda")), (B, 128, 128, 40), 120),
    "mb.f7 40->240->80 s2@128": (lambda x: eng.mbconv_fused(x, blk[7], None), (B, 128, 128, 40), 80 / 4),
    "f12.project 672->112@64": (lambda x: eng.conv(x, blk[12]["pw2"]), (B, 64, 64, 672), 112),
    "f12.project+SE 672->112@64": (lambda x: eng.conv(x, blk[12]["pw2"], a_scale=torch.ones(B, 672, device="cuda"), a_act=3), (B, 64, 64, 672), 112),
    "f5.project 120->40@128": (lambda x: eng.conv(x, blk[5]["pw2"]), (B, 128, 128, 120), 40),
    "f5.project+SE 120->40@128": (lambda x: eng.conv(x, blk[5]["pw2"], a_scale=torch.ones(B, 120, device="cuda"), a_act=1), (B, 128, 128, 120), 40),
    "f14.dw 5x5 960@32": (lambda x: eng.dwconv(x, blk[14]["dw"]), (B, 32, 32, 960), 960),
}
import os
jobs = {k: v for k, v in jobs_all.items() if any(t in k for t in os.environ.get("LAYERS", "f2.expand,ffm.convblk,conv_out.conv").split(","))}
R = 5
for name, (fn, shp, cout) in jobs.items():
    x = rnd(*shp)
    pix = shp[0] * shp[1] * shp[2]
    nbytes = pix * (shp[3] + cout) * 2
    line = f"{name:32s} {nbytes / 1e6:7.1f} MB "
    for fl in flags:
        lib.cabinet_debug_flags(fl)
        fn(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(R):
            fn(x)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / R
        line += f"| flags {fl}: {ms * 1e3:7.1f} us {nbytes / ms / 1e6:6.0f} GB/s "
    lib.cabinet_debug_flags(0)
    print(line)
