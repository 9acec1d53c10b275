import ctypes, sys, torch
sys.path.insert(0, ".")
from cabinet_b200 import _lib
from cabinet_b200.engine import Map
from cabinet_b200.synthetic import build_model
model = build_model(8, "large").cuda(); eng = model.engine(); lib = _lib.load()
B = 16
def rnd(n, h, w, c):
    t = torch.randn(n, h, w, c, device="cuda").to(torch.bfloat16); return Map(t, n, h, w, c, c)
jobs = {"sb.conv2": (lambda x: eng.conv(x, eng.sb2), (B, 512, 512, 64)), "conv_out.conv": (lambda x: eng.conv(x, eng.head_conv), (B, 128, 128, 256)),
        "f2.expand": (lambda x: eng.conv(x, eng.blocks[1]["pw1"]), (B, 512, 512, 16))}
for name, (fn, shp) in jobs.items():
    x = rnd(*shp); fn(x); torch.cuda.synchronize()
    for flags in (8, 12):
        lib.cabinet_debug_flags(flags); fn(x); torch.cuda.synchronize(); lib.cabinet_debug_flags(0)
        buf = (ctypes.c_longlong * 8000)(); lib.cabinet_debug_read(buf, 8000)
        import numpy as np
        a = np.array(buf[:8000]).reshape(-1, 4)[40:400]
        wait = a[:, 1] - a[:, 0]; issue = a[:, 2] - a[:, 1]; commit = a[:, 3] - a[:, 2]; period = np.diff(a[:, 0])
        print(f"{name} flags={flags}: per k-block cycles: wait(full) med {np.median(wait):.0f} mean {wait.mean():.0f} | issue med {np.median(issue):.0f} | commit med {np.median(commit):.0f} | period med {np.median(period):.0f} mean {period.mean():.0f}")
