"""Phase timeline of the fused MBConv kernel (CTA 0, compute thread 0): python tools/dbg_mbconv.py [block]
Needs a library built with the stamps compiled in: make -C cabinet_b200/csrc clean all NVCCFLAGS+=-DCAB_MB_DEBUG"""
import sys

import torch

from cabinet_b200 import _lib
from cabinet_b200.synthetic import build_model, make_input

blk = sys.argv[1] if len(sys.argv) > 1 else "f2"
model = build_model(8, "large").cuda()
x = make_input(16, 1024, 1024).cuda()
eng = model.engine()
lib = _lib.load()
import ctypes
lib.cabinet_mbconv_debug.argtypes = [ctypes.c_void_p]
model(x)
torch.cuda.synchronize()
buf = torch.zeros(2048, dtype=torch.int64, device="cuda")
orig = lib.cabinet_mbconv_fused


class Hook:
    def __init__(self):
        self.i = 0

    def __call__(self, *a):
        self.i += 1
        name = {1: "f2", 2: "f3", 3: "f4", 4: "f5", 5: "f6", 6: "f7"}.get(self.i)
        lib.cabinet_mbconv_debug(buf.data_ptr() if name == blk else None)
        return orig(*a)


h = Hook()
eng.lib = type("L", (), {"__getattr__": lambda s, n: h if n == "cabinet_mbconv_fused" else getattr(lib, n)})()
model(x)
torch.cuda.synchronize()
lib.cabinet_mbconv_debug(None)
t = buf.cpu().view(-1, 16)
names = ["wait_d1", "epi1", "bar1", "dw", "bar2", "wait_d2", "epi2", "bar3"]
t0 = int(t[0, 0])
print("chunk  start   " + " ".join(f"{n:>8s}" for n in names))
for g in range(2, 26):
    r = t[g]
    if int(r[0]) == 0:
        break
    d = [int(r[i + 1] - r[i]) if int(r[i + 1]) and int(r[i]) else 0 for i in range(8)]
    ex = [int(r[i] - r[1]) if int(r[i]) else 0 for i in (9, 10, 11, 12, 13)] + [int(r[14] - r[8]) if int(r[14]) else 0]
    print(f"{g:4d} {int(r[0]) - t0:8d} " + " ".join(f"{v:8d}" for v in d) + "  | epi1 w0 p0 w1 p1 w2:" + " ".join(f"{v:6d}" for v in ex[:5]) + f" store {ex[5]}")
