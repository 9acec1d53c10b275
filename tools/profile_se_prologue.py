"""f12.project (672 -> 112, 64x64, batch 16, + identity) with the SE A-operand prologue vs scale_act + plain GEMM."""
import sys

import torch

sys.path.insert(0, ".")
from cabinet_b200 import _lib  # noqa: E402
from cabinet_b200._lib import ACT_HSWISH, ACT_NONE, BF16, check  # noqa: E402

lib = _lib.load()
N, H, W, Cin, Cout = 16, 64, 64, 672, 112
st = torch.cuda.current_stream().cuda_stream
x = torch.randn(N, H, W, Cin, device="cuda").bfloat16()
res = torch.randn(N, H, W, Cout, device="cuda").bfloat16()
y = torch.empty(N, H, W, Cout, device="cuda", dtype=torch.bfloat16)
scale = torch.rand(N, Cin, device="cuda")
n16, c64 = -(-Cout // 16) * 16, -(-Cin // 64) * 64
w = torch.zeros(n16, 1, c64, device="cuda")
w[:Cout, :, :Cin] = torch.randn(Cout, 1, Cin, device="cuda") * Cin ** -0.5
w = w.bfloat16().contiguous()
b = torch.randn(Cout, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def fused():
    check(lib.cabinet_conv_tc_se(x.data_ptr(), Cin, N, H, W, Cin, scale.data_ptr(), ACT_HSWISH, w.data_ptr(), Cout, 1, 1, 1, 0,
                                 b.data_ptr(), res.data_ptr(), Cout, y.data_ptr(), BF16, Cout, H, W, ACT_NONE, st), "se")


def plain():
    check(lib.cabinet_scale_act(x.data_ptr(), Cin, BF16, scale.data_ptr(), N, H * W, Cin, ACT_HSWISH, 0, st), "sa")
    check(lib.cabinet_conv_tc(x.data_ptr(), Cin, N, H, W, Cin, w.data_ptr(), Cout, 1, 1, 1, 0, b.data_ptr(), res.data_ptr(),
                              Cout, y.data_ptr(), BF16, Cout, H, W, ACT_NONE, st), "tc")


for name, fn in (("fused prologue", fused), ("scale_act + gemm", plain)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:18s} {e0.elapsed_time(e1) * 100:.1f} us", flush=True)
