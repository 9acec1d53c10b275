"""GPU: each C-ABI kernel against the same op in plain fp32 torch on the CPU (seeded inputs)."""

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from cabinet_b200 import _lib  # noqa: E402
from cabinet_b200._lib import ACT_HSIGMOID, ACT_HSWISH, ACT_NONE, ACT_RELU, ACT_SIGMOID, BF16, F32, check  # noqa: E402
from tests.gpu_util import from_map, rel_l2, to_map, tol  # noqa: E402

DTYPES = [torch.float32, torch.bfloat16]


def gen(*shape, seed=0, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


def stream():
    return torch.cuda.current_stream().cuda_stream


def dt(t):
    return BF16 if t == torch.bfloat16 else F32


def act_ref(x, act):
    return {ACT_NONE: lambda v: v, ACT_RELU: F.relu, ACT_HSWISH: lambda v: v * F.relu6(v + 3) / 6,
            ACT_HSIGMOID: lambda v: F.relu6(v + 3) / 6, ACT_SIGMOID: torch.sigmoid}[act](x)


def q(x, dtype):
    """Round through the storage dtype so the reference sees the same inputs as the kernel."""
    return x.to(dtype).float()


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("cin,cout,k,s,p,H,W,act,res", [
    (16, 64, 1, 1, 0, 17, 23, ACT_RELU, False),
    (24, 24, 1, 1, 0, 9, 31, ACT_NONE, True),
    (64, 64, 3, 2, 1, 33, 18, ACT_RELU, False),
    (72, 40, 3, 1, 1, 8, 8, ACT_HSWISH, False),
    (200, 80, 1, 1, 0, 5, 7, ACT_NONE, True),
])
def test_conv2d_simt(dtype, cin, cout, k, s, p, H, W, act, res):
    lib = _lib.load()
    N = 2
    x = q(gen(N, cin, H, W, seed=1), dtype)
    w = q(gen(cout, cin, k, k, seed=2, scale=(cin * k * k) ** -0.5), dtype)
    b = gen(cout, seed=3, scale=0.1)
    OH, OW = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    r = q(gen(N, cout, OH, OW, seed=4), dtype) if res else None
    ref = act_ref(F.conv2d(x, w, b, s, p), act)
    if res:
        ref = ref + r
    xm = to_map(x, dtype, ld=cin + 8, off=8)  # strided input (channel slice of a wider buffer)
    ym = to_map(torch.zeros(N, cout, OH, OW), dtype, ld=cout + 16, off=16)
    rm = to_map(r, dtype) if res else None
    wp = w.permute(0, 2, 3, 1).reshape(cout, -1).contiguous().to("cuda", dtype)
    bd = b.cuda()
    check(lib.cabinet_conv2d_simt(xm.ptr, xm.dt, H * W * xm.ld, W * xm.ld, xm.ld, 1, 0, wp.data_ptr(), dt(dtype),
                                  wp.shape[1], 1, 0, bd.data_ptr(), rm.ptr if res else None, rm.ld if res else 0,
                                  ym.ptr, ym.dt, ym.ld, 0, 1, N, H, W, cin, cout, k, k, s, p, OH, OW, act, 1.0,
                                  stream()), "conv")
    torch.cuda.synchronize()
    assert rel_l2(from_map(ym), ref) < tol(dtype)
    assert float(ym.t[..., :16].abs().max()) == 0  # the neighbouring slice is untouched


def test_conv2d_simt_reads_nchw_fp32_input():
    lib = _lib.load()
    x = gen(2, 3, 37, 41, seed=5)
    for (cout, k, s, p) in [(16, 3, 2, 1), (64, 7, 2, 3)]:
        w = gen(cout, 3, k, k, seed=6, scale=0.2)
        b = gen(cout, seed=7, scale=0.1)
        ref = F.relu(F.conv2d(x, w, b, s, p))
        OH, OW = ref.shape[2:]
        ym = to_map(torch.zeros_like(ref), torch.bfloat16)
        xd, wp, bd = x.cuda(), w.permute(0, 2, 3, 1).reshape(cout, -1).contiguous().cuda(), b.cuda()
        check(lib.cabinet_conv2d_simt(xd.data_ptr(), F32, 3 * 37 * 41, 41, 1, 37 * 41, 0, wp.data_ptr(), F32,
                                      wp.shape[1], 1, 0, bd.data_ptr(), None, 0, ym.ptr, BF16, ym.ld, 0, 1, 2, 37, 41,
                                      3, cout, k, k, s, p, OH, OW, ACT_RELU, 1.0, stream()), "stem")
        torch.cuda.synchronize()
        assert rel_l2(from_map(ym), ref) < 4e-3


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("C,k,s,H,W,act,gap", [
    (16, 3, 1, 13, 21, ACT_RELU, False), (64, 3, 2, 32, 32, ACT_RELU, False), (72, 5, 2, 19, 27, ACT_NONE, True),
    (120, 5, 1, 16, 9, ACT_NONE, True), (240, 3, 2, 7, 5, ACT_HSWISH, False), (960, 5, 1, 4, 4, ACT_NONE, True),
    (8, 3, 1, 1, 1, ACT_RELU, True),
])
def test_dwconv(dtype, C, k, s, H, W, act, gap):
    lib = _lib.load()
    N = 2
    x = q(gen(N, C, H, W, seed=1), dtype)
    w = gen(C, 1, k, k, seed=2, scale=0.3)
    b = gen(C, seed=3, scale=0.1)
    p = (k - 1) // 2
    ref = act_ref(F.conv2d(x, w, b, s, p, 1, C), act)
    OH, OW = ref.shape[2:]
    xm, ym = to_map(x, dtype), to_map(torch.zeros_like(ref), dtype)
    wp, bd = w.view(C, -1).t().contiguous().cuda(), b.cuda()
    g = torch.zeros(N, C, device="cuda") if gap else None
    check(lib.cabinet_dwconv(xm.ptr, xm.ld, wp.data_ptr(), bd.data_ptr(), ym.ptr, ym.ld, dt(dtype), N, H, W, C, k, s,
                             OH, OW, act, g.data_ptr() if gap else None, stream()), "dwconv")
    torch.cuda.synchronize()
    assert rel_l2(from_map(ym), ref) < tol(dtype)
    if gap:
        assert rel_l2(g.cpu(), ref.sum(dim=(2, 3))) < 1e-4


@pytest.mark.parametrize("C,Cmid,gate,bias", [(72, 24, ACT_HSIGMOID, True), (960, 240, ACT_HSIGMOID, True),
                                               (256, 64, ACT_SIGMOID, False)])
def test_gate_mlp_and_scale_act(C, Cmid, gate, bias):
    lib = _lib.load()
    N, HW = 3, 35
    sums = gen(N, C, seed=1) * HW
    w1, w2 = gen(Cmid, C, seed=2, scale=C ** -0.5), gen(C, Cmid, seed=3, scale=Cmid ** -0.5)
    b1, b2 = (gen(Cmid, seed=4, scale=0.1), gen(C, seed=5, scale=0.5)) if bias else (None, None)
    h = F.relu(F.linear(sums / HW, w1, b1))
    ref = act_ref(F.linear(h, w2, b2), gate)
    d = lambda t: None if t is None else t.cuda()  # noqa: E731
    sd, w1d, w2d, b1d, b2d = d(sums), d(w1), d(w2), d(b1), d(b2)
    out = torch.empty(N, C, device="cuda")
    check(lib.cabinet_gate_mlp(sd.data_ptr(), 1.0 / HW, w1d.data_ptr(), b1d.data_ptr() if bias else None,
                               w2d.data_ptr(), b2d.data_ptr() if bias else None, out.data_ptr(), N, C, Cmid, gate,
                               stream()), "gate")
    torch.cuda.synchronize()
    assert rel_l2(out.cpu(), ref) < 1e-5
    for dtype in DTYPES:
        x = q(gen(N, C, 5, 7, seed=6), dtype)
        for act, plus in [(ACT_HSWISH, False), (ACT_RELU, False), (ACT_NONE, True)]:
            xm = to_map(x, dtype)
            check(lib.cabinet_scale_act(xm.ptr, xm.ld, xm.dt, out.data_ptr(), N, 35, C, act, int(plus), stream()), "sa")
            torch.cuda.synchronize()
            sc = ref.view(N, C, 1, 1)
            want = x * sc + x if plus else act_ref(x * sc, act)
            assert rel_l2(from_map(xm), want) < tol(dtype)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("H,W", [(32, 32), (3, 4), (9, 13), (68, 120)])
def test_psp_pool_and_concat(dtype, H, W):
    lib = _lib.load()
    N, C = 2, 128
    x = q(gen(N, C, H, W, seed=1), dtype)
    pools = [F.adaptive_avg_pool2d(x, (s, s)) for s in (1, 3, 6, 8)]
    ref_pooled = torch.cat([p.permute(0, 2, 3, 1).reshape(N, -1, C) for p in pools], dim=1)
    ref_cat = torch.cat([x] + [F.interpolate(p, size=(H, W), mode="bilinear", align_corners=False) for p in pools], 1)
    xm = to_map(x, dtype)
    pooled = torch.empty(N, 110, C, device="cuda")
    cat = to_map(torch.zeros(N, 5 * C, H, W), dtype)
    scratch = torch.zeros(128 + N * (110 + 256 * C), device="cuda")
    for _ in range(2):  # the second launch checks that the tickets were left at zero
        pooled.fill_(-1.0)
        check(lib.cabinet_psp_pool(xm.ptr, xm.ld, xm.dt, pooled.data_ptr(), N, H, W, C, scratch.data_ptr(),
                                   scratch.numel() * 4, stream()), "pool")
    check(lib.cabinet_psp_concat(xm.ptr, xm.ld, pooled.data_ptr(), cat.ptr, cat.ld, xm.dt, N, H, W, C, stream()), "cat")
    torch.cuda.synchronize()
    assert rel_l2(pooled.cpu(), ref_pooled) < 1e-5
    assert rel_l2(from_map(cat), ref_cat) < tol(dtype)


@pytest.mark.parametrize("dtype", DTYPES)
def test_softmax_and_cab_combine(dtype):
    lib = _lib.load()
    s = gen(37, 300, seed=1, scale=3.0).cuda()
    p = torch.empty(37, 300, dtype=dtype, device="cuda")
    check(lib.cabinet_softmax_rows(s.data_ptr(), p.data_ptr(), dt(dtype), 37, 300, stream()), "softmax")
    g, x, r = (q(gen(2, 256, 6, 5, seed=i), dtype) for i in (2, 3, 4))
    gm, xm, rm = to_map(g, dtype), to_map(x, dtype), to_map(r, dtype)
    out = to_map(torch.zeros(2, 256, 6, 5), dtype, ld=256 + 64, off=64)
    gamma = torch.tensor([0.37], device="cuda")
    check(lib.cabinet_cab_combine(gm.ptr, xm.ptr, rm.ptr, out.ptr, out.ld, gamma.data_ptr(), dt(dtype), 2 * 30, 256,
                                  stream()), "combine")
    torch.cuda.synchronize()
    assert rel_l2(p.float().cpu(), F.softmax(s.cpu(), dim=-1)) < tol(dtype)
    assert rel_l2(from_map(out), 0.37 * g + x + x * torch.sigmoid(r)) < tol(dtype)


@pytest.mark.parametrize("dtype", DTYPES)
def test_channel_sum(dtype):
    lib = _lib.load()
    x = q(gen(3, 256, 20, 13, seed=1), dtype)
    xm = to_map(x, dtype)
    out = torch.zeros(3, 256, device="cuda")
    scratch = torch.zeros(256 + 3 * 64 * 256, device="cuda")
    for _ in range(2):  # the second launch checks that the tickets were left at zero
        out.fill_(-1.0)
        check(lib.cabinet_channel_sum(xm.ptr, xm.ld, xm.dt, 3, 260, 256, out.data_ptr(), scratch.data_ptr(),
                                      scratch.numel() * 4, stream()), "sum")
    torch.cuda.synchronize()
    assert rel_l2(out.cpu(), x.sum(dim=(2, 3))) < 1e-5


@pytest.mark.parametrize("C,IH,IW,OH,OW", [(256, 4, 4, 16, 16), (19, 3, 4, 9, 13), (8, 2, 3, 17, 33), (8, 1, 1, 8, 8),
                                               (16, 5, 7, 20, 28), (8, 1, 1, 4, 4), (24, 32, 32, 128, 128)])
def test_bilinear_nhwc(C, IH, IW, OH, OW):
    lib = _lib.load()
    for din, dout in [(torch.bfloat16, torch.bfloat16), (torch.float32, torch.float32)]:
        x = q(gen(2, C, IH, IW, seed=1), din)
        ref = F.interpolate(x, size=(OH, OW), mode="bilinear", align_corners=False)
        xm = to_map(x, din)
        ym = to_map(torch.zeros_like(ref), dout, ld=C + 8, off=4 if dout == torch.float32 else 8)
        check(lib.cabinet_bilinear_nhwc(xm.ptr, xm.ld, xm.dt, ym.ptr, ym.ld, ym.dt, 2, IH, IW, C, OH, OW, stream()), "bl")
        torch.cuda.synchronize()
        assert rel_l2(from_map(ym), ref) < tol(dout)


@pytest.mark.parametrize("C,IH,IW,OH,OW", [(8, 16, 16, 128, 128), (19, 9, 13, 70, 100), (8, 5, 7, 37, 51)])
def test_logits_tail_nchw_argmax_hist(C, IH, IW, OH, OW):
    lib = _lib.load()
    N = 2
    x = gen(N, C, IH, IW, seed=1)
    ref = F.interpolate(x, size=(OH, OW), mode="bilinear", align_corners=False)
    xd = x.permute(0, 2, 3, 1).contiguous().cuda()
    for odt in (torch.float32, torch.bfloat16):
        y = torch.empty(N, C, OH, OW, dtype=odt, device="cuda")
        check(lib.cabinet_upsample_logits_nchw(xd.data_ptr(), N, IH, IW, C, y.data_ptr(), dt(odt), OH, OW, stream()), "up")
        torch.cuda.synchronize()
        assert rel_l2(y.float().cpu(), ref) < (1e-5 if odt == torch.float32 else 4e-3)
        if odt == torch.float32:
            y32 = y.cpu()
    labels = torch.randint(0, C, (N, OH, OW), generator=torch.Generator().manual_seed(3))
    labels[:, OH // 3, :] = 255
    mask = torch.empty(N, OH, OW, dtype=torch.uint8, device="cuda")
    hist = torch.zeros(C, C, dtype=torch.int64, device="cuda")
    for ldt, lab in ((0, labels.cuda()), (1, labels.to(torch.uint8).cuda())):
        check(lib.cabinet_upsample_argmax(xd.data_ptr(), N, IH, IW, C, mask.data_ptr(), OH, OW, lab.data_ptr(), ldt, 255,
                                          hist.data_ptr(), stream()), "argmax")
    torch.cuda.synchronize()
    # bit-exact vs argmax of OUR fp32 upsample (same arithmetic); and vs torch up to fp32 round-off ties
    want = torch.argmax(y32, dim=1)
    assert torch.equal(mask.cpu().long(), want)
    assert (mask.cpu().long() == torch.argmax(ref, dim=1)).float().mean() > 0.9995
    from oracle.evaluator_oracle import compute_hist
    h = sum(compute_hist(want[i].numpy(), labels[i].numpy(), C, 255) for i in range(N))
    np.testing.assert_array_equal(hist.cpu().numpy(), 2 * h)  # accumulated twice (int64 + uint8 labels)
    # standalone confusion matrix incl. out-of-range predictions/labels (clipped like the reference)
    pred = torch.randint(-2, C + 3, (5000,), generator=torch.Generator().manual_seed(4))
    lab = torch.randint(0, C + 2, (5000,), generator=torch.Generator().manual_seed(5))
    lab[::7] = 255
    hist2 = torch.zeros(C, C, dtype=torch.int64, device="cuda")
    pd, ld_ = pred.cuda(), lab.cuda()
    check(lib.cabinet_confusion_hist(pd.data_ptr(), 0, ld_.data_ptr(), 0, 5000, C, 255, hist2.data_ptr(), stream()), "hist")
    torch.cuda.synchronize()
    np.testing.assert_array_equal(hist2.cpu().numpy(), compute_hist(pred.numpy(), lab.numpy(), C, 255))


def test_empty_batch_is_a_noop():
    lib = _lib.load()
    check(lib.cabinet_softmax_rows(1, 1, F32, 0, 5, stream()), "softmax")
    check(lib.cabinet_confusion_hist(1, 0, 1, 0, 0, 4, 255, 1, stream()), "hist")


TC_CASES = [
    # cin, cout, k, s, p, H, W, act, res, out_fp32
    (16, 64, 1, 1, 0, 17, 23, ACT_RELU, False, False),      # K = 16: single 16-wide MMA step, OOB-filled K tail
    (24, 72, 1, 1, 0, 9, 31, ACT_HSWISH, False, False),     # K, N not multiples of 16/64
    (72, 24, 1, 1, 0, 16, 16, ACT_NONE, True, False),       # linear bottleneck + residual
    (200, 80, 1, 1, 0, 5, 7, ACT_NONE, True, False),        # 4 K blocks with a ragged tail
    (160, 960, 1, 1, 0, 8, 8, ACT_HSWISH, False, False),    # N = 960 -> 4 cout tiles of 240
    (256, 19, 1, 1, 0, 6, 10, ACT_NONE, False, True),       # class head: N = 19, fp32 output
    (256, 8, 1, 1, 0, 16, 16, ACT_NONE, False, True),
    (64, 64, 3, 1, 1, 20, 36, ACT_RELU, False, False),      # 3x3 stride 1: 9 shifted box loads, OOB = zero padding
    (64, 64, 3, 2, 1, 33, 18, ACT_RELU, False, False),      # 3x3 stride 2, odd height: parity tensor maps
    (64, 64, 3, 2, 1, 64, 64, ACT_RELU, False, False),
    (128, 256, 3, 1, 1, 3, 4, ACT_RELU, False, False),      # map smaller than one tile
    (320, 256, 3, 1, 1, 32, 32, ACT_RELU, False, False),    # long K loop (45 k-blocks): ring wrap-around
    (640, 128, 1, 1, 0, 32, 32, ACT_NONE, False, False),
]


@pytest.mark.parametrize("cin,cout,k,s,p,H,W,act,res,out_fp32", TC_CASES)
def test_conv_tc(cin, cout, k, s, p, H, W, act, res, out_fp32):
    lib = _lib.load()
    dtype = torch.bfloat16
    N = 3
    x = q(gen(N, cin, H, W, seed=1), dtype)
    w = q(gen(cout, cin, k, k, seed=2, scale=(cin * k * k) ** -0.5), dtype)
    b = gen(cout, seed=3, scale=0.1)
    OH, OW = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    r = q(gen(N, cout, OH, OW, seed=4), dtype) if res else None
    ref = act_ref(F.conv2d(x, w, b, s, p), act)
    if res:
        ref = ref + r
    xm = to_map(x, dtype, ld=cin + 16, off=8)   # channel slice of a wider (concat) buffer
    odt = torch.float32 if out_fp32 else dtype
    ym = to_map(torch.zeros(N, cout, OH, OW), odt, ld=cout + 24, off=16 if not out_fp32 else 3)
    ym.t.fill_(7.0)
    rm = to_map(r, dtype) if res else None
    n16, c64 = -(-cout // 16) * 16, -(-cin // 64) * 64
    pk = torch.zeros(n16, k * k, c64)
    pk[:cout, :, :cin] = w.permute(0, 2, 3, 1).reshape(cout, k * k, cin)
    pk = pk.to("cuda", dtype).contiguous()
    bd = b.cuda()
    check(lib.cabinet_conv_tc(xm.ptr, xm.ld, N, H, W, cin, pk.data_ptr(), cout, k, k, s, p, bd.data_ptr(),
                              rm.ptr if res else None, rm.ld if res else 0, ym.ptr, ym.dt, ym.ld, OH, OW, act,
                              stream()), "conv_tc")
    torch.cuda.synchronize()
    got = from_map(ym)
    err = rel_l2(got, ref)
    print(f"conv_tc cin={cin} cout={cout} k={k} s={s} {H}x{W}: rel_l2 {err:.3e}")
    assert err < (2e-5 if out_fp32 else 6e-3)  # fp32 out: accumulation order only; bf16 out: one rounding
    # neighbouring channels of the wider output buffer are untouched
    full = ym.t.float()
    assert float((full[..., : ym.off] - 7.0).abs().max()) == 0
    assert float((full[..., ym.off + cout:] - 7.0).abs().max()) == 0


@pytest.mark.parametrize("N,H,W", [(2, 64, 64), (1, 37, 52), (3, 128, 96), (1, 21, 8)])
def test_stem_tc2(N, H, W):
    """Stems without the im2col tile (overlapping-row UMMA descriptor over the interleaved bf16 window) vs torch."""
    lib = _lib.load()
    x = gen(N, 3, H, W, seed=1)
    w7, b7 = gen(64, 3, 7, 7, seed=2, scale=0.1), gen(64, seed=3, scale=0.1)
    w3, b3 = gen(16, 3, 3, 3, seed=4, scale=0.2), gen(16, seed=5, scale=0.1)
    xq = q(x, torch.bfloat16)
    ref_sb = F.relu(F.conv2d(xq, q(w7, torch.bfloat16), b7, 2, 3))
    t = F.conv2d(xq, q(w3, torch.bfloat16), b3, 2, 1)
    ref_st = t * F.relu6(t + 3) / 6
    OH, OW = ref_sb.shape[2:]
    pk = torch.zeros(80, 3, 7, 8)
    pk[:64, :, :, 1:8] = w7
    pk[64:, :, 2:5, 3:6] = w3
    bias = torch.cat([b7, b3])
    wn = torch.zeros(80, 7, 8, 4)               # W[o][ky][kx][c]; c = 3 is the constant-1 channel
    wn[..., :3] = pk.permute(0, 2, 3, 1)
    hi = bias.to(torch.bfloat16).float()
    wn[:, 3, 4, 3], wn[:, 3, 5, 3] = hi, bias - hi
    wk = wn.view(80, 7, 2, 2, 8).permute(1, 2, 3, 0, 4).contiguous().to("cuda", torch.bfloat16)
    xd = x.cuda()
    y_sb = to_map(torch.zeros(N, 64, OH, OW), torch.bfloat16)
    y_st = to_map(torch.zeros(N, 16, OH, OW), torch.bfloat16)
    check(lib.cabinet_stem_tc2(xd.data_ptr(), N, H, W, wk.data_ptr(), y_sb.ptr, y_sb.ld, y_st.ptr, y_st.ld, OH, OW,
                               stream()), "stem_tc2")
    torch.cuda.synchronize()
    e1, e2 = rel_l2(from_map(y_sb), ref_sb), rel_l2(from_map(y_st), ref_st)
    print(f"stem_tc2 {N}x{H}x{W}: sb {e1:.3e} stem {e2:.3e}")
    assert e1 < 6e-3 and e2 < 6e-3


@pytest.mark.parametrize("N,H,W", [(2, 64, 64), (1, 37, 52), (3, 128, 96), (1, 21, 8)])
def test_stem_tc(N, H, W):
    """Fused 7x7/3x3 stems on tcgen05 vs the two fp32 torch convolutions."""
    lib = _lib.load()
    x = gen(N, 3, H, W, seed=1)
    w7, b7 = gen(64, 3, 7, 7, seed=2, scale=0.1), gen(64, seed=3, scale=0.1)
    w3, b3 = gen(16, 3, 3, 3, seed=4, scale=0.2), gen(16, seed=5, scale=0.1)
    xq = q(x, torch.bfloat16)
    ref_sb = F.relu(F.conv2d(xq, q(w7, torch.bfloat16), b7, 2, 3))
    t = F.conv2d(xq, q(w3, torch.bfloat16), b3, 2, 1)
    ref_st = t * F.relu6(t + 3) / 6
    OH, OW = ref_sb.shape[2:]
    pk = torch.zeros(80, 3, 7, 8)
    pk[:64, :, :, 1:8] = w7
    pk[64:, :, 2:5, 3:6] = w3
    wk = torch.zeros(80, 192)
    wk[:, :168] = pk.reshape(80, 168)
    bias = torch.cat([b7, b3])
    wk[:, 168] = bias.to(torch.bfloat16).float()
    wk[:, 169] = bias - wk[:, 168]
    wk, bias = wk.to("cuda", torch.bfloat16).contiguous(), bias.cuda()
    xd = x.cuda()
    y_sb = to_map(torch.zeros(N, 64, OH, OW), torch.bfloat16)
    y_st = to_map(torch.zeros(N, 16, OH, OW), torch.bfloat16)
    check(lib.cabinet_stem_tc(xd.data_ptr(), N, H, W, wk.data_ptr(), bias.data_ptr(), y_sb.ptr, y_sb.ld, y_st.ptr,
                              y_st.ld, OH, OW, stream()), "stem_tc")
    torch.cuda.synchronize()
    e1, e2 = rel_l2(from_map(y_sb), ref_sb), rel_l2(from_map(y_st), ref_st)
    print(f"stem_tc {N}x{H}x{W}: sb {e1:.3e} stem {e2:.3e}")
    assert e1 < 6e-3 and e2 < 6e-3


@pytest.mark.parametrize("C,k,s,H,W,act,gap", [
    (16, 3, 1, 13, 21, ACT_RELU, False), (64, 3, 2, 32, 32, ACT_RELU, False), (72, 5, 2, 19, 27, ACT_NONE, True),
    (120, 5, 1, 16, 9, ACT_NONE, True), (240, 3, 2, 7, 5, ACT_HSWISH, False), (960, 5, 1, 4, 4, ACT_NONE, True),
    (8, 3, 1, 1, 1, ACT_RELU, True), (200, 3, 1, 40, 33, ACT_HSWISH, True), (672, 5, 2, 33, 64, ACT_NONE, True),
])
def test_dwconv_tma(C, k, s, H, W, act, gap):
    lib = _lib.load()
    dtype = torch.bfloat16
    N = 2
    x = q(gen(N, C, H, W, seed=1), dtype)
    w = gen(C, 1, k, k, seed=2, scale=0.3)
    b = gen(C, seed=3, scale=0.1)
    p = (k - 1) // 2
    ref = act_ref(F.conv2d(x, w, b, s, p, 1, C), act)
    OH, OW = ref.shape[2:]
    xm, ym = to_map(x, dtype, ld=C + 8, off=8), to_map(torch.zeros_like(ref), dtype, ld=C + 16, off=8)
    wp, bd = w.view(C, -1).t().contiguous().cuda(), b.cuda()
    runs = []
    for _ in range(3 if gap else 1):
        g = torch.zeros((N, C), dtype=torch.int64, device="cuda") if gap else None  # fixed-point accumulators (2^-24)
        check(lib.cabinet_dwconv_tma(xm.ptr, xm.ld, wp.data_ptr(), bd.data_ptr(), ym.ptr, ym.ld, N, H, W, C, k, s, OH, OW,
                                     act, g.data_ptr() if gap else None, stream()), "dwconv_tma")
        torch.cuda.synchronize()
        runs.append(g)
    assert rel_l2(from_map(ym), ref) < tol(dtype)
    if gap:
        assert rel_l2(g.cpu().double().mul(2.0 ** -24).float(), ref.sum(dim=(2, 3))) < 1e-4
        assert all(torch.equal(r, runs[0]) for r in runs)  # deterministic: integer accumulation of fixed-order partials


@pytest.mark.parametrize("N,C,J,act,bias", [(16, 960, 240, ACT_RELU, True), (3, 72, 24, ACT_HSIGMOID, True),
                                             (19, 64, 256, ACT_SIGMOID, False), (1, 8, 8, ACT_NONE, True)])
def test_gate_fc(N, C, J, act, bias):
    lib = _lib.load()
    x, W = gen(N, C, seed=1), gen(J, C, seed=2, scale=C ** -0.5)
    b = gen(J, seed=3, scale=0.2) if bias else None
    ref = act_ref(F.linear(x * 0.25, W, b), act)
    xd, Wd, bd = x.cuda(), W.cuda(), (b.cuda() if bias else None)
    out = torch.empty(N, J, device="cuda")
    check(lib.cabinet_gate_fc(xd.data_ptr(), 0.25, Wd.data_ptr(), bd.data_ptr() if bias else None, out.data_ptr(), N, C,
                              J, act, 0, stream()), "gate_fc")
    torch.cuda.synchronize()
    assert rel_l2(out.cpu(), ref) < 1e-5
    # in_fixed: the input arrives as int64 fixed-point sums (2^-24), as the depthwise kernels accumulate them
    xf = (x.double() * 2.0 ** 24).round().to(torch.int64).cuda()
    check(lib.cabinet_gate_fc(xf.data_ptr(), 0.25, Wd.data_ptr(), bd.data_ptr() if bias else None, out.data_ptr(), N, C,
                              J, act, 1, stream()), "gate_fc")
    torch.cuda.synchronize()
    assert rel_l2(out.cpu(), ref) < 1e-5


@pytest.mark.parametrize("N,L", [(2, 1024), (3, 12), (1, 300), (2, 128), (1, 2040)])
def test_attention_tc(N, L):
    """Fused tcgen05 attention vs softmax(q k^T / sqrt(d)) v in fp32 (bf16-rounded inputs)."""
    lib = _lib.load()
    d = 128
    bf = torch.bfloat16
    qf, kf, vf = (q(gen(N, L, d, seed=i, scale=s), bf) for i, s in ((1, 1.0), (2, 1.5), (3, 1.0)))
    ref = torch.softmax(qf @ kf.transpose(1, 2) * d ** -0.5, dim=-1) @ vf
    qd, kd, vd = (t.to("cuda", bf).contiguous() for t in (qf, kf, vf))
    vt = torch.empty(N, d, (L + 7) // 8 * 8, dtype=bf, device="cuda")
    ctx = torch.zeros(N, L, d, dtype=bf, device="cuda")
    check(lib.cabinet_attention_tc(qd.data_ptr(), d, kd.data_ptr(), d, vd.data_ptr(), d, vt.data_ptr(), ctx.data_ptr(),
                                   d, N, L, d, d ** -0.5, stream()), "attention_tc")
    torch.cuda.synchronize()
    err = rel_l2(ctx.float().cpu(), ref)
    print(f"attention_tc N={N} L={L}: rel_l2 {err:.3e}")
    assert err < 8e-3  # P is rounded to bf16 before the P V product (like the reference under bf16 autocast)


@pytest.mark.parametrize("cin,cout,H,W,a_act,res", [(72, 40, 16, 16, ACT_RELU, False), (120, 40, 9, 13, ACT_RELU, True),
                                                      (672, 112, 8, 8, ACT_HSWISH, True), (960, 160, 5, 7, ACT_HSWISH, False),
                                                      (16, 16, 33, 18, ACT_NONE, False)])
def test_conv_tc_se_prologue(cin, cout, H, W, a_act, res):
    """Project 1x1 with the SE apply fused in front: y = W * act(x * s[n, c]) + b (+ res)."""
    lib = _lib.load()
    dtype = torch.bfloat16
    N = 3
    x = q(gen(N, cin, H, W, seed=1), dtype)
    sc = torch.rand(N, cin, generator=torch.Generator().manual_seed(5))
    w = q(gen(cout, cin, 1, 1, seed=2, scale=cin ** -0.5), dtype)
    b = gen(cout, seed=3, scale=0.1)
    r = q(gen(N, cout, H, W, seed=4), dtype) if res else None
    a = q(act_ref(x * sc.view(N, cin, 1, 1), a_act), dtype)  # the kernel re-rounds the transformed tile to bf16
    ref = F.conv2d(a, w, b) + (r if res else 0)
    xm = to_map(x, dtype)
    ym = to_map(torch.zeros(N, cout, H, W), dtype)
    rm = to_map(r, dtype) if res else None
    n16, c64 = -(-cout // 16) * 16, -(-cin // 64) * 64
    pk = torch.zeros(n16, 1, c64)
    pk[:cout, :, :cin] = w.reshape(cout, 1, cin)
    pk, bd, scd = pk.to("cuda", dtype).contiguous(), b.cuda(), sc.cuda().contiguous()
    check(lib.cabinet_conv_tc_se(xm.ptr, xm.ld, N, H, W, cin, scd.data_ptr(), a_act, pk.data_ptr(), cout, 1, 1, 1, 0,
                                 bd.data_ptr(), rm.ptr if res else None, rm.ld if res else 0, ym.ptr, ym.dt, ym.ld, H, W,
                                 ACT_NONE, stream()), "conv_tc_se")
    torch.cuda.synchronize()
    err = rel_l2(from_map(ym), ref)
    print(f"conv_tc_se cin={cin} cout={cout}: rel_l2 {err:.3e}")
    assert err < 6e-3


def test_normalize_u8_matches_totensor_normalize():
    lib = _lib.load()
    N, H, W = 2, 18, 22
    x = torch.randint(0, 256, (N, H, W, 3), generator=torch.Generator().manual_seed(1), dtype=torch.uint8)
    mean, std = (0.480, 0.499, 0.457), (0.225, 0.208, 0.228)
    ref = (x.permute(0, 3, 1, 2).float() / 255 - torch.tensor(mean).view(1, 3, 1, 1)) / torch.tensor(std).view(1, 3, 1, 1)
    xd = x.cuda()
    y = torch.empty(N, 3, H, W, device="cuda")
    check(lib.cabinet_normalize_u8(xd.data_ptr(), y.data_ptr(), N, H, W, *mean, *std, stream()), "normalize_u8")
    torch.cuda.synchronize()
    assert float((y.cpu() - ref).abs().max()) < 2e-6


@pytest.mark.parametrize("C,H,W,act", [(16, 32, 48, ACT_RELU), (16, 13, 21, ACT_RELU), (32, 9, 7, ACT_HSWISH), (8, 40, 16, ACT_RELU)])
def test_mbconv_noexpand_fused(C, H, W, act):
    """y = x + W2 * act(dw3x3(x) + b1) + b2 in one kernel vs the three torch ops."""
    lib = _lib.load()
    dtype = torch.bfloat16
    N = 2
    x = q(gen(N, C, H, W, seed=1), dtype)
    wd, bd = gen(C, 1, 3, 3, seed=2, scale=0.3), gen(C, seed=3, scale=0.1)
    wp, bp = gen(C, C, 1, 1, seed=4, scale=C ** -0.5), gen(C, seed=5, scale=0.1)
    ref = x + F.conv2d(act_ref(F.conv2d(x, wd, bd, 1, 1, 1, C), act), wp, bp)
    xm, ym = to_map(x, dtype, ld=C + 8, off=8), to_map(torch.zeros_like(ref), dtype)
    wdd, bdd = wd.view(C, -1).t().contiguous().cuda(), bd.cuda()
    wpd, bpd = wp.view(C, C).contiguous().cuda(), bp.cuda()
    check(lib.cabinet_mbconv_noexpand_fused(xm.ptr, xm.ld, wdd.data_ptr(), bdd.data_ptr(), wpd.data_ptr(), bpd.data_ptr(),
                                            ym.ptr, ym.ld, N, H, W, C, act, stream()), "mbconv1")
    torch.cuda.synchronize()
    assert rel_l2(from_map(ym), ref) < tol(dtype)


def _pack_aux(wd, be, bd):
    """[chunks][k*k + 2][64] fp32: depthwise taps, expand bias, depthwise bias (include/cabinet_b200.h)."""
    cexp, kk = wd.shape[0], wd.shape[2] * wd.shape[3]
    nc = -(-cexp // 64)
    aux = torch.zeros(nc * 64, kk + 2)
    aux[:cexp, :kk] = wd.reshape(cexp, kk)
    aux[:cexp, kk], aux[:cexp, kk + 1] = be, bd
    return aux.view(nc, 64, kk + 2).permute(0, 2, 1).contiguous().cuda()


def _pack_expand(we, be):
    """[ceil16(Cexp)][64 * (Cin // 64 + 1)] bf16: expand weights + the bias as two extra K columns (hi, lo) --
    include/cabinet_b200.h."""
    cexp, cin = we.shape[:2]
    pk = torch.zeros(-(-cexp // 16) * 16, 64 * (cin // 64 + 1))
    pk[:cexp, :cin] = we.reshape(cexp, cin)
    hi = be.to(torch.bfloat16).float()
    pk[:cexp, cin], pk[:cexp, cin + 1] = hi, be - hi
    return pk.to("cuda", torch.bfloat16).contiguous()


def _pack_tc(w, dtype=torch.bfloat16):
    cout, cin = w.shape[:2]
    n16, c64 = -(-cout // 16) * 16, -(-cin // 64) * 64
    pk = torch.zeros(n16, 1, c64)
    pk[:cout, 0, :cin] = w.reshape(cout, cin)
    return pk.to("cuda", dtype).contiguous()


@pytest.mark.parametrize("cin,cexp,cout,k,s,H,W,act,res", [
    (16, 64, 24, 3, 2, 40, 56, ACT_RELU, False),      # Large f2
    (24, 72, 24, 3, 1, 24, 40, ACT_RELU, True),       # Large f3 (ragged second chunk of 8 channels)
    (40, 240, 80, 3, 2, 20, 36, ACT_HSWISH, False),   # Large f7 (4 chunks, streamed weights, 2 store groups)
    (40, 120, 40, 5, 1, 19, 23, ACT_RELU, True),      # k5 with project
    (16, 16, 16, 3, 1, 9, 7, ACT_RELU, True),         # single narrow chunk, tile larger than the image
    (24, 88, 24, 5, 2, 33, 17, ACT_HSWISH, False),    # Small f3-like k5 stride 2, odd sizes
    (56, 128, 56, 3, 1, 16, 16, ACT_HSWISH, True),
    (16, 72, 24, 3, 2, 30, 34, ACT_RELU, False),      # Small f2: two chunks at stride 2 (narrow-tile fallback)
    (40, 240, 40, 5, 1, 14, 18, ACT_HSWISH, True),    # Small f5-like: four k5 chunks, streamed taps
    (80, 200, 80, 3, 1, 20, 28, ACT_HSWISH, True),    # Large f8: two K blocks of input channels, 8 x 8 tiles
    (80, 184, 80, 3, 1, 64, 64, ACT_HSWISH, True),    # Large f9 / f10 at their config-2 size
    (72, 136, 72, 3, 1, 13, 9, ACT_RELU, True),       # odd size, ragged last chunk with two K blocks
    (64, 128, 64, 3, 1, 16, 16, ACT_RELU, True),      # Cin a multiple of 64: the bias slots open a K block of their own
])
def test_mbconv_fused_project(cin, cexp, cout, k, s, H, W, act, res):
    """expand -> depthwise -> project (+identity) in one kernel vs the torch ops with the same bf16 roundings."""
    lib = _lib.load()
    dtype = torch.bfloat16
    N = 3
    x = q(gen(N, cin, H, W, seed=1), dtype)
    we, be = q(gen(cexp, cin, 1, 1, seed=2, scale=cin ** -0.5), dtype), gen(cexp, seed=3, scale=0.2)
    wd, bd = gen(cexp, 1, k, k, seed=4, scale=1.0 / k), gen(cexp, seed=5, scale=0.1)
    wp, bp = q(gen(cout, cexp, 1, 1, seed=6, scale=cexp ** -0.5), dtype), gen(cout, seed=7, scale=0.1)
    pad = (k - 1) // 2
    h = q(act_ref(F.conv2d(x, we, be), act), dtype)
    d = q(act_ref(F.conv2d(h, wd, bd, s, pad, 1, cexp), act), dtype)
    ref = F.conv2d(d, wp, bp)
    if res:
        ref = ref + x
    OH, OW = ref.shape[2:]
    xm = to_map(x, dtype, ld=cin + 8, off=8)
    ym = to_map(torch.zeros_like(ref), dtype, ld=cout + 24, off=16)
    ym.t.fill_(7.0)
    aux, bpd = _pack_aux(wd, be, bd), bp.cuda()
    pe, pp = _pack_expand(we, be), _pack_tc(wp)
    check(lib.cabinet_mbconv_fused(xm.ptr, xm.ld, N, H, W, cin, pe.data_ptr(), aux.data_ptr(), cexp, act, k, s, act,
                                   pp.data_ptr(), bpd.data_ptr(), cout, 1 if res else 0, ym.ptr, ym.ld, OH, OW, None,
                                   stream()), "mbconv_fused")
    torch.cuda.synchronize()
    err = rel_l2(from_map(ym), ref)
    print(f"mbconv_fused {cin}->{cexp}->{cout} k{k} s{s} {H}x{W}: rel_l2 {err:.3e}")
    assert err < 8e-3  # three bf16 roundings (h, d, y); a flipped rounding of h or d moves y by ~1 bf16 ulp
    full = ym.t.float()
    assert float((full[..., : ym.off] - 7.0).abs().max()) == 0
    assert float((full[..., ym.off + cout:] - 7.0).abs().max()) == 0


@pytest.mark.parametrize("act_dw", [ACT_NONE, ACT_RELU])
@pytest.mark.parametrize("cin,cexp,k,s,H,W,act", [
    (24, 72, 5, 2, 40, 56, ACT_RELU),       # Large f4
    (40, 120, 5, 1, 24, 40, ACT_RELU),      # Large f5 / f6
    (40, 240, 3, 1, 11, 13, ACT_HSWISH),
    (16, 72, 3, 2, 17, 9, ACT_RELU),
    (80, 480, 3, 1, 64, 64, ACT_HSWISH),    # Large f11
    (112, 672, 3, 1, 21, 35, ACT_HSWISH),   # Large f12, odd size
    (48, 144, 5, 1, 17, 23, ACT_HSWISH),    # Small f7-like
    (96, 128, 5, 1, 12, 20, ACT_HSWISH),    # k5 with two K blocks
    (160, 64, 3, 1, 16, 24, ACT_HSWISH),    # three K blocks
])
def test_mbconv_fused_dw_out(cin, cexp, k, s, H, W, act, act_dw):
    """expand -> depthwise with the pre-SE output and its pooling sums (blocks with squeeze-excite)."""
    lib = _lib.load()
    dtype = torch.bfloat16
    N = 2
    x = q(gen(N, cin, H, W, seed=1), dtype)
    we, be = q(gen(cexp, cin, 1, 1, seed=2, scale=cin ** -0.5), dtype), gen(cexp, seed=3, scale=0.2)
    wd, bd = gen(cexp, 1, k, k, seed=4, scale=1.0 / k), gen(cexp, seed=5, scale=0.1)
    pad = (k - 1) // 2
    h = q(act_ref(F.conv2d(x, we, be), act), dtype)
    pre = F.conv2d(h, wd, bd, s, pad, 1, cexp)   # the pooling sums are taken BEFORE act_dw (SE pools the BN output)
    ref = act_ref(pre, act_dw)
    OH, OW = ref.shape[2:]
    xm = to_map(x, dtype)
    ym = to_map(torch.zeros_like(ref), dtype, ld=cexp + 8, off=0)
    ym.t.fill_(7.0)
    aux = _pack_aux(wd, be, bd)
    pe = _pack_expand(we, be)
    sums = []
    for _ in range(2):
        acc = torch.zeros((N, cexp), dtype=torch.int64, device="cuda")  # fixed-point accumulators (2^-24)
        check(lib.cabinet_mbconv_fused(xm.ptr, xm.ld, N, H, W, cin, pe.data_ptr(), aux.data_ptr(), cexp, act, k, s,
                                       act_dw, None, None, 0, 0, ym.ptr, ym.ld, OH, OW, acc.data_ptr(), stream()),
              "mbconv_fused")
        torch.cuda.synchronize()
        sums.append(acc)
    assert torch.equal(sums[0], sums[1])   # deterministic
    gap = sums[0].double().mul(2.0 ** -24).float()
    err = rel_l2(from_map(ym), ref)
    gerr = rel_l2(gap.cpu(), pre.sum(dim=(2, 3)))
    print(f"mbconv_fused(dw out) {cin}->{cexp} k{k} s{s} {H}x{W}: rel_l2 {err:.3e} gap {gerr:.3e}")
    assert err < 6e-3 and gerr < 1e-3
    assert float((ym.t.float()[..., cexp:] - 7.0).abs().max()) == 0


def _pack_expand_t(we, be):
    """[nc * 128][64 * (Cin // 64 + 1)] bf16, row (c, l) = channel c * CH + l % CH (include/cabinet_b200.h: cabinet_mbconv_t)."""
    cexp, cin = we.shape[:2]
    ch = 64 if cexp <= 64 else 128
    nc = 1 if cexp <= 64 else -(-cexp // 128)
    flat = torch.zeros(nc * ch, 64 * (cin // 64 + 1))
    flat[:cexp, :cin] = we.reshape(cexp, cin)
    hi = be.to(torch.bfloat16).float()
    flat[:cexp, cin], flat[:cexp, cin + 1] = hi, be - hi
    pk = flat.view(nc, ch, -1).repeat(1, 128 // ch, 1).reshape(nc * 128, -1)
    return pk.to("cuda", torch.bfloat16).contiguous()


def _pack_aux_t(wd, bd):
    """[nc][k*k + 1][128] fp32: depthwise taps + bias, rows mapped like _pack_expand_t."""
    cexp, kk = wd.shape[0], wd.shape[2] * wd.shape[3]
    ch = 64 if cexp <= 64 else 128
    nc = 1 if cexp <= 64 else -(-cexp // 128)
    flat = torch.zeros(nc * ch, kk + 1)
    flat[:cexp, :kk] = wd.reshape(cexp, kk)
    flat[:cexp, kk] = bd
    return flat.view(nc, ch, kk + 1).repeat(1, 128 // ch, 1).permute(0, 2, 1).contiguous().cuda()


@pytest.mark.parametrize("cin,cexp,cout,k,s,H,W,act,res", [
    (16, 64, 24, 3, 2, 40, 56, ACT_RELU, False),      # Large f2 (64 channels replicated over the 128 lanes)
    (24, 72, 24, 3, 1, 24, 40, ACT_RELU, True),       # Large f3
    (40, 240, 80, 3, 2, 20, 36, ACT_HSWISH, False),   # Large f7 (two chunks)
    (16, 16, 16, 3, 1, 9, 7, ACT_RELU, True),         # tile larger than the image
    (56, 128, 56, 3, 1, 16, 16, ACT_HSWISH, True),
    (16, 72, 24, 3, 2, 30, 34, ACT_RELU, False),      # Small f2
    (80, 200, 80, 3, 1, 20, 28, ACT_HSWISH, True),    # Large f8: two K blocks
    (80, 184, 80, 3, 1, 64, 64, ACT_HSWISH, True),    # Large f9 / f10 at their config-2 size
    (72, 392, 72, 3, 1, 13, 9, ACT_RELU, True),       # four chunks: streamed weights, odd size
    (64, 128, 64, 3, 1, 16, 16, ACT_RELU, True),      # Cin a multiple of 64: the bias slots open a K block of their own
    (24, 48, 128, 3, 1, 33, 17, ACT_HSWISH, False),   # 128 output channels
])
def test_mbconv_t_project(cin, cexp, cout, k, s, H, W, act, res):
    """Channel-major fused block (expand -> depthwise from TMEM -> project) vs torch ops; the expanded activation stays
    fp32 here, so only d and y carry bf16 roundings."""
    lib = _lib.load()
    dtype = torch.bfloat16
    N = 3
    x = q(gen(N, cin, H, W, seed=1), dtype)
    we, be = q(gen(cexp, cin, 1, 1, seed=2, scale=cin ** -0.5), dtype), gen(cexp, seed=3, scale=0.2)
    wd, bd = gen(cexp, 1, k, k, seed=4, scale=1.0 / k), gen(cexp, seed=5, scale=0.1)
    wp, bp = q(gen(cout, cexp, 1, 1, seed=6, scale=cexp ** -0.5), dtype), gen(cout, seed=7, scale=0.1)
    pad = (k - 1) // 2
    h = act_ref(F.conv2d(x, we, be), act)
    d = q(act_ref(F.conv2d(h, wd, bd, s, pad, 1, cexp), act), dtype)
    ref = F.conv2d(d, wp, bp)
    if res:
        ref = ref + x
    OH, OW = ref.shape[2:]
    xm = to_map(x, dtype, ld=cin + 8, off=8)
    ym = to_map(torch.zeros_like(ref), dtype, ld=cout + 24, off=16)
    ym.t.fill_(7.0)
    aux, bpd = _pack_aux_t(wd, bd), bp.cuda()
    pe, pp = _pack_expand_t(we, be), _pack_tc(wp)
    check(lib.cabinet_mbconv_t(xm.ptr, xm.ld, N, H, W, cin, pe.data_ptr(), aux.data_ptr(), cexp, act, k, s, act,
                               pp.data_ptr(), bpd.data_ptr(), cout, 1 if res else 0, ym.ptr, ym.ld, OH, OW, None,
                               None, stream()), "mbconv_t")
    torch.cuda.synchronize()
    err = rel_l2(from_map(ym), ref)
    print(f"mbconv_t {cin}->{cexp}->{cout} k{k} s{s} {H}x{W}: rel_l2 {err:.3e}")
    assert err < 6e-3
    full = ym.t.float()
    assert float((full[..., : ym.off] - 7.0).abs().max()) == 0
    assert float((full[..., ym.off + cout:] - 7.0).abs().max()) == 0


@pytest.mark.parametrize("act_dw", [ACT_NONE, ACT_RELU])
@pytest.mark.parametrize("cin,cexp,k,s,H,W,act", [
    (40, 120, 5, 1, 24, 40, ACT_RELU),      # Large f5 / f6
    (40, 240, 3, 1, 11, 13, ACT_HSWISH),
    (16, 72, 3, 2, 17, 9, ACT_RELU),
    (80, 480, 3, 1, 64, 64, ACT_HSWISH),    # Large f11
    (112, 672, 3, 1, 21, 35, ACT_HSWISH),   # Large f12, odd size
    (160, 960, 5, 1, 32, 32, ACT_HSWISH),   # Large f14 / f15: three K blocks, eight chunks
    (48, 144, 5, 1, 17, 23, ACT_HSWISH),    # Small f7-like
    (16, 64, 5, 1, 12, 20, ACT_RELU),       # k5 with replicated lanes
    (16, 48, 3, 2, 22, 14, ACT_RELU),       # stride 2 with replicated lanes, 48 of 64 channels
    (24, 72, 5, 2, 40, 56, ACT_RELU),       # Large f4 (k5 stride 2)
    (112, 672, 5, 2, 23, 31, ACT_HSWISH),   # Large f13, odd size
    (16, 64, 5, 2, 18, 10, ACT_HSWISH),     # k5 stride 2 with replicated lanes
])
def test_mbconv_t_dw_out(cin, cexp, k, s, H, W, act, act_dw):
    """Channel-major expand -> depthwise with the pre-SE output and its pooling sums (squeeze-excite blocks)."""
    lib = _lib.load()
    dtype = torch.bfloat16
    N = 2
    x = q(gen(N, cin, H, W, seed=1), dtype)
    we, be = q(gen(cexp, cin, 1, 1, seed=2, scale=cin ** -0.5), dtype), gen(cexp, seed=3, scale=0.2)
    wd, bd = gen(cexp, 1, k, k, seed=4, scale=1.0 / k), gen(cexp, seed=5, scale=0.1)
    pad = (k - 1) // 2
    h = act_ref(F.conv2d(x, we, be), act)
    pre = F.conv2d(h, wd, bd, s, pad, 1, cexp)
    ref = act_ref(pre, act_dw)
    OH, OW = ref.shape[2:]
    xm = to_map(x, dtype)
    ym = to_map(torch.zeros_like(ref), dtype, ld=cexp + 8, off=0)
    ym.t.fill_(7.0)
    aux, pe = _pack_aux_t(wd, bd), _pack_expand_t(we, be)
    sums = []
    for _ in range(2):
        acc = torch.zeros((N, cexp), dtype=torch.int64, device="cuda")
        check(lib.cabinet_mbconv_t(xm.ptr, xm.ld, N, H, W, cin, pe.data_ptr(), aux.data_ptr(), cexp, act, k, s,
                                   act_dw, None, None, 0, 0, ym.ptr, ym.ld, OH, OW, acc.data_ptr(), None, stream()),
              "mbconv_t")
        torch.cuda.synchronize()
        sums.append(acc)
    assert torch.equal(sums[0], sums[1])   # deterministic
    gap = sums[0].double().mul(2.0 ** -24).float()
    err = rel_l2(from_map(ym), ref)
    gerr = rel_l2(gap.cpu(), pre.sum(dim=(2, 3)))
    print(f"mbconv_t(dw out) {cin}->{cexp} k{k} s{s} {H}x{W}: rel_l2 {err:.3e} gap {gerr:.3e}")
    assert err < 4e-3 and gerr < 1e-3
    assert float((ym.t.float()[..., cexp:] - 7.0).abs().max()) == 0


@pytest.mark.parametrize("cin,cexp,k,H,W,act,nsplit", [
    (40, 120, 5, 128, 128, ACT_RELU, 8),     # Large f5 / f6 at their config-2 size
    (80, 480, 3, 64, 64, ACT_HSWISH, 2),     # Large f11
    (112, 672, 3, 21, 35, ACT_HSWISH, 3),    # Large f12, odd size (ragged last flat box, ragged 16-column pieces)
    (160, 960, 5, 32, 32, ACT_HSWISH, 1),    # Large f14 / f15: three K blocks
    (24, 72, 5, 9, 4, ACT_RELU, 1),          # image as narrow as the border (W = 2p)
    (48, 144, 3, 200, 13, ACT_HSWISH, 5),    # tall image: one row per flat box would be too few pixels -> 19 rows per box
])
def test_expand_sums(cin, cexp, k, H, W, act, nsplit):
    """Pooling sums of the depthwise BN output from border-corrected sums of the expanded activation (no depthwise conv
    is run) vs the sum over the real depthwise output (mobilenetv3.py:68-83,137-143); deterministic."""
    lib = _lib.load()
    dtype = torch.bfloat16
    N = 3
    x = q(gen(N, cin, H, W, seed=1), dtype)
    we, be = q(gen(cexp, cin, 1, 1, seed=2, scale=cin ** -0.5), dtype), gen(cexp, seed=3, scale=0.2)
    wd, bd = gen(cexp, 1, k, k, seed=4, scale=1.0 / k), gen(cexp, seed=5, scale=0.1)
    pad = (k - 1) // 2
    h = act_ref(F.conv2d(x, we, be), act)
    ref = F.conv2d(h, wd, bd, 1, pad, 1, cexp).sum(dim=(2, 3))
    xm = to_map(x, dtype, ld=cin + 8, off=8)
    pe, aux = _pack_expand_t(we, be), _pack_aux_t(wd, bd)
    sums = []
    for _ in range(2):
        acc = torch.zeros((N, cexp), dtype=torch.int64, device="cuda")
        check(lib.cabinet_expand_sums(xm.ptr, xm.ld, N, H, W, cin, pe.data_ptr(), aux.data_ptr(), cexp, act, k, nsplit,
                                      acc.data_ptr(), stream()), "expand_sums")
        torch.cuda.synchronize()
        sums.append(acc)
    assert torch.equal(sums[0], sums[1])
    gap = sums[0].double().mul(2.0 ** -24).float().cpu()
    err = float((gap - ref).abs().max() / ref.abs().max())
    print(f"expand_sums {cin}->{cexp} k{k} {H}x{W}: max err / max |sum| {err:.3e}")
    assert err < 2e-5


@pytest.mark.parametrize("cin,cexp,cout,k,H,W,act,res", [
    (40, 120, 40, 5, 24, 40, ACT_RELU, True),        # Large f5 / f6 (k5 project mode: 8 x 8 tiles)
    (80, 480, 112, 3, 20, 28, ACT_HSWISH, False),    # Large f11
    (112, 672, 112, 3, 21, 35, ACT_HSWISH, True),    # Large f12
    (16, 64, 24, 5, 13, 9, ACT_HSWISH, False),       # k5 with replicated lanes
])
def test_mbconv_t_se_block(cin, cexp, cout, k, H, W, act, res):
    """Whole squeeze-excite block in one launch once the gate is known: project(act(gate * (dw(h) + b))) (+ x)."""
    lib = _lib.load()
    dtype = torch.bfloat16
    N = 2
    x = q(gen(N, cin, H, W, seed=1), dtype)
    we, be = q(gen(cexp, cin, 1, 1, seed=2, scale=cin ** -0.5), dtype), gen(cexp, seed=3, scale=0.2)
    wd, bd = gen(cexp, 1, k, k, seed=4, scale=1.0 / k), gen(cexp, seed=5, scale=0.1)
    wp, bp = q(gen(cout, cexp, 1, 1, seed=6, scale=cexp ** -0.5), dtype), gen(cout, seed=7, scale=0.1)
    gate = torch.rand((N, cexp), generator=torch.Generator().manual_seed(8))
    pad = (k - 1) // 2
    h = act_ref(F.conv2d(x, we, be), act)
    d = q(act_ref(F.conv2d(h, wd, bd, 1, pad, 1, cexp) * gate[:, :, None, None], act), dtype)
    ref = F.conv2d(d, wp, bp)
    if res:
        ref = ref + x
    xm = to_map(x, dtype)
    ym = to_map(torch.zeros_like(ref), dtype, ld=cout + 8, off=0)
    ym.t.fill_(7.0)
    aux, bpd, gd = _pack_aux_t(wd, bd), bp.cuda(), gate.cuda()
    pe, pp = _pack_expand_t(we, be), _pack_tc(wp)
    check(lib.cabinet_mbconv_t(xm.ptr, xm.ld, N, H, W, cin, pe.data_ptr(), aux.data_ptr(), cexp, act, k, 1, act,
                               pp.data_ptr(), bpd.data_ptr(), cout, 1 if res else 0, ym.ptr, ym.ld, H, W, None,
                               gd.data_ptr(), stream()), "mbconv_t")
    torch.cuda.synchronize()
    err = rel_l2(from_map(ym), ref)
    print(f"mbconv_t SE block {cin}->{cexp}->{cout} k{k} {H}x{W}: rel_l2 {err:.3e}")
    assert err < 6e-3
    assert float((ym.t.float()[..., cout:] - 7.0).abs().max()) == 0


@pytest.mark.parametrize("C,J,gate,bias,fixed,plus,taps", [(120, 32, ACT_HSIGMOID, True, True, False, 1),
                                                          (256, 64, ACT_SIGMOID, False, False, True, 9),
                                                          (72, 24, ACT_HSIGMOID, True, True, False, 1)])
def test_gate_scale_weights(C, J, gate, bias, fixed, plus, taps):
    """gate MLP + per-image weight scaling in one launch == gate_fc x 2 + scale_weights."""
    lib = _lib.load()
    N, HW, rows = 3, 77, 48
    cin_pad = -(-C // 64) * 64
    sums = gen(N, C, seed=1) * HW
    w1, w2 = gen(J, C, seed=2, scale=C ** -0.5), gen(C, J, seed=3, scale=J ** -0.5)
    b1, b2 = (gen(J, seed=4, scale=0.1), gen(C, seed=5, scale=0.5)) if bias else (None, None)
    g = act_ref(F.linear(F.relu(F.linear(sums / HW, w1, b1)), w2, b2), gate) + (1.0 if plus else 0.0)
    w = torch.zeros(rows, taps, cin_pad)
    w[:, :, :C] = gen(rows, taps, C, seed=6)
    wq = w.to(torch.bfloat16)
    ref = (wq.float()[None] * F.pad(g, (0, cin_pad - C))[:, None, None, :]).to(torch.bfloat16)
    d = lambda t: None if t is None else t.cuda()  # noqa: E731
    gap = (sums.double() * 2.0 ** 24).round().to(torch.int64).cuda() if fixed else sums.cuda()
    w1d, w2d, b1d, b2d, wd = d(w1), d(w2), d(b1), d(b2), wq.cuda()
    out = torch.full((N, rows, taps, cin_pad), 3.0, dtype=torch.bfloat16, device="cuda")
    check(lib.cabinet_gate_scale_weights(gap.data_ptr(), int(fixed), 1.0 / HW, w1d.data_ptr(), b1d.data_ptr() if bias else None,
                                         w2d.data_ptr(), b2d.data_ptr() if bias else None, gate, C, J, wd.data_ptr(),
                                         out.data_ptr(), N, rows, taps, cin_pad, int(plus), stream()), "gate_scale_weights")
    torch.cuda.synchronize()
    assert rel_l2(out.float().cpu(), ref.float()) < 3e-3
    assert C == cin_pad or float(out[..., C:].abs().max()) == 0


def test_conv_tc_per_image_weights():
    """conv(x * (1 + a_n), W) == conv(x, W * (1 + a_n)): cabinet_scale_weights + cabinet_conv_tc_imgw vs torch."""
    lib = _lib.load()
    dtype = torch.bfloat16
    N, cin, cout, k, H, W = 3, 40, 24, 3, 19, 21
    x = q(gen(N, cin, H, W, seed=1), dtype)
    w = q(gen(cout, cin, k, k, seed=2, scale=(cin * k * k) ** -0.5), dtype)
    b = gen(cout, seed=3, scale=0.1)
    a = torch.rand(N, cin, generator=torch.Generator().manual_seed(4))
    ref = torch.relu(F.conv2d(x * (1 + a)[:, :, None, None], w, b, 1, 1))
    n16, c64 = -(-cout // 16) * 16, -(-cin // 64) * 64
    pk = torch.zeros(n16, k * k, c64)
    pk[:cout, :, :cin] = w.permute(0, 2, 3, 1).reshape(cout, k * k, cin)
    pk = pk.to("cuda", dtype).contiguous()
    wimg = torch.full((N, n16, k * k, c64), 7.0, dtype=dtype, device="cuda")
    ad, bd = a.cuda(), b.cuda()
    check(lib.cabinet_scale_weights(pk.data_ptr(), ad.data_ptr(), wimg.data_ptr(), N, n16, k * k, c64, cin, 1, stream()), "sw")
    torch.cuda.synchronize()
    exp = pk.float()[None] * torch.cat([1 + ad, torch.zeros(N, c64 - cin, device="cuda")], 1)[:, None, None, :]
    assert float((wimg.float() - exp.to(dtype).float()).abs().max()) == 0
    xm, ym = to_map(x, dtype), to_map(torch.zeros_like(ref), dtype)
    check(lib.cabinet_conv_tc_imgw(xm.ptr, xm.ld, N, H, W, cin, wimg.data_ptr(), n16 * k * k * c64, cout, k, k, 1, 1,
                                   bd.data_ptr(), None, 0, ym.ptr, ym.dt, ym.ld, H, W, ACT_RELU, stream()), "imgw")
    torch.cuda.synchronize()
    assert rel_l2(from_map(ym), ref) < 8e-3


def test_conv_tc_per_image_weights_1x1_with_residual():
    """Flat (1x1) tiles: every 128-pixel tile belongs to one image when H * W % 128 == 0 (SE gate folded into the
    project weights, relu(s * d) = s * relu(d))."""
    lib = _lib.load()
    dtype = torch.bfloat16
    N, cin, cout, H, W = 4, 120, 40, 16, 24   # H * W = 384 = 3 tiles per image
    d = q(torch.relu(gen(N, cin, H, W, seed=1)), dtype)
    w = q(gen(cout, cin, 1, 1, seed=2, scale=cin ** -0.5), dtype)
    b = gen(cout, seed=3, scale=0.1)
    sgate = torch.rand(N, cin, generator=torch.Generator().manual_seed(4))
    r = q(gen(N, cout, H, W, seed=5), dtype)
    ref = F.conv2d(d * sgate[:, :, None, None], w, b) + r
    n16, c64 = -(-cout // 16) * 16, -(-cin // 64) * 64
    pk = torch.zeros(n16, 1, c64)
    pk[:cout, 0, :cin] = w.reshape(cout, cin)
    pk = pk.to("cuda", dtype).contiguous()
    wimg = torch.empty((N, n16, 1, c64), dtype=dtype, device="cuda")
    sd, bd = sgate.cuda(), b.cuda()
    check(lib.cabinet_scale_weights(pk.data_ptr(), sd.data_ptr(), wimg.data_ptr(), N, n16, 1, c64, cin, 0, stream()), "sw")
    xm, rm, ym = to_map(d, dtype), to_map(r, dtype), to_map(torch.zeros_like(ref), dtype)
    check(lib.cabinet_conv_tc_imgw(xm.ptr, xm.ld, N, H, W, cin, wimg.data_ptr(), n16 * c64, cout, 1, 1, 1, 0,
                                   bd.data_ptr(), rm.ptr, rm.ld, ym.ptr, ym.dt, ym.ld, H, W, ACT_NONE, stream()), "imgw")
    torch.cuda.synchronize()
    assert rel_l2(from_map(ym), ref) < 8e-3
    # a map whose pixel count is not a multiple of 128 must be refused (a tile would straddle two images)
    rc = lib.cabinet_conv_tc_imgw(xm.ptr, xm.ld, N, 5, 7, cin, wimg.data_ptr(), n16 * c64, cout, 1, 1, 1, 0, bd.data_ptr(),
                                  None, 0, ym.ptr, ym.dt, ym.ld, 5, 7, ACT_NONE, stream())
    assert rc != 0


def test_conv_tc_split_act():
    """One GEMM for three projections of the same input: ReLU on the first 2 x 32 output channels only."""
    lib = _lib.load()
    dtype = torch.bfloat16
    N, cin, cout, H, W = 2, 48, 96, 12, 20
    x = q(gen(N, cin, H, W, seed=1), dtype)
    w = q(gen(cout, cin, 1, 1, seed=2, scale=cin ** -0.5), dtype)
    b = gen(cout, seed=3, scale=0.3)
    full = F.conv2d(x, w, b)
    ref = torch.cat([torch.relu(full[:, :64]), full[:, 64:]], 1)
    pk = torch.zeros(cout, 1, 64)
    pk[:, 0, :cin] = w.reshape(cout, cin)
    pk = pk.to("cuda", dtype).contiguous()
    bd = b.cuda()
    xm, ym = to_map(x, dtype), to_map(torch.zeros_like(ref), dtype)
    check(lib.cabinet_conv_tc_split_act(xm.ptr, xm.ld, N, H, W, cin, pk.data_ptr(), cout, 1, 1, 1, 0, bd.data_ptr(), ym.ptr,
                                        ym.dt, ym.ld, H, W, ACT_RELU, 64, stream()), "split_act")
    torch.cuda.synchronize()
    got = from_map(ym)
    assert rel_l2(got, ref) < 6e-3
    assert float(got[:, :64].min()) >= 0 and float(got[:, 64:].min()) < 0


# ------------------------------------------------------------------ evaluator tail kernels (MscEvalV0 general mode)
@pytest.mark.parametrize("N,IH,IW,C,OH,OW,flip,dy0,dx0,dh,dw,weights", [
    (2, 8, 8, 8, 64, 64, False, 0, 0, 64, 64, False),      # exact x8, whole window
    (2, 8, 16, 8, 64, 128, True, 5, 3, 80, 140, True),     # exact x8 + flip TTA + unaligned window + overlap weights
    (1, 8, 8, 19, 64, 64, True, -7, -9, 50, 48, True),     # C % 4 != 0, padded chip: destination clipped on all sides
    (2, 5, 7, 4, 37, 53, False, 2, 1, 40, 60, True),       # generic bilinear, OW % 8 != 0
    (1, 5, 7, 5, 37, 53, True, -3, 0, 30, 50, True),       # generic + flip (per-pixel mirrored path) + clipping
    (1, 6, 9, 8, 44, 72, True, 0, 4, 44, 80, False),       # generic rows, OW % 8 == 0: mirrored group path
])
def test_upsample_softmax_accum(N, IH, IW, C, OH, OW, flip, dy0, dx0, dh, dw, weights):
    """eval_chip + window accumulation (reference: evaluate.py:74-87,139-146) against torch fp32 on the CPU."""
    lib = _lib.load()
    x = gen(N, IH, IW, C, seed=1, scale=2.0)
    xf = gen(N, IH, IW, C, seed=2, scale=2.0) if flip else None
    prob0 = torch.rand(N, C, dh, dw, generator=torch.Generator().manual_seed(3))
    wy = 1.0 / torch.randint(1, 4, (OH,), generator=torch.Generator().manual_seed(4)).float() if weights else None
    wx = 1.0 / torch.randint(1, 4, (OW,), generator=torch.Generator().manual_seed(5)).float() if weights else None
    up = lambda t: F.interpolate(t.permute(0, 3, 1, 2), (OH, OW), mode="bilinear", align_corners=False)  # noqa: E731
    p = F.softmax(up(x), dim=1)
    if flip:
        p = (p + F.softmax(torch.flip(up(xf), dims=(3,)), dim=1)) * 0.5
    if weights:
        p = p * wy.view(1, 1, OH, 1) * wx.view(1, 1, 1, OW)
    want = prob0.clone()
    ys = [y for y in range(OH) if 0 <= dy0 + y < dh]
    xs = [c for c in range(OW) if 0 <= dx0 + c < dw]
    want[:, :, dy0 + ys[0]:dy0 + ys[-1] + 1, dx0 + xs[0]:dx0 + xs[-1] + 1] += 0.75 * p[:, :, ys[0]:ys[-1] + 1, xs[0]:xs[-1] + 1]
    xd, pd = x.cuda(), prob0.cuda()
    xfd = xf.cuda() if flip else None
    wyd, wxd = (wy.cuda(), wx.cuda()) if weights else (None, None)
    check(lib.cabinet_upsample_softmax_accum(xd.data_ptr(), xfd.data_ptr() if flip else None, N, IH, IW, C, OH, OW,
                                             pd.data_ptr(), pd.stride(0), pd.stride(1), pd.stride(2), dy0, dx0, dh, dw,
                                             wyd.data_ptr() if weights else None, wxd.data_ptr() if weights else None,
                                             0.75, stream()), "softmax_accum")
    torch.cuda.synchronize()
    got = pd.cpu()
    assert rel_l2(got - prob0, want - prob0) < 2e-5
    assert float((got - want).abs().max()) < 2e-5
    untouched = torch.ones(dh, dw, dtype=torch.bool)
    untouched[dy0 + ys[0]:dy0 + ys[-1] + 1, dx0 + xs[0]:dx0 + xs[-1] + 1] = False
    assert torch.equal(got[:, :, untouched], prob0[:, :, untouched])  # nothing outside the window is written


@pytest.mark.parametrize("N,C,SH,SW,crop,H,W", [
    (2, 3, 40, 56, (0, 0, 40, 56), 80, 112),     # x2 up (input rescale of the 2.0 scale)
    (1, 8, 48, 64, (4, 6, 36, 45), 72, 90),      # un-pad crop + resize back
    (2, 5, 33, 47, (0, 0, 33, 47), 22, 31),      # down
    (1, 4, 20, 24, (0, 0, 20, 24), 20, 24),      # identity
])
def test_prob_resize_accum(N, C, SH, SW, crop, H, W):
    """un-pad + F.interpolate back + `probs +=` (reference: evaluate.py:152-158,218) against torch on the CPU."""
    lib = _lib.load()
    src = gen(N, C, SH, SW, seed=6)
    dst0 = gen(N, C, H, W, seed=7)
    y0, x0, ch, cw = crop
    want = dst0 + F.interpolate(src[:, :, y0:y0 + ch, x0:x0 + cw], (H, W), mode="bilinear", align_corners=False)
    sd, dd = src.cuda(), dst0.cuda()
    check(lib.cabinet_prob_resize_accum(sd.data_ptr(), N, C, SH, SW, y0, x0, ch, cw, dd.data_ptr(), H, W, stream()), "resize")
    torch.cuda.synchronize()
    assert float((dd.cpu() - want).abs().max()) < 1e-5
    with pytest.raises(ValueError):
        check(lib.cabinet_prob_resize_accum(sd.data_ptr(), N, C, SH, SW, y0, x0, SH + 1, cw, dd.data_ptr(), H, W, stream()), "resize")


@pytest.mark.parametrize("C,H,W", [(8, 37, 53), (19, 64, 64), (1, 5, 5)])
def test_argmax_hist_nchw(C, H, W):
    """torch.argmax(probs, 1) + compute_hist (reference: evaluate.py:222-228,162-191): bit-exact."""
    from oracle.evaluator_oracle import compute_hist

    lib = _lib.load()
    N = 3
    probs = torch.rand(N, C, H, W, generator=torch.Generator().manual_seed(8))
    probs[:, :, ::3, ::5] = 0.25  # ties: the first maximum must win
    labels = torch.randint(0, C + 2, (N, H, W), generator=torch.Generator().manual_seed(9))
    labels[:, H // 2, :] = 255
    want = torch.argmax(probs, dim=1)
    h = sum(compute_hist(want[i].numpy(), labels[i].numpy(), C, 255) for i in range(N))
    pd = probs.cuda()
    for ldt, lab in ((0, labels.cuda()), (1, labels.to(torch.uint8).cuda())):
        mask = torch.full((N, H, W), 77, dtype=torch.uint8, device="cuda")
        hist = torch.zeros(C, C, dtype=torch.int64, device="cuda")
        check(lib.cabinet_argmax_hist_nchw(pd.data_ptr(), N, C, H * W, mask.data_ptr(), lab.data_ptr(), ldt, 255,
                                           hist.data_ptr(), stream()), "argmax_hist")
        torch.cuda.synchronize()
        assert torch.equal(mask.cpu().long(), want)
        np.testing.assert_array_equal(hist.cpu().numpy(), h)
    hist = torch.zeros(C, C, dtype=torch.int64, device="cuda")
    check(lib.cabinet_argmax_hist_nchw(pd.data_ptr(), N, C, H * W, None, labels.cuda().data_ptr(), 0, 255,
                                       hist.data_ptr(), stream()), "argmax_hist")  # hist only
    np.testing.assert_array_equal(hist.cpu().numpy(), h)
    check(lib.cabinet_argmax_hist_nchw(pd.data_ptr(), 0, C, H * W, None, 1, 0, 255, hist.data_ptr(), stream()), "empty")


@pytest.mark.parametrize("cin,cout,k,p,H,W,UH,UW,act", [
    (128, 256, 1, 0, 32, 32, 8, 8, ACT_RELU),      # the FFM case: flat 1x1 tile walk, exact x4
    (128, 256, 1, 0, 9, 13, 3, 4, ACT_RELU),       # odd sizes: generic bilinear ratio, tiles straddle images
    (64, 48, 1, 0, 20, 12, 5, 3, ACT_NONE),        # Cout = 48 (three 16-column chunks), no activation
    (32, 64, 3, 1, 16, 24, 4, 6, ACT_RELU),        # 3x3: patch tile walk (per-image coordinates)
])
def test_conv_tc_upsample_add_epilogue(cin, cout, k, p, H, W, UH, UW, act):
    """y = act(conv(x) + b + bilinear(up -> H x W)): the FFM convblk with the x4 upsample commuted behind the 1x1
    conv (reference: cabinet.py:228-231,143-144), against fp32 torch on the CPU."""
    lib = _lib.load()
    dtype = torch.bfloat16
    N = 3
    x = q(gen(N, cin, H, W, seed=1), dtype)
    w = q(gen(cout, cin, k, k, seed=2, scale=(cin * k * k) ** -0.5), dtype)
    b = gen(cout, seed=3, scale=0.1)
    up = gen(N, cout, UH, UW, seed=4)
    ref = act_ref(F.conv2d(x, w, b, 1, p) + F.interpolate(up, (H, W), mode="bilinear", align_corners=False), act)
    xm = to_map(x, dtype, ld=cin + 16, off=8)
    ym = to_map(torch.zeros(N, cout, H, W), dtype, ld=cout + 24, off=16)
    ym.t.fill_(7.0)
    upd = up.permute(0, 2, 3, 1).contiguous().cuda()
    n16, c64 = -(-cout // 16) * 16, -(-cin // 64) * 64
    pk = torch.zeros(n16, k * k, c64)
    pk[:cout, :, :cin] = w.permute(0, 2, 3, 1).reshape(cout, k * k, cin)
    pk = pk.to("cuda", dtype).contiguous()
    bd = b.cuda()
    check(lib.cabinet_conv_tc_up(xm.ptr, xm.ld, N, H, W, cin, pk.data_ptr(), cout, k, k, 1, p, bd.data_ptr(),
                                 upd.data_ptr(), UH, UW, ym.ptr, ym.ld, H, W, act, stream()), "conv_tc_up")
    torch.cuda.synchronize()
    err = rel_l2(from_map(ym), ref)
    print(f"conv_tc_up cin={cin} cout={cout} k={k} {H}x{W} <- {UH}x{UW}: rel_l2 {err:.3e}")
    assert err < 6e-3
    full = ym.t.float()
    assert float((full[..., : ym.off] - 7.0).abs().max()) == 0 and float((full[..., ym.off + cout:] - 7.0).abs().max()) == 0
    with pytest.raises(ValueError):  # Cout must be a multiple of 16
        check(lib.cabinet_conv_tc_up(xm.ptr, xm.ld, N, H, W, cin, pk.data_ptr(), cout - 3, k, k, 1, p, bd.data_ptr(),
                                     upd.data_ptr(), UH, UW, ym.ptr, ym.ld, H, W, act, stream()), "conv_tc_up")
