"""CPU: drives the engine's kernel schedule against a recording test double of the C-ABI.

No arithmetic happens (the double returns 0 without touching memory): this checks the host logic —
shape bookkeeping, concat-slice strides, argument counts/types of every ctypes call, launch counting —
so that a GPU run is not wasted on a Python typo.  Numerical parity lives in the ``-m gpu`` tests.
"""

import ctypes

import pytest
import torch

from cabinet_b200 import _lib, engine as engine_mod
from cabinet_b200.synthetic import build_model


class RecordingLib:
    def __init__(self):
        self.calls = []

    def __getattr__(self, name):
        if name not in _lib.SIGNATURES:
            raise AttributeError(name)
        argtypes, _ = _lib.SIGNATURES[name]

        def fn(*args):
            assert len(args) == len(argtypes), f"{name}: {len(args)} args, ABI has {len(argtypes)}"
            for a, t in zip(args, argtypes):
                t.from_param(a) if hasattr(t, "from_param") else t(a)  # raises on a wrong Python type
            self.calls.append((name, args))
            return 0

        return fn


@pytest.fixture
def cpu_engine(monkeypatch):
    rec = RecordingLib()
    monkeypatch.setattr(_lib, "load", lambda: rec)
    monkeypatch.setattr(engine_mod.Engine, "stream", property(lambda self: None))

    def make(mode, C, precision):
        model = build_model(C, mode)
        real_next = next

        class P:  # pretend the parameters live on a CUDA device
            is_cuda = True
            device = torch.device("cpu")

        monkeypatch.setattr(engine_mod, "next", lambda it: P, raising=False)
        eng = engine_mod.Engine(model, precision)
        monkeypatch.setattr(engine_mod, "next", real_next, raising=False)
        return eng, rec

    return make


@pytest.mark.parametrize("mode,C,shape,precision", [
    ("small", 8, (1, 3, 96, 128), "fp32"),
    ("large", 19, (2, 3, 64, 64), "bf16"),
    ("large", 8, (1, 3, 70, 100), "bf16"),
])
def test_schedule_shapes_and_abi(cpu_engine, mode, C, shape, precision):
    eng, rec = cpu_engine(mode, C, precision)
    x = torch.zeros(shape)
    final, aux = eng.forward(x)
    assert final.shape == (shape[0], C, shape[2], shape[3]) and aux.shape == final.shape
    names = [c[0] for c in rec.calls]
    n_blocks = 15 if mode == "large" else 11
    assert (names.count("cabinet_dwconv") + names.count("cabinet_dwconv_tma")
            + names.count("cabinet_mbconv_noexpand_fused") + names.count("cabinet_mbconv_fused")
            + names.count("cabinet_mbconv_t")) == n_blocks + 3
    assert names.count("cabinet_upsample_logits_nchw") == 2
    assert names.count("cabinet_psp_pool") == 2 and names.count("cabinet_softmax_rows") + names.count("cabinet_attention_tc") == 1
    # the single gap-sum / scratch memset (+ the V transpose inside attention_tc)
    # (+ one zeroed scratch per SE-block channel_sum of the fp32 parity mode; the FFM one shares the memset)
    extra = 1 + names.count("cabinet_attention_tc") + names.count("cabinet_channel_sum") - 1
    assert eng.launches == len(rec.calls) + extra
    rec.calls.clear()
    mask = eng.forward_mask(x)
    assert mask.shape == (shape[0], shape[2], shape[3]) and mask.dtype == torch.uint8
    hist = torch.zeros(C, C, dtype=torch.int64)
    eng.forward_hist(x, torch.zeros(shape[0], shape[2], shape[3], dtype=torch.int64), hist)
    with pytest.raises(ValueError):
        eng.forward_hist(x, torch.zeros(shape[0], shape[2], shape[3], dtype=torch.int32), hist)


def test_dropin_refuses_cpu_and_train_mode():
    m = build_model(8, "small")
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.zeros(1, 3, 64, 64))
    with pytest.raises(ValueError):
        m(torch.zeros(1, 4, 64, 64))


@pytest.mark.parametrize("precision", ["bf16"])
def test_mask_path_skips_aux_head_and_optional_schedules(cpu_engine, precision):
    """Host logic of the schedule variants: mask / class-map callers do not launch the auxiliary head; the folded low
    path calls cabinet_conv_tc_up; reverse_layers ORs CABINET_CONV_REVERSE_TILES into the layer's act argument."""
    eng, rec = cpu_engine("large", 8, precision)
    x = torch.zeros(2, 3, 64, 64)
    eng.forward(x)
    full = [c[0] for c in rec.calls]
    rec.calls.clear()
    eng.forward_mask(x)
    mask = [c[0] for c in rec.calls]
    assert full.count("cabinet_bilinear_nhwc") == 2 and mask.count("cabinet_bilinear_nhwc") == 1  # low_up only
    assert mask.count("cabinet_upsample_argmax") == 1 and mask.count("cabinet_upsample_logits_nchw") == 0
    n_conv = lambda names: sum(n.startswith("cabinet_conv_tc") or n == "cabinet_conv2d_simt" for n in names)  # noqa: E731
    assert n_conv(full) - n_conv(mask) == 2  # b1 and b4
    rec.calls.clear()
    cm = eng.class_map8(x)
    assert tuple(cm.shape) == (2, 8, 8, 8) and cm.dtype == torch.float32
    assert [c[0] for c in rec.calls] == mask[:-1]  # the same trunk without the argmax tail
    # folded low path
    rec.calls.clear()
    eng.fold_low_up = True
    eng.forward_mask(x)
    names = [c[0] for c in rec.calls]
    assert names.count("cabinet_conv_tc_up") == 1 and names.count("cabinet_bilinear_nhwc") == 0
    up = next(c[1] for c in rec.calls if c[0] == "cabinet_conv_tc_up")
    assert up[5] == 128 and up[7] == 256 and (up[14], up[15]) == (2, 2) and (up[18], up[19]) == (8, 8)
    eng.fold_low_up = False
    # reversed tile walk
    rec.calls.clear()
    eng.reverse_layers = frozenset({"sb.conv2", "mobile.f4.dw"})
    eng.fuse_mbconv = False
    eng.forward_mask(x)
    flagged = [c for c in rec.calls if c[0] in ("cabinet_conv_tc_se", "cabinet_dwconv_tma") and (c[1][-2 if c[0] == "cabinet_conv_tc_se" else -3] & 0x100)]
    assert len(flagged) == 2


def test_general_mode_window_geometry_matches_oracle(monkeypatch):
    """Host glue of the fused multi-scale evaluator: for padded, exact and multi-window images the (window, destination
    offset, clip extent, weight slice) arguments handed to cabinet_upsample_softmax_accum reproduce the oracle's
    pad_tensor / chip_windows geometry (reference: evaluate.py:60-72,95-146)."""
    from cabinet_b200 import evaluator as ev_mod
    from oracle import evaluator_oracle

    rec = RecordingLib()
    monkeypatch.setattr(_lib, "load", lambda: rec)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda dev=None: type("S", (), {"cuda_stream": 0})())

    class FakeModel(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.p = torch.nn.Parameter(torch.zeros(1))
            self.seen = []

        def class_map8(self, x):
            self.seen.append(tuple(x.shape))
            return torch.zeros(x.shape[0], x.shape[2] // 8, x.shape[3] // 8, 4)

    for (H, W), cs, flip in (((40, 56), 64, True), ((64, 64), 64, False), ((100, 150), 64, True), ((50, 200), 64, False)):
        model = FakeModel()
        ev = ev_mod.MscEvalV0(model, None, 4, 255, (1.0,), flip, cropsize=cs, device=torch.device("cpu"))
        ev.max_chip_batch = 8
        N = 2
        image, dst = torch.zeros(N, 3, H, W), torch.zeros(N, 4, H, W)
        rec.calls.clear()
        ev._crop_eval_into(image, dst)
        calls = [c[1] for c in rec.calls if c[0] == "cabinet_upsample_softmax_accum"]
        # oracle geometry
        if H < cs or W < cs:
            tgt = (cs, cs) if max(H, W) < cs else (cs if H < W else H, cs if W < H else W)
            _, (hst, hed, wst, wed) = evaluator_oracle.pad_tensor(image, tgt)
            fh, fw = tgt
        else:
            hst = wst = 0
            fh, fw = H, W
        wins = evaluator_oracle.chip_windows(fh, fw, cs)
        assert len(calls) == len(wins)
        inv_base = {}
        for a, (y0, y1, x0, x1) in zip(calls, wins):
            assert (a[6], a[7]) == (cs, cs) and (a[2], a[3], a[4], a[5]) == (N, cs // 8, cs // 8, 4)
            assert (a[12], a[13]) == (y0 - hst, x0 - wst) and (a[14], a[15]) == (H, W)
            assert (a[9], a[10], a[11]) == (dst.stride(0), dst.stride(1), dst.stride(2))
            assert (a[1] is not None) == flip
            # weight slices start at the window origin of the per-axis inverse-count vectors
            inv_base.setdefault("y", a[16] - 4 * y0)
            inv_base.setdefault("x", a[17] - 4 * x0)
            assert a[16] - 4 * y0 == inv_base["y"] and a[17] - 4 * x0 == inv_base["x"]
        per_win = N * (2 if flip else 1)
        assert sum(s[0] for s in model.seen) == len(wins) * per_win and all(s[2:] == (cs, cs) for s in model.seen)
        assert all(s[0] <= max(8, per_win) for s in model.seen)


def test_packed_state_roundtrip_on_the_host(cpu_engine, tmp_path, monkeypatch):
    """Engine.packed_state() -> torch.save -> torch.load -> Engine(packed=...) reproduces every packed tensor (the
    device-side cache test lives in tests/test_checkpoint.py)."""
    eng, rec = cpu_engine("large", 8, "bf16")
    state = eng.packed_state()
    assert set(state) == set(engine_mod.Engine.PACKED_ATTRS)
    torch.save({"packed": state}, tmp_path / "w.pack")
    loaded = torch.load(tmp_path / "w.pack", weights_only=False)["packed"]

    def tensors(o, out):
        if isinstance(o, torch.Tensor):
            out.append(o)
        elif isinstance(o, dict):
            for k in sorted(o, key=str):
                tensors(o[k], out)
        elif isinstance(o, (list, tuple)):
            for v in o:
                tensors(v, out)
        elif hasattr(o, "__dict__"):
            tensors(vars(o), out)
        return out

    a, b = tensors(state, []), tensors(loaded, [])
    assert len(a) == len(b) > 150 and all(torch.equal(x, y) for x, y in zip(a, b))

    class P:
        is_cuda = True
        device = torch.device("cpu")

    model = build_model(8, "large")
    monkeypatch.setattr(engine_mod, "next", lambda it: P, raising=False)
    eng2 = engine_mod.Engine(model, "bf16", packed=loaded)
    assert eng2.qkv[0].shape == eng.qkv[0].shape and eng2.key_ch == eng.key_ch
    with pytest.raises(ValueError):
        engine_mod.Engine(model, "bf16", packed={k: v for k, v in loaded.items() if k != "stem"})


@pytest.mark.parametrize("mode,C,shape,precision", [("small", 8, (2, 3, 64, 96), "fp32"), ("large", 5, (2, 3, 64, 64), "bf16")])
def test_train_engine_schedule_and_abi(monkeypatch, mode, C, shape, precision):
    """Host logic of the training step against the recording double: every ctypes call of the forward and of the
    backward tape has the ABI's argument count / types, every trained parameter gets a gradient buffer, the gradient
    regions of the concat buffers are covered before they are read."""
    from cabinet_b200 import train_engine as te

    rec = RecordingLib()
    rec.cabinet_train_scratch_floats = lambda M, Cc, nq: 16
    rec.cabinet_conv_wgrad_scratch_floats = lambda *a: 16
    rec.cabinet_conv_wgrad_tc_scratch_floats = lambda *a: 16
    monkeypatch.setattr(_lib, "load", lambda: rec)
    monkeypatch.setattr(te.TrainEngine, "stream", property(lambda self: None))
    model = build_model(C, mode).train()

    class P:
        is_cuda = True
        device = torch.device("cpu")

    real_next = next
    monkeypatch.setattr(te, "next", lambda it: P, raising=False)
    eng = te.TrainEngine(model, precision)
    monkeypatch.setattr(te, "next", real_next, raising=False)
    x = torch.zeros(shape)
    final, aux = eng.forward(x)
    assert final.shape == (shape[0], C, shape[2], shape[3]) and aux.shape == final.shape
    n_fwd = len(rec.calls)
    names = [c[0] for c in rec.calls]
    n_bn = sum(1 for m in model.modules() if isinstance(m, torch.nn.BatchNorm2d))
    assert names.count("cabinet_bn_train_stats") == n_bn
    grads = eng.backward(torch.zeros_like(final), torch.zeros_like(aux))
    bwd = [c[0] for c in rec.calls[n_fwd:]]
    assert bwd.count("cabinet_bn_train_backward") == n_bn
    n_conv = names.count("cabinet_conv2d_simt") + names.count("cabinet_conv_tc") - 2   # minus the two attention GEMMs
    assert bwd.count("cabinet_conv_wgrad") + bwd.count("cabinet_conv_wgrad_tc") == n_conv
    # the two stems read the network input: no data gradient; bf16 mode: stride-1 data gradients are conv_tc calls
    assert bwd.count("cabinet_conv_dgrad") + bwd.count("cabinet_conv_tc") + bwd.count("cabinet_conv_tc_view") // 4 == n_conv - 2
    assert names.count("cabinet_im2col_nchw") == (1 if precision == "bf16" else 0)
    if precision == "bf16":
        assert names.count("cabinet_conv_tc") >= n_conv - 4 and bwd.count("cabinet_conv_tc") >= n_conv - 6
        assert names.count("cabinet_dwconv_tma") > 0 and bwd.count("cabinet_dwconv_tma") > 0
        assert bwd.count("cabinet_conv_wgrad_tc") >= n_conv - 8   # all but the stems, the stride-2 and the class-logit convs
    trained = {id(p): n for n, p in model.named_parameters() if not n.startswith("mobile.classifier")}
    assert set(grads) == set(trained), [trained[i] for i in set(trained) - set(grads)]
    assert all(grads[id(p)].shape == p.shape for p in model.parameters() if id(p) in grads)


def test_fused_block_routing(cpu_engine):
    """Host logic of the round-2 block kernels: which Large blocks go to the channel-major kernel, the one-kernel SE path
    (expand_sums -> two gate layers -> mbconv_t with the gate as input) for the stride-1 3x3 SE blocks, and the fall-backs
    when the flags are off."""
    eng, rec = cpu_engine("large", 8, "bf16")
    x = torch.zeros(1, 3, 128, 128)
    eng.forward_mask(x)
    calls = [(c[0], c[1]) for c in rec.calls]
    names = [c[0] for c in calls]
    assert names.count("cabinet_stem_tc2") == 1 and names.count("cabinet_stem_tc") == 0
    # f2 (stride 2, 64 expanded channels) stays on the pixel-major kernel, f3 .. f15 are channel-major
    assert names.count("cabinet_mbconv_fused") == 1 and names.count("cabinet_mbconv_t") == 13
    assert names.count("cabinet_expand_sums") == 2          # f11, f12
    with_gate = [a for n, a in calls if n == "cabinet_mbconv_t" and a[-2] is not None]
    assert len(with_gate) == 2 and all(a[13] is not None for a in with_gate)   # project mode with se_scale
    assert sorted(a[8] for a in with_gate) == [480, 672]                       # Cexp of f11 / f12
    # a sums pass is followed by the two gate layers on the fixed-point accumulator it filled
    i = names.index("cabinet_expand_sums")
    assert names[i + 1:i + 4] == ["cabinet_gate_fc", "cabinet_gate_fc", "cabinet_mbconv_t"]
    assert calls[i + 1][1][0] == calls[i][1][-2] and calls[i + 1][1][-2] == 1     # in = gap_sum, in_fixed = 1
    rec.calls.clear()
    eng.se_from_sums = False
    eng.use_stem2 = False
    eng.forward_mask(x)
    names = [c[0] for c in rec.calls]
    assert names.count("cabinet_expand_sums") == 0 and names.count("cabinet_stem_tc") == 1
    assert names.count("cabinet_mbconv_t") == 13            # f11 / f12: expand + depthwise, then scale / project kernels
    rec.calls.clear()
    eng.use_mbconv_t = False
    eng.forward_mask(x)
    names = [c[0] for c in rec.calls]
    assert names.count("cabinet_mbconv_t") == 0 and names.count("cabinet_mbconv_fused") >= 8
