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
            + names.count("cabinet_mbconv_noexpand_fused") + names.count("cabinet_mbconv_fused")) == n_blocks + 3
    assert names.count("cabinet_upsample_logits_nchw") == 2
    assert names.count("cabinet_psp_pool") == 2 and names.count("cabinet_softmax_rows") + names.count("cabinet_attention_tc") == 1
    # the single gap-sum / scratch memset (+ the V transpose inside attention_tc)
    extra = 1 + names.count("cabinet_attention_tc")
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
