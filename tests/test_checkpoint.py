"""Checkpoint / weight interop (cabinet_b200/checkpoint.py; reference: evaluate.py:259-267, train.py:126-176)."""

import pytest
import torch

from cabinet_b200 import checkpoint
from cabinet_b200.synthetic import build_model, make_input


def test_raw_and_wrapped_checkpoints(tmp_path):
    m = build_model(8, "small")
    sd = m.state_dict()
    torch.save(sd, tmp_path / "model_best.pth")
    torch.save({"model_state": sd, "epoch": 3, "optimizer_state": {}}, tmp_path / "checkpoint_last.pth")
    for name in ("model_best.pth", "checkpoint_last.pth"):
        got = checkpoint.load_model_weights(tmp_path / name)
        assert list(got) == list(sd) and all(torch.equal(got[k], sd[k]) for k in sd)
        build_model(8, "small", seed=5).load_state_dict(got, strict=True)


def test_warm_start_takes_name_and_shape_matches_only(tmp_path):
    src = build_model(8, "small", seed=1)
    sd = dict(src.state_dict())
    sd["not.in.the.model"] = torch.zeros(3)
    torch.save({"model_state": sd}, tmp_path / "uavid.pth")
    dst = build_model(19, "small", seed=2)   # other dataset: the two class heads are sized by n_classes
    before = {k: v.clone() for k, v in dst.state_dict().items()}
    loaded, mismatch, unknown = checkpoint.load_pretrained(dst, tmp_path / "uavid.pth")
    # the two class heads and the backbone's (unused) classifier output layer are sized by n_classes
    assert sorted(mismatch) == ["ab.b4.bias", "ab.b4.weight", "conv_out.conv_out.weight", "mobile.classifier.3.bias",
                                "mobile.classifier.3.weight"]
    assert unknown == ["not.in.the.model"]
    assert len(loaded) == len(before) - 5
    after = dst.state_dict()
    for k in loaded:
        assert torch.equal(after[k], sd[k])
    for k in mismatch:
        assert torch.equal(after[k], before[k])  # left at their fresh initialisation


def test_packed_cache_path():
    assert str(checkpoint.packed_cache_path("/x/model_best.pth")) == "/x/model_best.pth.cabinet_b200.bf16.pack"


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_packed_weight_cache_roundtrip(tmp_path, precision):
    x = make_input(2, 96, 128).cuda()
    m = build_model(8, "large").cuda()
    m.precision = precision
    want = [t.clone() for t in m(x)]
    path = checkpoint.save_packed(m, checkpoint.packed_cache_path(tmp_path / "w.pth", precision))
    m2 = build_model(8, "large").cuda()
    m2.precision = precision
    assert checkpoint.load_packed(m2, path)
    eng = m2.__dict__["_engine"]
    got = m2(x)
    assert m2.__dict__["_engine"] is eng                  # the cached pack is the one that ran (no repack)
    assert all(torch.equal(a, b) for a, b in zip(got, want))  # both modes: every reduction has a fixed order
    m3 = build_model(8, "large", seed=3).cuda()           # other weights: the digest differs, the cache is ignored
    m3.precision = precision
    assert not checkpoint.load_packed(m3, path)
    m4 = build_model(8, "large").cuda()
    m4.precision = "fp32" if precision == "bf16" else "bf16"
    assert not checkpoint.load_packed(m4, path)           # other precision
    assert not checkpoint.load_packed(m2, tmp_path / "missing.pack")
    with torch.no_grad():
        m2.ab.a2block.gamma.fill_(0.25)                   # a weight change after loading still invalidates the pack
    assert not torch.equal(m2(x)[0], want[0])


def test_packed_state_is_plain_and_loads_with_weights_only(tmp_path):
    """the pack cache must not need full unpickling.  Layer objects are stored as tagged dicts."""
    from cabinet_b200.engine import ConvLayer, DwLayer, GateLayer, _from_plain, _to_plain

    conv = torch.nn.Conv2d(8, 16, 3, stride=2, padding=1, bias=False)
    bn = torch.nn.BatchNorm2d(16).eval()
    dw = torch.nn.Conv2d(16, 16, 5, padding=2, groups=16, bias=False)
    tree = {"blocks": [dict(spec={"k": 3, "se": True}, act=2, pw1=ConvLayer(conv, bn, 1, torch.bfloat16, "a"),
                            dw=DwLayer(dw, bn, 0, "b"), fused_pw=(torch.ones(2), torch.zeros(2)))],
            "gate": GateLayer(torch.randn(4, 16), None, torch.randn(16, 4), torch.randn(16), 3), "none": None}
    plain = _to_plain(tree)
    torch.save({"header": {"format": 3}, "packed": plain}, tmp_path / "p.pack")
    back = _from_plain(torch.load(tmp_path / "p.pack", weights_only=True)["packed"])
    b0, b1 = tree["blocks"][0], back["blocks"][0]
    assert isinstance(b1["pw1"], ConvLayer) and isinstance(b1["dw"], DwLayer) and isinstance(back["gate"], GateLayer)
    assert isinstance(b1["fused_pw"], tuple) and b1["spec"] == b0["spec"] and back["none"] is None
    for a, b in ((b0["pw1"], b1["pw1"]), (b0["dw"], b1["dw"]), (tree["gate"], back["gate"])):
        assert vars(a).keys() == vars(b).keys()
        for k, v in vars(a).items():
            w = vars(b)[k]
            assert torch.equal(v, w) if isinstance(v, torch.Tensor) else v == w
