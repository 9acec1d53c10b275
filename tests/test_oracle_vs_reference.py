"""CPU, build container only: live comparison against the imported reference (skipped on the GPU box)."""

from pathlib import Path

import pytest
import torch

REF = Path("/root/reference/src/models/cabinet.py")
pytestmark = pytest.mark.skipif(not REF.is_file(), reason="/root/reference not mounted")


@pytest.fixture(scope="module")
def ref_cls():
    import sys

    sys.dont_write_bytecode = True
    from oracle.make_golden import import_reference

    return import_reference()[0]


@pytest.mark.parametrize("mode,C", [("small", 8), ("large", 19)])
def test_state_dict_bit_identical(ref_cls, mode, C):
    from cabinet_b200 import BACKBONE_CFGS, CABiNet

    torch.manual_seed(0)
    a = ref_cls(C, mode=mode, cfgs=BACKBONE_CFGS[mode]).state_dict()
    torch.manual_seed(0)
    b = CABiNet(C, mode=mode, cfgs=BACKBONE_CFGS[mode]).state_dict()
    assert list(a) == list(b)
    for k in a:
        assert a[k].dtype == b[k].dtype and torch.equal(a[k], b[k]), k


def test_strict_load_both_directions_and_get_params(ref_cls):
    from cabinet_b200 import BACKBONE_CFGS, CABiNet

    ref = ref_cls(8, mode="small", cfgs=BACKBONE_CFGS["small"])
    mine = CABiNet(8, mode="small", cfgs=BACKBONE_CFGS["small"])
    mine.load_state_dict(ref.state_dict(), strict=True)
    ref.load_state_dict(mine.state_dict(), strict=True)
    names_m = {id(p): n for n, p in mine.named_parameters()}
    names_r = {id(p): n for n, p in ref.named_parameters()}
    for gm, gr in zip(mine.get_params(), ref.get_params()):
        assert [names_m[id(p)] for p in gm] == [names_r[id(p)] for p in gr]


def test_oracle_live_small_512(ref_cls):
    """BASELINE config 1 shape (Small, 1x3x512x512, 8 classes, fp32 CPU)."""
    from cabinet_b200.constants import BACKBONE_CFGS
    from cabinet_b200.synthetic import make_input, perturb_state_dict
    from oracle import cabinet_oracle

    torch.manual_seed(0)
    ref = ref_cls(8, mode="small", cfgs=BACKBONE_CFGS["small"])
    sd = perturb_state_dict(ref.state_dict())
    ref.load_state_dict(sd)
    ref.eval()
    x = make_input(1, 512, 512)
    with torch.no_grad():
        f_ref, a_ref = ref(x)
    f, a = cabinet_oracle.cabinet_forward(sd, x, BACKBONE_CFGS["small"])
    torch.testing.assert_close(f, f_ref, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(a, a_ref, rtol=1e-5, atol=1e-6)
