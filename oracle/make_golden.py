"""Generates ``tests/golden/*.npz`` by running the UNMODIFIED reference imported from /root/reference.

Run in the build container only (the reference does not travel to the GPU box):

    python oracle/make_golden.py

Model fixtures: seeded reference model (``torch.manual_seed(0)``) + ``perturb_state_dict`` (seed 123)
+ seeded N(0,1) input (seed 7) -> final/aux logits and stage activations, plus the sha256 of the
weights so the GPU box can prove it rebuilt the same model from the seed.
Evaluator fixtures: the reference ``MscEvalV0`` (imported through a 2-module hydra/omegaconf stub,
those packages are not installed offline) around a tiny deterministic conv model.
"""

from __future__ import annotations

import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.dont_write_bytecode = True

from cabinet_b200.constants import BACKBONE_CFGS  # noqa: E402
from cabinet_b200.synthetic import make_input, make_labels, perturb_state_dict, state_dict_digest  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"

MODEL_CASES = [  # name, mode, classes, (N, H, W)
    ("small_c8_96x128", "small", 8, (1, 96, 128)),
    ("large_c19_64x64", "large", 19, (1, 64, 64)),
    ("large_c8_70x100", "large", 8, (1, 70, 100)),  # odd sizes: 70->35->18->9->5->3, sb 9x13
]


def import_reference():
    sys.path.insert(0, "/root/reference")
    hydra = types.ModuleType("hydra")
    hydra.main = lambda *a, **k: (lambda f: f)
    oc = types.ModuleType("omegaconf")
    oc.DictConfig = dict
    oc.OmegaConf = type("OmegaConf", (), {"to_yaml": staticmethod(lambda c: str(c))})
    sys.modules.setdefault("hydra", hydra)
    sys.modules.setdefault("omegaconf", oc)
    from src.models.cabinet import CABiNet
    from src.scripts.evaluate import MscEvalV0

    return CABiNet, MscEvalV0


class TinySegModel(torch.nn.Module):
    """Deterministic stand-in model for the evaluator fixtures: fixed 3x3 conv -> C logits."""

    def __init__(self, n_classes, seed=5):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.w = torch.nn.Parameter(torch.randn(n_classes, 3, 3, 3, generator=g))

    def forward(self, x):
        y = torch.nn.functional.conv2d(x, self.w, padding=1)
        return y, y


EVAL_CASES = [  # name, classes, (N,H,W) per batch, batches, scales, flip, cropsize
    ("eval_fast_100_c64", 5, (2, 100, 100), 2, (1.0,), False, 64),
    ("eval_ms_flip_96_c48", 4, (1, 96, 120), 2, (0.5, 1.0, 1.5), True, 48),
    ("eval_pad_40_c64", 3, (2, 40, 56), 1, (1.0, 0.75), True, 64),
]


def main():
    GOLDEN.mkdir(parents=True, exist_ok=True)
    RefCABiNet, MscEvalV0 = import_reference()
    torch.set_num_threads(8)
    for name, mode, C, (N, H, W) in MODEL_CASES:
        torch.manual_seed(0)
        ref = RefCABiNet(C, mode=mode, cfgs=BACKBONE_CFGS[mode])
        sd = perturb_state_dict(ref.state_dict())
        ref.load_state_dict(sd)
        ref.eval()
        stages = {}
        hooks = [
            ref.sb.register_forward_hook(lambda m, i, o: stages.__setitem__("feat_sb", o)),
            ref.mobile.register_forward_hook(lambda m, i, o: stages.__setitem__("mobile_feat", o)),
            ref.ab.register_forward_hook(lambda m, i, o: stages.update(low=o[0], high=o[1])),
            ref.ffm.register_forward_hook(lambda m, i, o: stages.__setitem__("feat_fuse", o)),
            ref.conv_out.register_forward_hook(lambda m, i, o: stages.__setitem__("final8", o)),
        ]
        x = make_input(N, H, W)
        with torch.no_grad():
            final, aux = ref(x)
        for h in hooks:
            h.remove()
        np.savez_compressed(
            GOLDEN / f"model_{name}.npz", mode=mode, n_classes=C, shape=np.array([N, H, W]),
            digest=state_dict_digest(sd), final=final.numpy(), aux=aux.numpy(),
            **{f"stage_{k}": v.detach().numpy() for k, v in stages.items()})
        print(name, "final", tuple(final.shape), float(final.abs().mean()))

    for name, C, (N, H, W), nb, scales, flip, crop in EVAL_CASES:
        model = TinySegModel(C).eval()
        batches = [(make_input(N, H, W, seed=70 + b), make_labels(N, H, W, C, seed=110 + b)) for b in range(nb)]
        ev = MscEvalV0(model, batches, n_classes=C, ignore_label=255, scales=scales, flip=flip, cropsize=crop,
                       device=torch.device("cpu"))
        res = ev.evaluate()
        np.savez_compressed(GOLDEN / f"evaluator_{name}.npz", n_classes=C, shape=np.array([N, H, W]), batches=nb,
                            scales=np.array(scales), flip=flip, cropsize=crop, hist=res["confusion_matrix"],
                            miou=res["mIoU"], acc=res["accuracy"])
        print(name, "mIoU", res["mIoU"], "sum", res["confusion_matrix"].sum())


if __name__ == "__main__":
    main()
