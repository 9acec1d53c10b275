"""CPU: the oracle restatement reproduces the imported reference's committed outputs (tests/golden)."""

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from cabinet_b200.constants import BACKBONE_CFGS
from cabinet_b200.synthetic import build_model, make_input, make_labels, state_dict_digest
from oracle import cabinet_oracle, evaluator_oracle, primitives_np
from oracle.make_golden import EVAL_CASES, MODEL_CASES, TinySegModel


@pytest.mark.parametrize("name,mode,C,shape", MODEL_CASES)
def test_forward_matches_reference_golden(golden_dir, name, mode, C, shape):
    g = np.load(golden_dir / f"model_{name}.npz")
    model = build_model(C, mode)  # drop-in tree, same seed + perturbation as the reference side
    sd = model.state_dict()
    assert state_dict_digest(sd) == str(g["digest"]), "seeded drop-in weights differ from the reference's"
    stages = {}
    final, aux = cabinet_oracle.cabinet_forward(sd, make_input(*shape), BACKBONE_CFGS[mode], stages)
    # same ATen primitives, same order of operations: fp32 round-off only
    np.testing.assert_allclose(final.numpy(), g["final"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(aux.numpy(), g["aux"], rtol=1e-5, atol=1e-6)
    for k, v in stages.items():
        np.testing.assert_allclose(v.numpy(), g[f"stage_{k}"], rtol=1e-5, atol=1e-6, err_msg=k)


@pytest.mark.parametrize("case", EVAL_CASES, ids=[c[0] for c in EVAL_CASES])
def test_evaluator_matches_reference_golden(golden_dir, case):
    name, C, (N, H, W), nb, scales, flip, crop = case
    g = np.load(golden_dir / f"evaluator_{name}.npz")
    model = TinySegModel(C).eval()
    batches = [(make_input(N, H, W, seed=70 + b), make_labels(N, H, W, C, seed=110 + b)) for b in range(nb)]
    res = evaluator_oracle.evaluate(model, batches, C, 255, scales, flip, crop)
    np.testing.assert_array_equal(res["confusion_matrix"], g["hist"])  # integer counts: bit-exact
    assert res["mIoU"] == pytest.approx(float(g["miou"]), rel=0, abs=1e-12)
    assert res["accuracy"] == pytest.approx(float(g["acc"]), rel=0, abs=1e-12)


@pytest.mark.parametrize("insz,outsz", [(32, 128), (128, 1024), (3, 34), (68, 270), (9, 70), (1, 7), (8, 3)])
def test_bilinear_rule_matches_aten(insz, outsz):
    x = torch.randn(2, 3, insz, insz + 1, generator=torch.Generator().manual_seed(1))
    ref = F.interpolate(x, size=(outsz, outsz + 3), mode="bilinear", align_corners=False).numpy()
    got = primitives_np.bilinear_resize(x.numpy(), outsz, outsz + 3)
    # the source coordinate is an fp32 number of magnitude ~in_size: ATen's FMA contraction moves it by 1 ulp
    # (7.6e-6 at 128), which moves the tap weight by the same amount -> atol ~ 1e-4 on N(0,1) data
    np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("insz", [32, 34, 3, 2, 68, 5])
@pytest.mark.parametrize("s", [1, 3, 6, 8])
def test_adaptive_pool_rule_matches_aten(insz, s):
    x = torch.randn(2, 4, insz, insz + 2, generator=torch.Generator().manual_seed(2))
    ref = F.adaptive_avg_pool2d(x, (s, s)).numpy()
    got = primitives_np.adaptive_avg_pool(x.numpy(), s)
    np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-5)


def test_compute_hist_edge_cases():
    C = 4
    pred = np.array([[0, 1, 2, 3], [3, 3, 9, -2]])  # out-of-range preds are clipped, like the reference
    label = np.array([[0, 1, 255, 3], [0, 7, 2, 1]])  # 255 ignored, 7 clipped to 3
    h = evaluator_oracle.compute_hist(pred, label, C)
    assert h.sum() == 7 and h[0, 0] == 1 and h[3, 3] == 2 and h[3, 2] == 1 and h[0, 1] == 1 and h[3, 0] == 1
    assert evaluator_oracle.compute_hist(np.zeros((0, 5)), np.zeros((0, 5)), C).sum() == 0
    assert evaluator_oracle.compute_hist(np.zeros((2, 2)), np.full((2, 2), 255), C).sum() == 0


# ------------------------------------------------------------------ training loss (OhemCELoss)
from oracle.loss_oracle import OHEM_CASES, make_case, ohem_ce_loss  # noqa: E402


@pytest.mark.parametrize("case", OHEM_CASES, ids=[c[0] for c in OHEM_CASES])
def test_loss_oracle_matches_reference_golden(golden_dir, case):
    """oracle/loss_oracle.py against outputs of the imported reference OhemCELoss (oracle/make_golden_loss.py)."""
    name, shape, thresh, n_min, ignore, weighted, scale, quant = case
    g = np.load(golden_dir / f"ohem_{name}.npz")
    logits, labels, weight = make_case(shape, ignore, weighted, scale, quant)
    assert float(logits.double().sum()) == float(g["logits_sum"]) and int(labels.sum()) == int(g["labels_sum"])
    x = logits.clone().requires_grad_(True)
    loss = ohem_ce_loss(x, labels, thresh, n_min, 255, weight)
    loss.backward()
    assert float(loss) == pytest.approx(float(g["loss"]), rel=1e-6, abs=1e-7)
    grad = x.grad if x.grad is not None else torch.zeros_like(x)
    np.testing.assert_allclose(grad.numpy(), g["grad"], rtol=1e-5, atol=1e-8)


# ------------------------------------------------------------------ training step (forward + backward, train-mode BN)
from oracle.train_oracle import FULL_GRAD_KEYS, STAT_KEYS, TRAIN_CASES, train_step  # noqa: E402


@pytest.mark.parametrize("case", TRAIN_CASES, ids=[c[0] for c in TRAIN_CASES])
def test_train_step_oracle_matches_reference_golden(golden_dir, case):
    """oracle/train_oracle.py (train-mode BN forward, OHEM loss on both outputs, autograd) against the imported
    reference's loss, per-parameter gradient norms, four full gradients and updated BN running statistics
    (oracle/make_golden_train.py; reference: src/scripts/train.py:430-436)."""
    name, mode, C, (N, H, W), thresh, n_min = case
    g = np.load(golden_dir / f"train_step_{name}.npz")
    model = build_model(C, mode)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    assert state_dict_digest(sd) == str(g["digest"])
    loss, grads, running = train_step(sd, make_input(N, H, W), make_labels(N, H, W, C), BACKBONE_CFGS[mode], thresh, n_min)
    assert float(loss) == pytest.approx(float(g["loss"]), rel=1e-5)
    want = dict(zip([str(k) for k in g["grad_keys"]], g["grad_norms"]))
    assert set(want) == set(grads)
    worst = 0.0
    for k, ref_norm in want.items():
        if ref_norm < 0:  # the backbone's unused classifier receives no gradient
            assert grads[k] is None and k.startswith("mobile.classifier")
            continue
        got = float(grads[k].norm())
        worst = max(worst, abs(got - ref_norm) / max(ref_norm, 1e-12))
    assert worst < 1e-3, worst
    for k in FULL_GRAD_KEYS:
        ref = torch.from_numpy(g["grad__" + k])
        assert float((grads[k] - ref).norm() / ref.norm()) < 1e-4, k
    for k in STAT_KEYS:
        np.testing.assert_allclose(running[k][0].numpy(), g["mean__" + k], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(running[k][1].numpy(), g["var__" + k], rtol=1e-5, atol=1e-6)
