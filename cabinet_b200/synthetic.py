"""Seeded synthetic weights / inputs / labels shared by tests, ``bench.py`` and ``smoke()``.

There are no pretrained blobs and no datasets offline, and a *fresh* reference model is a
degenerate parity target (``gamma = 0`` silences the whole attention path and fresh BN is the
identity).  ``perturb_state_dict`` therefore randomises, from one seeded generator
and in ``state_dict`` key order (identical between the reference and the drop-in): every BN's
running stats and affine, every conv/linear bias, and sets ``gamma = 0.5``.
"""

from __future__ import annotations

import hashlib

import torch


def perturb_state_dict(sd: dict, seed: int = 123, gamma: float = 0.5) -> dict:
    """Returns a new CPU fp32 state_dict (same keys/order) with the perturbation applied."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    bn_prefixes = {k[: -len(".running_mean")] for k in sd if k.endswith(".running_mean")}
    for k, v in sd.items():
        v = v.detach().cpu().clone()
        prefix, _, leaf = k.rpartition(".")
        if leaf == "gamma":
            v.fill_(gamma)
        elif prefix in bn_prefixes:
            if leaf == "running_mean" or leaf == "bias":
                v.copy_(0.1 * torch.randn(v.shape, generator=g))
            elif leaf == "running_var" or leaf == "weight":
                v.copy_(0.5 + torch.rand(v.shape, generator=g))
        elif leaf == "bias":
            v.copy_(0.1 * torch.randn(v.shape, generator=g))
        out[k] = v
    return out


def state_dict_digest(sd: dict) -> str:
    """sha256 over keys + raw bytes, used to check that two builds hold the same weights."""
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(v.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def make_input(n: int, h: int, w: int, seed: int = 7) -> torch.Tensor:
    """Normalised-image-like input: N(0,1) fp32 NCHW (dataset tensors are mean/std normalised)."""
    return torch.randn(n, 3, h, w, generator=torch.Generator().manual_seed(seed))


def make_labels(n: int, h: int, w: int, n_classes: int, seed: int = 11, ignore_label: int = 255) -> torch.Tensor:
    """int64 labels with an ignore stripe (rows h//4 .. h//4 + max(1, h//32))."""
    lb = torch.randint(0, n_classes, (n, h, w), generator=torch.Generator().manual_seed(seed))
    lb[:, h // 4: h // 4 + max(1, h // 32), :] = ignore_label
    return lb


def build_model(n_classes: int, mode: str, seed: int = 0, perturb: bool = True):
    """Seeded drop-in model with perturbed weights (CPU, eval)."""
    from .constants import BACKBONE_CFGS
    from .modules import CABiNet

    torch.manual_seed(seed)
    m = CABiNet(n_classes, mode=mode, cfgs=BACKBONE_CFGS[mode])
    if perturb:
        m.load_state_dict(perturb_state_dict(m.state_dict()))
    return m.eval()
