"""Helpers for the -m gpu tests: NCHW cpu tensors <-> NHWC device Maps, error metrics."""

import torch

from cabinet_b200.engine import Map


def to_map(x_nchw: torch.Tensor, dtype, ld=None, off=0, dev="cuda") -> Map:
    N, C, H, W = x_nchw.shape
    ld = ld or C
    buf = torch.zeros((N, H, W, ld), dtype=dtype, device=dev)
    buf[..., off:off + C] = x_nchw.permute(0, 2, 3, 1).to(dev, dtype)
    return Map(buf, N, H, W, C, ld, off)


def from_map(m: Map) -> torch.Tensor:
    return m.nhwc().float().permute(0, 3, 1, 2).cpu()


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def tol(dtype):
    """Op-level tolerances: fp32 kernels vs fp32 torch 1e-5 rel-L2; bf16 storage one rounding (2^-9) of noise."""
    return 2e-5 if dtype == torch.float32 else 6e-3
