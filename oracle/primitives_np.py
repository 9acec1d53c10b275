"""ORACLE (test infrastructure) — numpy restatement of the index rules the CUDA kernels implement.

The two resampling primitives of the path are ATen functions (third-party to the reference):
``upsample_bilinear2d(align_corners=False)`` and ``adaptive_avg_pool2d``.  Their published
index rules are restated here in plain numpy and pinned against ATen in
``tests/test_oracle_golden.py``; the reference calls them at ``src/models/cabinet.py:228-245``
and ``src/models/cab.py:54,68-72``.
"""

import numpy as np


def bilinear_taps(out_size: int, in_size: int):
    """Source taps of ``F.interpolate(mode='bilinear', align_corners=False)`` along one axis.

    src = max(0, (dst + 0.5) * in/out - 0.5); i0 = floor(src); i1 = min(i0 + 1, in - 1); w1 = src - i0.
    Computed in float32 like ATen's CUDA/CPU kernels (area_pixel_compute_source_index).
    """
    scale = np.float32(in_size) / np.float32(out_size)
    dst = np.arange(out_size, dtype=np.float32)
    src = np.maximum(np.float32(0), (dst + np.float32(0.5)) * scale - np.float32(0.5)).astype(np.float32)
    i0 = np.minimum(np.floor(src).astype(np.int64), in_size - 1)
    i1 = np.minimum(i0 + 1, in_size - 1)
    w1 = (src - i0.astype(np.float32)).astype(np.float32)
    return i0, i1, w1


def bilinear_resize(x: np.ndarray, oh: int, ow: int) -> np.ndarray:
    """x: (..., H, W) float32 -> (..., oh, ow)."""
    H, W = x.shape[-2:]
    y0, y1, wy = bilinear_taps(oh, H)
    x0, x1, wx = bilinear_taps(ow, W)
    top = x[..., y0, :]
    bot = x[..., y1, :]
    rows = top * (1 - wy)[:, None] + bot * wy[:, None]
    return (rows[..., :, x0] * (1 - wx) + rows[..., :, x1] * wx).astype(np.float32)


def adaptive_bins(out_size: int, in_size: int):
    """Bin [start, end) of ``adaptive_avg_pool2d`` along one axis: floor(i*in/out), ceil((i+1)*in/out)."""
    i = np.arange(out_size, dtype=np.int64)
    start = (i * in_size) // out_size
    end = -((-(i + 1) * in_size) // out_size)
    return start, end


def adaptive_avg_pool(x: np.ndarray, s: int) -> np.ndarray:
    """x: (..., H, W) -> (..., s, s) with (possibly overlapping) ATen bins."""
    H, W = x.shape[-2:]
    hs, he = adaptive_bins(s, H)
    ws, we = adaptive_bins(s, W)
    out = np.empty(x.shape[:-2] + (s, s), dtype=np.float32)
    for i in range(s):
        for j in range(s):
            out[..., i, j] = x[..., hs[i]:he[i], ws[j]:we[j]].mean(axis=(-2, -1))
    return out
