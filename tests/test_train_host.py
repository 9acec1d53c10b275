"""CPU: host-side pieces of the training step -- the separable resampling operators (whose transposes the backward
kernels apply) against ATen's F.interpolate / adaptive_avg_pool2d, and the CSR packing."""

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from cabinet_b200.train_engine import adaptive_pool_matrix, bilinear_matrix, csr


@pytest.mark.parametrize("n_in,n_out", [(4, 16), (32, 128), (128, 1024), (17, 68), (68, 270), (3, 5), (8, 5), (1, 7), (2, 2)])
def test_bilinear_matrix_is_aten_align_corners_false(n_in, n_out):
    m = torch.from_numpy(bilinear_matrix(n_in, n_out)).float()
    x = torch.randn(2, 3, n_in, n_in, generator=torch.Generator().manual_seed(n_in * 131 + n_out))
    ref = F.interpolate(x, (n_out, n_out), mode="bilinear", align_corners=False)
    got = torch.einsum("oi,ncij,pj->ncop", m, x, m)
    assert torch.allclose(got, ref, atol=2e-5, rtol=1e-5)   # (ATen's CPU path rounds the tap weights a few ulp differently)
    assert np.allclose(bilinear_matrix(n_in, n_out).sum(1), 1.0)  # rows are convex weights


@pytest.mark.parametrize("n_in,n_out", [(32, 1), (32, 3), (32, 6), (32, 8), (2, 6), (3, 8), (68, 6), (120, 8), (5, 5)])
def test_adaptive_pool_matrix_is_aten(n_in, n_out):
    m = torch.from_numpy(adaptive_pool_matrix(n_in, n_out)).float()
    x = torch.randn(2, 3, n_in, n_in, generator=torch.Generator().manual_seed(n_in * 17 + n_out))
    ref = F.adaptive_avg_pool2d(x, (n_out, n_out))
    got = torch.einsum("oi,ncij,pj->ncop", m, x, m)
    assert torch.allclose(got, ref, atol=2e-6, rtol=1e-5)


def test_csr_round_trip_and_transpose():
    m = bilinear_matrix(5, 12)
    for mat in (m, m.T):
        start, idx, w = csr(mat)
        dense = np.zeros_like(mat)
        for o in range(mat.shape[0]):
            for a in range(start[o], start[o + 1]):
                dense[o, idx[a]] += w[a]
        assert np.allclose(dense, mat.astype(np.float32))
        assert start.dtype == np.int32 and idx.dtype == np.int32 and w.dtype == np.float32
    s, i, w = csr(np.zeros((3, 4)))  # an all-zero operator still yields valid (non-empty) arrays
    assert list(s) == [0, 0, 0, 0] and len(i) == 1 and len(w) == 1
