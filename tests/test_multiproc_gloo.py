"""CPU, world_size 2 over gloo: the N > 1 path of the evaluator — contiguous sharding of the units (images) over
ranks, per-rank int64 confusion matrices, ONE all-reduce, metrics identical on every rank and equal to the
single-process reference result (oracle compute_hist over all images)."""

import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cabinet_b200.evaluator import metrics_from_hist, reduce_hist, shard_range
from oracle.evaluator_oracle import compute_hist, metrics

C, N_IMG, H, W = 6, 7, 24, 31


def make_data():
    g = torch.Generator().manual_seed(3)
    preds = torch.randint(0, C, (N_IMG, H, W), generator=g)
    labels = torch.randint(0, C, (N_IMG, H, W), generator=g)
    labels[:, 5, :] = 255
    return preds, labels


def worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    preds, labels = make_data()
    lo, hi = shard_range(N_IMG, rank, world)
    hist = torch.zeros(C, C, dtype=torch.int64)
    for i in range(lo, hi):
        hist += torch.from_numpy(compute_hist(preds[i].numpy(), labels[i].numpy(), C, 255))
    reduce_hist(hist)
    res = metrics_from_hist(hist)
    out[rank] = (hist.numpy().copy(), float(res["mIoU"]), (lo, hi))
    dist.destroy_process_group()


def test_shard_ranges_cover_everything_once():
    for n in (0, 1, 7, 16, 17):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(120)
def test_two_rank_hist_allreduce_matches_single_process():
    world, port = 2, 29500 + os.getpid() % 2000
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(worker, args=(world, port, out), nprocs=world, join=True)
    preds, labels = make_data()
    want = sum(compute_hist(preds[i].numpy(), labels[i].numpy(), C, 255) for i in range(N_IMG))
    for r in range(world):
        hist, miou, _ = out[r]
        np.testing.assert_array_equal(hist, want)  # integer counts: bit-exact on every rank
        assert miou == pytest.approx(float(metrics(want)["mIoU"]), abs=1e-12)
    assert out[0][2] == (0, 4) and out[1][2] == (4, 7)
