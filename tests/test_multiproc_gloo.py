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


# ------------------------------------------------------------------ training config: bucketed gradient all-reduce
def grad_worker(rank, world, port, out):
    from cabinet_b200.grad_sync import GradBuckets
    from cabinet_b200.synthetic import build_model

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    model = build_model(8, "small").train()
    gb = GradBuckets(model.named_parameters(), bucket_bytes=1 << 20)
    for i, (n, p) in enumerate(model.named_parameters()):
        if p.grad is not None:
            p.grad.add_(float(rank + 1) * (1 + i % 5))  # in place, like autograd's accumulation into the view
    gb.all_reduce()
    named = dict(model.named_parameters())
    ok = all(torch.allclose(named[n].grad, torch.full_like(named[n], 1.5 * (1 + i % 5)))
             for i, (n, p) in enumerate(model.named_parameters()) if p.grad is not None)
    out[rank] = (ok, len(gb.buckets), gb.skipped, [len(x) for x in gb.names], gb.names[0][0],
                 sum(b.numel() for b in gb.buckets))
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_bucketed_gradient_allreduce():
    """Config-5 plumbing: gradients live in flat buckets (reverse parameter order), one async all-reduce per bucket,
    averaged over the ranks; the backbone's unused classifier is left out."""
    from cabinet_b200.grad_sync import GradBuckets
    from cabinet_b200.synthetic import build_model

    world, port = 2, 31500 + os.getpid() % 2000
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(grad_worker, args=(world, port, out), nprocs=world, join=True)
    model = build_model(8, "small")
    n_all = sum(p.numel() for p in model.parameters())
    n_cls = sum(p.numel() for n, p in model.named_parameters() if n.startswith("mobile.classifier"))
    for r in range(world):
        ok, n_buckets, skipped, sizes, first, total = out[r]
        assert ok and n_buckets >= 2 and sum(sizes) == len(list(model.parameters())) - 4
        assert sorted(skipped) == ["mobile.classifier.0.bias", "mobile.classifier.0.weight", "mobile.classifier.3.bias",
                                   "mobile.classifier.3.weight"]
        assert total == n_all - n_cls
        assert first == list(dict(model.named_parameters()))[-1]  # the last parameter's gradient is ready first
    # single process: views are installed, all_reduce is a no-op
    gb = GradBuckets(model.named_parameters())
    p = model.conv_out.conv_out.weight
    p.grad.fill_(2.0)
    gb.all_reduce()
    assert float(p.grad.mean()) == 2.0 and p.grad.data_ptr() >= gb.buckets[0].data_ptr()
    gb.zero_()
    assert float(p.grad.abs().max()) == 0.0


def test_grad_buckets_survive_optimizer_zero_grad():
    """``optimizer.zero_grad()`` defaults to set_to_none=True, which detaches every ``p.grad`` from its flat
    bucket; the next backward then allocates fresh gradients.  The buckets must pick those up again (two steps)."""
    from cabinet_b200.grad_sync import GradBuckets

    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.ReLU(), torch.nn.Linear(7, 3))
    gb = GradBuckets(net.named_parameters(), bucket_bytes=64, skip=lambda n: False)
    opt = torch.optim.SGD(net.parameters(), lr=0.1)
    x = torch.randn(4, 5)
    for step in range(2):
        opt.zero_grad()  # set_to_none=True: the views are gone
        net(x).square().sum().backward()
        want = [p.grad.clone() for p in net.parameters()]
        with pytest.raises(RuntimeError):
            gb.attach(strict=True)
        gb.all_reduce()  # re-attaches (copies the fresh gradients into the buckets) before reducing
        flat = torch.cat([b for b in gb.buckets])
        assert float(flat.abs().sum()) > 0
        for p, w in zip(net.parameters(), want):
            assert torch.equal(p.grad, w)
            assert any(b.data_ptr() <= p.grad.data_ptr() < b.data_ptr() + b.numel() * 4 for b in gb.buckets)
        opt.step()
    gb.zero_()  # in place: the views stay
    assert gb.attach(strict=True) == 0 and all(float(p.grad.abs().max()) == 0 for p in net.parameters())
