"""Host-side helpers that need no GPU."""
def test_device_prefetcher_passes_batches_through_on_cpu():
    import torch

    from cabinet_b200.prefetch import DevicePrefetcher

    batches = [(torch.full((2, 3), float(i)), torch.full((2,), i)) for i in range(4)]
    out = list(DevicePrefetcher(batches, "cpu"))
    assert len(out) == 4 and all(torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) for a, b in zip(out, batches))
    assert [t[0].shape for t in DevicePrefetcher([torch.zeros(1)], "cpu")] == [torch.Size([1])]
