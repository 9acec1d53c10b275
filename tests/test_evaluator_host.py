"""CPU: host-side logic of the device evaluator (window grid / overlap counts, rank sharding) against the oracle."""

import numpy as np
import pytest

from cabinet_b200.evaluator import MscEvalV0, shard_range
from oracle import evaluator_oracle


@pytest.mark.parametrize("fh,fw,cs", [(64, 64, 64), (96, 120, 64), (144, 90 + 64, 64), (1024, 2048, 1024),
                                      (2160, 3840, 1024), (1620, 2880, 1024), (48, 180, 48)])
def test_window_grid_matches_reference_windows_and_counts(fh, fw, cs):
    """The count map of crop_eval (reference: evaluate.py:127-149) is the outer product of per-axis counts."""
    ys, inv_y = MscEvalV0.window_grid(fh, cs)
    xs, inv_x = MscEvalV0.window_grid(fw, cs)
    wins = evaluator_oracle.chip_windows(fh, fw, cs)
    assert [(y, y + cs, x, x + cs) for y in ys for x in xs] == [tuple(w) for w in wins]
    count = np.zeros((fh, fw), dtype=np.float32)
    for y0, y1, x0, x1 in wins:
        count[y0:y1, x0:x1] += 1
    np.testing.assert_array_equal(np.outer(inv_y, inv_x), (1.0 / np.maximum(count, 1)).astype(np.float32))
    assert count.min() >= 1


def test_shard_range_partitions_every_item_once():
    for n in (0, 1, 7, 16, 33):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


class _ConstantModel:
    """Stand-in of the reference test's ConstantModel (tests/integration/test_training_pipeline.py:259-274)."""

    def __init__(self, n_classes, value):
        self.n_classes, self.value = n_classes, value

    def __call__(self, x):
        import torch

        logits = torch.full((x.shape[0], self.n_classes, x.shape[2], x.shape[3]), -1e9)
        logits[:, self.value] = 1e9
        return logits, logits


@pytest.mark.parametrize("n_classes,cropsize,size,value", [(4, 64, 100, 0), (3, 48, 96, 1)])
def test_oracle_overlap_normalisation_is_uniform(n_classes, cropsize, size, value):
    """reference tests/integration/test_training_pipeline.py:276-338: a constant model must give a spatially uniform
    probability map whatever the window overlap (checks the oracle's count normalisation)."""
    import torch

    prob = evaluator_oracle.crop_eval(_ConstantModel(n_classes, value), torch.zeros(1, 3, size, size), n_classes,
                                      cropsize, False)
    assert (prob.argmax(dim=1) == value).all()
    p = prob[0, value]
    assert float(p.max() - p.min()) < 1e-5


def test_cpulist_parser_of_the_numa_binding():
    from cabinet_b200.affinity import _parse_cpulist

    assert _parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert _parse_cpulist("5") == {5} and _parse_cpulist("") == set()
