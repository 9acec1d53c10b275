"""Config surface of the CABiNet forward path.

Mirrors the values (not the code) of the reference's
``src/models/constants.py:10-27`` and ``configs/model/mobilenetv3_{large,small}.yaml``.
"""

# reference: src/models/constants.py:10-19 (attention_planes = expanded width of the last block)
MODEL_CONFIG = {
    "large": {"attention_planes": 960, "output_channel": 1280},
    "small": {"attention_planes": 576, "output_channel": 1024},
}

# reference: src/models/constants.py:26-27
EVAL_STRIDE_RATE = 5 / 6.0
DEFAULT_EVAL_SCALES = [0.5, 0.75, 1.0, 1.25, 1.5, 1.75]
DEFAULT_IGNORE_LABEL = 255

# reference: configs/model/mobilenetv3_large.yaml:5-21 / mobilenetv3_small.yaml:5-17
# rows are [kernel, expand_ratio, out_channels, use_se, use_hs, stride]
BACKBONE_CFGS = {
    "large": [
        [3, 1, 16, 0, 0, 1],
        [3, 4, 24, 0, 0, 2],
        [3, 3, 24, 0, 0, 1],
        [5, 3, 40, 1, 0, 2],
        [5, 3, 40, 1, 0, 1],
        [5, 3, 40, 1, 0, 1],
        [3, 6, 80, 0, 1, 2],
        [3, 2.5, 80, 0, 1, 1],
        [3, 2.3, 80, 0, 1, 1],
        [3, 2.3, 80, 0, 1, 1],
        [3, 6, 112, 1, 1, 1],
        [3, 6, 112, 1, 1, 1],
        [5, 6, 160, 1, 1, 2],
        [5, 6, 160, 1, 1, 1],
        [5, 6, 160, 1, 1, 1],
    ],
    "small": [
        [3, 1, 16, 1, 0, 2],
        [3, 4.5, 24, 0, 0, 2],
        [3, 3.67, 24, 0, 0, 1],
        [5, 4, 40, 1, 1, 2],
        [5, 6, 40, 1, 1, 1],
        [5, 6, 40, 1, 1, 1],
        [5, 3, 48, 1, 1, 1],
        [5, 3, 48, 1, 1, 1],
        [5, 6, 96, 1, 1, 2],
        [5, 6, 96, 1, 1, 1],
        [5, 6, 96, 1, 1, 1],
    ],
}

PSP_SIZES = (1, 3, 6, 8)  # reference: src/models/cab.py:54
BN_EPS = 1e-5  # torch.nn.BatchNorm2d default, used everywhere in the reference


def make_divisible(v, divisor=8, min_value=None):
    """Channel rounding rule of the backbone (reference: src/models/mobilenetv3.py:18-35)."""
    floor = divisor if min_value is None else min_value
    r = max(floor, int(v + divisor / 2) // divisor * divisor)
    return r + divisor if r < 0.9 * v else r


def resolve_blocks(cfgs, width_mult=1.0):
    """Expand cfg rows into concrete block specs.

    Returns (stem_out, [dict(inp, exp, out, k, s, se, hs, identity, expand)], last_exp).
    reference: src/models/mobilenetv3.py:172-185 (note: the final 1x1 conv re-uses the loop's last exp_size, F4).
    """
    inp = make_divisible(16 * width_mult)
    stem = inp
    blocks, exp = [], inp
    for k, t, c, se, hs, s in cfgs:
        out = make_divisible(c * width_mult)
        exp = make_divisible(inp * t)
        blocks.append(
            dict(inp=inp, exp=exp, out=out, k=int(k), s=int(s), se=bool(se), hs=bool(hs),
                 identity=(int(s) == 1 and inp == out), expand=(inp != exp))
        )
        inp = out
    return stem, blocks, exp
