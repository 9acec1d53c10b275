"""Per-call trace of one bf16 training step (Large, batch 8, 1024x1024): python tools/trace_train.py [min_ms]"""
import sys

import torch

sys.path.insert(0, ".")
from cabinet_b200.loss import OhemCELoss  # noqa: E402
from cabinet_b200.synthetic import build_model, make_input, make_labels  # noqa: E402

MIN_MS = float(sys.argv[1]) if len(sys.argv) > 1 else 0.12
B, S, C = 8, 1024, 8
model = build_model(C, "large").cuda().train()
model.train_precision = "bf16"
model.logits_dtype = torch.bfloat16
x, lb = make_input(B, S, S).cuda(), make_labels(B, S, S, C).cuda()
crit = OhemCELoss(0.7, B * S * S // 16, 255)


def step():
    model.zero_grad(set_to_none=True)
    out, out16 = model(x)
    (crit(out, lb) + crit(out16, lb)).backward()


for _ in range(2):
    step()
eng = model.train_engine()
# keep (M, C) of the BatchNorm calls: achieved bandwidth per call (backward = 5 tensor passes, statistics 1, apply 2)
shapes = []
convs = []
orig_call = eng._call


def spy(name, *args):
    if name == "cabinet_bn_train_backward":
        shapes.append((len(eng.trace), name, int(args[11]), int(args[12]), 5))
    elif name == "cabinet_bn_train_stats":
        shapes.append((len(eng.trace), name, int(args[3]), int(args[4]), 1))
    elif name == "cabinet_conv_wgrad_tc":  # N, H, W, Cin, Cout, k, stride: bytes = x + dy read once
        N, H, W, Cin, Cout, k, st = (int(args[i]) for i in (5, 6, 7, 8, 9, 10, 12))
        convs.append((len(eng.trace), name, f"N{N} {H}x{W} {Cin}->{Cout} k{k} s{st}", 2 * N * H * W * (Cin + Cout / (st * st))))
    elif name == "cabinet_dwconv_wgrad":
        N, H, W, C, k, st = (int(args[i]) for i in (6, 7, 8, 9, 10, 11))
        convs.append((len(eng.trace), name, f"N{N} {H}x{W} C{C} k{k} s{st}", 2 * N * H * W * C * (1 + 1 / (st * st))))
    return orig_call(name, *args)


eng._call = spy
eng.start_trace()
step()
rows = eng.stop_trace()
eng._call = orig_call
for idx, name, M, C, passes in shapes:
    ms = rows[idx][2]
    if ms >= 0.05:
        print(f"{idx:5d} {name:28s} M {M:8d} C {C:4d} {ms:6.3f} ms  {passes * M * C * 2 / ms / 1e6:7.0f} GB/s")
for idx, name, desc, nbytes in convs:
    ms = rows[idx][2]
    if ms >= 0.03:
        print(f"{idx:5d} {name:24s} {desc:32s} {ms:6.3f} ms  {nbytes / ms / 1e6:7.0f} GB/s")
tot = sum(r[2] for r in rows)
print(f"{len(rows)} calls, {tot:.2f} ms traced")
fam = {}
for i, (n, ph, ms) in enumerate(rows):
    fam.setdefault((n, ph), [0, 0.0])
    fam[(n, ph)][0] += 1
    fam[(n, ph)][1] += ms
    if ms >= MIN_MS:
        print(f"{i:5d} {ph} {n:34s} {ms:7.3f} ms")
for (n, ph), (c, ms) in sorted(fam.items(), key=lambda kv: -kv[1][1])[:20]:
    print(f"{ph} {n:34s} x{c:3d} {ms:7.2f} ms")
