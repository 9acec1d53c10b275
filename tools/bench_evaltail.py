"""Device time of the evaluator-tail kernels at 1024x1024 chips (CUDA events, L2 flushed by the working set)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from cabinet_b200 import _lib  # noqa: E402

lib = _lib.load()
N, C, S = 16, 8, 1024
st = torch.cuda.current_stream().cuda_stream
m = torch.randn(N, S // 8, S // 8, C, device="cuda")
mf = torch.randn(N, S // 8, S // 8, C, device="cuda")
prob = torch.zeros(N, C, S, S, device="cuda")
inv = torch.full((S,), 0.5, device="cuda")
labels = torch.randint(0, C, (N, S, S), device="cuda", dtype=torch.uint8)
hist = torch.zeros(C, C, dtype=torch.int64, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timeit(name, fn, nbytes, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:34s} {ms * 1e3:8.1f} us  {nbytes / ms / 1e6:8.0f} GB/s algorithmic ({nbytes / 1e6:.0f} MB)", flush=True)


pb = prob.numel() * 4
for flip in (False, True):
    timeit(f"upsample_softmax_accum flip={flip}",
           lambda: _lib.check(lib.cabinet_upsample_softmax_accum(m.data_ptr(), mf.data_ptr() if flip else None, N, S // 8,
                                                                 S // 8, C, S, S, prob.data_ptr(), prob.stride(0),
                                                                 prob.stride(1), prob.stride(2), 0, 0, S, S,
                                                                 inv.data_ptr(), inv.data_ptr(), 1.0, st)), 2 * pb)
src = torch.randn(N, C, 1536, 1536, device="cuda")
timeit("prob_resize_accum 1536->1024", lambda: _lib.check(lib.cabinet_prob_resize_accum(
    src.data_ptr(), N, C, 1536, 1536, 0, 0, 1536, 1536, prob.data_ptr(), S, S, st)), src.numel() * 4 + 2 * pb)
timeit("prob_resize_accum identity", lambda: _lib.check(lib.cabinet_prob_resize_accum(
    prob.data_ptr(), N, C, S, S, 0, 0, S, S, prob.data_ptr(), S, S, st)), 3 * pb)
timeit("argmax_hist_nchw", lambda: _lib.check(lib.cabinet_argmax_hist_nchw(
    prob.data_ptr(), N, C, S * S, None, labels.data_ptr(), 1, 255, hist.data_ptr(), st)), pb + labels.numel())
