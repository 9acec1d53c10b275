"""Graph-replay step time with one engine flag toggled: python tools/bench_flags.py <attr> [values...]"""
import sys, torch
sys.path.insert(0, ".")
from cabinet_b200.synthetic import build_model, make_input
attr = sys.argv[1]
vals = [eval(v) for v in sys.argv[2:]] or [False, True, False, True]
model = build_model(8, "large").cuda()
model.logits_dtype = torch.bfloat16
model.use_cuda_graph = True
x = make_input(16, 1024, 1024).cuda()
for val in vals:
    eng = model.engine()
    setattr(eng, attr, val)
    eng._graphs.clear(); eng._graph_seen.clear()
    for _ in range(4): model(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): model(x)
    e1.record(); torch.cuda.synchronize()
    print(attr, val, f"{e0.elapsed_time(e1)/20:.3f} ms/step")
