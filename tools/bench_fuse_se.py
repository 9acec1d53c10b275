import sys, torch
sys.path.insert(0, ".")
from cabinet_b200.synthetic import build_model, make_input
model = build_model(8, "large").cuda()
model.logits_dtype = torch.bfloat16
model.use_cuda_graph = True
x = make_input(16, 1024, 1024).cuda()
for fuse in (False, "sel", True, False, "sel", True):
    model._engine = None if hasattr(model, "_engine") else None
    eng = model.engine()
    eng.fuse_se = fuse is True
    eng.fuse_se_blocks = frozenset({"mobile.f12", "mobile.f13"}) if fuse == "sel" else frozenset()
    eng._graphs.clear(); eng._graph_seen.clear()
    for _ in range(4): model(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): model(x)
    e1.record(); torch.cuda.synchronize()
    print("fuse_se", fuse, f"{e0.elapsed_time(e1)/20:.3f} ms/step")
