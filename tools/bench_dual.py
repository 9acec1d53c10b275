import sys, torch
sys.path.insert(0, ".")
from cabinet_b200.synthetic import build_model, make_input
B = 16
model = build_model(8, "large").cuda(); model.logits_dtype = torch.bfloat16; model.use_cuda_graph = True
x = make_input(B, 1024, 1024).cuda()
ref = None
for dual in (False, True):
    eng = model.engine(); eng.dual_stream = dual
    for _ in range(3): out = model(x)
    torch.cuda.synchronize()
    if ref is None: ref = out[0].clone()
    else: print("identical:", torch.equal(ref, out[0]))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): model(x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"dual_stream={dual}: {ms:.3f} ms/step {B/ms*1e3:.0f} img/s")
