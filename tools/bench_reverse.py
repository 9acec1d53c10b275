import sys, torch
sys.path.insert(0, ".")
from cabinet_b200.synthetic import build_model, make_input
model = build_model(8, "large").cuda()
model.logits_dtype = torch.bfloat16
model.use_cuda_graph = True
x = make_input(16, 1024, 1024).cuda()
ref = None
for layers in ((), ("sb.conv2",), ("mobile.f4.dw",), ("ffm.convblk",), ("sb.conv2", "mobile.f4.dw"), (), ("sb.conv2", "mobile.f4.dw")):
    model.repack()  # fresh engine (and graph pool) per configuration
    eng = model.engine()
    eng.reverse_layers = frozenset(layers)
    for _ in range(4): out = model(x)
    torch.cuda.synchronize()
    if ref is None: ref = out[0].clone()
    same = torch.equal(ref, out[0])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): model(x)
    e1.record(); torch.cuda.synchronize()
    print("reverse", layers, f"{e0.elapsed_time(e1)/20:.3f} ms/step identical={same}", flush=True)
