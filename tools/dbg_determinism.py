"""Graph-vs-eager bit equality probe (mirrors tests/test_gpu_parity.py::test_cuda_graph_and_sub_batch_...)."""
import sys, torch
sys.path.insert(0, ".")
from cabinet_b200.synthetic import build_model, make_input
for fold in (True, False, True):
    model = build_model(8, "large").cuda()
    model.logits_dtype = torch.bfloat16
    model.engine().fold_ffm = fold
    x = make_input(4, 256, 256).cuda()
    f0 = model(x)[0].clone()
    model.sub_batch = 2
    f1 = model(x)[0].clone()
    model.use_cuda_graph = True
    for _ in range(3):
        f2 = model(x)[0].clone()
    x2 = make_input(4, 256, 256, seed=99).cuda()
    f3 = model(x2)[0].clone()
    f3b = model(x2)[0].clone()
    model.use_cuda_graph, model.sub_batch = False, 0
    f4 = model(x2)[0].clone()
    model.sub_batch = 2
    f5 = model(x2)[0].clone()
    d = lambda a, b: float((a.float() - b.float()).abs().max())
    print("fold", fold, "eager-vs-sub", d(f0, f1), "graph(x)", d(f0, f2), "graph(x2) vs eager", d(f3, f4), "replay2", d(f3b, f4),
          "eager sub(x2) vs eager", d(f5, f4), "fold flag now", model.engine().fold_ffm)
