import sys, torch
sys.path.insert(0, ".")
from cabinet_b200.synthetic import build_model, make_input, make_labels
from cabinet_b200.evaluator import MscEvalV0
C, S, B, K = 8, 1024, 16, 10
model = build_model(C, "large").cuda()
x = make_input(B, S, S).pin_memory(); lb = make_labels(B, S, S, C).to(torch.uint8).pin_memory()
valid = int((lb != 255).sum())
for pipelined in (False, True):
    ev = MscEvalV0(model, [(x, lb)] * K, C, 255, (1.0,), False, cropsize=S); ev.pipelined = pipelined
    res = ev.evaluate()
    print(pipelined, res["confusion_matrix"].sum(), valid * K)
hist = torch.zeros(C, C, dtype=torch.int64, device="cuda")
m = model.accumulate_hist(x.cuda(), lb.cuda(), hist); print("direct", int(hist.sum()), valid)
