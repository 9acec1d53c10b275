import sys, time, torch
sys.path.insert(0, ".")
from cabinet_b200.loss import OhemCELoss
from cabinet_b200.synthetic import build_model, make_input, make_labels
B, S, C = 8, 1024, 8
model = build_model(C, "large").cuda().train()
model.train_precision = "bf16"; model.logits_dtype = torch.bfloat16
x, lb = make_input(B, S, S).cuda(), make_labels(B, S, S, C).cuda()
crit = OhemCELoss(0.7, B * S * S // 16, 255)
def step():
    model.zero_grad(set_to_none=True)
    out, out16 = model(x)
    (crit(out, lb) + crit(out16, lb)).backward()
for _ in range(3): step()
torch.cuda.synchronize()
for _ in range(3):
    t0 = time.perf_counter(); step(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"host enqueue {1e3*(t1-t0):.2f} ms, total {1e3*(t2-t0):.2f} ms")
# forward only / backward only host time
model.zero_grad(set_to_none=True)
torch.cuda.synchronize(); t0=time.perf_counter(); out,out16=model(x); t1=time.perf_counter(); torch.cuda.synchronize(); t2=time.perf_counter()
print(f"fwd host {1e3*(t1-t0):.2f} total {1e3*(t2-t0):.2f}")
loss = crit(out, lb) + crit(out16, lb)
torch.cuda.synchronize(); t0=time.perf_counter(); loss.backward(); t1=time.perf_counter(); torch.cuda.synchronize(); t2=time.perf_counter()
print(f"bwd host {1e3*(t1-t0):.2f} total {1e3*(t2-t0):.2f}")
