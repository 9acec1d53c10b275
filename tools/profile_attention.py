"""Attention kernel alone at the bench shape (16 images, 32x32 tokens, d = 128): python tools/profile_attention.py"""
import sys, torch
sys.path.insert(0, ".")
from cabinet_b200.engine import Map
from cabinet_b200.synthetic import build_model
model = build_model(8, "large").cuda()
eng = model.engine()
mk = lambda: Map(torch.randn(16, 32, 32, 128, device="cuda").to(torch.bfloat16), 16, 32, 32, 128, 128)
q, k, v = mk(), mk(), mk()
for _ in range(3):
    eng.attention(q, k, v)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    eng.attention(q, k, v)
e1.record(); torch.cuda.synchronize()
print(f"attention (transpose + kernel): {e0.elapsed_time(e1) / 10 * 1e3:.1f} us")
