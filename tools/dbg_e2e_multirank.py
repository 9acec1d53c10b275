"""Where does the end-to-end evaluator spend its time under torchrun?  torchrun --nproc-per-node 2 tools/dbg_e2e_multirank.py"""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, ".")
from cabinet_b200.synthetic import build_model, make_input, make_labels
from cabinet_b200.evaluator import MscEvalV0, reduce_hist, metrics_from_hist

rank = int(os.environ.get("LOCAL_RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(rank)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
B, S, C, K = 16, 1024, 8, 10
model = build_model(C, "large").cuda().eval()
model.use_cuda_graph = True
x = make_input(B, S, S, seed=rank).pin_memory(); lb = make_labels(B, S, S, C, seed=rank).to(torch.uint8).pin_memory()
masks = [torch.empty((B, S, S), dtype=torch.uint8).pin_memory() for _ in range(K)]
ev = MscEvalV0(model, [(x, lb)] * 4, C, 255, (1.0,), False, cropsize=S)
ev.evaluate(masks_out=masks)
ev.dl = [(x, lb)] * K
for trial in range(2):
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    dev = torch.device("cuda", rank)
    hist = torch.zeros((C, C), dtype=torch.int64, device=dev)
    t0 = time.perf_counter()
    ev._fast_pipelined(dev, hist, masks)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    reduce_hist(hist)
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    metrics_from_hist(hist)
    t4 = time.perf_counter()
    print(f"rank {rank} trial {trial}: enqueue {1e3*(t1-t0):.1f} ms, drain {1e3*(t2-t1):.1f} ms, all-reduce {1e3*(t3-t2):.1f} ms, metrics {1e3*(t4-t3):.1f} ms", flush=True)
if world > 1: dist.destroy_process_group()
