"""All-reduce time of the config-5 gradient volume (3 fp32 buckets, 14 MB) and of one bf16 bucket: torchrun ... tools/nccl_probe.py"""
import os, torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
bufs = [torch.ones(n, device="cuda") for n in (1_200_000, 1_200_000, 1_100_000)]
half = torch.ones(3_500_000, device="cuda", dtype=torch.bfloat16)
one = torch.ones(3_500_000, device="cuda")
def timed(fn, n=20):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
def three():
    hs = [dist.all_reduce(b, async_op=True) for b in bufs]
    for h in hs: h.wait()
t3 = timed(three); t1 = timed(lambda: dist.all_reduce(one)); th = timed(lambda: dist.all_reduce(half))
if rank == 0:
    print(f"world {world}: 3 fp32 buckets {t3:.3f} ms, one 14 MB fp32 {t1:.3f} ms, one 7 MB bf16 {th:.3f} ms")
dist.destroy_process_group()
