"""All-reduce cost at the learner's gradient size (dev tool): NCCL vs torch symmetric-memory kernels, N ranks."""
import os, sys, time
import torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
dist.init_process_group("nccl", device_id=dev)
n = 2311697  # critic parameters at cfg-3
x = torch.randn(n, device=dev)
def timeit(fn, reps=200):
    for _ in range(10): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
t = timeit(lambda: dist.all_reduce(x))
if rank == 0: print(f"NCCL all_reduce {n*4/1e6:.1f} MB x{world}: {t:.1f} us")
for nn in (2311697 // 8, 2311697 * 4):
    y = torch.randn(nn, device=dev)
    t = timeit(lambda: dist.all_reduce(y))
    if rank == 0: print(f"NCCL all_reduce {nn*4/1e6:.1f} MB: {t:.1f} us")
try:
    import torch.distributed._symmetric_memory as symm
    npad = (n + 1023) // 1024 * 1024
    buf = symm.empty(npad, dtype=torch.float32, device=dev)
    hdl = symm.rendezvous(buf, dist.group.WORLD.group_name)
    buf.normal_()
    for name in ("one_shot_all_reduce", "two_shot_all_reduce_", "multimem_all_reduce_"):
        op = getattr(torch.ops.symm_mem, name, None)
        if op is None:
            if rank == 0: print(name, "not available")
            continue
        try:
            t = timeit(lambda: op(buf, "sum", dist.group.WORLD.group_name))
            if rank == 0: print(f"symm_mem.{name} {npad*4/1e6:.1f} MB: {t:.1f} us")
        except Exception as e:
            if rank == 0: print(name, "failed:", repr(e)[:200])
except Exception as e:
    if rank == 0: print("symmetric memory unavailable:", repr(e)[:300])
dist.barrier()
dist.destroy_process_group()
