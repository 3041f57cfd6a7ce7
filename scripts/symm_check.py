"""Latency of an 8.7 KB all-reduce: NCCL vs torch symmetric-memory one-shot (run under torchrun)."""
import os, time, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = 1084
x = torch.randn(n, dtype=torch.float64, device="cuda")
def timeit(f, it=200):
    for _ in range(20): f()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3
t_nccl = timeit(lambda: dist.all_reduce(x))
try:
    t = symm.empty(n, dtype=torch.float64, device="cuda")
    hdl = symm.rendezvous(t, dist.group.WORLD.group_name)
    t.copy_(x)
    ops = [o for o in dir(torch.ops.symm_mem)]
    out = torch.ops.symm_mem.one_shot_all_reduce(t, "sum", dist.group.WORLD.group_name)
    ref = x.clone(); dist.all_reduce(ref)
    ok = torch.allclose(out, ref, rtol=1e-12, atol=1e-12)
    t_symm = timeit(lambda: torch.ops.symm_mem.one_shot_all_reduce(t, "sum", dist.group.WORLD.group_name))
    if rank == 0: print(f"world {world}: nccl {t_nccl:.1f} us, symm one-shot {t_symm:.1f} us, equal {ok}")
except Exception as ex:
    if rank == 0: print(f"world {world}: nccl {t_nccl:.1f} us, symm unavailable: {type(ex).__name__}: {ex}")
dist.destroy_process_group()
