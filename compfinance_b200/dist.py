"""Multi-GPU plumbing: how the path range is sharded over ranks and how per-rank results are combined.

The hot path shards naturally: paths are independent and both RNGs are random-access by
path index, so rank r of G simply runs paths [first, first + count) of the same Sobol / mrg32k3a stream.
The only exchange is one sum-all-reduce of the result vector [payoff sums, aggregate, adjoints]
(NCCL over NVLink on GPUs; gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(n_paths: int, rank: int, world: int, antithetic: bool = False, align: int = 256):
    """Contiguous shard [first, first + count) of rank `rank`.

    Shard boundaries are multiples of `align` paths (one thread block = 256 consecutive Sobol indices, so
    every block of every rank stays on a single pair of Sobol bases) and, for mrg32k3a, even, so that an
    antithetic pair (2q, 2q+1) is never split across ranks (the reference guarantees the same with its
    64-path batches, mcBase.h:312).  The last rank takes the remainder."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank / world")
    step = max(align, 2 if antithetic else 1)
    if antithetic and step % 2:
        step *= 2
    per = (n_paths // world) // step * step
    if per == 0:                       # fewer paths than ranks * align: fall back to pair granularity
        step = 2 if antithetic else 1
        per = (n_paths // world) // step * step
    first = rank * per
    count = per if rank < world - 1 else n_paths - per * (world - 1)
    return first, count


def allreduce_sum_(t: torch.Tensor):
    """In-place sum over ranks of the result vector (no-op for a single process)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


class PeerSum:
    """Exchange buffers of the fused reduce + all-reduce kernel (include/cf_b200.h: cf_plan_set_peers).

    One symmetric allocation per rank -- [2][world][n_out] doubles (two alternating epochs, one row per sender), then `world` uint32 flag
    words -- mapped into every process of the group by torch's symmetric memory (CUDA IPC / fabric handles over
    NVLink).  The kernel does the rest: no collective is called per step.  Raises when symmetric memory is not
    available (the caller falls back to one NCCL all-reduce per step)."""

    def __init__(self, n_out: int):
        import torch.distributed._symmetric_memory as symm
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        n_flag = (4 * self.world + 7) // 8 + 1
        self.t = symm.empty(2 * self.world * n_out + n_flag, dtype=torch.float64, device="cuda")
        self.t.zero_()
        torch.cuda.synchronize()
        self.handle = symm.rendezvous(self.t, dist.group.WORLD.group_name)
        self.bufs = [int(p) for p in self.handle.buffer_ptrs]
        self.flags = [p + 2 * self.world * n_out * 8 for p in self.bufs]
        dist.barrier()
        torch.cuda.synchronize()

    def attach(self, lib, plan):
        import ctypes as C
        b = (C.c_void_p * self.world)(*self.bufs)
        f = (C.c_void_p * self.world)(*self.flags)
        rc = lib.cf_plan_set_peers(plan, self.world, self.rank, b, f)
        if rc != 0:
            raise RuntimeError(lib.cf_last_error().decode())

    def detach(self, lib, plan):
        lib.cf_plan_set_peers(plan, 0, 0, None, None)
