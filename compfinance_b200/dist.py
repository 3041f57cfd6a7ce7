"""Multi-GPU plumbing: how the path range is sharded over ranks and how per-rank results are combined.

The hot path shards naturally: paths are independent and both RNGs are random-access by
path index, so rank r of G simply runs paths [first, first + count) of the same Sobol / mrg32k3a stream.
The only exchange is one sum-all-reduce of the result vector [payoff sums, aggregate, adjoints]
(NCCL over NVLink on GPUs; gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(n_paths: int, rank: int, world: int, antithetic: bool = False, align: int = 256):
    """Contiguous shard [first, first + count) of rank `rank` (the engine's own rule: cf_shard_range, cf_api.cu).

    Shard boundaries are multiples of `align` paths (one thread block = 256 consecutive Sobol indices, so
    every block of every rank stays on a single pair of Sobol bases) and, for mrg32k3a, even, so that an
    antithetic pair (2q, 2q+1) is never split across ranks (the reference guarantees the same with its
    64-path batches, mcBase.h:312).  The last rank takes the remainder."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank / world")
    step = max(align, 2)
    if step % 2:
        step *= 2
    per = (n_paths // world) // step * step
    if per == 0:                       # fewer paths than ranks * align: fall back to pair granularity
        step = 2
        per = (n_paths // world) // step * step
    first = rank * per
    count = per if rank < world - 1 else n_paths - per * (world - 1)
    return first, count


def allreduce_sum_(t: torch.Tensor):
    """In-place sum over ranks of the result vector (no-op for a single process)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def connect_comm(eng, capacity: int):
    """Set up the engine's communicator over the processes of the torch.distributed job (include/cf_b200.h:
    cf_comm_create / cf_comm_connect): every rank allocates its receive block, the CUDA IPC handles are gathered in rank
    order with one all_gather_object (plumbing: any transport would do), every rank maps its peers' blocks.  From then on
    the sum over ranks is part of the engine's reduction kernels (peer memory over NVLink); no collective is called
    per step.  Raises when the blocks cannot be shared (the caller falls back to one NCCL all-reduce per step)."""
    world, rank = dist.get_world_size(), dist.get_rank()
    handle = eng.comm_create(world, rank, int(capacity))
    handles = [None] * world
    dist.all_gather_object(handles, handle)
    eng.comm_connect(handles)
    dist.barrier()
