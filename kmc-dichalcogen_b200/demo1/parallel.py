"""Multi-GPU plumbing: one process per GPU (torch.distributed), replicas sharded by contiguous global
index, no data-path collective.  The only exchange is one all-reduce of the per-class event histogram
(and of timing scalars) at the end -- NCCL over NVLink on GPUs, gloo in the CPU tests.

The reference has no parallelism at all (one trajectory, one process); trajectories are independent,
so replica r's result depends only on (seed, r), never on how many GPUs share the batch.
"""
import os


def world():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def replica_range(n_total, world_size, rank):
    """Contiguous share [lo, hi) of n_total replicas for `rank`; sizes differ by at most one."""
    base, extra = divmod(int(n_total), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def init(backend=None, device=None):
    """Initialise torch.distributed from the torchrun environment (no-op for a single process)."""
    import torch.distributed as dist
    rank, size, local = world()
    if size > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        kw = {}
        if backend is None:
            import torch
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl" and device is not None:
            kw["device_id"] = device
        dist.init_process_group(backend, **kw)
    return rank, size, local


def init_quiet(backend=None, device=None):
    """init() + a first barrier with file descriptor 1 pointed at stderr: NCCL prints its version banner on
    stdout when the communicator comes up, and our stdout carries JSON."""
    import sys
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        out = init(backend=backend, device=device)
        barrier()
        if device is not None:
            import torch
            torch.cuda.synchronize(device)
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)
    return out


def reduce_histogram(hist, device=None):
    """Sum an int64 histogram (numpy) over all ranks; returns numpy on every rank."""
    import torch
    import torch.distributed as dist
    t = torch.as_tensor(hist, dtype=torch.int64)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        if device is not None:
            t = t.to(device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def gather_float64(values, device=None):
    """Concatenation, in rank order, of every rank's float64 vector (lengths may differ); numpy on every rank.
    Used for statistics that must not depend on how the replicas were sharded: the caller reduces the gathered
    per-replica values with an exactly rounded sum (math.fsum) instead of summing per-rank partial sums."""
    import numpy as np
    import torch
    import torch.distributed as dist
    v = np.ascontiguousarray(values, dtype=np.float64).ravel()
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return v
    world = dist.get_world_size()
    n = torch.tensor([v.size], dtype=torch.int64)
    if device is not None:
        n = n.to(device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(x.item()) for x in sizes]
    buf = torch.zeros(max(sizes), dtype=torch.float64)
    buf[:v.size] = torch.from_numpy(v)
    if device is not None:
        buf = buf.to(device)
    parts = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    return np.concatenate([pt.cpu().numpy()[:k] for pt, k in zip(parts, sizes)])


def sum_over_ranks(value, device=None):
    """Sum of a Python float over all ranks."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        if device is not None:
            t = t.to(device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.cpu().item())


def max_over_ranks(value, device=None):
    """Max of a Python float over all ranks (timings are reported as the slowest rank's)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        if device is not None:
            t = t.to(device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.cpu().item())


def barrier():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def shutdown():
    """Tear the process group down (quietens NCCL's exit warning); a no-op for a single process."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()
