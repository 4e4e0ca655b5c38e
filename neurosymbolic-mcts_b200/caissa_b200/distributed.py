"""Multi-GPU plumbing: one process per GPU, replicas only on the hot path.

Self-play games are independent, so every rank runs its own game pool and evaluator handle with
replicated weights and there is NO collective on the evaluation path.  The two
off-path exchanges are here, over torch.distributed (NCCL on NVLink for GPU tensors, gloo for the
CPU tests):

  gather_samples     finished-game training samples (5763-f32 records, src/training_data.rs:20-60)
                     from all ranks to every rank (all_gather of padded fixed-size blocks)
  broadcast_weights  a new generation's flat weight blob from rank 0 to all ranks
"""
import numpy as np
import torch
import torch.distributed as dist

RECORD_FLOATS = 5763


def _device():
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def shard_range(n_items, rank, world):
    """Contiguous, balanced [lo, hi) of `n_items` independent units (games / positions) for `rank`."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


_PINNED = {}


def _pinned(name, shape):
    """A reusable pinned host tensor (CUDA ranks): staging in pageable memory halves the rate of both copies."""
    t = _PINNED.get(name)
    n = int(np.prod(shape))
    if t is None or t.numel() < n:
        t = torch.empty(max(n, 1), dtype=torch.float32).pin_memory()
        _PINNED[name] = t
    return t[:n].view(*shape)


def gather_samples(local, dst=None, transport="all_gather"):
    """local: float32 [n_local, 5763] (n_local may differ per rank, may be 0) -> float32 [sum n, 5763] in rank-major
    order.  dst=None: identical result on every rank.  dst=r: only rank r (the trainer's rank; the reference writes one file
    per generation, src/training_data.rs:24-60) copies the gathered records to the host, the others return an empty array.
    transport: "all_gather" (default) is ONE NCCL collective kernel; "gather" is NCCL's send / receive form.  Measured beside
    a two-pool self-play run (tools/exchange_probe.py, 2 x B200, 25 flushes/s of 512 records per rank): all_gather costs the
    pool 0.3 % of its sims/s, gather 13 % — the send / receive kernels hold SMs while they wait for their peer."""
    local = np.ascontiguousarray(local, dtype=np.float32).reshape(-1, RECORD_FLOATS)
    dev = _device()
    cuda = dev.type == "cuda"
    world, rank = dist.get_world_size(), dist.get_rank()
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    counts[rank] = local.shape[0]
    dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    counts = [int(x) for x in counts.cpu().tolist()]
    n_max, total = max(counts), sum(counts)
    if n_max == 0:
        return np.zeros((0, RECORD_FLOATS), dtype=np.float32)
    block = torch.zeros((n_max, RECORD_FLOATS), dtype=torch.float32, device=dev)
    if local.shape[0]:
        if cuda:
            stage = _pinned("in", local.shape)
            stage.copy_(torch.from_numpy(local))
            block[:local.shape[0]].copy_(stage, non_blocking=True)
        else:
            block[:local.shape[0]] = torch.from_numpy(local)
    if dst is None or transport == "all_gather":
        parts = torch.empty((world, n_max, RECORD_FLOATS), dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(parts, block) if cuda else dist.all_gather(list(parts.unbind(0)), block)
        parts = list(parts.unbind(0))
    else:
        parts = [torch.empty_like(block) for _ in range(world)] if rank == dst else None
        dist.gather(block, parts, dst=dst)
    if dst is not None and rank != dst:
        if cuda:
            torch.cuda.current_stream().synchronize()
        return np.zeros((0, RECORD_FLOATS), dtype=np.float32)
    out = _pinned("out", (total, RECORD_FLOATS)) if cuda else torch.empty((total, RECORD_FLOATS), dtype=torch.float32)
    lo = 0
    for r in range(world):
        if counts[r]:
            out[lo:lo + counts[r]].copy_(parts[r][:counts[r]], non_blocking=cuda)
        lo += counts[r]
    if cuda:
        torch.cuda.current_stream().synchronize()
    return out.numpy().copy() if cuda else out.numpy()


def broadcast_weights(blob, n_floats):
    """Rank 0 passes the blob, the others pass None; every rank returns the same float32 [n_floats]."""
    dev = _device()
    t = torch.empty(n_floats, dtype=torch.float32, device=dev)
    if dist.get_rank() == 0:
        t.copy_(torch.from_numpy(np.ascontiguousarray(blob, dtype=np.float32)))
    dist.broadcast(t, src=0)
    return t.cpu().numpy()
