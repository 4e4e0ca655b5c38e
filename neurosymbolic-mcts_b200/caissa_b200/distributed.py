"""Multi-GPU plumbing: one process per GPU, replicas only on the hot path.

Self-play games are independent, so every rank runs its own game pool and evaluator handle with
replicated weights and there is NO collective on the evaluation path.  The two
off-path exchanges are here, over torch.distributed (NCCL on NVLink for GPU tensors, gloo for the
CPU tests):

  gather_samples     finished-game training samples (5763-f32 records, src/training_data.rs:20-60)
                     from all ranks to every rank (all_gather of padded fixed-size blocks)
  gather_samples_packed  the same exchange with the records in a lossless sparse wire form (~0.9 KB instead of 23 KB)
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


def pack_records(dense):
    """Lossless sparse form of float32 records [n, 5763] for the wire: a training record is 1088 plane entries that are 0 or
    1, three scalars and a dense 4672-wide policy with ~35 non-zero entries (src/training_data.rs:20-60), i.e. ~96 % zeros.
    Layout (little endian): n u32 | total u32 | count u32[n] | value bits u32[total] | column u16[total].  Entries are
    compared as BIT PATTERNS, so -0.0, NaN payloads and denormals survive; ~0.9 KB per record instead of 23,052 B."""
    dense = np.ascontiguousarray(dense, dtype=np.float32).reshape(-1, RECORD_FLOATS)
    bits = dense.view(np.uint32)
    rows, cols = np.nonzero(bits)
    n, total = dense.shape[0], rows.shape[0]
    out = np.empty(8 + 4 * n + 4 * total + 2 * total, dtype=np.uint8)
    out[:8].view(np.uint32)[:] = (n, total)
    out[8:8 + 4 * n].view(np.uint32)[:] = np.bincount(rows, minlength=n)
    o = 8 + 4 * n
    out[o:o + 4 * total].view(np.uint32)[:] = bits[rows, cols]
    out[o + 4 * total:].view(np.uint16)[:] = cols
    return out


def unpack_records(buf):
    """Inverse of pack_records: uint8 buffer (trailing padding allowed) -> float32 [n, 5763], bit for bit."""
    buf = np.ascontiguousarray(buf, dtype=np.uint8)
    n, total = (int(x) for x in buf[:8].view(np.uint32))
    counts = buf[8:8 + 4 * n].view(np.uint32).astype(np.int64)
    o = 8 + 4 * n
    vals = buf[o:o + 4 * total].view(np.uint32)
    cols = buf[o + 4 * total:o + 6 * total].view(np.uint16).astype(np.int64)
    dense = np.zeros((n, RECORD_FLOATS), dtype=np.float32)
    dense.view(np.uint32)[np.repeat(np.arange(n, dtype=np.int64), counts), cols] = vals
    return dense


def _pinned_u8(name, nbytes):
    t = _PINNED.get(name)
    if t is None or t.numel() < nbytes:
        t = torch.empty(max(nbytes, 1), dtype=torch.uint8).pin_memory()
        _PINNED[name] = t
    return t[:nbytes]


def gather_samples_packed(local, dst=None, timings=None, size_group=None):
    """gather_samples with the records in their sparse wire form (pack_records): the same result, bit for bit, with ~25 x
    fewer bytes through NCCL.  What the exchange costs a self-play pool that runs beside it is the time its collective
    kernels hold SMs, so the wire size is what matters (bench.py `nccl_exchange`, 8 x B200: dense records at the produced
    rate cost a two-pool run 24 % of its sims/s, packed ones see the `nccl_exchange` record of the last 8-GPU line).
    -> (float32 [sum n, 5763] on `dst` / every rank, empty elsewhere; bytes this rank put on the wire).  `timings`: a dict
    that receives pack_s / wire_s / unpack_s of this rank.  `size_group`: a gloo group of the same ranks; the block sizes
    are then agreed on the HOST, which also lines the ranks up before the one NCCL kernel of the flush is launched — a
    collective kernel that starts early spins on its SMs until the last rank arrives, and beside a running self-play pool
    that skew (milliseconds between free-running host threads) is what the exchange costs."""
    import time
    t0 = time.perf_counter()
    blob = pack_records(local)
    t1 = time.perf_counter()
    dev = _device()
    cuda = dev.type == "cuda"
    world, rank = dist.get_world_size(), dist.get_rank()
    if size_group is not None:
        sizes = torch.zeros(world, dtype=torch.int64)
        sizes[rank] = blob.shape[0]
        dist.all_reduce(sizes, op=dist.ReduceOp.SUM, group=size_group)
        sizes = [int(x) for x in sizes.tolist()]
    else:
        sizes = torch.zeros(world, dtype=torch.int64, device=dev)
        sizes[rank] = blob.shape[0]
        dist.all_reduce(sizes, op=dist.ReduceOp.SUM)
        sizes = [int(x) for x in sizes.cpu().tolist()]
    b_max = (max(sizes) + 15) & ~15
    block = torch.zeros(b_max, dtype=torch.uint8, device=dev)
    if cuda:
        stage = _pinned_u8("in8", blob.shape[0])
        stage.copy_(torch.from_numpy(blob))
        block[:blob.shape[0]].copy_(stage, non_blocking=True)
    else:
        block[:blob.shape[0]] = torch.from_numpy(blob)
    parts = torch.empty((world, b_max), dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(parts, block) if cuda else dist.all_gather(list(parts.unbind(0)), block)
    if dst is not None and rank != dst:
        if cuda:
            torch.cuda.current_stream().synchronize()
        if timings is not None:
            timings.update(pack_s=t1 - t0, wire_s=time.perf_counter() - t1, unpack_s=0.0)
        return np.zeros((0, RECORD_FLOATS), dtype=np.float32), blob.shape[0]
    if cuda:
        host = _pinned_u8("out8", world * b_max).view(world, b_max)
        host.copy_(parts, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        host = host.numpy()
    else:
        host = parts.numpy()
    t2 = time.perf_counter()
    out = [unpack_records(host[r, :sizes[r]]) for r in range(world)]
    res = np.concatenate(out, axis=0)
    if timings is not None:
        timings.update(pack_s=t1 - t0, wire_s=t2 - t1, unpack_s=time.perf_counter() - t2)
    return res, blob.shape[0]


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
