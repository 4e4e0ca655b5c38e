"""N > 1 host logic on CPU: two gloo ranks exchange training samples and weights exactly the way
the NCCL ranks do on the GPU box (no collective exists on the evaluation path itself)."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, os.path.join(REPO, "neurosymbolic-mcts_b200"))
    import torch.distributed as dist
    from caissa_b200 import distributed as D
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(100 + rank)
    n_local = [3, 0, 5][rank % 3] if world > 1 else 3
    local = rng.random((n_local, D.RECORD_FLOATS), dtype=np.float32) + rank
    allrec = D.gather_samples(local)
    on0 = D.gather_samples(local, dst=0)
    assert np.array_equal(D.gather_samples(local, dst=0, transport="gather"), on0)
    # the sparse wire form: lossless on arbitrary bit patterns, and on records shaped like the real ones
    sparse = np.zeros((n_local + 2, D.RECORD_FLOATS), dtype=np.float32)
    sparse[:, :1088] = rng.random((n_local + 2, 1088)) < 0.05
    sparse[:, 1088:1091] = rng.normal(size=(n_local + 2, 3))
    for i in range(n_local + 2):
        sparse[i, 1091 + rng.choice(4672, size=30, replace=False)] = rng.random(30)
    sparse[0, 7] = -0.0
    sparse[-1, 5762] = np.float32(1e-42)                      # a denormal in the last column
    for x in (local, sparse, sparse[:0]):
        assert np.array_equal(D.unpack_records(D.pack_records(x)).view(np.uint32), np.ascontiguousarray(x).view(np.uint32))
    assert D.pack_records(sparse).shape[0] < sparse.nbytes // 20
    packed_all, _ = D.gather_samples_packed(sparse)
    packed_on0, wire = D.gather_samples_packed(sparse, dst=0, size_group=dist.new_group(backend="gloo"))
    assert np.array_equal(D.gather_samples(sparse).view(np.uint32), packed_all.view(np.uint32))
    assert packed_on0.shape[0] == (packed_all.shape[0] if rank == 0 else 0) and wire < sparse.nbytes // 20
    blob = D.broadcast_weights(np.arange(1000, dtype=np.float32) if rank == 0 else None, 1000)
    lo, hi = D.shard_range(1024 + 1, rank, world)
    np.savez(os.path.join(out_dir, "r%d.npz" % rank), local=local, allrec=allrec, on0=on0, blob=blob, lo=lo, hi=hi)
    dist.destroy_process_group()


def test_two_rank_gloo_sample_gather_and_weight_broadcast(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, 29611, str(tmp_path)), nprocs=world, join=True)
    r = [np.load(tmp_path / ("r%d.npz" % i)) for i in range(world)]
    want = np.concatenate([r[0]["local"], r[1]["local"]], axis=0)
    for i in range(world):
        assert np.array_equal(r[i]["allrec"], want)                 # rank-major, ragged counts (3 and 0)
        assert np.array_equal(r[i]["blob"], np.arange(1000, dtype=np.float32))
        assert np.array_equal(r[i]["on0"], want) if i == 0 else r[i]["on0"].shape[0] == 0   # gather to the trainer's rank
    # the shards of the independent units tile [0, n) exactly
    assert r[0]["lo"] == 0 and r[0]["hi"] == r[1]["lo"] and r[1]["hi"] == 1025
