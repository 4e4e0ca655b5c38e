"""What does the NCCL sample gather cost the self-play pool?  torchrun --nproc-per-node N tools/exchange_probe.py
For each (pipeline, flushes per second, records per flush): sims/s of a pool with a side thread gathering records to rank 0."""
import os
import sys
import threading
import time

import numpy as np
import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "neurosymbolic-mcts_b200"))
import caissa_b200 as cb  # noqa: E402
from caissa_b200 import distributed as D  # noqa: E402
from oracle import weights as W  # noqa: E402

rank, local_rank, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local_rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
ev = cb.Evaluator(6, 128, "bf16", device=local_rank)
ev.load_weights(W.flatten(W.make_weights(6, 128, 1803, "B"), 6, 128))
T = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
records = np.random.default_rng(rank).random((16384, D.RECORD_FLOATS), dtype=np.float32)


def run(pipeline, games, rate, chunk, dst):
    n_flush = int(T * rate)
    done = [0]

    def loop():
        torch.cuda.set_device(local_rank)
        for _ in range(n_flush):
            D.gather_samples(records[:chunk], dst=dst)
            done[0] += 1
            time.sleep(max(0.0, 1.0 / rate - 0.002))
    dist.barrier()
    th = threading.Thread(target=loop) if n_flush else None
    if th:
        th.start()
    st = ev.selfplay(games=games, sims_per_move=200, seed=1 + rank, max_seconds=T, pipeline=pipeline)
    if th:
        th.join()
    t = torch.tensor([st["sims"] / st["seconds"]], dtype=torch.float64, device="cuda")
    dist.all_reduce(t)
    return t.item()


D.gather_samples(records[:8], dst=0)
FAST = os.environ.get("PROBE_FAST") == "1"
for pipeline, games in (((4, 2048),) if FAST else ((3, 1024), (4, 2048))):
    base = run(pipeline, games, 0, 0, 0)
    for rate, chunk in (((25, 64), (100, 128), (25, 512)) if FAST else ((25, 512), (5, 2560), (1, 12800))):
        for dst in ((None,) if FAST else (0, None)):
            v = run(pipeline, games, rate, chunk, dst)
            if rank == 0:
                print("NCCL_MAX_CTAS=%s pipeline %d: %2d flushes/s x %5d records/rank (%s): %.2f M sims/s (%.1f %% of %.2f M)" % (
                    os.environ.get("NCCL_MAX_CTAS", "-"), pipeline, rate, chunk, "gather to rank 0" if dst == 0 else "all-gather", v / 1e6, 100 * v / base, base / 1e6), flush=True)
dist.destroy_process_group()
