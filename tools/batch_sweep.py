"""Config 2 of BASELINE.json: leaf-eval batch sweep 1..4096 on one B200 (6x128, bf16) with the
reference's traced CPU module beside it on a bounded sample.  Prints one JSON document."""
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "neurosymbolic-mcts_b200"))
import caissa_b200 as cb  # noqa: E402
from oracle import pyoracle, weights as W  # noqa: E402

blocks, hidden = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (6, 128)
dtype = sys.argv[3] if len(sys.argv) > 3 else "bf16"
boards, idx, off, _ = pyoracle.random_positions(20261019, 4736)
q = pyoracle.static_q(boards)
blob = W.flatten(W.make_weights(blocks, hidden, 1803, "B"), blocks, hidden)
ev = cb.Evaluator(blocks, hidden, dtype)
ev.load_weights(blob)
rows = []
cpu = None
try:
    import torch
    path = os.path.join(REPO, "oracle", "_ref", "oraclenet_%dx%d_B.pt" % (blocks, hidden))
    if os.path.exists(path):
        torch.set_num_threads(os.cpu_count())
        cpu = torch.jit.load(path, map_location="cpu").eval()
except Exception:
    cpu = None
for B in [1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 1184, 2048, 2368, 4096, 4736]:
    b, qq, o = boards[:B], q[:B], off[:B + 1]
    ii = idx[:o[B]]
    ev.stage_leaves(b, qq, ii, o)
    ev.time_resident(5)
    ms = ev.time_resident(30, flush_l2=True)
    for _ in range(3):
        ev.eval_leaves(b, qq, ii, o)
    t0 = time.perf_counter()
    n = 20
    for _ in range(n):
        ev.eval_leaves(b, qq, ii, o)
    e2e = (time.perf_counter() - t0) / n
    row = {"batch": B, "resident_us": float(np.median(ms) * 1e3), "resident_evals_per_s": B / float(np.median(ms) * 1e-3),
           "e2e_sync_us": e2e * 1e6, "e2e_sync_evals_per_s": B / e2e}
    if cpu is not None and B in (1, 16, 64, 256, 1024):
        planes = torch.from_numpy(pyoracle.planes_batch(b))
        scal = torch.from_numpy(np.stack([qq, np.ones(B, dtype=np.float32)], axis=1))
        with torch.no_grad():
            cpu(planes, scal)
            t0 = time.perf_counter()
            reps = 3 if B >= 256 else 10
            for _ in range(reps):
                cpu(planes, scal)
        row["cpu_reference_evals_per_s"] = B * reps / (time.perf_counter() - t0)
    rows.append(row)
print(json.dumps({"net": "%dx%d" % (blocks, hidden), "dtype": dtype, "cpu_cores": os.cpu_count(),
                  "l2": "flushed between resident steps", "rows": rows}))
