"""Resident forward time of the 6x128 net at small batches (median of 30 replays, no L2 flush / with flush)."""
import os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "neurosymbolic-mcts_b200"))
import caissa_b200 as cb
from oracle import pyoracle, weights as W
boards, idx, off, _ = pyoracle.random_positions(20261019, 1024)
q = pyoracle.static_q(boards)
ev = cb.Evaluator(6, 128, sys.argv[1] if len(sys.argv) > 1 else "bf16")
ev.load_weights(W.flatten(W.make_weights(6, 128, 1803, "B"), 6, 128))
out = []
for B in (1, 4, 16, 64, 128, 148, 256, 296, 444, 592, 1024):
    ev.stage_leaves(boards[:B], q[:B], idx[:off[B]], off[:B + 1])
    ev.time_resident(5, flush_l2=False)
    a = float(np.median(ev.time_resident(30, flush_l2=False))) * 1e3
    b = float(np.median(ev.time_resident(30, flush_l2=True))) * 1e3
    out.append("B=%4d: %6.1f us back to back, %6.1f us flushed (%.0f evals/s)" % (B, a, b, B / b * 1e6))
print("\n".join(out))
