"""Where does the small-batch time of the resident 128-channel kernel go?  Times the staged forward at a few batch sizes
with the weight stream on / off (CB_DBG_R=16: the producer arrives on w_full without loading) and with the MMAs off
(CB_DBG_R=1), deep ring on / off.  Debug switches select the DBG instantiation (4-stage ring)."""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "neurosymbolic-mcts_b200"))
import caissa_b200 as cb  # noqa: E402
from oracle import pyoracle, weights as W  # noqa: E402

boards, idx, off, _ = pyoracle.random_positions(20261019, 1024)
q = pyoracle.static_q(boards)
ev = cb.Evaluator(6, 128, "bf16")
ev.load_weights(W.flatten(W.make_weights(6, 128, 1803, "B"), 6, 128))
for B in (1, 4, 16, 128, 256, 296, 592):
    ev.stage_leaves(boards[:B], q[:B], idx[:off[B]], off[:B + 1])
    row = []
    for name, env in (("default", {}), ("no deep ring", {"CB_NO_DEEP_RING": "1"}), ("dbg inst.", {"CB_DBG_R": "64"}),
                      ("no weight stream", {"CB_DBG_R": "16"}), ("no MMA", {"CB_DBG_R": "1"}), ("no MMA, no weights", {"CB_DBG_R": "17"}),
                      ("no epilogue main pass", {"CB_DBG_R": "32"})):
        for k in ("CB_NO_DEEP_RING", "CB_DBG_R"):
            os.environ.pop(k, None)
        os.environ.update(env)
        ev.time_resident(3, flush_l2=False)
        ms = ev.time_resident(30, flush_l2=False)
        row.append("%s %.1f us" % (name, float(np.median(ms)) * 1e3))
    print("B=%d: %s" % (B, " | ".join(row)), flush=True)
