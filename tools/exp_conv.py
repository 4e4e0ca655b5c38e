"""Experiment: per-layer conv timing under CB_DBG switches (set in the environment)."""
import os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "neurosymbolic-mcts_b200"))
import caissa_b200 as cb
from oracle import pyoracle, weights as W
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
blocks, hidden = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (6, 128)
boards, idx, off, _ = pyoracle.random_positions(20261019, B)
q = pyoracle.static_q(boards)
ev = cb.Evaluator(blocks, hidden, "bf16")
ev.load_weights(W.flatten(W.make_weights(blocks, hidden, 1803, "B"), blocks, hidden))
os.environ.setdefault("X", "")
ev.stage_leaves(boards, q, idx, off)
ev.time_conv_layers_detail(3)
per = ev.time_conv_layers_detail(20) * 1e3
relu = [per[i] for i in range(1, len(per)) if i % 2 == 1]
se = [per[i] for i in range(1, len(per)) if i % 2 == 0]
ms = ev.time_resident(20, flush_l2=False)
tw, _ = ev.time_tower(10)
print("CB_DBG=%s B=%d net %dx%d conv0 %.1f us | relu layers %.1f us | se layers %.1f us | forward %.1f us | tower %.1f us" % (
    os.environ.get("CB_DBG", "0"), B, blocks, hidden, per[0], np.mean(relu), np.mean(se), ms[5:].mean() * 1e3, tw[3:].mean() * 1e3))
