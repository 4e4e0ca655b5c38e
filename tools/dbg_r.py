"""Debug driver for the resident tower: one forward at batch B under CB_DBG_R bits, prints errors instead of raising."""
import os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "neurosymbolic-mcts_b200"))
import caissa_b200 as cb
from oracle import pyoracle, weights as W
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
boards, idx, off, _ = pyoracle.random_positions(20261019, B)
q = pyoracle.static_q(boards)
ev = cb.Evaluator(2, 128, "bf16")
ev.load_weights(W.flatten(W.make_weights(2, 128, 1405, "B"), 2, 128))
try:
    pri, v, vf, k = ev.eval_leaves(boards, q, idx, off)
    print("CB_DBG_R=%s B=%d ok v[:4]=%s" % (os.environ.get("CB_DBG_R"), B, v[:4]))
except Exception as e:
    print("CB_DBG_R=%s B=%d FAILED: %s" % (os.environ.get("CB_DBG_R"), B, e))
