"""One launch of the stand-alone gate kernel (cb_debug_device_tier1_gates) on 2500 playout positions, for ncu."""
import os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "neurosymbolic-mcts_b200")]
import caissa_b200 as cb
from oracle import pyoracle as O, weights as W
ev = cb.Evaluator(2, 128, "bf16", device=0)
ev.load_weights(W.flatten(W.make_weights(2, 128, 1405, "B"), 2, 128))
a, _, _, _ = O.random_positions(2718, 2500, 200)
v, d = ev.debug_device_tier1_gates(a, koth=False)
print("hits", int((v == 1).sum()))
