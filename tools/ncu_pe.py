"""One launch of the stand-alone exchange-line kernel on 1024 self-play-like positions (for ncu)."""
import os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "neurosymbolic-mcts_b200"))
import caissa_b200 as cb
from oracle import weights as W, pyoracle as O
ev = cb.Evaluator(2, 128, "bf16")
ev.load_weights(W.flatten(W.make_weights(2, 128, 1803, "B"), 2, 128))
boards, _, _, _ = O.random_positions(1, 1024, 200)
for _ in range(3):
    q = ev.principal_exchange(boards)
print(q[:8])
