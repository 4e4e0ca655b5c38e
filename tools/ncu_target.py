"""Small target for ncu: a few B=1024 leaf evaluations (6x128, bf16) through the C ABI."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "neurosymbolic-mcts_b200"))
import caissa_b200 as cb  # noqa: E402
from oracle import pyoracle, weights as W  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 4
blocks, hidden = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (6, 128)
dtype = sys.argv[5] if len(sys.argv) > 5 else "bf16"
boards, idx, off, _ = pyoracle.random_positions(20261019, B)
q = pyoracle.static_q(boards)
ev = cb.Evaluator(blocks, hidden, dtype)
ev.load_weights(W.flatten(W.make_weights(blocks, hidden, 1803, "B"), blocks, hidden))
for _ in range(iters):
    out = ev.eval_leaves(boards, q, idx, off)
print("ok", float(out[2].mean()))
