"""ncu target: device-side self-play (1024 games, 200 sims/move) for a bounded number of simulations."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "neurosymbolic-mcts_b200"))
import caissa_b200 as cb
from oracle import weights as W
ev = cb.Evaluator(6, 128, "bf16")
ev.load_weights(W.flatten(W.make_weights(6, 128, 1803, "B"), 6, 128))
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 450
st = ev.selfplay(games=1024, sims_per_move=200, material_on=False, seed=3, max_sims=1024 * steps, pipeline=3)
print("ok", st["sims"], st["moves"])
