"""Two-pool device self-play (pipeline 4): pool size sweep."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "neurosymbolic-mcts_b200"))
import caissa_b200 as cb
from oracle import weights as W
ev = cb.Evaluator(6, 128, "bf16")
ev.load_weights(W.flatten(W.make_weights(6, 128, 1803, "B"), 6, 128))
for games in (2048, 2112, 2176, 2240, 2304, 1920, 1792):
    for material, qs in ((False, 0), (True, "pe")):
        st = ev.selfplay(games=games, sims_per_move=200, material_on=material, qsearch=qs, seed=3, max_seconds=2.5, pipeline=4)
        print("games 2 x %d (forward on %d SMs) material %d qsearch %-3s : %.2fM sims/s, %.1f us per pool step" % (
            games // 2, (games // 2 + 7) // 8, material, qs, st["sims"] / st["seconds"] / 1e6, st["seconds"] / st["steps"] / 2 * 1e6))
