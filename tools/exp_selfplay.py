"""Self-play throughput under different pool shapes (games, pipeline mode, host threads)."""
import os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "neurosymbolic-mcts_b200"))
import caissa_b200 as cb
from oracle import weights as W
ev = cb.Evaluator(6, 128, "bf16")
ev.load_weights(W.flatten(W.make_weights(6, 128, 1803, "B"), 6, 128))
ev2 = cb.Evaluator(6, 128, "bf16")
ev2.load_weights(W.flatten(W.make_weights(6, 128, 1803, "B"), 6, 128))
ncpu = os.cpu_count()
for games, pipeline, threads, aux in [(1024, 1, ncpu - 2, None), (1024, 2, ncpu - 2, None), (1024, 2, ncpu - 2, ev2), (2048, 2, ncpu - 2, None),
                                      (2048, 2, ncpu - 2, ev2), (4096, 2, ncpu - 2, ev2), (1024, 2, 8, ev2), (2048, 2, 8, ev2), (2048, 2, 5, ev2)]:
    st = ev.selfplay(games=games, sims_per_move=200, material_on=False, seed=3, max_seconds=3.0, threads=threads, pipeline=pipeline, aux=aux)
    print("games %4d pipeline %d aux %d threads %2d : %.2fM sims/s, mean batch %.0f, eval %.2fs tree %.2fs of %.2fs, steps %d" % (
        games, pipeline, int(aux is not None), threads, st["sims"] / st["seconds"] / 1e6, st["mean_batch"], st["eval_seconds"], st["tree_seconds"],
        st["seconds"], st["steps"]))

for games in (1024, 2048, 4096):
    st = ev.selfplay(games=games, sims_per_move=200, material_on=False, seed=3, max_seconds=3.0, pipeline=3)
    print("DEVICE tree ops: games %4d : %.2fM sims/s, mean batch %.0f, forward %.2fs tree kernels %.2fs of %.2fs, steps %d, moves %d, games %d" % (
        games, st["sims"] / st["seconds"] / 1e6, st["mean_batch"], st["eval_seconds"], st["tree_seconds"], st["seconds"], st["steps"],
        st["moves"], st["games_finished"]))

cand, curr = ev, ev2
m = cand.match(curr, games=1024, total_games=1 << 30, sims_per_move=200, seed=3, max_seconds=3.0)
print("DEVICE match: 1024 concurrent games, two models per step: %.2fM sims/s, %d games finished (W %d / L %d / D %d)" % (
    m["sims"] / m["seconds"] / 1e6, m["games"], m["candidate_wins"], m["current_wins"], m["draws"]))
