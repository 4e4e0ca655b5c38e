"""Device self-play throughput by material mode: off, stand pat, principal exchange (cb_selfplay_config.qsearch);
and the stand-alone exchange-line kernel on a batch of positions."""
import os, sys, time
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "neurosymbolic-mcts_b200"))
import caissa_b200 as cb
from oracle import weights as W, pyoracle as O
ev = cb.Evaluator(6, 128, "bf16")
ev.load_weights(W.flatten(W.make_weights(6, 128, 1803, "B"), 6, 128))
for material, qs in [(False, 0), (True, 0), (True, "pe")]:
    st = ev.selfplay(games=1024, sims_per_move=200, material_on=material, qsearch=qs, seed=3, max_seconds=3.0, pipeline=3)
    print("material %d qsearch %-3s : %.2fM sims/s, forward %.2fs tree kernels %.2fs of %.2fs, steps %d, moves %d" % (
        material, qs, st["sims"] / st["seconds"] / 1e6, st["eval_seconds"], st["tree_seconds"], st["seconds"], st["steps"], st["moves"]))
for games, pipeline, material, qs in [(1024, 3, False, 0), (2048, 3, False, 0), (2048, 4, False, 0), (2048, 4, True, 0), (2048, 4, True, "pe"), (2368, 4, False, 0), (1024, 4, False, 0)]:
    st = ev.selfplay(games=games, sims_per_move=200, material_on=material, qsearch=qs, seed=3, max_seconds=3.0, pipeline=pipeline)
    print("games %d pipeline %d material %d qsearch %-3s : %.2fM sims/s, steps %d (%.1f us per pool step), moves %d" % (
        games, pipeline, material, qs, st["sims"] / st["seconds"] / 1e6, st["steps"], st["seconds"] / st["steps"] / (2 if pipeline == 4 else 1) * 1e6, st["moves"]))
boards, _, _, _ = O.random_positions(1, 4096, 200)
q = ev.principal_exchange(boards)
t = time.time()
for _ in range(20):
    q = ev.principal_exchange(boards)
dt = (time.time() - t) / 20
_, nodes = cb.principal_exchange(boards, want_nodes=True)
t = time.time(); cb.principal_exchange(boards); th = time.time() - t
print("exchange line, 4096 positions (mean %.2f positions visited, max %d): device call %.3f ms (host buffers), one host thread %.2f ms" % (
    nodes.mean(), nodes.max(), dt * 1e3, th * 1e3))
