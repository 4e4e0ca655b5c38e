"""Diagnostic: cost per position entered of the stand-alone gate kernel (cb_debug_device_tier1_gates)."""
import os, sys, time
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "neurosymbolic-mcts_b200"), os.path.join(REPO, "tests")]
import caissa_b200 as cb
from oracle import pyoracle as O, weights as W

blob = W.flatten(W.make_weights(2, 128, 1405, "B"), 2, 128)
ev = cb.Evaluator(2, 128, "bf16", device=0)
ev.load_weights(blob)
a, _, _, _ = O.random_positions(2718, 2500, 200)
mn = np.array([O.mate_search(O.array_to_board(r), 5, 0)[2] for r in a])
hn = np.array([O.koth_center_in_n(O.array_to_board(r), 3)[1] for r in a])
for koth, nodes in ((False, mn), (True, mn + hn)):
    ev.debug_device_tier1_gates(a[:64], koth=koth)
    for rep in range(2):
        t0 = time.perf_counter()
        ev.debug_device_tier1_gates(a, koth=koth)
        dt = time.perf_counter() - t0
    # sorted by cost, to see the tail: time of the 148 * 4 cheapest-only subset
    order = np.argsort(nodes)
    cheap = a[order[:2000]]
    t0 = time.perf_counter(); ev.debug_device_tier1_gates(cheap, koth=koth); dtc = time.perf_counter() - t0
    print("koth", koth, "all 2500: %.2f ms, total nodes (approx) %d, max %d -> %.2f us per node of the longest job; cheapest 2000: %.2f ms, nodes %d max %d"
          % (dt * 1e3, nodes.sum(), nodes.max(), dt * 1e6 / nodes.max(), dtc * 1e3, nodes[order[:2000]].sum(), nodes[order[:2000]].max()), flush=True)
