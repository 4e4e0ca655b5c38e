"""Self-play rate with the tree operations on the device: one pool / two pools, vanilla / exchange-line dM / light gates."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "neurosymbolic-mcts_b200"))
import caissa_b200 as cb
if os.environ.get("CB_LIB"):      # A/B runs: another build of the library
    cb._LIB_PATH = os.environ["CB_LIB"]
from oracle import weights as W
ev = cb.Evaluator(6, 128, "bf16")
ev.load_weights(W.flatten(W.make_weights(6, 128, 1803, "B"), 6, 128))
T = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
for name, kw in (("one pool, vanilla", dict(games=1024, pipeline=3, material_on=False)),
                 ("one pool, dM = exchange line", dict(games=1024, pipeline=3, material_on=True, qsearch="pe")),
                 ("one pool, exchange line + light gates", dict(games=1024, pipeline=3, material_on=True, qsearch="pe", tier1_light=True)),
                 ("two pools, vanilla", dict(games=2048, pipeline=4, material_on=False)),
                 ("two pools, dM = exchange line", dict(games=2048, pipeline=4, material_on=True, qsearch="pe"))):
    st = ev.selfplay(sims_per_move=200, seed=3, max_seconds=T, **kw)
    print("%-40s %.3f M sims/s, forward %.2fs tree kernels %.2fs of %.2fs, steps %d" % (
        name, st["sims"] / st["seconds"] / 1e6, st["eval_seconds"], st["tree_seconds"], st["seconds"], st["steps"]), flush=True)
