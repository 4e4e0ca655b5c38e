"""Self-play rates of the tree-side configurations on one B200 (A/B of changes to the chess primitives).

    python tools/tree_rate_probe.py [seconds per run]"""
import json, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "neurosymbolic-mcts_b200")]
import caissa_b200 as cb
from oracle import weights as W
secs = float(sys.argv[1]) if len(sys.argv) > 1 else 6.0
os.environ["CB_MCTS_TIMES"] = "1"
ev = cb.Evaluator(6, 128, "bf16", device=0)
ev.load_weights(W.flatten(W.make_weights(6, 128, 1405, "B"), 6, 128))
rows = []
for label, kw in (("vanilla, one pool, 1024 games", dict(pipeline=3, games=1024)),
                  ("light gates + dM by the exchange line, two pools, 2048 games", dict(pipeline=4, games=2048, tier1_light=True, material_on=True, qsearch="pe")),
                  ("deep gates, one pool, 1024 games", dict(pipeline=3, games=1024, tier1_light="deep")),
                  ("deep gates, one pool, 4096 games", dict(pipeline=3, games=4096, tier1_light="deep")),
                  ("light gates + dM by cap1, one pool, 1024 games", dict(pipeline=3, games=1024, tier1_light=True, material_on=True, qsearch="cap1"))):
    st = ev.selfplay(sims_per_move=200, max_seconds=secs, seed=5, **kw)
    row = dict(config=label, sims_per_s=round(st["sims"] / st["seconds"]), steps_per_s=round(st["steps"] / st["seconds"]))
    rows.append(row)
    print(json.dumps(row), flush=True)
json.dump(rows, open(os.path.join(REPO, "gpurun_out", "tree_rate_probe.json"), "w"), indent=1)
