"""Self-play rate with the reference's Tier-1 gate at full depth (tier1_light = 2) against the light gates, one B200.

    python tools/deep_gates_probe.py [seconds per run]

6x128 net (random init), 1024 games, 200 sims/move, tree operations on the device (pipeline 3 / 4).  The gate jobs are
suspended after CB_GATE_BUDGET positions per step; the sweep shows the trade between the length of a step and the steps a
game waits for its gate."""
import json, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "neurosymbolic-mcts_b200")]
import caissa_b200 as cb
from oracle import weights as W

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
blob = W.flatten(W.make_weights(6, 128, 1405, "B"), 6, 128)
ev = cb.Evaluator(6, 128, "bf16", device=0)
ev.load_weights(blob)
rows = []


def run(label, budget=None, **kw):
    if budget is not None:
        os.environ["CB_GATE_BUDGET"] = str(budget)
    st = ev.selfplay(games=kw.pop("games", 1024), sims_per_move=200, max_seconds=secs, seed=5, **kw)
    row = dict(config=label, budget=budget, sims_per_s=round(st["sims"] / st["seconds"]), steps_per_s=round(st["steps"] / st["seconds"]),
               evals_per_s=round(st["nn_evals"] / st["seconds"]), gate_hits=st["gate_hits"], gate_waits=st["gate_waits"],
               wait_fraction=round(st["gate_waits"] / max(1, st["steps"] * 1024), 4), games=st["games_finished"], moves=st["moves"])
    rows.append(row)
    print(json.dumps(row), flush=True)


run("vanilla, one pool", pipeline=3)
run("light gates, one pool", pipeline=3, tier1_light=True)
for b in (8, 24, 48, 96, 192):
    run("deep gates (mate-in-5 by checks), one pool", budget=b, pipeline=3, tier1_light="deep")
run("deep gates, two pools", budget=48, pipeline=4, games=2048, tier1_light="deep")
run("deep gates + dM by the exchange line, one pool", budget=48, pipeline=3, tier1_light="deep", material_on=True, qsearch="pe")
run("KOTH, light gates", pipeline=3, tier1_light=True, koth=True)
for b in (48, 192):
    run("KOTH, deep gates (hill in 3 + mate-in-5)", budget=b, pipeline=3, tier1_light="deep", koth=True)
with open(os.path.join(REPO, "gpurun_out", "deep_gates_probe.json"), "w") as f:
    json.dump(rows, f, indent=1)
