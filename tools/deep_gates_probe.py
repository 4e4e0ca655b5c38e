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
    st = ev.selfplay(games=kw.get("games", 1024), sims_per_move=200, max_seconds=secs, seed=5, **{k: v for k, v in kw.items() if k != 'games'})
    row = dict(config=label, budget=budget, sims_per_s=round(st["sims"] / st["seconds"]), steps_per_s=round(st["steps"] / st["seconds"]),
               evals_per_s=round(st["nn_evals"] / st["seconds"]), gate_hits=st["gate_hits"], gate_waits=st["gate_waits"],
               wait_fraction=round(st["gate_waits"] / max(1, st["steps"] * kw.get("games", 1024)), 4), games=st["games_finished"], moves=st["moves"])
    rows.append(row)
    print(json.dumps(row), flush=True)


if len(sys.argv) > 2 and sys.argv[2] == "cap1":
    run("light gates + dM by the exchange line, one pool, 1024 games", pipeline=3, games=1024, tier1_light=True, material_on=True, qsearch="pe")
    run("light gates + dM by cap1, one pool, 1024 games", pipeline=3, games=1024, tier1_light=True, material_on=True, qsearch="cap1")
    run("light gates + dM by cap1, one pool, 4096 games", pipeline=3, games=4096, tier1_light=True, material_on=True, qsearch="cap1")
    run("light gates + dM by cap1, two pools, 2048 games", pipeline=4, games=2048, tier1_light=True, material_on=True, qsearch="cap1")
    run("deep gates + dM by cap1, one pool, 4096 games", budget=24, pipeline=3, games=4096, tier1_light="deep", material_on=True, qsearch="cap1")
    with open(os.path.join(REPO, "gpurun_out", "cap1_probe.json"), "w") as f:
        json.dump(rows, f, indent=1)
    sys.exit(0)
if len(sys.argv) > 2 and sys.argv[2] == "warps":
    # warps per game in the gate kernel (each takes every W-th move of the gate's root)
    for W in ("1", "4", "8"):
        os.environ["CB_GATE_WARPS"] = W
        run("deep gates, 1024 games, %s warps per game" % W, budget=24, pipeline=3, games=1024, tier1_light="deep")
        run("deep gates, 4096 games, %s warps per game" % W, budget=24, pipeline=3, games=4096, tier1_light="deep")
    os.environ["CB_GATE_WARPS"] = "4"
    run("KOTH, deep gates, 1024 games, 4 warps per game", budget=24, pipeline=3, games=1024, tier1_light="deep", koth=True)
    os.environ["CB_GATE_WARPS"] = "8"
    run("KOTH, deep gates, 1024 games, 8 warps per game", budget=24, pipeline=3, games=1024, tier1_light="deep", koth=True)
    with open(os.path.join(REPO, "gpurun_out", "deep_gates_probe_warps.json"), "w") as f:
        json.dump(rows, f, indent=1)
    sys.exit(0)
if len(sys.argv) > 2 and sys.argv[2] == "stretch":
    # the adaptive stretch of the per-step budget (CB_GATE_STRETCH positions more while fewer than half of the games are done)
    for stretch in ("0", "400"):
        os.environ["CB_GATE_STRETCH"] = stretch
        run("deep gates, 1024 games, stretch %s" % stretch, budget=24, pipeline=3, games=1024, tier1_light="deep")
        run("deep gates, 4096 games, stretch %s" % stretch, budget=24, pipeline=3, games=4096, tier1_light="deep")
        run("light gates + dM by cap1, 1024 games, stretch %s" % stretch, budget=24, pipeline=3, games=1024, tier1_light=True, material_on=True, qsearch="cap1")
    with open(os.path.join(REPO, "gpurun_out", "deep_gates_probe_stretch.json"), "w") as f:
        json.dump(rows, f, indent=1)
    sys.exit(0)
if len(sys.argv) > 2 and sys.argv[2] == "long":
    # steady-state figures: runs long enough to leave the openings (a game is ~90 moves x 200 simulations)
    for G, b in ((1024, 12), (1024, 24), (4096, 12), (4096, 24), (8192, 24)):
        run("deep gates, one pool, %d games" % G, budget=b, pipeline=3, games=G, tier1_light="deep")
    run("deep gates + dM by the exchange line, one pool, 4096 games", budget=24, pipeline=3, games=4096, tier1_light="deep", material_on=True, qsearch="pe")
    run("KOTH, deep gates (hill in 3 + mate-in-5), 4096 games", budget=24, pipeline=3, games=4096, tier1_light="deep", koth=True)
    run("light gates, one pool, 1024 games", pipeline=3, games=1024, tier1_light=True)
    run("light gates + dM by the exchange line, one pool, 1024 games", pipeline=3, games=1024, tier1_light=True, material_on=True, qsearch="pe")
    run("light gates + dM by cap1, one pool, 1024 games", pipeline=3, games=1024, tier1_light=True, material_on=True, qsearch="cap1")
    run("light gates + dM by cap1, one pool, 4096 games", pipeline=3, games=4096, tier1_light=True, material_on=True, qsearch="cap1")
    run("deep gates + dM by cap1, one pool, 4096 games", budget=24, pipeline=3, games=4096, tier1_light="deep", material_on=True, qsearch="cap1")
    with open(os.path.join(REPO, "gpurun_out", "deep_gates_probe_long.json"), "w") as f:
        json.dump(rows, f, indent=1)
    sys.exit(0)
if len(sys.argv) > 2 and sys.argv[2] == "games":
    for G in (1024, 2048, 4096):
        for b in (24, 48):
            run("deep gates, one pool, %d games" % G, budget=b, pipeline=3, games=G, tier1_light="deep")
    run("deep gates, two pools, 4096 games", budget=48, pipeline=4, games=4096, tier1_light="deep")
    run("light gates, one pool, 4096 games", pipeline=3, games=4096, tier1_light=True)
    with open(os.path.join(REPO, "gpurun_out", "deep_gates_probe_games.json"), "w") as f:
        json.dump(rows, f, indent=1)
    sys.exit(0)
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
