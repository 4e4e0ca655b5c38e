"""Diagnostic: forward / tree split of a self-play step with the deep gates (CB_MCTS_TIMES)."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "neurosymbolic-mcts_b200")]
import caissa_b200 as cb
from oracle import weights as W
os.environ["CB_MCTS_TIMES"] = "1"
ev = cb.Evaluator(6, 128, "bf16", device=0)
ev.load_weights(W.flatten(W.make_weights(6, 128, 1405, "B"), 6, 128))
for label, kw in (("light", dict(tier1_light=True)), ("deep b1", dict(tier1_light="deep", budget=1)), ("deep b8", dict(tier1_light="deep", budget=8)),
                  ("deep b24", dict(tier1_light="deep", budget=24)), ("deep b48", dict(tier1_light="deep", budget=48)),
                  ("deep b24 inline", dict(tier1_light="deep", budget=24, inline=1))):
    b = kw.pop("budget", None)
    if b: os.environ["CB_GATE_BUDGET"] = str(b)
    if kw.pop("inline", 0): os.environ["CB_GATE_INLINE"] = "1"
    print(label, flush=True)
    sys.stderr.flush()
    st = ev.selfplay(games=1024, sims_per_move=200, max_seconds=2.0, seed=5, pipeline=3, **kw)
    print("   sims/s %.0f steps/s %.0f waits/step %.1f evals/step %.1f" % (st["sims"] / st["seconds"], st["steps"] / st["seconds"],
          st["gate_waits"] / st["steps"], st["nn_evals"] / st["steps"]), flush=True)
