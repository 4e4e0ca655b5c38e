"""CB_DTYPE_FP32 on the tensor cores (fp16-pair operands): error against the golden fixture and resident rate."""
import os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "neurosymbolic-mcts_b200"))
import caissa_b200 as cb
from oracle import pyoracle, weights as W
for name in ("6x128_B", "2x128_B", "6x128_A"):
    g = np.load(os.path.join(REPO, "tests", "golden", "net_%s.npz" % name))
    nb, hid = int(g["blocks"]), int(g["hidden"])
    blob = W.flatten(W.make_weights(nb, hid, int(g["weight_seed"]), str(g["variant"])), nb, hid)
    ev = cb.Evaluator(nb, hid, "fp32")
    ev.load_weights(blob)
    n = g["boards"].shape[0]
    pol, v, k = ev.predict_batch(g["boards"], g["q"], np.ones(n, dtype=np.uint8))
    nf = g["log_policy_full"].shape[0]
    p = np.exp(g["log_policy_full"].astype(np.float64))
    klv = (p * (g["log_policy_full"] - np.log(np.maximum(pol[:nf], 1e-30)))).sum(1).max()
    bb = ev.debug_backbone(4)
    print("%s: |dv_logit| %.3e  KL %.3e  backbone %.3e  (launches per forward: %s)" %
          (name, np.abs(v - g["v_logit"]).max(), klv, np.abs(bb - g["backbone"]).max(), ev.launches_per_forward() if hasattr(ev, "launches_per_forward") else "?"))
boards, idx, off, _ = pyoracle.random_positions(20261019, 1024)
q = pyoracle.static_q(boards)
ev = cb.Evaluator(6, 128, "fp32")
ev.load_weights(W.flatten(W.make_weights(6, 128, 1803, "B"), 6, 128))
for B in (16, 148, 592, 1024):
    ev.stage_leaves(boards[:B], q[:B], idx[:off[B]], off[:B + 1])
    ev.time_resident(3, flush_l2=False)
    ms = float(np.median(ev.time_resident(20, flush_l2=True)))
    print("B=%4d: %.1f us per forward, %.0f evals/s" % (B, ms * 1e3, B / ms * 1e3))
