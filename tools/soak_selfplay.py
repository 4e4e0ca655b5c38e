"""Soak: device self-play at full size with the sample file on, then validate every record written."""
import os, sys, tempfile
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "neurosymbolic-mcts_b200"))
import caissa_b200 as cb
from oracle import weights as W
ev = cb.Evaluator(6, 128, "bf16")
ev.load_weights(W.flatten(W.make_weights(6, 128, 1803, "B"), 6, 128))
CONFIGS = ((4, 2048, dict(material_on=True, qsearch="pe", tier1_light=True)), (3, 1024, dict(koth=True, tier1_light=True)), (4, 2048, dict()))
if len(sys.argv) > 2 and sys.argv[2] == "deep":   # the gate at full depth / dM by cap1: jobs suspended and resumed for the whole run
    CONFIGS = ((3, 2048, dict(tier1_light=2, material_on=True, qsearch="pe")), (3, 1024, dict(tier1_light=2, koth=True)),
               (3, 2048, dict(tier1_light=True, material_on=True, qsearch="cap1")), (4, 2048, dict(tier1_light=2, material_on=True, qsearch="cap1")))
for pipeline, games, kw in CONFIGS:
    path = os.path.join(tempfile.mkdtemp(), "s.bin")
    st = ev.selfplay(games=games, sims_per_move=64, max_plies=120, seed=5, max_seconds=float(sys.argv[1]) if len(sys.argv) > 1 else 8.0,
                     pipeline=pipeline, sample_path=path, **kw)
    rec = np.fromfile(path, dtype=np.float32).reshape(-1, 5763)
    planes, mat, flag, z, pol = rec[:, :1088], rec[:, 1088], rec[:, 1089], rec[:, 1090], rec[:, 1091:]
    ok = (rec.shape[0] == st["samples"] and np.isin(planes, (0.0, 1.0)).all() and np.isin(z, (-1.0, 0.0, 1.0)).all()
          and np.allclose(pol.sum(1), 1.0, atol=1e-4) and (pol >= 0).all() and np.isfinite(mat).all() and np.isin(flag, (0.0, 1.0)).all()
          and st["overflow"] == 0
          and st["games_finished"] == st["white_wins"] + st["black_wins"] + st["draws"])
    print("pipeline %d games %d %s: %.2fM sims/s, %d games finished (W %d B %d D %d), %d samples, terminal hits %d, gate hits %d, "
          "job waits %d, incomplete dM searches %d, records valid: %s" % (
        pipeline, games, kw, st["sims"] / st["seconds"] / 1e6, st["games_finished"], st["white_wins"], st["black_wins"], st["draws"],
        st["samples"], st["terminal_hits"], st["gate_hits"], st["gate_waits"], int((flag == 0.0).sum()), ok), flush=True)
    assert ok
    os.remove(path)
