"""Regenerates tests/golden/selfplay_samples.bin on a GPU box: the first 12 training samples (5763-f32 records,
src/training_data.rs:24-60) of a short device self-play run with dM from the principal-exchange line.

    python tests/golden/make_selfplay_fixture.py          # needs a B200 and the built library
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "neurosymbolic-mcts_b200"))
import caissa_b200 as cb  # noqa: E402
from oracle import weights as W  # noqa: E402

ev = cb.Evaluator(2, 128, "bf16")
ev.load_weights(W.flatten(W.make_weights(2, 128, 1803, "B"), 2, 128))
path = os.path.join(tempfile.mkdtemp(), "s.bin")
st = ev.selfplay(games=8, sims_per_move=24, max_plies=30, material_on=True, qsearch="pe", seed=7, max_games=8, pipeline=3,
                 sample_path=path)
rec = np.fromfile(path, dtype=np.float32).reshape(-1, 5763)
assert rec.shape[0] == st["samples"] >= 12
# a spread over the games' plies: every (n // 12)-th record
pick = rec[:: rec.shape[0] // 12][:12]
pick.tofile(os.path.join(HERE, "selfplay_samples.bin"))
print("wrote", pick.shape, "of", rec.shape[0], "samples;", st["games_finished"], "games finished")
