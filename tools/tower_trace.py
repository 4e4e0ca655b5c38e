"""Timeline of CTA 0 of the fused transposed tower (debug): per work item, microseconds relative to the first stamp."""
import os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "neurosymbolic-mcts_b200"))
import caissa_b200 as cb
from oracle import pyoracle, weights as W
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
boards, idx, off, _ = pyoracle.random_positions(20261019, B)
q = pyoracle.static_q(boards)
ev = cb.Evaluator(6, 128, "bf16")
ev.load_weights(W.flatten(W.make_weights(6, 128, 1803, "B"), 6, 128))
ev.stage_leaves(boards, q, idx, off)
ev.time_resident(3)
tr = ev.debug_tower_trace().astype(np.float64)
t0 = tr[tr > 0].min()
us = (tr - t0) / 1965.0

names = ["prod_ready", "prod_done", "mma_start", "mma_end", "epi_full", "epi_pub", "st_issue", "st_done"]
print("item L c " + " ".join("%10s" % n for n in names) + "   mma_us  epi_us  store_us  gap_to_next_mma")
for i in range(us.shape[0]):
    nxt = us[i + 1, 2] - us[i, 3] if i + 1 < us.shape[0] else 0
    print("%3d %2d %d " % (i, i // 2, i % 2) + " ".join("%10.2f" % v for v in us[i]) +
          "   %6.2f  %6.2f  %6.2f  %6.2f" % (us[i, 3] - us[i, 2], us[i, 5] - us[i, 4], us[i, 7] - us[i, 6], nxt))

try:
    ev.time_tower(3)
    ct = ev.debug_cta_times().astype(np.float64)
    t0 = ct[:, 0].min()
    rel = (ct - t0) / 1e3
    print("CTA times (us, relative to the first entry): entry min/max %.2f/%.2f | first MMA min/max %.2f/%.2f | last epilogue min/median/max %.2f/%.2f/%.2f | exit max %.2f"
          % (rel[:, 0].min(), rel[:, 0].max(), rel[:, 1].min(), rel[:, 1].max(), rel[:, 2].min(), np.median(rel[:, 2]), rel[:, 2].max(), rel[:, 3].max()))
    dur = rel[:, 2] - rel[:, 1]
    print("per-CTA first MMA -> last epilogue: min %.2f median %.2f max %.2f us" % (dur.min(), np.median(dur), dur.max()))
    tw, _ = ev.time_tower(10)
    print("tower launch (CUDA events): %s us" % np.round(tw * 1e3, 1))
except Exception as e:
    print("cta times unavailable:", e)

# effective SM clock during the kernel: SM cycles of CTA 0 (first MMA stamp -> last epilogue stamp) over the
# global-timer nanoseconds of the same span in a second launch
try:
    span_cyc = float(tr[:, 5].max() - tr[tr[:, 2] > 0, 2].min())
    span_ns = float(ct[0, 2] - ct[0, 1])
    print("effective SM clock under this kernel: %.0f MHz (%.0f cycles / %.1f us)" % (span_cyc / span_ns * 1e3, span_cyc, span_ns / 1e3))
except Exception as e:
    print("clock estimate unavailable:", e)
