"""Power draw and SM clock (NVML, 20 ms samples) while the resident forward graph replays back to back."""
import os, sys, threading, time
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "neurosymbolic-mcts_b200"))
import caissa_b200 as cb
import pynvml as nv
from oracle import pyoracle, weights as W
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
boards, idx, off, _ = pyoracle.random_positions(20261019, B)
q = pyoracle.static_q(boards)
ev = cb.Evaluator(6, 128, "bf16")
ev.load_weights(W.flatten(W.make_weights(6, 128, 1803, "B"), 6, 128))
ev.stage_leaves(boards, q, idx, off)
nv.nvmlInit()
h = nv.nvmlDeviceGetHandleByIndex(0)
rows, stop = [], False
def sample():
    while not stop:
        rows.append((time.perf_counter(), nv.nvmlDeviceGetPowerUsage(h) / 1e3, nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM),
                     nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)))
        time.sleep(0.02)
t = threading.Thread(target=sample, daemon=True); t.start()
time.sleep(0.3)
t0 = time.perf_counter()
ms = []
while time.perf_counter() - t0 < 4.0:
    ms.append(ev.time_resident(500, flush_l2=False))
t1 = time.perf_counter()
stop = True; t.join()
ms = np.concatenate(ms)
busy = [r for r in rows if t0 + 1.0 <= r[0] <= t1]
print("batch %d: %d replays, %.1f us each (first 500: %.1f, last 500: %.1f)" % (B, len(ms), ms.mean() * 1e3, ms[:500].mean() * 1e3, ms[-500:].mean() * 1e3))
print("under load: power %.0f W (max %.0f), SM clock reported %.0f MHz, throttle reason bits 0x%x; limit %.0f W" % (
    np.mean([r[1] for r in busy]), max(r[1] for r in busy), np.mean([r[2] for r in busy]),
    int(np.bitwise_or.reduce([r[3] for r in busy])), nv.nvmlDeviceGetPowerManagementLimit(h) / 1e3))
