"""GPU triage: compare the tensor-core tower with the exact-mode tower layer by layer,
then against the golden fixtures.  Usage: python tools/gpu_debug.py [cfg] [dtype] [B]"""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "neurosymbolic-mcts_b200"))
import caissa_b200 as cb  # noqa: E402
from oracle import pyoracle, weights as W  # noqa: E402


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "2x128_B"
    dtype = sys.argv[2] if len(sys.argv) > 2 else "bf16"
    g = np.load(os.path.join(REPO, "tests", "golden", "net_%s.npz" % cfg))
    blocks, hidden = int(g["blocks"]), int(g["hidden"])
    B = int(sys.argv[3]) if len(sys.argv) > 3 else min(8, g["boards"].shape[0])
    blob = W.flatten(W.make_weights(blocks, hidden, int(g["weight_seed"]), str(g["variant"])), blocks, hidden)
    boards, q = g["boards"][:B], g["q"][:B]
    print(cb.load_library().cb_version().decode(), "cfg", cfg, "dtype", dtype, "B", B, flush=True)

    ref = cb.Evaluator(blocks, hidden, "fp32")
    ref.load_weights(blob)
    planes = ref.encode(boards)
    print("encode bit-exact vs oracle:", np.array_equal(planes, pyoracle.planes_batch(boards)), flush=True)
    pol_r, v_r, k_r = ref.predict_batch(boards, q)
    nf = min(B, g["log_policy_full"].shape[0])
    print("fp32 mode vs golden: max|dv|=%.3e  max|dlogp|=%.3e  k=%.6f/%.6f" % (
        np.abs(v_r - g["v_logit"][:B]).max(),
        np.abs(np.log(np.maximum(pol_r[:nf], 1e-30)) - g["log_policy_full"][:nf]).max(), k_r[0], g["k"][0]), flush=True)
    bb = ref.debug_backbone(min(B, 4))
    print("fp32 backbone vs golden: max|d|=%.3e (ref max %.3f)" % (np.abs(bb - g["backbone"][:bb.shape[0]]).max(),
                                                                   np.abs(g["backbone"]).max()), flush=True)

    tc = cb.Evaluator(blocks, hidden, dtype)
    tc.load_weights(blob)
    nconv = 2 * blocks + 2
    # buffer that holds the output of conv launch n (1-based): see enqueue_tower
    cur = 0
    bufs = [0]
    for i in range(blocks):
        nxt = 2 if cur == 0 else 0
        bufs += [1, nxt]
        cur = nxt
    bufs.append(1)
    for n in range(1, nconv + 1):
        ref.debug_set_layer_limit(n)
        tc.debug_set_layer_limit(n)
        ref.predict_batch(boards, q, want_policy=False)
        tc.predict_batch(boards, q, want_policy=False)
        a = ref.debug_act(B, bufs[n - 1])
        b = tc.debug_act(B, bufs[n - 1])
        d = np.abs(a - b)
        print("after conv launch %2d (buf %d): max|d|=%.4e  mean|d|=%.3e  ref max=%.3f mean=%.3f  nan=%d" % (
            n, bufs[n - 1], d.max(), d.mean(), np.abs(a).max(), np.abs(a).mean(), int(np.isnan(b).sum())), flush=True)
        if n == 1 and d.max() > 0.1:
            bad = np.argwhere(d > 0.1)
            print("  first mismatches (b,c,y,x):", bad[:10].tolist())
            print("  ref", a[tuple(bad[0])], "got", b[tuple(bad[0])])
            print("  per-board max", d.reshape(B, -1).max(axis=1).tolist())
            print("  per-row(y) max", d.max(axis=(0, 1, 3)).tolist())
            print("  per-col(x) max", d.max(axis=(0, 1, 2)).tolist())
            print("  per-chan max (first 16)", d.max(axis=(0, 2, 3))[:16].tolist())
    ref.debug_set_layer_limit(0)
    tc.debug_set_layer_limit(0)
    pol_t, v_t, k_t = tc.predict_batch(boards, q)
    kl = (pol_r * (np.log(np.maximum(pol_r, 1e-30)) - np.log(np.maximum(pol_t, 1e-30)))).sum(axis=1)
    print("%s vs fp32 mode: max|dv_logit|=%.3e  max KL=%.3e  k=%.6f" % (dtype, np.abs(v_t - v_r).max(), kl.max(), k_t[0]))
    print("%s vs golden   : max|dv_logit|=%.3e" % (dtype, np.abs(v_t - g["v_logit"][:B]).max()))
    mi, mo = g["move_idx"], g["move_off"][:B + 1]
    pri, v2, vf, k2 = tc.eval_leaves(boards, q, mi[:mo[B]], mo)
    print("eval_leaves priors vs golden: max|d|=%.3e; v_final vs tanh: %.3e" % (
        np.abs(pri - g["priors"][:mo[B]]).max(), np.abs(vf - np.tanh(v2 + k2 * q)).max()))


if __name__ == "__main__":
    main()
