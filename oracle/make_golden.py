"""Generate the golden fixtures (tests/golden/*.npz) and the traced reference
modules (oracle/_ref/*.pt) by RUNNING THE REFERENCE'S OWN CODE.

Runs only in the build container, where /root/reference exists.  The
reference network definition (python/model.py) is imported from where it lies
— never copied — loaded with the numpy-seeded weights of oracle/weights.py,
traced exactly as the reference exports it for its Rust evaluator
(python/orchestrate.py:219-230: eval(), torch.jit.trace with [1,17,8,8] and
[1,2] examples), and executed on CPU in fp32.  That traced module is "the
reference's own TorchScript CPU OracleNet" (src/neural_net.rs:115,246-252).

    python -m oracle.make_golden            # from the repo root

Outputs
  tests/golden/net_<cfg>.npz   committed, small: seed/config, boards, q, legal
                               CSR, reference log-policy / v_logit / k, a few
                               backbone activations, oracle-side priors.
  oracle/_ref/<cfg>.pt         git-ignored, travels to the GPU box: the traced
                               reference module for bench.py --impl reference
                               and cpu_baseline(kind="reference").
"""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from oracle import pyoracle, weights as W  # noqa: E402

REF_PY = "/root/reference/python"

# name, blocks, hidden, variant, weight seed, position seed, n positions, n with full policy
CONFIGS = [
    ("6x128_A", 6, 128, "A", 1001, 20261019, 64, 4),
    ("6x128_B", 6, 128, "B", 1003, 20261019, 256, 32),
    ("2x128_B", 2, 128, "B", 1005, 7, 64, 16),
    ("10x256_B", 10, 256, "B", 1007, 11, 32, 8),
]


def reference_model(blocks, hidden, weights):
    sys.path.insert(0, REF_PY)
    import model as ref_model  # the reference's python/model.py, imported in place
    m = ref_model.OracleNet(num_blocks=blocks, hidden_dim=hidden)
    sd = m.state_dict()
    for name, arr in weights.items():
        t = torch.from_numpy(arr)
        if name == "p_head.weight":
            t = t.reshape(73, hidden, 1, 1)
        elif name == "v_conv.weight":
            t = t.reshape(1, hidden, 1, 1)
        elif name == "v_out.weight":
            t = t.reshape(1, 256)
        elif name == "k_logit":
            t = t.reshape(())
        assert sd[name].shape == t.shape, (name, sd[name].shape, t.shape)
        sd[name] = t
    m.load_state_dict(sd)
    m.eval()
    # exactly export_model_for_rust (python/orchestrate.py:219-230)
    traced = torch.jit.trace(m, (torch.randn(1, 17, 8, 8), torch.randn(1, 2)))
    return m, traced


def main():
    os.makedirs(os.path.join(REPO, "tests", "golden"), exist_ok=True)
    os.makedirs(os.path.join(REPO, "oracle", "_ref"), exist_ok=True)
    torch.manual_seed(0)
    for name, blocks, hidden, variant, wseed, pseed, n, n_full in CONFIGS:
        boards, move_idx, move_off, moves_raw = pyoracle.random_positions(pseed, n, 160)
        planes = pyoracle.planes_batch(boards)
        # Variant B funnels the value path through ONE channel + ReLU, which can be dead (or never
        # clip) for a whole batch: walk the weight seed until it is mixed.
        while True:
            weights = W.make_weights(blocks, hidden, wseed, variant)
            model, traced = reference_model(blocks, hidden, weights)
            if variant != "B":
                break
            with torch.no_grad():
                xb = torch.relu(model.start_bn(model.start_conv(torch.from_numpy(planes))))
                for blk in model.res_blocks:
                    xb = blk(xb)
                alive_all = float((model.v_bn(model.v_conv(xb)) > 0).float().mean())
            if 0.3 < alive_all < 0.9:
                break
            wseed += 100
        traced.save(os.path.join(REPO, "oracle", "_ref", "oraclenet_%s.pt" % name))
        q = pyoracle.static_q(boards)
        flags = np.ones(n, dtype=np.float32)
        scalars = torch.from_numpy(np.stack([q, flags], axis=1))
        with torch.no_grad():
            logp, v_logit, k = traced(torch.from_numpy(planes), scalars)
            # backbone activation of the first positions, through the eager module
            x = torch.relu(model.start_bn(model.start_conv(torch.from_numpy(planes[:4]))))
            for blk in model.res_blocks:
                x = blk(x)
        logp = logp.numpy()
        v_logit = v_logit.numpy().reshape(-1)
        k = k.numpy().reshape(-1)

        # value path must be alive for variant B or parity on it is vacuous
        alive = alive_all if variant == "B" else 1.0
        probs = np.exp(logp)
        priors = np.zeros(move_idx.size, dtype=np.float32)
        for i in range(n):
            lo, hi = int(move_off[i]), int(move_off[i + 1])
            w_to_move = int(boards[i, 112])
            priors[lo:hi] = pyoracle.policy_to_move_priors(probs[i], list(moves_raw[lo:hi]), w_to_move)
        print("%-9s wseed=%d n=%d  v_logit[%.3f, %.3f] std %.3f  k=%.5f  v_feat alive=%.2f  policy std=%.2e"
              % (name, wseed, n, v_logit.min(), v_logit.max(), v_logit.std(), k[0], alive, logp.std()))
        np.savez_compressed(
            os.path.join(REPO, "tests", "golden", "net_%s.npz" % name),
            blocks=blocks, hidden=hidden, variant=variant, weight_seed=wseed, position_seed=pseed,
            boards=boards, q=q, move_idx=move_idx, move_off=move_off, moves_raw=moves_raw,
            log_policy_full=logp[:n_full].astype(np.float32), v_logit=v_logit.astype(np.float32),
            k=k.astype(np.float32), priors=priors, backbone=x.numpy().astype(np.float32),
            weights_checksum=np.float64(np.abs(W.flatten(weights, blocks, hidden).astype(np.float64)).sum()))


if __name__ == "__main__":
    main()
