"""Deterministic OracleNet weight sets and the flat blob layout.

TEST INFRASTRUCTURE (weight *generation*); the blob LAYOUT documented here is
the product's load format (`cb_load_weights`), the reference's own flat order
(python/export_weights_cuda.py:73-107 == cuda/nn_weights.cuh:61-86) generalised
to (num_blocks, hidden).

Weights are drawn with numpy (PCG64) so that the GPU box — which has no
/root/reference — can rebuild exactly the tensors the golden fixtures were
computed with.  Distributions follow python/model.py:148-169:

  variant "A"  the reference's initialisation law: kaiming-normal(fan_out) conv
               and linear weights, BN gamma=1 beta=0 mean=0 var=1, default
               uniform biases, p_head and v_out zeroed, k_logit=0.  (Parity on
               A is vacuous downstream of the zero heads — SURVEY.md App. D.3.)
  variant "B"  "de-zeroed": same body, heads alive, BN statistics and affine
               randomised, k_logit=0.3, value-head scale chosen so |v_logit|
               stays O(1) like a trained net.
"""
import numpy as np

POLICY_PLANES = 73
VALUE_FC = 256


def field_shapes(num_blocks, hidden):
    """Ordered (name, shape) list — THE blob layout."""
    H, S = hidden, hidden // 16
    f = [("start_conv.weight", (H, 17, 3, 3))]
    f += _bn("start_bn", H)
    for i in range(num_blocks):
        p = "res_blocks.%d." % i
        f.append((p + "conv1.weight", (H, H, 3, 3)))
        f += _bn(p + "bn1", H)
        f.append((p + "conv2.weight", (H, H, 3, 3)))
        f += _bn(p + "bn2", H)
        f.append((p + "se.fc.0.weight", (S, H)))
        f.append((p + "se.fc.2.weight", (H, S)))
    f.append(("p_conv.weight", (H, H, 3, 3)))
    f += _bn("p_bn", H)
    f.append(("p_head.weight", (POLICY_PLANES, H)))
    f.append(("p_head.bias", (POLICY_PLANES,)))
    f.append(("v_conv.weight", (H,)))
    f.append(("v_conv.bias", (1,)))
    f += _bn("v_bn", 1)
    f.append(("v_fc.weight", (VALUE_FC, 65)))
    f.append(("v_fc.bias", (VALUE_FC,)))
    f.append(("v_out.weight", (VALUE_FC,)))
    f.append(("v_out.bias", (1,)))
    f.append(("k_logit", (1,)))
    return f


def _bn(prefix, c):
    return [(prefix + ".weight", (c,)), (prefix + ".bias", (c,)),
            (prefix + ".running_mean", (c,)), (prefix + ".running_var", (c,))]


def blob_size(num_blocks, hidden):
    return int(sum(int(np.prod(s)) for _, s in field_shapes(num_blocks, hidden)))


def make_weights(num_blocks, hidden, seed, variant):
    rng = np.random.default_rng(seed)
    w = {}
    for name, shape in field_shapes(num_blocks, hidden):
        leaf = name.split(".")[-1]
        if leaf == "running_mean":
            w[name] = rng.normal(0, 0.1, shape) if variant == "B" else np.zeros(shape)
        elif leaf == "running_var":
            w[name] = rng.uniform(0.5, 1.5, shape) if variant == "B" else np.ones(shape)
        elif "bn" in name and leaf == "weight":
            w[name] = rng.uniform(0.8, 1.2, shape) if variant == "B" else np.ones(shape)
        elif "bn" in name and leaf == "bias":
            w[name] = rng.normal(0, 0.1, shape) if variant == "B" else np.zeros(shape)
        elif name == "k_logit":
            w[name] = np.full(shape, 0.3 if variant == "B" else 0.0)
        elif leaf == "weight":
            if len(shape) == 4:
                fan_out = shape[0] * shape[2] * shape[3]
            elif name == "v_conv.weight":
                fan_out = 1  # Conv2d(H,1,1): fan_out = 1 -> std sqrt(2)
            elif name == "v_out.weight":
                fan_out = 1
            else:
                fan_out = shape[0]
            w[name] = rng.normal(0, np.sqrt(2.0 / fan_out), shape)
        else:  # conv / linear biases: torch default U(-1/sqrt(fan_in), 1/sqrt(fan_in))
            fan_in = {"p_head.bias": hidden, "v_conv.bias": hidden, "v_fc.bias": 65, "v_out.bias": VALUE_FC}[name]
            w[name] = rng.uniform(-1, 1, shape) / np.sqrt(fan_in)
    if variant == "A":
        for n in ("p_head.weight", "p_head.bias", "v_out.weight", "v_out.bias"):
            w[n] = np.zeros_like(w[n])
    else:
        w["p_head.weight"] = rng.normal(0, 0.05, w["p_head.weight"].shape)
        w["p_head.bias"] = rng.normal(0, 0.1, w["p_head.bias"].shape)
        # keep the single value channel alive after its ReLU and |v_logit| = O(1)
        w["v_conv.weight"] = rng.normal(0, np.sqrt(2.0 / hidden), w["v_conv.weight"].shape)
        w["v_out.weight"] = rng.normal(0, 0.1, w["v_out.weight"].shape)
        w["v_out.bias"] = rng.normal(0, 0.1, w["v_out.bias"].shape)
    return {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in w.items()}


def flatten(weights, num_blocks, hidden):
    parts = [weights[name].reshape(-1) for name, _ in field_shapes(num_blocks, hidden)]
    blob = np.concatenate(parts).astype(np.float32)
    assert blob.size == blob_size(num_blocks, hidden)
    return blob


def unflatten(blob, num_blocks, hidden):
    out, off = {}, 0
    for name, shape in field_shapes(num_blocks, hidden):
        n = int(np.prod(shape))
        out[name] = np.asarray(blob[off:off + n], dtype=np.float32).reshape(shape)
        off += n
    assert off == blob.size
    return out
