"""Restatement of the reference network's eval-mode forward from the flat blob.

TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).  This is the CPU fp32 checker
for the floating-point half of the path; it is pinned against outputs of the
reference's own python/model.py (tests/golden/*.npz, made by
oracle/make_golden.py) to 1e-5.

Follows python/model.py:97-145 (OracleNet.forward), :30-51 (ResBlock),
:8-26 (SEBlock).  Eval mode returns (log_softmax policy [B,4672], raw v_logit
[B,1], k [B,1]); BatchNorm uses running statistics with eps 1e-5.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import weights as W

BN_EPS = 1e-5


def _bn(x, w, p):
    return F.batch_norm(x, w[p + ".running_mean"], w[p + ".running_var"], w[p + ".weight"], w[p + ".bias"],
                        training=False, eps=BN_EPS)


def forward(blob, num_blocks, hidden, planes, q, return_backbone=False):
    """planes float32 [B,17,8,8]; q float32 [B] (column 0 of the reference's scalars; column 1 is unused,
    python/model.py:106)."""
    w = {k: torch.from_numpy(v) for k, v in W.unflatten(np.asarray(blob), num_blocks, hidden).items()}
    x = torch.as_tensor(planes, dtype=torch.float32)
    qt = torch.as_tensor(q, dtype=torch.float32).reshape(-1, 1)
    with torch.no_grad():
        x = F.relu(_bn(F.conv2d(x, w["start_conv.weight"], padding=1), w, "start_bn"))
        for i in range(num_blocks):
            p = "res_blocks.%d." % i
            r = x
            o = F.relu(_bn(F.conv2d(x, w[p + "conv1.weight"], padding=1), w, p + "bn1"))
            o = _bn(F.conv2d(o, w[p + "conv2.weight"], padding=1), w, p + "bn2")
            y = o.mean(dim=(2, 3))
            y = torch.sigmoid(F.linear(F.relu(F.linear(y, w[p + "se.fc.0.weight"])), w[p + "se.fc.2.weight"]))
            x = F.relu(o * y[:, :, None, None] + r)
        backbone = x
        pol = F.relu(_bn(F.conv2d(x, w["p_conv.weight"], padding=1), w, "p_bn"))
        pol = F.conv2d(pol, w["p_head.weight"].reshape(W.POLICY_PLANES, hidden, 1, 1), w["p_head.bias"])
        pol = pol.permute(0, 2, 3, 1).reshape(pol.shape[0], -1)  # (h*8+w)*73 + c
        log_policy = F.log_softmax(pol, dim=1)
        v = F.conv2d(x, w["v_conv.weight"].reshape(1, hidden, 1, 1), w["v_conv.bias"])
        v = F.relu(_bn(v, w, "v_bn")).reshape(-1, 64)
        v = torch.cat([v, qt], dim=1)
        v = F.relu(F.linear(v, w["v_fc.weight"], w["v_fc.bias"]))
        v_logit = F.linear(v, w["v_out.weight"].reshape(1, -1), w["v_out.bias"])
        k = 0.47 * F.softplus(w["k_logit"].reshape(()))
        k = k.reshape(1, 1).expand(x.shape[0], 1)
    out = (log_policy.numpy(), v_logit.numpy().reshape(-1), k.numpy().reshape(-1).copy())
    if return_backbone:
        return out + (backbone.numpy(),)
    return out
