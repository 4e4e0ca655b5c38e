"""Parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle and the
golden fixtures recorded from the reference's own network.  Run on the B200 box: pytest -m gpu.

Bars (BASELINE.json north_star): planes / static PeSTO bit-exact; value |d| <= 1e-3 in the exact
(fp32) mode and <= 1e-2 in the 16-bit tensor-core modes; policy KL <= 1e-4.
"""
import ctypes

import numpy as np
import pytest

from oracle import net, pyoracle as O

pytestmark = pytest.mark.gpu

TOL_V_FP32 = 1e-3
TOL_V_16 = 1e-2
TOL_KL = 1e-4

KAT_FENS = [
    "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1",
    "rnbqkbnr/pppppppp/8/8/4P3/8/PPPP1PPP/RNBQKBNR b KQkq e3 0 1",
    "rnbqkbnr/ppp1pppp/8/3pP3/8/8/PPPP1PPP/RNBQKBNR w KQkq d6 0 3",
    "4k3/8/8/8/8/8/P7/4K3 w - - 0 1", "4k3/p7/8/8/8/8/8/4K3 b - - 0 1",
    "4k3/8/8/8/8/8/8/4K2R w K - 0 1", "4k2r/8/8/8/8/8/8/4K3 b k - 0 1", "r3k3/8/8/8/8/8/8/4K3 w q - 0 1",
    "4k3/8/8/3pP3/8/8/8/4K3 w - d6 0 1", "4k3/8/8/8/3Pp3/8/8/4K3 b - d3 0 1",
    "r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1",
    "8/P7/8/8/8/8/8/K6k w - - 0 1", "4k3/8/8/8/8/8/8/4K3 b - - 0 1",
]


def kl(p_ref, p):
    return (p_ref * (np.log(np.maximum(p_ref, 1e-30)) - np.log(np.maximum(p, 1e-30)))).sum(axis=1)


@pytest.fixture(scope="module")
def positions():
    return O.random_positions(20261019, 4096)


def make_eval(cb, g, blob, dtype):
    ev = cb.Evaluator(int(g["blocks"]), int(g["hidden"]), dtype)
    ev.load_weights(blob)
    assert ev.is_available()
    return ev


# ------------------------------------------------------------------ integer half: bit-exact
def test_encode_bit_exact(cb, positions):
    ev = cb.Evaluator(2, 128, "bf16")
    fens = O.boards_to_array([O.from_fen(f) for f in KAT_FENS])
    for boards in (fens, positions[0][:2048], positions[0][:1], positions[0][:777]):
        got = ev.encode(boards)
        assert got.dtype == np.float32 and np.array_equal(got, O.planes_batch(boards))
    # the literal cells the reference pins (tests/unit/board_encoding_tests.rs:119-171,298-329)
    got = ev.encode(fens)
    assert got[3, 0, 6, 0] == 1.0 and got[4, 0, 6, 0] == 1.0
    assert got[2, 12, 2, 3] == 1.0 and got[1, 12, 2, 4] == 1.0
    assert got[0, 13:17].min() == 1.0 and got[5, 13].min() == 1.0 and got[5, 14:17].max() == 0.0


def test_static_pesto_bit_exact(cb, positions):
    ev = cb.Evaluator(2, 128, "bf16")
    boards = positions[0]
    want = np.array([O.pst_eval_cp(O.array_to_board(b)) for b in boards], dtype=np.int32)
    assert np.array_equal(ev.pst_eval_cp(boards), want)
    assert len(set(want.tolist())) > 500


# ------------------------------------------------------------------ network: golden fixtures
@pytest.mark.parametrize("name,cuda_cores", [("6x128_A", False), ("6x128_B", False), ("2x128_B", False), ("6x128_B", True),
                                             ("2x128_B", True), ("10x256_B", True)])
def test_exact_mode_matches_reference(cb, golden, name, cuda_cores, monkeypatch):
    """CB_DTYPE_FP32, the mode that answers the north star's fp32 tolerance (|dV| <= 1e-3, KL <= 1e-4), held here to
    1e-4 / 1e-5.  128-channel nets run it on the tensor cores: the resident forward with fp16-PAIR operands (x = hi + lo,
    three MMAs per product, fp32 accumulation); CB_FP32_EXACT=1 and the 256-channel nets run fp32 FMA on CUDA cores."""
    if cuda_cores:
        monkeypatch.setenv("CB_FP32_EXACT", "1")
    g, blob = golden(name)
    ev = make_eval(cb, g, blob, "fp32")
    n = g["boards"].shape[0]
    pol, v, k = ev.predict_batch(g["boards"], g["q"], np.ones(n, dtype=np.uint8))
    assert ev.launches_per_forward() == (1 if not cuda_cores else 3 * int(g["blocks"]) + 5)
    nf = g["log_policy_full"].shape[0]
    assert np.abs(v - g["v_logit"]).max() <= 1e-4 < TOL_V_FP32
    assert np.abs(k - g["k"]).max() <= 1e-6
    assert kl(np.exp(g["log_policy_full"]), pol[:nf]).max() <= 1e-5 < TOL_KL
    np.testing.assert_allclose(pol.sum(axis=1), 1.0, atol=1e-4)
    bb = ev.debug_backbone(4)
    # (fp16 pairs carry 22 significant bits: measured 7.6e-5 on activations of magnitude ~30, fp32 FMA 1e-6)
    assert np.abs(bb - g["backbone"]).max() <= (1e-4 if cuda_cores else 1e-5 * max(float(np.abs(g["backbone"]).max()), 10.0))


@pytest.mark.parametrize("name,dtype", [("6x128_A", "bf16"), ("6x128_B", "bf16"), ("6x128_B", "fp16"),
                                        ("2x128_B", "bf16"), ("10x256_B", "bf16"), ("10x256_B", "fp16")])
def test_tensor_core_modes_match_reference(cb, golden, name, dtype):
    g, blob = golden(name)
    ev = make_eval(cb, g, blob, dtype)
    n = g["boards"].shape[0]
    q = g["q"]
    pol, v, k = ev.predict_batch(g["boards"], q, None)
    nf = g["log_policy_full"].shape[0]
    v_final_ref = np.tanh(g["v_logit"].astype(np.float64) + g["k"].astype(np.float64) * q)
    v_final = np.tanh(v.astype(np.float64) + k.astype(np.float64) * q)
    dv, dl = np.abs(v_final - v_final_ref).max(), np.abs(v - g["v_logit"]).max()
    klm = kl(np.exp(g["log_policy_full"]), pol[:nf]).max()
    print("%s %s: max|dV_final|=%.3e max|dv_logit|=%.3e max KL=%.3e" % (name, dtype, dv, dl, klm))
    # (the 10x256 net meets the bar in bf16 too since the pair kernel keeps the skip path in fp16: 9.2e-3; it was 1.2e-2
    #  with bf16-stored block inputs)
    assert dv <= TOL_V_16, "value tolerance (north star: <= 1e-2 in 16-bit modes)"
    assert klm <= TOL_KL
    assert np.abs(k - g["k"]).max() <= 1e-6
    if str(g["variant"]) == "A":      # zero heads: exactly uniform policy, zero logit
        assert np.abs(v).max() == 0.0 and np.abs(pol - 1.0 / 4672).max() < 1e-9
    # backbone activation within 16-bit rounding of the reference (not vacuous for variant A either)
    bb = ev.debug_backbone(4)
    scale = np.abs(g["backbone"]).max()
    assert np.abs(bb - g["backbone"]).max() <= (0.03 if dtype == "bf16" else 0.005) * max(scale, 1.0)


@pytest.mark.parametrize("env,name,dtype", [({"CB_NO_FUSE": "1"}, "6x128_B", "bf16"), ({"CB_NO_FUSE": "1"}, "10x256_B", "fp16"),
                                            ({"CB_TOWER_PAIR": "1"}, "6x128_B", "bf16"), ({"CB_TOWER_PAIR": "1"}, "2x128_B", "fp16"),
                                            ({"CB_CLUSTER": "1"}, "6x128_B", "bf16"), ({"CB_SIMT_POLICY": "1"}, "10x256_B", "fp16"),
                                            ({"CB_NO_GRAPH": "1"}, "6x128_B", "fp16"), ({"CB_NO_DEEP_RING": "1"}, "2x128_B", "bf16")],
                         ids=lambda x: "+".join(x) if isinstance(x, dict) else x)
def test_every_kernel_path_in_the_library_matches_reference(cb, golden, monkeypatch, env, name, dtype):
    """The switches that select the other kernels the library ships — one launch per layer (CB_NO_FUSE), the pair-layout
    global-round-trip tower on a 128-channel net (CB_TOWER_PAIR; it is the 256-channel nets' default), the 2-CTA weight
    multicast of the resident kernel (CB_CLUSTER), the SIMT policy head (CB_SIMT_POLICY), eager launches (CB_NO_GRAPH),
    the 4-stage ring for small batches (CB_NO_DEEP_RING): each must meet the same bars as the default path."""
    for k_, v_ in env.items():
        monkeypatch.setenv(k_, v_)
    g, blob = golden(name)
    ev = make_eval(cb, g, blob, dtype)
    q = g["q"]
    pol, v, k = ev.predict_batch(g["boards"], q, None)
    nf = g["log_policy_full"].shape[0]
    ref = np.tanh(g["v_logit"].astype(np.float64) + g["k"].astype(np.float64) * q)
    got = np.tanh(v.astype(np.float64) + k.astype(np.float64) * q)
    assert np.abs(got - ref).max() <= TOL_V_16
    assert kl(np.exp(g["log_policy_full"]), pol[:nf]).max() <= TOL_KL
    pri, v2, vf, _ = ev.eval_leaves(g["boards"], q, g["move_idx"], g["move_off"], material_on=True)
    assert np.abs(pri - g["priors"]).max() <= 2e-3 and np.abs(v2 - v).max() <= 1e-3
    # small batches too (the deep-ring / single-chain launch shapes of the resident kernel)
    pri1, v1, _, _ = ev.eval_leaves(g["boards"][:3], q[:3], g["move_idx"][:g["move_off"][3]], g["move_off"][:4], material_on=True)
    assert np.abs(v1 - v[:3]).max() <= 1e-3 and np.abs(pri1 - pri[:g["move_off"][3]]).max() <= 1e-3


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_fused_leaf_eval_matches_reference(cb, golden, dtype):
    g, blob = golden("6x128_B")
    ev = make_eval(cb, g, blob, dtype)
    n = g["boards"].shape[0]
    pri, v, vf, k = ev.eval_leaves(g["boards"], g["q"], g["move_idx"], g["move_off"], material_on=True)
    # (fp32 mode = fp16-pair operands on the tensor cores; its softmax uses the fast exp / log of the resident kernel)
    tol_p, tol_v = (1e-5, 1e-4) if dtype == "fp32" else (2e-3, TOL_V_16)
    print("%s: max |prior - reference| = %.3e" % (dtype, np.abs(pri - g["priors"]).max()))
    assert np.abs(pri - g["priors"]).max() <= tol_p
    for i in range(n):
        lo, hi = int(g["move_off"][i]), int(g["move_off"][i + 1])
        assert abs(pri[lo:hi].sum() - 1.0) < 1e-5
    want = np.array([O.value_fuse(v[i], k[i], g["q"][i], True) for i in range(n)])
    assert np.abs(vf - want).max() <= 1e-6                      # device tanhf vs the host's f64 tanh
    ref_final = np.array([O.value_fuse(g["v_logit"][i], g["k"][i], g["q"][i], True) for i in range(n)])
    assert np.abs(vf - ref_final).max() <= tol_v
    # material off (vanilla / KOTH config 5): V = tanh(v_logit)  (src/mcts/tactical_mcts.rs:787-790)
    pri2, v2, vf2, _ = ev.eval_leaves(g["boards"], None, g["move_idx"], g["move_off"], material_on=False)
    assert np.abs(vf2 - np.tanh(v2.astype(np.float64))).max() <= 1e-6
    assert np.array_equal(pri2, pri)                            # policy is independent of q


# ------------------------------------------------------------------ edge cases
def test_edge_cases(cb, golden):
    g, blob = golden("2x128_B")
    ev = make_eval(cb, g, blob, "bf16")
    boards, q = g["boards"], g["q"]
    # B = 1 (NeuralNetPolicy::predict) and odd B agree bit-for-bit with the same boards inside a big batch
    pol_all, v_all, _ = ev.predict_batch(boards, q)
    pol1, v1, _ = ev.predict_batch(boards[:1], q[:1])
    assert np.array_equal(pol1[0], pol_all[0]) and v1[0] == v_all[0]
    pol7, v7, _ = ev.predict_batch(boards[:7], q[:7])
    assert np.array_equal(pol7, pol_all[:7]) and np.array_equal(v7, v_all[:7])
    # q omitted == q zero
    _, v0a, _ = ev.predict_batch(boards[:8], None, want_policy=False)
    _, v0b, _ = ev.predict_batch(boards[:8], np.zeros(8, dtype=np.float32), want_policy=False)
    assert np.array_equal(v0a, v0b)
    # empty batch is a no-op
    L = cb.load_library()
    assert L.cb_predict_batch(ev._h, None, None, None, 0, None, None, None) in (0, -1)
    # ragged CSR: an empty list, the maximum list, out-of-range indices -> uniform fallback
    B = 4
    off = np.array([0, 0, 218, 221, 222], dtype=np.uint32)
    idx = np.zeros(222, dtype=np.uint16)
    idx[:218] = np.arange(218) * 21
    idx[218:221] = 60000                  # invalid -> contributes 0 -> sum 0 -> uniform (src/tensor.rs:202-211)
    idx[221] = 877
    pri, _, _, _ = ev.eval_leaves(boards[:B], q[:B], idx, off)
    assert abs(pri[:218].sum() - 1.0) < 1e-5
    np.testing.assert_allclose(pri[218:221], 1.0 / 3, rtol=1e-6)
    assert pri[221] == 1.0
    # duplicates are counted as listed
    off2 = np.array([0, 2], dtype=np.uint32)
    pri2, _, _, _ = ev.eval_leaves(boards[:1], q[:1], np.array([877, 877], dtype=np.uint16), off2)
    np.testing.assert_allclose(pri2, [0.5, 0.5], rtol=1e-6)


def test_prepared_eval_leaves_equals_eval_leaves(cb, golden):
    g, blob = golden("2x128_B")
    ev = make_eval(cb, g, blob, "bf16")
    want = ev.eval_leaves(g["boards"], g["q"], g["move_idx"], g["move_off"])
    run = ev.prepared_eval_leaves(g["boards"], g["q"], g["move_idx"], g["move_off"])
    for _ in range(2):
        got = run()
        for a, b in zip(want, got):
            assert np.array_equal(a, b)


def test_errors_are_codes_not_aborts(cb):
    ev = cb.Evaluator(2, 128, "bf16")
    with pytest.raises(cb.CaissaError, match="not loaded"):
        ev.predict_batch(np.zeros((2, 128), dtype=np.uint8))
    with pytest.raises(cb.CaissaError, match="expected"):
        ev.load_weights(np.zeros(10, dtype=np.float32))
    with pytest.raises(cb.CaissaError):
        cb.Evaluator(2, 96, "bf16")


def test_two_models_coexist(cb, golden):
    # evaluate_models keeps candidate and current alive at once (src/bin/evaluate_models.rs:558-584)
    ga, ba = golden("6x128_A")
    gb, bb = golden("6x128_B")
    ea, eb = make_eval(cb, ga, ba, "bf16"), make_eval(cb, gb, bb, "bf16")
    boards, q = gb["boards"][:32], gb["q"][:32]
    for _ in range(3):
        _, va, _ = ea.predict_batch(boards, q, want_policy=False)
        _, vb, _ = eb.predict_batch(boards, q, want_policy=False)
        assert np.abs(va).max() == 0.0
        assert np.abs(vb - gb["v_logit"][:32]).max() < 0.05


# ------------------------------------------------------------------ full-size properties
@pytest.mark.parametrize("B", [1024, 4096])
def test_full_size_properties(cb, golden, positions, B):
    g, blob = golden("6x128_B")
    ev = make_eval(cb, g, blob, "bf16")
    boards, idx, off, _ = positions
    boards, off = boards[:B], off[:B + 1]
    idx = idx[:off[B]]
    q = O.static_q(boards)
    pri, v, vf, k = ev.eval_leaves(boards, q, idx, off)
    # deterministic
    pri_b, v_b, vf_b, _ = ev.eval_leaves(boards, q, idx, off)
    assert np.array_equal(pri, pri_b) and np.array_equal(v, v_b) and np.array_equal(vf, vf_b)
    # priors are distributions; value fusion identity
    sums = np.add.reduceat(pri, off[:-1].astype(np.int64))
    assert np.abs(sums - 1.0).max() < 1e-5 and pri.min() >= 0.0
    assert np.abs(vf - np.tanh(v.astype(np.float64) + k.astype(np.float64) * q)).max() < 1e-6
    assert np.abs(vf).max() <= 1.0
    # batch-composition invariance: reversing the batch permutes the results, bit for bit
    rev = np.arange(B)[::-1]
    roff = np.zeros(B + 1, dtype=np.uint32)
    counts = (off[1:] - off[:-1])[rev]
    roff[1:] = np.cumsum(counts)
    ridx = np.concatenate([idx[off[i]:off[i + 1]] for i in rev])
    pri_r, v_r, _, _ = ev.eval_leaves(boards[rev], q[rev], ridx, roff)
    assert np.array_equal(v_r, v[rev])
    assert np.array_equal(pri_r[:counts[0]], pri[off[B - 1]:off[B]])
    # a sample of the batch against the CPU oracle network
    sel = np.arange(0, B, B // 16)
    logp, v_ref, k_ref = net.forward(blob, 6, 128, O.planes_batch(boards[sel]), q[sel])
    assert np.abs(np.tanh(v[sel] + k[sel] * q[sel]) - np.tanh(v_ref + k_ref * q[sel])).max() <= TOL_V_16
    pol, _, _ = ev.predict_batch(boards[sel], q[sel])
    assert kl(np.exp(logp), pol).max() <= TOL_KL


@pytest.mark.parametrize("dtype", ["bf16", "fp16"])
def test_pair_kernel_batch_shapes_are_bit_identical(cb, golden, positions, dtype):
    """The 256-channel forward on CTA pairs (tower_r2_umma.cuh): full units (two chains of 4 boards), tail units (one
    chain), ragged ends, several rounds per pair — every position must get exactly the result it gets inside the big
    batch, and the big batch must meet the parity bar on the golden boards."""
    g, blob = golden("10x256_B")
    ev = make_eval(cb, g, blob, dtype)
    boards, idx, off, _ = positions
    Bmax = 1400                                   # 175 units of 8 on 74 pairs: 2 full rounds + 27 units cut into 54 single chains
    q = O.static_q(boards[:Bmax])
    pri, v, vf, k = ev.eval_leaves(boards[:Bmax], q, idx[:off[Bmax]], off[:Bmax + 1])
    assert np.isfinite(pri).all() and np.isfinite(v).all() and np.abs(v).max() > 0.05
    for B in (1, 2, 3, 4, 5, 7, 8, 9, 16, 37, 296, 297, 300, 592, 593, 640, 1184, 1185):
        p_b, v_b, vf_b, _ = ev.eval_leaves(boards[:B], q[:B], idx[:off[B]], off[:B + 1])
        assert np.array_equal(v_b, v[:B]), "v_logit differs at batch %d" % B
        assert np.array_equal(vf_b, vf[:B]), "v_final differs at batch %d" % B
        assert np.array_equal(p_b, pri[:off[B]]), "priors differ at batch %d" % B
    # the golden boards inside a batch large enough for full units
    n = g["boards"].shape[0]
    big = np.concatenate([boards[:800], g["boards"]])
    pol, vg, kg = ev.predict_batch(big, np.concatenate([q[:800], g["q"]]), None)
    ref = np.tanh(g["v_logit"].astype(np.float64) + g["k"].astype(np.float64) * g["q"])
    got = np.tanh(vg[800:].astype(np.float64) + kg[800:].astype(np.float64) * g["q"])
    assert np.abs(got - ref).max() <= TOL_V_16
    nf = g["log_policy_full"].shape[0]
    assert kl(np.exp(g["log_policy_full"]), pol[800:800 + nf]).max() <= TOL_KL
    small, vs, _ = ev.predict_batch(g["boards"], g["q"], None)
    assert np.array_equal(vs, vg[800:]) and np.array_equal(small, pol[800:])


@pytest.mark.parametrize("dtype", ["bf16", "fp32"])
def test_ragged_batch_sizes_are_bit_identical_to_the_full_batch(cb, golden, positions, dtype):
    """Every way the resident tower splits a batch into units / chains (1..4 boards per chain, with and
    without y-halo rows, one or several units per SM) must give each position exactly the result it gets
    inside the full batch: priors, v_logit and v_final bit for bit, for the first B positions."""
    g, blob = golden("6x128_B")
    ev = make_eval(cb, g, blob, dtype)
    boards, idx, off, _ = positions
    Bmax = 2500
    q = O.static_q(boards[:Bmax])
    pri, v, vf, k = ev.eval_leaves(boards[:Bmax], q, idx[:off[Bmax]], off[:Bmax + 1])
    assert np.isfinite(pri).all() and np.isfinite(v).all()
    sizes = (1, 2, 3, 4, 5, 6, 7, 8, 9, 15, 16, 17, 31, 33, 147, 148, 149, 295, 296, 297, 443, 591, 1023, 1183, 1184,
             1185, 1337, 2049)
    for B in sizes if dtype == "bf16" else (1, 2, 3, 4, 5, 9, 148, 149, 297, 591, 593, 1185):
        p_b, v_b, vf_b, _ = ev.eval_leaves(boards[:B], q[:B], idx[:off[B]], off[:B + 1])
        assert np.array_equal(v_b, v[:B]), "v_logit differs at batch %d" % B
        assert np.array_equal(vf_b, vf[:B]), "v_final differs at batch %d" % B
        assert np.array_equal(p_b, pri[:off[B]]), "priors differ at batch %d" % B
    # compat path (full 4672-wide probabilities) on a ragged size
    pol, v_c, _ = ev.predict_batch(boards[:37], q[:37])
    assert np.array_equal(v_c, v[:37])
    np.testing.assert_allclose(pol.sum(axis=1), 1.0, atol=1e-4)
    for i in (0, 17, 36):
        lo, hi = int(off[i]), int(off[i + 1])
        want = pol[i][idx[lo:hi]]
        np.testing.assert_allclose(pri[lo:hi], want / want.sum(), rtol=2e-6, atol=1e-9)


def test_cluster_mode_is_bit_identical(cb, golden, positions, monkeypatch):
    """CB_CLUSTER=1: pairs of CTAs share the weight stream (TMA multicast, commit multicast).  Same numbers."""
    g, blob = golden("6x128_B")
    boards, idx, off, _ = positions
    ev = make_eval(cb, g, blob, "bf16")
    monkeypatch.setenv("CB_CLUSTER", "1")
    evc = make_eval(cb, g, blob, "bf16")
    monkeypatch.delenv("CB_CLUSTER")
    for B in (1024, 512, 128, 16, 2048):     # two chains per CTA, one chain per CTA, one board per CTA; 2048: no clusters
        q = O.static_q(boards[:B])
        a = ev.eval_leaves(boards[:B], q, idx[:off[B]], off[:B + 1])
        b = evc.eval_leaves(boards[:B], q, idx[:off[B]], off[:B + 1])
        for x, y in zip(a, b):
            assert np.array_equal(x, y), B


def test_resident_path_equals_host_path(cb, golden, positions):
    g, blob = golden("6x128_B")
    ev = make_eval(cb, g, blob, "bf16")
    boards, idx, off, _ = positions
    B = 512
    q = O.static_q(boards[:B])
    want = ev.eval_leaves(boards[:B], q, idx[:off[B]], off[:B + 1])
    ev.stage_leaves(boards[:B], q, idx[:off[B]], off[:B + 1])
    ms = ev.time_resident(3, material_on=True, flush_l2=True)
    assert ms.min() > 0
    got = ev.read_resident()
    for a, b in zip(want, got):
        assert np.array_equal(a, b)
    conv_ms, n = ev.time_conv_layers(2)
    assert n == 13 and conv_ms > 0
    tower_ms, launches = ev.time_tower(3)
    assert launches == 1 and tower_ms.min() > 0            # the fused persistent tower kernel
    assert ev.launches_per_forward() == 1                  # planes, tower and both heads in one launch


def test_submit_wait_equals_eval_leaves(cb, golden, positions):
    """cb_eval_leaves_submit + cb_eval_leaves_wait == cb_eval_leaves bit for bit; one host thread drives two handles with
    their batches interleaved (the second handle's copies run under the first one's kernel)."""
    g, blob = golden("6x128_B")
    a, b = make_eval(cb, g, blob, "bf16"), make_eval(cb, g, blob, "bf16")
    boards, idx, off, _ = positions
    q = O.static_q(boards)
    cuts = [(0, 700), (700, 1724), (1724, 1725), (1725, 2048)]
    want = [a.eval_leaves(boards[s:e], q[s:e], idx[off[s]:off[e]], off[s:e + 1] - off[s]) for s, e in cuts]
    pairs = [(a if i % 2 == 0 else b).prepared_submit_wait(boards[s:e], q[s:e], idx[off[s]:off[e]], off[s:e + 1] - off[s])
             for i, (s, e) in enumerate(cuts)]
    pairs[0][0]()
    pairs[1][0]()
    got = [None] * 4
    got[0] = [x.copy() for x in pairs[0][1]()]
    pairs[2][0]()
    got[1] = [x.copy() for x in pairs[1][1]()]
    pairs[3][0]()
    got[2] = [x.copy() for x in pairs[2][1]()]
    got[3] = [x.copy() for x in pairs[3][1]()]
    for w, gt in zip(want, got):
        for x, y in zip(w, gt):
            assert np.array_equal(x, y)
    with pytest.raises(cb.CaissaError):
        pairs[0][1]()                                  # nothing in flight
    pairs[0][0]()
    with pytest.raises(cb.CaissaError):
        pairs[0][0]()                                  # one batch per handle
    pairs[0][1]()


# ------------------------------------------------------------------ device-side tree operations
@pytest.mark.parametrize("name,material,dtype,koth", [("6x128_A", False, "bf16", False), ("6x128_B", False, "bf16", False),
                                                      ("6x128_B", True, "bf16", False), ("6x128_B", False, "bf16", True),
                                                      ("2x128_B", True, "fp32", False), ("10x256_B", False, "bf16", False),
                                                      ("6x128_B", "pe", "bf16", False), ("2x128_B", "pe", "fp32", True),
                                                      ("6x128_B", "gates", "bf16", False), ("6x128_B", "gates", "bf16", True)])
def test_device_tree_ops_equal_the_host_driver(cb, golden, name, material, dtype, koth, tmp_path):
    """Self-play with select / expand / backup / move choice on the device (pipeline 3) must walk exactly the
    trees of the host driver: same evaluator outputs (they do not depend on batch composition), same chess
    rules (shared inline code), same explicitly rounded PUCT arithmetic, same RNG stream => identical counters."""
    g, blob = golden(name)
    ev = make_eval(cb, g, blob, dtype)
    kw = dict(games=24, sims_per_move=20, max_plies=40, material_on=bool(material), koth=koth, seed=11,
              max_sims=24 * 20 * 60, qsearch="pe" if material == "pe" else 0,   # "pe": dM by the exchange line (warp form)
              tier1_light=material == "gates")                                  # "gates": light Tier-1 gates + adjudication
    host = ev.selfplay(threads=1, pipeline=1, sample_path=str(tmp_path / "host.bin"), **kw)
    dev = ev.selfplay(pipeline=3, sample_path=str(tmp_path / "dev.bin"), **kw)
    for k in ("sims", "steps", "nn_evals", "terminal_hits", "moves", "games_finished", "samples", "white_wins",
              "black_wins", "draws"):
        assert host[k] == dev[k], (k, host[k], dev[k])
    assert dev["moves"] > 24 * 40 and dev["games_finished"] >= 24 and dev["overflow"] == 0
    # every step of a game is an evaluation (a simulation's leaf, or the policy request of a new game's root) or a decided leaf
    assert dev["sims"] == dev["nn_evals"] + dev["terminal_hits"]
    # the training samples too (5763-f32 records, src/training_data.rs:20-60); games that end in the same step
    # may be written in a different order
    rh = np.fromfile(tmp_path / "host.bin", dtype=np.float32).reshape(-1, 5763)
    rd = np.fromfile(tmp_path / "dev.bin", dtype=np.float32).reshape(-1, 5763)
    assert rh.shape == rd.shape and rh.shape[0] == host["samples"] > 0
    key = lambda r: r[np.lexsort(r.T[::-1])]
    assert np.array_equal(key(rh), key(rd))


SEARCH_FENS = ["r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1",
               "r3k2r/Pppp1ppp/1b3nbN/nP6/BBP1P3/q4N2/Pp1P2PP/R2Q1RK1 w kq - 0 1",
               "8/2p5/3p4/KP5r/1R3p1k/8/4P1P1/8 w - - 0 1",
               "rnbq1k1r/pp1Pbppp/2p5/8/2B5/8/PPP1NnPP/RNBQK2R w KQ - 1 8",
               "r1b2bnr/p2q1p1p/3kP1p1/2p5/p7/N4N2/PP1PPPPP/R1BK1B1R w - - 0 10",
               "4k3/8/8/3pP3/8/8/8/4K3 w - d6 0 1", "1n2k3/P1P5/8/8/8/8/8/4K3 w - - 0 1", "8/8/8/8/8/3k4/8/4K3 w - - 0 1",
               "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1"]


def search_corpus(n_random, koth=None, deep=False):
    """Positions for the one-move searches; with the gates on, roots the game loop would adjudicate are left out
    (src/bin/self_play.rs:241-283; see tests/test_host_search.py)."""
    boards, _, _, _ = O.random_positions(99, n_random, 160)
    boards = np.concatenate([O.boards_to_array([O.from_fen(f) for f in SEARCH_FENS]), boards])
    if koth is not None and deep:
        def adjudicated(b):
            return (koth and O.koth_center_in_n(b, 3)[0] is not None) or O.mate_search(b, 5, 2)[0] != 0 or \
                   (koth and O.light_gates(b, koth=True) == -1)
        boards = boards[[i for i in range(boards.shape[0]) if not adjudicated(O.array_to_board(boards[i]))]]
    elif koth is not None:
        boards = boards[[i for i in range(boards.shape[0]) if O.light_gates(O.array_to_board(boards[i]), koth=koth) == 2]]
    return boards


@pytest.mark.parametrize("mock,mode", [((0.0, 0.5), dict()), ((20.0, 0.5), dict()), ((-20.0, 0.5), dict(koth=True)),
                                       ((0.0, 0.5), dict(koth=True, tier1_light=True)), ((0.0, 0.5), dict(tier1_light=True)),
                                       ((0.0, 0.5), dict(shuffle_children=True)), ((0.0, 0.5), dict(tier1_light="deep")),
                                       ((0.0, 0.5), dict(tier1_light="deep", koth=True)), ((-20.0, 0.5), dict(tier1_light="deep"))],
                         ids=["uniform", "biased+", "biased-koth", "uniform-koth-gates", "uniform-gates", "uniform-shuffle",
                              "uniform-deep-gates", "uniform-koth-deep-gates", "biased-deep-gates"])
def test_device_search_equals_the_oracle_with_the_mock_evaluator(cb, golden, mock, mode, monkeypatch):
    """The device tree kernels against oracle/mcts.c with the reference's mock evaluator (InferenceServer::new_mock_biased:
    uniform priors, constant v_logit — 0 and a saturated +-20 give exact values on both sides): ties everywhere, so child
    order, the tactical phase, the root's force-visit and the first-maximum rule decide every visit count.  200 simulations
    from each of >= 64 positions; visit counts AND f64 value sums must be identical."""
    g, blob = golden("2x128_B")
    ev = make_eval(cb, g, blob, "bf16")
    if mode.get("tier1_light") == "deep" and mode.get("koth"):
        monkeypatch.setenv("CB_GATE_WARPS", "4")      # the gate jobs shared by four warps per game (every fourth root move each)
    boards = search_corpus(72, koth=mode.get("koth", False) if mode.get("tier1_light") else None, deep=mode.get("tier1_light") == "deep")
    assert boards.shape[0] >= 64      # (deep gates: gate jobs are suspended after CB_GATE_BUDGET positions per step and resumed)
    kids, root_visits = ev.debug_device_search(boards, sims_per_move=200, seed=3, mock=mock, **mode)
    okw = dict(mode)
    okw["shuffle"] = okw.pop("shuffle_children", False)
    cfg = O.mcts_config(sims=200, **okw)
    for i in range(boards.shape[0]):
        seed = (3 * 0x9E3779B97F4A7C15 + 0x632BE59BD9B4E019 * (i + 1)) % (1 << 64)
        t = O.Tree(O.array_to_board(boards[i]), seed)
        t.search(cfg, ("mock", mock[0], mock[1]))
        want = t.root_children()
        assert [(m, v) for m, v, _ in kids[i]] == [(m, v) for m, v, _ in want], (i, mode)
        assert [s for _, _, s in kids[i]] == [s for _, _, s in want], (i, mode)
        assert root_visits[i] == t.root_stats()["visits"]


@pytest.mark.parametrize("material,dtype", [(False, "bf16"), ("pe", "bf16"), (False, "fp32"), ("cap1", "bf16")])
def test_device_search_equals_the_oracle_with_the_network(cb, golden, material, dtype):
    """The same comparison with the golden 6x128 network as evaluator: the oracle's callback evaluates each leaf through
    cb_eval_leaves (one board per call; the forward is bit-identical whatever the batch), so both searches see the same
    priors and values and must build the same tree — 64 positions x 200 simulations."""
    g, blob = golden("6x128_B")
    ev = make_eval(cb, g, blob, dtype)     # fp32: the fp16-pair instantiation of the forward under the tree kernels
    boards = search_corpus(56)[:64]
    mode = dict(material_on=bool(material), qsearch=material if material else 0)
    kids, root_visits = ev.debug_device_search(boards, sims_per_move=200, seed=3, **mode)

    def net(b, moves, q, material_on):
        bu = np.frombuffer(bytes(b), dtype=np.uint8)[None]
        idx = np.array([O.move_to_index_stm(m, b.w_to_move) for m in moves], dtype=np.uint16)
        pri, _, vf, _ = ev.eval_leaves(bu, np.array([q], dtype=np.float32), idx, np.array([0, len(moves)], dtype=np.uint32),
                                       material_on=material_on)
        return pri, float(vf[0])
    cfg = O.mcts_config(sims=200, **mode)
    tactical = 0
    for i in range(boards.shape[0]):
        t = O.Tree(O.array_to_board(boards[i]), 1)
        t.search(cfg, net)
        want = t.root_children()
        assert [(m, v) for m, v, _ in kids[i]] == [(m, v) for m, v, _ in want], i
        assert [s for _, _, s in kids[i]] == [s for _, _, s in want], i
        tactical += t.root_stats()["tactical"]
    assert tactical > 500          # the tactical phase did steer leaves


def test_device_search_with_the_deep_gates_equals_the_oracle_with_the_network(cb, golden, monkeypatch):
    """The reference's self-play configuration of Tier 1 (mate-in-5 by checks at every leaf, src/bin/self_play.rs:213) under
    the golden network, with a gate budget of 2 positions per step so that most gate jobs are suspended and resumed."""
    monkeypatch.setenv("CB_GATE_BUDGET", "2")
    g, blob = golden("6x128_B")
    ev = make_eval(cb, g, blob, "bf16")
    boards = search_corpus(56, koth=False, deep=True)[:48]
    mode = dict(material_on=True, qsearch="pe", tier1_light="deep")
    kids, root_visits = ev.debug_device_search(boards, sims_per_move=200, seed=3, **mode)

    def net(b, moves, q, material_on):
        bu = np.frombuffer(bytes(b), dtype=np.uint8)[None]
        idx = np.array([O.move_to_index_stm(m, b.w_to_move) for m in moves], dtype=np.uint16)
        pri, _, vf, _ = ev.eval_leaves(bu, np.array([q], dtype=np.float32), idx, np.array([0, len(moves)], dtype=np.uint32),
                                       material_on=material_on)
        return pri, float(vf[0])
    cfg = O.mcts_config(sims=200, **mode)
    for i in range(boards.shape[0]):
        t = O.Tree(O.array_to_board(boards[i]), 1)
        t.search(cfg, net)
        want = t.root_children()
        assert [(m, v) for m, v, _ in kids[i]] == [(m, v) for m, v, _ in want], i
        assert [s for _, _, s in kids[i]] == [s for _, _, s in want], i


def test_device_search_equals_the_host_driver_hook(cb, golden):
    """cb_debug_host_search with the network behind its callback == cb_debug_device_search (three-way closure)."""
    g, blob = golden("6x128_B")
    ev = make_eval(cb, g, blob, "bf16")
    boards = search_corpus(24)[:32]
    kids, _ = ev.debug_device_search(boards, sims_per_move=64, seed=3, material_on=True, qsearch="pe", tier1_light=False)

    def net(board, moves, idx, q, material_on):
        pri, _, vf, _ = ev.eval_leaves(board[None], np.array([q], dtype=np.float32), np.array(idx, dtype=np.uint16),
                                       np.array([0, len(moves)], dtype=np.uint32), material_on=material_on)
        return pri, float(vf[0])
    hk, _ = cb.debug_host_search(boards, net, sims_per_move=64, seed=3, material_on=True, qsearch="pe")
    assert hk == kids


def test_device_deep_gates_equal_the_host_form(cb, golden):
    """tier1_gates_run (the tree kernels' resumable any / all minimax over the warp move generator) against
    cb_tier1_gates (host/gates.hpp, itself equal to the oracle in tests/test_host.py): value and terminal distance of the
    leaf gate as self-play runs it (mate-in-5 by checks, hill in 3), of the adjudication search (exhaustive to mate-in-2)
    and of a run resumed every 3 positions."""
    g, blob = golden("2x128_B")
    ev = make_eval(cb, g, blob, "bf16")
    a, _, _, _ = O.random_positions(2718, 2500, 200)
    fens = ["6k1/5ppp/8/8/8/8/8/4R1K1 w - - 0 1", "4r1k1/8/8/8/8/8/5PPP/6K1 b - - 0 1", "7k/R7/5K2/8/8/8/8/8 w - - 0 1",
            "r1bqkbnr/pppp1ppp/2n5/4p2Q/2B1P3/8/PPPP1PPP/RNB1K1NR w KQkq - 4 4", "k7/1R6/K7/8/8/8/8/8 b - - 0 1",
            "6k1/5ppp/8/7Q/8/8/8/5RK1 w - - 0 1", "8/8/8/8/8/2K5/8/k7 w - - 0 1", "8/8/8/8/3k4/8/8/K7 w - - 0 1", "8/8/8/8/8/8/1K6/k7 w - - 0 1",
            "rnbqkbnr/1pppppp1/p6p/8/8/4P3/PPPPKPPP/RNBQ1BNR w kq - 0 3", "rnbqkbnr/pppp1pp1/4p2p/8/8/4P3/PPPPKPPP/RNBQ1BNR w kq - 0 3",
            "4k3/8/R7/1R6/8/8/8/6K1 w - - 0 1", "7k/8/5BK1/8/8/8/8/R7 w - - 0 1"]
    boards = np.concatenate([a, O.boards_to_array([O.from_fen(f) for f in fens])])
    for koth in (False, True):
        for depths in (dict(mate_depth=5, exhaustive_depth=0, koth_depth=3), dict(mate_depth=3, exhaustive_depth=2, koth_depth=3)):
            sub = boards if depths["exhaustive_depth"] == 0 else boards[-600:]
            want_v, want_d = cb.tier1_gates(sub, koth=koth, **depths)
            for budget in (0, 3):
                got_v, got_d = ev.debug_device_tier1_gates(sub, koth=koth, budget=budget, **depths)
                assert np.array_equal(got_v, want_v), (koth, depths, budget)
                hit = want_v != 2
                assert np.array_equal(got_d[hit], want_d[hit]), (koth, depths, budget)
            assert (want_v == 1).sum() >= 5


@pytest.mark.parametrize("koth", [False, True], ids=["standard", "koth"])
def test_device_self_play_with_the_deep_gates_equals_the_host_driver(cb, golden, tmp_path, monkeypatch, koth):
    """Self-play with the reference's Tier-1 configuration (leaf gate mate-in-5 by checks + hill in 3; every new root
    adjudicated by hill in 3 / mate_search(5, exhaustive 2)) on the device against the host driver.  With the gate jobs
    inline in the tree kernel and a budget no job exceeds, the device runs in lock step with the host: identical counters
    and sample files.  With the jobs in their own kernel (the default) and / or a small budget, games advance at different
    paces (suspended jobs, one step between an adjudication and the next game) and every game must still be the same game:
    the finished games' samples are a subset of the host run's."""
    g, blob = golden("6x128_B")
    ev = make_eval(cb, g, blob, "bf16")
    kw = dict(games=16, sims_per_move=20, max_plies=40, koth=koth, seed=11, tier1_light="deep", material_on=True, qsearch="pe",
              max_sims=16 * 20 * 60)
    monkeypatch.setenv("CB_GATE_BUDGET", "100000000")
    monkeypatch.setenv("CB_GATE_INLINE", "1")
    host = ev.selfplay(threads=1, pipeline=1, sample_path=str(tmp_path / "host.bin"), **kw)
    dev = ev.selfplay(pipeline=3, sample_path=str(tmp_path / "dev.bin"), **kw)
    for k in ("steps", "nn_evals", "terminal_hits", "moves", "games_finished", "samples", "white_wins", "black_wins", "draws"):
        assert host[k] == dev[k], (k, host[k], dev[k])
    assert dev["games_finished"] >= 16 and dev["overflow"] == 0 and dev["gate_waits"] == 0
    rh = np.fromfile(tmp_path / "host.bin", dtype=np.float32).reshape(-1, 5763)
    rd = np.fromfile(tmp_path / "dev.bin", dtype=np.float32).reshape(-1, 5763)
    key = lambda r: r[np.lexsort(r.T[::-1])]
    assert rh.shape == rd.shape and rh.shape[0] > 0 and np.array_equal(key(rh), key(rd))
    have = {r.tobytes() for r in rh}
    monkeypatch.delenv("CB_GATE_INLINE")
    # (CB_GATE_STRETCH: positions a job may go beyond its budget while fewer than half of the games are done with the step)
    # (CB_GATE_WARPS: warps per game in the gate kernel, each with every W-th move of the gate's root)
    for name, budget, stretch, warps in (("split", "100000000", "400", "1"), ("slow", "24" if koth else "3", "0", "1"),
                                         ("stretched", "2", "400", "1"), ("shared", "6", "0", "4")):
        monkeypatch.setenv("CB_GATE_BUDGET", budget)
        monkeypatch.setenv("CB_GATE_STRETCH", stretch)
        monkeypatch.setenv("CB_GATE_WARPS", warps)
        run = ev.selfplay(pipeline=3, sample_path=str(tmp_path / (name + ".bin")), **kw)
        assert run["games_finished"] > 0 and run["overflow"] == 0 and (name != "slow" or run["gate_waits"] > 0) and \
            (name != "split" or run["gate_waits"] == 0)
        rs = np.fromfile(tmp_path / (name + ".bin"), dtype=np.float32).reshape(-1, 5763)
        assert rs.shape[0] > 0 and all(r.tobytes() in have for r in rs), name
        if name == "split":
            assert rs.shape[0] >= 0.9 * rh.shape[0]


@pytest.mark.parametrize("material,qsearch", [(False, 0), (True, "pe")])
def test_two_pool_self_play_walks_the_same_trees(cb, golden, material, qsearch, tmp_path):
    """pipeline 4: the games split in two pools whose tree kernels run beside each other's forward kernel
    (programmatic dependent launch, forward on a bounded grid with a board offset).  Games are independent and
    the evaluator does not depend on batch composition, so every game must be the one pipeline 3 plays."""
    g, blob = golden("6x128_B")
    ev = make_eval(cb, g, blob, "bf16")
    kw = dict(games=48, sims_per_move=20, max_plies=40, material_on=material, qsearch=qsearch, seed=11, max_sims=48 * 20 * 60)
    one = ev.selfplay(pipeline=3, sample_path=str(tmp_path / "one.bin"), **kw)
    two = ev.selfplay(pipeline=4, sample_path=str(tmp_path / "two.bin"), **kw)
    for k in ("sims", "steps", "nn_evals", "terminal_hits", "moves", "games_finished", "samples", "white_wins",
              "black_wins", "draws"):
        assert one[k] == two[k], (k, one[k], two[k])
    assert two["games_finished"] >= 48
    ra = np.fromfile(tmp_path / "one.bin", dtype=np.float32).reshape(-1, 5763)
    rb = np.fromfile(tmp_path / "two.bin", dtype=np.float32).reshape(-1, 5763)
    key = lambda r: r[np.lexsort(r.T[::-1])]
    assert ra.shape == rb.shape and np.array_equal(key(ra), key(rb))
    with pytest.raises(cb.CaissaError):
        ev.selfplay(pipeline=4, games=47, sims_per_move=4, max_sims=47 * 8)


def test_device_move_generator_equals_the_host_list_move_for_move(cb, golden):
    """The warp form of the move generator (one piece per lane, scans for the order) against the host list —
    same moves in the same order, which is what keeps device trees identical to the host driver's."""
    g, blob = golden("2x128_B")
    ev = make_eval(cb, g, blob, "bf16")
    a, _, _, _ = O.random_positions(20261019, 3000, 200)
    fens = ["r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1",      # castling both ways
            "r3k2r/Pppp1ppp/1b3nbN/nP6/BBP1P3/q4N2/Pp1P2PP/R2Q1RK1 w kq - 0 1",           # promotions, captures
            "r3k2r/8/8/8/8/8/8/R3K2R b KQkq - 0 1", "r3k2r/8/8/8/8/5n2/8/R3K2R w KQkq - 0 1",   # castling through check
            "4k3/8/8/3pP3/8/8/8/4K3 w - d6 0 1", "8/P7/8/8/8/8/7p/K6k b - - 0 1", "4k3/8/8/8/8/8/8/4K3 w - - 0 1"]
    boards = np.concatenate([a, O.boards_to_array([O.from_fen(f) for f in fens])])
    moves, counts = ev.debug_device_legal_moves(boards)
    for i in range(boards.shape[0]):
        want, _ = cb.chess_legal_moves(boards[i])
        assert counts[i] == len(want) and np.array_equal(moves[i, :counts[i]], want), i


def test_device_light_gates_equal_the_host_form(cb, golden):
    """The warp form decides "gives check" from attack sets (exact test only for moves that can uncover a line, en
    passant, castling, promotions); the host form makes every move.  Same answers, both rule sets."""
    g, blob = golden("2x128_B")
    ev = make_eval(cb, g, blob, "bf16")
    a, _, _, _ = O.random_positions(31337, 4000, 200)
    fens = ["6k1/5ppp/8/8/8/8/8/4R1K1 w - - 0 1", "4r1k1/8/8/8/8/8/5PPP/6K1 b - - 0 1",
            "r1bqkbnr/pppp1ppp/2n5/4p2Q/2B1P3/8/PPPP1PPP/RNB1K1NR w KQkq - 4 4", "k7/1R6/K7/8/8/8/8/8 b - - 0 1",
            "6k1/5ppp/8/8/8/8/5PPP/3R2K1 w - - 0 1",                       # Rd8#
            "5rk1/5ppp/8/8/8/8/1B6/K5R1 w - - 0 1",                        # Rxg7+ is not mate; discovered lines
            "7k/5P2/6K1/8/8/8/8/8 w - - 0 1",                              # f8=Q# / f8=R#: promotion mates
            "4k3/8/8/8/8/8/8/R3K2R w KQ - 0 1",                            # castling gives no mate here
            "7k/6pp/8/3B4/8/8/6PP/5R1K w - - 0 1",                         # Rf8#
            "8/8/8/8/8/2K5/8/k7 w - - 0 1", "8/8/8/8/3k4/8/8/K7 w - - 0 1"]
    boards = np.concatenate([a, O.boards_to_array([O.from_fen(f) for f in fens])])
    for koth in (False, True):
        got = ev.debug_device_light_gates(boards, koth=koth)
        want = cb.light_gates(boards, koth=koth)
        assert np.array_equal(got, want), np.nonzero(got != want)[0][:10]
        assert (got == 1).sum() >= 8


def test_principal_exchange_on_device(cb, golden):
    """The warp form of the exchange line (one piece per lane) against the host form and the oracle, and the
    material scalar of a self-play sample file against the oracle."""
    g, blob = golden("2x128_B")
    ev = make_eval(cb, g, blob, "bf16")
    a, _, _, _ = O.random_positions(20261019, 3000, 200)
    b, _, _, _ = O.random_positions(5, 1100, 100)
    boards = np.concatenate([a, b])
    dev = ev.principal_exchange(boards)
    assert np.array_equal(dev, cb.principal_exchange(boards))
    assert np.array_equal(dev[:600], O.pe_q(boards[:600]))
    assert np.array_equal(ev.principal_exchange(boards[:1]), dev[:1]) and ev.principal_exchange(boards[:0]).shape == (0,)
    assert (dev != O.static_q(boards, clip=1e9)).mean() > 0.2      # the line does change dM on this corpus


def test_model_match_on_device(cb, golden):
    """Candidate-vs-current evaluation match (src/bin/evaluate_models.rs): accounting, determinism, SPRT trigger."""
    g, blob = golden("6x128_B")
    ga, blob_a = golden("6x128_A")
    cand = make_eval(cb, g, blob, "bf16")
    curr = make_eval(cb, ga, blob_a, "bf16")
    kw = dict(games=32, total_games=48, sims_per_move=16, max_plies=30, seed=5)
    a = cand.match(curr, **kw)
    b = cand.match(curr, **kw)
    assert a["games"] == 48 == a["candidate_wins"] + a["current_wins"] + a["draws"]
    for k in ("games", "candidate_wins", "current_wins", "draws"):
        assert a[k] == b[k]                                   # same seed, same match
    assert a["decision"] == 0 and a["sims"] > 48 * 16
    want = cb.sprt_llr(a["candidate_wins"], a["current_wins"], a["draws"])
    assert (want is None) == (a["llr_valid"] == 0) and (want is None or abs(want - a["llr"]) < 1e-12)
    # a model against itself with SPRT on and generous error rates: the test stops on a decision or plays all games
    s = cand.match(cand2 := make_eval(cb, g, blob, "bf16"), games=32, total_games=64, sims_per_move=8, max_plies=20, seed=9,
                   use_sprt=True, elo0=0.0, elo1=400.0, alpha=0.3, beta=0.3)
    assert s["games"] <= 64 and s["candidate_wins"] + s["current_wins"] + s["draws"] == s["games"]
    if s["decision"] != 0:
        assert cb.sprt_decision(s["candidate_wins"], s["current_wins"], s["draws"], 0.0, 400.0, 0.3, 0.3) == \
            {1: "H1", -1: "H0"}[s["decision"]]
    del cand2


def test_device_cap1_qsearch_equals_the_host_form(cb, golden, tmp_path, monkeypatch):
    """cap1_run (iterative alpha-beta over an explicit stack, lane d keeping ply d) against cb_cap1_qsearch (recursive
    host form, itself equal to the oracle in tests/test_host.py): score and completion flag on 5000 positions; then the
    device self-play with dM by cap1 against the host driver (same counters, same sample files incl. the completion flag)
    with a per-step budget no search exceeds, and with a budget of 5 positions (searches suspended and resumed: the finished
    games are the host's games)."""
    monkeypatch.setenv("CB_GATE_BUDGET", "100000000")
    g, blob = golden("6x128_B")
    ev = make_eval(cb, g, blob, "bf16")
    a, _, _, _ = O.random_positions(4711, 5000, 220)
    fens = ["r3k3/8/8/1N6/8/8/8/4K3 w - - 0 1", "6k1/5ppp/8/8/8/8/8/3RK3 w - - 0 1", "k7/8/8/8/8/8/1B6/R3K3 w - - 0 1",
            "7k/1p1b1Qp1/1r6/p4pq1/1b1Np2p/4P3/K1rNR1PP/3R4 w - - 0 36", "8/8/8/KPp4r/8/8/8/7k w - c6 0 2", "4k3/P7/8/8/8/8/8/4K3 w - - 0 1"]
    boards = np.concatenate([a, O.boards_to_array([O.from_fen(f) for f in fens])])
    want_q, want_c, nodes, _ = cb.cap1_qsearch(boards)
    got_q, got_c = ev.cap1_qsearch_device(boards)
    assert np.array_equal(got_q, want_q) and np.array_equal(got_c, want_c)
    assert (want_c == 0).sum() >= 1 and nodes.max() > 300 and (want_q != cb.principal_exchange(boards)).sum() > 1000
    kw = dict(games=24, sims_per_move=20, max_plies=40, material_on=True, qsearch="cap1", seed=11, max_sims=24 * 20 * 50)
    host = ev.selfplay(threads=1, pipeline=1, sample_path=str(tmp_path / "host.bin"), **kw)
    dev = ev.selfplay(pipeline=3, sample_path=str(tmp_path / "dev.bin"), **kw)
    # ("sims": with jobs the device counts completed simulations, the host driver steps x games)
    for k in ("steps", "nn_evals", "terminal_hits", "moves", "games_finished", "samples", "white_wins", "black_wins", "draws"):
        assert host[k] == dev[k], (k, host[k], dev[k])
    assert dev["sims"] == dev["nn_evals"] + dev["terminal_hits"] - dev["moves"] - 24 or dev["sims"] <= host["sims"]
    rh = np.fromfile(tmp_path / "host.bin", dtype=np.float32).reshape(-1, 5763)
    rd = np.fromfile(tmp_path / "dev.bin", dtype=np.float32).reshape(-1, 5763)
    key = lambda r: r[np.lexsort(r.T[::-1])]
    assert rh.shape == rd.shape and rh.shape[0] > 0 and np.array_equal(key(rh), key(rd))
    monkeypatch.setenv("CB_GATE_BUDGET", "5")
    monkeypatch.setenv("CB_GATE_STRETCH", "0")
    slow = ev.selfplay(pipeline=3, sample_path=str(tmp_path / "slow.bin"), **kw)
    rs = np.fromfile(tmp_path / "slow.bin", dtype=np.float32).reshape(-1, 5763)
    have = {r.tobytes() for r in rh}
    assert slow["gate_waits"] > 0 and rs.shape[0] > 0 and all(r.tobytes() in have for r in rs)
    # and under the gate at full depth (jobs in the gate kernel: the leaf's gate, then its cap1 search)
    deep = ev.selfplay(pipeline=3, sample_path=str(tmp_path / "deep.bin"), tier1_light="deep", **kw)
    hdeep = ev.selfplay(threads=1, pipeline=1, sample_path=str(tmp_path / "hdeep.bin"), tier1_light="deep", **kw)
    rdp = np.fromfile(tmp_path / "deep.bin", dtype=np.float32).reshape(-1, 5763)
    have = {r.tobytes() for r in np.fromfile(tmp_path / "hdeep.bin", dtype=np.float32).reshape(-1, 5763)}
    assert deep["gate_waits"] > 0 and hdeep["games_finished"] > 0 and rdp.shape[0] > 0 and all(r.tobytes() in have for r in rdp)


def test_model_match_with_the_deep_gates(cb, golden, monkeypatch):
    """The match runner with the Tier-1 gate at full depth (the reference's evaluate_models default): the accounting holds,
    and pacing (the per-step budget of the gate jobs) changes which games finish first, not what a game is - a model against
    itself under King-of-the-Hill rules, where the gate decides many games, scores the same with two budgets."""
    g, blob = golden("6x128_B")
    cand = make_eval(cb, g, blob, "bf16")
    curr = make_eval(cb, g, blob, "bf16")
    kw = dict(games=16, total_games=16, sims_per_move=16, max_plies=40, seed=5, tier1_light=2, koth=True)
    runs = []
    for budget in ("1000000", "8"):
        monkeypatch.setenv("CB_GATE_BUDGET", budget)
        r = cand.match(curr, **kw)
        assert r["games"] == 16 == r["candidate_wins"] + r["current_wins"] + r["draws"] and r["sims"] > 16 * 16
        runs.append((r["candidate_wins"], r["current_wins"], r["draws"]))
    assert runs[0] == runs[1]                 # 16 slots, 16 games: every slot plays exactly its first game
    light = cand.match(curr, **dict(kw, tier1_light=1))
    assert light["games"] == 16


def test_two_pool_self_play_with_the_deep_gates(cb, golden, tmp_path, monkeypatch):
    """pipeline 4 with tier1_light = 2 (gate jobs inline in the tree kernels that run beside the other pool's forward): the
    games are the one-pool run's games."""
    g, blob = golden("6x128_B")
    ev = make_eval(cb, g, blob, "bf16")
    kw = dict(games=16, sims_per_move=20, max_plies=30, seed=4, tier1_light="deep", max_sims=16 * 20 * 50)
    monkeypatch.setenv("CB_GATE_BUDGET", "6")
    one = ev.selfplay(pipeline=3, sample_path=str(tmp_path / "one.bin"), **kw)
    two = ev.selfplay(pipeline=4, sample_path=str(tmp_path / "two.bin"), **kw)
    r1 = np.fromfile(tmp_path / "one.bin", dtype=np.float32).reshape(-1, 5763)
    r2 = np.fromfile(tmp_path / "two.bin", dtype=np.float32).reshape(-1, 5763)
    assert one["games_finished"] > 0 and two["games_finished"] > 0 and one["overflow"] == two["overflow"] == 0
    a, b = {r.tobytes() for r in r1}, {r.tobytes() for r in r2}
    small, big = (a, b) if len(a) <= len(b) else (b, a)
    assert len(small) > 0 and len(small & big) >= 0.8 * len(small)     # the slower run's finished games are the faster run's


# ------------------------------------------------------------------ reference-shaped host interface
def test_neural_net_policy_and_inference_server(cb, golden):
    import threading
    g, blob = golden("6x128_B")
    nn = cb.NeuralNetPolicy(6, 128, "bf16")
    assert nn.load("/nonexistent/model.bin").startswith("Model file not found")
    assert not nn.is_available() and nn.predict(bytes(128), True, 0.0) is None
    assert nn.load(blob) is None and nn.is_available()
    policy, v_logit, k = nn.predict(g["boards"][0], True, float(g["q"][0]))
    assert policy.shape == (4672,) and abs(policy.sum() - 1.0) < 1e-4        # tests/unit/inference_server_tests.rs
    assert abs(v_logit - g["v_logit"][0]) < 0.05 and abs(k - g["k"][0]) < 1e-6
    server = cb.InferenceServer(nn, 16)
    out = [None] * 48

    def game(i):
        out[i] = server.predict_async(g["boards"][i], True, float(g["q"][i])).get(timeout=30)
    threads = [threading.Thread(target=game, args=(i,)) for i in range(48)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    server.shutdown()
    for i in range(48):
        assert out[i] is not None and abs(out[i][1] - g["v_logit"][i]) < 0.05
    t = nn.timing()
    assert t["positions"] >= 49 and t["calls"] < t["positions"]              # requests were batched
