"""Host side of the path: the product's chess rules (legal-move lists = the path's legal-move
mask) against the oracle and the reference's perft table; the C++ mirror of the reference's
evaluator interface; the self-play driver's invariants."""
import os
import subprocess

import numpy as np
import pytest

from oracle import pyoracle as O

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
START = "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1"
POS1 = "r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1"
POS2 = "8/2p5/3p4/KP5r/1R3p1k/8/4P1P1/8 w - - 0 1"
POS3 = "r3k2r/Pppp1ppp/1b3nbN/nP6/BBP1P3/q4N2/Pp1P2PP/R2Q1RK1 w kq - 0 1"
POS4 = "rnbq1k1r/pp1Pbppp/2p5/8/2B5/8/PPP1NnPP/RNBQK2R w KQ - 1 8"
POS5 = "r4rk1/1pp1qppp/p1np1n2/2b1p1B1/2B1P1b1/P1NP1N2/1PP1QPPP/R4RK1 w - - 0 10"


def u8(fen):
    return O.boards_to_array([O.from_fen(fen)])[0]


@pytest.mark.parametrize("fen,depth,nodes", [  # tests/perft_tests.rs:68-291
    (START, 4, 197281), (POS1, 3, 97862), (POS2, 5, 674624), (POS3, 4, 422333), (POS4, 3, 62379), (POS5, 3, 89890)])
def test_product_chess_perft(cb, fen, depth, nodes):
    assert cb.chess_perft(u8(fen), depth) == nodes


def test_product_legal_lists_equal_oracle(cb):
    assert np.array_equal(cb.chess_startpos(), u8(START))
    boards, idx, off, raw = O.random_positions(4242, 1500)
    for i in range(1500):
        mv, pi = cb.chess_legal_moves(boards[i])
        want = sorted(zip(raw[off[i]:off[i + 1]].tolist(), idx[off[i]:off[i + 1]].tolist()))
        assert sorted(zip(mv.tolist(), pi.tolist())) == want          # same set, same policy indices: bit-exact mask
        if len(mv):
            m = int(mv[i % len(mv)])
            assert np.array_equal(cb.chess_make_move(boards[i], m),
                                  O.boards_to_array([O.make_move(O.array_to_board(boards[i]), m)])[0])


PE_FENS = [START, POS1, POS2, POS3, POS4, POS5,
           "4k3/8/8/8/4n3/3P4/8/4K3 w - - 0 1", "4k3/8/8/3p4/2p5/8/8/3QK3 w - - 0 1",       # ext_qsearch_tests.rs:321,333
           "4k3/8/8/8/8/8/8/r3K2R w K - 0 1", "1n2k3/2P5/8/8/8/8/8/4K3 w - - 0 1",          # :345,:561
           "r3k3/8/8/1N6/8/8/8/4K3 w - - 0 1", "4k3/8/8/8/8/8/8/4K3 w - - 0 1",             # :368,:550
           "rnbqkb1r/pppp1ppp/8/4p3/2B1N3/8/PPPP1PPP/R1BQK1NR b KQkq - 0 4",                # :421
           "r1b2rk1/pp2nppp/2pBp3/q7/1nP4P/5BN1/P4P2/R2QK2R w KQ - 0 15",                   # :425
           "7k/1p1b1Qp1/1r6/p4pq1/1b1Np2p/4P3/K1rNR1PP/3R4 w - - 0 36",                     # :402
           "4k3/8/8/3pP3/8/8/8/4K3 w - d6 0 1",                                             # en passant is the only capture
           "4r1k1/8/8/8/8/8/4R3/4K3 b - - 0 1", "4r2k/8/8/8/7b/8/4R3/4K3 w - - 0 1",        # pinned capturer
           "3r2k1/8/8/8/8/8/3p4/4K3 w - - 0 1", "1r2k3/P7/8/8/8/8/8/4K3 w - - 0 1"]          # defended piece next to the king; promotion choices


def test_principal_exchange_equals_the_oracle(cb):
    """Product (key order, no list) against oracle/qsearch.c (list + stable sort): score and positions visited."""
    a, _, _, _ = O.random_positions(20261019, 2500, 200)
    b, _, _, _ = O.random_positions(99, 1500, 120)
    boards = np.concatenate([a, b, np.stack([u8(f) for f in PE_FENS])])
    q, nodes = cb.principal_exchange(boards, want_nodes=True)
    first_illegal = deep = 0
    for i in range(boards.shape[0]):
        ob = O.array_to_board(boards[i])
        want_q, want_nodes = O.forced_principal_exchange(ob)
        assert q[i] == np.float32(want_q) and nodes[i] == want_nodes, i
        caps, long_list = O.captures_mvvlva(ob)
        assert not long_list
        first_illegal += bool(caps) and caps[0] not in O.legal_moves(ob)
        deep += want_nodes >= 4
    assert first_illegal > 20 and deep > 200     # the corpus exercises the retry and real exchanges
    assert cb.principal_exchange(boards[:0]).shape == (0,)


def test_cap1_qsearch_equals_the_oracle(cb):
    """cb_cap1_qsearch (the capture by MVV-LVA keys, no list, no sort) against oracle/qsearch.c's list-and-sort statement of
    src/search/quiescence.rs:808-1012: score, completion flag, node count and depth of every position."""
    a, _, _, _ = O.random_positions(4711, 4000, 220)
    boards = np.concatenate([a, np.stack([u8(f) for f in PE_FENS + [
        "r3k3/8/8/1N6/8/8/8/4K3 w - - 0 1", "6k1/5ppp/8/8/8/8/8/3RK3 w - - 0 1", "k7/8/8/8/8/8/1B6/R3K3 w - - 0 1",
        "7k/1p1b1Qp1/1r6/p4pq1/1b1Np2p/4P3/K1rNR1PP/3R4 w - - 0 36"]])])
    q, completed, nodes, depth = cb.cap1_qsearch(boards)
    for i in range(boards.shape[0]):
        w = O.forced_cap1(O.array_to_board(boards[i]))
        assert (q[i], bool(completed[i]), int(nodes[i]), int(depth[i])) == (np.float32(w[0]), w[1], w[2], w[3]), i
    assert (q != cb.principal_exchange(boards)).sum() > 800 and nodes.max() > 300     # the checks do change dM on this corpus


def test_light_gates_equal_the_oracle(cb):
    """cb_light_gates (legal list + early-exit reply search) against the oracle's plain statement, both rule sets."""
    a, _, _, _ = O.random_positions(31337, 3000, 200)
    boards = np.concatenate([a, np.stack([u8(f) for f in PE_FENS + [
        "6k1/5ppp/8/8/8/8/8/4R1K1 w - - 0 1", "4r1k1/8/8/8/8/8/5PPP/6K1 b - - 0 1",
        "r1bqkbnr/pppp1ppp/2n5/4p2Q/2B1P3/8/PPPP1PPP/RNB1K1NR w KQkq - 4 4", "k7/1R6/K7/8/8/8/8/8 b - - 0 1",
        "8/8/8/8/8/2K5/8/k7 w - - 0 1", "8/8/8/8/3k4/8/8/K7 w - - 0 1", "8/8/8/2k5/8/2K5/8/8 w - - 0 1"]])])
    for koth in (False, True):
        got = cb.light_gates(boards, koth=koth)
        want = np.array([O.light_gates(O.array_to_board(r), koth) for r in boards], dtype=np.int8)
        assert np.array_equal(got, want)
        assert (got == 1).sum() >= 5
    assert (cb.light_gates(boards, koth=True) == -1).sum() >= 1


def test_move_flags_equal_make_move(cb):
    """cbchess::move_flags_hd (legality and "gives check" of a pseudo-legal move without making it: what the tree kernels'
    gate searches run per move) against make_move_hd + in_check_hd on every pseudo-legal move of > 20,000 positions:
    pins, evasions, en passant (including the rank pin through both pawns), castling, promotions, discovered checks."""
    special = ["r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1", "8/2p5/3p4/KP5r/1R3p1k/8/4P1P1/8 w - - 0 1",
               "r3k2r/Pppp1ppp/1b3nbN/nP6/BBP1P3/q4N2/Pp1P2PP/R2Q1RK1 w kq - 0 1", "rnbq1k1r/pp1Pbppp/2p5/8/2B5/8/PPP1NnPP/RNBQK2R w KQ - 1 8",
               "8/8/8/KPp4r/8/8/8/7k w - c6 0 2", "4k3/8/8/3pP3/8/8/8/4K3 w - d6 0 1", "8/8/8/8/k2Pp2R/8/8/4K3 b - d3 0 1",
               "r3k2r/8/8/8/8/8/8/R3K2R w KQkq - 0 1", "r3k2r/8/8/8/8/8/8/R3K2R b KQkq - 0 1", "4k3/P7/8/8/8/8/8/4K3 w - - 0 1",
               "3k4/P7/8/8/8/8/8/4K3 w - - 0 1", "8/8/8/8/8/5k2/4p3/4K3 b - - 0 1", "k7/8/8/8/8/8/1p6/R3K3 b - - 0 1",
               "5k2/8/8/8/8/8/8/4K2R w K - 0 1", "8/8/8/2k5/3Pp3/8/8/4K2B b - d3 0 1"]
    total = 0
    for seed in range(9):
        a, _, _, _ = O.random_positions(1000 + seed, 2500, 220)
        bad, n = cb.debug_move_flags(np.concatenate([a, np.stack([u8(f) for f in special])]))
        assert bad == 0
        total += n
    assert total > 600_000


def test_deep_gates_equal_the_oracle(cb):
    """The Tier-1 gates at full depth on the host (csrc/host/gates.hpp: mate search to mate-in-5, hill in 3) against
    oracle/gates.c, which is pinned to the reference's own tests: value and terminal distance of the leaf gate as self-play
    runs it (checking moves only), the adjudication search (exhaustive to mate-in-2) and the hill distances."""
    a, _, _, _ = O.random_positions(4242, 700, 200)
    boards = np.concatenate([a, np.stack([u8(f) for f in PE_FENS + [
        "6k1/5ppp/8/8/8/8/8/4R1K1 w - - 0 1", "4r1k1/8/8/8/8/8/5PPP/6K1 b - - 0 1", "7k/R7/5K2/8/8/8/8/8 w - - 0 1",
        "r1bqkbnr/pppp1ppp/2n5/4p2Q/2B1P3/8/PPPP1PPP/RNB1K1NR w KQkq - 4 4", "k7/1R6/K7/8/8/8/8/8 b - - 0 1",
        "6k1/5ppp/8/7Q/8/8/8/5RK1 w - - 0 1", "8/8/8/8/8/2K5/8/k7 w - - 0 1", "8/8/8/8/3k4/8/8/K7 w - - 0 1", "8/8/8/8/8/8/1K6/k7 w - - 0 1",
        "rnbqkbnr/1pppppp1/p6p/8/8/4P3/PPPPKPPP/RNBQ1BNR w kq - 0 3", "rnbqkbnr/pppp1pp1/4p2p/8/8/4P3/PPPPKPPP/RNBQ1BNR w kq - 0 3"]])])
    obs = [O.array_to_board(r) for r in boards]
    for koth in (False, True):
        got_v, got_d = cb.tier1_gates(boards, koth=koth)                                   # (5, 0, 3): self-play's leaf gate
        want = [O.tier1_gates(b, koth=koth) for b in obs]
        assert np.array_equal(got_v, np.array([w[0] for w in want], dtype=np.int8))
        hit = got_v != 2
        assert np.array_equal(got_d[hit], np.array([w[1] for w in want], dtype=np.uint8)[hit])
        assert (got_v == 1).sum() >= 8
    deep = cb.mate_search(boards, 5, 0)
    assert np.array_equal(deep, np.array([max(0, O.mate_search(b, 5, 0)[0] - 1_000_000) for b in obs], dtype=np.int32))
    assert (deep >= 3).sum() >= 3                                                          # mates longer than one move are in the corpus
    sub = boards[:200]
    adj = cb.mate_search(sub, 3, 2)                                                        # exhaustive levels (costly on the oracle side)
    assert np.array_equal(adj, np.array([max(0, O.mate_search(b, 3, 2)[0] - 1_000_000) for b in obs[:200]], dtype=np.int32))
    hill = cb.koth_center_in_n(boards, 3)
    want_h = np.array([-1 if O.koth_center_in_n(b, 3)[0] is None else O.koth_center_in_n(b, 3)[0] for b in obs], dtype=np.int8)
    assert np.array_equal(hill, want_h)
    assert set(np.unique(hill)) >= {-1, 0, 1, 2}


@pytest.fixture(scope="module")
def host_test_exe(cb):
    exe = os.path.join(REPO, "tests", "cpp", "_host_mirror_test")
    lib_dir = os.path.dirname(cb.lib_path())
    subprocess.check_call(["g++", "-std=c++17", "-O1", os.path.join(REPO, "tests", "cpp", "host_mirror_test.cpp"),
                           "-o", exe, "-L", lib_dir, "-lcaissa_b200", "-Wl,-rpath," + lib_dir, "-lpthread"])
    yield exe
    os.remove(exe)


def test_cpp_mirror_mock_servers(host_test_exe):
    # tests/unit/inference_server_tests.rs: reply shape and (policy, v_logit, k) semantics
    p = subprocess.run([host_test_exe, "mock"], capture_output=True, text=True, timeout=60)
    assert p.returncode == 0, p.stderr
    assert "mock ok" in p.stdout


@pytest.mark.gpu
def test_cpp_mirror_on_gpu(cb, golden, host_test_exe, tmp_path):
    g, blob = golden("6x128_B")
    n = 48
    blob.tofile(tmp_path / "w.bin")
    g["boards"][:n].tofile(tmp_path / "boards.bin")
    p = subprocess.run([host_test_exe, "gpu", str(tmp_path / "w.bin"), str(tmp_path / "boards.bin"), str(n)],
                       capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stderr
    rows = [l.split() for l in p.stdout.splitlines() if l and l[0].isdigit()]
    assert len(rows) == n
    ev = cb.Evaluator(6, 128, "bf16")
    ev.load_weights(blob)
    _, v_ref, k_ref = ev.predict_batch(g["boards"][:n], np.full(n, 0.25, dtype=np.float32), want_policy=False)
    for i, r in enumerate(rows):
        v_server, v_direct, psum, k = float(r[1]), float(r[2]), float(r[3]), float(r[4])
        assert v_server == v_direct == pytest.approx(float(v_ref[i]), abs=1e-6)   # batching does not change results
        assert abs(psum - 1.0) < 1e-3 and abs(k - k_ref[i]) < 1e-6


@pytest.mark.gpu
def test_selfplay_driver(cb, golden, tmp_path):
    g, blob = golden("6x128_B")
    ev = cb.Evaluator(6, 128, "bf16")
    ev.load_weights(blob)
    path = str(tmp_path / "games.bin")
    st = ev.selfplay(games=64, sims_per_move=24, max_plies=30, material_on=True, seed=7, max_games=40, threads=4,
                     sample_path=path, max_seconds=120)
    assert st["games_finished"] >= 40
    assert st["sims"] == st["steps"] * 64 == st["nn_evals"] + st["terminal_hits"]   # one leaf per game per step
    assert 40 < st["mean_batch"] <= 64
    assert st["white_wins"] + st["black_wins"] + st["draws"] == st["games_finished"]
    rec = np.fromfile(path, dtype=np.float32).reshape(-1, 5763)                     # src/training_data.rs:20-60
    assert rec.shape[0] == st["samples"] > 0
    planes, material, flag, z, pol = rec[:, :1088], rec[:, 1088], rec[:, 1089], rec[:, 1090], rec[:, 1091:]
    assert set(np.unique(planes)) <= {0.0, 1.0} and (planes[:, 5 * 64:6 * 64].sum(axis=1) == 1).all()  # one STM king
    assert (flag == 1.0).all() and set(np.unique(z)) <= {-1.0, 0.0, 1.0}
    np.testing.assert_allclose(pol.sum(axis=1), 1.0, atol=1e-4)
    # first sample of a game is the start position: its planes are the oracle's, its policy lives on legal moves
    start_planes = O.board_to_planes(O.from_fen(START))
    first = np.where((planes == start_planes).all(axis=1))[0]
    assert len(first) >= 1
    legal = {O.move_to_index_stm(m, True) for m in O.legal_moves(O.from_fen(START))}
    assert set(np.nonzero(pol[first[0]])[0].tolist()) <= legal
    assert abs(material[first[0]]) < 1e-6
    # determinism of the whole pipeline for a fixed seed and a single host thread
    a = ev.selfplay(games=16, sims_per_move=8, max_plies=12, seed=3, max_sims=16 * 8 * 12, threads=1)
    b = ev.selfplay(games=16, sims_per_move=8, max_plies=12, seed=3, max_sims=16 * 8 * 12, threads=1)
    for k in ("sims", "nn_evals", "terminal_hits", "moves", "games_finished"):
        assert a[k] == b[k]


# ------------------------------------------------------------------ SPRT gate (src/mcts/sprt.rs)
def test_sprt_matches_the_reference_known_answers(cb):
    """The cases of tests/unit/evaluate_models_tests.rs:494-680 (elo0 0, elo1 10, alpha = beta = 0.05) and the
    closed form of src/mcts/sprt.rs:70-100 recomputed here."""
    import math
    assert cb.sprt_llr(60, 30, 10) > 0.0                       # test_sprt_llr_winning
    assert cb.sprt_llr(30, 60, 10) < 0.0                       # test_sprt_llr_losing
    assert cb.sprt_llr(0, 0, 100) is None                      # all draws: zero variance
    assert abs(cb.sprt_llr(2, 2, 96)) < 2.0
    assert cb.sprt_llr(1, 0, 0) is None                        # fewer than 2 games
    assert cb.sprt_decision(80, 10, 10) == "H1"
    assert cb.sprt_decision(10, 80, 10) == "H0"
    assert cb.sprt_decision(26, 24, 50) == "inconclusive"
    upper, lower = math.log(0.95 / 0.05), math.log(0.05 / 0.95)
    # the in-flight dilution regression: 21W/4L/137D triggers H1, 24W/9L/156D no longer does
    assert cb.sprt_llr(21, 4, 137) >= upper and cb.sprt_decision(21, 4, 137) == "H1"
    assert cb.sprt_llr(24, 9, 156) < upper and cb.sprt_decision(24, 9, 156) == "inconclusive"
    for w, l, d in [(60, 30, 10), (21, 4, 137), (3, 5, 2), (500, 480, 1020)]:
        n = w + l + d
        s0, s1 = 1 / (1 + 10 ** (-0.0 / 400)), 1 / (1 + 10 ** (-10.0 / 400))
        s = (w + 0.5 * d) / n
        var = (w + d / 4) / n - s * s
        want = (s1 - s0) * (2 * s - s0 - s1) / (2 * var / n)
        assert abs(cb.sprt_llr(w, l, d) - want) <= 1e-12 * max(1.0, abs(want))
    assert lower < 0 < upper
