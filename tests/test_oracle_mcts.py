"""The tree-operations oracle (oracle/mcts.c) against what the reference's own tests hold for this part of the path.

The Rust crate cannot be built here (no toolchain), and the reference's tests for the search hold behaviours and
inequalities rather than exact visit counts; every one that does not need the deep (depth > 1) gates or the wider
q-searches is replayed below, citing the test it restates.  The generation ORDER has no test in the reference at all:
it is restated from src/move_generation.rs:287-328,461-1000 and checked here against a hand derivation from that code
(Kiwipete) plus rule pins (perft through the new generator, set equality with the perft-pinned generator)."""
import numpy as np
import pytest

from oracle import pyoracle as O

START = "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1"
KIWI = "r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1"
CLASSICAL = ("mock", 0.0, 0.326)     # no network: uniform priors, tanh(0.326 * dM) — src/mcts/tactical_mcts.rs:34,776-785
UNIFORM = ("mock", 0.0, 0.5)         # InferenceServer::new_mock_uniform, src/mcts/inference_server.rs:103-122


def ucis(moves):
    return [O.uci(m) for m in moves]


# ------------------------------------------------------------------ generation order
def test_start_position_order():
    mv, nt = O.gen_legal_ref(O.from_fen(START))
    assert nt == 0
    # src/magic_bitboard.rs:248-265: the double push precedes the single push; :160-189: knight targets +15, +17, ...
    assert ucis(mv) == ["a2a4", "a2a3", "b2b4", "b2b3", "c2c4", "c2c3", "d2d4", "d2d3", "e2e4", "e2e3", "f2f4", "f2f3",
                        "g2g4", "g2g3", "h2h4", "h2h3", "b1a3", "b1c3", "g1f3", "g1h3"]


def test_kiwipete_order_hand_derived():
    """Derived by hand from src/move_generation.rs (see the module docstring): captures by piece class (pawns, knights,
    bishops, rooks, queens with the rook set first, king), promotions last; quiets with castling before the king steps."""
    mv, nt = O.gen_legal_ref(O.from_fen(KIWI))
    assert nt == 8 and len(mv) == 48
    assert ucis(mv[:8]) == ["g2h3", "d5e6", "e5d7", "e5f7", "e5g6", "e2a6", "f3h3", "f3f6"]
    assert ucis(mv[8:]) == ["a2a4", "a2a3", "b2b3", "g2g4", "g2g3", "d5d6",
                            "c3b5", "c3a4", "c3b1", "c3d1", "e5c6", "e5c4", "e5g4", "e5d3",
                            "d2c1", "d2e3", "d2f4", "d2g5", "d2h6", "e2d1", "e2f1", "e2d3", "e2c4", "e2b5",
                            "a1b1", "a1c1", "a1d1", "h1f1", "h1g1",
                            "f3d3", "f3e3", "f3g3", "f3f4", "f3f5", "f3g4", "f3h5",
                            "e1g1", "e1c1", "e1d1", "e1f1"]


def test_promotions_close_the_capture_list_and_ep_sits_in_target_order():
    # capturing and pushing promotions (Q, R, N, B each, src/magic_bitboard.rs:457-474) come after the king's captures
    mv, nt = O.gen_legal_ref(O.from_fen("1n2k3/P1P5/8/8/8/8/8/4K3 w - - 0 1"))
    assert ucis(mv[:nt]) == ["a7b8q", "a7b8r", "a7b8n", "a7b8b", "a7a8q", "a7a8r", "a7a8n", "a7a8b",
                             "c7b8q", "c7b8r", "c7b8n", "c7b8b", "c7c8q", "c7c8r", "c7c8n", "c7c8b"]
    # en passant is tested inside the pawn's capture loop (src/move_generation.rs:472-489): a-file side target first
    mv, nt = O.gen_legal_ref(O.from_fen("4k3/8/8/3pPp2/8/8/8/4K3 w - d6 0 1"))
    assert ucis(mv[:nt]) == ["e5d6"]
    mv, nt = O.gen_legal_ref(O.from_fen("4k3/8/5n2/3pP3/8/8/8/4K3 w - d6 0 1"))
    assert ucis(mv[:nt]) == ["e5d6", "e5f6"]
    # black: tables run towards lower squares, src/magic_bitboard.rs:214-229
    mv, nt = O.gen_legal_ref(O.from_fen("4k3/8/8/8/3pP3/2N5/8/4K3 b - e3 0 1"))
    assert ucis(mv[:nt]) == ["d4c3", "d4e3"]


def test_ref_order_generator_is_a_permutation_of_the_perft_pinned_one():
    boards, _, _, _ = O.random_positions(20261019, 2500, 200)
    n_tact = 0
    for i in range(boards.shape[0]):
        b = O.array_to_board(boards[i])
        mv, nt = O.gen_legal_ref(b)
        assert sorted(mv) == sorted(O.legal_moves(b)), i
        them = 0 if not b.w_to_move else 1
        for j, m in enumerate(mv):     # the first nt moves are exactly the captures, en-passant captures and promotions
            to, frm, promo = (m >> 6) & 63, m & 63, (m >> 12) & 7
            is_cap = bool(b.occ[them] >> to & 1) or promo != 0 or \
                (bool(b.pieces[(1 - them) * 6] >> frm & 1) and b.en_passant == to)
            assert is_cap == (j < nt), (i, O.uci(m))
        n_tact += nt
    assert n_tact > 3000


@pytest.mark.parametrize("fen,depth,want", [
    (START, 3, 8902),                                                          # tests/perft_tests.rs
    (KIWI, 2, 2039),
    ("8/2p5/3p4/KP5r/1R3p1k/8/4P1P1/8 w - - 0 1", 3, 2812),
    ("r3k2r/Pppp1ppp/1b3nbN/nP6/BBP1P3/q4N2/Pp1P2PP/R2Q1RK1 w kq - 0 1", 2, 264),
    ("rnbq1k1r/pp1Pbppp/2p5/8/2B5/8/PPP1NnPP/RNBQK2R w KQ - 1 8", 2, 1486)])
def test_perft_through_the_ref_order_generator(fen, depth, want):
    def perft(b, d):
        mv, _ = O.gen_legal_ref(b)
        if d == 1:
            return len(mv)
        return sum(perft(O.make_move(b, m), d - 1) for m in mv)
    assert perft(O.from_fen(fen), depth) == want


# ------------------------------------------------------------------ tactical list (src/mcts/tactical.rs)
def test_tactical_known_answers():
    assert O.tactical_moves(O.from_fen(START)) == []                                      # tactical.rs:233-242
    b = O.from_fen("rnbqkbnr/pppp1ppp/8/4p3/4P3/8/PPPP1PPP/RNBQKBNR w KQkq - 0 2")
    assert O.lib().orc_tactical_score(b, O.mv("e4e5")) == 9.0                             # :275-284 (the test's literal move)
    b = O.from_fen("4k3/P7/8/8/8/8/8/4K3 w - - 0 1")
    assert O.lib().orc_tactical_score(b, O.mv("a7a8q")) == 8.0                            # :287-309
    assert O.lib().orc_tactical_score(b, O.mv("a7a8n")) == 2.0
    t = O.tactical_moves(O.from_fen("4k3/8/8/3p4/8/5B2/8/4K3 w - - 0 1"))                 # :312-327: no SEE filter
    assert [O.uci(m) for m, _ in t] == ["f3d5"]
    t = O.tactical_moves(O.from_fen("4k3/8/3q4/4P3/8/8/8/4K3 w - - 0 1"))                 # tactical_detection_tests.rs:84-107
    assert [O.uci(m) for m, _ in t] == ["e5d6"] and t[0][1] == 89.0
    # descending, ties in generation order (stable sort_by): Kiwipete
    t = O.tactical_moves(O.from_fen(KIWI))
    assert [s for _, s in t] == sorted([s for _, s in t], reverse=True)
    assert [O.uci(m) for m, _ in t] == ["e2a6", "f3f6", "g2h3", "d5e6", "e5d7", "e5f7", "e5g6", "f3h3"]


# ------------------------------------------------------------------ search behaviour
def root_q(t):
    s = t.root_stats()
    return -(s["total_value"] / s["visits"]) if s["visits"] else 0.0


def test_every_root_child_is_visited_and_counts_add_up():
    t = O.Tree(O.from_fen(START))
    assert t.search(O.mcts_config(sims=50), UNIFORM) == 50
    ch = t.root_children()
    # selection.rs:65-73 (every root child once), then strict '>' keeps the first maximum among equals
    assert [v for _, v, _ in ch] == [3] * 10 + [2] * 10
    s = t.root_stats()
    assert s["visits"] == 50 and s["evals"] == 51 and s["tactical"] == 0      # + the root's own policy request (selection.rs:315-332)
    assert t.best_move() == O.mv("a2a4")                                       # tactical_mcts.rs:896-905: first strict maximum


def test_tactical_children_are_visited_first_in_mvv_lva_order():
    """selection.rs:37-63,107-149: before any PUCT step a node visits each capture / promotion child once, best
    MVV-LVA first — at the root this precedes the force-visit of the remaining children."""
    b = O.from_fen(KIWI)
    want = [m for m, _ in O.tactical_moves(b)]
    for k in range(1, 9):
        t = O.Tree(b)
        t.search(O.mcts_config(sims=k), UNIFORM)
        visited = [m for m, v, _ in t.root_children() if v > 0]
        assert sorted(visited) == sorted(want[:k]), k
    t = O.Tree(b)
    t.search(O.mcts_config(sims=48), UNIFORM)
    assert all(v == 1 for _, v, _ in t.root_children())
    assert t.root_stats()["tactical"] >= 8 and t.root_stats()["evals"] == 48       # no PUCT step yet: no root policy request


def test_backup_signs_and_visit_counts():
    # node_tests.rs:309-411 through the search: one simulation from the start position with a mock value of tanh(0.7)
    import math
    t = O.Tree(O.from_fen(START))
    v = math.atanh(0.7)
    t.search(O.mcts_config(sims=1), ("mock", v, 0.5))
    (m, vis, tv), s = t.root_children()[0], t.root_stats()
    assert vis == 1 and s["visits"] == 1
    assert abs(tv - (-0.7)) < 1e-7 and abs(s["total_value"] - 0.7) < 1e-7
    t.search(O.mcts_config(sims=500), ("mock", v, 0.5))
    s = t.root_stats()
    assert s["visits"] == 501 and sum(x for _, x, _ in t.root_children()) == 501


def test_q_value_sign_convention():
    # tactical_mcts_tests.rs:220-255: after 1.f4, e7e5 loses a pawn to fxe5 => its Q (black's view) is negative
    t = O.Tree(O.from_fen("rnbqkbnr/pppppppp/8/8/5P2/8/PPPPP1PP/RNBQKBNR b KQkq f3 0 1"))
    t.search(O.mcts_config(sims=50, material_on=True, qsearch="pe"), CLASSICAL)
    e7e5 = [(v, tv) for m, v, tv in t.root_children() if m == O.mv("e7e5")][0]
    assert e7e5[0] > 0 and e7e5[1] / e7e5[0] < 0.0


def test_classical_fallback_material():
    # tactical_mcts_tests.rs:551-627
    for fen, check in [("4k3/8/8/3Q4/8/8/8/4K3 w - - 0 1", lambda q: q > 0.0),
                       ("3qk3/8/8/8/8/8/8/4K3 w - - 0 1", lambda q: q < 0.0),
                       ("4k3/pppppppp/8/8/8/8/PPPPPPPP/4K3 w - - 0 1", lambda q: abs(q) < 0.5)]:
        t = O.Tree(O.from_fen(fen))
        t.search(O.mcts_config(sims=50, material_on=True, qsearch="pe"), CLASSICAL)
        assert check(root_q(t)), (fen, root_q(t))


def test_mock_uniform_matches_classical_and_bias_changes_evaluation():
    # tactical_mcts_tests.rs:980-1049
    fen = "4k3/8/8/3Q4/8/8/8/4K3 w - - 0 1"
    a, b = O.Tree(O.from_fen(fen)), O.Tree(O.from_fen(fen))
    a.search(O.mcts_config(sims=100, material_on=True, qsearch="pe"), CLASSICAL)
    b.search(O.mcts_config(sims=100, material_on=True, qsearch="pe"), UNIFORM)
    assert a.best_move() == b.best_move() and abs(root_q(a) - root_q(b)) < 0.15 and b.root_stats()["evals"] > 0
    # :1051-1110
    fen = "4k3/pppppppp/8/8/8/8/PPPPPPPP/4K3 w - - 0 1"
    p, n = O.Tree(O.from_fen(fen)), O.Tree(O.from_fen(fen))
    p.search(O.mcts_config(sims=50, material_on=True, qsearch="pe"), ("mock", 2.0, 0.5))
    n.search(O.mcts_config(sims=50, material_on=True, qsearch="pe"), ("mock", -2.0, 0.5))
    assert abs(root_q(p) - root_q(n)) > 0.05


def test_material_value_toggle():
    # tactical_mcts_tests.rs:1190-1258 (classical fallback without material = 0 for every leaf: k = 0, v_logit = 0)
    fen = "4k3/8/8/3Q4/8/8/8/4K3 w - - 0 1"
    on, off = O.Tree(O.from_fen(fen)), O.Tree(O.from_fen(fen))
    on.search(O.mcts_config(sims=50, material_on=True, qsearch="pe"), CLASSICAL)
    off.search(O.mcts_config(sims=50, material_on=False), ("mock", 0.0, 0.0))
    assert root_q(on) > 0.3 and abs(root_q(off)) < 0.01


def test_terminal_positions_and_single_legal_move():
    # tactical_mcts_tests.rs:133-170: stalemate / checkmate roots return no move
    for fen in ["k7/8/1Q6/8/8/8/8/K7 b - - 0 1", "k7/1Q6/1K6/8/8/8/8/8 b - - 0 1"]:
        t = O.Tree(O.from_fen(fen))
        assert t.search(O.mcts_config(sims=20), UNIFORM) == 0 and t.best_move() == 0
    # :1355-1366
    t = O.Tree(O.from_fen("k7/8/1K6/8/8/8/8/8 b - - 0 1"))
    t.search(O.mcts_config(sims=50), UNIFORM)
    assert t.best_move() != 0


KOTH_FEN = "r1b2bnr/p2q1p1p/3kP1p1/2p5/p7/N4N2/PP1PPPPP/R1BK1B1R w - - 0 10"   # tactical_mcts_tests.rs:468-475


def test_gate_resolved_nodes_and_proven_losses():
    """tactical_mcts_tests.rs:317-400 with the gate at depth 1 (their position needs the depth-3 gate; this one, from
    :468-475, has the hill one step away): children where Black steps onto the hill at once are proven losses — one
    visit each (tactical phase or the root's force-visit), never selected by PUCT (selection.rs:178-191)."""
    t = O.Tree(O.from_fen(KOTH_FEN))
    t.search(O.mcts_config(sims=200, material_on=True, qsearch="pe", koth=True, tier1_light=True), CLASSICAL)
    proven, others = 0, 0
    for m, v, tv in t.root_children():
        nb = O.make_move(O.from_fen(KOTH_FEN), m)
        if O.light_gates(nb, koth=True) == 1:
            proven += 1
            assert v == 1 and tv == -1.0, O.uci(m)       # Q consistent with the gate value, :259-313
        elif v > 0:
            others += 1
    assert proven > 0 and others > 0


def test_deep_koth_gate_values_are_not_diluted_and_gated_nodes_stay_unexpanded():
    """tactical_mcts_tests.rs:258-352 on the reference's own position, which needs the gate at full depth (after a White move
    that does not attend to it Black forces the hill in 2): every root child with a terminal_or_mate_value keeps
    Q = -value however often it is visited, has no children, and there are such children.  (The reference runs the
    test with its default q-search, `extended`, for the classical leaf values; the assertions do not depend on them.)"""
    fen = "r1b4r/ppp1k2p/n3pq1n/5pp1/8/1PP1K3/P3PPPP/RNQ2BNR w - - 0 11"
    for sims in (100, 1000):
        t = O.Tree(O.from_fen(fen))
        t.search(O.mcts_config(sims=sims, material_on=True, qsearch="pe", koth=True, tier1_light="deep"), CLASSICAL)
        kids, info = t.root_children(), t.root_child_info()
        gated = 0
        for (m, visits, total), (tv, dist, n_children) in zip(kids, info):
            if tv is None:
                continue
            gated += 1
            assert n_children == 0, O.uci(m)
            if visits:
                assert abs(total / visits + tv) < 0.01, (O.uci(m), total, visits, tv)
            nb = O.make_move(O.from_fen(fen), m)
            assert O.tier1_gates(nb, koth=True)[0] == int(tv)          # and the value is the gate's
        assert gated > 0
        # the gate at depth 1 sees none of this: these are hills in 2 and 3
        assert all(O.light_gates(O.make_move(O.from_fen(fen), m), koth=True) == 2 for (m, _, _), (tv, _, _) in zip(kids, info) if tv == 1.0)


def test_koth_early_termination_and_best_move():
    # :433-465, :737-760 (training search): a king step onto the hill ends the search at once and is the move returned
    t = O.Tree(O.from_fen("3N4/k6p/p6P/2P1R3/2n3P1/P1rK4/8/5B2 w - - 4 47"))
    it = t.search(O.mcts_config(sims=1000, material_on=True, qsearch="pe", koth=True, tier1_light=True), CLASSICAL)
    assert it < 100 and O.uci(t.best_move()) in ("d3d4", "d3e4")
    # :516-547: every move loses at once => a move is still returned, quickly
    t = O.Tree(O.from_fen("8/8/8/8/8/3k4/8/4K3 w - - 0 1"))
    it = t.search(O.mcts_config(sims=200, koth=True, tier1_light=True), CLASSICAL)
    assert it < 50 and t.best_move() != 0
    # :403-430 reads the root's own gate, which the game loop applies before the search (self_play.rs:241-283)
    assert O.light_gates(O.from_fen("8/k7/8/8/4K3/8/8/8 b - - 0 1"), koth=True) == -1


def test_koth_only_safe_move():
    # :468-513: e2e4 is the only move that does not lose to a king step onto the hill
    t = O.Tree(O.from_fen("r1b2bnr/p2q1p1p/3kP1p1/2p5/p7/N4N2/PP1PPPPP/R1BK1B1R w - - 0 10"))
    t.search(O.mcts_config(sims=200, material_on=True, qsearch="pe", koth=True, tier1_light=True), CLASSICAL)
    ch = t.root_children()
    e4 = [v for m, v, _ in ch if O.uci(m) == "e2e4"][0]
    assert O.uci(t.best_move()) == "e2e4" and e4 > sum(v for _, v, _ in ch) * 3 // 4


def test_tree_reuse():
    # tree_reuse_tests.rs:35-140, tactical_mcts_tests.rs:762-806: the played child keeps its visits and grandchildren
    t = O.Tree(O.from_fen(START))
    t.search(O.mcts_config(sims=100), UNIFORM)
    m, v, _ = t.root_children()[0]
    assert not t.advance(O.mv("e2e5")) and t.advance(m)
    assert t.root_stats()["visits"] == v
    kept = sum(x for _, x, _ in t.root_children())
    assert kept == v - 1                      # every visit of the node but its own evaluation went to a child
    t.search(O.mcts_config(sims=30), UNIFORM)
    assert t.root_stats()["visits"] == v + 30 and t.best_move() != 0


def test_game_loop():
    """self_play.rs:191-445 with mock evaluators: games end, samples carry z from the mover's side, policies are
    visit fractions over legal moves, the first sample is the start position."""
    start = O.from_fen(START)
    cfg = O.mcts_config(sims=24, max_plies=60)
    samples, res, plies = O.play_game(start, cfg, 7, UNIFORM)
    assert len(samples) == plies and plies >= 10 and res in (-1, 0, 1)
    assert np.array_equal(samples[0]["board"], np.frombuffer(bytes(start), dtype=np.uint8))
    for i, s in enumerate(samples):
        b = O.array_to_board(s["board"])
        assert bool(b.w_to_move) == (i % 2 == 0)
        assert s["z"] == (res if b.w_to_move else -res)
        assert abs(float(s["prob"].sum()) - 1.0) < 1e-5
        legal = {O.move_to_index_stm(m, b.w_to_move) for m in O.legal_moves(b)}
        assert set(int(x) for x in s["idx"]) == legal
    again, res2, plies2 = O.play_game(start, cfg, 7, UNIFORM)
    assert res2 == res and plies2 == plies and all(np.array_equal(a["prob"], b["prob"]) for a, b in zip(samples, again))
    other, _, _ = O.play_game(start, cfg, 8, UNIFORM)
    assert any(not np.array_equal(a["board"], b["board"]) for a, b in zip(samples, other))
    # a decisive game: the mover always thinks it is winning => captures happen; with the gates on, a mate in one
    # at the root is adjudicated (self_play.rs:241-283)
    s3, r3, p3 = O.play_game(O.from_fen("6k1/5ppp/8/8/8/8/5PPP/3R2K1 w - - 0 1"), O.mcts_config(sims=16, tier1_light=True), 3, UNIFORM)
    assert r3 == 1 and p3 == 0 and s3 == []
