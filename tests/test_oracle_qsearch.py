"""The oracle's principal-exchange line against the reference's own tests for it
(tests/unit/ext_qsearch_tests.rs:306-350,546-567: node counts and inequalities — the reference holds no exact
scores) and against an independent brute-force statement of "first legal capture in MVV-LVA order"."""
import numpy as np
import pytest

from oracle import pyoracle as O

START = "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1"


def pe(fen):
    return O.forced_principal_exchange(O.from_fen(fen))


def test_starting_position():  # ext_qsearch_tests.rs:308-315
    score, nodes = pe(START)
    assert abs(score) < 0.5 and nodes == 1


def test_free_piece():  # :317-326
    score, nodes = pe("4k3/8/8/8/4n3/3P4/8/4K3 w - - 0 1")
    assert score > 0.5 and nodes <= 5


def test_defended_piece():  # :328-339 (the reference asserts the node bound only: the c4 pawn does not defend d5)
    fen = "4k3/8/8/3p4/2p5/8/8/3QK3 w - - 0 1"
    score, nodes = pe(fen)
    assert nodes <= 5
    assert score >= O.pst_eval_cp(O.from_fen(fen)) / 100.0


def test_truly_defended_piece_is_left_alone():  # QxP, PxQ loses the queen: the stand pat is the answer
    fen = "4k3/8/4p3/3p4/8/8/8/3QK3 w - - 0 1"
    score, nodes = pe(fen)
    assert nodes == 3
    assert score == pytest.approx(O.pst_eval_cp(O.from_fen(fen)) / 100.0)


def test_exchange_sequence():  # :341-350
    _, nodes = pe("4k3/8/8/8/8/8/8/r3K2R w K - 0 1")
    assert nodes <= 5


def test_no_captures():  # :546-555
    score, nodes = pe("4k3/8/8/8/8/8/8/4K3 w - - 0 1")
    assert nodes == 1 and score == 0.0


def test_promotion_capture():  # :557-567
    score, _ = pe("1n2k3/2P5/8/8/8/8/8/4K3 w - - 0 1")
    assert score > 2.0


def test_check_winning_material_is_invisible_to_the_exchange_line():  # :364-383 (the PE side of the comparison)
    fen = "r3k3/8/8/1N6/8/8/8/4K3 w - - 0 1"
    score, nodes = pe(fen)
    assert nodes == 1 and score == pytest.approx(O.pst_eval_cp(O.from_fen(fen)) / 100.0)


def test_mvv_lva_order_and_ties():  # src/move_generation.rs:426-446
    # two pawns can take the queen (tie: lower square first), a knight can take it too (worse attacker), then PxP
    b = O.from_fen("4k3/8/8/3q1p2/2P1P3/5N2/8/4K3 w - - 0 1")
    moves, long_list = O.captures_mvvlva(b)
    assert not long_list
    assert moves == [O.mv("c4d5"), O.mv("e4d5"), O.mv("e4f5")]
    b = O.from_fen("4k3/8/8/3q4/2P1P3/2N5/8/4K3 w - - 0 1")
    assert O.captures_mvvlva(b)[0] == [O.mv("c4d5"), O.mv("e4d5"), O.mv("c3d5")]
    # promotions score 0 unless they capture; a capturing promotion scores 10*victim like a pawn capture
    b = O.from_fen("1n2k3/2P5/8/8/8/8/8/4K3 w - - 0 1")
    assert O.captures_mvvlva(b)[0] == [O.mv("c7b8q"), O.mv("c7b8r"), O.mv("c7b8n"), O.mv("c7b8b"),
                                       O.mv("c7c8q"), O.mv("c7c8r"), O.mv("c7c8n"), O.mv("c7c8b")]
    # en passant scores 0 (empty target), behind every real capture except PxP (0 as well, generated first)
    b = O.from_fen("4k3/8/8/3pP3/8/8/8/4K3 w - d6 0 1")
    assert O.captures_mvvlva(b)[0] == [O.mv("e5d6")]


def _brute_force_pe(b, alpha, beta, depth):
    """The same line stated differently: rank the LEGAL captures/promotions by (score desc, generation order)."""
    sp = O.pst_eval_cp(b)
    if sp >= beta:
        return beta
    alpha = max(alpha, sp)
    if depth == 0:
        return alpha
    caps, _ = O.captures_mvvlva(b)
    legal = set(O.legal_moves(b))
    for m in caps:
        if m in legal:
            s = -_brute_force_pe(O.make_move(b, m), -beta, -alpha, depth - 1)
            if s >= beta:
                return beta
            return max(alpha, s)
    return alpha


def test_line_equals_the_legal_list_formulation_on_random_positions():
    boards, _, _, _ = O.random_positions(20261019, 400, 200)
    deep = 0
    for row in boards:
        b = O.array_to_board(row)
        cp, nodes, long_list = O.principal_exchange(b)
        assert not long_list
        assert cp == _brute_force_pe(b, -100000, 100000, 20)
        deep += nodes > 2
    assert deep > 40   # the corpus does exercise real exchanges


def test_captures_are_exactly_the_pseudo_legal_captures_and_promotions():
    boards, _, _, _ = O.random_positions(7, 300, 200)
    for row in boards:
        b = O.array_to_board(row)
        caps, _ = O.captures_mvvlva(b)
        assert len(set(caps)) == len(caps)
        them = b.occ[0] if not b.w_to_move else b.occ[1]
        legal = O.legal_moves(b)
        want = {m for m in legal if (them >> ((m >> 6) & 63)) & 1 or (m >> 12) & 7
                or (((m >> 6) & 63) == b.en_passant and any(b.pieces[(0 if b.w_to_move else 6)] >> (m & 63) & 1 for _ in [0]))}
        assert want <= set(caps)
        keys = []
        for m in caps:
            victim = next((i % 6 for i in range(12) if b.pieces[i] >> ((m >> 6) & 63) & 1), None)
            attacker = next(i % 6 for i in range(12) if b.pieces[i] >> (m & 63) & 1)
            keys.append(0 if victim is None else 10 * victim - attacker)
        assert keys == sorted(keys, reverse=True)


# ------------------------------------------------------------------ light Tier-1 gates (depth 1)
def test_mate_in_one_known_answers():  # tests/unit/mate_search_tests.rs:26-33,346-351,448-474
    for fen in ("6k1/5ppp/8/8/8/8/8/4R1K1 w - - 0 1",                       # Re8#
                "4r1k1/8/8/8/8/8/5PPP/6K1 b - - 0 1",                       # Re1#
                "r1bqkbnr/pppp1ppp/2n5/4p2Q/2B1P3/8/PPPP1PPP/RNB1K1NR w KQkq - 4 4"):   # Qxf7#
        assert O.mate_in_1(O.from_fen(fen)), fen
    for fen in (START,
                "r1bqk2r/pppp1ppp/2n2n2/2b1p3/2B1P3/5N2/PPPP1PPP/RNBQK2R w KQkq - 4 4",   # "no mate within depth 1"
                "k7/1R6/K7/8/8/8/8/8 b - - 0 1",                            # stalemate
                "7k/R7/5K2/8/8/8/8/8 w - - 0 1"):                           # a mate in TWO
        assert not O.mate_in_1(O.from_fen(fen)), fen


def test_hill_in_one_known_answers():  # tests/unit/koth_tests.rs:19-37,104-118
    assert O.koth_in_1(O.from_fen("8/8/8/8/3K4/8/8/k7 w - - 0 1"))          # already there: Some(0)
    assert O.koth_in_1(O.from_fen("8/8/8/8/8/2K5/8/k7 w - - 0 1"))          # Kc3-d4: Some(1)
    assert not O.koth_in_1(O.from_fen("8/8/8/8/8/8/1K6/k7 w - - 0 1"))      # two steps away
    assert not O.koth_in_1(O.from_fen("8/8/8/8/3k4/8/8/K7 w - - 0 1"))      # the opponent is already on the hill
    assert not O.koth_in_1(O.from_fen("8/8/8/2k5/8/2K5/8/8 w - - 0 1"))     # d4 is covered by the other king
    b = O.from_fen("8/8/8/8/3k4/8/8/K7 w - - 0 1")
    assert O.light_gates(b, koth=True) == -1 and O.light_gates(b, koth=False) == 2
    assert O.light_gates(O.from_fen("6k1/5ppp/8/8/8/8/8/4R1K1 w - - 0 1")) == 1


# ------------------------------------------------------------------ cap = 1 quiescence search (variant "cap1")
def test_cap1_known_answers_of_the_reference():
    """tests/unit/ext_qsearch_tests.rs:354-409,571-586: the start position is quiet and completes; a checking knight fork is
    found where the exchange line sees nothing; the back-rank check is worth more than three pawns (the mate is returned as the
    window's beta: 1000 pawn units); the "R3 blunder" position stays under 5000 nodes; a discovered check is searched."""
    score, completed, nodes, _ = O.forced_cap1(O.from_fen("rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1"))
    assert abs(score) < 0.5 and completed and nodes >= 1
    fork = O.from_fen("r3k3/8/8/1N6/8/8/8/4K3 w - - 0 1")
    assert O.forced_cap1(fork)[0] > O.forced_principal_exchange(fork)[0]
    score, _, nodes, _ = O.forced_cap1(O.from_fen("6k1/5ppp/8/8/8/8/8/3RK3 w - - 0 1"))
    assert score > 3.0 and score == 1000.0 and nodes == 2
    assert O.forced_cap1(O.from_fen("7k/1p1b1Qp1/1r6/p4pq1/1b1Np2p/4P3/K1rNR1PP/3R4 w - - 0 36"))[2] < 5000
    score, completed, nodes, depth = O.forced_cap1(O.from_fen("k7/8/8/8/8/8/1B6/R3K3 w - - 0 1"))
    assert completed and nodes >= 2 and depth >= 1
    for fen in ["rnbqkb1r/pppp1ppp/8/4p3/2B1N3/8/PPPP1PPP/R1BQK1NR b KQkq - 0 4", "r1b2rk1/pp2nppp/2pBp3/q7/1nP4P/5BN1/P4P2/R2QK2R w KQ - 0 15",
                "r1bq1rk1/ppp2ppp/2np1n2/2b1p3/2B1P3/2NP1N2/PPP2PPP/R1BQ1RK1 b - - 5 7"]:   # :410-455 (printed table: must run)
        O.forced_cap1(O.from_fen(fen))


def test_cap1_is_the_exchange_line_plus_one_check_per_side():
    """Structure, on 1500 playout positions: a position without any checking move available to either side along the line
    scores what the exchange line scores whenever no check exists at the root and cap1 used a single path; and the node count
    is bounded by the narrow tree (two children per node, evasions aside)."""
    boards, _, _, _ = O.random_positions(99, 1500, 200)
    same = 0
    for r in boards:
        b = O.array_to_board(r)
        s, completed, nodes, depth = O.forced_cap1(b)
        assert depth <= 20 and nodes >= 1 and -1000.0 <= s <= 1000.0
        same += s == O.forced_principal_exchange(b)[0]
    assert 0.3 * len(boards) < same < len(boards)
