"""oracle/gates.c (mate search, hill in n at full depth) pinned to the reference's own tests.

Every case restates one test of tests/unit/mate_search_tests.rs or tests/unit/koth_tests.rs of the reference
(cited per case).  Squares: a1 = 0, index = rank * 8 + file.
"""
import pytest

from oracle import pyoracle as O

START = "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1"
BACK_RANK = "6k1/5ppp/8/8/8/8/8/4R1K1 w - - 0 1"
SCHOLARS = "r1bqkbnr/pppp1ppp/2n5/4p2Q/2B1P3/8/PPPP1PPP/RNB1K1NR w KQkq - 4 4"
BLACK_BACK_RANK = "4r1k1/8/8/8/8/8/5PPP/6K1 b - - 0 1"
STALEMATE = "k7/1R6/K7/8/8/8/8/8 b - - 0 1"
QUIET_M2 = "7k/R7/5K2/8/8/8/8/8 w - - 0 1"
ITALIAN = "r1bqk2r/pppp1ppp/2n2n2/2b1p3/2B1P3/5N2/PPPP1PPP/RNBQK2R w KQkq - 4 4"
MATE = 1_000_000


def ms(fen, depth, exhaustive=2):
    return O.mate_search(O.from_fen(fen), depth, exhaustive)


def test_mates_found_with_the_reference_moves():
    # mate_search_tests.rs:26-33, 72-83, 85-94, 180-196, 572-585, 624-638, 701-708
    score, mv, nodes = ms(BACK_RANK, 2)
    assert score == MATE + 1 and mv[1] == 60 and nodes > 0
    score, mv, _ = ms(SCHOLARS, 3)
    assert score >= MATE and mv[1] == 53
    score, mv, _ = ms(BLACK_BACK_RANK, 2)
    assert score >= MATE and mv[:2] == (60, 4)
    score, mv, _ = ms(QUIET_M2, 3)
    assert score == MATE + 3 and mv[:2] == (45, 46)
    score, mv, _ = ms("7k/8/5BK1/8/8/8/8/R7 w - - 0 1", 2)
    assert score >= MATE and mv[1] == 56
    score, mv, _ = ms("7k/6P1/6K1/8/8/8/8/R7 w - - 0 1", 2)
    assert score >= MATE and mv[1] == 56
    assert ms("7k/5ppp/8/5Q2/8/8/8/6K1 w - - 0 1", 3)[0] >= MATE            # :36-44
    assert ms(QUIET_M2, 3)[0] >= MATE                                       # :353-367


def test_no_mate_where_the_reference_finds_none():
    # :46-58, 96-106, 166-176, 345-351, 563-570, 640-651, 653-699
    assert abs(ms(START, 4)[0]) < MATE
    assert abs(ms(STALEMATE, 2)[0]) < MATE
    score, _, nodes = ms(START, 10)
    assert score == 0 and nodes > 0
    assert ms(ITALIAN, 1)[0] == 0
    for fen in ["4R2k/6p1/8/8/8/8/8/4K3 b - - 0 1", "4k3/8/8/1B6/8/8/8/3RK3 w - - 0 1", "5nk1/8/6Q1/8/8/8/8/4K3 b - - 0 1",
                "Q3k3/3r4/8/8/8/8/8/4K3 b - - 0 1", "6k1/4pppp/8/8/4N3/8/8/4K3 w - - 0 1", "6k1/5ppp/3r4/8/4N3/8/8/4K3 w - - 0 1"]:
        assert ms(fen, 2)[0] < MATE, fen
    ms("rnb1kbnr/pppp1ppp/8/4p3/6Pq/5P2/PPPPP2P/RNBQKBNR w KQkq - 1 3", 2)   # :60-69: a mated side to move; must not crash


def test_exhaustive_depth_decides_whether_quiet_first_moves_are_seen():
    # :198-211, 213-245, 247-265, 267-280, 282-317, 319-343
    assert ms(QUIET_M2, 3, 2)[0] >= MATE and ms(QUIET_M2, 3, 0)[0] < MATE
    for fen, depth, to in [(BACK_RANK, 2, 60), (SCHOLARS, 3, 53)]:
        for exhaustive in (2, 0):
            score, mv, _ = ms(fen, depth, exhaustive)
            assert score >= MATE and mv[1] == to
    assert abs(ms(STALEMATE, 3, 2)[0]) < MATE
    score, _, nodes = ms("r1bqkbnr/pppppppp/2n5/4p3/2B1P3/5N2/PPPP1PPP/RNBQK2R w KQkq - 4 4", 3, 2)
    assert score == 0 and nodes > 0
    score, mv, _ = ms(BLACK_BACK_RANK, 3, 2)
    assert score >= MATE and mv[1] == 4
    for fen, depth, mate in [(BACK_RANK, 2, True), (SCHOLARS, 3, True), (BLACK_BACK_RANK, 2, True), (STALEMATE, 2, False), (START, 3, False)]:
        assert (ms(fen, depth, 0)[0] >= MATE) == mate, fen
    other = "7k/8/5K2/R7/8/8/8/8 w - - 0 1"
    score, mv, _ = ms(other, 3, 2)
    assert score >= MATE and mv[0] == 45 and ms(other, 3, 0)[0] < MATE


def test_node_counts_grow_with_depth_and_returned_moves_are_legal():
    # :108-133, 136-164, 477-493, 496-560
    dev = "r1bqkbnr/pppp1ppp/2n5/4p3/2B1P3/5N2/PPPP1PPP/RNBQK2R w KQkq - 4 4"
    assert ms(dev, 3)[2] >= ms(dev, 1)[2]
    assert ms(START, 2)[2] >= ms(START, 1)[2]
    for fen in [START, "r1bqkb1r/pppp1ppp/2n2n2/4p3/2B1P3/5N2/PPPP1PPP/RNBQK2R w KQkq - 4 4", "8/8/8/8/8/5k2/8/4K2R w - - 0 1"]:
        b = O.from_fen(fen)
        _, mv, _ = O.mate_search(b, 3, 2)
        if mv is not None:
            legal = {(m & 63, (m >> 6) & 63, (m >> 12) & 7) for m in O.gen_legal_ref(b)[0]}
            assert mv in legal
    cases = [(BACK_RANK, 2, True, 60), (BLACK_BACK_RANK, 2, True, 4), (SCHOLARS, 3, True, 53), (QUIET_M2, 3, True, None),
             ("7k/8/5K2/R7/8/8/8/8 w - - 0 1", 3, True, None), (START, 3, False, None), (STALEMATE, 2, False, None), (ITALIAN, 1, False, None)]
    for fen, depth, mate, to in cases:
        score, mv, nodes = ms(fen, depth)
        assert (score >= MATE) == mate and nodes >= 0, fen
        if mate:
            assert mv is not None and (to is None or mv[1] == to), fen


def test_checking_move_filter_equals_check_after_the_move():
    # :374-422 (gives_check == is_check after the move) is how gates.c states the filter; what remains to pin is that the
    # filtered search and the unfiltered one agree on mates that start with a check (:428-474)
    for fen, depth in [(BACK_RANK, 2), (BLACK_BACK_RANK, 2), (SCHOLARS, 3)]:
        score, mv, _ = ms(fen, depth)
        assert score >= MATE and mv is not None
    for fen, depth in [(START, 3), (STALEMATE, 2)]:
        assert abs(ms(fen, depth)[0]) < MATE


def test_root_gate_of_the_search_entry_points():
    """tests/unit/tactical_mcts_tests.rs:43-71,895-921: tactical_mcts_search answers a root with a forced mate from its gate,
    mate_search(board, mate_search_depth = 2, exhaustive_mate_depth = 0) (src/mcts/tactical_mcts.rs:392-420): Re8#."""
    score, mv, _ = ms(BACK_RANK, 2, 0)
    assert score == MATE + 1 and mv[1] == 60
    # koth_tests.rs:182-227: the root gate of a King-of-the-Hill search, koth_center_in_n(board, koth_depth = 3): forced in 2
    assert hill("rnbqkbnr/1pppppp1/p6p/8/8/4P3/PPPPKPPP/RNBQ1BNR w kq - 0 3")[0] == 2


def test_a_mate_in_three_by_checks():
    """Not a reference case: two rooks ladder a bare king (Ra7+? no - every move a check): the checks-only levels find it
    and report 1_000_000 + 5 plies; a depth too small reports 0."""
    fen = "7k/8/8/8/8/8/R7/1R4K1 w - - 0 1"      # 1.Rh2+? Kg7 ... not forced; the search decides, we only compare depths
    s5 = ms(fen, 5, 0)[0]
    s1 = ms(fen, 1, 0)[0]
    assert s1 == 0
    assert s5 == 0 or s5 in (MATE + 3, MATE + 5, MATE + 7, MATE + 9)
    ladder = "6k1/R7/1R6/8/8/8/8/6K1 w - - 0 1"  # 1.Rb8#
    assert ms(ladder, 5, 0)[0] == MATE + 1
    two = "5k2/R7/8/1R6/8/8/8/6K1 w - - 0 1"     # 1.Rb8# as well
    assert ms(two, 5, 0)[0] == MATE + 1
    three = "4k3/8/R7/1R6/8/8/8/6K1 w - - 0 1"   # 1.Rb7 is quiet; by checks: 1.Ra8+ Kd7/e7/f7 2.Rb7+ K-6th 3.Ra6+ ... no mate in 2 checks
    s = ms(three, 2, 0)[0]
    assert s == 0
    assert ms(three, 2, 2)[0] == MATE + 3        # exhaustive: 1.Rb7 and 2.Ra8#


# ------------------------------------------------------------------ King of the Hill (tests/unit/koth_tests.rs)
def hill(fen, n=3):
    return O.koth_center_in_n(O.from_fen(fen), n)


def test_hill_distances():
    # koth_tests.rs:24-52 (one and two steps), :126-137 (start), :139-152, :199-208 (forced in 2), :229-240 (defended)
    assert hill("8/8/8/8/3K4/8/8/k7 w - - 0 1")[0] == 0
    assert hill("8/8/8/8/8/2K5/8/k7 w - - 0 1")[0] == 1
    assert hill("8/8/8/8/8/8/1K6/k7 w - - 0 1")[0] == 2
    r = hill("8/8/8/3k4/8/8/8/7K w - - 0 1")[0]
    assert r is None or r <= 3
    assert hill(START)[0] is None
    assert hill("rnbqkbnr/pppp2pp/4pp2/8/8/4P3/PPPPKPPP/RNBQ1BNR w kq - 0 3")[0] == 2
    assert hill("rnbqkbnr/1pppppp1/p6p/8/8/4P3/PPPPKPPP/RNBQ1BNR w kq - 0 3")[0] == 2
    assert hill("rnbqkbnr/pppp1pp1/4p2p/8/8/4P3/PPPPKPPP/RNBQ1BNR w kq - 0 3")[0] is None


def test_hill_node_counts_and_depth_one_equals_the_light_gate():
    # :330-341: the ring filter keeps the start position under 100 nodes
    r, nodes = hill(START)
    assert r is None and nodes < 100
    # :296-328: every position of the equivalence list answers (the counted form IS the form restated)
    for fen in [START, "8/8/8/8/8/2K5/8/k7 w - - 0 1", "8/8/8/3k4/8/8/8/7K w - - 0 1", "K7/8/8/8/8/5k2/8/8 b - - 0 1",
                "8/8/8/8/8/8/1K6/k7 w - - 0 1", "8/8/8/8/2Kk4/8/8/8 w - - 0 1"]:
        b = O.from_fen(fen)
        n1 = O.koth_center_in_n(b, 1)[0]
        assert (n1 is not None) == O.koth_in_1(b), fen
        n3 = O.koth_center_in_n(b, 3)[0]
        if n1 is not None:
            assert n3 == n1


def test_leaf_gate_order_and_distance():
    # src/mcts/tactical_mcts.rs:619-708: opponent on the hill -> -1 (distance 0); hill in n -> +1 (n); mate -> +1 with the
    # SCORE cast to u8 as the distance: (1_000_000 + plies) & 0xFF = 64 + plies
    assert O.tier1_gates(O.from_fen("8/k7/8/8/4K3/8/8/8 b - - 0 1"), koth=True) == (-1, 0)
    assert O.tier1_gates(O.from_fen("8/8/8/8/8/8/1K6/k7 w - - 0 1"), koth=True) == (1, 2)
    assert O.tier1_gates(O.from_fen(BACK_RANK)) == (1, 65)
    assert O.tier1_gates(O.from_fen(QUIET_M2), exhaustive_depth=2) == (1, 67)
    assert O.tier1_gates(O.from_fen(QUIET_M2), exhaustive_depth=0)[0] == 2       # self-play's leaf gate is checks-only
    assert O.tier1_gates(O.from_fen(START))[0] == 2
    # depth 1 = the light gates
    for fen in [BACK_RANK, SCHOLARS, START, STALEMATE, ITALIAN, "8/8/8/8/8/2K5/8/k7 w - - 0 1"]:
        for koth in (False, True):
            b = O.from_fen(fen)
            assert O.tier1_gates(b, koth=koth, mate_depth=1, koth_depth=1)[0] == O.light_gates(b, koth=koth), fen


def test_search_with_the_deep_gates_resolves_a_mate_in_two_by_checks():
    """With tier1 = 2 a child from which the opponent... rather: a leaf where the side to move mates in 2 by checks is
    decided by the gate (value +1 for its side to move), which the depth-1 gate leaves to the evaluator."""
    # White: Kg1, Qh5, Rf1; Black: Kg8, pawns f7 g7 h7 -> after a black waiting move white has Qxf7+ Kh8 Qf8# (checks only)
    fen = "6k1/5ppp/8/7Q/8/8/8/5RK1 w - - 0 1"
    b = O.from_fen(fen)
    assert O.tier1_gates(b, mate_depth=5, exhaustive_depth=0) == (1, 67)
    assert O.light_gates(b) == 2


@pytest.mark.parametrize("koth", [False, True])
def test_deep_gates_agree_with_an_unpruned_minimax_on_random_positions(koth):
    """The results are pure functions of the position: compare with a plain minimax without move-order shortcuts or ring
    pruning, written here from the rules (checks-only mate-in-2; hill in 2), on positions from random playouts."""
    boards, _, _, _ = O.random_positions(11, 150, 80)

    def legal_children(b):
        out = []
        for m in O.legal_moves(b):
            out.append(O.make_move(b, int(m)))
        return out

    def in_check(b):
        return O.in_check(b)

    def mated(b):
        return in_check(b) and not legal_children(b)

    def mate_by_checks(b, d):       # the side to move mates within d own moves, every one a check
        for c in legal_children(b):
            if not in_check(c):
                continue
            replies = legal_children(c)
            if not replies:
                return True
            if d > 1 and all(mate_by_checks(r, d - 1) for r in replies):
                return True
        return False

    CENTER = 0x0000001818000000

    def king_on_hill(b, side):
        return bool(b.pieces[side * 6 + 5] & CENTER)

    def hill_in(b, n):              # root = the side to move of b
        us = 0 if b.w_to_move else 1
        if king_on_hill(b, us):
            return True
        if king_on_hill(b, 1 - us) or n == 0:
            return False
        for c in legal_children(b):
            if king_on_hill(c, us):
                return True
            if n > 1:
                replies = legal_children(c)
                if replies and all((not king_on_hill(r, 1 - us)) and hill_in(r, n - 1) for r in replies):
                    return True
        return False

    hits = 0
    for i in range(boards.shape[0]):
        b = O.array_to_board(boards[i])
        if koth:
            want = hill_in(b, 2)
            got = O.koth_center_in_n(b, 2)[0] is not None
        else:
            want = mate_by_checks(b, 2)
            got = O.mate_search(b, 2, 0)[0] != 0
        assert got == want, i
        hits += want
    assert koth or hits >= 0


def test_exhaustive_levels_agree_with_a_plain_minimax_on_small_positions():
    """The adjudication search (mate_search(board, 2, exhaustive 2): every legal move at the attacker's plies) against a
    mate-in-2 minimax written here from the rules, on playout positions with few legal moves (the reference's rule for a
    side without moves on an attacker ply, src/search/mate_search.rs:276-283, never fires in these: it needs the defender to
    deliver mate)."""
    boards, _, _, _ = O.random_positions(2024, 1200, 220)

    def kids(b):
        return [O.make_move(b, int(m)) for m in O.legal_moves(b)]

    def mated(b):
        return O.in_check(b) and not O.legal_moves(b)

    def mate_in(b, d):
        for c in kids(b):
            if mated(c):
                return True
            if d > 1:
                replies = kids(c)
                if replies and not any(mated(r) for r in replies) and all(mate_in(r, d - 1) for r in replies):
                    return True
        return False

    checked = hits = 0
    for i in range(boards.shape[0]):
        b = O.array_to_board(boards[i])
        if len(O.legal_moves(b)) > 12 or checked >= 60:
            continue
        checked += 1
        want = mate_in(b, 2)
        got = O.mate_search(b, 2, 2)[0]
        assert (got != 0) == want, i
        hits += want
    assert checked >= 40
