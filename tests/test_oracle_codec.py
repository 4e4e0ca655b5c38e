"""The CPU oracle against the reference's own known answers for the integer half of the path.

Every literal below is transcribed from a reference test (file:line given per case):
tests/unit/board_encoding_tests.rs, tests/unit/tensor_extended_tests.rs, src/tensor.rs:215-246,
tests/unit/pesto_qsearch_tests.rs:14-80, cuda/test/test_nn_ops.cu:619-645.
"""
import os
import numpy as np
import pytest

from oracle import pyoracle as O

START = "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1"
N_, B_, R_, Q_ = 1, 2, 3, 4


def planes(fen):
    return O.board_to_planes(O.from_fen(fen)).reshape(17, 8, 8)


def count(p, i):
    return int((p[i] == 1.0).sum())


# ---- A/B. plane assignment and rank orientation (board_encoding_tests.rs:39-195)
def test_startpos_plane_counts():  # :39-82, :352-382
    p = planes(START)
    assert [count(p, i) for i in range(12)] == [8, 2, 2, 2, 1, 1, 8, 2, 2, 2, 1, 1]
    assert count(p, 12) == 0
    for i in range(13, 17):
        assert count(p, i) == 64
    assert set(np.unique(p)) <= {0.0, 1.0}


def test_black_to_move_swaps_sides():  # :85-112
    p = planes("rnbqkbnr/pppppppp/8/8/4P3/8/PPPP1PPP/RNBQKBNR b KQkq e3 0 1")
    assert count(p, 0) == 8 and count(p, 6) == 8  # black pawns are STM, white pawns opponent
    assert count(p, 7) == 2 and count(p, 11) == 1


def test_white_rank_mapping():  # :119-131  a2 pawn -> (0, 6, 0)
    p = planes("4k3/8/8/8/8/8/P7/4K3 w - - 0 1")
    assert p[0, 6, 0] == 1.0 and count(p, 0) == 1


def test_black_rank_mapping():  # :134-148  black a7 pawn -> (0, 6, 0)
    p = planes("4k3/p7/8/8/8/8/8/4K3 b - - 0 1")
    assert p[0, 6, 0] == 1.0 and count(p, 0) == 1


def test_kings_symmetry():  # :151-171  STM king -> (5, 7, 4) for either side
    assert planes("4k3/8/8/8/8/8/8/4K3 w - - 0 1")[5, 7, 4] == 1.0
    assert planes("4k3/8/8/8/8/8/8/4K3 b - - 0 1")[5, 7, 4] == 1.0


def test_mirror_positions_identical():  # :174-195
    a = planes("4k3/8/8/8/8/8/4P3/4K3 w - - 0 1")
    b = planes("4k3/4p3/8/8/8/8/8/4K3 b - - 0 1")
    assert np.array_equal(a, b)


# ---- C. castling plane order (board_encoding_tests.rs:199-275)
@pytest.mark.parametrize("fen,ones", [
    ("4k3/8/8/8/8/8/8/4K2R w K - 0 1", [13]),      # :200-220
    ("4k2r/8/8/8/8/8/8/4K3 b k - 0 1", [13]),      # :223-243
    ("r3k3/8/8/8/8/8/8/4K3 w q - 0 1", [16]),      # :246-266
    (START, [13, 14, 15, 16]),                     # :269-277
])
def test_castling_planes(fen, ones):
    p = planes(fen)
    for i in range(13, 17):
        assert count(p, i) == (64 if i in ones else 0)


# ---- D. en passant (board_encoding_tests.rs:298-345)
def test_en_passant_cells():
    p = planes("rnbqkbnr/ppp1pppp/8/3pP3/8/8/PPPP1PPP/RNBQKBNR w KQkq d6 0 3")
    assert p[12, 2, 3] == 1.0 and count(p, 12) == 1          # :298-313  d6 -> (12,2,3)
    p = planes("rnbqkbnr/pppppppp/8/8/4P3/8/PPPP1PPP/RNBQKBNR b KQkq e3 0 1")
    assert p[12, 2, 4] == 1.0 and count(p, 12) == 1          # :318-329  e3, black -> (12,2,4)
    assert planes("4k3/8/8/3pP3/8/8/8/4K3 w - d6 0 1")[12, 2, 3] == 1.0   # :332-345
    assert planes("4k3/8/8/8/3Pp3/8/8/4K3 b - d3 0 1")[12, 2, 3] == 1.0


def test_only_kings():  # :385-405
    p = planes("4k3/8/8/8/8/8/8/4K3 w - - 0 1")
    assert [count(p, i) for i in range(17)] == [0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0]


def test_specific_knights():  # :408-434  g1 knight (white) and g8 knight (black) -> (1,7,6)
    assert planes("4k3/8/8/8/8/8/8/4K1N1 w - - 0 1")[1, 7, 6] == 1.0
    assert planes("4k1n1/8/8/8/8/8/8/4K3 b - - 0 1")[1, 7, 6] == 1.0


# ---- policy index (src/tensor.rs:220-245, tensor_extended_tests.rs:9-225)
@pytest.mark.parametrize("frm,to,promo,expected", [
    (28, 36, 0, 28 * 73 + 0),    # src/tensor.rs:224  e4e5 = 2044
    (28, 31, 0, 28 * 73 + 16),   # :228  e4h4 = 2060
    (28, 45, 0, 28 * 73 + 56),   # :235  knight e4f6 = 2100
    (48, 56, R_, 48 * 73 + 70),  # :243  a7a8=R = 3574
    (0, 56, 0, 0 * 73 + 6),      # tensor_extended_tests.rs:17-21
    (0, 9, 0, 7), (0, 1, 0, 14), (56, 49, 0, 56 * 73 + 21), (36, 28, 0, 36 * 73 + 28),   # :24-49
    (63, 54, 0, 63 * 73 + 35), (7, 6, 0, 7 * 73 + 42), (7, 14, 0, 7 * 73 + 49),          # :52-70
    (48, 56, N_, 48 * 73 + 64), (49, 56, N_, 49 * 73 + 65), (48, 57, N_, 48 * 73 + 66),  # :105-123
    (52, 60, B_, 52 * 73 + 67), (52, 59, B_, 52 * 73 + 68), (52, 61, B_, 52 * 73 + 69),  # :126-144
    (48, 56, R_, 48 * 73 + 70), (49, 56, R_, 49 * 73 + 71), (48, 57, R_, 48 * 73 + 72),  # :147-165
    (48, 56, Q_, 48 * 73 + 0), (48, 57, Q_, 48 * 73 + 7),                                # :170-183
    (12, 28, 0, 877),            # board_encoding_tests.rs:440-448  e2e4
    (4, 6, 0, 307),              # cuda/test/test_nn_ops.cu:633-645  O-O
])
def test_move_to_index_known_answers(frm, to, promo, expected):
    assert O.move_to_index(frm, to, promo) == expected


def test_knight_all_directions():  # tensor_extended_tests.rs:75-101
    for to, k in [(44, 0), (37, 1), (21, 2), (12, 3), (10, 4), (17, 5), (33, 6), (42, 7)]:
        assert O.move_to_index(27, to, 0) == 27 * 73 + 56 + k


def test_slide_distances_and_range():  # tensor_extended_tests.rs:188-225
    for dist in range(1, 8):
        assert O.move_to_index(0, dist * 8, 0) == dist - 1
    for frm, to, promo in [(0, 8, 0), (63, 55, 0), (28, 45, 0), (48, 56, R_), (55, 63, Q_)]:
        assert 0 <= O.move_to_index(frm, to, promo) < 4672


def test_black_flip_equivalence():  # board_encoding_tests.rs:451-484
    assert O.move_to_index_stm(O.mv("e7e5"), False) == 877 == O.move_to_index_stm(O.mv("e2e4"), True)
    assert O.move_to_index_stm(O.mv("d5e4"), False) == O.move_to_index(27, 36, 0)


def test_index_is_injective_over_legal_moves():
    boards, idx, off, raw = O.random_positions(99, 200)
    for i in range(200):
        seg = idx[off[i]:off[i + 1]]
        assert len(set(seg.tolist())) == len(seg) and seg.max() < 4672


# ---- priors (src/tensor.rs:184-213)
def test_priors_renormalise_and_uniform_fallback():
    moves = [O.mv("e2e4"), O.mv("d2d4"), O.mv("g1f3")]
    pol = np.zeros(4672, dtype=np.float32)
    pol[O.move_to_index_stm(moves[0], True)] = 0.2
    pol[O.move_to_index_stm(moves[1], True)] = 0.6
    pr = O.policy_to_move_priors(pol, moves, True)
    np.testing.assert_allclose(pr, [0.25, 0.75, 0.0], rtol=1e-6)
    pr = O.policy_to_move_priors(np.zeros(4672, dtype=np.float32), moves, True)
    np.testing.assert_allclose(pr, [1 / 3] * 3, rtol=1e-6)
    # duplicates are counted as listed
    pr = O.policy_to_move_priors(pol, [moves[0], moves[0]], True)
    np.testing.assert_allclose(pr, [0.5, 0.5], rtol=1e-6)


# ---- static PeSTO (tests/unit/pesto_qsearch_tests.rs:14-80: inequalities) + hand-computed values
def test_pst_eval_reference_inequalities():
    assert abs(O.pst_eval_cp(O.from_fen(START))) < 50                                   # :14-24
    assert O.pst_eval_cp(O.from_fen("4k3/8/8/3Q4/8/8/8/4K3 w - - 0 1")) > 800           # :27-40
    w = O.pst_eval_cp(O.from_fen("4k3/8/8/3Q4/8/8/8/4K3 w - - 0 1"))
    b = O.pst_eval_cp(O.from_fen("4k3/8/8/3Q4/8/8/8/4K3 b - - 0 1"))
    assert w == -b and w > 0                                                            # :43-60 STM sign


def test_pst_eval_exact_small_cases():
    # startpos is symmetric -> exactly 0; lone kings e1/e8 symmetric -> 0
    assert O.pst_eval_cp(O.from_fen(START)) == 0
    assert O.pst_eval_cp(O.from_fen("4k3/8/8/8/8/8/8/4K3 w - - 0 1")) == 0
    # K e1 + P e2 vs K e8 (white to move), phase 0 -> pure endgame table:
    # pawn e2: 94 + EG_PAWN[e2^56 = index 52] = 94 + 13 = 107
    assert O.pst_eval_cp(O.from_fen("4k3/8/8/8/8/8/4P3/4K3 w - - 0 1")) == 107
    assert O.pst_eval_cp(O.from_fen("4k3/8/8/8/8/8/4P3/4K3 b - - 0 1")) == -107


def test_value_fuse_matches_f64_tanh():  # src/mcts/tactical_mcts.rs:762-765, :787-790
    import math
    assert O.value_fuse(0.3, 0.4, 1.5, True) == math.tanh(float(np.float32(0.3)) + float(np.float32(0.4)) * 1.5)
    assert O.value_fuse(0.3, 0.4, 1.5, False) == math.tanh(float(np.float32(0.3)))


# ------------------------------------------------------------------ independent statement of the 73-plane layout
@pytest.mark.skipif(not os.path.isdir("/root/reference/python"), reason="reference checkout not present (GPU box)")
def test_policy_index_commutes_with_the_reference_mirror_table():
    """python/augmentation.py states the policy layout a second time (direction order, knight order, underpromotion
    offsets) as permutation tables.  move_to_index must commute with the file mirror: index(mirror(move)) ==
    HFLIP_POLICY[index(move)] for every move kind.  (The quarter-turn table of that file is not used as a pin: its
    direction map (+2) and its square map (file -> rank) turn in opposite senses, e.g. a1-b1 lands on h1-h2 by the
    square table but is sent to the "south" plane by the direction table.)"""
    import sys
    sys.path.insert(0, "/root/reference/python")
    try:
        import augmentation as aug
    finally:
        sys.path.remove("/root/reference/python")
    hflip = lambda s: (s // 8) * 8 + (7 - s % 8)
    assert [hflip(s) for s in range(64)] == aug.HFLIP_SQ.tolist()
    checked = set()
    for frm in range(64):
        for to in range(64):
            for promo in (0, 1, 2, 3, 4):
                if promo and not (frm // 8 == 6 and to // 8 == 7 and abs(to % 8 - frm % 8) <= 1):
                    continue
                idx = O.move_to_index(frm, to, promo)
                if idx < 0:
                    continue
                assert O.move_to_index(hflip(frm), hflip(to), promo) == aug.HFLIP_POLICY[idx]
                checked.add(idx)
    # every queen-slide / knight entry reachable from some square, and all nine underpromotion planes
    assert len(checked) >= 1792 + 64 and {i % 73 for i in checked} == set(range(73))
