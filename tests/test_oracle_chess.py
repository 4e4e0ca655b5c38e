"""The oracle's chess substrate against the reference's perft table (tests/perft_tests.rs:68-291)
and the make-move side effects that reach the network input (src/make_move.rs:99-185)."""
import pytest

from oracle import pyoracle as O

START = "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1"
POS1 = "r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1"
POS2 = "8/2p5/3p4/KP5r/1R3p1k/8/4P1P1/8 w - - 0 1"
POS3 = "r3k2r/Pppp1ppp/1b3nbN/nP6/BBP1P3/q4N2/Pp1P2PP/R2Q1RK1 w kq - 0 1"
POS4 = "rnbq1k1r/pp1Pbppp/2p5/8/2B5/8/PPP1NnPP/RNBQK2R w KQ - 1 8"
POS5 = "r4rk1/1pp1qppp/p1np1n2/2b1p1B1/2B1P1b1/P1NP1N2/1PP1QPPP/R4RK1 w - - 0 10"

FAST = [
    (START, 1, 20), (START, 2, 400), (START, 3, 8902), (START, 4, 197281),          # :68-89
    (POS1, 1, 48), (POS1, 2, 2039), (POS1, 3, 97862),                               # :105-123
    (POS2, 1, 14), (POS2, 2, 191), (POS2, 3, 2812), (POS2, 4, 43238), (POS2, 5, 674624),  # :141-168
    (POS3, 1, 6), (POS3, 2, 264), (POS3, 3, 9467), (POS3, 4, 422333),               # :178-203
    (POS4, 1, 44), (POS4, 2, 1486), (POS4, 3, 62379),                               # :223-238
    (POS5, 1, 46), (POS5, 2, 2079), (POS5, 3, 89890),                               # :254-275
]
MEDIUM = [(START, 5, 4865609), (POS1, 4, 4085603), (POS4, 4, 2103487), (POS5, 4, 3894594)]  # :92,:126,:241,:278
DEEP = [(START, 6, 119060324), (POS1, 5, 193690690), (POS2, 6, 11030083), (POS3, 5, 15833292),
        (POS3, 6, 706045033), (POS4, 5, 89941194), (POS5, 5, 164075551)]


@pytest.mark.parametrize("fen,depth,nodes", FAST + MEDIUM)
def test_perft(fen, depth, nodes):
    assert O.perft(O.from_fen(fen), depth) == nodes


@pytest.mark.slow
@pytest.mark.skipif("not config.getoption('--run-slow', default=False)")
@pytest.mark.parametrize("fen,depth,nodes", DEEP)
def test_perft_deep(fen, depth, nodes):
    assert O.perft(O.from_fen(fen), depth) == nodes


def test_en_passant_set_after_every_double_push():  # src/make_move.rs:99-109
    b = O.make_move(O.from_fen(START), O.mv("e2e4"))
    assert b.en_passant == O.sq("e3") and b.w_to_move == 0
    b = O.make_move(b, O.mv("g8f6"))
    assert b.en_passant == 0xFF


def test_castling_rights_updates():  # src/make_move.rs:128-185
    b = O.from_fen("r3k2r/8/8/8/8/8/8/R3K2R w KQkq - 0 1")
    assert O.make_move(b, O.mv("e1g1")).castling == 4 | 8           # king move drops both white rights
    assert O.make_move(b, O.mv("h1h2")).castling == 2 | 4 | 8       # rook leaves h1 -> WK gone
    assert O.make_move(b, O.mv("a1a8")).castling == 1 | 4           # capture on a8 drops BQ, a1 left drops WQ
    k = O.make_move(b, O.mv("e1c1"))
    assert k.pieces[3] & (1 << O.sq("d1")) and k.pieces[5] & (1 << O.sq("c1"))


def test_promotions_generate_all_four_pieces():  # src/magic_bitboard.rs:457-474
    moves = O.legal_moves(O.from_fen("8/P7/8/8/8/8/8/K6k w - - 0 1"))
    promos = sorted((m >> 12) & 7 for m in moves if (m & 63) == O.sq("a7"))
    assert promos == [1, 2, 3, 4]


def test_random_positions_are_reproducible_and_varied():
    a = O.random_positions(20261019, 512)
    b = O.random_positions(20261019, 512)
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all()
    boards = a[0]
    assert 0.3 < boards[:, 112].mean() < 0.7                  # both colours to move
    assert (boards[:, 113] != 0xFF).mean() > 0.02             # en-passant squares occur
    assert len(set(boards[:, 114].tolist())) > 4              # several castling-right states
    n = a[2][1:] - a[2][:-1]
    assert n.min() >= 1 and n.max() <= 218 and 15 < n.mean() < 45
