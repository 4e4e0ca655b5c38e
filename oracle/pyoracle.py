"""ctypes binding of oracle/_build/liboracle.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module (see oracle/oracle.h).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")

POLICY = 4672
PLANES = 1088
MAX_MOVES = 256


def build(force=False):
    """Compile the C restatement with gcc (seconds)."""
    srcs = [os.path.join(_HERE, f) for f in ("codec.c", "chess.c", "qsearch.c", "oracle.h", "pesto_tables.h")]
    if not force and os.path.exists(_SO) and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in srcs):
        return _SO
    subprocess.check_call(["make", "-C", _HERE, "-B"], stdout=subprocess.DEVNULL)
    return _SO


class Board(ctypes.Structure):
    _fields_ = [
        ("pieces", ctypes.c_uint64 * 12),
        ("occ", ctypes.c_uint64 * 2),
        ("w_to_move", ctypes.c_uint8),
        ("en_passant", ctypes.c_uint8),
        ("castling", ctypes.c_uint8),
        ("halfmove", ctypes.c_uint8),
        ("pad", ctypes.c_uint8 * 12),
    ]


assert ctypes.sizeof(Board) == 128

_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        bp = ctypes.POINTER(Board)
        L.orc_parse_fen.argtypes = [ctypes.c_char_p, bp]
        L.orc_parse_fen.restype = ctypes.c_int
        L.orc_startpos.argtypes = [bp]
        L.orc_board_to_planes.argtypes = [bp, ctypes.POINTER(ctypes.c_float)]
        L.orc_move_to_index.argtypes = [ctypes.c_int] * 3
        L.orc_move_to_index.restype = ctypes.c_int
        L.orc_move_to_index_stm.argtypes = [ctypes.c_uint16, ctypes.c_int]
        L.orc_move_to_index_stm.restype = ctypes.c_int
        L.orc_policy_to_move_priors.argtypes = [
            ctypes.POINTER(ctypes.c_float), ctypes.c_int, ctypes.POINTER(ctypes.c_uint16),
            ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_float)]
        L.orc_pst_eval_cp.argtypes = [bp]
        L.orc_pst_eval_cp.restype = ctypes.c_int
        L.orc_value_fuse.argtypes = [ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_int]
        L.orc_value_fuse.restype = ctypes.c_double
        L.orc_gen_legal.argtypes = [bp, ctypes.POINTER(ctypes.c_uint16)]
        L.orc_gen_legal.restype = ctypes.c_int
        L.orc_make_move.argtypes = [bp, ctypes.c_uint16, bp]
        L.orc_in_check.argtypes = [bp, ctypes.c_int]
        L.orc_in_check.restype = ctypes.c_int
        L.orc_perft.argtypes = [bp, ctypes.c_int]
        L.orc_perft.restype = ctypes.c_uint64
        L.orc_random_positions.argtypes = [
            ctypes.c_uint64, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.orc_random_positions.restype = ctypes.c_int
        ip = ctypes.POINTER(ctypes.c_int)
        up = ctypes.POINTER(ctypes.c_uint32)
        L.orc_gen_captures_mvvlva.argtypes = [bp, ctypes.POINTER(ctypes.c_uint16), ip]
        L.orc_gen_captures_mvvlva.restype = ctypes.c_int
        L.orc_principal_exchange.argtypes = [bp, ctypes.c_int, ctypes.c_int, ctypes.c_int, up, ip]
        L.orc_principal_exchange.restype = ctypes.c_int
        L.orc_forced_principal_exchange.argtypes = [bp, up, ip]
        L.orc_forced_principal_exchange.restype = ctypes.c_float
        for name in ("orc_mate_in_1", "orc_koth_in_1"):
            getattr(L, name).argtypes = [bp]
            getattr(L, name).restype = ctypes.c_int
        L.orc_light_gates.argtypes = [bp, ctypes.c_int]
        L.orc_light_gates.restype = ctypes.c_int
        _lib = L
    return _lib


# ---------------------------------------------------------------- helpers
PROMO = {None: 0, "n": 1, "b": 2, "r": 3, "q": 4}


def sq(name):
    return (ord(name[1]) - ord("1")) * 8 + (ord(name[0]) - ord("a"))


def mv(uci):
    """'e2e4' / 'a7a8r' -> packed move."""
    return sq(uci[0:2]) | (sq(uci[2:4]) << 6) | (PROMO[uci[4] if len(uci) > 4 else None] << 12)


def from_fen(fen):
    b = Board()
    if lib().orc_parse_fen(fen.encode(), ctypes.byref(b)) != 0:
        raise ValueError("bad FEN: " + fen)
    return b


def boards_to_array(boards):
    """list[Board] -> uint8 [n,128]."""
    out = np.zeros((len(boards), 128), dtype=np.uint8)
    for i, b in enumerate(boards):
        out[i] = np.frombuffer(bytes(b), dtype=np.uint8)
    return out


def array_to_board(row):
    return Board.from_buffer_copy(np.ascontiguousarray(row, dtype=np.uint8).tobytes())


def board_to_planes(b):
    out = np.zeros(PLANES, dtype=np.float32)
    lib().orc_board_to_planes(ctypes.byref(b), out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
    return out


def planes_batch(boards_u8):
    """uint8 [n,128] -> float32 [n,17,8,8]."""
    n = boards_u8.shape[0]
    out = np.zeros((n, PLANES), dtype=np.float32)
    for i in range(n):
        b = array_to_board(boards_u8[i])
        lib().orc_board_to_planes(ctypes.byref(b), out[i].ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
    return out.reshape(n, 17, 8, 8)


def move_to_index(frm, to, promo=0):
    return lib().orc_move_to_index(frm, to, promo)


def move_to_index_stm(m, w_to_move):
    return lib().orc_move_to_index_stm(m, int(w_to_move))


def legal_moves(b):
    buf = (ctypes.c_uint16 * MAX_MOVES)()
    n = lib().orc_gen_legal(ctypes.byref(b), buf)
    return [buf[i] for i in range(n)]


def make_move(b, m):
    out = Board()
    lib().orc_make_move(ctypes.byref(b), m, ctypes.byref(out))
    return out


def perft(b, depth):
    return lib().orc_perft(ctypes.byref(b), depth)


def pst_eval_cp(b):
    return lib().orc_pst_eval_cp(ctypes.byref(b))


def value_fuse(v_logit, k, q, material_on=True):
    return lib().orc_value_fuse(float(v_logit), float(k), float(q), int(material_on))


def policy_to_move_priors(policy, moves, w_to_move):
    policy = np.ascontiguousarray(policy, dtype=np.float32)
    mv_arr = (ctypes.c_uint16 * len(moves))(*moves)
    out = np.zeros(len(moves), dtype=np.float32)
    lib().orc_policy_to_move_priors(
        policy.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), policy.size, mv_arr, len(moves),
        int(w_to_move), out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
    return out


def random_positions(seed, n, max_plies=160):
    """Seeded random legal playouts -> (boards u8 [n,128], move_idx u16, move_off u32 [n+1], moves_raw u16)."""
    boards = np.zeros((n, 128), dtype=np.uint8)
    idx = np.zeros(n * MAX_MOVES, dtype=np.uint16)
    raw = np.zeros(n * MAX_MOVES, dtype=np.uint16)
    off = np.zeros(n + 1, dtype=np.uint32)
    got = lib().orc_random_positions(seed, n, max_plies, boards.ctypes.data, idx.ctypes.data,
                                     off.ctypes.data, raw.ctypes.data)
    assert got == n
    total = int(off[n])
    return boards, idx[:total].copy(), off, raw[:total].copy()


def captures_mvvlva(b):
    """MVV-LVA ordered pseudo-legal captures + promotions -> (moves, list longer than the pinned 20)."""
    buf = (ctypes.c_uint16 * MAX_MOVES)()
    long_list = ctypes.c_int(0)
    n = lib().orc_gen_captures_mvvlva(ctypes.byref(b), buf, ctypes.byref(long_list))
    return [buf[i] for i in range(n)], bool(long_list.value)


def principal_exchange(b, alpha=-100000, beta=100000, max_depth=20):
    """-> (score_cp, nodes, long_list)."""
    nodes = ctypes.c_uint32(0)
    long_list = ctypes.c_int(0)
    cp = lib().orc_principal_exchange(ctypes.byref(b), alpha, beta, max_depth, ctypes.byref(nodes), ctypes.byref(long_list))
    return cp, nodes.value, bool(long_list.value)


def forced_principal_exchange(b):
    """-> (pawn units as f32, nodes)."""
    nodes = ctypes.c_uint32(0)
    long_list = ctypes.c_int(0)
    s = lib().orc_forced_principal_exchange(ctypes.byref(b), ctypes.byref(nodes), ctypes.byref(long_list))
    return float(s), nodes.value


def mate_in_1(b):
    return bool(lib().orc_mate_in_1(ctypes.byref(b)))


def koth_in_1(b):
    return bool(lib().orc_koth_in_1(ctypes.byref(b)))


def light_gates(b, koth=False):
    """-1 / +1 for the side to move, 2 when no gate fires."""
    return lib().orc_light_gates(ctypes.byref(b), int(koth))


def pe_q(boards_u8):
    """Principal-exchange material scalar of every board (pawn units, side to move)."""
    q = np.zeros(boards_u8.shape[0], dtype=np.float32)
    for i in range(boards_u8.shape[0]):
        q[i] = forced_principal_exchange(array_to_board(boards_u8[i]))[0]
    return q


def static_q(boards_u8, clip=20.0):
    """Stand-pat material scalar fed as q_result: pst_eval_cp/100 (pawn units, side to move)."""
    q = np.zeros(boards_u8.shape[0], dtype=np.float32)
    for i in range(boards_u8.shape[0]):
        q[i] = pst_eval_cp(array_to_board(boards_u8[i])) / 100.0
    return np.clip(q, -clip, clip).astype(np.float32)
