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
    srcs = [os.path.join(_HERE, f) for f in ("codec.c", "chess.c", "qsearch.c", "mcts.c", "oracle.h", "pesto_tables.h")]
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
        L.orc_forced_cap1.argtypes = [bp, ip, up, ip, ip]
        L.orc_forced_cap1.restype = ctypes.c_float
        for name in ("orc_mate_in_1", "orc_koth_in_1"):
            getattr(L, name).argtypes = [bp]
            getattr(L, name).restype = ctypes.c_int
        L.orc_light_gates.argtypes = [bp, ctypes.c_int]
        L.orc_light_gates.restype = ctypes.c_int
        # the Tier-1 gates at full depth (gates.c)
        L.orc_mate_search.argtypes = [bp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_uint16), ip]
        L.orc_mate_search.restype = ctypes.c_int
        L.orc_koth_center_in_n.argtypes = [bp, ctypes.c_int, up]
        L.orc_koth_center_in_n.restype = ctypes.c_int
        L.orc_tier1_gates.argtypes = [bp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ip]
        L.orc_tier1_gates.restype = ctypes.c_int
        # tree operations (mcts.c)
        L.orc_gen_pseudo_ref.argtypes = [bp, ctypes.POINTER(ctypes.c_uint16), ip]
        L.orc_gen_pseudo_ref.restype = ctypes.c_int
        L.orc_gen_legal_ref.argtypes = [bp, ctypes.POINTER(ctypes.c_uint16), ip]
        L.orc_gen_legal_ref.restype = ctypes.c_int
        L.orc_tactical_score.argtypes = [bp, ctypes.c_uint16]
        L.orc_tactical_score.restype = ctypes.c_double
        L.orc_tactical_order.argtypes = [bp, ctypes.POINTER(ctypes.c_uint16), ctypes.c_int, ip]
        L.orc_mcts_new.argtypes = [bp, ctypes.c_uint64]
        L.orc_mcts_new.restype = ctypes.c_void_p
        L.orc_mcts_free.argtypes = [ctypes.c_void_p]
        L.orc_mcts_search.argtypes = [ctypes.c_void_p, ctypes.POINTER(MctsConfig), ctypes.c_void_p, ctypes.c_void_p]
        L.orc_mcts_search.restype = ctypes.c_int
        L.orc_mcts_root_children.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint16), up, ctypes.POINTER(ctypes.c_double)]
        L.orc_mcts_root_children.restype = ctypes.c_int
        lp = ctypes.POINTER(ctypes.c_long)
        L.orc_mcts_root_stats.argtypes = [ctypes.c_void_p, up, ctypes.POINTER(ctypes.c_double), lp, lp, lp]
        L.orc_mcts_best_move.argtypes = [ctypes.c_void_p]
        L.orc_mcts_best_move.restype = ctypes.c_uint16
        L.orc_mcts_advance.argtypes = [ctypes.c_void_p, ctypes.c_uint16]
        L.orc_mcts_advance.restype = ctypes.c_int
        L.orc_play_game.argtypes = [bp, ctypes.POINTER(MctsConfig), ctypes.c_uint64, ctypes.c_void_p, ctypes.c_void_p,
                                    ctypes.POINTER(GameSample), ctypes.c_int, ip, ip]
        L.orc_play_game.restype = ctypes.c_int
        _lib = L
    return _lib


class MctsConfig(ctypes.Structure):
    _fields_ = [("sims", ctypes.c_int), ("material_on", ctypes.c_int), ("qsearch", ctypes.c_int), ("koth", ctypes.c_int),
                ("tier1_light", ctypes.c_int), ("shuffle", ctypes.c_int), ("max_plies", ctypes.c_int),
                ("c_puct", ctypes.c_double), ("explore_base", ctypes.c_double),
                ("mate_depth", ctypes.c_int), ("exhaustive_mate_depth", ctypes.c_int), ("koth_depth", ctypes.c_int)]


class GameSample(ctypes.Structure):
    _fields_ = [("board", Board), ("n", ctypes.c_int), ("idx", ctypes.c_uint16 * MAX_MOVES),
                ("prob", ctypes.c_float * MAX_MOVES), ("material", ctypes.c_float), ("value_target", ctypes.c_float),
                ("w_to_move", ctypes.c_int)]


EVAL_FN = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.POINTER(Board), ctypes.POINTER(ctypes.c_uint16), ctypes.c_int,
                           ctypes.c_float, ctypes.c_int, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_double))


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


def in_check(b, color=None):
    """Is the king of `color` (0 white, 1 black; default: the side to move) attacked?"""
    L = lib()
    L.orc_in_check.argtypes = [ctypes.POINTER(Board), ctypes.c_int]
    L.orc_in_check.restype = ctypes.c_int
    return bool(L.orc_in_check(ctypes.byref(b), (0 if b.w_to_move else 1) if color is None else int(color)))


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


def forced_cap1(b):
    """forced_cap1_pesto_balance -> (pawn units as f32, completed, nodes, depth_used)."""
    c, d, ll = ctypes.c_int(1), ctypes.c_int(0), ctypes.c_int(0)
    n = ctypes.c_uint32(0)
    s = lib().orc_forced_cap1(ctypes.byref(b), ctypes.byref(c), ctypes.byref(n), ctypes.byref(d), ctypes.byref(ll))
    return float(s), bool(c.value), n.value, d.value


def cap1_q(boards_u8):
    """cap1 material scalar of every board (pawn units, side to move)."""
    q = np.zeros(boards_u8.shape[0], dtype=np.float32)
    for i in range(boards_u8.shape[0]):
        q[i] = forced_cap1(array_to_board(boards_u8[i]))[0]
    return q


def mate_in_1(b):
    return bool(lib().orc_mate_in_1(ctypes.byref(b)))


def koth_in_1(b):
    return bool(lib().orc_koth_in_1(ctypes.byref(b)))


def light_gates(b, koth=False):
    """-1 / +1 for the side to move, 2 when no gate fires."""
    return lib().orc_light_gates(ctypes.byref(b), int(koth))


def mate_search(b, max_depth, exhaustive_depth=2):
    """mate_search(board, max_depth, exhaustive_depth) -> (score, (from, to, promo) or None, nodes)."""
    mv = ctypes.c_uint16(0)
    nodes = ctypes.c_int(0)
    score = lib().orc_mate_search(ctypes.byref(b), int(max_depth), int(exhaustive_depth), ctypes.byref(mv), ctypes.byref(nodes))
    m = mv.value
    return score, ((m & 63, (m >> 6) & 63, (m >> 12) & 7) if m else None), nodes.value


def koth_center_in_n(b, max_n=3):
    """koth_center_in_n_counted -> (n or None, nodes)."""
    nodes = ctypes.c_uint32(0)
    n = lib().orc_koth_center_in_n(ctypes.byref(b), int(max_n), ctypes.byref(nodes))
    return (None if n < 0 else n), nodes.value


def tier1_gates(b, koth=False, mate_depth=5, exhaustive_depth=0, koth_depth=3):
    """The leaf gate at full depth -> (-1 / +1 / 2, terminal_distance)."""
    dist = ctypes.c_int(0)
    v = lib().orc_tier1_gates(ctypes.byref(b), int(koth), int(mate_depth), int(exhaustive_depth), int(koth_depth), ctypes.byref(dist))
    return v, dist.value


def gate_node_stats(reset=False):
    """Cost of the orc_tier1_gates calls so far: (calls, positions entered, largest single call)."""
    c, n, m = ctypes.c_long(0), ctypes.c_long(0), ctypes.c_long(0)
    lib().orc_gate_node_stats(ctypes.byref(c), ctypes.byref(n), ctypes.byref(m), int(reset))
    return c.value, n.value, m.value


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


# ---------------------------------------------------------------- tree operations (mcts.c)
def mcts_config(sims=200, material_on=False, qsearch=0, koth=False, tier1_light=False, shuffle=False, max_plies=200,
                c_puct=1.414, explore_base=0.80, mate_depth=0, exhaustive_mate_depth=0, koth_depth=0):
    """tier1_light: False / True (gates at depth 1) / 2 or "deep" (the reference's gate: mate-in-5 checks only, hill in 3)."""
    tier1 = 2 if tier1_light == "deep" else int(tier1_light)
    return MctsConfig(int(sims), int(bool(material_on)), 1 if qsearch in (1, "pe") else 2 if qsearch in (2, "cap1") else 0, int(koth), tier1,
                      int(shuffle), int(max_plies), float(c_puct), float(explore_base), int(mate_depth),
                      int(exhaustive_mate_depth), int(koth_depth))


def gen_legal_ref(b):
    """Legal moves in the reference's child order -> (moves, n_tactical)."""
    buf = (ctypes.c_uint16 * MAX_MOVES)()
    nt = ctypes.c_int(0)
    n = lib().orc_gen_legal_ref(ctypes.byref(b), buf, ctypes.byref(nt))
    return [buf[i] for i in range(n)], nt.value


def tactical_moves(b):
    """identify_tactical_moves: [(move, score)] in visiting order."""
    moves, nt = gen_legal_ref(b)
    arr = (ctypes.c_uint16 * max(1, len(moves)))(*moves)
    order = (ctypes.c_int * max(1, nt))()
    lib().orc_tactical_order(ctypes.byref(b), arr, nt, order)
    return [(moves[order[i]], lib().orc_tactical_score(ctypes.byref(b), moves[order[i]])) for i in range(nt)]


def _evaluator(evaluator):
    """evaluator: ('mock', v_logit[, k]) -> C mock; or a Python callable (board, moves, q, material_on) -> (priors, value)."""
    if isinstance(evaluator, tuple) and evaluator[0] == "mock":
        user = (ctypes.c_float * 2)(float(evaluator[1]), float(evaluator[2]) if len(evaluator) > 2 else 0.5)
        fn = ctypes.cast(lib().orc_eval_mock, ctypes.c_void_p)
        return fn, ctypes.cast(user, ctypes.c_void_p), user

    def trampoline(_user, bptr, mptr, n, q, material_on, priors, value):
        pri, val = evaluator(bptr.contents, [mptr[i] for i in range(n)], float(q), bool(material_on))
        for i in range(n):
            priors[i] = pri[i]
        value[0] = float(val)
    cb = EVAL_FN(trampoline)
    return ctypes.cast(cb, ctypes.c_void_p), None, cb


class Tree:
    """One search tree of the restated reference search (tactical_mcts_search_for_training_with_reuse)."""

    def __init__(self, board, seed=1):
        self.h = lib().orc_mcts_new(ctypes.byref(board), seed)

    def search(self, cfg, evaluator):
        fn, user, keep = _evaluator(evaluator)
        return lib().orc_mcts_search(self.h, ctypes.byref(cfg), fn, user)

    def root_children(self):
        mv_ = (ctypes.c_uint16 * MAX_MOVES)()
        vis = (ctypes.c_uint32 * MAX_MOVES)()
        tv = (ctypes.c_double * MAX_MOVES)()
        n = lib().orc_mcts_root_children(self.h, mv_, vis, tv)
        return [(mv_[i], vis[i], tv[i]) for i in range(n)]

    def root_child_info(self):
        """Per root child: (terminal_or_mate_value or None, terminal_distance or None, number of children)."""
        L = lib()
        L.orc_mcts_root_child_info.argtypes = [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 4
        L.orc_mcts_root_child_info.restype = ctypes.c_int
        out = []
        i = 0
        while True:
            ht, td, nc = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
            tv = ctypes.c_double(0)
            if not L.orc_mcts_root_child_info(self.h, i, ctypes.byref(ht), ctypes.byref(tv), ctypes.byref(td), ctypes.byref(nc)):
                return out
            out.append((tv.value if ht.value else None, td.value if ht.value and td.value >= 0 else None, nc.value))
            i += 1

    def root_stats(self):
        v = ctypes.c_uint32(0)
        tv = ctypes.c_double(0)
        e, th, tc = ctypes.c_long(0), ctypes.c_long(0), ctypes.c_long(0)
        lib().orc_mcts_root_stats(self.h, ctypes.byref(v), ctypes.byref(tv), ctypes.byref(e), ctypes.byref(th), ctypes.byref(tc))
        return dict(visits=v.value, total_value=tv.value, evals=e.value, terminal_hits=th.value, tactical=tc.value)

    def best_move(self):
        return lib().orc_mcts_best_move(self.h)

    def advance(self, move):
        return bool(lib().orc_mcts_advance(self.h, move))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_mcts_free(self.h)
            self.h = None


def play_game(start, cfg, seed, evaluator, max_samples=512):
    """-> (samples list of dict, result_white, plies)."""
    fn, user, keep = _evaluator(evaluator)
    buf = (GameSample * max_samples)()
    res, plies = ctypes.c_int(0), ctypes.c_int(0)
    n = lib().orc_play_game(ctypes.byref(start), ctypes.byref(cfg), seed, fn, user, buf, max_samples, ctypes.byref(res),
                            ctypes.byref(plies))
    out = []
    for i in range(min(n, max_samples)):
        s = buf[i]
        out.append(dict(board=np.frombuffer(bytes(s.board), dtype=np.uint8).copy(), idx=np.array(s.idx[:s.n], dtype=np.uint16),
                        prob=np.array(s.prob[:s.n], dtype=np.float32), material=s.material, z=s.value_target))
    return out, res.value, plies.value


def uci(m):
    s_ = lambda q_: "abcdefgh"[q_ & 7] + str((q_ >> 3) + 1)
    return s_(m & 63) + s_((m >> 6) & 63) + ["", "n", "b", "r", "q"][(m >> 12) & 7]
