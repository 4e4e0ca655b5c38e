"""caissa_b200 — ctypes binding of libcaissa_b200.so plus Python mirrors of the
reference's evaluator interface (`NeuralNetPolicy`, `InferenceServer`).

The compute lives in the CUDA library behind the C ABI of
include/caissa_b200.h; this module adds no arithmetic of its own and has no
CPU fallback: if the library or a B200 is missing, calls raise.
"""
import ctypes
import os
import queue
import threading
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libcaissa_b200.so")

POLICY_SIZE = 4672
PLANES = 1088
MAX_MOVES = 256
DTYPE_FP32, DTYPE_BF16, DTYPE_FP16 = 0, 1, 2
DTYPES = {"fp32": 0, "bf16": 1, "fp16": 2}


class CbTiming(ctypes.Structure):
    _fields_ = [("calls", ctypes.c_uint64), ("positions", ctypes.c_uint64), ("h2d_ms", ctypes.c_double),
                ("forward_ms", ctypes.c_double), ("d2h_ms", ctypes.c_double)]


class SelfplayConfig(ctypes.Structure):
    _fields_ = [("games", ctypes.c_int), ("sims_per_move", ctypes.c_int), ("max_plies", ctypes.c_int),
                ("material_on", ctypes.c_int), ("koth", ctypes.c_int), ("threads", ctypes.c_int),
                ("c_puct", ctypes.c_double), ("explore_base", ctypes.c_double), ("seed", ctypes.c_uint64),
                ("max_sims", ctypes.c_longlong), ("max_games", ctypes.c_longlong), ("max_seconds", ctypes.c_double),
                ("sample_path", ctypes.c_char_p), ("pipeline", ctypes.c_int), ("aux_handle", ctypes.c_void_p),
                ("qsearch", ctypes.c_int), ("tier1_light", ctypes.c_int), ("shuffle_children", ctypes.c_int),
                ("mock_eval", ctypes.c_int), ("mock_v_logit", ctypes.c_float), ("mock_k", ctypes.c_float),
                ("tier1_mate_depth", ctypes.c_int), ("tier1_exhaustive_depth", ctypes.c_int), ("tier1_koth_depth", ctypes.c_int)]


class SelfplayStats(ctypes.Structure):
    _fields_ = [(n, ctypes.c_longlong) for n in ("sims", "nn_evals", "terminal_hits", "moves", "games_finished",
                                                 "samples", "steps", "white_wins", "black_wins", "draws")] + \
               [(n, ctypes.c_double) for n in ("seconds", "eval_seconds", "tree_seconds", "mean_batch")] + \
               [("overflow", ctypes.c_longlong), ("gate_hits", ctypes.c_longlong), ("gate_waits", ctypes.c_longlong)]


class MatchConfig(ctypes.Structure):
    _fields_ = [("games", ctypes.c_int), ("total_games", ctypes.c_int), ("sims_per_move", ctypes.c_int),
                ("max_plies", ctypes.c_int), ("material_on", ctypes.c_int), ("koth", ctypes.c_int),
                ("c_puct", ctypes.c_double), ("explore_base", ctypes.c_double), ("seed", ctypes.c_uint64),
                ("use_sprt", ctypes.c_int), ("elo0", ctypes.c_double), ("elo1", ctypes.c_double),
                ("alpha", ctypes.c_double), ("beta", ctypes.c_double), ("max_seconds", ctypes.c_double),
                ("qsearch", ctypes.c_int), ("tier1_light", ctypes.c_int)]


class MatchStats(ctypes.Structure):
    _fields_ = [(n, ctypes.c_longlong) for n in ("games", "candidate_wins", "current_wins", "draws", "sims")] + \
               [("decision", ctypes.c_int), ("llr_valid", ctypes.c_int), ("llr", ctypes.c_double), ("seconds", ctypes.c_double)]


class CaissaError(RuntimeError):
    pass


def selfplay_config(games=1024, sims_per_move=200, max_plies=200, material_on=False, koth=False, threads=0, c_puct=1.414,
                    explore_base=0.80, seed=1, max_sims=0, max_games=0, max_seconds=0.0, sample_path=None, pipeline=1, aux=None,
                    qsearch=0, tier1_light=False, shuffle_children=False, mock=None, tier1_mate_depth=0, tier1_exhaustive_depth=0,
                    tier1_koth_depth=0):
    """tier1_light: False, True (gates at depth 1) or 2 / "deep" (the reference's gate: mate-in-5 by checks, hill in 3)."""
    tier1_light = 2 if tier1_light == "deep" else int(tier1_light)
    return SelfplayConfig(games, sims_per_move, max_plies, int(bool(material_on)), int(koth), threads, c_puct, explore_base, seed,
                          max_sims, max_games, max_seconds, sample_path.encode() if sample_path else None, pipeline,
                          aux._h if aux is not None else None, _QSEARCH[qsearch], int(tier1_light), int(shuffle_children),
                          int(mock is not None), float(mock[0]) if mock else 0.0, float(mock[1]) if mock else 0.5,
                          int(tier1_mate_depth), int(tier1_exhaustive_depth), int(tier1_koth_depth))


class _SearchOut:
    """Output buffers of the cb_debug_*_search hooks."""

    def __init__(self, n):
        self.n = n
        self.moves = np.zeros((n, MAX_MOVES), dtype=np.uint16)
        self.visits = np.zeros((n, MAX_MOVES), dtype=np.uint32)
        self.sums = np.zeros((n, MAX_MOVES), dtype=np.float64)
        self.counts = np.zeros(n, dtype=np.int32)
        self.root_visits = np.zeros(n, dtype=np.uint32)

    def ptrs(self):
        return (self.moves.ctypes.data, self.visits.ctypes.data, self.sums.ctypes.data, self.counts.ctypes.data,
                self.root_visits.ctypes.data)

    def unpack(self):
        kids = [[(int(self.moves[i, c]), int(self.visits[i, c]), float(self.sums[i, c])) for c in range(self.counts[i])]
                for i in range(self.n)]
        return kids, self.root_visits.copy()


EVAL_CALLBACK = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint16),
                                 ctypes.POINTER(ctypes.c_uint16), ctypes.c_int, ctypes.c_float, ctypes.c_int,
                                 ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float))


def _eval_callback(evaluator):
    """evaluator(board_bytes[128] as np.uint8, moves, policy_idx, q, material_on) -> (priors, v_final)"""
    def tramp(_user, bptr, mptr, iptr, n, q, material_on, priors, v_final):
        board = np.frombuffer(ctypes.string_at(bptr, 128), dtype=np.uint8)
        pri, val = evaluator(board, [mptr[i] for i in range(n)], [iptr[i] for i in range(n)], float(q), bool(material_on))
        for i in range(n):
            priors[i] = pri[i]
        v_final[0] = float(val)
    return EVAL_CALLBACK(tramp)


def debug_host_search(boards, evaluator, **kw):
    """One move's search with the HOST driver's tree code and a Python evaluator (no GPU needed) ->
    list of [(move, visits, value_sum)], root visit counts (cb_debug_host_search)."""
    L = load_library()
    boards = np.ascontiguousarray(boards, dtype=np.uint8).reshape(-1, 128)
    n = boards.shape[0]
    cfg = selfplay_config(games=n, **kw)
    out = _SearchOut(n)
    cb = _eval_callback(evaluator)
    rc = L.cb_debug_host_search(ctypes.byref(cfg), boards.ctypes.data, n, cb, None, *out.ptrs())
    if rc != 0:
        raise CaissaError("cb_debug_host_search failed: %d" % rc)
    return out.unpack()


def debug_host_game(start, seed, evaluator, max_samples=512, **kw):
    """One whole game of the host driver with a Python evaluator -> (samples, result_white, plies)."""
    L = load_library()
    start = np.ascontiguousarray(start, dtype=np.uint8).reshape(128)
    cfg = selfplay_config(games=1, **kw)
    sb = np.zeros((max_samples, 128), dtype=np.uint8)
    idx = np.zeros((max_samples, MAX_MOVES), dtype=np.uint16)
    prob = np.zeros((max_samples, MAX_MOVES), dtype=np.float32)
    cnt = np.zeros(max_samples, dtype=np.int32)
    mat = np.zeros(max_samples, dtype=np.float32)
    res, plies = ctypes.c_int(0), ctypes.c_int(0)
    cb = _eval_callback(evaluator)
    n = L.cb_debug_host_game(ctypes.byref(cfg), start.ctypes.data, ctypes.c_uint64(seed), cb, None, sb.ctypes.data, idx.ctypes.data,
                             prob.ctypes.data, cnt.ctypes.data, mat.ctypes.data, max_samples, ctypes.byref(res), ctypes.byref(plies))
    if n < 0:
        raise CaissaError("cb_debug_host_game failed: %d" % n)
    samples = [dict(board=sb[i].copy(), idx=idx[i, :cnt[i]].copy(), prob=prob[i, :cnt[i]].copy(), material=float(mat[i]))
               for i in range(min(n, max_samples))]
    return samples, res.value, plies.value


_lib = None

_F = ctypes.POINTER(ctypes.c_float)
_VP = ctypes.c_void_p

# name -> (restype, argtypes): every symbol include/caissa_b200.h declares
SIGNATURES = {
    "cb_create": (ctypes.c_int, [ctypes.POINTER(_VP), ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "cb_load_weights": (ctypes.c_int, [_VP, _VP, ctypes.c_size_t]),
    "cb_weight_count": (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int]),
    "cb_is_available": (ctypes.c_int, [_VP]),
    "cb_predict_batch": (ctypes.c_int, [_VP, _VP, _VP, _VP, ctypes.c_int, _VP, _VP, _VP]),
    "cb_eval_leaves": (ctypes.c_int, [_VP, _VP, _VP, ctypes.c_int, ctypes.c_int, _VP, _VP, _VP, _VP, _VP, _VP]),
    "cb_eval_leaves_submit": (ctypes.c_int, [_VP, _VP, _VP, ctypes.c_int, ctypes.c_int, _VP, _VP]),
    "cb_eval_leaves_wait": (ctypes.c_int, [_VP, _VP, _VP, _VP, _VP]),
    "cb_encode": (ctypes.c_int, [_VP, _VP, ctypes.c_int, _VP]),
    "cb_pst_eval_cp": (ctypes.c_int, [_VP, _VP, ctypes.c_int, _VP]),
    "cb_principal_exchange": (ctypes.c_int, [_VP, ctypes.c_int, _VP, _VP]),
    "cb_principal_exchange_device": (ctypes.c_int, [_VP, _VP, ctypes.c_int, _VP]),
    "cb_light_gates": (ctypes.c_int, [_VP, ctypes.c_int, ctypes.c_int, _VP]),
    "cb_tier1_gates": (ctypes.c_int, [_VP, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _VP, _VP]),
    "cb_debug_move_flags": (ctypes.c_longlong, [_VP, ctypes.c_int, _VP]),
    "cb_cap1_qsearch": (ctypes.c_int, [_VP, ctypes.c_int, _VP, _VP, _VP, _VP]),
    "cb_cap1_qsearch_device": (ctypes.c_int, [_VP, _VP, ctypes.c_int, _VP, _VP]),
    "cb_mate_search": (ctypes.c_int, [_VP, ctypes.c_int, ctypes.c_int, ctypes.c_int, _VP]),
    "cb_koth_center_in_n": (ctypes.c_int, [_VP, ctypes.c_int, ctypes.c_int, _VP]),
    "cb_debug_device_legal_moves": (ctypes.c_int, [_VP, _VP, ctypes.c_int, _VP, _VP]),
    "cb_debug_device_light_gates": (ctypes.c_int, [_VP, _VP, ctypes.c_int, ctypes.c_int, _VP]),
    "cb_debug_device_tier1_gates": (ctypes.c_int, [_VP, _VP, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                   ctypes.c_int, _VP, _VP]),
    "cb_stage_leaves": (ctypes.c_int, [_VP, _VP, _VP, ctypes.c_int, _VP, _VP]),
    "cb_time_resident": (ctypes.c_int, [_VP, ctypes.c_int, ctypes.c_int, ctypes.c_int, _VP]),
    "cb_time_sustained": (ctypes.c_int, [_VP, ctypes.c_int, ctypes.c_double, ctypes.POINTER(ctypes.c_float),
                                         ctypes.POINTER(ctypes.c_int)]),
    "cb_time_conv_layers": (ctypes.c_int, [_VP, ctypes.c_int, _VP, _VP]),
    "cb_time_conv_layers_detail": (ctypes.c_int, [_VP, ctypes.c_int, _VP, _VP, _VP]),
    "cb_time_tower": (ctypes.c_int, [_VP, ctypes.c_int, _VP, _VP]),
    "cb_read_resident": (ctypes.c_int, [_VP, _VP, _VP, _VP, _VP]),
    "cb_debug_backbone": (ctypes.c_int, [_VP, ctypes.c_int, _VP]),
    "cb_debug_act": (ctypes.c_int, [_VP, ctypes.c_int, ctypes.c_int, _VP]),
    "cb_debug_tower_trace": (ctypes.c_int, [_VP, _VP, ctypes.c_int]),
    "cb_debug_cta_times": (ctypes.c_int, [_VP, _VP, ctypes.c_int]),
    "cb_debug_set_layer_limit": (ctypes.c_int, [_VP, ctypes.c_int]),
    "cb_chess_startpos": (ctypes.c_int, [_VP]),
    "cb_chess_legal_moves": (ctypes.c_int, [_VP, _VP, _VP]),
    "cb_chess_make_move": (ctypes.c_int, [_VP, ctypes.c_uint16, _VP]),
    "cb_chess_perft": (ctypes.c_uint64, [_VP, ctypes.c_int]),
    "cb_selfplay_run": (ctypes.c_int, [_VP, ctypes.POINTER(SelfplayConfig), ctypes.POINTER(SelfplayStats)]),
    "cb_match_run": (ctypes.c_int, [_VP, _VP, ctypes.POINTER(MatchConfig), ctypes.POINTER(MatchStats)]),
    "cb_sprt_llr": (ctypes.c_int, [ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_double, ctypes.c_double,
                                   ctypes.POINTER(ctypes.c_double)]),
    "cb_sprt_decision": (ctypes.c_int, [ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_double, ctypes.c_double,
                                        ctypes.c_double, ctypes.c_double]),
    "cb_selfplay_device": (ctypes.c_int, [_VP, ctypes.POINTER(SelfplayConfig), ctypes.POINTER(SelfplayStats)]),
    "cb_debug_host_search": (ctypes.c_int, [ctypes.POINTER(SelfplayConfig), _VP, ctypes.c_int, EVAL_CALLBACK, _VP,
                                            _VP, _VP, _VP, _VP, _VP]),
    "cb_debug_host_game": (ctypes.c_int, [ctypes.POINTER(SelfplayConfig), _VP, ctypes.c_uint64, EVAL_CALLBACK, _VP,
                                          _VP, _VP, _VP, _VP, _VP, ctypes.c_int, ctypes.POINTER(ctypes.c_int),
                                          ctypes.POINTER(ctypes.c_int)]),
    "cb_debug_device_search": (ctypes.c_int, [_VP, ctypes.POINTER(SelfplayConfig), _VP, ctypes.c_int, _VP, _VP, _VP, _VP, _VP]),
    "cb_launches_per_forward": (ctypes.c_int, [_VP]),
    "cb_get_timing": (ctypes.c_int, [_VP, ctypes.POINTER(CbTiming)]),
    "cb_reset_timing": (None, [_VP]),
    "cb_last_error": (ctypes.c_char_p, [_VP]),
    "cb_version": (ctypes.c_char_p, []),
    "cb_destroy": (None, [_VP]),
}


def lib_path():
    return _LIB_PATH


def load_library():
    """dlopen the CUDA library (fails loudly if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise CaissaError("libcaissa_b200.so not built: run `python neurosymbolic-mcts_b200/build.py`")
        L = ctypes.CDLL(_LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_VP)


def _boards_u8(boards):
    b = np.ascontiguousarray(boards, dtype=np.uint8)
    if b.ndim != 2 or b.shape[1] != 128:
        raise ValueError("boards must be uint8 [B,128] packed cb_board records")
    return b


class Evaluator:
    """Thin owner of one cb_handle (one CUDA stream, one set of device buffers)."""

    def __init__(self, num_blocks=6, hidden=128, dtype="bf16", device=0):
        self._L = load_library()
        self._h = _VP()
        self.num_blocks, self.hidden = num_blocks, hidden
        self.dtype = DTYPES[dtype] if isinstance(dtype, str) else int(dtype)
        rc = self._L.cb_create(ctypes.byref(self._h), device, num_blocks, hidden, self.dtype)
        if rc != 0:
            self._h = _VP()
            raise CaissaError("cb_create failed with %d (no CUDA device / unsupported shape; there is no CPU fallback)" % rc)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.cb_destroy(self._h)
            self._h = _VP()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise CaissaError("caissa_b200 error %d: %s" % (rc, (self._L.cb_last_error(self._h) or b"").decode()))

    def weight_count(self):
        return self._L.cb_weight_count(self.num_blocks, self.hidden)

    def load_weights(self, blob):
        blob = np.ascontiguousarray(blob, dtype=np.float32)
        self._check(self._L.cb_load_weights(self._h, _ptr(blob), blob.size))

    def is_available(self):
        return bool(self._L.cb_is_available(self._h))

    def encode(self, boards):
        b = _boards_u8(boards)
        out = np.empty((b.shape[0], 17, 8, 8), dtype=np.float32)
        self._check(self._L.cb_encode(self._h, _ptr(b), b.shape[0], _ptr(out)))
        return out

    def pst_eval_cp(self, boards):
        b = _boards_u8(boards)
        out = np.empty(b.shape[0], dtype=np.int32)
        self._check(self._L.cb_pst_eval_cp(self._h, _ptr(b), b.shape[0], _ptr(out)))
        return out

    def debug_device_legal_moves(self, boards):
        """The tree kernels' legal-move lists: (moves uint16 [n,256], counts int32 [n])."""
        b = _boards_u8(boards)
        moves = np.zeros((b.shape[0], 256), dtype=np.uint16)
        counts = np.zeros(b.shape[0], dtype=np.int32)
        self._check(self._L.cb_debug_device_legal_moves(self._h, _ptr(b), b.shape[0], _ptr(moves), _ptr(counts)))
        return moves, counts

    def debug_device_light_gates(self, boards, koth=False):
        """The tree kernels' light gates (warp form): int8 per board (+1, -1, 2 = none)."""
        b = _boards_u8(boards)
        out = np.zeros(b.shape[0], dtype=np.int8)
        self._check(self._L.cb_debug_device_light_gates(self._h, _ptr(b), b.shape[0], int(koth), _ptr(out)))
        return out

    def principal_exchange(self, boards):
        """Principal-exchange material balance of every board, on the device (pawn units, f32)."""
        b = _boards_u8(boards)
        out = np.empty(b.shape[0], dtype=np.float32)
        self._check(self._L.cb_principal_exchange_device(self._h, _ptr(b), b.shape[0], _ptr(out)))
        return out

    def predict_batch(self, boards, q=None, qflag=None, want_policy=True):
        b = _boards_u8(boards)
        B = b.shape[0]
        qa = None if q is None else np.ascontiguousarray(q, dtype=np.float32)
        fa = None if qflag is None else np.ascontiguousarray(qflag, dtype=np.uint8)
        pol = np.empty((B, POLICY_SIZE), dtype=np.float32) if want_policy else None
        v = np.empty(B, dtype=np.float32)
        k = np.empty(B, dtype=np.float32)
        self._check(self._L.cb_predict_batch(self._h, _ptr(b), _ptr(qa), _ptr(fa), B, _ptr(pol), _ptr(v), _ptr(k)))
        return pol, v, k

    def eval_leaves(self, boards, q, move_idx, move_off, material_on=True):
        b = _boards_u8(boards)
        B = b.shape[0]
        qa = None if q is None else np.ascontiguousarray(q, dtype=np.float32)
        mi = np.ascontiguousarray(move_idx, dtype=np.uint16)
        mo = np.ascontiguousarray(move_off, dtype=np.uint32)
        assert mo.size == B + 1
        pri = np.empty(int(mo[B]), dtype=np.float32)
        v = np.empty(B, dtype=np.float32)
        vf = np.empty(B, dtype=np.float32)
        k = np.empty(B, dtype=np.float32)
        self._check(self._L.cb_eval_leaves(self._h, _ptr(b), _ptr(qa), B, int(material_on), _ptr(mi), _ptr(mo),
                                           _ptr(pri), _ptr(v), _ptr(vf), _ptr(k)))
        return pri, v, vf, k

    def prepared_eval_leaves(self, boards, q, move_idx, move_off, material_on=True):
        """cb_eval_leaves on fixed host buffers: returns run() -> (priors, v_logit, v_final, k), which calls the C ABI
        with argument pointers and output arrays made once (what a host that reuses its request buffers does; the
        per-call cost of the numpy / ctypes marshalling above is ~20 us).  Every run() still copies the inputs host ->
        device and the results device -> host inside cb_eval_leaves."""
        b = _boards_u8(boards)
        B = b.shape[0]
        qa = None if q is None else np.ascontiguousarray(q, dtype=np.float32)
        mi = np.ascontiguousarray(move_idx, dtype=np.uint16)
        mo = np.ascontiguousarray(move_off, dtype=np.uint32)
        assert mo.size == B + 1
        outs = (np.empty(int(mo[B]), dtype=np.float32), np.empty(B, dtype=np.float32), np.empty(B, dtype=np.float32),
                np.empty(B, dtype=np.float32))
        keep = (b, qa, mi, mo)                                    # the pointers below borrow these arrays
        args = (self._h, _ptr(b), _ptr(qa), B, int(material_on), _ptr(mi), _ptr(mo)) + tuple(_ptr(o) for o in outs)
        fn, check = self._L.cb_eval_leaves, self._check

        def run(_keep=keep):
            check(fn(*args))
            return outs
        return run

    def prepared_submit_wait(self, boards, q, move_idx, move_off, material_on=True):
        """cb_eval_leaves_submit / cb_eval_leaves_wait on fixed host buffers -> (submit(), wait() -> outputs): the two
        halves of prepared_eval_leaves, for a single host thread that keeps several handles busy."""
        b = _boards_u8(boards)
        B = b.shape[0]
        qa = None if q is None else np.ascontiguousarray(q, dtype=np.float32)
        mi = np.ascontiguousarray(move_idx, dtype=np.uint16)
        mo = np.ascontiguousarray(move_off, dtype=np.uint32)
        assert mo.size == B + 1
        outs = (np.empty(int(mo[B]), dtype=np.float32), np.empty(B, dtype=np.float32), np.empty(B, dtype=np.float32),
                np.empty(B, dtype=np.float32))
        keep = (b, qa, mi, mo)
        sargs = (self._h, _ptr(b), _ptr(qa), B, int(material_on), _ptr(mi), _ptr(mo))
        wargs = (self._h,) + tuple(_ptr(o) for o in outs)
        fs, fw, check = self._L.cb_eval_leaves_submit, self._L.cb_eval_leaves_wait, self._check

        def submit(_keep=keep):
            check(fs(*sargs))

        def wait():
            check(fw(*wargs))
            return outs
        return submit, wait

    def stage_leaves(self, boards, q, move_idx, move_off):
        b = _boards_u8(boards)
        qa = None if q is None else np.ascontiguousarray(q, dtype=np.float32)
        mi = None if move_idx is None else np.ascontiguousarray(move_idx, dtype=np.uint16)
        mo = None if move_off is None else np.ascontiguousarray(move_off, dtype=np.uint32)
        self._staged = (b.shape[0], 0 if mo is None else int(mo[-1]))
        self._check(self._L.cb_stage_leaves(self._h, _ptr(b), _ptr(qa), b.shape[0], _ptr(mi), _ptr(mo)))

    def time_resident(self, iters, material_on=True, flush_l2=False):
        ms = np.zeros(iters, dtype=np.float32)
        self._check(self._L.cb_time_resident(self._h, int(material_on), iters, int(flush_l2), _ptr(ms)))
        return ms

    def time_sustained(self, seconds, material_on=True):
        """-> (total device ms, launches) of back-to-back replays of the staged forward for >= seconds."""
        ms, n = ctypes.c_float(0), ctypes.c_int(0)
        self._check(self._L.cb_time_sustained(self._h, int(material_on), float(seconds), ctypes.byref(ms), ctypes.byref(n)))
        return ms.value, n.value

    def time_conv_layers(self, iters):
        ms = ctypes.c_float(0)
        n = ctypes.c_int(0)
        self._check(self._L.cb_time_conv_layers(self._h, iters, ctypes.byref(ms), ctypes.byref(n)))
        return ms.value, n.value

    def time_conv_layers_detail(self, iters):
        per = np.zeros(2 * self.num_blocks + 2, dtype=np.float32)
        self._check(self._L.cb_time_conv_layers_detail(self._h, iters, None, None, _ptr(per)))
        return per

    def time_tower(self, iters):
        ms = np.zeros(iters, dtype=np.float32)
        n = ctypes.c_int(0)
        self._check(self._L.cb_time_tower(self._h, iters, _ptr(ms), ctypes.byref(n)))
        return ms, n.value

    def read_resident(self):
        B, nm = self._staged
        pri = np.empty(nm, dtype=np.float32)
        v = np.empty(B, dtype=np.float32)
        vf = np.empty(B, dtype=np.float32)
        k = np.empty(B, dtype=np.float32)
        self._check(self._L.cb_read_resident(self._h, _ptr(pri), _ptr(v), _ptr(vf), _ptr(k)))
        return pri, v, vf, k

    def debug_backbone(self, B):
        out = np.empty((B, self.hidden, 8, 8), dtype=np.float32)
        self._check(self._L.cb_debug_backbone(self._h, B, _ptr(out)))
        return out

    def debug_act(self, B, buf):
        out = np.empty((B, self.hidden, 8, 8), dtype=np.float32)
        self._check(self._L.cb_debug_act(self._h, B, buf, _ptr(out)))
        return out

    def debug_tower_trace(self):
        items = 2 * (2 * self.num_blocks + 3)
        out = np.zeros((items, 8), dtype=np.int64)
        n = self._L.cb_debug_tower_trace(self._h, _ptr(out), items)
        if n < 0:
            self._check(n)
        return out[:n]

    def debug_cta_times(self, max_ctas=256):
        out = np.zeros((max_ctas, 4), dtype=np.uint64)
        n = self._L.cb_debug_cta_times(self._h, _ptr(out), max_ctas)
        if n < 0:
            self._check(n)
        return out[:n]

    def debug_set_layer_limit(self, n):
        self._check(self._L.cb_debug_set_layer_limit(self._h, n))

    def cap1_qsearch_device(self, boards):
        """The cap = 1 quiescence search on the GPU, one warp per board -> (q f32 pawn units, completed u8)."""
        b = _boards_u8(boards)
        q = np.empty(b.shape[0], dtype=np.float32)
        c = np.empty(b.shape[0], dtype=np.uint8)
        self._check(self._L.cb_cap1_qsearch_device(self._h, _ptr(b), b.shape[0], _ptr(q), _ptr(c)))
        return q, c

    def debug_device_tier1_gates(self, boards, koth=False, mate_depth=5, exhaustive_depth=0, koth_depth=3, budget=0):
        """The tree kernels' gate at full depth (warp form) -> (int8 value per board: +1 / -1 / 2, uint8 terminal_distance)."""
        b = _boards_u8(boards)
        out = np.empty(b.shape[0], dtype=np.int8)
        dist = np.empty(b.shape[0], dtype=np.uint8)
        self._check(self._L.cb_debug_device_tier1_gates(self._h, _ptr(b), b.shape[0], int(koth), int(mate_depth), int(exhaustive_depth),
                                                        int(koth_depth), int(budget), _ptr(out), _ptr(dist)))
        return out, dist

    def selfplay(self, games=1024, sims_per_move=200, max_plies=200, material_on=False, koth=False, threads=0,
                 c_puct=1.414, explore_base=0.80, seed=1, max_sims=0, max_games=0, max_seconds=0.0, sample_path=None,
                 pipeline=1, aux=None, qsearch=0, tier1_light=False, shuffle_children=False, mock=None, tier1_mate_depth=0,
                 tier1_exhaustive_depth=0, tier1_koth_depth=0):
        """Run the self-play pool on this evaluator; returns the stats as a dict.  `aux`: a second Evaluator
        with the same weights that serves the second half-pool in pipeline mode 2.  `qsearch`: 0 stand-pat
        material, 1 principal exchange ("pe").  `mock`: (v_logit, k) replaces the network by the reference's mock
        evaluator (device pipelines)."""
        cfg = selfplay_config(games=games, sims_per_move=sims_per_move, max_plies=max_plies, material_on=material_on, koth=koth,
                              threads=threads, c_puct=c_puct, explore_base=explore_base, seed=seed, max_sims=max_sims,
                              max_games=max_games, max_seconds=max_seconds, sample_path=sample_path, pipeline=pipeline,
                              aux=aux, qsearch=qsearch, tier1_light=tier1_light, shuffle_children=shuffle_children, mock=mock,
                              tier1_mate_depth=tier1_mate_depth, tier1_exhaustive_depth=tier1_exhaustive_depth,
                              tier1_koth_depth=tier1_koth_depth)
        st = SelfplayStats()
        self._check(self._L.cb_selfplay_run(self._h, ctypes.byref(cfg), ctypes.byref(st)))
        return {n: getattr(st, n) for n, _ in SelfplayStats._fields_}

    def debug_device_search(self, boards, **kw):
        """One move's search with the device tree kernels from every position -> list of [(move, visits, value_sum)],
        root visit counts (cb_debug_device_search)."""
        boards = np.ascontiguousarray(boards, dtype=np.uint8).reshape(-1, 128)
        n = boards.shape[0]
        cfg = selfplay_config(games=n, pipeline=3, **kw)
        out = _SearchOut(n)
        self._check(self._L.cb_debug_device_search(self._h, ctypes.byref(cfg), boards.ctypes.data, n, *out.ptrs()))
        return out.unpack()

    def match(self, current, games=256, total_games=256, sims_per_move=200, max_plies=200, material_on=False, koth=False,
              c_puct=1.414, explore_base=0.90, seed=1, use_sprt=False, elo0=0.0, elo1=10.0, alpha=0.05, beta=0.05,
              max_seconds=0.0, qsearch=0, tier1_light=False):
        """Evaluation match: this evaluator is the candidate, `current` the incumbent (cb_match_run)."""
        cfg = MatchConfig(games, total_games, sims_per_move, max_plies, int(material_on), int(koth), c_puct, explore_base,
                          seed, int(use_sprt), elo0, elo1, alpha, beta, max_seconds, _QSEARCH[qsearch], int(tier1_light))
        st = MatchStats()
        self._check(self._L.cb_match_run(self._h, current._h, ctypes.byref(cfg), ctypes.byref(st)))
        return {n: getattr(st, n) for n, _ in MatchStats._fields_}

    def launches_per_forward(self):
        return self._L.cb_launches_per_forward(self._h)

    def timing(self):
        t = CbTiming()
        self._check(self._L.cb_get_timing(self._h, ctypes.byref(t)))
        return {"calls": t.calls, "positions": t.positions, "h2d_ms": t.h2d_ms, "forward_ms": t.forward_ms,
                "d2h_ms": t.d2h_ms}


_QSEARCH = {0: 0, 1: 1, 2: 2, "static": 0, "stand-pat": 0, "pe": 1, "principal-exchange": 1, "cap1": 2}


def principal_exchange(boards, want_nodes=False):
    """Principal-exchange material balance on the host (cb_principal_exchange): f32 pawn units [, nodes]."""
    b = _boards_u8(boards)
    out = np.empty(b.shape[0], dtype=np.float32)
    nodes = np.empty(b.shape[0], dtype=np.uint32)
    rc = load_library().cb_principal_exchange(_ptr(b), b.shape[0], _ptr(out), _ptr(nodes))
    if rc:
        raise CaissaError("cb_principal_exchange failed")
    return (out, nodes) if want_nodes else out


def light_gates(boards, koth=False):
    """Light Tier-1 gates (cb_light_gates): int8 per board, +1 / -1 for the side to move, 2 = no gate."""
    b = _boards_u8(boards)
    out = np.empty(b.shape[0], dtype=np.int8)
    if load_library().cb_light_gates(_ptr(b), b.shape[0], int(koth), _ptr(out)):
        raise CaissaError("cb_light_gates failed")
    return out


def tier1_gates(boards, koth=False, mate_depth=5, exhaustive_depth=0, koth_depth=3):
    """The Tier-1 leaf gate at full depth (cb_tier1_gates): (int8 value per board: +1 / -1 / 2 = none, uint8 terminal_distance)."""
    b = _boards_u8(boards)
    out = np.empty(b.shape[0], dtype=np.int8)
    dist = np.empty(b.shape[0], dtype=np.uint8)
    if load_library().cb_tier1_gates(_ptr(b), b.shape[0], int(koth), int(mate_depth), int(exhaustive_depth), int(koth_depth),
                                     _ptr(out), _ptr(dist)):
        raise CaissaError("cb_tier1_gates failed")
    return out, dist


def cap1_qsearch(boards):
    """The cap = 1 quiescence search on the host (cb_cap1_qsearch) -> (q f32 pawn units, completed u8, nodes u32, depth u8)."""
    b = _boards_u8(boards)
    n = b.shape[0]
    q, c = np.empty(n, dtype=np.float32), np.empty(n, dtype=np.uint8)
    nodes, d = np.empty(n, dtype=np.uint32), np.empty(n, dtype=np.uint8)
    if load_library().cb_cap1_qsearch(_ptr(b), n, _ptr(q), _ptr(c), _ptr(nodes), _ptr(d)):
        raise CaissaError("cb_cap1_qsearch failed")
    return q, c, nodes, d


def debug_move_flags(boards):
    """(disagreements, moves compared) of move_flags_hd against make-move + attack test (cb_debug_move_flags)."""
    b = _boards_u8(boards)
    n = ctypes.c_longlong(0)
    bad = load_library().cb_debug_move_flags(_ptr(b), b.shape[0], ctypes.byref(n))
    return int(bad), int(n.value)


def mate_search(boards, max_depth=5, exhaustive_depth=2):
    """mate_search of the reference (cb_mate_search): plies of the forced mate found per board (odd), 0 = none."""
    b = _boards_u8(boards)
    out = np.empty(b.shape[0], dtype=np.int32)
    if load_library().cb_mate_search(_ptr(b), b.shape[0], int(max_depth), int(exhaustive_depth), _ptr(out)):
        raise CaissaError("cb_mate_search failed")
    return out


def koth_center_in_n(boards, max_n=3):
    """koth_center_in_n (cb_koth_center_in_n): own moves to the hill against every defence per board, -1 = None."""
    b = _boards_u8(boards)
    out = np.empty(b.shape[0], dtype=np.int8)
    if load_library().cb_koth_center_in_n(_ptr(b), b.shape[0], int(max_n), _ptr(out)):
        raise CaissaError("cb_koth_center_in_n failed")
    return out


def sprt_llr(wins, losses, draws, elo0=0.0, elo1=10.0):
    """SprtState::compute_llr (src/mcts/sprt.rs): the LLR, or None when it is undefined."""
    v = ctypes.c_double(0.0)
    rc = load_library().cb_sprt_llr(wins, losses, draws, elo0, elo1, ctypes.byref(v))
    return v.value if rc == 1 else None


def sprt_decision(wins, losses, draws, elo0=0.0, elo1=10.0, alpha=0.05, beta=0.05):
    """SprtState::check_decision: 'H1', 'H0' or 'inconclusive'."""
    return {1: "H1", -1: "H0", 0: "inconclusive"}[load_library().cb_sprt_decision(wins, losses, draws, elo0, elo1, alpha, beta)]


def chess_startpos():
    b = np.zeros(128, dtype=np.uint8)
    load_library().cb_chess_startpos(_ptr(b))
    return b


def chess_legal_moves(board):
    b = np.ascontiguousarray(board, dtype=np.uint8)
    mv = np.zeros(MAX_MOVES, dtype=np.uint16)
    idx = np.zeros(MAX_MOVES, dtype=np.uint16)
    n = load_library().cb_chess_legal_moves(_ptr(b), _ptr(mv), _ptr(idx))
    return mv[:n].copy(), idx[:n].copy()


def chess_make_move(board, move):
    b = np.ascontiguousarray(board, dtype=np.uint8)
    out = np.zeros(128, dtype=np.uint8)
    load_library().cb_chess_make_move(_ptr(b), int(move), _ptr(out))
    return out


def chess_perft(board, depth):
    b = np.ascontiguousarray(board, dtype=np.uint8)
    return load_library().cb_chess_perft(_ptr(b), depth)


class NeuralNetPolicy:
    """Mirror of the reference's `NeuralNetPolicy` (src/neural_net.rs:86-363): same method names,
    argument meaning and error behaviour (errors answer None per item, never raise from predict*).

    `load` takes a flat f32 weight blob (.bin, or an ndarray) instead of a TorchScript file; boards
    are packed 128-byte records (uint8 [128])."""

    def __init__(self, num_blocks=6, hidden=128, dtype="bf16", device=0):
        self._cfg = (num_blocks, hidden, dtype, device)
        self._ev = None
        self._warned = False

    def load(self, path_or_blob):
        try:
            if isinstance(path_or_blob, (str, os.PathLike)):
                if not os.path.exists(path_or_blob):
                    return "Model file not found: %s" % path_or_blob
                blob = np.fromfile(path_or_blob, dtype=np.float32)
            else:
                blob = np.asarray(path_or_blob, dtype=np.float32)
            ev = Evaluator(*self._cfg)
            ev.load_weights(blob)
            self._ev = ev
            return None  # Ok(())
        except Exception as e:  # Err(String)
            return "Failed to load model: %s" % e

    def is_available(self):
        return self._ev is not None and self._ev.is_available()

    def predict(self, board, qsearch_completed, q_result):
        res = self.predict_batch([board], [qsearch_completed], [q_result])
        return res[0]

    def predict_batch(self, boards, qsearch_flags, q_results):
        n = len(boards)
        if self._ev is None:
            return [None] * n
        try:
            b = np.stack([np.frombuffer(bytes(x), dtype=np.uint8) if not isinstance(x, np.ndarray) else x for x in boards])
            pol, v, k = self._ev.predict_batch(b, np.asarray(q_results, dtype=np.float32),
                                               np.asarray(qsearch_flags, dtype=np.uint8))
            return [(pol[i], float(v[i]), float(k[i])) for i in range(n)]
        except Exception as e:
            if not self._warned:
                print("[NN-ERROR] predict_batch failed: %s" % e)
                self._warned = True
            return [None] * n

    def timing(self):
        return self._ev.timing() if self._ev else None

    def print_inference_timing(self):
        t = self.timing()
        if t and t["calls"]:
            c = t["calls"]
            print("NN timing over %d calls: to-gpu %.3f ms, forward %.3f ms, from-gpu %.3f ms (means)"
                  % (c, t["h2d_ms"] / c, t["forward_ms"] / c, t["d2h_ms"] / c))


class InferenceServer:
    """Mirror of src/mcts/inference_server.rs:29-140: one worker thread owns the evaluator, blocks on
    the first request, then gathers more until `batch_size` or 500 us, evaluates, and answers each
    request on its own one-slot reply queue with `(policy, v_logit, k)` or None."""

    def __init__(self, nn, batch_size):
        self._q = queue.Queue()
        self._nn = nn
        self._batch = batch_size
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def _run(self):
        while True:
            first = self._q.get()
            if first is None:
                break
            reqs = [first]
            t0 = time.perf_counter()
            while len(reqs) < self._batch and time.perf_counter() - t0 < 500e-6:
                try:
                    r = self._q.get_nowait()
                except queue.Empty:
                    time.sleep(0)
                    continue
                if r is None:
                    self._q.put(None)
                    break
                reqs.append(r)
            res = self._nn.predict_batch([r[0] for r in reqs], [r[1] for r in reqs], [r[2] for r in reqs])
            for r, out in zip(reqs, res):
                r[3].put(out)
        self._nn.print_inference_timing()

    def predict_async(self, board, qsearch_completed, q_result):
        reply = queue.Queue(maxsize=1)
        self._q.put((board, qsearch_completed, q_result, reply))
        return reply

    def shutdown(self):
        self._q.put(None)
        self._t.join(timeout=5)
