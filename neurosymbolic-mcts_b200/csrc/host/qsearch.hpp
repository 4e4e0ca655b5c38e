// qsearch.hpp — the material term dM of the leaf evaluation: the static PeSTO stand pat (src/eval.rs:93-122) and
// the principal-exchange quiescence line on top of it (src/search/quiescence.rs:756-804, the reference's
// `--qsearch pe` variant of src/mcts/tactical_mcts.rs:715-731).  One statement of the algorithm for the host
// self-play driver (selfplay.cpp, compiled by g++) and for the device tree kernels (mcts_device.cuh, compiled by
// nvcc): a scalar form, and a warp form built from the same per-piece routine.
//
// The exchange line takes, at every ply, the FIRST LEGAL move of the reference's MVV-LVA sorted capture list
// (src/move_generation.rs:426-446).  No list is built or sorted here: every capture gets a key
//     (45 - score) << 16 | generation rank          score = 10 * victim - attacker, 0 on an empty target square
// whose ascending order IS the sorted order (ties keep generation order: the reference's sort_unstable is an
// insertion sort up to 20 elements), and the line repeatedly takes the smallest key above the last rejected one.
// Generation rank = class << 13 | from << 7 | sub with classes in the reference's concatenation order
// (src/move_generation.rs:287-309): 0 pawn captures, 1 knights, 2 bishops, 3 rooks, 4 queens, 5 kings,
// 6 promotions (capturing or not); `sub` orders one piece's moves as the reference's tables / bit loops do.
//
// Product code: it does not use oracle/.  tests/ check it against oracle/qsearch.c (an independent list + sort
// statement) through cb_principal_exchange, and the warp form against the scalar form through device self-play.
#pragma once
#include "chess_rules.hpp"
#include "../pesto_tables.h"

// The PeSTO tables are __device__ objects in a CUDA translation unit and plain statics elsewhere, so these
// functions are device functions under nvcc and host functions under g++.
#if defined(__CUDACC__)
#define CB_Q __device__ __forceinline__
#else
#define CB_Q inline
#endif

namespace cbchess {

CB_Q int pesto_cp(const cb_board& b) {   // src/eval.rs:93-122: tapered PeSTO, side-to-move relative, truncating /24
    int mg[2] = {0, 0}, eg[2] = {0, 0}, phase = 0;
    for (int c = 0; c < 2; ++c)
        for (int pt = 0; pt < 6; ++pt)
            for (uint64_t bb = b.pieces[c * 6 + pt]; bb; bb &= bb - 1) {
                const int sq = lsb(bb), t = c == 0 ? (sq ^ 56) : sq;
                mg[c] += CB_MG_VALUE[pt] + CB_MG_PST[pt][t];
                eg[c] += CB_EG_VALUE[pt] + CB_EG_PST[pt][t];
                phase += CB_PHASE_INC[pt];
            }
    const int mgp = phase < 24 ? phase : 24;
    const int score = ((mg[0] - mg[1]) * mgp + (eg[0] - eg[1]) * (24 - mgp)) / 24;
    return b.w_to_move ? score : -score;
}

constexpr uint32_t kPeNoCapture = 0xFFFFFFFFu;
constexpr int kPeMaxDepth = 20;        // src/search/quiescence.rs:802
constexpr int kPeWindow = 100000;

CB_Q uint32_t pe_key(int score, int cls, int from, int sub) {
    return (uint32_t)(45 - score) << 16 | (uint32_t)cls << 13 | (uint32_t)from << 7 | (uint32_t)sub;
}

// The smallest-key capture or promotion of the piece of type `pt` on `from` whose key is above `above`
// (kPeNoCapture if there is none); *mv receives the move.  `above` = 0 accepts everything (keys are >= 5 << 16).
CB_Q uint32_t pe_piece_best(const cb_board& b, int from, int pt, uint32_t above, Move* mv) {
    const Tables& T = tables();
    const int us = b.w_to_move ? 0 : 1, them = 1 - us;
    const uint64_t opp = b.occ[them], occ = b.occ[0] | b.occ[1];
    uint32_t best = kPeNoCapture;
    Move bm = 0;
#define CB_PE_TRY(key_, move_)                          \
    do {                                                \
        const uint32_t k_ = (key_);                     \
        if (k_ > above && k_ < best) { best = k_; bm = (move_); } \
    } while (0)
    if (pt == PAWN) {
        const int r = from >> 3;
        const bool promo = us == 0 ? r == 6 : r == 1;
        for (uint64_t tg = T.pawn[us][from] & (opp | (b.en_passant != 0xFF ? 1ull << b.en_passant : 0ull)); tg; tg &= tg - 1) {
            const int to = lsb(tg);
            // table order (src/magic_bitboard.rs:214-229): the a-file side target first
            const int which = (to & 7) < (from & 7) ? 0 : 1;
            const int victim = piece_on(b, them, 1ull << to);
            const int score = victim < 0 ? 0 : 10 * victim;
            if (promo) {
                const int pp[4] = {QUEEN, ROOK, KNIGHT, BISHOP};   // src/magic_bitboard.rs:457-474
                for (int i = 0; i < 4; ++i) CB_PE_TRY(pe_key(score, 6, from, which * 4 + i), make_mv(from, to, pp[i]));
            } else {
                CB_PE_TRY(pe_key(score, 0, from, which), make_mv(from, to, 0));
            }
        }
        if (promo) {
            const int one = us == 0 ? from + 8 : from - 8;
            if (!(occ >> one & 1)) {
                const int pp[4] = {QUEEN, ROOK, KNIGHT, BISHOP};
                for (int i = 0; i < 4; ++i) CB_PE_TRY(pe_key(0, 6, from, 8 + i), make_mv(from, one, pp[i]));
            }
        }
    } else if (pt == KNIGHT || pt == KING) {
        // table order of src/magic_bitboard.rs:160-189 (knight) and :120-149 (king), as square deltas
        const int nd[8] = {15, 17, 6, 10, -10, -6, -17, -15}, kd[8] = {-1, 7, 8, 9, 1, -7, -8, -9};
        for (uint64_t tg = (pt == KNIGHT ? T.knight[from] : T.king[from]) & opp; tg; tg &= tg - 1) {
            const int to = lsb(tg), d = to - from;
            int idx = 0;
            for (int i = 0; i < 8; ++i)
                if ((pt == KNIGHT ? nd[i] : kd[i]) == d) idx = i;
            CB_PE_TRY(pe_key(10 * piece_on(b, them, 1ull << to) - pt, pt == KNIGHT ? 1 : 5, from, idx), make_mv(from, to, 0));
        }
    } else {
        // sliders: attack set from the lowest square up; the queen's rook set before its bishop set
        const uint64_t ra = pt != BISHOP ? rook_att(from, occ) & opp : 0ull;
        const uint64_t ba = pt != ROOK ? bishop_att(from, occ) & opp : 0ull;
        const int cls = pt == BISHOP ? 2 : pt == ROOK ? 3 : 4;
        for (uint64_t tg = ra; tg; tg &= tg - 1) {
            const int to = lsb(tg);
            CB_PE_TRY(pe_key(10 * piece_on(b, them, 1ull << to) - pt, cls, from, to), make_mv(from, to, 0));
        }
        for (uint64_t tg = ba; tg; tg &= tg - 1) {
            const int to = lsb(tg);
            CB_PE_TRY(pe_key(10 * piece_on(b, them, 1ull << to) - pt, cls, from, (pt == QUEEN ? 64 : 0) + to), make_mv(from, to, 0));
        }
    }
#undef CB_PE_TRY
    *mv = bm;
    return best;
}

// Unwind of the straight line (src/search/quiescence.rs:781-793): level d entered its child with its raised
// alpha ap[d]; its beta is the negated raised alpha of the level above (100000 at the root).
CB_Q int pe_unwind(const int* ap, int depth, int ret) {
    for (int d = depth - 1; d >= 0; --d) {
        const int beta = d == 0 ? kPeWindow : -ap[d - 1];
        const int s = -ret;
        ret = s >= beta ? beta : (s > ap[d] ? s : ap[d]);
    }
    return ret;
}

// forced_principal_exchange in centipawns (the caller divides by 100 in f32, src/search/quiescence.rs:803).
CB_Q int principal_exchange_cp(const cb_board& root, uint32_t* nodes_out = nullptr) {
    cb_board b = root;
    int ap[kPeMaxDepth + 1];
    int alpha = -kPeWindow, beta = kPeWindow, depth = 0, ret;
    uint32_t nodes = 0;
    for (;;) {
        ++nodes;
        const int sp = pesto_cp(b);
        if (sp >= beta) { ret = beta; break; }
        if (sp > alpha) alpha = sp;
        if (depth == kPeMaxDepth) { ret = alpha; break; }
        const int us = b.w_to_move ? 0 : 1;
        cb_board nb;
        bool found = false;
        for (uint32_t above = 0;;) {
            uint32_t best = kPeNoCapture;
            Move bm = 0;
            for (int pt = 0; pt < 6; ++pt)
                for (uint64_t bb = b.pieces[us * 6 + pt]; bb; bb &= bb - 1) {
                    Move m;
                    const uint32_t k = pe_piece_best(b, lsb(bb), pt, above, &m);
                    if (k < best) { best = k; bm = m; }
                }
            if (best == kPeNoCapture) break;
            make_move_hd(b, bm, &nb);
            if (!in_check_hd(nb, us)) { found = true; break; }   // Board::is_legal, src/board.rs:352
            above = best;
        }
        if (!found) { ret = alpha; break; }
        ap[depth++] = alpha;
        const int na = -beta;
        beta = -alpha;
        alpha = na;
        b = nb;
    }
    if (nodes_out) *nodes_out = nodes;
    return pe_unwind(ap, depth, ret);
}

#if !defined(__CUDACC__)
// The cap = 1 quiescence search (the reference's `--qsearch cap1`, src/search/quiescence.rs:808-1012), host form: at a node
// that is not in check the stand pat, the FIRST LEGAL capture of the MVV-LVA order (the same keys as the exchange line) and -
// once per side and branch - the first legal checking move of the generation order that is not that capture; in check every
// legal evasion.  Window (-100000, 100000), depth 20 (:1000-1012).  "That capture" is identified the way the reference does
// it (:933-944): by the from-square of the first entry of the capture list whose target is the played capture's target - an
// illegal capture of the same target that sorts earlier makes the test miss, as written.
struct Cap1Result {
    int score;
    bool completed;
    uint32_t nodes;
    int depth;
};
inline Cap1Result cap1_qsearch(const cb_board& b, int alpha, int beta, int max_depth, bool w_check_used, bool b_check_used) {
    Cap1Result r = {0, true, 1, 0};
    const int us = b.w_to_move ? 0 : 1;
    const bool in_check = in_check_hd(b, us);
    if (!in_check) {
        const int sp = pesto_cp(b);
        if (sp >= beta) { r.score = beta; return r; }
        if (sp > alpha) alpha = sp;
    }
    if (max_depth == 0) { r.score = alpha; r.completed = !in_check; return r; }
    cb_board nb;
    auto child = [&](bool wu, bool bu) -> bool {   // searches `nb`; true = beta cut-off (r.score is set)
        const Cap1Result c = cap1_qsearch(nb, -beta, -alpha, max_depth - 1, wu, bu);
        r.nodes += c.nodes;
        if (c.depth + 1 > r.depth) r.depth = c.depth + 1;
        if (!c.completed) r.completed = false;
        const int sc = -c.score;
        if (sc >= beta) { r.score = beta; return true; }
        if (sc > alpha) alpha = sc;
        return false;
    };
    Move mv[kMaxMoves];
    if (in_check) {
        const int n = gen_pseudo(b, mv);
        bool any = false;
        for (int i = 0; i < n; ++i) {
            make_move_hd(b, mv[i], &nb);
            if (in_check_hd(nb, us)) continue;
            any = true;
            if (child(w_check_used, b_check_used)) return r;
        }
        r.score = any ? alpha : -1000000;
        return r;
    }
    int best_to = -1, skip_from = -1;
    uint16_t tried[kMaxMoves];   // the capture list's entries up to the played one, in order
    int n_tried = 0;
    for (uint32_t above = 0;;) {
        uint32_t best = kPeNoCapture;
        Move bm = 0;
        for (int pt = 0; pt < 6; ++pt)
            for (uint64_t bb = b.pieces[us * 6 + pt]; bb; bb &= bb - 1) {
                Move m;
                const uint32_t k = pe_piece_best(b, lsb(bb), pt, above, &m);
                if (k < best) { best = k; bm = m; }
            }
        if (best == kPeNoCapture) break;
        tried[n_tried++] = bm;
        make_move_hd(b, bm, &nb);
        if (!in_check_hd(nb, us)) { best_to = mv_to(bm); break; }
        above = best;
    }
    if (best_to >= 0) {
        for (int i = 0; i < n_tried && skip_from < 0; ++i)
            if (mv_to(tried[i]) == best_to) skip_from = mv_from(tried[i]);
        if (child(w_check_used, b_check_used)) return r;
    }
    if (!(us == 0 ? w_check_used : b_check_used)) {
        const int n = gen_pseudo(b, mv);
        for (int i = 0; i < n; ++i) {
            if (best_to >= 0 && mv_from(mv[i]) == skip_from && mv_to(mv[i]) == best_to) continue;
            make_move_hd(b, mv[i], &nb);
            if (!in_check_hd(nb, 1 - us) || in_check_hd(nb, us)) continue;
            if (child(us == 0 ? true : w_check_used, us == 1 ? true : b_check_used)) return r;
            break;
        }
    }
    r.score = alpha;
    return r;
}
inline Cap1Result forced_cap1(const cb_board& b) { return cap1_qsearch(b, -kPeWindow, kPeWindow, kPeMaxDepth, false, false); }
#endif

#if defined(__CUDACC__)
// k-th (0-based) set bit of a 64-bit set, k < popcount: binary search on popcounts
__device__ __forceinline__ int nth_set_bit(uint64_t bb, int k) {
    uint32_t w = (uint32_t)bb;
    int base = 0, c = __popc(w);
    if (k >= c) { w = (uint32_t)(bb >> 32); k -= c; base = 32; }
    c = __popc(w & 0xFFFFu); if (k >= c) { w >>= 16; k -= c; base += 16; }
    c = __popc(w & 0xFFu);   if (k >= c) { w >>= 8;  k -= c; base += 8; }
    c = __popc(w & 0xFu);    if (k >= c) { w >>= 4;  k -= c; base += 4; }
    c = __popc(w & 0x3u);    if (k >= c) { w >>= 2;  k -= c; base += 2; }
    return base + (k >= (int)(w & 1u) ? 1 : 0);
}

// Warp form of pesto_cp: two squares per lane.  Also fills `mailbox` (64 bytes of shared memory: 0 empty, else
// 1 + colour * 6 + piece) for the capture scan that follows; ends with a __syncwarp.
__device__ __forceinline__ int pesto_cp_warp(const cb_board& b, uint8_t* mailbox, int lane) {
    int code0 = 0, code1 = 0;
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        const uint64_t bb = b.pieces[i];
        if (((uint32_t)bb >> lane) & 1u) code0 = i + 1;
        if (((uint32_t)(bb >> 32) >> lane) & 1u) code1 = i + 1;
    }
    mailbox[lane] = (uint8_t)code0;
    mailbox[lane + 32] = (uint8_t)code1;
    int mg = 0, eg = 0, ph = 0;
    if (code0) {
        const int c = code0 > 6, pt = code0 - 1 - 6 * c, t = c == 0 ? (lane ^ 56) : lane, sgn = c == 0 ? 1 : -1;
        mg += sgn * (CB_MG_VALUE[pt] + CB_MG_PST[pt][t]);
        eg += sgn * (CB_EG_VALUE[pt] + CB_EG_PST[pt][t]);
        ph += CB_PHASE_INC[pt];
    }
    if (code1) {
        const int sq = lane + 32;
        const int c = code1 > 6, pt = code1 - 1 - 6 * c, t = c == 0 ? (sq ^ 56) : sq, sgn = c == 0 ? 1 : -1;
        mg += sgn * (CB_MG_VALUE[pt] + CB_MG_PST[pt][t]);
        eg += sgn * (CB_EG_VALUE[pt] + CB_EG_PST[pt][t]);
        ph += CB_PHASE_INC[pt];
    }
    mg = __reduce_add_sync(0xffffffffu, mg);
    eg = __reduce_add_sync(0xffffffffu, eg);
    ph = __reduce_add_sync(0xffffffffu, ph);
    const int mgp = ph < 24 ? ph : 24;
    const int score = (mg * mgp + eg * (24 - mgp)) / 24;
    return b.w_to_move ? score : -score;
}

// pe_piece_best with the victims read from the mailbox.
__device__ __forceinline__ uint32_t pe_piece_best_mb(const cb_board& b, const uint8_t* mailbox, int from, int pt, uint32_t above, Move* mv) {
    const Tables& T = tables();
    const int us = b.w_to_move ? 0 : 1, them = 1 - us;
    const uint64_t opp = b.occ[them], occ = b.occ[0] | b.occ[1];
    uint32_t best = kPeNoCapture;
    Move bm = 0;
    auto victim_of = [&](int to) { return (int)mailbox[to] - 1 - 6 * them; };   // -1 - 6*them .. on an empty square
    auto consider = [&](uint32_t k, Move m) {
        if (k > above && k < best) { best = k; bm = m; }
    };
    if (pt == PAWN) {
        const int r = from >> 3;
        const bool promo = us == 0 ? r == 6 : r == 1;
        for (uint64_t tg = T.pawn[us][from] & (opp | (b.en_passant != 0xFF ? 1ull << b.en_passant : 0ull)); tg; tg &= tg - 1) {
            const int to = lsb(tg);
            const int which = (to & 7) < (from & 7) ? 0 : 1;
            const int score = mailbox[to] ? 10 * victim_of(to) : 0;
            if (promo) {
                const int pp[4] = {QUEEN, ROOK, KNIGHT, BISHOP};
#pragma unroll
                for (int i = 0; i < 4; ++i) consider(pe_key(score, 6, from, which * 4 + i), make_mv(from, to, pp[i]));
            } else {
                consider(pe_key(score, 0, from, which), make_mv(from, to, 0));
            }
        }
        if (promo) {
            const int one = us == 0 ? from + 8 : from - 8;
            if (!(occ >> one & 1)) {
                const int pp[4] = {QUEEN, ROOK, KNIGHT, BISHOP};
#pragma unroll
                for (int i = 0; i < 4; ++i) consider(pe_key(0, 6, from, 8 + i), make_mv(from, one, pp[i]));
            }
        }
    } else if (pt == KNIGHT || pt == KING) {
        // position of the square delta in the reference's table order, 4 bits per entry, indexed by a delta code
        for (uint64_t tg = (pt == KNIGHT ? T.knight[from] : T.king[from]) & opp; tg; tg &= tg - 1) {
            const int to = lsb(tg), d = to - from;
            int idx;
            if (pt == KNIGHT) idx = d == 15 ? 0 : d == 17 ? 1 : d == 6 ? 2 : d == 10 ? 3 : d == -10 ? 4 : d == -6 ? 5 : d == -17 ? 6 : 7;
            else idx = d == -1 ? 0 : d == 7 ? 1 : d == 8 ? 2 : d == 9 ? 3 : d == 1 ? 4 : d == -7 ? 5 : d == -8 ? 6 : 7;
            consider(pe_key(10 * victim_of(to) - pt, pt == KNIGHT ? 1 : 5, from, idx), make_mv(from, to, 0));
        }
    } else {
        const uint64_t ra = pt != BISHOP ? rook_att(from, occ) & opp : 0ull;
        const uint64_t ba = pt != ROOK ? bishop_att(from, occ) & opp : 0ull;
        const int cls = pt == BISHOP ? 2 : pt == ROOK ? 3 : 4;
        for (uint64_t tg = ra; tg; tg &= tg - 1) {
            const int to = lsb(tg);
            consider(pe_key(10 * victim_of(to) - pt, cls, from, to), make_mv(from, to, 0));
        }
        for (uint64_t tg = ba; tg; tg &= tg - 1) {
            const int to = lsb(tg);
            consider(pe_key(10 * victim_of(to) - pt, cls, from, (pt == QUEEN ? 64 : 0) + to), make_mv(from, to, 0));
        }
    }
    *mv = bm;
    return best;
}

// make_move_hd for the moves of the exchange line (captures, en passant, promotions — never castling or a double
// push), applied in place to a copy of the position.  `pt` / `victim` are the mover's and the captured piece's
// types (victim < 0: empty target square).  One lane.
__device__ __forceinline__ void pe_apply_capture(cb_board& b, Move m, int pt, int victim) {
    const int from = mv_from(m), to = mv_to(m), promo = mv_promo(m);
    const int us = b.w_to_move ? 0 : 1, them = 1 - us;
    const uint64_t fbit = 1ull << from, tbit = 1ull << to;
    if (victim >= 0) {
        b.pieces[them * 6 + victim] &= ~tbit;
        b.occ[them] &= ~tbit;
    } else if (pt == PAWN && ((from ^ to) & 7)) {   // en passant
        const uint64_t ep = 1ull << (us == 0 ? to - 8 : to + 8);
        b.pieces[them * 6 + PAWN] &= ~ep;
        b.occ[them] &= ~ep;
    }
    b.pieces[us * 6 + pt] &= ~fbit;
    b.pieces[us * 6 + (promo ? promo : pt)] |= tbit;
    b.occ[us] = (b.occ[us] & ~fbit) | tbit;
    uint8_t cr = b.castling;
    if (pt == KING) cr &= us == 0 ? (uint8_t)~(WK | WQ) : (uint8_t)~(BK | BQ);
    if (from == 0 || to == 0) cr &= (uint8_t)~WQ;
    if (from == 7 || to == 7) cr &= (uint8_t)~WK;
    if (from == 56 || to == 56) cr &= (uint8_t)~BQ;
    if (from == 63 || to == 63) cr &= (uint8_t)~BK;
    b.castling = cr;
    b.en_passant = 0xFF;
    b.halfmove = 0;
    b.w_to_move = (uint8_t)!b.w_to_move;
}

// Is the king of colour `color` attacked in `b`?  Eight rays, knights, pawns and the other king on eleven lanes.
__device__ __forceinline__ bool in_check_warp(const cb_board& b, int color, int lane) {
    const Tables& T = tables();
    const uint64_t k = b.pieces[color * 6 + KING];
    const uint64_t* pc = &b.pieces[(1 - color) * 6];
    bool hit = false;
    if (k) {
        const int ks = lsb(k);
        if (lane < 8) hit = (ray_attack(lane, ks, b.occ[0] | b.occ[1]) & ((lane & 1 ? pc[BISHOP] : pc[ROOK]) | pc[QUEEN])) != 0;
        else if (lane == 8) hit = (T.knight[ks] & pc[KNIGHT]) != 0;
        else if (lane == 9) hit = (T.pawn[color][ks] & pc[PAWN]) != 0;
        else if (lane == 10) hit = (T.king[ks] & pc[KING]) != 0;
    }
    return __any_sync(0xffffffffu, hit);
}

// Warp form of principal_exchange_cp.  `b0` (shared memory) holds the position and is overwritten; `b1` is a
// second shared scratch board, `mailbox` 64 bytes of shared scratch.  One own piece per lane (at most 16) scans
// its captures; the lane holding the smallest key applies the move to a copy; the legality test runs on eleven
// lanes.  All lanes must call; all lanes return the score.
__device__ __forceinline__ int principal_exchange_cp_warp(cb_board& b0, cb_board& b1, uint8_t* mailbox, int lane) {
    cb_board* cur = &b0;
    cb_board* nxt = &b1;
    int ap[kPeMaxDepth + 1];
    int alpha = -kPeWindow, beta = kPeWindow, depth = 0, ret;
    for (;;) {
        __syncwarp();
        const cb_board& b = *cur;
        const int sp = pesto_cp_warp(b, mailbox, lane);
        if (sp >= beta) { ret = beta; break; }
        if (sp > alpha) alpha = sp;
        if (depth == kPeMaxDepth) { ret = alpha; break; }
        __syncwarp();
        const int us = b.w_to_move ? 0 : 1;
        const uint64_t own = b.occ[us];
        const bool has_piece = lane < __popcll(own);
        int from = 0, pt = 0;
        if (has_piece) {
            from = nth_set_bit(own, lane);
            pt = (int)mailbox[from] - 1 - 6 * us;
        }
        Move m = 0;
        uint32_t key = has_piece ? pe_piece_best_mb(b, mailbox, from, pt, 0, &m) : kPeNoCapture;
        bool found = false;
        for (;;) {
            const uint32_t kmin = __reduce_min_sync(0xffffffffu, key);
            if (kmin == kPeNoCapture) break;
            if (lane < 16) reinterpret_cast<uint64_t*>(nxt)[lane] = reinterpret_cast<const uint64_t*>(cur)[lane];
            __syncwarp();
            const bool winner = key == kmin;
            if (winner) pe_apply_capture(*nxt, m, pt, mailbox[mv_to(m)] ? (int)mailbox[mv_to(m)] - 1 - 6 * (1 - us) : -1);
            __syncwarp();
            if (!in_check_warp(*nxt, us, lane)) { found = true; break; }   // Board::is_legal, src/board.rs:352
            if (winner) key = pe_piece_best_mb(b, mailbox, from, pt, kmin, &m);
        }
        if (!found) { ret = alpha; break; }
        ap[depth++] = alpha;
        const int na = -beta;
        beta = -alpha;
        alpha = na;
        cb_board* t = cur; cur = nxt; nxt = t;
    }
    return pe_unwind(ap, depth, ret);
}
#endif

}  // namespace cbchess
