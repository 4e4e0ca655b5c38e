// chess_rules.hpp — the bitboard chess rules of chess.hpp as inline host+device code, so that the
// host self-play driver (selfplay.cpp) and the device-side tree kernels (mcts_device.cuh) run the SAME
// move generator: identical legal sets in identical order (the reference's generation order), identical
// make-move side effects (src/make_move.rs:99-185), identical policy indices (src/tensor.rs:82-178).
// Attack tables (6 KB) are a static object on the host and a global-memory object on the device (L1-resident;
// the warp forms index them with a different square in every lane, which the constant cache would serialise).
#pragma once
#include <stdint.h>
#include <string.h>

#include "chess.hpp"

#if defined(__CUDACC__)
#define CB_HD __host__ __device__ __forceinline__
#else
#define CB_HD inline
#endif

namespace cbchess {

struct Tables {
    uint64_t knight[64], king[64], pawn[2][64];
    uint64_t ray[8][64];  // N NE E SE S SW W NW
    uint64_t line[4][64]; // file, rank, diagonal (NE-SW), antidiagonal (NW-SE) through the square, without the square
};

inline void build_tables(Tables* t) {
    static const int ndx[8] = {1, 2, 2, 1, -1, -2, -2, -1}, ndy[8] = {2, 1, -1, -2, -2, -1, 1, 2};
    static const int rdx[8] = {0, 1, 1, 1, 0, -1, -1, -1}, rdy[8] = {1, 1, 0, -1, -1, -1, 0, 1};
    for (int sq = 0; sq < 64; ++sq) {
        const int r = sq >> 3, f = sq & 7;
        t->knight[sq] = t->king[sq] = t->pawn[0][sq] = t->pawn[1][sq] = 0;
        for (int i = 0; i < 8; ++i) {
            int rr = r + ndy[i], ff = f + ndx[i];
            if (rr >= 0 && rr < 8 && ff >= 0 && ff < 8) t->knight[sq] |= 1ull << (rr * 8 + ff);
            rr = r + rdy[i]; ff = f + rdx[i];
            if (rr >= 0 && rr < 8 && ff >= 0 && ff < 8) t->king[sq] |= 1ull << (rr * 8 + ff);
            uint64_t line = 0;
            for (rr = r + rdy[i], ff = f + rdx[i]; rr >= 0 && rr < 8 && ff >= 0 && ff < 8; rr += rdy[i], ff += rdx[i])
                line |= 1ull << (rr * 8 + ff);
            t->ray[i][sq] = line;
        }
        t->line[0][sq] = t->ray[0][sq] | t->ray[4][sq];
        t->line[1][sq] = t->ray[2][sq] | t->ray[6][sq];
        t->line[2][sq] = t->ray[1][sq] | t->ray[5][sq];
        t->line[3][sq] = t->ray[7][sq] | t->ray[3][sq];
        for (int c = 0; c < 2; ++c) {
            const int rr = r + (c == 0 ? 1 : -1);
            if (rr >= 0 && rr < 8) {
                if (f > 0) t->pawn[c][sq] |= 1ull << (rr * 8 + f - 1);
                if (f < 7) t->pawn[c][sq] |= 1ull << (rr * 8 + f + 1);
            }
        }
    }
}

const Tables& host_tables();   // chess.cpp
#if defined(__CUDACC__)
__device__ Tables d_chess_tables;   // one CUDA translation unit (cb_api.cu) includes this; uploaded per device there
#endif

CB_HD const Tables& tables() {
#if defined(__CUDA_ARCH__)
    return d_chess_tables;
#else
    return host_tables();
#endif
}

CB_HD int lsb(uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __ffsll(static_cast<long long>(b)) - 1;
#else
    return __builtin_ctzll(b);
#endif
}
CB_HD int msb(uint64_t b) {
#if defined(__CUDA_ARCH__)
    return 63 - __clzll(static_cast<long long>(b));
#else
    return 63 - __builtin_clzll(b);
#endif
}

// classical ray attacks: directions 7,0,1,2 grow towards higher squares, the others lower
CB_HD uint64_t ray_attack(int dir, int sq, uint64_t occ) {
    const Tables& T = tables();
    uint64_t att = T.ray[dir][sq];
    const uint64_t blockers = att & occ;
    if (blockers) {
        const bool up = (dir == 0 || dir == 1 || dir == 2 || dir == 7);
        const int first = up ? lsb(blockers) : msb(blockers);
        att ^= T.ray[dir][first];
    }
    return att;
}
// Sliding attacks along one line by subtraction ("hyperbola quintessence"): with o the occupied squares of the line and r the
// slider, (o - 2r) flips the squares from r up to the first blocker above it; the same on the bit-reversed board gives the
// squares below it.  A dozen instructions per line instead of two table-and-scan rays.
CB_HD uint64_t bit_reverse64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __brevll(x);
#else
    x = __builtin_bswap64(x);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((x & 0x0F0F0F0F0F0F0F0Full) << 4);
    x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
    x = ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);
    return x;
#endif
}
CB_HD uint64_t line_attack(uint64_t mask, int sq, uint64_t occ) {
    const uint64_t o = occ & mask;
    const uint64_t up = o - (2ull << sq);
    const uint64_t down = bit_reverse64(bit_reverse64(o) - (2ull << (63 - sq)));
    return (up ^ down) & mask;
}
CB_HD uint64_t rook_att(int sq, uint64_t occ) {
    const Tables& T = tables();
    return line_attack(T.line[0][sq], sq, occ) | line_attack(T.line[1][sq], sq, occ);
}
CB_HD uint64_t bishop_att(int sq, uint64_t occ) {
    const Tables& T = tables();
    return line_attack(T.line[2][sq], sq, occ) | line_attack(T.line[3][sq], sq, occ);
}

CB_HD bool attacked(const cb_board& b, int sq, int by) {
    const Tables& T = tables();
    const uint64_t* pc = &b.pieces[by * 6];
    const uint64_t occ = b.occ[0] | b.occ[1];
    if (T.pawn[1 - by][sq] & pc[PAWN]) return true;
    if (T.knight[sq] & pc[KNIGHT]) return true;
    if (T.king[sq] & pc[KING]) return true;
    if (rook_att(sq, occ) & (pc[ROOK] | pc[QUEEN])) return true;
    if (bishop_att(sq, occ) & (pc[BISHOP] | pc[QUEEN])) return true;
    return false;
}

CB_HD int piece_on(const cb_board& b, int color, uint64_t bit) {
    const uint64_t* pc = &b.pieces[color * 6];
    int p = -1;   // (a square holds one piece: the order of the tests is free; no early exit, no divergence)
    p = (pc[5] & bit) ? 5 : p;
    p = (pc[4] & bit) ? 4 : p;
    p = (pc[3] & bit) ? 3 : p;
    p = (pc[2] & bit) ? 2 : p;
    p = (pc[1] & bit) ? 1 : p;
    p = (pc[0] & bit) ? 0 : p;
    return p;
}

CB_HD int add_pawn(Move* out, int n, int from, int to, int promo_rank) {
    if ((to >> 3) == promo_rank) {
        out[n++] = make_mv(from, to, QUEEN);
        out[n++] = make_mv(from, to, ROOK);
        out[n++] = make_mv(from, to, KNIGHT);
        out[n++] = make_mv(from, to, BISHOP);
    } else {
        out[n++] = make_mv(from, to, 0);
    }
    return n;
}

// Target order of one knight / king, as square deltas: the reference's tables list them this way
// (init_knight_moves src/magic_bitboard.rs:160-189, init_king_moves :120-149), and a piece's moves are generated in
// table order.  A delta is valid when the target is in the piece's attack set (a wrapped delta never is).
CB_HD int knight_delta(int i) { return i == 0 ? 15 : i == 1 ? 17 : i == 2 ? 6 : i == 3 ? 10 : i == 4 ? -10 : i == 5 ? -6 : i == 6 ? -17 : -15; }
CB_HD int king_delta(int i) { return i == 0 ? -1 : i == 1 ? 7 : i == 2 ? 8 : i == 3 ? 9 : i == 4 ? 1 : i == 5 ? -7 : i == 6 ? -8 : -9; }
CB_HD uint64_t step_targets_bit(uint64_t att, int from, int d) {
    const int to = from + d;
    return (to >= 0 && to < 64) ? (att & (1ull << to)) : 0ull;
}

// Pseudo-legal moves in the reference's generation order (gen_pseudo_legal_moves, src/move_generation.rs:287-328):
//   capture list  pawn captures incl. en passant (per pawn from the lowest square, the a-file side target first), knights
//                 (table order), bishops, rooks (attack-set bits ascending), queens (rook set, then bishop set), king
//                 (table order), then every promotion (per pawn: capturing ones first; Q, R, N, B)
//   quiet list    pawns (the DOUBLE push before the single one, src/magic_bitboard.rs:248-265), knights, bishops, rooks,
//                 queens, then the king: castling (king side, queen side) before its steps
// Children of a tree node are created in this order (src/mcts/tactical_mcts.rs:813-836), which decides PUCT ties, the
// root's force-visit order and the order of the tactical phase.  *n_caps (nullable) = length of the capture list.
CB_HD int gen_pseudo(const cb_board& b, Move* out, int* n_caps = nullptr) {
    const Tables& T = tables();
    int n = 0;
    const int us = b.w_to_move ? 0 : 1, them = 1 - us;
    const uint64_t own = b.occ[us], opp = b.occ[them], occ = own | opp;
    const uint64_t* pc = &b.pieces[us * 6];
    const int fwd = us == 0 ? 8 : -8, promo_from_rank = us == 0 ? 6 : 1, start_rank = us == 0 ? 1 : 6;
    const uint64_t ep = b.en_passant != 0xFF ? 1ull << b.en_passant : 0ull;
    for (int pass = 0; pass < 3; ++pass) {   // 0 captures, 1 promotions, 2 quiets
        for (uint64_t bb = pc[PAWN]; bb; bb &= bb - 1) {
            const int from = lsb(bb), one = from + fwd;
            const bool promo = (from >> 3) == promo_from_rank;
            if (pass == 0 && !promo) {
                for (uint64_t t = T.pawn[us][from] & (opp | ep); t; t &= t - 1) out[n++] = make_mv(from, lsb(t), 0);
            } else if (pass == 1 && promo) {
                for (uint64_t t = T.pawn[us][from] & (opp | ep); t; t &= t - 1) n = add_pawn(out, n, from, lsb(t), one >> 3);
                if (!(occ >> one & 1)) n = add_pawn(out, n, from, one, one >> 3);
            } else if (pass == 2 && !promo && !(occ >> one & 1)) {
                if ((from >> 3) == start_rank && !(occ >> (one + fwd) & 1)) out[n++] = make_mv(from, one + fwd, 0);
                out[n++] = make_mv(from, one, 0);
            }
        }
        if (pass == 1) {
            if (n_caps) *n_caps = n;
            continue;
        }
        const uint64_t mask = pass == 0 ? opp : ~occ;
        for (uint64_t bb = pc[KNIGHT]; bb; bb &= bb - 1) {
            const int from = lsb(bb);
            const uint64_t att = T.knight[from] & mask;
            for (int i = 0; i < 8; ++i)
                if (step_targets_bit(att, from, knight_delta(i))) out[n++] = make_mv(from, from + knight_delta(i), 0);
        }
        for (uint64_t bb = pc[BISHOP]; bb; bb &= bb - 1) {
            const int from = lsb(bb);
            for (uint64_t t = bishop_att(from, occ) & mask; t; t &= t - 1) out[n++] = make_mv(from, lsb(t), 0);
        }
        for (uint64_t bb = pc[ROOK]; bb; bb &= bb - 1) {
            const int from = lsb(bb);
            for (uint64_t t = rook_att(from, occ) & mask; t; t &= t - 1) out[n++] = make_mv(from, lsb(t), 0);
        }
        for (uint64_t bb = pc[QUEEN]; bb; bb &= bb - 1) {
            const int from = lsb(bb);
            for (uint64_t t = rook_att(from, occ) & mask; t; t &= t - 1) out[n++] = make_mv(from, lsb(t), 0);
            for (uint64_t t = bishop_att(from, occ) & mask; t; t &= t - 1) out[n++] = make_mv(from, lsb(t), 0);
        }
        if (pass == 2) {
            const int ksq = us == 0 ? 4 : 60;
            if ((pc[KING] >> ksq & 1) && !attacked(b, ksq, them)) {
                const int ks = us == 0 ? WK : BK, qs = us == 0 ? WQ : BQ;
                if ((b.castling & ks) && (pc[ROOK] >> (ksq + 3) & 1) && !(occ & (3ull << (ksq + 1))) &&
                    !attacked(b, ksq + 1, them) && !attacked(b, ksq + 2, them))
                    out[n++] = make_mv(ksq, ksq + 2, 0);
                if ((b.castling & qs) && (pc[ROOK] >> (ksq - 4) & 1) && !(occ & (7ull << (ksq - 3))) &&
                    !attacked(b, ksq - 1, them) && !attacked(b, ksq - 2, them))
                    out[n++] = make_mv(ksq, ksq - 2, 0);
            }
        }
        for (uint64_t bb = pc[KING]; bb; bb &= bb - 1) {
            const int from = lsb(bb);
            const uint64_t att = T.king[from] & mask;
            for (int i = 0; i < 8; ++i)
                if (step_targets_bit(att, from, king_delta(i))) out[n++] = make_mv(from, from + king_delta(i), 0);
        }
    }
    return n;
}

CB_HD bool in_check_hd(const cb_board& b, int color) {
    const uint64_t k = b.pieces[color * 6 + KING];
    return k && attacked(b, lsb(k), 1 - color);
}

CB_HD void make_move_hd(const cb_board& src, Move m, cb_board* out) {
    cb_board b = src;
    const int from = mv_from(m), to = mv_to(m), promo = mv_promo(m);
    const int us = b.w_to_move ? 0 : 1, them = 1 - us;
    const uint64_t fbit = 1ull << from, tbit = 1ull << to;
    const int pt = piece_on(b, us, fbit);
    const int captured = (b.occ[them] & tbit) ? piece_on(b, them, tbit) : -1;
    const bool reset_clock = pt == PAWN || captured >= 0;
    uint64_t occ_us = (b.occ[us] & ~fbit) | tbit, occ_them = b.occ[them] & ~tbit;   // kept up to date move by move
    if (captured >= 0) b.pieces[them * 6 + captured] &= ~tbit;
    if (pt == PAWN && b.en_passant != 0xFF && to == b.en_passant && captured < 0 && ((from ^ to) & 7)) {
        const uint64_t epb = 1ull << (us == 0 ? to - 8 : to + 8);
        b.pieces[them * 6 + PAWN] &= ~epb;
        occ_them &= ~epb;
    }
    b.pieces[us * 6 + pt] &= ~fbit;
    b.pieces[us * 6 + (promo ? promo : pt)] |= tbit;
    if (pt == KING && (to - from == 2 || from - to == 2)) {
        const int rf = to > from ? from + 3 : from - 4, rt = to > from ? from + 1 : from - 1;
        b.pieces[us * 6 + ROOK] &= ~(1ull << rf);
        b.pieces[us * 6 + ROOK] |= 1ull << rt;
        occ_us = (occ_us & ~(1ull << rf)) | (1ull << rt);
    }
    if (pt == KING) b.castling &= us == 0 ? (uint8_t)~(WK | WQ) : (uint8_t)~(BK | BQ);
    if (from == 0 || to == 0) b.castling &= (uint8_t)~WQ;
    if (from == 7 || to == 7) b.castling &= (uint8_t)~WK;
    if (from == 56 || to == 56) b.castling &= (uint8_t)~BQ;
    if (from == 63 || to == 63) b.castling &= (uint8_t)~BK;
    b.en_passant = 0xFF;
    if (pt == PAWN && (to - from == 16 || from - to == 16)) b.en_passant = (uint8_t)((from + to) / 2);
    b.halfmove = reset_clock ? 0 : (uint8_t)(b.halfmove < 255 ? b.halfmove + 1 : 255);
    b.w_to_move = (uint8_t)!b.w_to_move;
    b.occ[us] = occ_us;
    b.occ[them] = occ_them;
    *out = b;
}

// ---------------------------------------------------------------------------------------------------------------
// Legality and "gives check" of a pseudo-legal move WITHOUT making it: the same answers as
// make_move_hd + in_check_hd (tests/test_host.py::test_move_flags_equal_make_move compares them on every pseudo-legal
// move of > 20,000 positions), at a fraction of the instructions — the tree kernels' gate searches test thousands of
// moves per leaf.  MoveCtx holds what the moves of one position share.
struct MoveCtx {
    uint64_t occ;          // all pieces
    uint64_t own_klines;   // every square on a rank, file or diagonal of the mover's king (empty board)
    uint64_t opp_klines;   // ... of the other king
    uint64_t rvis, bvis;   // squares a rook / bishop standing on the other king's square would see
    int us, ks, eks;       // side to move, its king square, the other king's square (-1: no king)
    bool in_check;         // the side to move is in check
};
CB_HD MoveCtx move_ctx(const cb_board& b) {
    MoveCtx c;
    c.us = b.w_to_move ? 0 : 1;
    c.occ = b.occ[0] | b.occ[1];
    const uint64_t k = b.pieces[c.us * 6 + KING], ek = b.pieces[(1 - c.us) * 6 + KING];
    c.ks = k ? lsb(k) : -1;
    c.eks = ek ? lsb(ek) : -1;
    c.own_klines = k ? (rook_att(c.ks, 0ull) | bishop_att(c.ks, 0ull)) : 0ull;
    c.opp_klines = ek ? (rook_att(c.eks, 0ull) | bishop_att(c.eks, 0ull)) : 0ull;
    c.rvis = ek ? rook_att(c.eks, c.occ) : 0ull;
    c.bvis = ek ? bishop_att(c.eks, c.occ) : 0ull;
    c.in_check = k && attacked(b, c.ks, 1 - c.us);
    return c;
}
// bit 0: the move is legal; bit 1 (only when legal and want_check): it gives check
CB_HD int move_flags_hd(const cb_board& b, const MoveCtx& c, Move m, bool want_check) {
    const Tables& T = tables();
    const int from = mv_from(m), to = mv_to(m), promo = mv_promo(m);
    const int us = c.us, them = 1 - us;
    const uint64_t fb = 1ull << from, tb = 1ull << to;
    const uint64_t* own = &b.pieces[us * 6];
    const uint64_t* opp = &b.pieces[them * 6];
    const int pt = piece_on(b, us, fb);
    const bool is_ep = pt == PAWN && b.en_passant != 0xFF && to == b.en_passant && ((from ^ to) & 7) && !(b.occ[them] & tb);
    const bool castle = pt == KING && (to - from == 2 || from - to == 2);
    uint64_t occ2 = (c.occ & ~fb) | tb;
    const uint64_t epb = is_ep ? 1ull << (us == 0 ? to - 8 : to + 8) : 0ull;
    occ2 &= ~epb;
    uint64_t rfb = 0, rtb = 0;
    if (castle) {
        rfb = 1ull << (to > from ? from + 3 : from - 4);
        rtb = 1ull << (to > from ? from + 1 : from - 1);
        occ2 = (occ2 & ~rfb) | rtb;
    }
    bool legal;
    if (c.ks < 0) legal = true;
    else if (pt != KING && !is_ep && !c.in_check && !(c.own_klines >> from & 1)) legal = true;   // cannot expose the king
    else {
        const int k2 = pt == KING ? to : c.ks;
        const uint64_t keep = ~tb;   // a captured piece is gone
        legal = !((T.pawn[us][k2] & opp[PAWN] & keep & ~epb) || (T.knight[k2] & opp[KNIGHT] & keep) || (T.king[k2] & opp[KING]) ||
                  (rook_att(k2, occ2) & (opp[ROOK] | opp[QUEEN]) & keep) || (bishop_att(k2, occ2) & (opp[BISHOP] | opp[QUEEN]) & keep));
    }
    if (!legal) return 0;
    if (!want_check || c.eks < 0) return 1;
    const uint64_t ek = 1ull << c.eks;
    bool chk;
    if (!promo && !is_ep && !castle && !(c.opp_klines >> from & 1)) {   // no line can open: only the mover itself can check
        chk = pt == PAWN ? (T.pawn[us][to] & ek) != 0 : pt == KNIGHT ? (T.knight[to] & ek) != 0 : pt == BISHOP ? (c.bvis >> to & 1) != 0
            : pt == ROOK ? (c.rvis >> to & 1) != 0 : pt == QUEEN ? ((c.rvis | c.bvis) >> to & 1) != 0 : false;
    } else {
        const int npt = promo ? promo : pt;
        uint64_t o[6];
        for (int q = 0; q < 6; ++q) o[q] = own[q];
        o[pt] &= ~fb;
        o[npt] |= tb;
        o[ROOK] = (o[ROOK] & ~rfb) | rtb;
        chk = (T.pawn[them][c.eks] & o[PAWN]) || (T.knight[c.eks] & o[KNIGHT]) || (rook_att(c.eks, occ2) & (o[ROOK] | o[QUEEN])) ||
              (bishop_att(c.eks, occ2) & (o[BISHOP] | o[QUEEN]));
    }
    return chk ? 3 : 1;
}

// Legal moves in generation order: a pseudo-legal move stays if the mover's king is not attacked after
// it (src/make_move.rs:24, src/board.rs:352).  *n_tactical (nullable) = how many of them come from the capture list:
// the candidates of the tactical phase (src/mcts/tactical.rs:155-177).
CB_HD int gen_legal_hd(const cb_board& b, Move* out, int* n_tactical = nullptr) {
    Move tmp[kMaxMoves];
    int ncap = 0;
    const int np = gen_pseudo(b, tmp, &ncap);
    const int us = b.w_to_move ? 0 : 1;
    int n = 0, nt = 0;
    for (int i = 0; i < np; ++i) {
        cb_board nb;
        make_move_hd(b, tmp[i], &nb);
        if (!in_check_hd(nb, us)) {
            out[n++] = tmp[i];
            if (i < ncap) ++nt;
        }
    }
    if (n_tactical) *n_tactical = nt;
    return n;
}

// calculate_mvv_lva of the tactical phase (src/mcts/tactical.rs:179-216) as an integer: 10 * victim - attacker + promoted
// piece with P1 N3 B3 R5 Q9 K0 (the en-passant target square is empty: victim 0).  Range [-9, 98].
CB_HD int tactical_score_hd(const cb_board& b, Move m) {
    const int us = b.w_to_move ? 0 : 1;
    const int val[6] = {1, 3, 3, 5, 9, 0};
    const int victim = piece_on(b, 1 - us, 1ull << mv_to(m)), attacker = piece_on(b, us, 1ull << mv_from(m));
    int s = (victim >= 0 ? 10 * val[victim] : 0) - (attacker >= 0 ? val[attacker] : 0);
    if (mv_promo(m)) s += val[mv_promo(m)];
    return s;
}

// Rank of every tactical child in the order the reference visits them (stable descending sort by score,
// src/mcts/tactical.rs:166-172): rank[i] for the first n_tactical moves of a legal list in generation order.
CB_HD void tactical_ranks_hd(const cb_board& b, const Move* legal, int n_tactical, uint8_t* rank) {
    int sc[kMaxMoves];
    for (int i = 0; i < n_tactical; ++i) sc[i] = tactical_score_hd(b, legal[i]);
    for (int i = 0; i < n_tactical; ++i) {
        int r = 0;
        for (int j = 0; j < n_tactical; ++j) r += sc[j] > sc[i] || (sc[j] == sc[i] && j < i);
        rank[i] = (uint8_t)r;
    }
}

CB_HD int policy_index_hd(Move m, bool white_to_move) {
    int from = mv_from(m), to = mv_to(m);
    const int promo = mv_promo(m);
    if (!white_to_move) { from ^= 56; to ^= 56; }
    const int dx = (to & 7) - (from & 7), dy = (to >> 3) - (from >> 3);
    const int adx = dx < 0 ? -dx : dx, ady = dy < 0 ? -dy : dy;
    int plane;
    if (promo != 0 && promo != QUEEN) {
        if (adx > 1) return -1;
        plane = 64 + (dx == 0 ? 0 : dx < 0 ? 1 : 2) + (promo == KNIGHT ? 0 : promo == BISHOP ? 3 : 6);
    } else if (adx * ady == 2) {
        const int kdx[8] = {1, 2, 2, 1, -1, -2, -2, -1}, kdy[8] = {2, 1, -1, -2, -2, -1, 1, 2};
        plane = -1;
        for (int i = 0; i < 8; ++i)
            if (kdx[i] == dx && kdy[i] == dy) plane = 56 + i;
    } else {
        if (!(dx == 0 || dy == 0 || adx == ady) || (dx == 0 && dy == 0)) return -1;
        const int sx = (dx > 0) - (dx < 0), sy = (dy > 0) - (dy < 0);
        // N NE E SE S SW W NW
        const int dir_of[3][3] = {{5, 6, 7}, {4, -1, 0}, {3, 2, 1}};  // [sx+1][sy+1]
        plane = dir_of[sx + 1][sy + 1] * 7 + ((adx > ady ? adx : ady) - 1);
    }
    return plane < 0 ? -1 : from * 73 + plane;
}

CB_HD uint64_t position_key_hd(const cb_board& b) {
    uint64_t h = 0x9E3779B97F4A7C15ull;
    for (int i = 0; i < 12; ++i) {
        h ^= b.pieces[i] + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
        h *= 0xFF51AFD7ED558CCDull;
    }
    h ^= (uint64_t)b.w_to_move | ((uint64_t)b.castling << 8) | ((uint64_t)b.en_passant << 16);
    h *= 0xC4CEB9FE1A85EC53ull;
    return h ^ (h >> 29);
}

}  // namespace cbchess
