// chess_rules.hpp — the bitboard chess rules of chess.hpp as inline host+device code, so that the
// host self-play driver (selfplay.cpp) and the device-side tree kernels (mcts_device.cuh) run the SAME
// move generator: identical legal sets in identical order (captures / promotions first), identical
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
CB_HD uint64_t rook_att(int sq, uint64_t occ) {
    return ray_attack(0, sq, occ) | ray_attack(2, sq, occ) | ray_attack(4, sq, occ) | ray_attack(6, sq, occ);
}
CB_HD uint64_t bishop_att(int sq, uint64_t occ) {
    return ray_attack(1, sq, occ) | ray_attack(3, sq, occ) | ray_attack(5, sq, occ) | ray_attack(7, sq, occ);
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
    for (int p = 0; p < 6; ++p)
        if (b.pieces[color * 6 + p] & bit) return p;
    return -1;
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

// Pseudo-legal moves, captures and promotions first (src/move_generation.rs:323-328).
CB_HD int gen_pseudo(const cb_board& b, Move* out) {
    const Tables& T = tables();
    int n = 0;
    const int us = b.w_to_move ? 0 : 1, them = 1 - us;
    const uint64_t own = b.occ[us], opp = b.occ[them], occ = own | opp;
    const uint64_t* pc = &b.pieces[us * 6];
    const int fwd = us == 0 ? 8 : -8, promo_rank = us == 0 ? 7 : 0, start_rank = us == 0 ? 1 : 6;
    for (int pass = 0; pass < 2; ++pass) {
        const bool caps = pass == 0;
        for (uint64_t bb = pc[PAWN]; bb; bb &= bb - 1) {
            const int from = lsb(bb), one = from + fwd;
            if (caps) {
                for (uint64_t t = T.pawn[us][from] & opp; t; t &= t - 1) n = add_pawn(out, n, from, lsb(t), promo_rank);
                if (b.en_passant != 0xFF && (T.pawn[us][from] >> b.en_passant & 1)) out[n++] = make_mv(from, b.en_passant, 0);
                if ((one >> 3) == promo_rank && !(occ >> one & 1)) n = add_pawn(out, n, from, one, promo_rank);
            } else if ((one >> 3) != promo_rank && !(occ >> one & 1)) {
                out[n++] = make_mv(from, one, 0);
                if ((from >> 3) == start_rank && !(occ >> (one + fwd) & 1)) out[n++] = make_mv(from, one + fwd, 0);
            }
        }
        const uint64_t mask = caps ? opp : ~occ;
        for (uint64_t bb = pc[KNIGHT]; bb; bb &= bb - 1) {
            const int from = lsb(bb);
            for (uint64_t t = T.knight[from] & mask; t; t &= t - 1) out[n++] = make_mv(from, lsb(t), 0);
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
            for (uint64_t t = (rook_att(from, occ) | bishop_att(from, occ)) & mask; t; t &= t - 1)
                out[n++] = make_mv(from, lsb(t), 0);
        }
        for (uint64_t bb = pc[KING]; bb; bb &= bb - 1) {
            const int from = lsb(bb);
            for (uint64_t t = T.king[from] & mask; t; t &= t - 1) out[n++] = make_mv(from, lsb(t), 0);
        }
        if (!caps) {
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
    const int captured = piece_on(b, them, tbit);
    const bool reset_clock = pt == PAWN || captured >= 0;
    if (captured >= 0) b.pieces[them * 6 + captured] &= ~tbit;
    if (pt == PAWN && b.en_passant != 0xFF && to == b.en_passant && captured < 0 && ((from ^ to) & 7))
        b.pieces[them * 6 + PAWN] &= ~(1ull << (us == 0 ? to - 8 : to + 8));
    b.pieces[us * 6 + pt] &= ~fbit;
    b.pieces[us * 6 + (promo ? promo : pt)] |= tbit;
    if (pt == KING && (to - from == 2 || from - to == 2)) {
        const int rf = to > from ? from + 3 : from - 4, rt = to > from ? from + 1 : from - 1;
        b.pieces[us * 6 + ROOK] &= ~(1ull << rf);
        b.pieces[us * 6 + ROOK] |= 1ull << rt;
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
    for (int c = 0; c < 2; ++c) {
        uint64_t o = 0;
        for (int p = 0; p < 6; ++p) o |= b.pieces[c * 6 + p];
        b.occ[c] = o;
    }
    *out = b;
}

// Legal moves in generation order: a pseudo-legal move stays if the mover's king is not attacked after
// it (src/make_move.rs:24, src/board.rs:352).
CB_HD int gen_legal_hd(const cb_board& b, Move* out) {
    Move tmp[kMaxMoves];
    const int np = gen_pseudo(b, tmp);
    const int us = b.w_to_move ? 0 : 1;
    int n = 0;
    for (int i = 0; i < np; ++i) {
        cb_board nb;
        make_move_hd(b, tmp[i], &nb);
        if (!in_check_hd(nb, us)) out[n++] = tmp[i];
    }
    return n;
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
