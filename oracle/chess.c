/* chess.c — small bitboard legal move generator, make-move, perft and a seeded
 * random-playout position source.  TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * This is a fresh implementation of the rules of chess, not a transcription
 * of the reference's magic-bitboard generator; it is pinned to the reference
 * by the perft known answers of tests/perft_tests.rs:68-291 (replayed in
 * tests/test_oracle_chess.py).  Board-state side effects that reach the
 * network input follow src/make_move.rs:24-225: the en-passant square is set
 * after EVERY double pawn push (:99-109), castling rights drop on king moves
 * (:128-161), on a rook leaving its corner (:162-174) and on any capture that
 * lands on a corner (:177-185). */
#include "oracle.h"
#include <string.h>

enum { P = 0, N = 1, B = 2, R = 3, Q = 4, K = 5 };
enum { WK = 1, WQ = 2, BK = 4, BQ = 8 };

static uint64_t knight_att[64], king_att[64], pawn_att[2][64];
static int tables_ready = 0;

static void init_tables(void) {
    if (tables_ready) return;
    for (int sq = 0; sq < 64; ++sq) {
        const int r = sq >> 3, f = sq & 7;
        uint64_t kn = 0, kg = 0;
        static const int ndx[8] = {1, 2, 2, 1, -1, -2, -2, -1}, ndy[8] = {2, 1, -1, -2, -2, -1, 1, 2};
        for (int i = 0; i < 8; ++i) {
            int rr = r + ndy[i], ff = f + ndx[i];
            if (rr >= 0 && rr < 8 && ff >= 0 && ff < 8) kn |= 1ULL << (rr * 8 + ff);
        }
        for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx) {
                if (!dx && !dy) continue;
                int rr = r + dy, ff = f + dx;
                if (rr >= 0 && rr < 8 && ff >= 0 && ff < 8) kg |= 1ULL << (rr * 8 + ff);
            }
        knight_att[sq] = kn;
        king_att[sq] = kg;
        for (int c = 0; c < 2; ++c) {
            uint64_t pa = 0;
            int rr = r + (c == 0 ? 1 : -1);
            if (rr >= 0 && rr < 8) {
                if (f > 0) pa |= 1ULL << (rr * 8 + f - 1);
                if (f < 7) pa |= 1ULL << (rr * 8 + f + 1);
            }
            pawn_att[c][sq] = pa;
        }
    }
    tables_ready = 1;
}

static uint64_t slide(int sq, uint64_t occ, const int* dxs, const int* dys) {
    uint64_t att = 0;
    for (int d = 0; d < 4; ++d) {
        int r = (sq >> 3) + dys[d], f = (sq & 7) + dxs[d];
        while (r >= 0 && r < 8 && f >= 0 && f < 8) {
            const uint64_t bit = 1ULL << (r * 8 + f);
            att |= bit;
            if (occ & bit) break;
            r += dys[d];
            f += dxs[d];
        }
    }
    return att;
}
static uint64_t rook_att(int sq, uint64_t occ) {
    static const int dx[4] = {1, -1, 0, 0}, dy[4] = {0, 0, 1, -1};
    return slide(sq, occ, dx, dy);
}
static uint64_t bishop_att(int sq, uint64_t occ) {
    static const int dx[4] = {1, 1, -1, -1}, dy[4] = {1, -1, 1, -1};
    return slide(sq, occ, dx, dy);
}

/* Is `sq` attacked by any piece of colour `by`? */
static int attacked(const orc_board* b, int sq, int by) {
    const uint64_t* pc = &b->pieces[by * 6];
    const uint64_t occ = b->occ[0] | b->occ[1];
    if (pawn_att[1 - by][sq] & pc[P]) return 1;
    if (knight_att[sq] & pc[N]) return 1;
    if (king_att[sq] & pc[K]) return 1;
    if (rook_att(sq, occ) & (pc[R] | pc[Q])) return 1;
    if (bishop_att(sq, occ) & (pc[B] | pc[Q])) return 1;
    return 0;
}

int orc_in_check(const orc_board* b, int color) {
    init_tables();
    const uint64_t k = b->pieces[color * 6 + K];
    if (!k) return 0;
    return attacked(b, __builtin_ctzll(k), 1 - color);
}

static int piece_on(const orc_board* b, int color, int sq) {
    const uint64_t bit = 1ULL << sq;
    for (int p = 0; p < 6; ++p)
        if (b->pieces[color * 6 + p] & bit) return p;
    return -1;
}

void orc_make_move(const orc_board* src, orc_move m, orc_board* out) {
    init_tables();
    orc_board b = *src;
    const int from = ORC_FROM(m), to = ORC_TO(m), promo = ORC_PROMO(m);
    const int us = b.w_to_move ? 0 : 1, them = 1 - us;
    const uint64_t fbit = 1ULL << from, tbit = 1ULL << to;
    const int pt = piece_on(&b, us, from);
    int captured = piece_on(&b, them, to);
    int reset_clock = (pt == P) || captured >= 0;

    if (captured >= 0) b.pieces[them * 6 + captured] &= ~tbit;
    if (pt == P && b.en_passant != 0xFF && to == b.en_passant && captured < 0 && ((from ^ to) & 7)) {
        const int victim = us == 0 ? to - 8 : to + 8;
        b.pieces[them * 6 + P] &= ~(1ULL << victim);
    }
    b.pieces[us * 6 + pt] &= ~fbit;
    b.pieces[us * 6 + (promo ? promo : pt)] |= tbit;

    if (pt == K && (to - from == 2 || from - to == 2)) {
        const int rf = to > from ? from + 3 : from - 4;
        const int rt = to > from ? from + 1 : from - 1;
        b.pieces[us * 6 + R] &= ~(1ULL << rf);
        b.pieces[us * 6 + R] |= 1ULL << rt;
    }
    if (pt == K) b.castling &= us == 0 ? (uint8_t)~(WK | WQ) : (uint8_t)~(BK | BQ);
    /* a rook leaving a corner, or anything captured on a corner */
    if (from == 0 || to == 0) b.castling &= (uint8_t)~WQ;
    if (from == 7 || to == 7) b.castling &= (uint8_t)~WK;
    if (from == 56 || to == 56) b.castling &= (uint8_t)~BQ;
    if (from == 63 || to == 63) b.castling &= (uint8_t)~BK;

    b.en_passant = 0xFF;
    if (pt == P && (to - from == 16 || from - to == 16)) b.en_passant = (uint8_t)((from + to) / 2);
    b.halfmove = reset_clock ? 0 : (uint8_t)(b.halfmove < 255 ? b.halfmove + 1 : 255);
    b.w_to_move = (uint8_t)!b.w_to_move;
    for (int c = 0; c < 2; ++c) {
        uint64_t o = 0;
        for (int p = 0; p < 6; ++p) o |= b.pieces[c * 6 + p];
        b.occ[c] = o;
    }
    *out = b;
}

static int push_pawn_moves(orc_move* out, int n, int from, int to, int promo_rank) {
    if ((to >> 3) == promo_rank) {
        out[n++] = ORC_MV(from, to, Q);
        out[n++] = ORC_MV(from, to, R);
        out[n++] = ORC_MV(from, to, N);
        out[n++] = ORC_MV(from, to, B);
    } else {
        out[n++] = ORC_MV(from, to, 0);
    }
    return n;
}

/* Pseudo-legal generation in two passes (captures and promotions, then quiet
 * moves — the reference's ordering class, src/move_generation.rs:323-328),
 * followed by the make-and-test legality filter (src/board.rs:352). */
static int gen_pseudo(const orc_board* b, orc_move* out) {
    int n = 0;
    const int us = b->w_to_move ? 0 : 1, them = 1 - us;
    const uint64_t own = b->occ[us], opp = b->occ[them], occ = own | opp;
    const uint64_t* pc = &b->pieces[us * 6];
    const int fwd = us == 0 ? 8 : -8, promo_rank = us == 0 ? 7 : 0, start_rank = us == 0 ? 1 : 6;

    for (int pass = 0; pass < 2; ++pass) {
        const int caps = pass == 0;
        /* pawns */
        for (uint64_t bb = pc[P]; bb; bb &= bb - 1) {
            const int from = __builtin_ctzll(bb);
            const int one = from + fwd;
            if (caps) {
                uint64_t targets = pawn_att[us][from] & opp;
                for (; targets; targets &= targets - 1)
                    n = push_pawn_moves(out, n, from, __builtin_ctzll(targets), promo_rank);
                if (b->en_passant != 0xFF && (pawn_att[us][from] & (1ULL << b->en_passant)))
                    out[n++] = ORC_MV(from, b->en_passant, 0);
                if (one >= 0 && one < 64 && (one >> 3) == promo_rank && !(occ & (1ULL << one)))
                    n = push_pawn_moves(out, n, from, one, promo_rank);
            } else {
                if (one >= 0 && one < 64 && (one >> 3) != promo_rank && !(occ & (1ULL << one))) {
                    out[n++] = ORC_MV(from, one, 0);
                    const int two = one + fwd;
                    if ((from >> 3) == start_rank && !(occ & (1ULL << two))) out[n++] = ORC_MV(from, two, 0);
                }
            }
        }
        const uint64_t mask = caps ? opp : ~occ;
        for (uint64_t bb = pc[N]; bb; bb &= bb - 1) {
            const int from = __builtin_ctzll(bb);
            for (uint64_t t = knight_att[from] & mask; t; t &= t - 1) out[n++] = ORC_MV(from, __builtin_ctzll(t), 0);
        }
        for (uint64_t bb = pc[B]; bb; bb &= bb - 1) {
            const int from = __builtin_ctzll(bb);
            for (uint64_t t = bishop_att(from, occ) & mask; t; t &= t - 1) out[n++] = ORC_MV(from, __builtin_ctzll(t), 0);
        }
        for (uint64_t bb = pc[R]; bb; bb &= bb - 1) {
            const int from = __builtin_ctzll(bb);
            for (uint64_t t = rook_att(from, occ) & mask; t; t &= t - 1) out[n++] = ORC_MV(from, __builtin_ctzll(t), 0);
        }
        for (uint64_t bb = pc[Q]; bb; bb &= bb - 1) {
            const int from = __builtin_ctzll(bb);
            for (uint64_t t = (rook_att(from, occ) | bishop_att(from, occ)) & mask; t; t &= t - 1)
                out[n++] = ORC_MV(from, __builtin_ctzll(t), 0);
        }
        for (uint64_t bb = pc[K]; bb; bb &= bb - 1) {
            const int from = __builtin_ctzll(bb);
            for (uint64_t t = king_att[from] & mask; t; t &= t - 1) out[n++] = ORC_MV(from, __builtin_ctzll(t), 0);
        }
        if (!caps) {
            /* castling: rights, empty path, king not in / through / into check */
            const int ksq = us == 0 ? 4 : 60;
            if ((pc[K] & (1ULL << ksq)) && !attacked(b, ksq, them)) {
                const int ks = us == 0 ? WK : BK, qs = us == 0 ? WQ : BQ;
                if ((b->castling & ks) && (pc[R] & (1ULL << (ksq + 3))) &&
                    !(occ & ((1ULL << (ksq + 1)) | (1ULL << (ksq + 2)))) &&
                    !attacked(b, ksq + 1, them) && !attacked(b, ksq + 2, them))
                    out[n++] = ORC_MV(ksq, ksq + 2, 0);
                if ((b->castling & qs) && (pc[R] & (1ULL << (ksq - 4))) &&
                    !(occ & ((1ULL << (ksq - 1)) | (1ULL << (ksq - 2)) | (1ULL << (ksq - 3)))) &&
                    !attacked(b, ksq - 1, them) && !attacked(b, ksq - 2, them))
                    out[n++] = ORC_MV(ksq, ksq - 2, 0);
            }
        }
    }
    return n;
}

int orc_gen_legal(const orc_board* b, orc_move* out) {
    init_tables();
    orc_move tmp[ORC_MAX_MOVES];
    const int np = gen_pseudo(b, tmp);
    const int us = b->w_to_move ? 0 : 1;
    int n = 0;
    for (int i = 0; i < np; ++i) {
        orc_board nb;
        orc_make_move(b, tmp[i], &nb);
        if (!orc_in_check(&nb, us)) out[n++] = tmp[i];
    }
    return n;
}

uint64_t orc_perft(const orc_board* b, int depth) {
    if (depth == 0) return 1;
    orc_move mv[ORC_MAX_MOVES];
    const int n = orc_gen_legal(b, mv);
    if (depth == 1) return (uint64_t)n;
    uint64_t total = 0;
    for (int i = 0; i < n; ++i) {
        orc_board nb;
        orc_make_move(b, mv[i], &nb);
        total += orc_perft(&nb, depth - 1);
    }
    return total;
}

static uint64_t rng_next(uint64_t* s) { /* xorshift64* */
    uint64_t x = *s;
    x ^= x >> 12; x ^= x << 25; x ^= x >> 27;
    *s = x;
    return x * 0x2545F4914F6CDD1DULL;
}

int orc_random_positions(uint64_t seed, int n, int max_plies, orc_board* boards,
                         uint16_t* move_idx, uint32_t* move_off, orc_move* moves_raw) {
    init_tables();
    uint64_t s = seed ? seed : 0x9E3779B97F4A7C15ULL;
    int written = 0;
    uint32_t off = 0;
    if (move_off) move_off[0] = 0;
    while (written < n) {
        orc_board b;
        orc_startpos(&b);
        const int game_len = 1 + (int)(rng_next(&s) % (uint64_t)(max_plies > 0 ? max_plies : 1));
        for (int ply = 0; ply <= game_len && written < n; ++ply) {
            orc_move mv[ORC_MAX_MOVES];
            const int nm = orc_gen_legal(&b, mv);
            if (nm == 0 || b.halfmove >= 100) break;
            /* keep roughly one position in three along the game, always the last */
            if (ply == game_len || rng_next(&s) % 3 == 0) {
                boards[written] = b;
                for (int i = 0; i < nm; ++i) {
                    if (move_idx) move_idx[off + i] = (uint16_t)orc_move_to_index_stm(mv[i], b.w_to_move);
                    if (moves_raw) moves_raw[off + i] = mv[i];
                }
                off += (uint32_t)nm;
                written++;
                if (move_off) move_off[written] = off;
            }
            orc_board nb;
            orc_make_move(&b, mv[rng_next(&s) % (uint64_t)nm], &nb);
            b = nb;
        }
    }
    return written;
}
