/* codec.c — CPU restatement of the reference's feature / policy / value codec.
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Plain scalar C on purpose: this is
 * the checker, written for obviousness, not speed. */
#include "oracle.h"
#include "pesto_tables.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

enum { WK = 1, WQ = 2, BK = 4, BQ = 8 };

static void recompute_occ(orc_board* b) {
    for (int c = 0; c < 2; ++c) {
        uint64_t o = 0;
        for (int p = 0; p < 6; ++p) o |= b->pieces[c * 6 + p];
        b->occ[c] = o;
    }
}

/* src/board.rs:115-199 — FEN fields: placement, side, castling, ep, halfmove. */
int orc_parse_fen(const char* fen, orc_board* out) {
    memset(out, 0, sizeof(*out));
    out->en_passant = 0xFF;
    const char* s = fen;
    int rank = 7, file = 0;
    for (; *s && *s != ' '; ++s) {
        char c = *s;
        if (c == '/') { rank--; file = 0; continue; }
        if (c >= '1' && c <= '8') { file += c - '0'; continue; }
        const char* sym = "PNBRQKpnbrqk";
        const char* hit = strchr(sym, c);
        if (!hit || rank < 0 || file > 7) return -1;
        out->pieces[hit - sym] |= 1ULL << (rank * 8 + file);
        file++;
    }
    if (*s != ' ') return -1;
    ++s;
    if (*s == 'w') out->w_to_move = 1; else if (*s == 'b') out->w_to_move = 0; else return -1;
    ++s;
    if (*s != ' ') return -1;
    ++s;
    for (; *s && *s != ' '; ++s) {
        switch (*s) {
            case 'K': out->castling |= WK; break;
            case 'Q': out->castling |= WQ; break;
            case 'k': out->castling |= BK; break;
            case 'q': out->castling |= BQ; break;
            case '-': break;
            default: return -1;
        }
    }
    if (*s == ' ') ++s;
    if (*s && *s != '-') {
        if (s[0] < 'a' || s[0] > 'h' || s[1] < '1' || s[1] > '8') return -1;
        out->en_passant = (uint8_t)((s[1] - '1') * 8 + (s[0] - 'a'));
        s += 2;
    } else if (*s == '-') ++s;
    if (*s == ' ') ++s;
    if (*s) out->halfmove = (uint8_t)atoi(s);
    recompute_occ(out);
    return 0;
}

void orc_startpos(orc_board* out) {
    orc_parse_fen("rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1", out);
}

/* src/tensor.rs:21-77.  Planes 0-5 side to move P..K, 6-11 opponent, 12 en
 * passant, 13-16 castling (stm-K, stm-Q, opp-K, opp-Q) broadcast.  Row is
 * 7-rank when white is to move, rank when black is; column is the file. */
void orc_board_to_planes(const orc_board* b, float* planes) {
    memset(planes, 0, sizeof(float) * ORC_PLANES);
    const int white = b->w_to_move != 0;
    const int stm = white ? 0 : 1;
    for (int side = 0; side < 2; ++side) {
        const int color = side == 0 ? stm : 1 - stm;
        for (int pt = 0; pt < 6; ++pt) {
            const uint64_t bb = b->pieces[color * 6 + pt];
            for (int sq = 0; sq < 64; ++sq) {
                if (!((bb >> sq) & 1)) continue;
                const int rank = sq >> 3, file = sq & 7;
                const int row = white ? 7 - rank : rank;
                planes[(side * 6 + pt) * 64 + row * 8 + file] = 1.0f;
            }
        }
    }
    if (b->en_passant != 0xFF) {
        const int rank = b->en_passant >> 3, file = b->en_passant & 7;
        const int row = white ? 7 - rank : rank;
        planes[12 * 64 + row * 8 + file] = 1.0f;
    }
    const int order_w[4] = {WK, WQ, BK, BQ}, order_b[4] = {BK, BQ, WK, WQ};
    const int* order = white ? order_w : order_b;
    for (int i = 0; i < 4; ++i)
        if (b->castling & order[i])
            for (int j = 0; j < 64; ++j) planes[(13 + i) * 64 + j] = 1.0f;
}

/* src/tensor.rs:82-178.  index = from*73 + plane; planes 0-55 slides
 * dir*7+(dist-1) with dir N,NE,E,SE,S,SW,W,NW; 56-63 knight; 64-72
 * under-promotion; queen promotion counts as a slide. */
int orc_move_to_index(int from, int to, int promo) {
    const int dx = (to & 7) - (from & 7);
    const int dy = (to >> 3) - (from >> 3);
    int plane;
    if (promo != 0 && promo != 4) {
        int dir, piece;
        if (dx == 0) dir = 0; else if (dx == -1) dir = 1; else if (dx == 1) dir = 2; else return -1;
        if (promo == 1) piece = 0; else if (promo == 2) piece = 3; else if (promo == 3) piece = 6; else return -1;
        plane = 64 + dir + piece;
    } else if (abs(dx * dy) == 2) {
        static const int kdx[8] = {1, 2, 2, 1, -1, -2, -2, -1};
        static const int kdy[8] = {2, 1, -1, -2, -2, -1, 1, 2};
        plane = -1;
        for (int i = 0; i < 8; ++i)
            if (kdx[i] == dx && kdy[i] == dy) plane = 56 + i;
        if (plane < 0) return -1;
    } else {
        int dir;
        if (dx == 0 && dy > 0) dir = 0;
        else if (dx > 0 && dy > 0) dir = 1;
        else if (dx > 0 && dy == 0) dir = 2;
        else if (dx > 0 && dy < 0) dir = 3;
        else if (dx == 0 && dy < 0) dir = 4;
        else if (dx < 0 && dy < 0) dir = 5;
        else if (dx < 0 && dy == 0) dir = 6;
        else if (dx < 0 && dy > 0) dir = 7;
        else return -1;
        const int dist = abs(dx) > abs(dy) ? abs(dx) : abs(dy);
        plane = dir * 7 + (dist - 1);
    }
    return from * 73 + plane;
}

/* Black moves are mirrored rank-wise first (src/move_types.rs:374-381,
 * src/board_utils.rs:144-146: sq -> 8*(7-sq/8)+sq%8 == sq^56). */
int orc_move_to_index_stm(orc_move m, int w_to_move) {
    int f = ORC_FROM(m), t = ORC_TO(m);
    if (!w_to_move) { f ^= 56; t ^= 56; }
    return orc_move_to_index(f, t, ORC_PROMO(m));
}

/* src/tensor.rs:184-213: gather, sequential f32 sum in list order, divide;
 * uniform 1/n if the sum is not positive; out-of-range index contributes 0. */
void orc_policy_to_move_priors(const float* policy, int n_policy, const orc_move* moves, int n,
                               int w_to_move, float* priors) {
    float total = 0.0f;
    for (int i = 0; i < n; ++i) {
        const int idx = orc_move_to_index_stm(moves[i], w_to_move);
        if (idx >= 0 && idx < n_policy) {
            priors[i] = policy[idx];
            total += policy[idx];
        } else {
            priors[i] = 0.0f;
        }
    }
    if (total > 0.0f) {
        for (int i = 0; i < n; ++i) priors[i] /= total;
    } else {
        const float u = 1.0f / (float)n;
        for (int i = 0; i < n; ++i) priors[i] = u;
    }
}

/* src/eval.rs:35-57 (table build: white looks up sq^56) and :93-122 (tapered
 * sum, i32 truncating division, sign for side to move). */
int orc_pst_eval_cp(const orc_board* b) {
    int mg[2] = {0, 0}, eg[2] = {0, 0}, phase = 0;
    for (int color = 0; color < 2; ++color)
        for (int pt = 0; pt < 6; ++pt) {
            uint64_t bb = b->pieces[color * 6 + pt];
            while (bb) {
                const int sq = __builtin_ctzll(bb);
                const int t = color == 0 ? (sq ^ 56) : sq;
                mg[color] += ORC_MG_VALUE[pt] + ORC_MG_PST[pt][t];
                eg[color] += ORC_EG_VALUE[pt] + ORC_EG_PST[pt][t];
                phase += ORC_PHASE_INC[pt];
                bb &= bb - 1;
            }
        }
    const int mgp = phase < 24 ? phase : 24;
    const int score = ((mg[0] - mg[1]) * mgp + (eg[0] - eg[1]) * (24 - mgp)) / 24;
    return b->w_to_move ? score : -score;
}

/* src/mcts/tactical_mcts.rs:762-765 (material on) and :787-790 (material off):
 * the host fuses in f64. */
double orc_value_fuse(float v_logit, float k, float q, int material_on) {
    if (material_on) return tanh((double)v_logit + (double)k * (double)q);
    return tanh((double)v_logit);
}
