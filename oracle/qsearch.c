/* qsearch.c — CPU restatement of the reference's principal-exchange quiescence line, the cheapest of the
 * three Tier-2 "material" searches that produce the dM input of the leaf evaluation
 * (src/mcts/tactical_mcts.rs:715-731, variant "pe").  TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Structure follows the reference on purpose: build the capture list in the reference's generation
 * order, sort it by MVV-LVA, recurse on the first legal entry.
 *
 *   capture list      src/move_generation.rs:287-309,426-434  pawn captures, knights, bishops, rooks, queens,
 *                                                              kings, then ALL promotions (capturing or not)
 *   pawns             src/move_generation.rs:461-560          per pawn (ascending square): capture targets in
 *                     src/magic_bitboard.rs:204-235           table order (white +7,+9; black -9,-7; en passant
 *                                                              counts), then the push; on the 7th rank every
 *                     src/magic_bitboard.rs:457-474           target yields Q,R,N,B into the promotion list
 *   knights / kings   src/move_generation.rs:575-606,650-703  table order of src/magic_bitboard.rs:120-189
 *   sliders           src/move_generation.rs:721-1000         attack set iterated from the lowest square; the
 *                                                              queen does its rook set first, then its bishop set
 *   bit iteration     src/bits.rs:27-38                       lowest set bit first
 *   MVV-LVA           src/move_generation.rs:436-446          10*victim - attacker with P,N,B,R,Q,K = 0..5; 0 when
 *                                                              the target square is empty (en passant, promotion push)
 *   sort              src/move_generation.rs:431              sort_unstable_by_key(-score).  Rust's sort_unstable is
 *                                                              an insertion sort for slices of <= 20 elements (both the
 *                                                              pdqsort and the ipnsort generations of the standard
 *                                                              library), i.e. equal keys keep generation order.  Longer
 *                                                              lists are NOT modelled: *long_list is set and the same
 *                                                              stable order is used (tie order unpinned there).
 *   the line          src/search/quiescence.rs:756-794        stand pat, cut-offs, first legal capture only
 *   wrapper           src/search/quiescence.rs:797-804        window (-100000, 100000), depth 20, cp / 100 as f32
 *
 * Pins: the known answers of tests/unit/ext_qsearch_tests.rs:306-350,546-575 (node counts and inequalities; the
 * reference holds no exact scores for this function), replayed in tests/test_oracle_qsearch.py.  The tie order of
 * equal MVV-LVA keys cannot be pinned against the Rust crate here (no Rust toolchain): it rests on the citations above.
 */
#include "oracle.h"
#include <string.h>

enum { P = 0, N = 1, B = 2, R = 3, Q = 4, K = 5 };

static int lsb(uint64_t b) { return __builtin_ctzll(b); }

static int type_at(const orc_board* b, int sq) { /* Board::get_piece(sq).1, either colour; -1 if empty */
    const uint64_t bit = 1ULL << sq;
    for (int i = 0; i < 12; ++i)
        if (b->pieces[i] & bit) return i % 6;
    return -1;
}

static uint64_t walk(int sq, uint64_t occ, const int* dx, const int* dy) {
    uint64_t att = 0;
    for (int d = 0; d < 4; ++d) {
        int r = (sq >> 3) + dy[d], f = (sq & 7) + dx[d];
        for (; r >= 0 && r < 8 && f >= 0 && f < 8; r += dy[d], f += dx[d]) {
            att |= 1ULL << (r * 8 + f);
            if (occ >> (r * 8 + f) & 1) break;
        }
    }
    return att;
}
static uint64_t rook_set(int sq, uint64_t occ) {
    static const int dx[4] = {0, 1, 0, -1}, dy[4] = {1, 0, -1, 0};
    return walk(sq, occ, dx, dy);
}
static uint64_t bishop_set(int sq, uint64_t occ) {
    static const int dx[4] = {1, 1, -1, -1}, dy[4] = {1, -1, -1, 1};
    return walk(sq, occ, dx, dy);
}

static int mvv_lva(const orc_board* b, int from, int to) {
    const int victim = type_at(b, to);
    if (victim < 0) return 0;
    return 10 * victim - type_at(b, from);
}

static int push4(orc_move* v, int n, int from, int to) {
    v[n++] = ORC_MV(from, to, Q);
    v[n++] = ORC_MV(from, to, R);
    v[n++] = ORC_MV(from, to, N);
    v[n++] = ORC_MV(from, to, B);
    return n;
}

int orc_gen_captures_mvvlva(const orc_board* b, orc_move* out, int* long_list) {
    orc_move caps[ORC_MAX_MOVES], promos[ORC_MAX_MOVES];
    int nc = 0, np = 0;
    const int us = b->w_to_move ? 0 : 1, them = 1 - us;
    const uint64_t own = b->occ[us], opp = b->occ[them], occ = own | opp;
    const uint64_t* pc = &b->pieces[us * 6];

    for (uint64_t bb = pc[P]; bb; bb &= bb - 1) {
        const int from = lsb(bb), r = from >> 3, f = from & 7;
        if (r == 0 || r == 7) continue; /* no table entries on the back ranks */
        const int promo = us == 0 ? r == 6 : r == 1;
        const int left = us == 0 ? from + 7 : from - 9, right = us == 0 ? from + 9 : from - 7;
        const int tgt[2] = {f > 0 ? left : -1, f < 7 ? right : -1};
        for (int i = 0; i < 2; ++i) {
            const int to = tgt[i];
            if (to < 0) continue;
            if ((opp >> to & 1) || (b->en_passant != 0xFF && b->en_passant == to)) {
                if (promo) np = push4(promos, np, from, to);
                else caps[nc++] = ORC_MV(from, to, 0);
            }
        }
        const int one = us == 0 ? from + 8 : from - 8;
        if (promo && !(occ >> one & 1)) np = push4(promos, np, from, one);
    }
    static const int n_off[8] = {15, 17, 6, 10, -10, -6, -17, -15};
    static const int n_df[8] = {-1, 1, -2, 2, -2, 2, -1, 1}, n_dr[8] = {2, 2, 1, 1, -1, -1, -2, -2};
    for (uint64_t bb = pc[N]; bb; bb &= bb - 1) {
        const int from = lsb(bb);
        for (int i = 0; i < 8; ++i) {
            const int r = (from >> 3) + n_dr[i], f = (from & 7) + n_df[i];
            if (r < 0 || r > 7 || f < 0 || f > 7) continue;
            const int to = from + n_off[i];
            if (opp >> to & 1) caps[nc++] = ORC_MV(from, to, 0);
        }
    }
    for (uint64_t bb = pc[B]; bb; bb &= bb - 1) {
        const int from = lsb(bb);
        for (uint64_t t = bishop_set(from, occ) & opp; t; t &= t - 1) caps[nc++] = ORC_MV(from, lsb(t), 0);
    }
    for (uint64_t bb = pc[R]; bb; bb &= bb - 1) {
        const int from = lsb(bb);
        for (uint64_t t = rook_set(from, occ) & opp; t; t &= t - 1) caps[nc++] = ORC_MV(from, lsb(t), 0);
    }
    for (uint64_t bb = pc[Q]; bb; bb &= bb - 1) {
        const int from = lsb(bb);
        for (uint64_t t = rook_set(from, occ) & opp; t; t &= t - 1) caps[nc++] = ORC_MV(from, lsb(t), 0);
        for (uint64_t t = bishop_set(from, occ) & opp; t; t &= t - 1) caps[nc++] = ORC_MV(from, lsb(t), 0);
    }
    static const int k_off[8] = {-1, 7, 8, 9, 1, -7, -8, -9};
    static const int k_df[8] = {-1, -1, 0, 1, 1, 1, 0, -1}, k_dr[8] = {0, 1, 1, 1, 0, -1, -1, -1};
    for (uint64_t bb = pc[K]; bb; bb &= bb - 1) {
        const int from = lsb(bb);
        for (int i = 0; i < 8; ++i) {
            const int r = (from >> 3) + k_dr[i], f = (from & 7) + k_df[i];
            if (r < 0 || r > 7 || f < 0 || f > 7) continue;
            const int to = from + k_off[i];
            if (opp >> to & 1) caps[nc++] = ORC_MV(from, to, 0);
        }
    }
    int n = 0, key[ORC_MAX_MOVES];
    for (int i = 0; i < nc; ++i) out[n++] = caps[i];
    for (int i = 0; i < np; ++i) out[n++] = promos[i];
    if (long_list) *long_list = n > 20;
    for (int i = 0; i < n; ++i) key[i] = -mvv_lva(b, ORC_FROM(out[i]), ORC_TO(out[i]));
    for (int i = 1; i < n; ++i) { /* insertion sort, equal keys stay in place */
        const orc_move m = out[i];
        const int k = key[i];
        int j = i;
        for (; j > 0 && key[j - 1] > k; --j) {
            out[j] = out[j - 1];
            key[j] = key[j - 1];
        }
        out[j] = m;
        key[j] = k;
    }
    return n;
}

int orc_principal_exchange(const orc_board* b, int alpha, int beta, int max_depth, uint32_t* nodes, int* long_list) {
    if (nodes) *nodes += 1;
    const int stand_pat = orc_pst_eval_cp(b);
    if (stand_pat >= beta) return beta;
    if (stand_pat > alpha) alpha = stand_pat;
    if (max_depth == 0) return alpha;
    orc_move caps[ORC_MAX_MOVES];
    int ll = 0;
    const int n = orc_gen_captures_mvvlva(b, caps, &ll);
    if (ll && long_list) *long_list = 1;
    const int mover = b->w_to_move ? 0 : 1;
    for (int i = 0; i < n; ++i) {
        orc_board nb;
        orc_make_move(b, caps[i], &nb);
        if (orc_in_check(&nb, mover)) continue; /* Board::is_legal, src/board.rs:352 */
        const int score = -orc_principal_exchange(&nb, -beta, -alpha, max_depth - 1, nodes, long_list);
        if (score >= beta) return beta;
        if (score > alpha) alpha = score;
        break;
    }
    return alpha;
}

float orc_forced_principal_exchange(const orc_board* b, uint32_t* nodes, int* long_list) {
    if (nodes) *nodes = 0;
    if (long_list) *long_list = 0;
    const int cp = orc_principal_exchange(b, -100000, 100000, 20, nodes, long_list);
    return (float)cp / 100.0f;
}

/* ---- cap = 1 quiescence search (variant "cap1"), src/search/quiescence.rs:808-1012 ----
 * At a node not in check: stand pat (cut-off / raise alpha), then the FIRST LEGAL capture of the MVV-LVA list, then - once
 * per side and branch - the first legal move of the generation-order list that gives check and is not "that capture"
 * (identified as written, :933-944: by the from-square of the first entry of the capture list whose target is the played
 * capture's target, and that target).  In check: every legal evasion, no stand pat; no legal move = -1_000_000.  Depth 0: in
 * check = not completed.  Returns the score; *completed / *nodes / *depth_used as the reference's tuple.
 * Pins: tests/unit/ext_qsearch_tests.rs:354-409,571-586 (inequalities, completion, node bound). */
int orc_cap1_qsearch(const orc_board* b, int alpha, int beta, int max_depth, int w_check_used, int b_check_used, int* completed,
                     uint32_t* nodes_out, int* depth_used, int* long_list) {
    uint32_t nodes = 1;
    int all_completed = 1, max_child_depth = 0, ret;
    const int us = b->w_to_move ? 0 : 1;
    const int in_check = orc_in_check(b, us);
#define CAP1_RETURN(v, c, d) do { ret = (v); *completed = (c); *nodes_out = nodes; *depth_used = (d); return ret; } while (0)
    if (!in_check) {
        const int stand_pat = orc_pst_eval_cp(b);
        if (stand_pat >= beta) CAP1_RETURN(beta, 1, 0);
        if (stand_pat > alpha) alpha = stand_pat;
    }
    if (max_depth == 0) CAP1_RETURN(alpha, in_check ? 0 : 1, 0);
    orc_move mv[ORC_MAX_MOVES];
    int ncap;
    if (in_check) {
        const int n = orc_gen_pseudo_ref(b, mv, &ncap);
        int any_legal = 0;
        for (int i = 0; i < n; ++i) {
            orc_board nb;
            orc_make_move(b, mv[i], &nb);
            if (orc_in_check(&nb, us)) continue;
            any_legal = 1;
            int cc, cd;
            uint32_t cn;
            const int score = -orc_cap1_qsearch(&nb, -beta, -alpha, max_depth - 1, w_check_used, b_check_used, &cc, &cn, &cd, long_list);
            nodes += cn;
            if (cd + 1 > max_child_depth) max_child_depth = cd + 1;
            if (!cc) all_completed = 0;
            if (score >= beta) CAP1_RETURN(beta, all_completed, max_child_depth);
            if (score > alpha) alpha = score;
        }
        if (!any_legal) CAP1_RETURN(-1000000, 1, 0);
        CAP1_RETURN(alpha, all_completed, max_child_depth);
    }
    orc_move caps[ORC_MAX_MOVES];
    int ll = 0;
    const int nc = orc_gen_captures_mvvlva(b, caps, &ll);
    if (ll && long_list) *long_list = 1;
    int best_to = -1;
    for (int i = 0; i < nc; ++i) {
        orc_board nb;
        orc_make_move(b, caps[i], &nb);
        if (orc_in_check(&nb, us)) continue;
        best_to = ORC_TO(caps[i]);
        int cc, cd;
        uint32_t cn;
        const int score = -orc_cap1_qsearch(&nb, -beta, -alpha, max_depth - 1, w_check_used, b_check_used, &cc, &cn, &cd, long_list);
        nodes += cn;
        if (cd + 1 > max_child_depth) max_child_depth = cd + 1;
        if (!cc) all_completed = 0;
        if (score >= beta) CAP1_RETURN(beta, all_completed, max_child_depth);
        if (score > alpha) alpha = score;
        break;
    }
    if (!(us == 0 ? w_check_used : b_check_used)) {
        int skip_from = -1;
        if (best_to >= 0)
            for (int i = 0; i < nc; ++i)
                if (ORC_TO(caps[i]) == best_to) { skip_from = ORC_FROM(caps[i]); break; }
        const int n = orc_gen_pseudo_ref(b, mv, &ncap);
        for (int i = 0; i < n; ++i) {
            if (best_to >= 0 && ORC_FROM(mv[i]) == skip_from && ORC_TO(mv[i]) == best_to) continue;
            orc_board nb;
            orc_make_move(b, mv[i], &nb);
            if (!orc_in_check(&nb, 1 - us)) continue; /* gives_check, src/board.rs:682-777 */
            if (orc_in_check(&nb, us)) continue;
            int cc, cd;
            uint32_t cn;
            const int score = -orc_cap1_qsearch(&nb, -beta, -alpha, max_depth - 1, us == 0 ? 1 : w_check_used, us == 1 ? 1 : b_check_used,
                                                &cc, &cn, &cd, long_list);
            nodes += cn;
            if (cd + 1 > max_child_depth) max_child_depth = cd + 1;
            if (!cc) all_completed = 0;
            if (score >= beta) CAP1_RETURN(beta, all_completed, max_child_depth);
            if (score > alpha) alpha = score;
            break;
        }
    }
    CAP1_RETURN(alpha, all_completed, max_child_depth);
#undef CAP1_RETURN
}

float orc_forced_cap1(const orc_board* b, int* completed, uint32_t* nodes, int* depth_used, int* long_list) { /* :1000-1012 */
    int c = 1, d = 0, ll = 0;
    uint32_t n = 0;
    const int cp = orc_cap1_qsearch(b, -100000, 100000, 20, 0, 0, &c, &n, &d, &ll);
    if (completed) *completed = c;
    if (nodes) *nodes = n;
    if (depth_used) *depth_used = d;
    if (long_list) *long_list = ll;
    return (float)cp / 100.0f;
}

/* ---- light Tier-1 gates: the reference's gate at depth 1 ----
 *   mate in one     src/search/mate_search.rs:64-108 with max_depth = 1: a legal (checking) move after which the
 *                   opponent is in check without a legal move (:129-200)
 *   hill in one     src/search/koth.rs:43-65 with max_n = 1: own king on a centre square, or (:213-247) a legal king
 *                   step onto one; false when the opponent's king already stands on one (:185-209)
 *   leaf value      src/mcts/tactical_mcts.rs:619-708: opponent on the hill -> -1, hill in one -> +1, mate in one -> +1
 * Pins: tests/unit/mate_search_tests.rs:26-33,346-351,448-474 and tests/unit/koth_tests.rs:19-37. */
#define ORC_CENTER 0x0000001818000000ULL

int orc_mate_in_1(const orc_board* b) {
    orc_move mv[ORC_MAX_MOVES], reply[ORC_MAX_MOVES];
    const int n = orc_gen_legal(b, mv);
    for (int i = 0; i < n; ++i) {
        orc_board nb;
        orc_make_move(b, mv[i], &nb);
        if (orc_in_check(&nb, nb.w_to_move ? 0 : 1) && orc_gen_legal(&nb, reply) == 0) return 1;
    }
    return 0;
}

int orc_koth_in_1(const orc_board* b) {
    const int us = b->w_to_move ? 0 : 1;
    const uint64_t k = b->pieces[us * 6 + K];
    if (k & ORC_CENTER) return 1;
    if (b->pieces[(1 - us) * 6 + K] & ORC_CENTER) return 0;
    orc_move mv[ORC_MAX_MOVES];
    const int n = orc_gen_legal(b, mv);
    for (int i = 0; i < n; ++i)
        if ((k >> ORC_FROM(mv[i]) & 1) && (ORC_CENTER >> ORC_TO(mv[i]) & 1)) return 1;
    return 0;
}

int orc_light_gates(const orc_board* b, int koth) { /* -1 / +1 for the side to move, 2 = no gate */
    if (koth) {
        if (b->pieces[(b->w_to_move ? 1 : 0) * 6 + K] & ORC_CENTER) return -1;
        if (orc_koth_in_1(b)) return 1;
    }
    return orc_mate_in_1(b) ? 1 : 2;
}
