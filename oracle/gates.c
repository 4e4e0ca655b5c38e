/* gates.c — the Tier-1 gates at full depth, restated from the reference (TEST INFRASTRUCTURE ONLY, see oracle.h).
 *
 *   mate search      src/search/mate_search.rs:64-108 (iterative deepening over mate-in-d, d = 1..max_depth; the first
 *                    exhaustive_depth levels try every legal attacker move, the deeper ones checking moves only),
 *                    :114-289 (solve_mate: pure minimax, any / all, the attacker-without-moves rule of :276-283 kept
 *                    as written: outside checks-only mode a side that has no legal move and stands in check counts as
 *                    "mate found" whichever side it is)
 *   hill in n        src/search/koth.rs:22-37 (rings), :43-65 (koth_center_in_n), :146-168 (the counted form),
 *                    :170-268 (solve_koth_counted: the root side's king must stay inside the ring that still reaches
 *                    the centre; the ring is computed from max_n, not from the iteration's n, as written)
 *   leaf value       src/mcts/tactical_mcts.rs:619-708 (hill first, then the mate; terminal_distance = the hill's n, or
 *                    the mate SCORE cast to u8, :701)
 *   adjudication     src/bin/self_play.rs:241-283 (hill in 3 and mate_search(board, 5, exhaustive 2))
 *
 * Moves are tried in the reference's generation order (orc_gen_pseudo_ref), so that the move returned and the node
 * counts are the reference's as well.  The transposition-table cache of :651-708 stores the result of this pure
 * function under the position and is not restated (same values; only the hit / miss statistics differ).
 *
 * Pins (tests/test_oracle_gates.py): every case of tests/unit/mate_search_tests.rs and tests/unit/koth_tests.rs that
 * calls these functions. */
#include "oracle.h"
#include <string.h>

#define K 5
#define CENTER 0x0000001818000000ULL

static int is_checkmate_d0(const orc_board* b, int checks_only) { /* mate_search.rs:129-238 */
    const int us = b->w_to_move ? 0 : 1;
    if (!orc_in_check(b, us)) return 0; /* in checks-only mode the reference skips this test: the last move was a check (:133-137) */
    (void)checks_only;
    orc_move mv[ORC_MAX_MOVES];
    return orc_gen_legal(b, mv) == 0;
}

static int solve_mate(const orc_board* b, int depth, int attacker, int checks_only, int* nodes) { /* :114-289 */
    ++*nodes;
    if (depth <= 0) return is_checkmate_d0(b, checks_only);
    const int us = b->w_to_move ? 0 : 1;
    orc_move mv[ORC_MAX_MOVES];
    int ncap;
    const int n = orc_gen_pseudo_ref(b, mv, &ncap);
    int has_legal = 0;
    for (int i = 0; i < n; ++i) {
        orc_board nb;
        orc_make_move(b, mv[i], &nb);
        if (attacker && checks_only && !orc_in_check(&nb, 1 - us)) continue; /* gives_check, src/board.rs:682-777 */
        if (orc_in_check(&nb, us)) continue;                                    /* is_legal_after_move, :785 */
        has_legal = 1;
        const int r = solve_mate(&nb, depth - 1, !attacker, checks_only, nodes);
        if (attacker) { if (r) return 1; }
        else if (!r) return 0;
    }
    if (!has_legal) {
        if (attacker && checks_only) return 0;
        return orc_in_check(b, us);
    }
    return !attacker;
}

int orc_mate_search(const orc_board* b, int max_depth, int exhaustive_depth, orc_move* best, int* nodes_out) { /* :64-108 */
    int nodes = 0;
    const int exhaustive_plies = exhaustive_depth * 2 - 1;
    const int us = b->w_to_move ? 0 : 1;
    if (best) *best = 0;
    for (int d = 1; d <= max_depth; ++d) {
        const int depth = 2 * d - 1;
        const int checks_only = depth > exhaustive_plies;
        orc_move mv[ORC_MAX_MOVES];
        int ncap;
        const int n = orc_gen_pseudo_ref(b, mv, &ncap);
        for (int i = 0; i < n; ++i) {
            orc_board nb;
            orc_make_move(b, mv[i], &nb);
            if (checks_only && !orc_in_check(&nb, 1 - us)) continue;
            if (orc_in_check(&nb, us)) continue;
            if (solve_mate(&nb, depth - 1, 0, checks_only, &nodes)) {
                if (best) *best = mv[i];
                if (nodes_out) *nodes_out = nodes;
                return 1000000 + depth;
            }
        }
    }
    if (nodes_out) *nodes_out = nodes;
    return 0;
}

/* ---- King of the Hill ---- */
#define RING_1 (0x00003C24243C0000ULL & ~CENTER)
#define RING_2 (0x007E424242427E00ULL & ~(RING_1 | CENTER))
#define RING_3 0xFF818181818181FFULL

static uint64_t cumulative_ring_mask(int remaining) { /* koth.rs:29-37 */
    switch (remaining) {
    case 0: return CENTER;
    case 1: return CENTER | RING_1;
    case 2: return CENTER | RING_1 | RING_2;
    default: return CENTER | RING_1 | RING_2 | RING_3;
    }
}

static const int8_t KING_STEPS[8][2] = {{1, 0}, {1, 1}, {0, 1}, {-1, 1}, {-1, 0}, {-1, -1}, {0, -1}, {1, -1}};
static uint64_t king_moves_bb(int sq) {
    uint64_t m = 0;
    const int r = sq >> 3, f = sq & 7;
    for (int i = 0; i < 8; ++i) {
        const int rr = r + KING_STEPS[i][0], ff = f + KING_STEPS[i][1];
        if (rr >= 0 && rr < 8 && ff >= 0 && ff < 8) m |= 1ULL << (rr * 8 + ff);
    }
    return m;
}

static int solve_koth(const orc_board* b, int ply, int max_ply, int max_n, uint32_t* nodes) { /* koth.rs:170-268 */
    ++*nodes;
    const int w_won = (b->pieces[K] & CENTER) != 0, b_won = (b->pieces[6 + K] & CENTER) != 0;
    const int root_turn = ply % 2 == 0;
    const int root_white = root_turn ? b->w_to_move : !b->w_to_move;
    if (root_white) { if (w_won) return 1; if (b_won) return 0; }
    else { if (b_won) return 1; if (w_won) return 0; }
    if (ply > max_ply) return 0;
    const int stm = b->w_to_move ? 0 : 1;
    orc_move mv[ORC_MAX_MOVES];
    int ncap;
    if (root_turn) {
        const int root_side = root_white ? 0 : 1;
        const uint64_t king = b->pieces[root_side * 6 + K];
        const int ks = __builtin_ctzll(king);
        int remaining = max_n - ply / 2 - 1;
        if (remaining < 0) remaining = 0;
        const uint64_t target = cumulative_ring_mask(remaining);
        if (!(king & target)) {
            for (uint64_t c = king_moves_bb(ks) & target & ~b->occ[root_side]; c; c &= c - 1) {
                orc_board nb;
                orc_make_move(b, ORC_MV(ks, __builtin_ctzll(c), 0), &nb);
                if (orc_in_check(&nb, stm)) continue;
                if (solve_koth(&nb, ply + 1, max_ply, max_n, nodes)) return 1;
            }
            return 0;
        }
        const int n = orc_gen_pseudo_ref(b, mv, &ncap);
        for (int i = 0; i < n; ++i) {
            orc_board nb;
            orc_make_move(b, mv[i], &nb);
            if (orc_in_check(&nb, stm)) continue;
            if (!(nb.pieces[root_side * 6 + K] & target)) continue;
            if (solve_koth(&nb, ply + 1, max_ply, max_n, nodes)) return 1;
        }
        return 0;
    }
    const int n = orc_gen_pseudo_ref(b, mv, &ncap);
    int legal = 0;
    for (int i = 0; i < n; ++i) {
        orc_board nb;
        orc_make_move(b, mv[i], &nb);
        if (orc_in_check(&nb, stm)) continue;
        legal = 1;
        if (!solve_koth(&nb, ply + 1, max_ply, max_n, nodes)) return 0;
    }
    return legal;
}

/* koth_center_in_n_counted, koth.rs:146-168: the number of own moves (0..max_n), or -1 (None) */
int orc_koth_center_in_n(const orc_board* b, int max_n, uint32_t* nodes_out) {
    const int us = b->w_to_move ? 0 : 1;
    uint32_t total = 1;
    int res = -1;
    if (!b->pieces[us * 6 + K]) res = -1;
    else if (b->pieces[us * 6 + K] & CENTER) res = 0;
    else
        for (int n = 1; n <= max_n; ++n) {
            uint32_t nodes = 0;
            const int r = solve_koth(b, 0, 2 * n - 1, max_n, &nodes);
            total += nodes;
            if (r) { res = n; break; }
        }
    if (nodes_out) *nodes_out = total;
    return res;
}

/* The Tier-1 leaf gate, src/mcts/tactical_mcts.rs:619-708: -1 / +1 for the side to move, 2 = no gate.
 * *dist = terminal_distance of a hit. */
static long g_gate_calls, g_gate_nodes, g_gate_max_nodes; /* positions entered by the leaf gates (a cost statistic for the tests) */
void orc_gate_node_stats(long* calls, long* nodes, long* max_nodes, int reset) {
    if (calls) *calls = g_gate_calls;
    if (nodes) *nodes = g_gate_nodes;
    if (max_nodes) *max_nodes = g_gate_max_nodes;
    if (reset) g_gate_calls = g_gate_nodes = g_gate_max_nodes = 0;
}

int orc_tier1_gates(const orc_board* b, int koth, int mate_depth, int exhaustive_depth, int koth_depth, int* dist) {
    int d = 0, v = 2;
    long total = 0;
    if (koth) {
        if (b->pieces[(b->w_to_move ? 1 : 0) * 6 + K] & CENTER) { v = -1; d = 0; goto done; }
        uint32_t hn = 0;
        const int n = orc_koth_center_in_n(b, koth_depth, &hn);
        total += hn;
        if (n >= 0) { v = 1; d = n; goto done; }
    }
    if (mate_depth > 0) {
        int mn = 0;
        const int score = orc_mate_search(b, mate_depth, exhaustive_depth, NULL, &mn);
        total += mn;
        if (score != 0) { v = 1; d = (int)(uint8_t)(unsigned)score; } /* mate_result.0.unsigned_abs() as u8, :701 */
    }
done:
    g_gate_calls++;
    g_gate_nodes += total;
    if (total > g_gate_max_nodes) g_gate_max_nodes = total;
    if (dist) *dist = d;
    return v;
}
