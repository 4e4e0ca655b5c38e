// gates.hpp — the light Tier-1 gates of the leaf evaluation: "the side to move mates in one" and,
// under King-of-the-Hill rules, "the side to move's king is on, or steps legally onto, a centre square".  They are the
// reference's Tier-1 gate at depth 1: mate_search(board, max_depth = 1) (src/search/mate_search.rs:64-108: at depth 1
// every candidate is a checking move whose reply position has no legal move) and koth_center_in_n(board, 1)
// (src/search/koth.rs:43-65 with :213-247 for one root move) — the same two quick checks the reference's own GPU
// search applies to a freshly expanded node (cuda/quick_checks.cuh:10-17).  A hit makes the node terminal with
// value +1 for the side to move (src/mcts/tactical_mcts.rs:636-647,695-707); the self-play driver adjudicates a game
// whose root position hits a gate (src/bin/self_play.rs:241-283).
//
// One inline host+device statement (scalar forms); the warp form of the tree kernels is light_gates_warp in
// mcts_device.cuh.  Product code: it does not
// use oracle/; tests/ check it against oracle/qsearch.c's independent statement and the reference's known answers.
#pragma once
#include "chess_rules.hpp"

namespace cbchess {

// Does the side to move have a legal move?  King steps first (the usual answer to a check), then everything else.
CB_HD bool has_legal_move_hd(const cb_board& b) {
    const Tables& T = tables();
    const int us = b.w_to_move ? 0 : 1;
    const uint64_t k = b.pieces[us * 6 + KING];
    cb_board nb;
    if (k) {
        const int ks = lsb(k);
        for (uint64_t t = T.king[ks] & ~b.occ[us]; t; t &= t - 1) {
            make_move_hd(b, make_mv(ks, lsb(t), 0), &nb);
            if (!in_check_hd(nb, us)) return true;
        }
    }
    Move tmp[kMaxMoves];
    const int n = gen_pseudo(b, tmp);
    for (int i = 0; i < n; ++i) {
        make_move_hd(b, tmp[i], &nb);
        if (!in_check_hd(nb, us)) return true;
    }
    return false;
}

// `m` is legal in `b`: does it checkmate?
CB_HD bool move_mates_hd(const cb_board& b, Move m) {
    cb_board nb;
    make_move_hd(b, m, &nb);
    return in_check_hd(nb, nb.w_to_move ? 0 : 1) && !has_legal_move_hd(nb);
}

// `legal[0..n)`: the legal moves of `b`.
CB_HD bool mate_in_1_hd(const cb_board& b, const Move* legal, int n) {
    for (int i = 0; i < n; ++i)
        if (move_mates_hd(b, legal[i])) return true;
    return false;
}

CB_HD bool koth_in_1_hd(const cb_board& b, const Move* legal, int n) {
    const int us = b.w_to_move ? 0 : 1;
    const uint64_t k = b.pieces[us * 6 + KING];
    if (k & kKothCenter) return true;                               // koth_center_in_n: Some(0)
    if (b.pieces[(1 - us) * 6 + KING] & kKothCenter) return false;  // solve_koth: the opponent already won
    if (!k) return false;
    const int ks = lsb(k);
    for (int i = 0; i < n; ++i)
        if (mv_from(legal[i]) == ks && (kKothCenter >> mv_to(legal[i]) & 1)) return true;
    return false;
}

// Value of a gate hit for the side to move of `b`, or 2 when no gate fires.  Order of
// src/mcts/tactical_mcts.rs:619-708: the hill (opponent on it: lost; reachable: won), then the mate.
CB_HD int light_gates_hd(const cb_board& b, const Move* legal, int n, bool koth) {
    if (koth) {
        if (b.pieces[(b.w_to_move ? 1 : 0) * 6 + KING] & kKothCenter) return -1;
        if (koth_in_1_hd(b, legal, n)) return 1;
    }
    return mate_in_1_hd(b, legal, n) ? 1 : 2;
}

// ---------------------------------------------------------------------------------------------------------------
// The Tier-1 gates at full depth (host form; the tree kernels' warp form is tier1_gates_warp in mcts_device.cuh).
//   mate_search_plies   src/search/mate_search.rs:64-108: iterative deepening over mate-in-d, d = 1..max_depth; levels up to
//                       exhaustive_depth try every legal attacker move, the deeper ones only moves that give check;
//                       solve_mate (:114-289) is a pure any / all minimax, so the answer does not depend on move order.
//                       Returns the mate length in plies (odd), 0 when there is none.  The rule of :276-283 is kept as
//                       written: outside checks-only mode a node whose side to move has no legal move and stands in
//                       check counts as "mate found" on attacker plies too.
//   koth_center_in_n_host  src/search/koth.rs:43-65,270-395: own moves needed to put the king on the hill against every
//                       defence (0..max_n), -1 when there is no forced way; the root side's king must stay inside the
//                       ring from which the centre is still in reach (computed from max_n, as written).
//   tier1_gates_host    the leaf gate, src/mcts/tactical_mcts.rs:619-708: -1 / +1 for the side to move, 2 = no gate;
//                       *dist = terminal_distance: the hill's n, or the mate SCORE 1_000_000 + plies cast to u8 (:701).
// Self-play runs the leaf gate with mate depth 5, exhaustive depth 0, hill depth 3 (src/bin/self_play.rs:213, the
// defaults of src/mcts/tactical_mcts.rs:115,124) and adjudicates a game whose root has a hill in 3 or
// mate_search(board, 5, exhaustive 2) (src/bin/self_play.rs:241-283).
constexpr int kGateMateScore = 1000000;
CB_HD int mate_dist_u8(int plies) { return (kGateMateScore + plies) & 0xFF; }

inline bool solve_mate_host(const cb_board& b, int depth, bool attacker, bool checks_only) {
    const int us = b.w_to_move ? 0 : 1;
    if (depth <= 0) return in_check_hd(b, us) && !has_legal_move_hd(b);
    Move mv[kMaxMoves];
    const int n = gen_pseudo(b, mv);
    bool any = false;
    for (int i = 0; i < n; ++i) {
        cb_board nb;
        make_move_hd(b, mv[i], &nb);
        if (in_check_hd(nb, us)) continue;
        if (attacker && (checks_only || depth == 1) && !in_check_hd(nb, 1 - us)) {   // at the last ply only a check can mate
            any = any || !checks_only;
            continue;
        }
        any = true;
        const bool r = solve_mate_host(nb, depth - 1, !attacker, checks_only);
        if (attacker ? r : !r) return attacker;
    }
    if (!any) return (attacker && checks_only) ? false : in_check_hd(b, us);
    return !attacker;
}

inline int mate_search_plies(const cb_board& b, int max_depth, int exhaustive_depth) {
    const int exhaustive_plies = exhaustive_depth * 2 - 1;
    for (int d = 1; d <= max_depth; ++d) {
        const int depth = 2 * d - 1;
        // the root is an attacker node that reports its depth; no legal move at the root = no mate (the loop finds none)
        const int us = b.w_to_move ? 0 : 1;
        const bool checks_only = depth > exhaustive_plies;
        Move mv[kMaxMoves];
        const int n = gen_pseudo(b, mv);
        for (int i = 0; i < n; ++i) {
            cb_board nb;
            make_move_hd(b, mv[i], &nb);
            if (in_check_hd(nb, us)) continue;
            if ((checks_only || depth == 1) && !in_check_hd(nb, 1 - us)) continue;
            if (solve_mate_host(nb, depth - 1, false, checks_only)) return depth;
        }
    }
    return 0;
}

CB_HD uint64_t koth_ring_mask(int remaining) {   // src/search/koth.rs:22-37
    // cumulative: the centre, the 4 x 4 block around it, the 6 x 6 block, the board
    return remaining <= 0 ? kKothCenter : remaining == 1 ? 0x00003C3C3C3C0000ull : remaining == 2 ? 0x007E7E7E7E7E7E00ull : ~0ull;
}

inline bool solve_koth_host(const cb_board& b, int ply, int max_ply, int max_n) {
    const bool root_turn = (ply & 1) == 0;
    const int stm = b.w_to_move ? 0 : 1, root = root_turn ? stm : 1 - stm;
    if (b.pieces[root * 6 + KING] & kKothCenter) return true;
    if (b.pieces[(1 - root) * 6 + KING] & kKothCenter) return false;
    if (ply > max_ply) return false;
    Move mv[kMaxMoves];
    cb_board nb;
    if (root_turn) {
        const uint64_t king = b.pieces[root * 6 + KING];
        const uint64_t target = koth_ring_mask(max_n - ply / 2 - 1);
        if (!(king & target)) {   // the king has to step inwards now
            const int ks = lsb(king);
            for (uint64_t t = tables().king[ks] & target & ~b.occ[root]; t; t &= t - 1) {
                make_move_hd(b, make_mv(ks, lsb(t), 0), &nb);
                if (!in_check_hd(nb, stm) && solve_koth_host(nb, ply + 1, max_ply, max_n)) return true;
            }
            return false;
        }
        const int n = gen_pseudo(b, mv);
        for (int i = 0; i < n; ++i) {
            make_move_hd(b, mv[i], &nb);
            if (in_check_hd(nb, stm) || !(nb.pieces[root * 6 + KING] & target)) continue;
            if (solve_koth_host(nb, ply + 1, max_ply, max_n)) return true;
        }
        return false;
    }
    const int n = gen_pseudo(b, mv);
    bool legal = false;
    for (int i = 0; i < n; ++i) {
        make_move_hd(b, mv[i], &nb);
        if (in_check_hd(nb, stm)) continue;
        legal = true;
        if (!solve_koth_host(nb, ply + 1, max_ply, max_n)) return false;
    }
    return legal;
}

inline int koth_center_in_n_host(const cb_board& b, int max_n) {
    const int us = b.w_to_move ? 0 : 1;
    const uint64_t k = b.pieces[us * 6 + KING];
    if (!k) return -1;
    if (k & kKothCenter) return 0;
    for (int n = 1; n <= max_n; ++n)
        if (solve_koth_host(b, 0, 2 * n - 1, max_n)) return n;
    return -1;
}

inline int tier1_gates_host(const cb_board& b, bool koth, int mate_depth, int exhaustive_depth, int koth_depth, int* dist = nullptr) {
    int d = 0, v = 2;
    if (koth) {
        if (b.pieces[(b.w_to_move ? 1 : 0) * 6 + KING] & kKothCenter) v = -1;
        else {
            const int n = koth_center_in_n_host(b, koth_depth);
            if (n >= 0) { v = 1; d = n; }
        }
    }
    if (v == 2 && mate_depth > 0) {
        const int plies = mate_search_plies(b, mate_depth, exhaustive_depth);
        if (plies) { v = 1; d = mate_dist_u8(plies); }
    }
    if (dist) *dist = d;
    return v;
}

}  // namespace cbchess
