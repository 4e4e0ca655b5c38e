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

}  // namespace cbchess
