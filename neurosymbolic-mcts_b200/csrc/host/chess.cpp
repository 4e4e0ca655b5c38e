// chess.cpp — see chess.hpp.  The rules themselves are the inline host+device code of chess_rules.hpp.
#include "chess.hpp"

#include <string.h>

#include "chess_rules.hpp"

namespace cbchess {

const Tables& host_tables() {
    static const Tables* t = [] {
        Tables* p = new Tables();
        build_tables(p);
        return p;
    }();
    return *t;
}

void startpos(cb_board* b) {
    memset(b, 0, sizeof(*b));
    const uint64_t back[6] = {0, 0x42, 0x24, 0x81, 0x08, 0x10};
    b->pieces[PAWN] = 0xFF00ull;
    b->pieces[6 + PAWN] = 0xFFull << 48;
    for (int p = 1; p < 6; ++p) {
        b->pieces[p] = back[p];
        b->pieces[6 + p] = back[p] << 56;
    }
    b->occ[0] = 0xFFFFull;
    b->occ[1] = 0xFFFFull << 48;
    b->w_to_move = 1;
    b->en_passant = 0xFF;
    b->castling = WK | WQ | BK | BQ;
}

bool in_check(const cb_board& b, int color) { return in_check_hd(b, color); }
void make_move(const cb_board& b, Move m, cb_board* out) { make_move_hd(b, m, out); }
int gen_legal(const cb_board& b, Move* out, int* n_tactical) { return gen_legal_hd(b, out, n_tactical); }
int policy_index(Move m, bool white_to_move) { return policy_index_hd(m, white_to_move); }
uint64_t position_key(const cb_board& b) { return position_key_hd(b); }

uint64_t perft(const cb_board& b, int depth) {
    if (depth == 0) return 1;
    Move mv[kMaxMoves];
    const int n = gen_legal(b, mv);
    if (depth == 1) return (uint64_t)n;
    uint64_t total = 0;
    for (int i = 0; i < n; ++i) {
        cb_board nb;
        make_move(b, mv[i], &nb);
        total += perft(nb, depth - 1);
    }
    return total;
}

}  // namespace cbchess
