// chess.hpp — minimal bitboard chess rules for the host side of the leaf-evaluation path:
// legal move lists (the "legal-move mask" of SURVEY.md §8 a16), make-move with the side effects
// that reach the network input (src/make_move.rs:99-185), terminal detection for the self-play
// driver, and the policy index of a move (src/tensor.rs:82-178).  Positions are cb_board records.
//
// Product code: it does not use oracle/.  tests/ cross-check it against the oracle's independent
// generator and the reference's perft table through the cb_chess_* C ABI.
#pragma once
#include <stdint.h>

#include "../../../include/caissa_b200.h"

namespace cbchess {

#if defined(__CUDACC__)
#define CB_HDI __host__ __device__ inline
#else
#define CB_HDI inline
#endif

typedef uint16_t Move;  // from | to << 6 | promo << 12   (promo: 0 none, 1 N, 2 B, 3 R, 4 Q)
CB_HDI int mv_from(Move m) { return m & 63; }
CB_HDI int mv_to(Move m) { return (m >> 6) & 63; }
CB_HDI int mv_promo(Move m) { return (m >> 12) & 7; }
CB_HDI Move make_mv(int f, int t, int p) { return (Move)(f | (t << 6) | (p << 12)); }

enum { PAWN = 0, KNIGHT, BISHOP, ROOK, QUEEN, KING };
enum { WK = 1, WQ = 2, BK = 4, BQ = 8 };
constexpr int kMaxMoves = CB_MAX_MOVES;
constexpr uint64_t kKothCenter = 0x0000001818000000ull;  // d4 e4 d5 e5 (src/board.rs:27)

void startpos(cb_board* b);
bool in_check(const cb_board& b, int color);
int gen_legal(const cb_board& b, Move* out, int* n_tactical = nullptr);   // the reference's generation order; see chess_rules.hpp
void make_move(const cb_board& b, Move m, cb_board* out);
uint64_t perft(const cb_board& b, int depth);
int policy_index(Move m, bool white_to_move);           // move_to_index(flip-if-black(m)); -1 if malformed
uint64_t position_key(const cb_board& b);               // for repetition detection

}  // namespace cbchess
