/* oracle.h — CPU restatement of the reference's leaf-evaluation codec and a
 * perft-pinned chess substrate.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load liboracle.so.  The product library
 * (libcaissa_b200.so) never links or dlopens it.
 *
 * Every function cites the reference file:line it restates (paths relative
 * to the reference checkout).  Parity pins: the known-answer literals of
 * tests/unit/board_encoding_tests.rs, tests/unit/tensor_extended_tests.rs,
 * src/tensor.rs:215-246 and the perft table of tests/perft_tests.rs are
 * replayed in tests/test_oracle_*.py.
 */
#ifndef CAISSA_ORACLE_H
#define CAISSA_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 128-byte packed position: the FFI board format (same field layout as the
 * reference's own C twin of Board, cuda/common.cuh:54-63; semantic source
 * src/board.rs:14-25).  Square index = rank*8 + file, a1 = 0. */
typedef struct {
    uint64_t pieces[12];  /* [color*6 + piece]; WHITE=0 BLACK=1; P N B R Q K = 0..5 */
    uint64_t occ[2];
    uint8_t  w_to_move;   /* 1 = white */
    uint8_t  en_passant;  /* target square or 0xFF */
    uint8_t  castling;    /* WK=1 WQ=2 BK=4 BQ=8 */
    uint8_t  halfmove;
    uint8_t  pad[12];
} orc_board;

/* Move packing: from | to<<6 | promo<<12, promo 0 none, 1 N, 2 B, 3 R, 4 Q
 * (the reference's piece constants, src/piece_types.rs:14-23). */
typedef uint16_t orc_move;
#define ORC_MV(f, t, p) ((orc_move)((f) | ((t) << 6) | ((p) << 12)))
#define ORC_FROM(m) ((int)((m) & 63))
#define ORC_TO(m) ((int)(((m) >> 6) & 63))
#define ORC_PROMO(m) ((int)(((m) >> 12) & 7))

#define ORC_POLICY 4672
#define ORC_PLANES 1088
#define ORC_MAX_MOVES 256

/* ---- codec (src/tensor.rs, src/eval.rs, src/mcts/tactical_mcts.rs) ---- */
int    orc_parse_fen(const char* fen, orc_board* out);            /* src/board.rs:115-199 */
void   orc_startpos(orc_board* out);
void   orc_board_to_planes(const orc_board* b, float* planes);    /* src/tensor.rs:21-77 */
int    orc_move_to_index(int from, int to, int promo);            /* src/tensor.rs:82-178; -1 where the reference panics */
int    orc_move_to_index_stm(orc_move m, int w_to_move);          /* + flip for black, src/move_types.rs:374-381 */
void   orc_policy_to_move_priors(const float* policy, int n_policy, const orc_move* moves, int n,
                                 int w_to_move, float* priors);   /* src/tensor.rs:184-213 */
int    orc_pst_eval_cp(const orc_board* b);                       /* src/eval.rs:93-122 */
double orc_value_fuse(float v_logit, float k, float q, int material_on); /* src/mcts/tactical_mcts.rs:762-765,787-790 */

/* ---- chess substrate (position source; pinned by tests/perft_tests.rs) ---- */
int      orc_gen_legal(const orc_board* b, orc_move* out);        /* legal list, captures/promotions first */
void     orc_make_move(const orc_board* b, orc_move m, orc_board* out); /* src/make_move.rs:24-225 rules */
int      orc_in_check(const orc_board* b, int color);
uint64_t orc_perft(const orc_board* b, int depth);
/* Seeded random legal playouts from the start position.  Emits up to n
 * positions (each with >=1 legal move) and their legal-move CSR as policy
 * indices (move_to_index of the STM-relative move).  Returns positions written. */
int      orc_random_positions(uint64_t seed, int n, int max_plies, orc_board* boards,
                              uint16_t* move_idx, uint32_t* move_off, orc_move* moves_raw);

/* ---- principal-exchange quiescence line (qsearch.c; the "pe" dM search) ---- */
/* MVV-LVA ordered captures + promotions, src/move_generation.rs:426-446.  *long_list = 1 when the list is longer
 * than the 20 elements for which the reference's sort is known to keep ties in generation order. */
int   orc_gen_captures_mvvlva(const orc_board* b, orc_move* out, int* long_list);
int   orc_principal_exchange(const orc_board* b, int alpha, int beta, int max_depth, uint32_t* nodes,
                             int* long_list);                       /* src/search/quiescence.rs:756-794; *nodes accumulates */
float orc_forced_principal_exchange(const orc_board* b, uint32_t* nodes, int* long_list); /* :797-804 */
/* cap = 1 quiescence search (qsearch.c; the "cap1" dM search), src/search/quiescence.rs:808-1012 */
int   orc_cap1_qsearch(const orc_board* b, int alpha, int beta, int max_depth, int w_check_used, int b_check_used, int* completed,
                       uint32_t* nodes, int* depth_used, int* long_list);
float orc_forced_cap1(const orc_board* b, int* completed, uint32_t* nodes, int* depth_used, int* long_list);
/* light Tier-1 gates = the reference's gate at depth 1 (qsearch.c) */
int   orc_mate_in_1(const orc_board* b);            /* src/search/mate_search.rs:64-108, max_depth 1 */
int   orc_koth_in_1(const orc_board* b);            /* src/search/koth.rs:43-65, max_n 1 */
int   orc_light_gates(const orc_board* b, int koth); /* src/mcts/tactical_mcts.rs:619-708: -1, +1 or 2 (none) */
/* the Tier-1 gates at full depth (gates.c) */
int   orc_mate_search(const orc_board* b, int max_depth, int exhaustive_depth, orc_move* best, int* nodes);
                                                    /* src/search/mate_search.rs:64-289: 1_000_000 + plies, or 0 */
int   orc_koth_center_in_n(const orc_board* b, int max_n, uint32_t* nodes); /* src/search/koth.rs:146-268: n, or -1 = None */
int   orc_tier1_gates(const orc_board* b, int koth, int mate_depth, int exhaustive_depth, int koth_depth, int* dist);
void  orc_gate_node_stats(long* calls, long* nodes, long* max_nodes, int reset); /* positions entered by orc_tier1_gates so far */
                                                    /* src/mcts/tactical_mcts.rs:619-708: -1, +1 or 2; *dist = terminal_distance */

/* ---- tree operations (mcts.c): generation order, tactical-first selection, f64 PUCT, backup, reuse, game loop ---- */
/* gen_pseudo_legal_moves in the reference's order, src/move_generation.rs:287-328: captures (pawns, knights, bishops,
 * rooks, queens, king) ++ promotions, then quiets; *n_captures = length of the first (capture + promotion) list */
int    orc_gen_pseudo_ref(const orc_board* b, orc_move* out, int* n_captures);
/* legal moves in child order (src/mcts/tactical_mcts.rs:813-836); *n_tactical (nullable) = how many come from the capture list */
int    orc_gen_legal_ref(const orc_board* b, orc_move* out, int* n_tactical);
double orc_tactical_score(const orc_board* b, orc_move m);       /* calculate_mvv_lva, src/mcts/tactical.rs:179-216 */
void   orc_tactical_order(const orc_board* b, const orc_move* moves, int n_tactical, int* order); /* :155-177 */

/* Evaluator callback: priors of `moves` (child order; src/tensor.rs:184-213) and the fused leaf value
 * tanh(v_logit + k*q) / tanh(v_logit) (src/mcts/tactical_mcts.rs:762-765,787-790). */
typedef void (*orc_eval_fn)(void* user, const orc_board* b, const orc_move* moves, int n, float q, int material_on,
                            float* priors, double* value);
typedef struct {
    int sims;             /* max_iterations */
    int material_on;      /* enable_material_value */
    int qsearch;          /* dM: 0 static PeSTO / 100, 1 principal exchange, 2 cap1 */
    int koth;
    int tier1_light;      /* Tier-1 gates: 1 at depth 1; 2 the reference's gate (leaves: mate_depth / exhaustive_mate_depth /
                             koth_depth below; the game loop: hill in 3, mate_search(5, exhaustive 2), self_play.rs:241-283) */
    int shuffle;          /* permute every expansion's child order (the reference's randomize_move_order is dead code in its
                             search loop: selection.rs:82-104 returns before shuffling, the children already exist) */
    int max_plies;        /* game loop: 200 */
    double c_puct;        /* 1.414 */
    double explore_base;  /* 0.80 */
    int mate_depth;            /* tier1_light = 2: mate_search_depth, 0 = 5 (src/bin/self_play.rs:213) */
    int exhaustive_mate_depth; /* 0 as in self-play (src/mcts/tactical_mcts.rs:115) */
    int koth_depth;            /* 0 = 3 (:124) */
} orc_mcts_config;
typedef struct orc_tree orc_tree;
orc_tree* orc_mcts_new(const orc_board* root, uint64_t seed);
void      orc_mcts_free(orc_tree* t);
int       orc_mcts_search(orc_tree* t, const orc_mcts_config* cfg, orc_eval_fn eval, void* user); /* returns iterations run */
int       orc_mcts_root_children(const orc_tree* t, orc_move* moves, uint32_t* visits, double* total_value);
int       orc_mcts_root_child_info(const orc_tree* t, int i, int* has_term, double* term_value, int* term_dist, int* n_children);
void      orc_mcts_root_stats(const orc_tree* t, uint32_t* visits, double* total_value, long* evals, long* terminal_hits, long* tactical);
orc_move  orc_mcts_best_move(const orc_tree* t);                  /* select_best_move_from_root; 0 = None */
int       orc_mcts_advance(orc_tree* t, orc_move played);         /* reuse_subtree */
void      orc_eval_mock(void* user /* const float[2]: v_logit, k */, const orc_board* b, const orc_move* moves, int n, float q,
                        int material_on, float* priors, double* value); /* InferenceServer::new_mock_biased */
typedef struct {
    orc_board board;
    int n;
    uint16_t idx[ORC_MAX_MOVES];
    float prob[ORC_MAX_MOVES];
    float material, value_target;
    int w_to_move;
} orc_game_sample;
/* play_game, src/bin/self_play.rs:191-445; returns the number of samples (moves played) */
int orc_play_game(const orc_board* start, const orc_mcts_config* cfg, uint64_t seed, orc_eval_fn eval, void* user,
                  orc_game_sample* samples, int max_samples, int* result_white, int* plies);

#ifdef __cplusplus
}
#endif
#endif
