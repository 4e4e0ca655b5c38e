/* caissa_b200.h — C ABI of libcaissa_b200.so, the B200-native batched MCTS
 * leaf evaluator for Caissawary (aaholmes/neurosymbolic-mcts).
 *
 * This is the drop-in boundary for the reference's evaluator: everything the
 * Rust `NeuralNetPolicy` (src/neural_net.rs:12-363) does through tch/libtorch
 * — encode boards, run OracleNet, return (policy, v_logit, k) — is reachable
 * through the functions below with plain pointers and sizes.  A thin Rust
 * `-sys` crate binds them 1:1; the C++ mirror in
 * neurosymbolic-mcts_b200/csrc/host and the ctypes binding in
 * neurosymbolic-mcts_b200/caissa_b200 call the same entry points.
 *
 * Conventions
 *   - Return value: 0 on success, a negative CB_E* code on failure.  The
 *     library never aborts and never throws; cb_last_error() gives the text
 *     (the reference answers errors with `None`, src/neural_net.rs:295-313).
 *   - Ownership: the caller owns every host buffer passed in; the library owns
 *     device memory, pinned staging and its CUDA stream.
 *   - Threading: one handle = one owner thread (mirrors `&mut self`, the handle
 *     is moved into the InferenceServer worker, src/mcts/inference_server.rs:
 *     29-32).  Distinct handles are independent (own stream, own buffers), so
 *     two models can coexist in a process (src/bin/evaluate_models.rs:558-584).
 *   - There is NO CPU fallback: without a CUDA device cb_create fails.
 */
#ifndef CAISSA_B200_H
#define CAISSA_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CB_POLICY_SIZE 4672 /* 8*8*73, src/tensor.rs:82-178 */
#define CB_PLANES 1088      /* 17*8*8, src/tensor.rs:21-77 */
#define CB_MAX_MOVES 256

enum {
    CB_OK = 0,
    CB_EINVAL = -1,   /* bad argument */
    CB_ECUDA = -2,    /* CUDA runtime / driver error, see cb_last_error */
    CB_ENODEV = -3,   /* no usable CUDA device (no CPU fallback exists) */
    CB_ESTATE = -4,   /* weights not loaded / nothing staged */
    CB_ENOMEM = -5
};

/* Arithmetic mode of the convolution tower. */
enum {
    CB_DTYPE_FP32 = 0, /* the mode of the fp32 tolerance (|dv| <= 1e-3): 128-channel nets run the resident forward on
                          tcgen05 with fp16-PAIR operands (x = hi + lo, three MMAs per product, fp32 accumulate; measured
                          1.2e-5); other nets, and every net under CB_FP32_EXACT=1, run fp32 FMA on CUDA cores (4e-7) */
    CB_DTYPE_BF16 = 1, /* tcgen05 kind::f16 with bf16 operands, fp32 accumulate in TMEM */
    CB_DTYPE_FP16 = 2  /* tcgen05 kind::f16 with fp16 operands, fp32 accumulate in TMEM */
};

/* 128-byte packed position; replaces `Board` (src/board.rs:14-25) at the FFI.
 * Same field order as the reference's own C twin (cuda/common.cuh:54-63).
 * Square = rank*8 + file, a1 = 0.  pieces[color*6 + piece], WHITE=0, BLACK=1,
 * P N B R Q K = 0..5. */
typedef struct cb_board {
    uint64_t pieces[12];
    uint64_t occ[2];
    uint8_t  w_to_move;  /* 1 = white to move */
    uint8_t  en_passant; /* target square, 0xFF = none (Option<u8>::None) */
    uint8_t  castling;   /* WK=1 WQ=2 BK=4 BQ=8 (src/move_types.rs:46-51) */
    uint8_t  halfmove;
    uint8_t  pad[12];
} cb_board;

typedef struct cb_handle cb_handle;

/* Per-phase timing, the counterpart of the reference's three public
 * TimingAccumulators (src/neural_net.rs:88-90,220,245,271). */
typedef struct cb_timing {
    uint64_t calls;
    uint64_t positions;
    double   h2d_ms;     /* transfer_to_gpu_timing   */
    double   forward_ms; /* forward_timing (device time, CUDA events) */
    double   d2h_ms;     /* transfer_from_gpu_timing */
} cb_timing;

/* Replaces NeuralNetPolicy::new (src/neural_net.rs:94-108).  `device` is the
 * CUDA ordinal; (num_blocks, hidden) select the OracleNet shape
 * (python/model.py:55: 6x128 default, 10x256 scaled; hidden in {128, 256}). */
int cb_create(cb_handle** out, int device, int num_blocks, int hidden, int dtype);

/* Replaces NeuralNetPolicy::load (src/neural_net.rs:110-125).  `blob` is the
 * flat f32 parameter vector in the reference's own export order
 * (python/export_weights_cuda.py:73-107) generalised to (num_blocks, hidden);
 * cb_weight_count gives the expected length.  BatchNorm is folded and the
 * operands are packed for the tensor cores here.  Caller keeps `blob`. */
int    cb_load_weights(cb_handle* h, const float* blob, size_t n_floats);
size_t cb_weight_count(int num_blocks, int hidden);

/* NeuralNetPolicy::is_available (src/neural_net.rs:131-133). */
int cb_is_available(const cb_handle* h);

/* Replaces NeuralNetPolicy::predict_batch (src/neural_net.rs:208-315); B=1 is
 * ::predict (:144-205).  Compat mode: policy_probs[B*4672] are PROBABILITIES
 * (exp already applied, :279-284), v_logit[B] the raw value logit, k[B] the
 * position-independent confidence scalar.  qflag (qsearch_completed) is
 * accepted and ignored exactly like the network does (python/model.py:106).
 * policy_probs may be NULL to skip the 18.7 KB/position read-back. */
int cb_predict_batch(cb_handle* h, const cb_board* boards, const float* q, const uint8_t* qflag, int B,
                     float* policy_probs, float* v_logit, float* k);

/* Fused fast path: predict_batch + policy_to_move_priors (src/tensor.rs:184-213)
 * + the value fusion of evaluate_leaf_node (src/mcts/tactical_mcts.rs:762-765
 * material_on: tanh(v_logit + k*q); :787-790 material off: tanh(v_logit)).
 * move_idx/move_off: CSR (B+1 offsets) of policy indices of each position's
 * legal moves, i.e. move_to_index(flip-if-black(mv)) (src/tensor.rs:82-178).
 * priors[move_off[B]] are renormalised over each list in list order with the
 * uniform fallback when the sum is 0.  v_final is fp32 device tanhf; hosts
 * that want the reference's f64 tanh recompute it from v_logit and k. */
int cb_eval_leaves(cb_handle* h, const cb_board* boards, const float* q, int B, int material_on,
                   const uint16_t* move_idx, const uint32_t* move_off,
                   float* priors, float* v_logit, float* v_final, float* k);

/* The same call in two halves, so that ONE host thread can keep several handles busy (the reference's multi-instance use,
 * SURVEY.md §3.4): submit copies the request into the handle's pinned staging block (the caller's buffers are free again
 * on return), enqueues the host->device copy, the forward and the device->host copy on the handle's stream and returns
 * without synchronising; wait blocks until that batch is done and copies the results out.  One batch in flight per handle;
 * cb_eval_leaves == submit + wait, bit for bit.  With two handles: submit(a), submit(b), wait(a), submit(a), wait(b), ... —
 * b's copies overlap a's kernel. */
int cb_eval_leaves_submit(cb_handle* h, const cb_board* boards, const float* q, int B, int material_on,
                          const uint16_t* move_idx, const uint32_t* move_off);
int cb_eval_leaves_wait(cb_handle* h, float* priors, float* v_logit, float* v_final, float* k);

/* Bit-exactness hook for board_to_planes (src/tensor.rs:21-77): planes[B*1088]
 * f32, NCHW, exactly the reference's Vec<f32>. */
int cb_encode(cb_handle* h, const cb_board* boards, int B, float* planes);

/* Static PeSTO stand-pat term of dM on the device, PestoEval::pst_eval_cp
 * (src/eval.rs:93-122): centipawns, side-to-move perspective, bit-exact. */
int cb_pst_eval_cp(cb_handle* h, const cb_board* boards, int B, int32_t* cp);

/* Device-resident benchmarking entry points (inputs already in HBM when the
 * timed region starts).  cb_stage_leaves copies one batch to the device;
 * cb_time_resident replays the captured forward graph `iters` times on the
 * handle's stream and returns the CUDA-event time of each iteration in
 * ms_each[iters] (events recorded on the launching stream).  If flush_l2 is
 * non-zero a 256 MiB buffer is rewritten before every iteration, outside the
 * timed events.  conv_ms_total (nullable) receives the event time spent in the
 * 3x3 convolution kernels alone over all iterations (separate un-graphed
 * replay), n_conv_launches (nullable) their number per iteration. */
int cb_stage_leaves(cb_handle* h, const cb_board* boards, const float* q, int B,
                    const uint16_t* move_idx, const uint32_t* move_off);
int cb_time_resident(cb_handle* h, int material_on, int iters, int flush_l2, float* ms_each);
/* The staged forward replayed back to back (no flush, no pause) for at least `seconds`: the SUSTAINED rate, i.e. under the
 * board's power cap.  Device time of all launches (CUDA events around chunks of 256) and their number. */
int cb_time_sustained(cb_handle* h, int material_on, double seconds, float* ms_total, int* launches);
int cb_time_conv_layers(cb_handle* h, int iters, float* conv_ms_total, int* n_conv_launches);
/* per_layer_ms[2*blocks+2] (nullable): mean event time of every conv launch, start conv first. */
int cb_time_conv_layers_detail(cb_handle* h, int iters, float* conv_ms_total, int* n_conv_launches,
                               float* per_layer_ms);
/* Event time of the convolution tower alone (the fused persistent kernel in the 16-bit modes,
 * the per-layer launches otherwise), one entry per iteration; n_launches = kernels per iteration. */
int cb_time_tower(cb_handle* h, int iters, float* ms_each, int* n_launches);
int cb_read_resident(cb_handle* h, float* priors, float* v_logit, float* v_final, float* k);

/* Debug / parity hook: copy the backbone activation (after the last residual
 * block) of the most recent forward as f32 NCHW [B*hidden*64]. */
int cb_debug_backbone(cb_handle* h, int B, float* out);
/* The device tree kernels' legal-move generator (one warp per board) on n boards: moves[i*256 ..] in generation
 * order and counts[i] — must equal cb_chess_legal_moves move for move (tests). */
int cb_debug_device_legal_moves(cb_handle* h, const cb_board* boards, int n, uint16_t* moves, int32_t* counts);
/* The device tree kernels' light Tier-1 gates (warp form) on n boards — must equal cb_light_gates (tests). */
int cb_debug_device_light_gates(cb_handle* h, const cb_board* boards, int n, int koth, int8_t* out);
/* The device tree kernels' gate at full depth (warp form, one warp per board, resumed in slices of `budget` positions as
 * the self-play steps do; 0 = one slice) — must equal cb_tier1_gates (tests). */
int cb_debug_device_tier1_gates(cb_handle* h, const cb_board* boards, int n, int koth, int mate_depth, int exhaustive_depth,
                                int koth_depth, int budget, int8_t* out, uint8_t* dist);
/* Same for activation ring buffer `buf` (0..2); with cb_debug_set_layer_limit(n) the tower stops
 * after n convolution launches so individual layers can be compared across arithmetic modes. */
int cb_debug_act(cb_handle* h, int B, int buf, float* out);
int cb_debug_set_layer_limit(cb_handle* h, int n_conv_launches);
/* Debug timeline of CTA 0 of the fused 128-channel tower for the staged batch: out[item*8 + slot] SM clock
 * stamps (slots documented at trace_stamp in tower_t_umma.cuh); returns the number of items written. */
int cb_debug_tower_trace(cb_handle* h, long long* out, int max_items);
/* Debug: global-timer stamps (ns) of every CTA of the resident tower for the staged batch:
 * out[cta*4 + {0 kernel entry, 1 first MMA, 2 last epilogue done, 3 exit}]; returns the CTA count. */
int cb_debug_cta_times(cb_handle* h, unsigned long long* out, int max_ctas);

/* Kernels launched per forward of the given kind (for gpu_launches accounting). */
int cb_launches_per_forward(const cb_handle* h);

/* ---- host-side chess rules used by the self-play driver (legal-move lists = the path's
 * "legal-move mask", src/mcts/selection.rs:82-104; perft-pinned, tests/perft_tests.rs). */
int      cb_chess_startpos(cb_board* out);
int      cb_chess_legal_moves(const cb_board* b, uint16_t* moves /*256: from|to<<6|promo<<12*/,
                              uint16_t* policy_idx /*256, nullable: move_to_index(flip-if-black)*/);
int      cb_chess_make_move(const cb_board* b, uint16_t move, cb_board* out);
uint64_t cb_chess_perft(const cb_board* b, int depth);

/* ---- self-play driver (src/bin/self_play.rs:29-624 semantics for the sims/s metric): a pool of
 * concurrent games, one leaf in flight per game per step, PUCT selection
 * Q + c*P*sqrt(max(N,1))/(1+n) (src/mcts/selection.rs:254-273) with the force-every-root-child
 * rule (:65-73), sign-flipping backup (src/mcts/node.rs:403-424), mate/stalemate terminals,
 * proportional-or-greedy move choice (src/bin/self_play.rs:343-351,628-650), subtree reuse, and
 * game end on mate/stalemate/3-fold/50-move/max plies (+ KOTH centre, :365-381).  Leaves are
 * evaluated in ONE cb_eval_leaves batch per step.  dM (material_on) is the static PeSTO stand pat, or with
 * qsearch = 1 the reference's principal-exchange line on top of it (`--qsearch pe`,
 * src/mcts/tactical_mcts.rs:715-731); tier1_light = 1 adds the Tier-1 gate at depth 1, tier1_light = 2 the gate at the depths
 * the reference's self-play runs (mate search to mate-in-5, hill in 3; on the device as resumable jobs).  qsearch = 2 computes dM
 * by the reference's cap1 search (on the device as resumable jobs too); its widest q-search (`extended`) is not built. */
typedef struct cb_selfplay_config {
    int games;            /* concurrent games in the pool */
    int sims_per_move;    /* 200 / 800 */
    int max_plies;        /* 200 */
    int material_on;      /* 1: V = tanh(v + k*q), q chosen by `qsearch`; 0: vanilla (config 5) */
    int koth;             /* King-of-the-Hill game-end rule */
    int threads;          /* host threads for tree operations (0 = hardware concurrency) */
    double c_puct;        /* 1.414 (exploration_constant, f64 like the reference; 0 = that default) */
    double explore_base;  /* 0.80 (0 = that default) */
    uint64_t seed;
    long long max_sims;   /* stop after this many simulations in total (0 = unlimited) */
    long long max_games;  /* stop after this many finished games (0 = unlimited) */
    double max_seconds;   /* wall-clock bound (0 = unlimited) */
    const char* sample_path; /* nullable: append finished games as 5763-f32 records (src/training_data.rs:24-60) */
    int pipeline;         /* 0/1: lock step (select, evaluate, expand); 2: two half-pools, the tree work of one
                             half runs on the host while the other half's leaves are on the GPU; 3: tree
                             operations on the device (cb_selfplay_device); 4: the same with the games split in
                             two pools that alternate, one pool's tree kernel running beside the other pool's
                             forward kernel (needs an even number of games; same games, same trees as 3) */
    cb_handle* aux_handle; /* nullable, pipeline 2 only: a second evaluator with the same weights (own stream and
                             buffers) that serves the second half-pool, so one half's copies and synchronisation
                             overlap the other half's kernel (the reference's multi-instance use) */
    int qsearch;          /* dM when material_on: 0 static PeSTO/100 (stand pat), 1 principal exchange
                             (QSearchVariant::PrincipalExchange; leaf input and the sample's material scalar,
                             src/bin/self_play.rs:315-329), 2 QSearchVariant::Cap1 (src/search/quiescence.rs:808-1012: one
                             capture + one check per side + every evasion) */
    int tier1_light;      /* 1: the light Tier-1 gates (the reference's gate at depth 1): a leaf where the side to
                             move mates in one — or, with koth, steps onto the hill / faces a king already on it —
                             is terminal (+1 / -1, src/mcts/tactical_mcts.rs:619-708) and a game whose root hits a
                             gate is adjudicated (src/bin/self_play.rs:241-283).
                             2: the reference's Tier-1 gate as self-play runs it: at the leaves mate_search to
                             mate-in-5 (checking moves only, src/search/mate_search.rs:64-289) and, with koth, the hill
                             in 3 (src/search/koth.rs:43-65); at the root of every move hill in 3 or
                             mate_search(board, 5, exhaustive 2); depths: the tier1_* fields at the end */
    int shuffle_children; /* 1: permute the child order of every expansion with the game's random stream.  The reference sets
                             randomize_move_order for self-play (src/bin/self_play.rs:224), but its shuffle never executes:
                             ensure_node_expanded (src/mcts/selection.rs:82-104) returns when the node already has children,
                             and the search loop only selects below nodes that evaluate_and_expand_node
                             (src/mcts/tactical_mcts.rs:813-836, no shuffle) has expanded.  0 is therefore what the reference
                             runs; 1 is the behaviour its flag describes */
    int mock_eval;        /* 1 (device pipelines only): replace the network by the reference's mock evaluator
                             (InferenceServer::new_mock_biased, src/mcts/inference_server.rs:103-122): uniform priors,
                             v_logit = mock_v_logit, k = mock_k */
    float mock_v_logit, mock_k;
    int tier1_mate_depth;       /* tier1_light = 2: mate_search_depth of the leaf gate; 0 = 5 (src/bin/self_play.rs:213) */
    int tier1_exhaustive_depth; /* ... its exhaustive_mate_depth; 0 as in self-play: every level tries checking moves only
                                   (src/mcts/tactical_mcts.rs:115) */
    int tier1_koth_depth;       /* ... koth_depth; 0 = 3 (src/mcts/tactical_mcts.rs:124) */
} cb_selfplay_config;

typedef struct cb_selfplay_stats {
    long long sims, nn_evals, terminal_hits, moves, games_finished, samples, steps;
    long long white_wins, black_wins, draws;
    double seconds, eval_seconds, tree_seconds;
    double mean_batch;
    long long overflow;   /* device pipelines: leaves left unexpanded because a game's node arena was full (0 in a sound run) */
    long long gate_hits;  /* device pipelines, tier1_light = 2: leaves decided by the gate at full depth */
    long long gate_waits; /* ... game-steps spent on a suspended gate job (the game completes no simulation in such a step) */
} cb_selfplay_stats;

int cb_selfplay_run(cb_handle* h, const cb_selfplay_config* cfg, cb_selfplay_stats* out);

/* ---- search hooks for the parity tests (tests compare them with oracle/mcts.c) ----
 * Evaluator callback of the host hooks: the leaf, its legal moves in child order (packed moves and policy indices) and dM;
 * writes the priors of those moves and the fused value. */
typedef void (*cb_eval_callback)(void* user, const cb_board* b, const uint16_t* moves, const uint16_t* policy_idx, int n, float q,
                                 int material_on, float* priors, float* v_final);
/* One move's search (cfg->sims_per_move simulations, src/mcts/tactical_mcts.rs:981-1075 on a fresh root) from each of n
 * positions with the HOST driver's tree code and the caller's evaluator — runs without a GPU.  Outputs per position i:
 * counts[i] root children, moves / visits / value_sums[i * CB_MAX_MOVES + c], root_visits[i]. */
int cb_debug_host_search(const cb_selfplay_config* cfg, const cb_board* boards, int n, cb_eval_callback eval, void* user,
                         uint16_t* moves, uint32_t* visits, double* value_sums, int32_t* counts, uint32_t* root_visits);
/* One whole game of the host driver (search, move choice, tree reuse, game end: src/bin/self_play.rs:191-445) from `start`
 * with the game's random stream seeded by `seed`; returns the number of samples, written like the training records. */
int cb_debug_host_game(const cb_selfplay_config* cfg, const cb_board* start, uint64_t seed, cb_eval_callback eval, void* user,
                       cb_board* sample_boards, uint16_t* sample_idx, float* sample_prob, int32_t* sample_n, float* sample_material,
                       int max_samples, int* result_white, int* plies);
/* The same one-move search with the DEVICE tree kernels and the handle's network (or cfg->mock_eval) as evaluator. */
int cb_debug_device_search(cb_handle* h, const cb_selfplay_config* cfg, const cb_board* boards, int n,
                           uint16_t* moves, uint32_t* visits, double* value_sums, int32_t* counts, uint32_t* root_visits);
/* The same pool with the tree operations on the device (cb_selfplay_run dispatches here
 * for pipeline = 3): select (src/mcts/selection.rs:65-73,254-273), expand (:82-104), backup
 * (src/mcts/node.rs:403-424), move choice / tree reuse / game end (src/bin/self_play.rs:300-388) run as two
 * kernels around the forward kernel, one warp per game; leaves and priors never cross PCIe.  Same statistics
 * and the same sample file: moves are logged on the device and the finished games written by the host. */
int cb_selfplay_device(cb_handle* h, const cb_selfplay_config* cfg, cb_selfplay_stats* out);

/* Sequential probability ratio test of the model gate (src/mcts/sprt.rs): trinomial GSPRT as in fishtest.
 * cb_sprt_llr returns 1 and writes the log-likelihood ratio, or 0 when it is undefined (fewer than 2 games,
 * zero variance: SprtState::compute_llr == None).  cb_sprt_decision: +1 accept H1 (LLR >= ln((1-beta)/alpha)),
 * -1 accept H0 (LLR <= ln(beta/(1-alpha))), 0 inconclusive (SprtState::check_decision). */
int cb_sprt_llr(long long wins, long long losses, long long draws, double elo0, double elo1, double* llr);
int cb_sprt_decision(long long wins, long long losses, long long draws, double elo0, double elo1, double alpha, double beta);

/* Evaluation match candidate vs current (src/bin/evaluate_models.rs:227-470 one game, :698-800 the SPRT loop),
 * tree operations on the device: `games` concurrent games, the candidate is white in even-numbered games, every
 * move is a fresh search of sims_per_move simulations by the mover's model, a child that mates is played at
 * once, otherwise moves are chosen as in self-play; game ends as in self-play.  With use_sprt the test runs
 * after every finished game (in finish order) and the decision at the trigger stands. */
typedef struct cb_match_config {
    int games;            /* concurrent games */
    int total_games;      /* stop after this many games */
    int sims_per_move;
    int max_plies;        /* 200 */
    int material_on;
    int koth;
    double c_puct;        /* 1.414 (0 = that default) */
    double explore_base;  /* 0.90, DEFAULT_EXPLORE_BASE of src/bin/evaluate_models.rs:85 (0 = that default) */
    uint64_t seed;
    int use_sprt;
    double elo0, elo1, alpha, beta;   /* 0, 10, 0.05, 0.05 */
    double max_seconds;   /* wall-clock bound (0 = unlimited) */
    int qsearch;          /* as cb_selfplay_config.qsearch (src/bin/evaluate_models.rs:343-357) */
    int tier1_light;      /* as cb_selfplay_config.tier1_light (2: the gate at the depths evaluate_models runs, mate-in-5 by
                             checks and hill in 3, src/bin/evaluate_models.rs:251-282); a root that hits a gate is scored as the
                             win it is (the reference's search returns the winning move there, src/mcts/tactical_mcts.rs:330-420) */
} cb_match_config;

typedef struct cb_match_stats {
    long long games, candidate_wins, current_wins, draws, sims;
    int decision;         /* +1 H1, -1 H0, 0 inconclusive */
    int llr_valid;
    double llr;
    double seconds;
} cb_match_stats;

/* Principal-exchange material balance: forced_principal_exchange of
 * src/search/quiescence.rs:756-804 — stand pat, then only the first legal capture of the MVV-LVA ordered list
 * (src/move_generation.rs:426-446) at every ply, window (-100000, 100000), depth 20 — in pawn units (cp / 100 as
 * f32).  cb_principal_exchange runs on the calling host thread (nodes nullable: positions visited per board);
 * cb_principal_exchange_device runs one warp per board on the handle's GPU (host buffers, like cb_eval_leaves).
 * Both are bit-identical. */
int cb_principal_exchange(const cb_board* boards, int n, float* q, uint32_t* nodes);
int cb_principal_exchange_device(cb_handle* h, const cb_board* boards, int n, float* q);
/* Light Tier-1 gates of n boards on the calling host thread: out[i] = +1 (the side to move mates in one, or with
 * koth has its king on / steps onto the hill), -1 (koth: the opponent's king is on the hill), 2 (no gate).
 * mate_search(board, 1) of src/search/mate_search.rs:64-108 and koth_center_in_n(board, 1) of
 * src/search/koth.rs:43-65, in the order of src/mcts/tactical_mcts.rs:619-708. */
int cb_light_gates(const cb_board* boards, int n, int koth, int8_t* out);
/* The Tier-1 leaf gate at full depth on the calling host thread (src/mcts/tactical_mcts.rs:619-708): out[i] as
 * cb_light_gates; dist[i] (nullable) = terminal_distance of a hit (the hill's n; for a mate the score
 * 1_000_000 + plies cast to u8, :701).  Self-play's leaf gate is (mate_depth 5, exhaustive_depth 0, koth_depth 3). */
int cb_tier1_gates(const cb_board* boards, int n, int koth, int mate_depth, int exhaustive_depth, int koth_depth, int8_t* out,
                   uint8_t* dist);
/* forced_cap1_pesto_balance (src/search/quiescence.rs:1000-1012) of n boards on the calling host thread: q[i] in pawn units,
 * completed[i] / nodes[i] / depth[i] (nullable) the other members of the reference's tuple. */
int cb_cap1_qsearch(const cb_board* boards, int n, float* q, uint8_t* completed, uint32_t* nodes, uint8_t* depth);
/* The same on the handle's GPU, one warp per board (the tree kernels' warp form: must equal cb_cap1_qsearch). */
int cb_cap1_qsearch_device(cb_handle* h, const cb_board* boards, int n, float* q, uint8_t* completed);
/* Test hook: the tree kernels' move tests that do not make the move (legal / gives check, csrc/host/chess_rules.hpp) against
 * make-move + attack test on every pseudo-legal move of n boards; returns the number of disagreements. */
long long cb_debug_move_flags(const cb_board* boards, int n, long long* n_moves);
/* mate_search (src/search/mate_search.rs:64-108): plies[i] = length of the forced mate found (odd), 0 = none. */
int cb_mate_search(const cb_board* boards, int n, int max_depth, int exhaustive_depth, int32_t* plies);
/* koth_center_in_n (src/search/koth.rs:43-65): out[i] = own moves to the hill against every defence (0..max_n), -1 = None. */
int cb_koth_center_in_n(const cb_board* boards, int n, int max_n, int8_t* out);

int cb_match_run(cb_handle* candidate, cb_handle* current, const cb_match_config* cfg, cb_match_stats* out);

int         cb_get_timing(const cb_handle* h, cb_timing* out);
void        cb_reset_timing(cb_handle* h);
const char* cb_last_error(const cb_handle* h);
const char* cb_version(void);
void        cb_destroy(cb_handle* h);

#ifdef __cplusplus
}
#endif
#endif
