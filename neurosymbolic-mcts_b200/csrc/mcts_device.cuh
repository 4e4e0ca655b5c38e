// mcts_device.cuh — the tree operations of the self-play pool ON THE DEVICE: one warp
// per game walks its tree, generates the leaf's legal moves and writes the leaf straight into the
// evaluator's input block; after the forward kernel a second kernel expands the leaf with the priors,
// backs the value up and — when the simulation budget of the move is spent — picks and plays the move,
// re-roots the tree and detects the end of the game.  No host work and no PCIe traffic per simulation.
//
// Semantics are the reference's search, statement for statement (the host driver host/selfplay.cpp is the scalar twin,
// oracle/mcts.c the independent restatement the tests compare both with):
//   child order      generation order, src/move_generation.rs:287-328 (host/chess_rules.hpp)
//   selection        src/mcts/selection.rs:37-63,107-149 (tactical phase: every capture / promotion child once, MVV-LVA
//                    order, in every mode), :65-73 (every root child once), :152-273 (PUCT in f64, sqrt(max(N,1)), Q = 0
//                    unvisited, first maximum wins; proven losses skipped, longest delay when every move loses)
//   priors           a fresh root is expanded without an evaluation and asks for its policy before its first PUCT step
//                    (:281-345): here the root of a new game is evaluated once for its priors only (no backup)
//   expansion        src/mcts/tactical_mcts.rs:813-836; terminal test src/mcts/node.rs:162-177 (made when the child is first
//                    reached instead of at creation: same values, same visit counts)
//   backup           src/mcts/node.rs:403-424 (f64 sums, sign flips every ply)
//   early stop       src/mcts/tactical_mcts.rs:524-560 (with the gates on)
//   move choice      src/bin/self_play.rs:300-356,628-650 (proportional to visits-1 with probability 0.8^(move-1), else the
//                    LAST most-visited child)
//   tree reuse / game end  src/bin/self_play.rs:358-388
// All floating-point steps are explicitly rounded single operations (no FMA contraction), so for identical evaluator
// outputs the device trees are bit-identical to the host driver's and the oracle's (tests/test_gpu_parity.py).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "host/chess_rules.hpp"
#include "host/qsearch.hpp"
#include "host/gates.hpp"

namespace cb {

using cbchess::Move;

struct MNode {
    double value_sum;      // total_value (src/mcts/node.rs:63): from the point of view of the player who moved INTO this node
    uint32_t first_child;
    uint32_t visits;
    float prior;           // move_priorities[move] of the parent (f32 from the evaluator; the reference widens it to f64)
    uint16_t n_children;
    uint16_t move;
    uint8_t state;         // 0 unexpanded, 1 expanded, 2 terminal or resolved by a gate
    int8_t term_value;     // terminal_or_mate_value for the side to move at a state-2 node (-1, 0, +1)
    uint8_t n_tact;        // as a parent: children that come from the capture list (the tactical phase's candidates)
    uint8_t t_done;        // as a parent: how many of them have been visited (tactical_moves_explored)
    uint8_t trank;         // as a child: place in the parent's tactical order (MVV-LVA, stable); 0xFF = not tactical
    uint8_t term_dist;     // terminal_distance of a state-2 node
    uint16_t pad;
};
static_assert(sizeof(MNode) == 32, "node layout");

constexpr uint8_t kMateIn1Dist = (uint8_t)1000001u;   // `mate_result.0.unsigned_abs() as u8` for 1_000_000 + 1 (tactical_mcts.rs:701)

constexpr int kMctsHist = 128;      // position keys since the last irreversible move (halfmove clock <= 100)
constexpr int kMctsMaxDepth = 512;  // longest root-to-leaf path kept
constexpr int kMctsWarps = 4;       // games per CTA

constexpr int kGatePlies = 10;   // mate-in-5: plies 0..8 move, ply 9 is the mated position; the hill in 3 enters ply 6
constexpr size_t kGateJobBytes = (kGatePlies + 1) * sizeof(cb_board) + (size_t)kGatePlies * CB_MAX_MOVES * 2 + 6 * 24 * 4 + 64 + 16;
constexpr size_t kGateJobActiveOffset = kGateJobBytes - 16;

enum { MS_EVALS = 0, MS_TERMINAL, MS_MOVES, MS_GAMES, MS_SAMPLES, MS_WWINS, MS_BWINS, MS_DRAWS, MS_OVERFLOW, MS_SIMS, MS_GATE_HITS, MS_GATE_WAITS,
       MS_COUNT = 16 };

struct MctsParams {
    int G, sims_per_move, max_plies, material_on, koth;
    int qsearch;                  // dM: 0 stand pat, 1 principal exchange (host/qsearch.hpp)
    int tier1_light;              // 1: light Tier-1 gates at the leaves + adjudication at the root (host/gates.hpp); 2: the
                                  // reference's gate at full depth (tier1_gates_run below), run as resumable jobs
    int t1_mate_depth, t1_exh_depth, t1_koth_depth;   // the leaf gate of tier1_light = 2 (5 / 0 / 3 in the reference's self-play)
    int gate_budget;              // positions a game's gate job may enter per step before it is suspended
    int gate_stretch;             // ... and how many more it may enter while fewer than gate_target games of the launch are done
    unsigned int gate_target;     //     with their step (mcts_gate_kernel only): when most games are inside long searches, a step
    unsigned int* gate_ready;     //     that stops after `gate_budget` positions would run the forward for a handful of leaves
    struct GateJob* gate_jobs;    // [G][gate_warps] (tier1_light = 2, dM by cap1): a game's jobs; game-level state in its first
    int gate_warps;               // warps per game in mcts_gate_kernel: each takes every gate_warps-th move of the gate's root
    int g0, gn;                   // the games [g0, g0 + gn) this launch works on (one pool of a two-pool run)
    unsigned int* pdl_done;       // nullable.  Set when the launch is a programmatic dependent of the forward kernel of the
                                  // OTHER pool: the last warp to finish (wrapping counter) waits for that kernel before it
                                  // exits, so that stream order after this launch still implies the forward's completion —
                                  // the other warps leave at once and free their SM for the blocks still queued
    double c_puct;                // 1.414 (f64 like the reference's exploration_constant)
    double explore_base;
    int shuffle;                  // permute every expansion's child order (see cb_selfplay_config.shuffle_children)
    int search_only;              // debug search (cb_debug_device_search): a game stops once its budget is spent, no move is played
    int mock_eval;                // the evaluator is mcts_mock_eval_kernel (InferenceServer::new_mock_biased), not the network
    uint32_t node_cap;              // nodes per arena
    // per-game state
    cb_board* root;                 // [G]
    MNode* nodes;                   // [G][2 arenas][node_cap]
    uint32_t* n_nodes;              // [G]
    uint8_t* arena;                 // [G] which arena holds the tree
    uint64_t* history;              // [G][kMctsHist]
    int* hist_n;                    // [G]
    uint64_t* rng;                  // [G]
    int* ply;                       // [G]
    int* sims_done;                 // [G]
    int* samples_in_game;           // [G]
    // per-step scratch
    uint32_t* path;                 // [G][kMctsMaxDepth]
    int* path_n;                    // [G]
    uint8_t* wants_eval;            // [G]
    uint16_t* leaf_moves;           // [G][CB_MAX_MOVES]
    uint8_t* leaf_trank;            // [G][CB_MAX_MOVES] tactical rank of every leaf move (0xFF: not tactical)
    uint8_t* leaf_ntact;            // [G]
    // evaluator input / output block (slot = game)
    cb_board* eval_boards;          // [G]
    float* eval_q;                  // [G]
    uint16_t* eval_idx;             // [G][CB_MAX_MOVES] policy indices of the leaf's legal moves
    uint16_t* eval_cnt;             // [G]
    const float* priors;            // [G][CB_MAX_MOVES]
    const float* v_final;           // [G]
    unsigned long long* stats;      // [MS_COUNT]
    uint64_t seed;
    const cb_board* start_boards;   // nullable: the position a game in slot g starts from (debug search); else the start position
    // optional training-sample log (nullptr: off): ring of kSampleSlotBytes records appended in program order
    // per game — one per move played (board, material, visit counts of the root's children) and one per
    // finished game (the result); the host drains it between graph replays
    uint8_t* sample_ring;
    uint32_t ring_slots;
    unsigned long long* ring_cursor;
    uint32_t* game_serial;          // [G] games finished so far in this slot
    // model-vs-model evaluation matches (src/bin/evaluate_models.rs:227-470,698-800): every move is a fresh
    // search by the side to move's model; both models evaluate every step's leaves, the expansion takes the
    // mover's outputs.  match_mode = 0: self-play.
    int match_mode;
    const float* priors_b;          // outputs of the second ("current") model, same layout as priors / v_final
    const float* v_final_b;
    uint8_t* cand_white;            // [G] the candidate plays white in the slot's current game (even game index)
    uint8_t* idle;                  // [G] the slot has no game left to play
    long long total_games;
    unsigned long long* games_started;
    int* result_log;                // finished games in finish order: +1 candidate won, -1 current won, 0 draw
    uint32_t result_cap;
    unsigned long long* result_cursor;
};

// Sample record: u32 game | u32 serial | u16 ply (bit 15: the material search did not complete) | u16 n (0xFFFF = game end) | f32 material or i32 result |
// cb_board (128 B) | n x (u16 policy index, u16 visits)
constexpr int kSampleSlotBytes = 1024;
static_assert(16 + 128 + 4 * 218 <= kSampleSlotBytes, "sample slot");

__device__ __forceinline__ uint64_t rng_next(uint64_t& s) {
    s ^= s >> 12; s ^= s << 25; s ^= s >> 27;
    return s * 0x2545F4914F6CDD1Dull;
}
__device__ __forceinline__ double rng_uniform(uint64_t& s) { return (double)(rng_next(s) >> 11) * (1.0 / 9007199254740992.0); }

// dM of the position in `b` (shared memory; overwritten when the exchange line runs) in pawn units.  All lanes.
// `always`: the material scalar of a training sample is computed whatever material_on says (src/bin/self_play.rs:315-329).
__device__ __forceinline__ float mcts_material_q(const MctsParams& p, cb_board& b, cb_board& scratch, uint8_t* mailbox, int lane,
                                                 bool always = false) {
    if (!p.material_on && !always) return 0.0f;
    const int cp = p.qsearch == 1 ? cbchess::principal_exchange_cp_warp(b, scratch, mailbox, lane) : cbchess::pesto_cp_warp(b, mailbox, lane);
    return __fdiv_rn((float)cp, 100.0f);
}

// Start a fresh game in slot g (lane 0 only).
__device__ __forceinline__ void mcts_reset_game(const MctsParams& p, int g, uint64_t seed) {
    cb_board b;
    memset(&b, 0, sizeof(b));
    if (p.start_boards) b = p.start_boards[g];
    else {
    const uint64_t back[6] = {0, 0x42, 0x24, 0x81, 0x08, 0x10};
    b.pieces[cbchess::PAWN] = 0xFF00ull;
    b.pieces[6 + cbchess::PAWN] = 0xFFull << 48;
    for (int q = 1; q < 6; ++q) {
        b.pieces[q] = back[q];
        b.pieces[6 + q] = back[q] << 56;
    }
    b.occ[0] = 0xFFFFull;
    b.occ[1] = 0xFFFFull << 48;
    b.w_to_move = 1;
    b.en_passant = 0xFF;
    b.castling = cbchess::WK | cbchess::WQ | cbchess::BK | cbchess::BQ;
    }
    p.root[g] = b;
    p.arena[g] = 0;
    MNode z;
    memset(&z, 0, sizeof(z));
    p.nodes[(size_t)g * 2 * p.node_cap] = z;
    p.n_nodes[g] = 1;
    p.history[(size_t)g * kMctsHist] = cbchess::position_key_hd(b);
    p.hist_n[g] = 1;
    p.rng[g] = seed | 1;
    p.ply[g] = 0;
    p.sims_done[g] = 0;
    p.samples_in_game[g] = 0;
    if (p.gate_jobs) reinterpret_cast<uint8_t*>(p.gate_jobs)[(size_t)g * p.gate_warps * kGateJobBytes + kGateJobActiveOffset] = 0;   // GateJob::active
}

__global__ void mcts_init_kernel(const MctsParams p) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= p.G) return;
    mcts_reset_game(p, g, p.seed * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull * (uint64_t)(g + 1));
    p.wants_eval[g] = 0;
    p.eval_cnt[g] = 0;
    p.path_n[g] = 0;
    if (p.match_mode) {
        p.cand_white[g] = (g & 1) == 0;         // game index g: the candidate is white in even games
        p.idle[g] = g >= p.total_games;
    }
}

// Value backed up along the stored path, sign flipping every ply (src/mcts/node.rs:403-424).  The updates of
// the path's nodes are independent, so the lanes take one node each (their memory round trips overlap); the
// arithmetic per node is the reference's: total_value += +-value (f64), visits += 1.  All lanes call this.
__device__ __forceinline__ void mcts_backup(MNode* nodes, const uint32_t* path, int n, double val, int lane) {
    for (int i = n - 1 - lane; i >= 0; i -= 32) {
        // node at path index i gets val negated (n - i) times
        const double v = ((n - i) & 1) ? -val : val;
        MNode& nd = nodes[path[i]];
        nd.value_sum = __dadd_rn(nd.value_sum, v);
        nd.visits++;
    }
    __syncwarp();
}

// should_early_terminate (src/mcts/tactical_mcts.rs:524-560; only with the gates on): a root child that is a decided
// win, a visited child with Q > 0.99 once every child is visited, or every child a proven loss.  All lanes; uniform.
__device__ __forceinline__ bool mcts_early_terminate(const MNode* nodes, int lane) {
    const MNode root = nodes[0];
    if (root.state != 1) return false;
    bool win1 = false, forced = false, unvisited = false, not_losing = false;
    for (uint32_t i = lane; i < root.n_children; i += 32) {
        const MNode ch = nodes[root.first_child + i];
        const bool term = ch.state == 2;
        win1 |= term && ch.term_value < 0;
        forced |= ch.visits > 0 && __ddiv_rn(ch.value_sum, (double)ch.visits) > 0.99;
        unvisited |= ch.visits == 0;
        not_losing |= !(term && ch.term_value > 0);
    }
    win1 = __any_sync(0xffffffffu, win1);
    forced = __any_sync(0xffffffffu, forced);
    unvisited = __any_sync(0xffffffffu, unvisited);
    not_losing = __any_sync(0xffffffffu, not_losing);
    if (win1) return true;
    if (forced && !unvisited) return true;
    return !unvisited && !not_losing && root.n_children > 0;
}

// A simulation is complete: count it, and with the gates on test the early stop (the move is then played at once).
__device__ __forceinline__ void mcts_sim_done(const MctsParams& p, const MNode* nodes, int g, int lane) {
    __syncwarp();
    const bool stop = p.tier1_light && mcts_early_terminate(nodes, lane);
    if (lane == 0) {
        p.sims_done[g] = stop ? p.sims_per_move : p.sims_done[g] + 1;
        atomicAdd(&p.stats[MS_SIMS], 1ull);
    }
    __syncwarp();
}

// cbchess::gen_pseudo with one own piece per lane (a side has at most 16): the same list in the same order.
// Lane k takes the k-th piece in generation order (pawns, knights, bishops, rooks, queens, king; ascending squares) and
// counts its captures, its promotions and its quiet moves; three warp scans give every lane its place in the capture
// block, the promotion block and the quiet block.  The castling moves (five squares tested for attack on five lanes)
// open the king's quiet moves.  Returns the list length; *n_caps = length of the capture list (captures + promotions).
__device__ __forceinline__ int gen_pseudo_warp(const cb_board& b, Move* out, int lane, int* n_caps = nullptr) {
    using namespace cbchess;
    const Tables& T = tables();
    const int us = b.w_to_move ? 0 : 1, them = 1 - us;
    const uint64_t opp = b.occ[them], occ = b.occ[0] | b.occ[1];
    const uint64_t* pc = &b.pieces[us * 6];
    const int fwd = us == 0 ? 8 : -8, promo_from_rank = us == 0 ? 6 : 1, start_rank = us == 0 ? 1 : 6;
    // castling first: uniform, needed by the king's lane count
    bool castle_k = false, castle_q = false;
    const int ksq = us == 0 ? 4 : 60, ks = us == 0 ? WK : BK, qs = us == 0 ? WQ : BQ;
    if ((pc[KING] >> ksq & 1) && (b.castling & (ks | qs))) {          // uniform
        const bool att = lane < 5 && attacked(b, ksq - 2 + lane, them);
        const uint32_t am = __ballot_sync(0xffffffffu, att);           // bit i: square ksq - 2 + i is attacked
        if (!(am >> 2 & 1)) {
            castle_k = (b.castling & ks) && (pc[ROOK] >> (ksq + 3) & 1) && !(occ & (3ull << (ksq + 1))) && !(am >> 3 & 3);
            castle_q = (b.castling & qs) && (pc[ROOK] >> (ksq - 4) & 1) && !(occ & (7ull << (ksq - 3))) && !(am & 3);
        }
    }
    int pt = -1, from = 0;
    {
        int k = lane;
#pragma unroll
        for (int t = 0; t < 6; ++t) {
            const int c = __popcll(pc[t]);
            if (pt < 0) {
                if (k < c) { pt = t; from = nth_set_bit(pc[t], k); }
                else k -= c;
            }
        }
    }
    uint64_t caps = 0, quiets = 0, caps2 = 0, quiets2 = 0;   // (queen: rook set, then bishop set)
    int ncap = 0, nprom = 0, nquiet = 0;
    bool promo = false, one_ok = false, two_ok = false;
    if (pt == PAWN) {
        const int one = from + fwd;
        promo = (from >> 3) == promo_from_rank;
        const bool one_empty = !(occ >> one & 1);
        caps = T.pawn[us][from] & (opp | (b.en_passant != 0xFF ? 1ull << b.en_passant : 0ull));
        if (promo) nprom = 4 * (__popcll(caps) + (one_empty ? 1 : 0));
        else {
            ncap = __popcll(caps);
            one_ok = one_empty;
            two_ok = one_ok && (from >> 3) == start_rank && !(occ >> (one + fwd) & 1);
            nquiet = (one_ok ? 1 : 0) + (two_ok ? 1 : 0);
        }
    } else if (pt == QUEEN) {
        const uint64_t ra = rook_att(from, occ), ba = bishop_att(from, occ);
        caps = ra & opp; caps2 = ba & opp;
        quiets = ra & ~occ; quiets2 = ba & ~occ;
        ncap = __popcll(caps) + __popcll(caps2);
        nquiet = __popcll(quiets) + __popcll(quiets2);
    } else if (pt >= 0) {
        const uint64_t att = pt == KNIGHT ? T.knight[from] : pt == KING ? T.king[from] : pt == BISHOP ? bishop_att(from, occ)
                                                                                                      : rook_att(from, occ);
        caps = att & opp;
        quiets = att & ~occ;
        ncap = __popcll(caps);
        nquiet = __popcll(quiets) + (pt == KING ? (castle_k ? 1 : 0) + (castle_q ? 1 : 0) : 0);
    }
    int cap_end = ncap, prom_end = nprom, quiet_end = nquiet;   // inclusive scans over the lanes
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int a = __shfl_up_sync(0xffffffffu, cap_end, d), pr = __shfl_up_sync(0xffffffffu, prom_end, d),
                  q = __shfl_up_sync(0xffffffffu, quiet_end, d);
        if (lane >= d) { cap_end += a; prom_end += pr; quiet_end += q; }
    }
    const int total_caps = __shfl_sync(0xffffffffu, cap_end, 31), total_proms = __shfl_sync(0xffffffffu, prom_end, 31),
              total_quiets = __shfl_sync(0xffffffffu, quiet_end, 31);
    int n = cap_end - ncap;                                        // this lane's first capture slot
    int nq = total_caps + total_proms + quiet_end - nquiet;        // ... first quiet slot
    if (pt == PAWN) {
        if (promo) {
            int np = total_caps + prom_end - nprom;
            const int one = from + fwd;
            for (uint64_t t = caps; t; t &= t - 1) np = add_pawn(out, np, from, lsb(t), one >> 3);
            if (!(occ >> one & 1)) np = add_pawn(out, np, from, one, one >> 3);
        } else {
            for (uint64_t t = caps; t; t &= t - 1) out[n++] = make_mv(from, lsb(t), 0);
            if (two_ok) out[nq++] = make_mv(from, from + 2 * fwd, 0);
            if (one_ok) out[nq++] = make_mv(from, from + fwd, 0);
        }
    } else if (pt == KNIGHT || pt == KING) {
        if (pt == KING) {
            if (castle_k) out[nq++] = make_mv(ksq, ksq + 2, 0);
            if (castle_q) out[nq++] = make_mv(ksq, ksq - 2, 0);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int d = pt == KNIGHT ? knight_delta(i) : king_delta(i);
            if (step_targets_bit(caps, from, d)) out[n++] = make_mv(from, from + d, 0);
            else if (step_targets_bit(quiets, from, d)) out[nq++] = make_mv(from, from + d, 0);
        }
    } else if (pt >= 0) {
        for (uint64_t t = caps; t; t &= t - 1) out[n++] = make_mv(from, lsb(t), 0);
        for (uint64_t t = caps2; t; t &= t - 1) out[n++] = make_mv(from, lsb(t), 0);
        for (uint64_t t = quiets; t; t &= t - 1) out[nq++] = make_mv(from, lsb(t), 0);
        for (uint64_t t = quiets2; t; t &= t - 1) out[nq++] = make_mv(from, lsb(t), 0);
    }
    if (n_caps) *n_caps = total_caps + total_proms;
    __syncwarp();
    return total_caps + total_proms + total_quiets;
}

// Legal moves of `board` in generation order: the warp generates the pseudo-legal list, the lanes test
// legality in parallel, a ballot compacts in order.  Returns the count (uniform); moves in `legal`; *n_tactical = how
// many of them come from the capture list (the candidates of the tactical phase).
__device__ __forceinline__ int mcts_gen_legal_warp(const cb_board& board, Move* pseudo, Move* legal, int lane, int* n_tactical = nullptr) {
    int ncap = 0;
    const int np = gen_pseudo_warp(board, pseudo, lane, &ncap);
    const int us = board.w_to_move ? 0 : 1;
    int n = 0, nt = 0;
    for (int base = 0; base < np; base += 32) {
        const int i = base + lane;
        bool ok = false;
        Move m = 0;
        if (i < np) {
            m = pseudo[i];
            cb_board nb;
            cbchess::make_move_hd(board, m, &nb);
            ok = !cbchess::in_check_hd(nb, us);
        }
        const uint32_t mask = __ballot_sync(0xffffffffu, ok);
        if (ok) legal[n + __popc(mask & ((1u << lane) - 1u))] = m;
        n += __popc(mask);
        nt += __popc(__ballot_sync(0xffffffffu, ok && i < ncap));
    }
    if (n_tactical) *n_tactical = nt;
    __syncwarp();
    return n;
}

// Warp form of cbchess::light_gates_hd (host/gates.hpp): the value of a gate hit for the side to move of `b`, or 2.
// `legal[0..n)` are the legal moves of `b`; `nb` and `pseudo` are shared scratch.  All lanes call and return alike.
//  * Which moves give check is decided without making them: a knight / pawn checks from the squares that attack the
//    king square, a slider from the squares a rook / bishop standing on the king square would see.  Moves that can
//    uncover a line (the mover stands on a rank, file or diagonal of the enemy king), en passant, castling and
//    promotions take the exact test (make the move, test the king).
//  * Every checking move is then examined by the whole warp: the king's steps on eight lanes first (the usual way
//    out), and only when none is legal the full reply list (warp generator, one reply per lane).
// *dist (nullable) = terminal_distance of a hit (src/mcts/tactical_mcts.rs:624,641,701): 0 king already on the hill,
// 1 hill in one, kMateIn1Dist for the mate.
__device__ __forceinline__ int light_gates_warp(const cb_board& b, const Move* legal, int n, bool koth, cb_board& nb, Move* pseudo,
                                                int lane, int* dist = nullptr) {
    using namespace cbchess;
    const Tables& T = tables();
    const int us = b.w_to_move ? 0 : 1, them = 1 - us;
    if (koth) {
        if (dist) *dist = 0;
        if (b.pieces[them * 6 + KING] & kKothCenter) return -1;
        const uint64_t k = b.pieces[us * 6 + KING];
        if (k & kKothCenter) return 1;
        if (dist) *dist = 1;
        const int ks = k ? lsb(k) : -1;
        bool step = false;
        for (int i = lane; i < n; i += 32) step |= mv_from(legal[i]) == ks && (kKothCenter >> mv_to(legal[i]) & 1);
        if (__any_sync(0xffffffffu, step)) return 1;
    }
    if (dist) *dist = kMateIn1Dist;
    const uint64_t ek = b.pieces[them * 6 + KING];
    if (!ek) return 2;
    const int eks = lsb(ek);
    const uint64_t occ = b.occ[0] | b.occ[1];
    const uint64_t rvis = rook_att(eks, occ), bvis = bishop_att(eks, occ);          // checking squares of the sliders
    const uint64_t klines = rook_att(eks, 0ull) | bishop_att(eks, 0ull);            // every square aligned with that king
    for (int base = 0; base < n; base += 32) {
        const int i = base + lane;
        bool chk = false;
        if (i < n) {
            const Move m = legal[i];
            const int from = mv_from(m), to = mv_to(m);
            const int pt = piece_on(b, us, 1ull << from);
            const bool exact = mv_promo(m) != 0 || (klines >> from & 1) || (pt == PAWN && to == b.en_passant) ||
                               (pt == KING && (to - from == 2 || from - to == 2));
            if (exact) {
                cb_board t;
                make_move_hd(b, m, &t);
                chk = in_check_hd(t, them);
            } else {
                chk = pt == PAWN ? (T.pawn[us][to] & ek) != 0 : pt == KNIGHT ? (T.knight[to] & ek) != 0
                    : pt == BISHOP ? (bvis >> to & 1) != 0 : pt == ROOK ? (rvis >> to & 1) != 0
                    : pt == QUEEN ? ((rvis | bvis) >> to & 1) != 0 : false;
            }
        }
        for (uint32_t cm = __ballot_sync(0xffffffffu, chk); cm; cm &= cm - 1) {
            const Move m = legal[base + __ffs(cm) - 1];
            __syncwarp();
            if (lane == 0) make_move_hd(b, m, &nb);
            __syncwarp();
            // the checked king's steps
            const uint64_t k2 = nb.pieces[them * 6 + KING];
            const int ks2 = lsb(k2);
            const uint64_t steps = T.king[ks2] & ~nb.occ[them];
            bool out = false;
            if (lane < __popcll(steps)) {
                cb_board t;
                make_move_hd(nb, make_mv(ks2, nth_set_bit(steps, lane), 0), &t);
                out = !in_check_hd(t, them);
            }
            if (__any_sync(0xffffffffu, out)) continue;
            // captures of the checker and interpositions
            const int np = gen_pseudo_warp(nb, pseudo, lane);
            bool reply = false;
            for (int rb = 0; rb < np && !reply; rb += 32) {
                bool ok = false;
                if (rb + lane < np) {
                    cb_board t;
                    make_move_hd(nb, pseudo[rb + lane], &t);
                    ok = !in_check_hd(t, them);
                }
                reply = __any_sync(0xffffffffu, ok);
            }
            if (!reply) return 1;
        }
    }
    return 2;
}

// ---------------------------------------------------------------------------------------------------------------
// The Tier-1 gates at full depth, warp form of cbchess::tier1_gates_host (host/gates.hpp): the hill in n
// (src/search/koth.rs:43-65,270-395) and the mate search (src/search/mate_search.rs:64-289) as ONE iterative any / all
// minimax over an explicit stack in global memory.  Both searches answer "is there a move such that for every reply
// ..." and their results do not depend on the order in which moves are tried, so a node is: the legal moves (warp
// generator), optionally filtered on the lanes (attacker plies of the checks-only levels and the last attacker ply:
// moves that give check; root-side plies of the hill search: moves after which the king is inside the ring that
// still reaches the centre), and a cursor.  Even plies are ANY nodes (the side that wants to win), odd plies ALL nodes.
//
// A search of a quiet position enters a handful of nodes; one with many checks can enter thousands.  The games of a
// pool advance in lock step, so a job that has entered `budget` positions in this step is SUSPENDED (its stack and
// cursors stay in global memory) and resumed in the game's next step: that game skips simulations meanwhile, the
// other games do not wait for it, and the trees come out the same because a game's simulations are sequential anyway.
struct GateStack {                 // the search stack of one job: in shared memory while the job runs
    cb_board boards[kGatePlies + 1];
    uint16_t list[kGatePlies][CB_MAX_MOVES];
};
struct GateJob {                   // a game's job in global memory: where the stack rests between steps
    GateStack st;                  // (a cap1 job keeps its Cap1Scratch boards here)
    int32_t frames[6][24];         // cap1 job: the frames of plies 0..20 (alpha, beta, state, cursor, depth, skip)
    uint16_t idx[16], cnt[16];     // gate job: cursor and length of every ply's move list
    uint8_t active;                // 0 none, 1 leaf gate, 2 adjudication of a new root, 3 the leaf's cap1 search (dM)
    uint8_t phase;                 // 0 hill, 1 mate
    uint8_t it;                    // iteration: the hill's n / the mate's d (0: a filed job that has not started)
    uint8_t ply;
    uint16_t n_moves;              // cap1 job: the leaf's legal moves / how many come from the capture list (the rest of the
    uint8_t n_tact;                // selection is done; the evaluation request waits for dM)
    uint8_t stripe_state;          // this warp's share of a gate job: 0 not started, 1 suspended, 2 done
    uint8_t res_value, res_dist;   // ... its result
    uint8_t pad[2];
    uint32_t remaining;            // (first job of the game) shares of the gate job that are not done yet
};
static_assert(sizeof(GateJob) == kGateJobBytes && offsetof(GateJob, active) == kGateJobActiveOffset, "gate job layout");
static_assert(sizeof(GateStack) % 16 == 0 && sizeof(cb_board) % 16 == 0, "stack copied as uint4");
struct GateCfg { int koth_n, mate_d, exh_d; };

// The per-step allowance of a job: `left` positions, then up to `stretch` more as long as fewer than `target` games of this
// launch have finished their step (*ready, counted by mcts_gate_kernel; nullptr: no stretching).
struct GateBudget {
    int left, stretch;
    const unsigned int* ready;
    unsigned int target;
    __device__ __forceinline__ bool spent() {
        if (left-- > 0) return false;
        if (!ready || stretch-- <= 0) return true;
        return *reinterpret_cast<const volatile unsigned int*>(ready) >= target;
    }
};
__device__ __forceinline__ GateBudget gate_budget_of(const MctsParams& p, bool stretchable) {
    GateBudget b = {p.gate_budget, p.gate_stretch, stretchable ? p.gate_ready : nullptr, p.gate_target};
    return b;
}

// Does the side to move of `b` (shared memory) have a legal move?  King steps on eight lanes first, then the list.  The
// moves are tested without making them (cbchess::move_flags_hd).
__device__ __forceinline__ bool has_legal_move_warp(const cb_board& b, Move* pseudo, int lane) {
    using namespace cbchess;
    const MoveCtx c = move_ctx(b);
    if (c.ks >= 0) {
        const uint64_t steps = tables().king[c.ks] & ~b.occ[c.us];
        bool out = false;
        if (lane < __popcll(steps)) out = move_flags_hd(b, c, make_mv(c.ks, nth_set_bit(steps, lane), 0), false) != 0;
        if (__any_sync(0xffffffffu, out)) return true;
    }
    const int np = gen_pseudo_warp(b, pseudo, lane);
    for (int base = 0; base < np; base += 32) {
        const bool ok = base + lane < np && move_flags_hd(b, c, pseudo[base + lane], false) != 0;
        if (__any_sync(0xffffffffu, ok)) return true;
    }
    return false;
}

// The moves of a gate node in one pass over the pseudo-legal list: every lane tests its move's legality and with it the
// node's filter (kind 1: the move gives check; kind 2: the mover's king ends inside `ring`).  The kept moves go to `list`
// in generation order; returns their number, *n_legal = the number of legal moves.
__device__ __forceinline__ int gate_moves_warp(const cb_board& b, Move* pseudo, uint16_t* list, int kind, uint64_t ring, int lane,
                                               int* n_legal) {
    using namespace cbchess;
    const MoveCtx c = move_ctx(b);
    const int np = gen_pseudo_warp(b, pseudo, lane);
    int n = 0, nl = 0;
    for (int base = 0; base < np; base += 32) {
        const int i = base + lane;
        bool legal = false, keep = false;
        Move m = 0;
        if (i < np) {
            m = pseudo[i];
            const int fl = move_flags_hd(b, c, m, kind == 1);
            legal = (fl & 1) != 0;
            keep = legal && (kind == 0 || (kind == 1 ? (fl & 2) != 0 : (ring >> (mv_from(m) == c.ks ? mv_to(m) : c.ks) & 1) != 0));
        }
        const uint32_t km = __ballot_sync(0xffffffffu, keep);
        if (keep) list[n + __popc(km & ((1u << lane) - 1u))] = m;
        n += __popc(km);
        nl += __popc(__ballot_sync(0xffffffffu, legal));
    }
    __syncwarp();
    *n_legal = nl;
    return n;
}

// Runs (fresh: st.boards[0] holds the position) or resumes the job (the stack is fetched from job.st).  true: finished,
// *value = +1 (the side to move of the position wins: hill reached or mate) or 2 (no gate), *dist = terminal_distance
// (src/mcts/tactical_mcts.rs:641,701).  false: suspended, the stack is back in job.st.  `st` / `pseudo`: the warp's
// shared memory.  All lanes call and return alike.
// root_first / root_stride: this call's share of the work - at the gate's root (ply 0) only the moves first, first + stride,
// ... of every iteration's move list are tried.  Both searches are "is there a root move such that ...", so the shares
// of a job are independent searches and the job's answer is the earliest hit among them (smallest *dist: the hill's n
// before the mate's 64 + plies, as the sequential search finds them).
__device__ __forceinline__ bool tier1_gates_run(GateJob& job, GateStack& st, const GateCfg gc, GateBudget budget, bool fresh, Move* pseudo,
                                                int lane, int* value, int* dist, int root_first = 0, int root_stride = 1) {
    using namespace cbchess;
    int phase, it, ply, r_idx = 0, r_cnt = 0;   // lane p keeps the cursor of ply p
    *value = 2;
    *dist = 0;
    if (fresh) {
        if (gc.koth_n > 0) {   // koth_center_in_n: Some(0)
            const int us = st.boards[0].w_to_move ? 0 : 1;
            if (st.boards[0].pieces[us * 6 + KING] & kKothCenter) { *value = 1; return true; }
        }
        phase = gc.koth_n > 0 ? 0 : 1;
        if (phase == 1 && gc.mate_d <= 0) return true;
        it = 1;
        ply = 0;
    } else {
        phase = job.phase;
        it = job.it;
        ply = job.ply;
        if (lane <= kGatePlies) { r_idx = job.idx[lane]; r_cnt = job.cnt[lane]; }
        // the boards up to `ply` and the move lists below it (the list of `ply` itself is generated on entry)
        const int nb4 = (ply + 1) * (int)(sizeof(cb_board) / 16);
        for (int i = lane; i < nb4; i += 32) reinterpret_cast<uint4*>(st.boards)[i] = reinterpret_cast<const uint4*>(job.st.boards)[i];
        for (int q = 0; q < ply; ++q) {
            const int n4 = (__shfl_sync(0xffffffffu, r_cnt, q) * 2 + 15) / 16;
            for (int i = lane; i < n4; i += 32) reinterpret_cast<uint4*>(st.list[q])[i] = reinterpret_cast<const uint4*>(job.st.list[q])[i];
        }
        __syncwarp();
    }
    bool need_enter = true;   // a job starts, and is suspended, in front of a position
    for (;;) {
        const bool hill = phase == 0;
        const int D = 2 * it - 1;                            // hill: max_ply; mate: depth in plies
        const bool c_only = !hill && D > 2 * gc.exh_d - 1;   // checks-only level
        int result = -1;
        if (need_enter) {
            if (budget.spent()) {
                if (lane == 0) { job.phase = (uint8_t)phase; job.it = (uint8_t)it; job.ply = (uint8_t)ply; }
                if (lane <= kGatePlies) { job.idx[lane] = (uint16_t)r_idx; job.cnt[lane] = (uint16_t)r_cnt; }
                const int nb4 = (ply + 1) * (int)(sizeof(cb_board) / 16);
                for (int i = lane; i < nb4; i += 32) reinterpret_cast<uint4*>(job.st.boards)[i] = reinterpret_cast<const uint4*>(st.boards)[i];
                for (int q = 0; q < ply; ++q) {
                    const int n4 = (__shfl_sync(0xffffffffu, r_cnt, q) * 2 + 15) / 16;
                    for (int i = lane; i < n4; i += 32) reinterpret_cast<uint4*>(job.st.list[q])[i] = reinterpret_cast<const uint4*>(st.list[q])[i];
                }
                __syncwarp();
                return false;
            }
            need_enter = false;
            const cb_board& cur = st.boards[ply];
            const int stm = cur.w_to_move ? 0 : 1;
            const bool any_node = (ply & 1) == 0;
            if (hill) {   // solve_koth's tests in front of the moves (koth.rs:272-300)
                const int root_side = any_node ? stm : 1 - stm;
                if (cur.pieces[root_side * 6 + KING] & kKothCenter) result = 1;
                else if (cur.pieces[(1 - root_side) * 6 + KING] & kKothCenter) result = 0;
                else if (ply > D) result = 0;
            } else if (ply == D) {   // depth 0: the last move was a check; mate iff there is no answer (mate_search.rs:129-238)
                result = has_legal_move_warp(cur, pseudo, lane) ? 0 : 1;
            }
            if (result < 0) {
                const int kind = any_node ? (hill ? 2 : (c_only || D - ply == 1) ? 1 : 0) : 0;
                int nl = 0;
                const int cnt = gate_moves_warp(cur, pseudo, st.list[ply], kind, kind == 2 ? koth_ring_mask(gc.koth_n - ply / 2 - 1) : 0ull,
                                                lane, &nl);
                if (cnt == 0) {
                    if (hill) result = 0;                                           // koth.rs:327,356,386-388
                    else if (any_node) result = (ply != 0 && !c_only && nl == 0 && in_check_hd(cur, stm)) ? 1 : 0;   // :276-283
                    else result = in_check_hd(cur, stm) ? 1 : 0;                    // mated / stalemate
                }
                if (lane == ply) { r_idx = ply == 0 ? root_first : 0; r_cnt = cnt; }
            }
        }
        if (result < 0) {
            const int idx = __shfl_sync(0xffffffffu, r_idx, ply), cnt = __shfl_sync(0xffffffffu, r_cnt, ply);
            if (idx < cnt) {
                if (lane == ply) r_idx = idx + (ply == 0 ? root_stride : 1);
                __syncwarp();
                if (lane == 0) make_move_hd(st.boards[ply], (Move)st.list[ply][idx], &st.boards[ply + 1]);
                __syncwarp();
                ++ply;
                need_enter = true;
                continue;
            }
            result = (ply & 1) ? 1 : 0;   // ALL node: every reply loses; ANY node: no move wins
        }
        while (ply > 0) {                 // hand the result up
            --ply;
            if (((ply & 1) == 0) ? result == 1 : result == 0) continue;   // decides the parent as well
            result = -1;                  // the parent tries its next move
            break;
        }
        if (result < 0) continue;
        if (result == 1) {
            *value = 1;
            *dist = hill ? it : mate_dist_u8(D);
            return true;
        }
        ++it;                             // nothing at this iteration
        if (it > (hill ? gc.koth_n : gc.mate_d)) {
            if (hill && gc.mate_d > 0) { phase = 1; it = 1; }
            else return true;
        }
        ply = 0;
        need_enter = true;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Warp form of cbchess::cap1_qsearch (host/qsearch.hpp; src/search/quiescence.rs:808-1012, `--qsearch cap1`): an
// iterative alpha-beta over an explicit stack.  A frame is: stand pat -> the first legal capture of the MVV-LVA order (the
// exchange line's keys, one own piece per lane) -> once per side and branch the first legal checking move of the
// generation order that is not "that capture" (identified as the reference does, :933-944) -> done; a frame in check
// walks its legal evasions.  Lane d keeps the scalars of ply d; the boards live in shared memory.
struct Cap1Scratch {
    cb_board boards[cbchess::kPeMaxDepth + 1];
    uint16_t list[CB_MAX_MOVES];   // the evasions of the frame in hand / the capture-list entries tried so far
};
static_assert(sizeof(Cap1Scratch) <= sizeof(GateStack), "cap1 borrows the gate stack");

// S.boards[0] holds the position (resume: the stack is fetched from the job).  true: finished, *cp_out = the score in
// centipawns, *completed_out = the reference's flag.  false (only with a job): the search has entered `budget` positions in
// this call and rests in the job.  All lanes.
__device__ __forceinline__ bool cap1_run(Cap1Scratch& S, GateJob* job, GateBudget budget, bool resume, Move* pseudo, uint8_t* mailbox, int lane,
                                         int* cp_out, bool* completed_out) {
    using namespace cbchess;
    enum { ENTER = 0, EVASIONS = 1, CAPTURE = 2, CHECK = 3, DONE = 4 };
    int f_alpha = 0, f_beta = 0, f_state = 0, f_idx = 0, f_maxd = 0, f_skip = 0;   // this lane's frame
    int d = 0;
    if (resume) {
        d = job->ply;
        if (lane <= kPeMaxDepth) {
            f_alpha = job->frames[0][lane]; f_beta = job->frames[1][lane]; f_state = job->frames[2][lane];
            f_idx = job->frames[3][lane]; f_maxd = job->frames[4][lane]; f_skip = job->frames[5][lane];
        }
        const int nb4 = (d + 1) * (int)(sizeof(cb_board) / 16);
        for (int i = lane; i < nb4; i += 32) reinterpret_cast<uint4*>(S.boards)[i] = reinterpret_cast<const uint4*>(&job->st)[i];
        __syncwarp();
    } else if (lane == 0) { f_alpha = -kPeWindow; f_beta = kPeWindow; f_state = ENTER; }
    for (;;) {
        int alpha = __shfl_sync(0xffffffffu, f_alpha, d);
        const int beta = __shfl_sync(0xffffffffu, f_beta, d);
        const int st = __shfl_sync(0xffffffffu, f_state, d);
        if (job && (st & 7) == ENTER && budget.spent()) {   // suspended in front of a position
            if (lane == 0) job->ply = (uint8_t)d;
            if (lane <= kPeMaxDepth) {
                job->frames[0][lane] = f_alpha; job->frames[1][lane] = f_beta; job->frames[2][lane] = f_state;
                job->frames[3][lane] = f_idx; job->frames[4][lane] = f_maxd; job->frames[5][lane] = f_skip;
            }
            const int nb4 = (d + 1) * (int)(sizeof(cb_board) / 16);
            for (int i = lane; i < nb4; i += 32) reinterpret_cast<uint4*>(&job->st)[i] = reinterpret_cast<const uint4*>(S.boards)[i];
            __syncwarp();
            return false;
        }
        int idx = __shfl_sync(0xffffffffu, f_idx, d), maxd = __shfl_sync(0xffffffffu, f_maxd, d), skip = __shfl_sync(0xffffffffu, f_skip, d);
        int stage = st & 7;
        const bool wu = (st & 8) != 0, bu = (st & 16) != 0;
        bool comp = (st & 32) != 0;
        const cb_board& b = S.boards[d];
        const int us = b.w_to_move ? 0 : 1;
        bool ret = false, push = false, ret_c = true;
        int ret_v = 0, ret_d = 0;
        bool cwu = wu, cbu = bu;   // the child's check budgets
        __syncwarp();
        if (stage == ENTER) {
            const bool inchk = in_check_warp(b, us, lane);
            comp = true;
            maxd = 0;
            idx = 0;
            skip = -1;
            if (!inchk) {
                const int sp = pesto_cp_warp(b, mailbox, lane);   // also fills the mailbox for the capture scan
                if (sp >= beta) { ret = true; ret_v = beta; }
                else if (sp > alpha) alpha = sp;
            }
            if (!ret && d == kPeMaxDepth) { ret = true; ret_v = alpha; ret_c = !inchk; }
            stage = inchk ? EVASIONS : CAPTURE;
        }
        if (!ret && stage == EVASIONS) {
            const int n = mcts_gen_legal_warp(b, pseudo, S.list, lane);
            if (n == 0) { ret = true; ret_v = -1000000; }
            else if (idx >= n) { ret = true; ret_v = alpha; ret_c = comp; ret_d = maxd; }
            else {
                if (lane == 0) make_move_hd(b, (Move)S.list[idx], &S.boards[d + 1]);
                ++idx;
                push = true;
            }
        } else if (!ret && stage == CAPTURE) {
            stage = CHECK;
            const uint64_t own = b.occ[us];
            const bool has_piece = lane < __popcll(own);
            int from = 0, pt = 0;
            if (has_piece) {
                from = nth_set_bit(own, lane);
                pt = (int)mailbox[from] - 1 - 6 * us;
            }
            Move m = 0;
            uint32_t key = has_piece ? pe_piece_best_mb(b, mailbox, from, pt, 0, &m) : kPeNoCapture;
            int n_tried = 0;
            Move played = 0;
            for (;;) {
                const uint32_t kmin = __reduce_min_sync(0xffffffffu, key);
                if (kmin == kPeNoCapture) break;
                if (lane < 16) reinterpret_cast<uint64_t*>(&S.boards[d + 1])[lane] = reinterpret_cast<const uint64_t*>(&b)[lane];
                __syncwarp();
                const bool winner = key == kmin;
                if (winner) {
                    pe_apply_capture(S.boards[d + 1], m, pt, mailbox[mv_to(m)] ? (int)mailbox[mv_to(m)] - 1 - 6 * (1 - us) : -1);
                    S.list[n_tried] = m;
                }
                ++n_tried;
                __syncwarp();
                if (!in_check_warp(S.boards[d + 1], us, lane)) { played = (Move)S.list[n_tried - 1]; break; }
                if (winner) key = pe_piece_best_mb(b, mailbox, from, pt, kmin, &m);
            }
            if (played) {
                const int best_to = mv_to(played);
                int sf = 64;                       // from-square of the first list entry with that target
                for (int i = n_tried - 1; i >= 0; --i)
                    if (mv_to((Move)S.list[i]) == best_to) sf = mv_from((Move)S.list[i]);
                skip = sf | (best_to << 8);
                push = true;
            }
        }
        if (!ret && !push && stage == CHECK) {
            stage = DONE;
            if (!(us == 0 ? wu : bu)) {
                const MoveCtx c = move_ctx(b);
                const int np = gen_pseudo_warp(b, pseudo, lane);
                for (int base = 0; base < np && !push; base += 32) {
                    const int i = base + lane;
                    bool keep = false;
                    if (i < np) {
                        const Move m = pseudo[i];
                        const bool same = skip >= 0 && mv_from(m) == (skip & 0xFF) && mv_to(m) == (skip >> 8);
                        keep = !same && move_flags_hd(b, c, m, true) == 3;
                    }
                    const uint32_t mask = __ballot_sync(0xffffffffu, keep);
                    if (mask) {
                        if (lane == 0) make_move_hd(b, pseudo[base + __ffs(mask) - 1], &S.boards[d + 1]);
                        push = true;
                        if (us == 0) cwu = true; else cbu = true;
                    }
                }
            }
        }
        if (!ret && !push) { ret = true; ret_v = alpha; ret_c = comp; ret_d = maxd; }   // DONE
        if (push) {
            if (lane == d) {
                f_alpha = alpha; f_state = stage | (wu ? 8 : 0) | (bu ? 16 : 0) | (comp ? 32 : 0);
                f_idx = idx; f_maxd = maxd; f_skip = skip;
            }
            if (lane == d + 1) { f_alpha = -beta; f_beta = -alpha; f_state = ENTER | (cwu ? 8 : 0) | (cbu ? 16 : 0); }
            __syncwarp();
            ++d;
            continue;
        }
        // hand the result up: a cut-off returns the parent's beta and cascades
        for (;;) {
            if (d == 0) {
                *completed_out = ret_c;
                *cp_out = ret_v;
                return true;
            }
            --d;
            int pa = __shfl_sync(0xffffffffu, f_alpha, d);
            const int pb = __shfl_sync(0xffffffffu, f_beta, d);
            const int pst = __shfl_sync(0xffffffffu, f_state, d);
            int pmax = __shfl_sync(0xffffffffu, f_maxd, d);
            bool pcomp = (pst & 32) != 0;
            if (ret_d + 1 > pmax) pmax = ret_d + 1;
            if (!ret_c) pcomp = false;
            const int sc = -ret_v;
            if (sc >= pb) { ret_v = pb; ret_c = pcomp; ret_d = pmax; continue; }
            if (sc > pa) pa = sc;
            if (lane == d) { f_alpha = pa; f_maxd = pmax; f_state = (pst & ~32) | (pcomp ? 32 : 0); }
            break;
        }
    }
}

// Phase A of a simulation for every game: walk down (tactical phase, root force-visit, PUCT), stop at an unexpanded or
// decided node.
// Per-warp shared scratch of the tree kernels.
struct MctsSmem {
    cb_board board[kMctsWarps];
    cb_board next[kMctsWarps];
    Move pseudo[kMctsWarps][CB_MAX_MOVES];
    Move legal[kMctsWarps][CB_MAX_MOVES];
    uint8_t mailbox[kMctsWarps][64];   // piece codes of the position under the exchange line
    GateStack gate[kMctsWarps];        // tier1_light = 2: the stack of the warp's gate job while it runs
    uint8_t reporter[kMctsWarps];      // mcts_gate_kernel: this warp accounts for its game's step (gate_ready)
};

// A self-play game in slot g is over (lane 0): count it, log the result record, start the next game.  `rs`: the game's
// random stream after the draws of its last move.
__device__ __forceinline__ void mcts_game_over(const MctsParams& p, int g, int result, uint64_t& rs) {
    if (p.match_mode) {
        // result from the candidate's point of view, logged in finish order for the SPRT; next game or idle
        const int cand = result == 0 ? 0 : ((result > 0) == (p.cand_white[g] != 0) ? 1 : -1);
        atomicAdd(&p.stats[cand > 0 ? MS_WWINS : cand < 0 ? MS_BWINS : MS_DRAWS], 1ull);
        atomicAdd(&p.stats[MS_GAMES], 1ull);
        const unsigned long long slot = atomicAdd(p.result_cursor, 1ull);
        p.result_log[slot % p.result_cap] = cand;
        const unsigned long long k = atomicAdd(p.games_started, 1ull);
        const uint64_t seed = rng_next(rs) ^ ((uint64_t)g << 32);
        mcts_reset_game(p, g, seed);
        p.cand_white[g] = (k & 1ull) == 0;
        if ((long long)k >= p.total_games) p.idle[g] = 1;
        return;
    }
    atomicAdd(&p.stats[result > 0 ? MS_WWINS : result < 0 ? MS_BWINS : MS_DRAWS], 1ull);
    atomicAdd(&p.stats[MS_SAMPLES], (unsigned long long)p.samples_in_game[g]);
    atomicAdd(&p.stats[MS_GAMES], 1ull);
    if (p.sample_ring) {
        const unsigned long long slot = atomicAdd(p.ring_cursor, 1ull);
        uint8_t* rec = p.sample_ring + (size_t)(slot % p.ring_slots) * kSampleSlotBytes;
        reinterpret_cast<uint32_t*>(rec)[0] = (uint32_t)g;
        reinterpret_cast<uint32_t*>(rec)[1] = p.game_serial[g];
        reinterpret_cast<uint16_t*>(rec)[4] = (uint16_t)p.ply[g];
        reinterpret_cast<uint16_t*>(rec)[5] = 0xFFFFu;
        reinterpret_cast<int32_t*>(rec)[3] = result;
        p.game_serial[g]++;
    }
    const uint64_t seed = rng_next(rs) ^ ((uint64_t)g << 32);
    mcts_reset_game(p, g, seed);
}

// The search run to completion in one call.
__device__ __forceinline__ int cap1_cp_warp(Cap1Scratch& S, Move* pseudo, uint8_t* mailbox, int lane, bool* completed_out) {
    int cp = 0;
    cap1_run(S, nullptr, GateBudget{0, 0, nullptr, 0u}, false, pseudo, mailbox, lane, &cp, completed_out);
    return cp;
}

// dM of the position in `b` = sm.board[wl] (overwritten when a search runs) by the configured search; *completed (nullable)
// = the search's completion flag (always true for the stand pat and the exchange line).  cap1 borrows the warp's gate stack,
// which is idle whenever a material term is computed (the leaf's gate job has finished; a suspended job rests in global
// memory).
__device__ __forceinline__ float mcts_material_q_sm(const MctsParams& p, MctsSmem& sm, int wl, int lane, bool always, bool* completed) {
    if (completed) *completed = true;
    if (p.qsearch != 2 || (!p.material_on && !always)) return mcts_material_q(p, sm.board[wl], sm.next[wl], sm.mailbox[wl], lane, always);
    Cap1Scratch& S = *reinterpret_cast<Cap1Scratch*>(&sm.gate[wl]);
    __syncwarp();
    if (lane < 16) reinterpret_cast<uint64_t*>(&S.boards[0])[lane] = reinterpret_cast<const uint64_t*>(&sm.board[wl])[lane];
    __syncwarp();
    bool c = true;
    const int cp = cap1_cp_warp(S, sm.pseudo[wl], sm.mailbox[wl], lane, &c);
    if (completed) *completed = c;
    return __fdiv_rn((float)cp, 100.0f);
}

// The part of a child node the walk needs, packed for one round of shuffles.
struct MChildLite {
    uint32_t first_child, visits, pk /* n_children | move << 16 */, st /* state | term_value << 8 | n_tact << 16 | t_done << 24 */;
};
__device__ __forceinline__ MChildLite lite_of(const MNode& n) {
    MChildLite l;
    l.first_child = n.first_child;
    l.visits = n.visits;
    l.pk = (uint32_t)n.n_children | ((uint32_t)n.move << 16);
    l.st = (uint32_t)n.state | ((uint32_t)(uint8_t)n.term_value << 8) | ((uint32_t)n.n_tact << 16) | ((uint32_t)n.t_done << 24);
    return l;
}

// MODE 0: everything in this launch (gate jobs run inline).  With tier1_light = 2 and one pool the work is split over two
// launches so that the gate jobs - long, latency-bound loops over a small body of code - run in a kernel of their own
// (mcts_gate_kernel) instead of inside the 400 KB tree kernel: MODE 1 (tree kernel) walks down and, at a fresh leaf, files
// the gate job; MODE 2 (gate kernel) runs the filed / suspended jobs and, when one finishes, completes the selection (decided
// leaf: backup; else the evaluation request) or adjudicates the root.
// In MODE 2 a game has gate_warps warps (`w` = this warp's index among them): each runs its share of the gate job (every
// gate_warps-th move of the gate's root, tier1_gates_run), and the warp that finishes the last share goes on alone.
template <int MODE>
__device__ __forceinline__ void mcts_select_body(const MctsParams& p, MctsSmem& sm, int g, int wl, int lane, int w = 0) {
    Move (&s_pseudo)[kMctsWarps][CB_MAX_MOVES] = sm.pseudo;
    Move (&s_legal)[kMctsWarps][CB_MAX_MOVES] = sm.legal;
    if ((p.match_mode || p.search_only) && p.idle[g]) {
        if (MODE != 2 && lane == 0) { p.wants_eval[g] = 0; p.eval_cnt[g] = 0; }
        if (MODE == 2 && w == 0 && lane == 0) sm.reporter[wl] = 1;
        return;
    }
    // tier1_light = 2: a gate job of this game may be waiting (tier1_gates_run)
    const bool deep = p.tier1_light >= 2;
    const int W = p.gate_warps > 0 ? p.gate_warps : 1;
    GateJob* job = p.gate_jobs ? &p.gate_jobs[(size_t)g * W] : nullptr;   // (allocated with tier1_light = 2 and with dM by cap1)
    int pending = job ? job->active : 0;
    if (MODE == 1 && pending) {   // the gate kernel carries the job on
        if (lane == 0) { p.wants_eval[g] = 0; p.eval_cnt[g] = 0; }
        return;
    }
    if (MODE == 2 && (!pending || pending == 3)) {
        if (w != 0) return;                          // nothing to share: the game's first warp
        if (lane == 0) sm.reporter[wl] = 1;
        if (!pending) return;
    }
    int pre_tv = 2, pre_td = 0;   // MODE 2: the gate's answer, put together from the shares
    if (MODE == 2 && pending != 3) {
        GateJob* mine = job + w;
        if (mine->stripe_state == 2) return;         // this share is done; others are still running
        const bool fresh_share = mine->stripe_state == 0;
        __syncwarp();
        if (fresh_share) {
            if (lane < 16) reinterpret_cast<uint64_t*>(&sm.gate[wl].boards[0])[lane] = reinterpret_cast<const uint64_t*>(&job->st.boards[0])[lane];
            __syncwarp();
        }
        const GateCfg gc = (pending == 2 && !p.match_mode) ? GateCfg{p.koth ? 3 : 0, 5, 2}
                                                           : GateCfg{p.koth ? p.t1_koth_depth : 0, p.t1_mate_depth, p.t1_exh_depth};
        int v = 2, d = 0;
        if (!tier1_gates_run(*mine, sm.gate[wl], gc, gate_budget_of(p, true), fresh_share, s_pseudo[wl], lane, &v, &d, w, W)) {
            if (lane == 0) { mine->stripe_state = 1; atomicAdd(&p.stats[MS_GATE_WAITS], 1ull); }
            return;
        }
        unsigned int old = 0;
        if (lane == 0) {
            mine->res_value = (uint8_t)v;
            mine->res_dist = (uint8_t)d;
            mine->stripe_state = 2;
            __threadfence();
            old = atomicSub(&job->remaining, 1u);
        }
        old = __shfl_sync(0xffffffffu, old, 0);
        if (old != 1u) return;                       // not the last share
        __threadfence();
        int best = 256;
        for (int k = 0; k < W; ++k) {
            const volatile GateJob* o = job + k;
            if (o->res_value == 1 && (int)o->res_dist < best) { pre_tv = 1; best = o->res_dist; }
        }
        pre_td = pre_tv == 1 ? best : 0;
        __syncwarp();
        if (lane == 0) {
            for (int k = 0; k < W; ++k) (job + k)->stripe_state = 0;
            sm.reporter[wl] = 1;
        }
        __syncwarp();
    }
    if (pending == 3) {
        // the selection is complete but for dM: the leaf's cap1 search was suspended (cap1_run); its result releases the request
        Cap1Scratch& S = *reinterpret_cast<Cap1Scratch*>(&sm.gate[wl]);
        int cp = 0;
        bool qc = true;
        if (!cap1_run(S, job, gate_budget_of(p, MODE == 2), true, s_pseudo[wl], sm.mailbox[wl], lane, &cp, &qc)) {
            if (lane == 0) { p.wants_eval[g] = 0; p.eval_cnt[g] = 0; atomicAdd(&p.stats[MS_GATE_WAITS], (unsigned long long)W); }
            return;
        }
        if (lane == 0) {
            job->active = 0;
            p.eval_q[g] = __fdiv_rn((float)cp, 100.0f);
            p.eval_cnt[g] = job->n_moves;
            p.leaf_ntact[g] = job->n_tact;
            p.wants_eval[g] = 1;
            atomicAdd(&p.stats[MS_EVALS], 1ull);
        }
        return;
    }
    if (pending == 2) {
        // the game loop's pre-check of the new root (src/bin/self_play.rs:241-283): a hill in 3 or mate_search(board, 5,
        // exhaustive 2) ends the game as a win of the side to move.  In a match the mover's search answers a root that its own
        // gate decides with the winning move (src/mcts/tactical_mcts.rs:330-420, the search's depths), move after move until
        // the win is on the board: the same result
        const GateCfg gc = p.match_mode ? GateCfg{p.koth ? p.t1_koth_depth : 0, p.t1_mate_depth, p.t1_exh_depth} : GateCfg{p.koth ? 3 : 0, 5, 2};
        int gv = pre_tv, gd = pre_td;
        if (MODE != 2 && !tier1_gates_run(*job, sm.gate[wl], gc, gate_budget_of(p, false), false, s_pseudo[wl], lane, &gv, &gd)) {
            if (lane == 0) { p.wants_eval[g] = 0; p.eval_cnt[g] = 0; atomicAdd(&p.stats[MS_GATE_WAITS], (unsigned long long)W); }
            return;
        }
        if (lane == 0) job->active = 0;
        __syncwarp();
        if (gv == 1) {   // adjudicated: the next game in this slot starts with this step's selection
            if (lane == 0) {
                uint64_t rs = p.rng[g];
                mcts_game_over(p, g, p.root[g].w_to_move ? 1 : -1, rs);
            }
            __syncwarp();
        }
        if (MODE == 2) return;   // the tree kernel selects the first leaf under this root in the next step
        pending = 0;
    }
    MNode* nodes = p.nodes + ((size_t)g * 2 + p.arena[g]) * p.node_cap;
    uint32_t* path = p.path + (size_t)g * kMctsMaxDepth;
    cb_board& board = sm.board[wl];
    uint32_t cur = 0;
    int path_n = 1;
    if (MODE == 2 || pending == 1) {   // the leaf gate of the simulation in flight was filed / suspended: same leaf, same path
        if (lane < 16) reinterpret_cast<uint64_t*>(&board)[lane] = reinterpret_cast<const uint64_t*>(&job->st.boards[0])[lane];
        __syncwarp();
        path_n = p.path_n[g];
        cur = path[path_n - 1];
    } else {
    if (lane < 16) reinterpret_cast<uint64_t*>(&board)[lane] = reinterpret_cast<const uint64_t*>(&p.root[g])[lane];
    __syncwarp();
    int depth = 0;
    if (lane == 0) path[0] = 0;
    // The node at the top of each level comes from the lane that scanned it as a child one level up (one
    // dependent memory round trip per level instead of two).
    MChildLite n = lite_of(nodes[0]);
    for (;;) {
        const uint32_t n_children = n.pk & 0xFFFFu, n_state = n.st & 0xFFu, n_tact = (n.st >> 16) & 0xFFu, t_done = n.st >> 24;
        if (n_state != 1 || depth >= kMctsMaxDepth - 1) break;
        uint32_t bi = 0xFFFFFFFFu;       // index of the chosen child
        MChildLite bch = {0, 0, 0, 0};   // ... as seen by the lane that scanned it
        if (t_done < n_tact) {
            // tactical phase (selection.rs:107-149): the next capture-list child in MVV-LVA order, whatever its statistics
            for (uint32_t i = lane; i < n_children; i += 32) {
                const MNode ch = nodes[n.first_child + i];
                if (ch.trank == t_done) { bi = i; bch = lite_of(ch); }
            }
            bi = __reduce_min_sync(0xffffffffu, bi);
            if (lane == 0) nodes[cur].t_done = (uint8_t)(t_done + 1);
        } else {
            bool forced = false;
            if (depth == 0) {   // every root child gets one visit first (selection.rs:65-73)
                for (uint32_t base = 0; base < n_children && !forced; base += 32) {
                    const uint32_t i = base + lane;
                    MNode ch;
                    ch.visits = 1;
                    if (i < n_children) ch = nodes[n.first_child + i];
                    const uint32_t mask = __ballot_sync(0xffffffffu, i < n_children && ch.visits == 0);
                    if (mask) {
                        bi = base + (__ffs(mask) - 1);
                        if (i == bi) bch = lite_of(ch);
                        forced = true;
                    }
                }
            }
            if (!forced) {
                // PUCT (selection.rs:152-273), f64: Q + c * P * sqrt(max(N, 1)) / (1 + n), first maximum; a child whose side
                // to move has a decided win is skipped, and when every child is one the longest delay is taken
                const double sq = sqrt((double)(n.visits > 1 ? n.visits : 1));
                double bu = -INFINITY;
                uint32_t li = 0xFFFFFFFFu;      // least-bad losing child of this lane
                int ld = -1;
                MChildLite lch = {0, 0, 0, 0};
                for (uint32_t i = lane; i < n_children; i += 32) {
                    const MNode ch = nodes[n.first_child + i];
                    if (ch.state == 2 && ch.term_value > 0) {
                        if ((int)ch.term_dist > ld) { ld = ch.term_dist; li = i; lch = lite_of(ch); }
                        continue;
                    }
                    const double q = ch.visits ? __ddiv_rn(ch.value_sum, (double)ch.visits) : 0.0;
                    const double u = __dadd_rn(q, __ddiv_rn(__dmul_rn(__dmul_rn(p.c_puct, (double)ch.prior), sq),
                                                            __dadd_rn(1.0, (double)ch.visits)));
                    if (u > bu) { bu = u; bi = i; bch = lite_of(ch); }
                }
                double wu = bu;
                uint32_t wi = bi;
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) {
                    const double ou = __shfl_xor_sync(0xffffffffu, wu, d);
                    const uint32_t oi = __shfl_xor_sync(0xffffffffu, wi, d);
                    if (oi != 0xFFFFFFFFu && (wi == 0xFFFFFFFFu || ou > wu || (ou == wu && oi < wi))) { wu = ou; wi = oi; }
                }
                if (wi == 0xFFFFFFFFu) {   // every child is a proven loss: first child with the largest distance
                    int wd = ld;
                    wi = li;
#pragma unroll
                    for (int d = 16; d >= 1; d >>= 1) {
                        const int od = __shfl_xor_sync(0xffffffffu, wd, d);
                        const uint32_t oi = __shfl_xor_sync(0xffffffffu, wi, d);
                        if (oi != 0xFFFFFFFFu && (wi == 0xFFFFFFFFu || od > wd || (od == wd && oi < wi))) { wd = od; wi = oi; }
                    }
                    bch = lch;                 // (the lane that owns the winner holds it as its own best)
                }
                bi = wi;
            }
        }
        if (bi == 0xFFFFFFFFu) break;             // no child (cannot happen for an expanded node)
        const uint32_t best = n.first_child + bi;
        const int owner = static_cast<int>(bi & 31u);   // child i was scanned by lane i & 31
        MChildLite nxt;
        nxt.first_child = __shfl_sync(0xffffffffu, bch.first_child, owner);
        nxt.visits = __shfl_sync(0xffffffffu, bch.visits, owner);
        nxt.pk = __shfl_sync(0xffffffffu, bch.pk, owner);
        nxt.st = __shfl_sync(0xffffffffu, bch.st, owner);
        if (lane == 0) {
            cb_board nb;
            cbchess::make_move_hd(board, (Move)(nxt.pk >> 16), &nb);
            board = nb;
            path[depth + 1] = best;
        }
        __syncwarp();
        cur = best;
        n = nxt;
        ++depth;
    }
    path_n = depth + 1;
    const uint8_t leaf_state = (uint8_t)(n.st & 0xFFu);
    const int8_t leaf_term = (int8_t)((n.st >> 8) & 0xFFu);
    if (leaf_state == 2 || leaf_state == 1) {   // decided (or the depth cap): back its value up again
        __syncwarp();
        mcts_backup(nodes, path, path_n, leaf_state == 2 ? (double)leaf_term : 0.0, lane);
        if (lane == 0) {
            atomicAdd(&p.stats[MS_TERMINAL], 1ull);
            p.wants_eval[g] = 0;
            p.eval_cnt[g] = 0;
        }
        mcts_sim_done(p, nodes, g, lane);
        return;
    }
    }
    MNode& leaf = nodes[cur];
    int n_tact = 0;
    const int n_moves = mcts_gen_legal_warp(board, s_pseudo[wl], s_legal[wl], lane, &n_tact);
    // mate / stalemate (MctsNode::new_child, src/mcts/node.rs:162-177), else the light Tier-1 gates; with the gates on, an
    // opponent king on the hill is tested before anything else (src/mcts/tactical_mcts.rs:619-632).  The root itself is
    // never gated: the game loop adjudicates it before the search (src/bin/self_play.rs:241-283)
    int tv = 2, td = 0;
    const int us = board.w_to_move ? 0 : 1;
    if (n_moves == 0) tv = cbchess::in_check_hd(board, us) ? -1 : 0;
    else if (p.tier1_light && cur != 0) {
        if (p.koth && (board.pieces[(1 - us) * 6 + cbchess::KING] & cbchess::kKothCenter)) tv = -1;
        else if (deep) {
            // the reference's leaf gate (src/mcts/tactical_mcts.rs:634-708): hill in koth_depth, then the mate search
            if (MODE == 1) {   // file the job for the gate kernel
                if (lane < 16) reinterpret_cast<uint64_t*>(&job->st.boards[0])[lane] = reinterpret_cast<const uint64_t*>(&board)[lane];
                if (lane == 0) {
                    for (int k = 0; k < W; ++k) (job + k)->stripe_state = 0;   // no share has started
                    job->remaining = (uint32_t)W;
                    job->active = 1;
                    p.path_n[g] = path_n;
                    p.wants_eval[g] = 0;
                    p.eval_cnt[g] = 0;
                }
                return;
            }
            if (MODE == 2) {   // the shares have answered
                tv = pre_tv;
                td = pre_td;
            } else {
                const bool fresh = pending != 1;
                __syncwarp();
                if (fresh) {
                    if (lane < 16) reinterpret_cast<uint64_t*>(&sm.gate[wl].boards[0])[lane] = reinterpret_cast<const uint64_t*>(&board)[lane];
                    __syncwarp();
                }
                const GateCfg gc = {p.koth ? p.t1_koth_depth : 0, p.t1_mate_depth, p.t1_exh_depth};
                if (!tier1_gates_run(*job, sm.gate[wl], gc, gate_budget_of(p, false), fresh, s_pseudo[wl], lane, &tv, &td)) {
                    if (lane == 0) {
                        job->active = 1;
                        p.path_n[g] = path_n;
                        p.wants_eval[g] = 0;
                        p.eval_cnt[g] = 0;
                        atomicAdd(&p.stats[MS_GATE_WAITS], (unsigned long long)W);
                    }
                    return;
                }
            }
            if (lane == 0) {
                job->active = 0;
                if (tv != 2) atomicAdd(&p.stats[MS_GATE_HITS], 1ull);
            }
            __syncwarp();
        } else {
            tv = light_gates_warp(board, s_legal[wl], n_moves, p.koth != 0, sm.next[wl], s_pseudo[wl], lane, &td);
        }
    }
    if (tv != 2) {
        if (lane == 0) {
            leaf.state = 2;
            leaf.term_value = (int8_t)tv;
            leaf.term_dist = (uint8_t)td;
        }
        __syncwarp();
        mcts_backup(nodes, path, path_n, (double)tv, lane);
        if (lane == 0) {
            atomicAdd(&p.stats[MS_TERMINAL], 1ull);
            p.wants_eval[g] = 0;
            p.eval_cnt[g] = 0;
        }
        mcts_sim_done(p, nodes, g, lane);
        return;
    }
    // tactical ranks of the capture-list moves (src/mcts/tactical.rs:155-177): stable descending MVV-LVA
    uint8_t* trank = p.leaf_trank + (size_t)g * CB_MAX_MOVES;
    for (int base = 0; base < n_moves; base += 32) {
        const int i = base + lane;
        if (i < n_moves) {
            uint8_t r = 0xFF;
            if (i < n_tact) {
                const int si = cbchess::tactical_score_hd(board, s_legal[wl][i]);
                int rk = 0;
                for (int j = 0; j < n_tact; ++j) {
                    const int sj = cbchess::tactical_score_hd(board, s_legal[wl][j]);
                    rk += sj > si || (sj == si && j < i);
                }
                r = (uint8_t)rk;
            }
            trank[i] = r;
        }
    }
    __syncwarp();
    if (p.shuffle && lane == 0) {   // SliceRandom::shuffle: for i in (1..len).rev() swap(i, uniform(0..=i))
        uint64_t rs = p.rng[g];
        for (int i = n_moves - 1; i >= 1; --i) {
            const int j = (int)(rng_next(rs) % (uint64_t)(i + 1));
            const Move tm = s_legal[wl][i]; s_legal[wl][i] = s_legal[wl][j]; s_legal[wl][j] = tm;
            const uint8_t tr = trank[i]; trank[i] = trank[j]; trank[j] = tr;
        }
        p.rng[g] = rs;
    }
    __syncwarp();
    const bool wtm = board.w_to_move != 0;
    for (int i = lane; i < n_moves; i += 32) {
        const Move m = s_legal[wl][i];
        p.leaf_moves[(size_t)g * CB_MAX_MOVES + i] = m;
        p.eval_idx[(size_t)g * CB_MAX_MOVES + i] = (uint16_t)cbchess::policy_index_hd(m, wtm);
    }
    if (lane < 16) reinterpret_cast<uint64_t*>(&p.eval_boards[g])[lane] = reinterpret_cast<const uint64_t*>(&board)[lane];
    __syncwarp();
    // a root asks for its policy only, with dM = 0 (selection.rs:315-332)
    float leaf_q = 0.0f;
    if (cur != 0) {
        if (p.qsearch == 2 && p.material_on && job) {
            // dM by cap1: a search of up to ~1000 positions, run under the per-step budget of the gate jobs
            Cap1Scratch& S = *reinterpret_cast<Cap1Scratch*>(&sm.gate[wl]);
            __syncwarp();
            if (lane < 16) reinterpret_cast<uint64_t*>(&S.boards[0])[lane] = reinterpret_cast<const uint64_t*>(&board)[lane];
            __syncwarp();
            int cp = 0;
            bool qc = true;
            if (!cap1_run(S, job, gate_budget_of(p, MODE == 2), false, s_pseudo[wl], sm.mailbox[wl], lane, &cp, &qc)) {
                if (lane == 0) {
                    job->active = 3;
                    job->n_moves = (uint16_t)n_moves;
                    job->n_tact = (uint8_t)n_tact;
                    p.path_n[g] = path_n;
                    p.wants_eval[g] = 0;
                    p.eval_cnt[g] = 0;
                    atomicAdd(&p.stats[MS_GATE_WAITS], (unsigned long long)W);
                }
                return;
            }
            leaf_q = __fdiv_rn((float)cp, 100.0f);
        } else {
            leaf_q = mcts_material_q_sm(p, sm, wl, lane, false, nullptr);   // `board` is dead after this
        }
    }
    if (lane == 0) {
        p.eval_q[g] = leaf_q;
        p.eval_cnt[g] = (uint16_t)n_moves;
        p.leaf_ntact[g] = (uint8_t)n_tact;
        p.path_n[g] = path_n;
        p.wants_eval[g] = 1;
        atomicAdd(&p.stats[MS_EVALS], 1ull);
    }
}

template <int MODE>
__global__ void __launch_bounds__(32 * kMctsWarps) mcts_select_kernel(const MctsParams p) {
    __shared__ MctsSmem sm;
    const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = p.g0 + blockIdx.x * kMctsWarps + wl;
    if (MODE == 1 && p.gate_ready && blockIdx.x == 0 && threadIdx.x == 0) *p.gate_ready = 0;   // counted by the gate kernel that follows
    if (g >= p.g0 + p.gn) return;
    mcts_select_body<MODE>(p, sm, g, wl, lane);
}

// tier1_light = 2, one pool: the games' gate jobs (and what follows a finished one), between the tree kernel and the forward.
// gate_warps warps per game.
__global__ void __launch_bounds__(32 * kMctsWarps) mcts_gate_kernel(const MctsParams p) {
    __shared__ MctsSmem sm;
    const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int W = p.gate_warps > 0 ? p.gate_warps : 1;
    const int widx = blockIdx.x * kMctsWarps + wl;
    const int g = p.g0 + widx / W, w = widx % W;
    if (g >= p.g0 + p.gn) return;
    if (lane == 0) sm.reporter[wl] = 0;
    __syncwarp();
    mcts_select_body<2>(p, sm, g, wl, lane, w);
    __syncwarp();
    // the warp that accounts for the game: its step is complete unless a job is still pending
    if (lane == 0 && p.gate_ready && sm.reporter[wl] && !p.gate_jobs[(size_t)g * W].active) atomicAdd(p.gate_ready, 1u);
}

// Keep the subtree under `child` as the new tree, copied breadth first into the other arena (same node
// order as the host driver's reroot).  All lanes.
__device__ __forceinline__ void mcts_reroot(const MctsParams& p, int g, uint32_t child, int lane) {
    const uint8_t a = p.arena[g];
    const MNode* old = p.nodes + ((size_t)g * 2 + a) * p.node_cap;
    MNode* nw = p.nodes + ((size_t)g * 2 + (a ^ 1)) * p.node_cap;
    if (lane == 0) nw[0] = old[child];
    __syncwarp();
    uint32_t tail = 1;
    for (uint32_t base = 0; base < tail;) {
        const uint32_t batch = tail - base < 32u ? tail - base : 32u;
        MNode on;
        on.state = 0; on.first_child = 0; on.n_children = 0;
        if ((uint32_t)lane < batch) on = nw[base + lane];
        uint32_t mask = __ballot_sync(0xffffffffu, (uint32_t)lane < batch && on.state == 1);
        while (mask) {
            const int l = __ffs(mask) - 1;
            mask &= mask - 1;
            const uint32_t fc = __shfl_sync(0xffffffffu, on.first_child, l);
            const uint32_t nc = __shfl_sync(0xffffffffu, (uint32_t)on.n_children, l);
            for (uint32_t i = lane; i < nc; i += 32) nw[tail + i] = old[fc + i];
            if (lane == 0) nw[base + l].first_child = tail;
            tail += nc;
        }
        __syncwarp();
        base += batch;
    }
    if (lane == 0) {
        p.n_nodes[g] = tail;
        p.arena[g] = a ^ 1;
    }
    __syncwarp();
}

// Phase C: expand the evaluated leaf with its priors, back the value up; at the end of the move's
// simulation budget choose and play the move, re-root, detect the end of the game.
__device__ __forceinline__ void mcts_expand_body(const MctsParams& p, MctsSmem& sm, int g, int wl, int lane) {
    Move (&s_pseudo)[kMctsWarps][CB_MAX_MOVES] = sm.pseudo;
    Move (&s_legal)[kMctsWarps][CB_MAX_MOVES] = sm.legal;
    MNode* nodes = p.nodes + ((size_t)g * 2 + p.arena[g]) * p.node_cap;
    if ((p.match_mode || p.search_only) && p.idle[g]) return;
    // whose search is this?  (match mode: the model of the side to move at the root)
    const bool second = p.match_mode && ((p.root[g].w_to_move != 0) != (p.cand_white[g] != 0));
    const float* priors = second ? p.priors_b : p.priors;
    const float* v_final = second ? p.v_final_b : p.v_final;
    if (p.wants_eval[g]) {
        const uint32_t* path = p.path + (size_t)g * kMctsMaxDepth;
        const int path_n = p.path_n[g];
        const uint32_t cur = path[path_n - 1];
        const int n = p.eval_cnt[g];
        const uint32_t first = p.n_nodes[g];
        if (first + (uint32_t)n <= p.node_cap) {
            for (int i = lane; i < n; i += 32) {
                MNode ch;
                ch.value_sum = 0.0; ch.first_child = 0; ch.visits = 0;
                ch.prior = priors[(size_t)g * CB_MAX_MOVES + i];
                ch.n_children = 0;
                ch.move = p.leaf_moves[(size_t)g * CB_MAX_MOVES + i];
                ch.state = 0; ch.term_value = 0; ch.n_tact = 0; ch.t_done = 0;
                ch.trank = p.leaf_trank[(size_t)g * CB_MAX_MOVES + i];
                ch.term_dist = 0; ch.pad = 0;
                nodes[first + i] = ch;
            }
            if (lane == 0) {
                nodes[cur].first_child = first;
                nodes[cur].n_children = (uint16_t)n;
                nodes[cur].n_tact = p.leaf_ntact[g];
                nodes[cur].t_done = 0;
                nodes[cur].state = 1;
                p.n_nodes[g] = first + (uint32_t)n;
            }
        } else if (lane == 0) {
            atomicAdd(&p.stats[MS_OVERFLOW], 1ull);   // arena full: the leaf stays unexpanded (evaluated again later)
        }
        __syncwarp();
        if (cur != 0) {    // (the root only asked for its priors: no value, no simulation)
            mcts_backup(nodes, path, path_n, (double)v_final[g], lane);
            mcts_sim_done(p, nodes, g, lane);
        }
        __syncwarp();
    }
    if (p.sims_done[g] < p.sims_per_move) return;
    if (p.search_only) {   // debug search: the tree stays as it is for the host to read
        if (lane == 0) p.idle[g] = 1;
        return;
    }

    // ------------------------------------------------------------------ play a move
    const MNode root = nodes[0];
    if (root.state != 1 || root.n_children == 0) {   // cannot happen for a live game; restart defensively
        if (lane == 0) {
            uint64_t s = p.rng[g];
            const uint64_t seed = rng_next(s);
            mcts_reset_game(p, g, seed);
        }
        return;
    }
    uint32_t pick = root.first_child;
    uint64_t rs = p.rng[g];
    bool forced_win = false;
    if (p.match_mode && lane == 0) {
        // a child where the opponent is mated is played at once (select_eval_move, evaluate_models.rs:98-116)
        for (uint32_t i = 0; i < root.n_children && !forced_win; ++i) {
            const MNode ch = nodes[root.first_child + i];
            if (ch.state == 2 && ch.term_value < 0) { pick = root.first_child + i; forced_win = true; }
        }
    }
    if (lane == 0 && !forced_win) {
        uint32_t best_v = 0;     // select_greedy: Iterator::max_by_key keeps the LAST maximum (src/bin/self_play.rs:648-650)
        for (uint32_t i = 0; i < root.n_children; ++i) {
            const uint32_t v = nodes[root.first_child + i].visits;
            if (v >= best_v) { best_v = v; pick = root.first_child + i; }
        }
        const int move_number = p.ply[g] / 2 + 1;
        if (rng_uniform(rs) < pow(p.explore_base, (double)(move_number - 1))) {
            uint32_t tot1 = 0;
            for (uint32_t i = 0; i < root.n_children; ++i) {
                const uint32_t v = nodes[root.first_child + i].visits;
                tot1 += v > 0 ? v - 1 : 0;
            }
            if (tot1 > 0) {
                const uint32_t thr = (uint32_t)(rng_next(rs) % tot1);
                uint32_t cum = 0;
                for (uint32_t i = 0; i < root.n_children; ++i) {
                    const uint32_t v = nodes[root.first_child + i].visits;
                    cum += v > 0 ? v - 1 : 0;
                    if (cum > thr) { pick = root.first_child + i; break; }
                }
            }
        }
    }
    if (lane == 0) p.samples_in_game[g]++;   // the (board, visit distribution) sample of this move
    pick = __shfl_sync(0xffffffffu, pick, 0);
    if (p.sample_ring) {
        unsigned long long slot = 0;
        if (lane == 0) slot = atomicAdd(p.ring_cursor, 1ull);
        slot = __shfl_sync(0xffffffffu, slot, 0);
        uint8_t* rec = p.sample_ring + (size_t)(slot % p.ring_slots) * kSampleSlotBytes;
        uint32_t total = 0;
        for (uint32_t i = lane; i < root.n_children; i += 32) total += nodes[root.first_child + i].visits;
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) total += __shfl_xor_sync(0xffffffffu, total, d);
        const bool wtm = p.root[g].w_to_move != 0;
        const uint32_t n = total > 0 ? root.n_children : 0;
        for (uint32_t i = lane; i < n; i += 32) {
            const MNode ch = nodes[root.first_child + i];
            const uint32_t v = ch.visits > 65535u ? 65535u : ch.visits;
            reinterpret_cast<uint32_t*>(rec + 144)[i] = (uint32_t)(uint16_t)cbchess::policy_index_hd(ch.move, wtm) | (v << 16);
        }
        if (lane < 16) {
            const uint64_t w = reinterpret_cast<const uint64_t*>(&p.root[g])[lane];
            reinterpret_cast<uint64_t*>(rec + 16)[lane] = w;
            reinterpret_cast<uint64_t*>(&sm.board[wl])[lane] = w;
        }
        __syncwarp();
        bool q_completed = true;
        const float root_q = mcts_material_q_sm(p, sm, wl, lane, true, &q_completed);
        if (lane == 0) {
            reinterpret_cast<uint32_t*>(rec)[0] = (uint32_t)g;
            reinterpret_cast<uint32_t*>(rec)[1] = p.game_serial[g];
            reinterpret_cast<uint16_t*>(rec)[4] = (uint16_t)(p.ply[g] | (q_completed ? 0 : 0x8000));   // bit 15: qsearch_completed = false
            reinterpret_cast<uint16_t*>(rec)[5] = (uint16_t)n;
            reinterpret_cast<float*>(rec)[3] = root_q;
        }
        __syncwarp();
    }
    cb_board& nb = sm.next[wl];
    bool white_moved = false;
    if (lane == 0) {
        const cb_board rb = p.root[g];
        white_moved = rb.w_to_move != 0;
        cbchess::make_move_hd(rb, nodes[pick].move, &nb);
        uint64_t* hist = p.history + (size_t)g * kMctsHist;
        int hn = nb.halfmove == 0 ? 0 : p.hist_n[g];
        if (hn < kMctsHist) hist[hn++] = cbchess::position_key_hd(nb);
        p.hist_n[g] = hn;
        p.root[g] = nb;
        p.ply[g]++;
        p.sims_done[g] = 0;
        atomicAdd(&p.stats[MS_MOVES], 1ull);
    }
    __syncwarp();
    if (p.match_mode) {
        // the next move is searched by the other model: fresh tree
        if (lane == 0) {
            MNode z;
            memset(&z, 0, sizeof(z));
            nodes[0] = z;
            p.n_nodes[g] = 1;
        }
        __syncwarp();
    } else {
        mcts_reroot(p, g, pick, lane);
    }
    // game end
    const int n_legal = mcts_gen_legal_warp(nb, s_pseudo[wl], s_legal[wl], lane);
    // a root where the side to move wins at once is adjudicated (src/bin/self_play.rs:241-283 at depth 1; in a match
    // the reference's search returns the winning move there, src/mcts/tactical_mcts.rs:330-357,400-420: same result)
    const int gate = (p.tier1_light == 1 && n_legal > 0) ? light_gates_warp(nb, s_legal[wl], n_legal, p.koth != 0, sm.board[wl], s_pseudo[wl], lane) : 2;
    if (lane == 0) {
        int result = 2;   // +1 white won, -1 black won, 0 draw, 2 continue
        if (p.koth && (nb.pieces[(white_moved ? 0 : 6) + cbchess::KING] & cbchess::kKothCenter)) result = white_moved ? 1 : -1;
        else if (n_legal == 0) result = cbchess::in_check_hd(nb, nb.w_to_move ? 0 : 1) ? (white_moved ? 1 : -1) : 0;
        else {
            const uint64_t* hist = p.history + (size_t)g * kMctsHist;
            const int hn = p.hist_n[g];
            int reps = 0;
            for (int i = 0; i < hn; ++i) reps += (hist[i] == hist[hn - 1]);
            if (reps >= 3 || nb.halfmove >= 100 || p.ply[g] > p.max_plies) result = 0;
            else if (gate == 1) result = white_moved ? -1 : 1;
        }
        if (result != 2) {
            mcts_game_over(p, g, result, rs);
        } else {
            p.rng[g] = rs;
            if (p.tier1_light >= 2) {   // the new root's pre-check runs as a job in front of its first simulation
                const int W = p.gate_warps > 0 ? p.gate_warps : 1;
                GateJob& job = p.gate_jobs[(size_t)g * W];
                job.st.boards[0] = nb;
                job.phase = p.koth ? 0 : 1;     // (the inline form resumes it from here; the gate kernel's shares start afresh)
                job.it = 1;
                job.ply = 0;
                for (int k = 0; k < W; ++k) (&job + k)->stripe_state = 0;
                job.remaining = (uint32_t)W;
                job.active = 2;
            }
        }
    }
}

__global__ void __launch_bounds__(32 * kMctsWarps) mcts_expand_kernel(const MctsParams p) {
    __shared__ MctsSmem sm;
    const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = p.g0 + blockIdx.x * kMctsWarps + wl;
    if (g >= p.g0 + p.gn) return;
    mcts_expand_body(p, sm, g, wl, lane);
}

// Steady state: expand the leaf of simulation n, then select the leaf of simulation n + 1, in one launch.
template <int MODE>
__global__ void __launch_bounds__(32 * kMctsWarps) mcts_step_kernel(const MctsParams p) {
    __shared__ MctsSmem sm;
    const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = p.g0 + blockIdx.x * kMctsWarps + wl;
    if (MODE == 1 && p.gate_ready && blockIdx.x == 0 && threadIdx.x == 0) *p.gate_ready = 0;   // counted by the gate kernel that follows
    if (g >= p.g0 + p.gn) return;
    mcts_expand_body(p, sm, g, wl, lane);
    __syncwarp();
    mcts_select_body<MODE>(p, sm, g, wl, lane);
    if (p.pdl_done) {
        __syncwarp();
        unsigned int last = 0;
        if (lane == 0) last = atomicInc(p.pdl_done, (unsigned int)p.gn - 1u) == (unsigned int)p.gn - 1u;
        if (__shfl_sync(0xffffffffu, last, 0)) asm volatile("griddepcontrol.wait;" ::: "memory");
    }
}

// The reference's mock evaluator (InferenceServer::new_mock_biased, src/mcts/inference_server.rs:103-122) in place of
// the network: an all-zero policy, i.e. the uniform fallback 1 / len of policy_to_move_priors (src/tensor.rs:202-211),
// a constant v_logit and k; the value is fused like the network's (tanhf(v + k * dM), or tanhf(v) without material).
__global__ void mcts_mock_eval_kernel(const MctsParams p, float v_logit, float k, float* priors, float* v_final) {
    const int g = p.g0 + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (g >= p.g0 + p.gn || !p.wants_eval[g]) return;
    const int n = p.eval_cnt[g];
    const float u = __fdiv_rn(1.0f, (float)n);
    for (int i = lane; i < n; i += 32) priors[(size_t)g * CB_MAX_MOVES + i] = u;
    if (lane == 0) v_final[g] = tanhf(p.material_on ? __fadd_rn(v_logit, __fmul_rn(k, p.eval_q[g])) : v_logit);
}

}  // namespace cb
