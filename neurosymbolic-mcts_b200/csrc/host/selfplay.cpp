// selfplay.cpp — per-GPU self-play game pool feeding the batched leaf evaluator, plus the
// cb_chess_* C ABI.  Semantics follow src/bin/self_play.rs:191-624 and src/mcts/{selection,node}.rs
// (citations at the declarations in include/caissa_b200.h); the structure does not: instead of a
// thread per game blocking on one request each, all games of the pool advance in lock step and
// their leaves are evaluated in ONE cb_eval_leaves call per step (one leaf in flight per game,
// exactly like the reference), so the evaluator sees batches of ~`games` positions.
//
// Host threading: the select / expand / backup work of a step is split across worker threads by
// game; the calling thread drives the GPU.  Trees are per-game arenas of 24-byte nodes.
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "../../../include/caissa_b200.h"
#include "qsearch.hpp"
#include "gates.hpp"
#include "chess.hpp"

using namespace cbchess;

namespace {

// Same 32-byte node as the device arenas (mcts_device.cuh: MNode).
struct Node {
    double value_sum = 0.0;   // total_value (f64 as the reference): from the point of view of the player who moved INTO this node
    uint32_t first_child = 0;
    uint32_t visits = 0;
    float prior = 0.0f;
    uint16_t n_children = 0;
    Move move = 0;
    uint8_t state = 0;        // 0 unexpanded, 1 expanded, 2 terminal or resolved by a gate
    int8_t term_value = 0;    // terminal_or_mate_value for the side to move at a state-2 node
    uint8_t n_tact = 0;       // as a parent: children from the capture list
    uint8_t t_done = 0;       // as a parent: how many of them have been visited
    uint8_t trank = 0xFF;     // as a child: place in the parent's tactical order
    uint8_t term_dist = 0;    // terminal_distance of a state-2 node
    uint16_t pad = 0;
};
static_assert(sizeof(Node) == 32, "node layout");

constexpr uint8_t kMateIn1Dist = (uint8_t)1000001u;   // `mate_result.0.unsigned_abs() as u8`, src/mcts/tactical_mcts.rs:701

struct Sample {
    cb_board board;
    float material;
    bool completed = true;   // qsearch_completed of the sample's material search
    std::vector<std::pair<uint16_t, float>> policy;
};

struct Rng {
    uint64_t s;
    uint64_t next() { s ^= s >> 12; s ^= s << 25; s ^= s >> 27; return s * 0x2545F4914F6CDD1Dull; }
    double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
};

// Tier-1 pre-check of the game loop (src/bin/self_play.rs:241-283): does the side to move of `b` win at once?  With the
// reference's gate: a hill in 3 or mate_search(board, 5, exhaustive 2); with the light gates the same at depth 1.
bool root_adjudicated(const cb_board& b, const Move* legal, int n, int tier1, bool koth) {
    if (tier1 >= 2) return (koth && cbchess::koth_center_in_n_host(b, 3) >= 0) || cbchess::mate_search_plies(b, 5, 2) != 0;
    return cbchess::light_gates_hd(b, legal, n, koth) == 1;
}

// dM of a position in pawn units: stand pat or the principal-exchange line (qsearch.hpp)
float material_q(const cb_board& b, int qsearch, bool* completed = nullptr) {
    if (completed) *completed = true;
    if (qsearch == 2) {   // QSearchVariant::Cap1 (src/mcts/tactical_mcts.rs:725-727)
        const cbchess::Cap1Result r = cbchess::forced_cap1(b);
        if (completed) *completed = r.completed;
        return (float)r.score / 100.0f;
    }
    return (float)(qsearch == 1 ? cbchess::principal_exchange_cp(b) : cbchess::pesto_cp(b)) / 100.0f;
}

// What the search needs of cb_selfplay_config / cb_match_config.
struct SearchCfg {
    int sims = 200, material_on = 0, qsearch = 0, koth = 0, tier1_light = 0, shuffle = 0, max_plies = 200;
    int mate_depth = 5, exhaustive_depth = 0, koth_depth = 3;   // the leaf gate of tier1_light = 2 (src/bin/self_play.rs:213)
    double c_puct = 1.414, explore_base = 0.80;
};

struct Game {
    cb_board root_board;
    std::vector<Node> nodes;        // nodes[0] = root
    std::vector<uint64_t> history;  // position keys since the last irreversible move
    std::vector<Sample> samples;
    Rng rng;
    int ply = 0;
    int sims_done = 0;
    // per-step scratch
    std::vector<uint32_t> path;
    cb_board leaf_board;
    Move leaf_moves[kMaxMoves];
    uint16_t leaf_idx[kMaxMoves];
    uint8_t leaf_trank[kMaxMoves];
    int leaf_n = 0, leaf_ntact = 0;
    float leaf_q = 0.0f;
    bool wants_eval = false;

    void reset(uint64_t seed, const cb_board* start = nullptr) {
        if (start) root_board = *start;
        else startpos(&root_board);
        nodes.clear();
        nodes.emplace_back();
        history.clear();
        history.push_back(position_key(root_board));
        samples.clear();
        rng.s = seed | 1;
        ply = 0;
        sims_done = 0;
    }
};

void backup(Game& g, double val) {   // src/mcts/node.rs:403-424: the sign flips every ply
    for (int i = (int)g.path.size() - 1; i >= 0; --i) {
        val = -val;
        Node& n = g.nodes[g.path[i]];
        n.value_sum += val;
        n.visits++;
    }
}

// should_early_terminate, src/mcts/tactical_mcts.rs:524-560 (only with the gates on)
bool early_terminate(const Game& g) {
    const Node& root = g.nodes[0];
    if (root.state != 1) return false;
    bool win1 = false, forced = false, all_visited = true, all_losing = root.n_children > 0;
    for (uint32_t i = 0; i < root.n_children; ++i) {
        const Node& ch = g.nodes[root.first_child + i];
        const bool term = ch.state == 2;
        if (term && ch.term_value < 0) win1 = true;
        if (ch.visits > 0 && ch.value_sum / (double)ch.visits > 0.99) forced = true;
        if (ch.visits == 0) all_visited = false;
        if (!(term && ch.term_value > 0)) all_losing = false;
    }
    if (win1) return true;
    if (forced && all_visited) return true;
    return all_visited && all_losing;
}

void sim_done(Game& g, const SearchCfg& c) {
    if (c.tier1_light && early_terminate(g)) g.sims_done = c.sims;
    else g.sims_done++;
}

// One simulation, phase A: walk down — tactical phase (src/mcts/selection.rs:37-63,107-149), every root child once
// (:65-73), PUCT in f64 with the proven-loss rules (:152-273) — and stop at an unexpanded or decided node.
void select_leaf(Game& g, const SearchCfg& c, std::atomic<long long>& terminal_hits) {
    g.path.clear();
    g.wants_eval = false;
    uint32_t cur = 0;
    cb_board board = g.root_board;
    g.path.push_back(0);
    int depth = 0;
    for (;;) {
        Node& n = g.nodes[cur];
        if (n.state != 1) break;
        uint32_t best = 0xFFFFFFFFu;
        if (n.t_done < n.n_tact) {
            for (uint32_t i = 0; i < n.n_children; ++i)
                if (g.nodes[n.first_child + i].trank == n.t_done) { best = n.first_child + i; break; }
            n.t_done++;
        } else {
            if (depth == 0)
                for (uint32_t i = 0; i < n.n_children; ++i)
                    if (g.nodes[n.first_child + i].visits == 0) { best = n.first_child + i; break; }
            if (best == 0xFFFFFFFFu) {
                const double sq = sqrt((double)(n.visits > 1 ? n.visits : 1));
                double best_u = -INFINITY;
                uint32_t losing = 0xFFFFFFFFu;
                int losing_dist = -1;
                for (uint32_t i = 0; i < n.n_children; ++i) {
                    const Node& ch = g.nodes[n.first_child + i];
                    if (ch.state == 2 && ch.term_value > 0) {
                        if ((int)ch.term_dist > losing_dist) { losing_dist = ch.term_dist; losing = n.first_child + i; }
                        continue;
                    }
                    const double q = ch.visits ? ch.value_sum / (double)ch.visits : 0.0;
                    const double u = q + c.c_puct * (double)ch.prior * sq / (1.0 + (double)ch.visits);
                    if (u > best_u) { best_u = u; best = n.first_child + i; }
                }
                if (best == 0xFFFFFFFFu) best = losing;
            }
        }
        if (best == 0xFFFFFFFFu) break;
        cb_board nb;
        make_move(board, g.nodes[best].move, &nb);
        board = nb;
        cur = best;
        g.path.push_back(cur);
        ++depth;
    }
    Node& leaf = g.nodes[cur];
    if (leaf.state != 0) {  // decided: back its value up again
        backup(g, leaf.state == 2 ? (double)leaf.term_value : 0.0);
        terminal_hits.fetch_add(1, std::memory_order_relaxed);
        sim_done(g, c);
        return;
    }
    g.leaf_n = gen_legal(board, g.leaf_moves, &g.leaf_ntact);
    // mate / stalemate (MctsNode::new_child terminal test, src/mcts/node.rs:162-177), else the light Tier-1 gates
    // (gates.hpp); with the gates on, an opponent king on the hill is tested before anything else
    // (src/mcts/tactical_mcts.rs:619-632).  The root itself is never gated (the game loop adjudicates it first).
    int tv = 2, td = 0;
    const int us = board.w_to_move ? 0 : 1;
    if (g.leaf_n == 0) tv = in_check(board, us) ? -1 : 0;
    else if (c.tier1_light >= 2 && cur != 0)
        tv = cbchess::tier1_gates_host(board, c.koth != 0, c.mate_depth, c.exhaustive_depth, c.koth_depth, &td);
    else if (c.tier1_light && cur != 0) {
        if (c.koth && (board.pieces[(1 - us) * 6 + KING] & kKothCenter)) tv = -1;
        else if (c.koth && cbchess::koth_in_1_hd(board, g.leaf_moves, g.leaf_n)) {
            tv = 1;
            td = (board.pieces[us * 6 + KING] & kKothCenter) ? 0 : 1;
        } else if (cbchess::mate_in_1_hd(board, g.leaf_moves, g.leaf_n)) {
            tv = 1;
            td = kMateIn1Dist;
        }
    }
    if (tv != 2) {
        leaf.state = 2;
        leaf.term_value = (int8_t)tv;
        leaf.term_dist = (uint8_t)td;
        backup(g, (double)tv);
        terminal_hits.fetch_add(1, std::memory_order_relaxed);
        sim_done(g, c);
        return;
    }
    for (int i = 0; i < g.leaf_n; ++i) g.leaf_trank[i] = 0xFF;
    cbchess::tactical_ranks_hd(board, g.leaf_moves, g.leaf_ntact, g.leaf_trank);
    if (c.shuffle)   // SliceRandom::shuffle: for i in (1..len).rev() swap(i, uniform(0..=i))
        for (int i = g.leaf_n - 1; i >= 1; --i) {
            const int j = (int)(g.rng.next() % (uint64_t)(i + 1));
            std::swap(g.leaf_moves[i], g.leaf_moves[j]);
            std::swap(g.leaf_trank[i], g.leaf_trank[j]);
        }
    for (int i = 0; i < g.leaf_n; ++i) g.leaf_idx[i] = (uint16_t)policy_index(g.leaf_moves[i], board.w_to_move != 0);
    g.leaf_board = board;
    // a root asks for its policy only, with dM = 0 (src/mcts/selection.rs:315-332)
    g.leaf_q = (cur == 0 || !c.material_on) ? 0.0f : material_q(board, c.qsearch);
    g.wants_eval = true;
}

// Phase C: expand the evaluated leaf with its priors and back the value up.
void expand_and_backup(Game& g, const SearchCfg& c, const float* priors, float v_final) {
    const uint32_t cur = g.path.back();
    const uint32_t first = (uint32_t)g.nodes.size();
    g.nodes.resize(g.nodes.size() + g.leaf_n);
    Node& leaf = g.nodes[cur];
    leaf.first_child = first;
    leaf.n_children = (uint16_t)g.leaf_n;
    leaf.n_tact = (uint8_t)g.leaf_ntact;
    leaf.t_done = 0;
    leaf.state = 1;
    for (int i = 0; i < g.leaf_n; ++i) {
        Node& ch = g.nodes[first + i];
        ch.move = g.leaf_moves[i];
        ch.prior = priors[i];
        ch.trank = g.leaf_trank[i];
    }
    if (cur == 0) return;   // the root only asked for its priors: no value, no simulation
    backup(g, (double)v_final);  // side-to-move perspective at the leaf
    sim_done(g, c);
}

// Keep the subtree under `child` as the new tree (tree reuse, src/bin/self_play.rs:358-363).
void reroot(Game& g, uint32_t child) {
    std::vector<Node> fresh;
    fresh.reserve(g.nodes.size() / 2 + 1);
    std::vector<std::pair<uint32_t, uint32_t>> todo;  // (old index, new index)
    fresh.push_back(g.nodes[child]);
    todo.emplace_back(child, 0u);
    for (size_t h = 0; h < todo.size(); ++h) {
        const uint32_t oldi = todo[h].first, newi = todo[h].second;
        const Node on = g.nodes[oldi];
        if (on.state == 1) {
            const uint32_t nf = (uint32_t)fresh.size();
            for (uint32_t i = 0; i < on.n_children; ++i) {
                fresh.push_back(g.nodes[on.first_child + i]);
                todo.emplace_back(on.first_child + i, nf + i);
            }
            fresh[newi].first_child = nf;
        }
    }
    g.nodes.swap(fresh);
}

// After the move's simulations: record the sample, choose and play a move (src/bin/self_play.rs:300-363), detect the end
// of the game (:365-388).  Returns 2 when the game goes on, else the result for white (+1, 0, -1).
int play_move(Game& g, const SearchCfg& c) {
    Node& root = g.nodes[0];
    uint32_t total = 0;
    for (uint32_t i = 0; i < root.n_children; ++i) total += g.nodes[root.first_child + i].visits;
    Sample s;
    s.board = g.root_board;
    s.material = material_q(g.root_board, c.qsearch, &s.completed);   // computed whatever material_on says (:315-329)
    if (total > 0)
        for (uint32_t i = 0; i < root.n_children; ++i) {
            const Node& ch = g.nodes[root.first_child + i];
            s.policy.emplace_back((uint16_t)policy_index(ch.move, g.root_board.w_to_move != 0), (float)ch.visits / (float)total);
        }
    g.samples.push_back(std::move(s));
    // proportional to (visits-1) with probability explore_base^(move-1), else greedy: the LAST most-visited child
    // (Iterator::max_by_key, :648-650)
    const int move_number = g.ply / 2 + 1;
    uint32_t pick = root.first_child;
    uint32_t best_v = 0;
    for (uint32_t i = 0; i < root.n_children; ++i)
        if (g.nodes[root.first_child + i].visits >= best_v) { best_v = g.nodes[root.first_child + i].visits; pick = root.first_child + i; }
    if (g.rng.uniform() < pow(c.explore_base, move_number - 1)) {
        uint32_t tot1 = 0;
        for (uint32_t i = 0; i < root.n_children; ++i) {
            const uint32_t v = g.nodes[root.first_child + i].visits;
            tot1 += v > 0 ? v - 1 : 0;
        }
        if (tot1 > 0) {
            const uint32_t thr = (uint32_t)(g.rng.next() % tot1);
            uint32_t cum = 0;
            for (uint32_t i = 0; i < root.n_children; ++i) {
                const uint32_t v = g.nodes[root.first_child + i].visits;
                cum += v > 0 ? v - 1 : 0;
                if (cum > thr) { pick = root.first_child + i; break; }
            }
        }
    }
    const Move mv = g.nodes[pick].move;
    const bool white_moved = g.root_board.w_to_move != 0;
    cb_board nb;
    make_move(g.root_board, mv, &nb);
    if (nb.halfmove == 0) g.history.clear();
    g.root_board = nb;
    g.history.push_back(position_key(nb));
    g.ply++;
    g.sims_done = 0;
    reroot(g, pick);
    int result = 2;
    Move tmp[kMaxMoves];
    if (c.koth && (nb.pieces[(white_moved ? 0 : 6) + KING] & kKothCenter)) result = white_moved ? 1 : -1;
    else {
        const int n_legal = gen_legal(nb, tmp);
        if (n_legal == 0) result = in_check(nb, nb.w_to_move ? 0 : 1) ? (white_moved ? 1 : -1) : 0;
        else {
            int reps = 0;
            for (uint64_t k : g.history) reps += (k == g.history.back());
            if (reps >= 3 || nb.halfmove >= 100 || g.ply > c.max_plies) result = 0;
            // the game goes on: a position where the side to move wins at once is adjudicated before it is
            // searched (Tier-1 pre-check of the self-play loop, src/bin/self_play.rs:241-283, at depth 1)
            else if (c.tier1_light && root_adjudicated(nb, tmp, n_legal, c.tier1_light, c.koth != 0)) result = white_moved ? -1 : 1;
        }
    }
    return result;
}

// Worker pool for the per-step tree work.  Dedicated spinning threads: a step's phases last
// ~100 us, so condition-variable wake-ups (tens of us each) would dominate.
struct Pool {
    int nthreads;
    std::vector<std::thread> threads;
    std::function<void(int, int)> job;
    std::atomic<int> generation{0}, pending{0};
    std::atomic<bool> stop{false};
    int begin = 0, end = 0;

    explicit Pool(int n) : nthreads(n) {
        for (int t = 0; t < n; ++t) threads.emplace_back([this, t] { loop(t); });
    }
    ~Pool() {
        stop.store(true);
        generation.fetch_add(1);
        for (auto& t : threads) t.join();
    }
    static void relax() {
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
    void loop(int t) {
        int seen = 0;
        for (;;) {
            int spins = 0;
            while (generation.load(std::memory_order_acquire) == seen) {
                relax();
                if (++spins > 4096) { std::this_thread::yield(); spins = 0; }
            }
            if (stop.load()) return;
            seen = generation.load(std::memory_order_acquire);
            const int n = end - begin, per = (n + nthreads - 1) / nthreads;
            const int b = begin + t * per, e = b + per < end ? b + per : end;
            if (b < e) job(b, e);
            pending.fetch_sub(1, std::memory_order_acq_rel);
        }
    }
    void run(int b, int e, std::function<void(int, int)> j) {
        job = std::move(j);
        begin = b;
        end = e;
        pending.store(nthreads, std::memory_order_release);
        generation.fetch_add(1, std::memory_order_acq_rel);
        while (pending.load(std::memory_order_acquire) != 0) relax();
    }
};

// One evaluation batch of a half-pool.  In pipelined mode a dedicated thread owns the evaluator
// handle and serves these, so the tree work of one half runs under the other half's evaluation.
struct EvalJob {
    std::vector<cb_board> boards;
    std::vector<float> q, priors, v_logit, v_final, kk;
    std::vector<uint16_t> move_idx;
    std::vector<uint32_t> move_off;
    std::vector<int> game_slot;
    int nb = 0, rc = CB_OK;
    std::atomic<int> state{0};  // 0 idle, 1 submitted, 2 done
    void init(int n) {
        boards.resize(n); q.resize(n); v_logit.resize(n); v_final.resize(n); kk.resize(n);
        priors.resize((size_t)n * kMaxMoves); move_idx.resize((size_t)n * kMaxMoves); move_off.resize(n + 1);
        game_slot.resize(n);
    }
};

}  // namespace

// One training sample in the reference's 5763-f32 record (src/training_data.rs:20-60): 1088 planes, material,
// qsearch flag, z from the mover's point of view, dense 4672 policy.  `rec` is scratch of 5763 floats.
void cbsp_write_sample(FILE* f, const cb_board& b, float material, const std::pair<uint16_t, float>* policy, size_t n_policy,
                       int result_white, std::vector<float>& rec, bool qsearch_completed) {
    rec.assign(5763, 0.0f);
    const bool white = b.w_to_move != 0;
    const int stm = white ? 0 : 1;
    for (int side = 0; side < 2; ++side)
        for (int pt = 0; pt < 6; ++pt)
            for (uint64_t bb = b.pieces[(side == 0 ? stm : 1 - stm) * 6 + pt]; bb; bb &= bb - 1) {
                const int sq = __builtin_ctzll(bb);
                const int row = white ? 7 - (sq >> 3) : (sq >> 3);
                rec[(side * 6 + pt) * 64 + row * 8 + (sq & 7)] = 1.0f;
            }
    if (b.en_passant != 0xFF) {
        const int sq = b.en_passant, row = white ? 7 - (sq >> 3) : (sq >> 3);
        rec[12 * 64 + row * 8 + (sq & 7)] = 1.0f;
    }
    const int ow[4] = {WK, WQ, BK, BQ}, ob[4] = {BK, BQ, WK, WQ};
    for (int i = 0; i < 4; ++i)
        if (b.castling & (white ? ow[i] : ob[i]))
            for (int j = 0; j < 64; ++j) rec[(13 + i) * 64 + j] = 1.0f;
    rec[1088] = material;
    rec[1089] = qsearch_completed ? 1.0f : 0.0f;  // (the stand pat and the exchange line always complete; cap1 can hit its depth in check)
    rec[1090] = result_white == 0 ? 0.0f : ((result_white > 0) == white ? 1.0f : -1.0f);
    for (size_t i = 0; i < n_policy; ++i)
        if (policy[i].first < CB_POLICY_SIZE) rec[1091 + policy[i].first] = policy[i].second;
    fwrite(rec.data(), 4, rec.size(), f);
}

namespace {

double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

namespace {

SearchCfg search_cfg(const cb_selfplay_config* cfg) {
    SearchCfg c;
    c.sims = cfg->sims_per_move;
    c.material_on = cfg->material_on;
    c.qsearch = cfg->qsearch;
    c.koth = cfg->koth;
    c.tier1_light = cfg->tier1_light;
    if (cfg->tier1_mate_depth > 0) c.mate_depth = cfg->tier1_mate_depth;
    c.exhaustive_depth = cfg->tier1_exhaustive_depth;
    if (cfg->tier1_koth_depth > 0) c.koth_depth = cfg->tier1_koth_depth;
    c.shuffle = cfg->shuffle_children;
    c.max_plies = cfg->max_plies > 0 ? cfg->max_plies : 200;
    c.c_puct = cfg->c_puct > 0 ? cfg->c_puct : 1.414;
    c.explore_base = cfg->explore_base > 0 ? cfg->explore_base : 0.80;
    return c;
}

// One move's search of the host driver with a caller-supplied evaluator.
void search_with_callback(Game& g, const SearchCfg& sc, cb_eval_callback eval, void* user) {
    std::atomic<long long> terminal_hits(0);
    float priors[kMaxMoves];
    const Node& root = g.nodes[0];
    if (root.state == 2) return;
    while (g.sims_done < sc.sims) {
        select_leaf(g, sc, terminal_hits);
        if (g.nodes[0].state == 2) return;   // the root itself is mate / stalemate: nothing to search
        if (g.wants_eval) {
            float v = 0.0f;
            eval(user, &g.leaf_board, g.leaf_moves, g.leaf_idx, g.leaf_n, g.leaf_q, sc.material_on, priors, &v);
            expand_and_backup(g, sc, priors, v);
        }
    }
}

}  // namespace

extern "C" {

int cb_chess_startpos(cb_board* out) {
    if (!out) return CB_EINVAL;
    startpos(out);
    return CB_OK;
}

int cb_chess_legal_moves(const cb_board* b, uint16_t* moves, uint16_t* policy_idx) {
    if (!b || !moves) return CB_EINVAL;
    const int n = gen_legal(*b, moves);
    if (policy_idx)
        for (int i = 0; i < n; ++i) policy_idx[i] = (uint16_t)policy_index(moves[i], b->w_to_move != 0);
    return n;
}

int cb_chess_make_move(const cb_board* b, uint16_t move, cb_board* out) {
    if (!b || !out) return CB_EINVAL;
    make_move(*b, move, out);
    return CB_OK;
}

uint64_t cb_chess_perft(const cb_board* b, int depth) { return b ? perft(*b, depth) : 0; }

int cb_principal_exchange(const cb_board* boards, int n, float* q, uint32_t* nodes) {
    if (n < 0 || (n > 0 && (!boards || !q))) return 1;
    for (int i = 0; i < n; ++i) {
        uint32_t nd = 0;
        q[i] = (float)cbchess::principal_exchange_cp(boards[i], &nd) / 100.0f;
        if (nodes) nodes[i] = nd;
    }
    return 0;
}

int cb_cap1_qsearch(const cb_board* boards, int n, float* q, uint8_t* completed, uint32_t* nodes, uint8_t* depth) {
    if (n < 0 || (n > 0 && (!boards || !q))) return 1;
    for (int i = 0; i < n; ++i) {
        const cbchess::Cap1Result r = cbchess::forced_cap1(boards[i]);
        q[i] = (float)r.score / 100.0f;
        if (completed) completed[i] = r.completed ? 1 : 0;
        if (nodes) nodes[i] = r.nodes;
        if (depth) depth[i] = (uint8_t)r.depth;
    }
    return 0;
}

int cb_light_gates(const cb_board* boards, int n, int koth, int8_t* out) {
    if (n < 0 || (n > 0 && (!boards || !out))) return 1;
    Move mv[kMaxMoves];
    for (int i = 0; i < n; ++i) {
        const int nm = gen_legal(boards[i], mv);
        out[i] = (int8_t)cbchess::light_gates_hd(boards[i], mv, nm, koth != 0);
    }
    return 0;
}

// move_flags_hd against make_move_hd + in_check_hd on every pseudo-legal move of n boards: returns the number of
// moves whose flags differ (0 in a sound build); *n_moves (nullable) = moves compared.
long long cb_debug_move_flags(const cb_board* boards, int n, long long* n_moves) {
    if (n < 0 || (n > 0 && !boards)) return -1;
    long long bad = 0, total = 0;
    Move mv[kMaxMoves];
    for (int i = 0; i < n; ++i) {
        const cb_board& b = boards[i];
        const cbchess::MoveCtx c = cbchess::move_ctx(b);
        const int us = b.w_to_move ? 0 : 1;
        const int np = cbchess::gen_pseudo(b, mv);
        for (int k = 0; k < np; ++k) {
            cb_board nb;
            cbchess::make_move_hd(b, mv[k], &nb);
            const bool legal = !cbchess::in_check_hd(nb, us);
            const int want = legal ? (cbchess::in_check_hd(nb, 1 - us) ? 3 : 1) : 0;
            bad += cbchess::move_flags_hd(b, c, mv[k], true) != want;
            bad += cbchess::move_flags_hd(b, c, mv[k], false) != (want & 1);
            ++total;
        }
    }
    if (n_moves) *n_moves = total;
    return bad;
}

int cb_tier1_gates(const cb_board* boards, int n, int koth, int mate_depth, int exhaustive_depth, int koth_depth, int8_t* out,
                   uint8_t* dist) {
    if (n < 0 || (n > 0 && (!boards || !out))) return 1;
    for (int i = 0; i < n; ++i) {
        int d = 0;
        out[i] = (int8_t)cbchess::tier1_gates_host(boards[i], koth != 0, mate_depth, exhaustive_depth, koth_depth, &d);
        if (dist) dist[i] = (uint8_t)d;
    }
    return 0;
}

int cb_mate_search(const cb_board* boards, int n, int max_depth, int exhaustive_depth, int32_t* plies) {
    if (n < 0 || (n > 0 && (!boards || !plies))) return 1;
    for (int i = 0; i < n; ++i) plies[i] = cbchess::mate_search_plies(boards[i], max_depth, exhaustive_depth);
    return 0;
}

int cb_koth_center_in_n(const cb_board* boards, int n, int max_n, int8_t* out) {
    if (n < 0 || (n > 0 && (!boards || !out))) return 1;
    for (int i = 0; i < n; ++i) out[i] = (int8_t)cbchess::koth_center_in_n_host(boards[i], max_n);
    return 0;
}

int cb_debug_host_search(const cb_selfplay_config* cfg, const cb_board* boards, int n, cb_eval_callback eval, void* user,
                         uint16_t* moves, uint32_t* visits, double* value_sums, int32_t* counts, uint32_t* root_visits) {
    if (!cfg || !boards || n < 0 || !eval || !counts) return CB_EINVAL;
    const SearchCfg sc = search_cfg(cfg);
    for (int i = 0; i < n; ++i) {
        Game g;
        g.reset(cfg->seed * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull * (uint64_t)(i + 1), &boards[i]);
        search_with_callback(g, sc, eval, user);
        const Node& root = g.nodes[0];
        const int nc = root.state == 1 ? root.n_children : 0;
        counts[i] = nc;
        if (root_visits) root_visits[i] = root.visits;
        for (int c = 0; c < nc; ++c) {
            const Node& ch = g.nodes[root.first_child + c];
            if (moves) moves[(size_t)i * kMaxMoves + c] = ch.move;
            if (visits) visits[(size_t)i * kMaxMoves + c] = ch.visits;
            if (value_sums) value_sums[(size_t)i * kMaxMoves + c] = ch.value_sum;
        }
    }
    return CB_OK;
}

int cb_debug_host_game(const cb_selfplay_config* cfg, const cb_board* start, uint64_t seed, cb_eval_callback eval, void* user,
                       cb_board* sample_boards, uint16_t* sample_idx, float* sample_prob, int32_t* sample_n, float* sample_material,
                       int max_samples, int* result_white, int* plies) {
    if (!cfg || !start || !eval) return CB_EINVAL;
    const SearchCfg sc = search_cfg(cfg);
    Game g;
    g.reset(seed, start);
    int result = 2;
    Move tmp[kMaxMoves];
    // the game loop's own pre-check of the start position (src/bin/self_play.rs:241-283 at depth 1)
    const int n0 = gen_legal(g.root_board, tmp);
    if (n0 == 0) result = in_check(g.root_board, g.root_board.w_to_move ? 0 : 1) ? (g.root_board.w_to_move ? -1 : 1) : 0;
    else if (sc.tier1_light && root_adjudicated(g.root_board, tmp, n0, sc.tier1_light, sc.koth != 0)) result = g.root_board.w_to_move ? 1 : -1;
    while (result == 2) {
        search_with_callback(g, sc, eval, user);
        result = play_move(g, sc);
    }
    const int ns = (int)g.samples.size();
    // a game that was adjudicated at its start has no samples; otherwise the sample of the last move was recorded
    for (int i = 0; i < ns && i < max_samples; ++i) {
        const Sample& s = g.samples[i];
        if (sample_boards) sample_boards[i] = s.board;
        if (sample_n) sample_n[i] = (int32_t)s.policy.size();
        if (sample_material) sample_material[i] = s.material;
        for (size_t k2 = 0; k2 < s.policy.size(); ++k2) {
            if (sample_idx) sample_idx[(size_t)i * kMaxMoves + k2] = s.policy[k2].first;
            if (sample_prob) sample_prob[(size_t)i * kMaxMoves + k2] = s.policy[k2].second;
        }
    }
    if (result_white) *result_white = result;
    if (plies) *plies = g.ply;
    return ns;
}

int cb_selfplay_run(cb_handle* h, const cb_selfplay_config* cfg, cb_selfplay_stats* out) {
    if (!h || !cfg || !out || cfg->games <= 0 || cfg->sims_per_move <= 0) return CB_EINVAL;
    if (!cb_is_available(h)) return CB_ESTATE;
    if (cfg->pipeline == 3 || cfg->pipeline == 4) return cb_selfplay_device(h, cfg, out);
    memset(out, 0, sizeof(*out));
    if (cfg->mock_eval) return CB_EINVAL;   // the mock evaluator exists on the device path only
    const int G = cfg->games;
    const SearchCfg sc = search_cfg(cfg);
    // the workers spin between phases: keep two cores for the driving thread and the CUDA runtime
    int nthreads = cfg->threads > 0 ? cfg->threads : (int)std::thread::hardware_concurrency() - 2;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > G) nthreads = G;

    const int halves = (cfg->pipeline >= 2 && G >= 2) ? 2 : 1;
    cb_handle* aux = (halves == 2 && cfg->aux_handle && cb_is_available(cfg->aux_handle)) ? cfg->aux_handle : nullptr;
    if (halves == 2 && nthreads > 2) nthreads -= aux ? 2 : 1;   // leave cores to the evaluation thread(s)

    std::vector<Game> games(G);
    for (int i = 0; i < G; ++i) games[i].reset(cfg->seed * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull * (uint64_t)(i + 1));
    Pool pool(nthreads);
    std::atomic<long long> terminal_hits(0), moves_played(0), games_finished(0), samples_written(0);
    std::atomic<long long> wwins(0), bwins(0), draws(0);
    std::mutex file_mu;
    FILE* sample_file = cfg->sample_path ? fopen(cfg->sample_path, "ab") : nullptr;

    // game ranges of the (one or two) half-pools and their evaluation jobs
    const int hb[3] = {0, halves == 2 ? G / 2 : G, G};
    EvalJob jobs[2];
    for (int hf = 0; hf < halves; ++hf) jobs[hf].init(hb[hf + 1] - hb[hf]);

    // Finish a game: assign z to its samples and write them in the reference's 5763-f32 format.
    auto finish_game = [&](Game& g, int result_white /* +1 white won, -1 black won, 0 draw */) {
        if (result_white > 0) wwins++; else if (result_white < 0) bwins++; else draws++;
        if (sample_file) {
            std::vector<float> rec(5763);
            std::lock_guard<std::mutex> l(file_mu);
            for (const Sample& s : g.samples)
                cbsp_write_sample(sample_file, s.board, s.material, s.policy.data(), s.policy.size(), result_white, rec, s.completed);
        }
        samples_written += (long long)g.samples.size();
        games_finished++;
    };

    // After sims_per_move simulations: record the sample, choose and play a move, detect game end.
    auto play = [&](Game& g, int gi) {
        Node& root = g.nodes[0];
        if (root.state != 1 || root.n_children == 0) {  // cannot happen for a live game; restart defensively
            g.reset(g.rng.next());
            return;
        }
        const int result = play_move(g, sc);
        moves_played++;
        if (result != 2) {
            finish_game(g, result);
            g.reset(g.rng.next() ^ ((uint64_t)gi << 32));
        }
    };

    const double t_start = now_s();
    double eval_s = 0.0, tree_s = 0.0;
    long long sims = 0, evals = 0, steps = 0;
    int failed = CB_OK;

    auto eval_job = [&](EvalJob& j, cb_handle* eh) {
        j.rc = CB_OK;
        if (j.nb > 0)
            j.rc = cb_eval_leaves(eh, j.boards.data(), j.q.data(), j.nb, cfg->material_on, j.move_idx.data(), j.move_off.data(),
                                  j.priors.data(), j.v_logit.data(), j.v_final.data(), j.kk.data());
    };
    // phase A for one half-pool: select leaves, pack the evaluation batch
    auto select_and_pack = [&](int hf) {
        EvalJob& j = jobs[hf];
        const int g0 = hb[hf], g1 = hb[hf + 1];
        pool.run(g0, g1, [&](int b, int e) {
            for (int i = b; i < e; ++i) select_leaf(games[i], sc, terminal_hits);
        });
        int nb = 0;
        uint32_t off = 0;
        for (int i = g0; i < g1; ++i) {
            Game& g = games[i];
            if (!g.wants_eval) continue;
            j.boards[nb] = g.leaf_board;
            j.q[nb] = g.leaf_q;
            j.move_off[nb] = off;
            memcpy(&j.move_idx[off], g.leaf_idx, sizeof(uint16_t) * g.leaf_n);
            off += (uint32_t)g.leaf_n;
            j.game_slot[i - g0] = nb;
            nb++;
        }
        j.move_off[nb] = off;
        j.nb = nb;
    };
    // phase C for one half-pool: expand + backup, then play moves where the budget is reached
    auto expand_half = [&](int hf) {
        EvalJob& j = jobs[hf];
        const int g0 = hb[hf], g1 = hb[hf + 1];
        pool.run(g0, g1, [&](int b, int e) {
            for (int i = b; i < e; ++i) {
                Game& g = games[i];
                if (g.wants_eval) {
                    const int s = j.game_slot[i - g0];
                    expand_and_backup(g, sc, &j.priors[j.move_off[s]], j.v_final[s]);
                }
                if (g.sims_done >= cfg->sims_per_move) play(g, i);
            }
        });
        sims += g1 - g0;
        evals += j.nb;
    };
    auto should_stop = [&] {
        if (cfg->max_sims > 0 && sims >= cfg->max_sims) return true;
        if (cfg->max_games > 0 && games_finished.load() >= cfg->max_games) return true;
        if (cfg->max_seconds > 0 && now_s() - t_start >= cfg->max_seconds) return true;
        return false;
    };

    if (halves == 1) {
        while (!should_stop() && failed == CB_OK) {
            const double t0 = now_s();
            select_and_pack(0);
            const double t1 = now_s();
            eval_job(jobs[0], h);
            failed = jobs[0].rc;
            const double t2 = now_s();
            if (failed == CB_OK) expand_half(0);
            tree_s += (t1 - t0) + (now_s() - t2);
            eval_s += t2 - t1;
            steps++;
        }
    } else {
        // software pipeline: the tree work of one half runs while the other half is on the GPU
        // One evaluation thread serving both halves in turn, or — with an auxiliary handle (same weights,
        // own stream and buffers) — one thread and handle per half, so that the copies / synchronisation of
        // one half's call overlap the other half's kernel on the GPU.
        std::atomic<bool> eval_stop(false);
        std::mutex eval_mu;
        auto serve = [&](int first, int stride, cb_handle* eh) {
            int next = first;
            for (;;) {
                EvalJob& j = jobs[next];
                int spins = 0;
                while (j.state.load(std::memory_order_acquire) != 1) {
                    if (eval_stop.load()) return;
                    Pool::relax();
                    if (++spins > 4096) { std::this_thread::yield(); spins = 0; }
                }
                const double t0 = now_s();
                eval_job(j, eh);
                {
                    std::lock_guard<std::mutex> l(eval_mu);
                    eval_s += now_s() - t0;
                }
                j.state.store(2, std::memory_order_release);
                next = (next + stride) & 1;
            }
        };
        std::vector<std::thread> eval_threads;
        if (aux) {
            eval_threads.emplace_back([&] { serve(0, 0, h); });
            eval_threads.emplace_back([&] { serve(1, 0, aux); });
        } else {
            eval_threads.emplace_back([&] { serve(0, 1, h); });
        }
        auto submit = [&](int hf) {
            const double t0 = now_s();
            select_and_pack(hf);
            tree_s += now_s() - t0;
            jobs[hf].state.store(1, std::memory_order_release);
        };
        auto complete = [&](int hf) {
            while (jobs[hf].state.load(std::memory_order_acquire) != 2) Pool::relax();
            jobs[hf].state.store(0, std::memory_order_relaxed);
            if (jobs[hf].rc != CB_OK) { failed = jobs[hf].rc; return; }
            const double t0 = now_s();
            expand_half(hf);
            tree_s += now_s() - t0;
        };
        submit(0);
        int cur = 0;
        long long half_steps = 0;
        for (;;) {
            submit(cur ^ 1);
            complete(cur);
            half_steps++;
            cur ^= 1;
            if (failed != CB_OK || should_stop()) break;
        }
        complete(cur);  // drain the last submitted half
        half_steps++;
        eval_stop.store(true);
        for (auto& t : eval_threads) t.join();
        steps = half_steps / 2;  // a step = one simulation for every game of the pool
    }
    if (failed != CB_OK) {
        if (sample_file) fclose(sample_file);
        return failed;
    }
    if (sample_file) fclose(sample_file);
    out->sims = sims;
    out->nn_evals = evals;
    out->terminal_hits = terminal_hits.load();
    out->moves = moves_played.load();
    out->games_finished = games_finished.load();
    out->samples = samples_written.load();
    out->steps = steps;
    out->white_wins = wwins.load();
    out->black_wins = bwins.load();
    out->draws = draws.load();
    out->seconds = now_s() - t_start;
    out->eval_seconds = eval_s;
    out->tree_seconds = tree_s;
    out->mean_batch = steps ? (double)evals / (double)steps : 0.0;
    return CB_OK;
}

}  // extern "C"
