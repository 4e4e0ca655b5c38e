/* mcts.c — CPU restatement of the reference's tree operations: move-generation ORDER, tactical-first
 * selection, PUCT in f64 with the proven-loss rules, expansion with terminal detection, backup, tree
 * reuse, and the self-play game loop.  TEST INFRASTRUCTURE ONLY (see oracle.h): the product never links it.
 *
 * What it follows (paths relative to the reference checkout):
 *   generation order   src/move_generation.rs:287-328,461-1000, tables src/magic_bitboard.rs:120-265,457-474
 *   tactical list      src/mcts/tactical.rs:155-216 (captures + promotions that are legal, MVV-LVA, stable sort)
 *   selection          src/mcts/selection.rs:19-79 (tactical phase, root force-visit), :152-273 (PUCT, f64,
 *                      proven-loss skip, longest-delay fallback), :281-345 (lazy priors)
 *   node               src/mcts/node.rs:118-206 (terminal test at creation), :403-424 (backup)
 *   search loop        src/mcts/tactical_mcts.rs:458-560 (iteration, early termination), :563-598 (leaf walk),
 *                      :600-810 (leaf value, gates first), :813-836 (expansion), :839-907 (best move),
 *                      :981-1075 (search with reuse), :1081-1099 (reuse_subtree)
 *   game loop          src/bin/self_play.rs:191-445,628-650
 *
 * Stated restrictions: the Tier-1 gates are the reference's at depth 1 (mate_search(board, 1),
 * koth_center_in_n(board, 1)); the random stream is this file's xorshift64* (the reference draws from
 * rand::StdRng / thread_rng, which cannot be reproduced without the crate) — the decision rules applied to
 * the draws are the reference's.  The evaluator is a callback: it receives the leaf, its legal moves in
 * child order and dM, and returns the priors of those moves (src/tensor.rs:184-213) and the fused value. */
#include "oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

enum { P = 0, N = 1, B = 2, R = 3, Q = 4, K = 5 };
enum { WK = 1, WQ = 2, BK = 4, BQ = 8 };
#define CENTER 0x0000001818000000ULL

/* ------------------------------------------------------------------ generation order */
static int type_at(const orc_board* b, int color, int sq) {
    for (int p = 0; p < 6; ++p)
        if (b->pieces[color * 6 + p] >> sq & 1) return p;
    return -1;
}
static int any_at(const orc_board* b, int sq) { return ((b->occ[0] | b->occ[1]) >> sq & 1) != 0; }

/* Is `sq` attacked by colour `by`?  (Board::is_square_attacked; ray walk) */
static int sq_attacked(const orc_board* b, int sq, int by) {
    static const int ndf[8] = {-1, 1, -2, 2, -2, 2, -1, 1}, ndr[8] = {2, 2, 1, 1, -1, -1, -2, -2};
    const int r = sq >> 3, f = sq & 7;
    const uint64_t* pc = &b->pieces[by * 6];
    const int pr = r + (by == 0 ? -1 : 1); /* a pawn of `by` standing one rank behind the square */
    if (pr >= 0 && pr < 8) {
        if (f > 0 && (pc[P] >> (pr * 8 + f - 1) & 1)) return 1;
        if (f < 7 && (pc[P] >> (pr * 8 + f + 1) & 1)) return 1;
    }
    for (int i = 0; i < 8; ++i) {
        const int rr = r + ndr[i], ff = f + ndf[i];
        if (rr >= 0 && rr < 8 && ff >= 0 && ff < 8 && (pc[N] >> (rr * 8 + ff) & 1)) return 1;
    }
    for (int dr = -1; dr <= 1; ++dr)
        for (int df = -1; df <= 1; ++df) {
            if (!dr && !df) continue;
            int rr = r + dr, ff = f + df, first = 1;
            while (rr >= 0 && rr < 8 && ff >= 0 && ff < 8) {
                const int s = rr * 8 + ff;
                if (first && (pc[K] >> s & 1)) return 1;
                first = 0;
                if (any_at(b, s)) {
                    const uint64_t line = (dr == 0 || df == 0) ? (pc[R] | pc[Q]) : (pc[B] | pc[Q]);
                    if (line >> s & 1) return 1;
                    break;
                }
                rr += dr;
                ff += df;
            }
        }
    return 0;
}

/* slider targets from the lowest square up (the reference iterates the bits of the magic attack set) */
static uint64_t slider_set(const orc_board* b, int from, int diag) {
    static const int rdf[4] = {1, -1, 0, 0}, rdr[4] = {0, 0, 1, -1}, bdf[4] = {1, 1, -1, -1}, bdr[4] = {1, -1, 1, -1};
    uint64_t set = 0;
    for (int d = 0; d < 4; ++d) {
        int rr = (from >> 3) + (diag ? bdr[d] : rdr[d]), ff = (from & 7) + (diag ? bdf[d] : rdf[d]);
        while (rr >= 0 && rr < 8 && ff >= 0 && ff < 8) {
            set |= 1ULL << (rr * 8 + ff);
            if (any_at(b, rr * 8 + ff)) break;
            rr += diag ? bdr[d] : rdr[d];
            ff += diag ? bdf[d] : rdf[d];
        }
    }
    return set;
}

typedef struct { orc_move caps[ORC_MAX_MOVES], promos[ORC_MAX_MOVES], quiets[ORC_MAX_MOVES]; int nc, np, nq; } gen_lists;

static void add_promotions(gen_lists* g, int from, int to) { /* src/magic_bitboard.rs:457-474 */
    g->promos[g->np++] = ORC_MV(from, to, Q);
    g->promos[g->np++] = ORC_MV(from, to, R);
    g->promos[g->np++] = ORC_MV(from, to, N);
    g->promos[g->np++] = ORC_MV(from, to, B);
}

static void split_targets(const orc_board* b, gen_lists* g, int from, int to, int us) {
    if (b->occ[1 - us] >> to & 1) g->caps[g->nc++] = ORC_MV(from, to, 0);
    else if (!(b->occ[us] >> to & 1)) g->quiets[g->nq++] = ORC_MV(from, to, 0);
}

/* gen_pseudo_legal_moves: (captures ++ promotions, quiets), src/move_generation.rs:287-328 */
int orc_gen_pseudo_ref(const orc_board* b, orc_move* out, int* n_captures) {
    gen_lists g;
    g.nc = g.np = g.nq = 0;
    const int us = b->w_to_move ? 0 : 1, them = 1 - us;
    const uint64_t* pc = &b->pieces[us * 6];
    /* pawns (:461-571): per pawn the two capture targets in table order (a-file side first), then the pushes
     * in table order — the DOUBLE push first (src/magic_bitboard.rs:248-265) */
    for (uint64_t bb = pc[P]; bb; bb &= bb - 1) {
        const int from = __builtin_ctzll(bb), f = from & 7;
        const int promo = us == 0 ? (from > 47 && from < 56) : (from > 7 && from < 16);
        const int home = us == 0 ? (from > 7 && from < 16) : (from > 47 && from < 56);
        const int fwd = us == 0 ? 8 : -8;
        int ct[2], nct = 0;
        if (us == 0) { if (f > 0) ct[nct++] = from + 7; if (f < 7) ct[nct++] = from + 9; }
        else { if (f > 0) ct[nct++] = from - 9; if (f < 7) ct[nct++] = from - 7; }
        for (int i = 0; i < nct; ++i) {
            const int to = ct[i];
            if (to < 0 || to > 63) continue;
            if ((b->occ[them] >> to & 1) || b->en_passant == to) {
                if (promo) add_promotions(&g, from, to);
                else g.caps[g.nc++] = ORC_MV(from, to, 0);
            }
        }
        int pt[2], npt = 0;
        if (home) pt[npt++] = from + 2 * fwd;
        pt[npt++] = from + fwd;
        for (int i = 0; i < npt; ++i) {
            const int to = pt[i];
            if (to < 0 || to > 63 || any_at(b, to)) continue;
            if (promo) add_promotions(&g, from, to);
            else if (home) { if (!any_at(b, from + fwd)) g.quiets[g.nq++] = ORC_MV(from, to, 0); }
            else g.quiets[g.nq++] = ORC_MV(from, to, 0);
        }
    }
    /* knights (:586-615), targets in the order of init_knight_moves (src/magic_bitboard.rs:160-189) */
    static const int ndf[8] = {-1, 1, -2, 2, -2, 2, -1, 1}, ndr[8] = {2, 2, 1, 1, -1, -1, -2, -2};
    for (uint64_t bb = pc[N]; bb; bb &= bb - 1) {
        const int from = __builtin_ctzll(bb);
        for (int i = 0; i < 8; ++i) {
            const int rr = (from >> 3) + ndr[i], ff = (from & 7) + ndf[i];
            if (rr >= 0 && rr < 8 && ff >= 0 && ff < 8) split_targets(b, &g, from, rr * 8 + ff, us);
        }
    }
    /* bishops, rooks, queens (:839-1000): attack-set bits ascending; a queen's rook set before its bishop set */
    for (uint64_t bb = pc[B]; bb; bb &= bb - 1) {
        const int from = __builtin_ctzll(bb);
        for (uint64_t t = slider_set(b, from, 1); t; t &= t - 1) split_targets(b, &g, from, __builtin_ctzll(t), us);
    }
    for (uint64_t bb = pc[R]; bb; bb &= bb - 1) {
        const int from = __builtin_ctzll(bb);
        for (uint64_t t = slider_set(b, from, 0); t; t &= t - 1) split_targets(b, &g, from, __builtin_ctzll(t), us);
    }
    for (uint64_t bb = pc[Q]; bb; bb &= bb - 1) {
        const int from = __builtin_ctzll(bb);
        for (uint64_t t = slider_set(b, from, 0); t; t &= t - 1) split_targets(b, &g, from, __builtin_ctzll(t), us);
        for (uint64_t t = slider_set(b, from, 1); t; t &= t - 1) split_targets(b, &g, from, __builtin_ctzll(t), us);
    }
    /* king (:630-724): castling first (king side, then queen side), then the steps in the order of
     * init_king_moves (src/magic_bitboard.rs:120-149) */
    {
        const int ksq = us == 0 ? 4 : 60;
        if ((b->castling & (us == 0 ? WK : BK)) && (pc[K] >> ksq & 1) && (pc[R] >> (ksq + 3) & 1) && !any_at(b, ksq + 1) &&
            !any_at(b, ksq + 2) && !sq_attacked(b, ksq, them) && !sq_attacked(b, ksq + 1, them) && !sq_attacked(b, ksq + 2, them))
            g.quiets[g.nq++] = ORC_MV(ksq, ksq + 2, 0);
        if ((b->castling & (us == 0 ? WQ : BQ)) && (pc[K] >> ksq & 1) && (pc[R] >> (ksq - 4) & 1) && !any_at(b, ksq - 1) &&
            !any_at(b, ksq - 2) && !any_at(b, ksq - 3) && !sq_attacked(b, ksq, them) && !sq_attacked(b, ksq - 1, them) &&
            !sq_attacked(b, ksq - 2, them))
            g.quiets[g.nq++] = ORC_MV(ksq, ksq - 2, 0);
    }
    static const int kdf[8] = {-1, -1, 0, 1, 1, 1, 0, -1}, kdr[8] = {0, 1, 1, 1, 0, -1, -1, -1};
    /* the king's captures come after the queens' in the capture list, its quiet steps after the castling moves */
    gen_lists kg;
    kg.nc = kg.np = kg.nq = 0;
    for (uint64_t bb = pc[K]; bb; bb &= bb - 1) {
        const int from = __builtin_ctzll(bb);
        for (int i = 0; i < 8; ++i) {
            const int rr = (from >> 3) + kdr[i], ff = (from & 7) + kdf[i];
            if (rr >= 0 && rr < 8 && ff >= 0 && ff < 8) split_targets(b, &kg, from, rr * 8 + ff, us);
        }
    }
    int n = 0;
    for (int i = 0; i < g.nc; ++i) out[n++] = g.caps[i];
    for (int i = 0; i < kg.nc; ++i) out[n++] = kg.caps[i];
    for (int i = 0; i < g.np; ++i) out[n++] = g.promos[i];
    *n_captures = n;
    for (int i = 0; i < g.nq; ++i) out[n++] = g.quiets[i];
    for (int i = 0; i < kg.nq; ++i) out[n++] = kg.quiets[i];
    return n;
}

static int legal_after(const orc_board* b, orc_move m, orc_board* nb) { /* apply_move_to_board + is_legal, src/board.rs:352 */
    orc_make_move(b, m, nb);
    return !orc_in_check(nb, b->w_to_move ? 0 : 1);
}

/* The legal moves in child order (src/mcts/tactical_mcts.rs:813-836); *n_tactical = how many of them come from the
 * capture list (captures, en passant, all promotions) — the candidates of the tactical phase. */
int orc_gen_legal_ref(const orc_board* b, orc_move* out, int* n_tactical) {
    orc_move ps[ORC_MAX_MOVES];
    int ncap = 0, n = 0, nt = 0;
    const int np = orc_gen_pseudo_ref(b, ps, &ncap);
    orc_board nb;
    for (int i = 0; i < np; ++i)
        if (legal_after(b, ps[i], &nb)) {
            out[n++] = ps[i];
            if (i < ncap) ++nt;
        }
    if (n_tactical) *n_tactical = nt;
    return n;
}

/* calculate_mvv_lva, src/mcts/tactical.rs:179-216 (values P1 N3 B3 R5 Q9 K0; the en-passant victim square is empty) */
static double piece_value(int pt) {
    static const double v[6] = {1.0, 3.0, 3.0, 5.0, 9.0, 0.0};
    return pt < 0 ? 0.0 : v[pt];
}
static int type_any(const orc_board* b, int sq) {
    int t = type_at(b, 0, sq);
    return t >= 0 ? t : type_at(b, 1, sq);
}
double orc_tactical_score(const orc_board* b, orc_move m) {
    double s = piece_value(type_any(b, ORC_TO(m))) * 10.0 - piece_value(type_any(b, ORC_FROM(m)));
    if (ORC_PROMO(m)) s += piece_value(ORC_PROMO(m));
    return s;
}

/* identify_tactical_moves: the order in which a node's tactical children are visited.  order[i] = index (into the
 * legal list `moves`, whose first n_tactical entries are the capture-list moves) of the i-th tactical move. */
void orc_tactical_order(const orc_board* b, const orc_move* moves, int n_tactical, int* order) {
    double sc[ORC_MAX_MOVES];
    for (int i = 0; i < n_tactical; ++i) { order[i] = i; sc[i] = orc_tactical_score(b, moves[i]); }
    for (int i = 1; i < n_tactical; ++i) { /* stable, descending (Vec::sort_by is a stable sort) */
        const int oi = order[i];
        int j = i - 1;
        while (j >= 0 && sc[order[j]] < sc[oi]) { order[j + 1] = order[j]; --j; }
        order[j + 1] = oi;
    }
}

/* ------------------------------------------------------------------ tree */
typedef struct node {
    orc_board state;
    orc_move action;
    uint32_t visits;
    double total_value, total_value_squared;
    int has_term;          /* terminal_or_mate_value.is_some() */
    double term_value;
    int has_dist;
    uint8_t term_dist;
    int is_terminal;
    int has_nn;
    double nn_value;
    struct node* parent;
    struct node** children;
    int n_children;
    int n_tactical_children; /* capture-list moves among the children (before any shuffle they are the first ones) */
    int tact_ready, n_tact;
    int* tact_order;       /* child indices in tactical order */
    uint8_t* tact_done;
    int policy_evaluated, has_raw_policy;
    float* prior;          /* per child, valid when policy_evaluated (move_priorities) or has_raw_policy (cached NN answer) */
    orc_move* pending_moves; /* legal list the evaluator saw (child order), kept until expansion */
    int n_pending, n_pending_tactical;
} node;

struct orc_tree {
    node* root;
    uint64_t rng;
    long evals, terminal_hits, tactical_picks, iterations;
};

static uint64_t rng_next(uint64_t* s) { /* xorshift64* — see the header for why it is not the reference's */
    uint64_t x = *s;
    x ^= x >> 12; x ^= x << 25; x ^= x >> 27;
    *s = x;
    return x * 0x2545F4914F6CDD1DULL;
}
static double rng_uniform(uint64_t* s) { return (double)(rng_next(s) >> 11) * (1.0 / 9007199254740992.0); }

static void mate_or_stalemate(const orc_board* b, int* mate, int* stale) { /* Board::is_checkmate_or_stalemate */
    orc_move mv[ORC_MAX_MOVES];
    const int n = orc_gen_legal_ref(b, mv, NULL);
    const int chk = orc_in_check(b, b->w_to_move ? 0 : 1);
    *mate = n == 0 && chk;
    *stale = n == 0 && !chk;
}

static node* node_new(node* parent, orc_move action, const orc_board* state) { /* new_root / new_child, node.rs:118-206 */
    node* n = (node*)calloc(1, sizeof(node));
    n->state = *state;
    n->action = action;
    n->parent = parent;
    int mate, stale;
    mate_or_stalemate(state, &mate, &stale);
    n->is_terminal = mate || stale;
    if (stale) { n->has_term = 1; n->term_value = 0.0; }
    else if (mate) { n->has_term = 1; n->term_value = -1.0; n->has_dist = 1; n->term_dist = 0; }
    return n;
}

static void node_free(node* n) {
    if (!n) return;
    for (int i = 0; i < n->n_children; ++i) node_free(n->children[i]);
    free(n->children); free(n->tact_order); free(n->tact_done); free(n->prior); free(n->pending_moves);
    free(n);
}

/* evaluate_and_expand_node / ensure_node_expanded: children for every legal move in generation order; with
 * `shuffle` the list the evaluator saw was already permuted (see leaf_moves) */
static void expand(node* n, const orc_mcts_config* cfg, uint64_t* rng) {
    if (n->n_children || n->is_terminal) return;
    orc_move mv[ORC_MAX_MOVES];
    int nm, nt;
    if (n->pending_moves) { nm = n->n_pending; nt = n->n_pending_tactical; memcpy(mv, n->pending_moves, sizeof(orc_move) * nm); }
    else {
        nm = orc_gen_legal_ref(&n->state, mv, &nt);
        if (cfg->shuffle) for (int i = nm - 1; i >= 1; --i) { /* SliceRandom::shuffle: for i in (1..len).rev() swap(i, gen_range(0..=i)) */
            const int j = (int)(rng_next(rng) % (uint64_t)(i + 1));
            const orc_move t = mv[i]; mv[i] = mv[j]; mv[j] = t;
        }
    }
    n->children = (node**)calloc(nm ? nm : 1, sizeof(node*));
    for (int i = 0; i < nm; ++i) {
        orc_board nb;
        orc_make_move(&n->state, mv[i], &nb);
        n->children[i] = node_new(n, mv[i], &nb);
    }
    n->n_children = nm;
    n->n_tactical_children = nt;
}

/* select_unexplored_tactical_move, selection.rs:107-149 */
static node* pick_tactical(orc_tree* t, node* n) {
    if (!n->tact_ready) {
        /* identify_tactical_moves works from the position, in generation order, whatever the children's order */
        orc_move lm[ORC_MAX_MOVES];
        int nt;
        orc_gen_legal_ref(&n->state, lm, &nt);
        int ord[ORC_MAX_MOVES];
        orc_tactical_order(&n->state, lm, nt, ord);
        n->tact_order = (int*)calloc(nt ? nt : 1, sizeof(int));
        n->tact_done = (uint8_t*)calloc(nt ? nt : 1, 1);
        for (int i = 0; i < nt; ++i) { /* find_child_for_move */
            n->tact_order[i] = -1;
            for (int c = 0; c < n->n_children; ++c)
                if (n->children[c]->action == lm[ord[i]]) { n->tact_order[i] = c; break; }
        }
        n->n_tact = nt;
        n->tact_ready = 1;
    }
    for (int i = 0; i < n->n_tact; ++i)
        if (!n->tact_done[i] && n->tact_order[i] >= 0) {
            n->tact_done[i] = 1;
            t->tactical_picks++;
            return n->children[n->tact_order[i]];
        }
    return NULL;
}

static double ucb(const node* c, uint32_t parent_visits, double prior, double cexp) { /* selection.rs:254-273 */
    const double q = c->visits == 0 ? 0.0 : c->total_value / (double)c->visits;
    const double pv = (double)parent_visits;
    const double u = cexp * prior * sqrt(pv > 1.0 ? pv : 1.0) / (1.0 + (double)c->visits);
    return q + u;
}

/* ensure_policy_evaluated, selection.rs:281-345: the cached answer of the node's own evaluation, else a policy-only
 * request with dM = 0 (a root that was never a leaf) */
static void ensure_policy(orc_tree* t, node* n, const orc_mcts_config* cfg, orc_eval_fn eval, void* user) {
    if (n->policy_evaluated) return;
    if (!n->has_raw_policy) {
        orc_move mv[ORC_MAX_MOVES];
        for (int i = 0; i < n->n_children; ++i) mv[i] = n->children[i]->action;
        n->prior = (float*)calloc(n->n_children ? n->n_children : 1, sizeof(float));
        double unused;
        eval(user, &n->state, mv, n->n_children, 0.0f, cfg->material_on, n->prior, &unused);
        t->evals++;
        n->has_raw_policy = 1;
    }
    n->policy_evaluated = 1;
}

/* select_ucb_with_policy, selection.rs:152-251 */
static node* pick_ucb(orc_tree* t, node* n, const orc_mcts_config* cfg, orc_eval_fn eval, void* user) {
    ensure_policy(t, n, cfg, eval, user);
    node *best = NULL, *best_losing = NULL;
    double best_u = -INFINITY;
    uint8_t best_losing_dist = 0;
    for (int i = 0; i < n->n_children; ++i) {
        node* c = n->children[i];
        if (c->has_term && c->term_value > 0.0) {
            const uint8_t d = c->has_dist ? c->term_dist : 0;
            if (!best_losing || d > best_losing_dist) { best_losing_dist = d; best_losing = c; }
            continue;
        }
        const double u = ucb(c, n->visits, (double)n->prior[i], cfg->c_puct);
        if (u > best_u) { best_u = u; best = c; }
    }
    return best ? best : best_losing;
}

/* select_child_with_tactical_priority, selection.rs:19-79 */
static node* pick_child(orc_tree* t, node* n, int depth, const orc_mcts_config* cfg, orc_eval_fn eval, void* user) {
    if (!n->n_children) return NULL;
    node* c = pick_tactical(t, n);
    if (c) return c;
    if (depth == 0)
        for (int i = 0; i < n->n_children; ++i)
            if (n->children[i]->visits == 0) return n->children[i];
    return pick_ucb(t, n, cfg, eval, user);
}

static node* select_leaf(orc_tree* t, const orc_mcts_config* cfg, orc_eval_fn eval, void* user) { /* tactical_mcts.rs:563-598 */
    node* cur = t->root;
    int depth = 0;
    for (;;) {
        if (cur->is_terminal || cur->has_term || !cur->n_children) return cur;
        node* c = pick_child(t, cur, depth, cfg, eval, user);
        if (!c) return cur;
        cur = c;
        ++depth;
    }
}

static float leaf_material(const orc_board* b, const orc_mcts_config* cfg) {
    if (cfg->qsearch == 1) return orc_forced_principal_exchange(b, NULL, NULL);
    if (cfg->qsearch == 2) return orc_forced_cap1(b, NULL, NULL, NULL, NULL); /* QSearchVariant::Cap1, tactical_mcts.rs:725-727 */
    return (float)orc_pst_eval_cp(b) / 100.0f;
}

/* evaluate_leaf_node, tactical_mcts.rs:600-810 (gates at depth 1) */
static double evaluate_leaf(orc_tree* t, node* n, const orc_mcts_config* cfg, orc_eval_fn eval, void* user) {
    if (n->has_term) { t->terminal_hits++; return n->term_value; }
    if (cfg->tier1_light) { /* hill first (:619-647), then the mate (:650-708); depths: 1, or the configured ones */
        const int deep = cfg->tier1_light >= 2;
        int dist = 0;
        const int v = orc_tier1_gates(&n->state, cfg->koth, deep ? (cfg->mate_depth > 0 ? cfg->mate_depth : 5) : 1,
                                      deep ? cfg->exhaustive_mate_depth : 0, deep ? (cfg->koth_depth > 0 ? cfg->koth_depth : 3) : 1, &dist);
        if (v != 2) {
            n->has_term = 1; n->term_value = (double)v; n->has_dist = 1; n->term_dist = (uint8_t)dist;
            if (v < 0) n->is_terminal = 1;
            t->terminal_hits++;
            return (double)v;
        }
    }
    if (!n->has_nn) {
        orc_move mv[ORC_MAX_MOVES];
        int nt;
        const int nm = orc_gen_legal_ref(&n->state, mv, &nt);
        if (cfg->shuffle) for (int i = nm - 1; i >= 1; --i) {
            const int j = (int)(rng_next(&t->rng) % (uint64_t)(i + 1));
            const orc_move x = mv[i]; mv[i] = mv[j]; mv[j] = x;
        }
        const float q = cfg->material_on ? leaf_material(&n->state, cfg) : 0.0f;
        n->prior = (float*)calloc(nm ? nm : 1, sizeof(float));
        double v = 0.0;
        eval(user, &n->state, mv, nm, q, cfg->material_on, n->prior, &v);
        t->evals++;
        n->pending_moves = (orc_move*)malloc(sizeof(orc_move) * (nm ? nm : 1));
        memcpy(n->pending_moves, mv, sizeof(orc_move) * nm);
        n->n_pending = nm;
        n->n_pending_tactical = nt;
        n->has_raw_policy = 1;
        n->nn_value = v;
        n->has_nn = 1;
    }
    return n->nn_value;
}

static void backpropagate(node* n, double value) { /* node.rs:403-424 */
    while (n) {
        n->visits += 1;
        const double reward = -value;
        n->total_value += reward;
        n->total_value_squared += reward * reward;
        value = -value;
        n = n->parent;
    }
}

static int should_early_terminate(const node* root, const orc_mcts_config* cfg) { /* tactical_mcts.rs:524-560 */
    if (!cfg->tier1_light) return 0;
    int win1 = 0, forced = 0, all_visited = 1, all_losing = root->n_children > 0;
    for (int i = 0; i < root->n_children; ++i) {
        const node* c = root->children[i];
        if (c->has_term && c->term_value < -0.5) win1 = 1;
        if (c->visits > 0 && c->total_value / (double)c->visits > 0.99) forced = 1;
        if (c->visits == 0) all_visited = 0;
        if (!(c->has_term && c->term_value > 0.0)) all_losing = 0;
    }
    if (win1) return 1;
    if (forced && all_visited) return 1;
    return all_visited && all_losing;
}

orc_tree* orc_mcts_new(const orc_board* root, uint64_t seed) {
    orc_tree* t = (orc_tree*)calloc(1, sizeof(orc_tree));
    t->root = node_new(NULL, 0, root);
    t->rng = seed | 1;
    return t;
}
void orc_mcts_free(orc_tree* t) {
    if (!t) return;
    node_free(t->root);
    free(t);
}

/* tactical_mcts_search_for_training_with_reuse, tactical_mcts.rs:981-1075.  Returns the iterations run. */
int orc_mcts_search(orc_tree* t, const orc_mcts_config* cfg, orc_eval_fn eval, void* user) {
    node* root = t->root;
    if (!root->is_terminal) { root->has_term = 0; root->has_dist = 0; } /* stale gate values of a reused root, :997-1004 */
    if (root->is_terminal) return 0;
    if (!root->n_children) expand(root, cfg, &t->rng);
    int it = 0;
    for (; it < cfg->sims;) {
        node* leaf = select_leaf(t, cfg, eval, user);
        const double value = evaluate_leaf(t, leaf, cfg, eval, user);
        if (!leaf->is_terminal && leaf->visits == 0 && !leaf->has_term) expand(leaf, cfg, &t->rng);
        backpropagate(leaf, value);
        ++it;
        t->iterations++;
        if (should_early_terminate(root, cfg)) break;
    }
    return it;
}

int orc_mcts_root_children(const orc_tree* t, orc_move* moves, uint32_t* visits, double* total_value) {
    const node* r = t->root;
    for (int i = 0; i < r->n_children; ++i) {
        if (moves) moves[i] = r->children[i]->action;
        if (visits) visits[i] = r->children[i]->visits;
        if (total_value) total_value[i] = r->children[i]->total_value;
    }
    return r->n_children;
}
/* what the reference's tests read off a root child: terminal_or_mate_value / terminal_distance (if set) and whether it was expanded */
int orc_mcts_root_child_info(const orc_tree* t, int i, int* has_term, double* term_value, int* term_dist, int* n_children) {
    const node* r = t->root;
    if (i < 0 || i >= r->n_children) return 0;
    const node* c = r->children[i];
    if (has_term) *has_term = c->has_term;
    if (term_value) *term_value = c->term_value;
    if (term_dist) *term_dist = c->has_dist ? c->term_dist : -1;
    if (n_children) *n_children = c->n_children;
    return 1;
}
void orc_mcts_root_stats(const orc_tree* t, uint32_t* visits, double* total_value, long* evals, long* terminal_hits, long* tactical) {
    if (visits) *visits = t->root->visits;
    if (total_value) *total_value = t->root->total_value;
    if (evals) *evals = t->evals;
    if (terminal_hits) *terminal_hits = t->terminal_hits;
    if (tactical) *tactical = t->tactical_picks;
}

/* select_best_move_from_root, tactical_mcts.rs:839-907 (0 = None) */
orc_move orc_mcts_best_move(const orc_tree* t) {
    const node* r = t->root;
    int have = 0;
    orc_move best = 0;
    uint8_t bd = 0;
    for (int i = 0; i < r->n_children; ++i) { /* a child whose side to move has lost: the fastest such win */
        const node* c = r->children[i];
        if (c->has_term && c->term_value < -0.5) {
            const uint8_t d = c->has_dist ? c->term_dist : 0;
            if (!have || d < bd) { have = 1; bd = d; best = c->action; }
        }
    }
    if (have) return best;
    int all_losing = r->n_children > 0;
    for (int i = 0; i < r->n_children; ++i)
        if (!(r->children[i]->has_term && r->children[i]->term_value > 0.0)) all_losing = 0;
    if (all_losing) {
        for (int i = 0; i < r->n_children; ++i) {
            const node* c = r->children[i];
            const uint8_t d = c->has_dist ? c->term_dist : 0;
            if (!have || d > bd) { have = 1; bd = d; best = c->action; }
        }
        if (have) return best;
    }
    uint32_t bv = 0;
    for (int i = 0; i < r->n_children; ++i)
        if (r->children[i]->visits > bv) { bv = r->children[i]->visits; best = r->children[i]->action; }
    return best;
}

/* reuse_subtree, tactical_mcts.rs:1081-1099: 1 when the move is a child of the root */
int orc_mcts_advance(orc_tree* t, orc_move played) {
    node* r = t->root;
    for (int i = 0; i < r->n_children; ++i)
        if (r->children[i]->action == played) {
            node* c = r->children[i];
            r->children[i] = NULL;
            c->parent = NULL;
            node_free(r);
            t->root = c;
            return 1;
        }
    return 0;
}

/* ------------------------------------------------------------------ mock evaluators (src/mcts/inference_server.rs:103-122) */
/* new_mock_biased(v_logit): all-zero policy => policy_to_move_priors falls back to uniform 1/len; k = 0.5; the value
 * fusion of tactical_mcts.rs:762-765,787-790 in f64.  user = const float[2] {v_logit, k}: k = 0.5 is the mock server,
 * {0, 0.326} the classical fallback of :776-785 (no network: uniform priors, tanh(0.326 * dM)). */
void orc_eval_mock(void* user, const orc_board* b, const orc_move* moves, int n, float q, int material_on, float* priors, double* value) {
    (void)b; (void)moves;
    const float* u = (const float*)user;
    for (int i = 0; i < n; ++i) priors[i] = 1.0f / (float)n;
    *value = orc_value_fuse(u[0], u[1], q, material_on);
}

/* ------------------------------------------------------------------ game loop (src/bin/self_play.rs:191-445) */
static uint64_t pos_key(const orc_board* b) { /* repetition identity: placement, side, castling, en passant */
    uint64_t h = 0xcbf29ce484222325ULL;
    for (int i = 0; i < 12; ++i) { h ^= b->pieces[i]; h *= 0x100000001b3ULL; h ^= h >> 29; }
    h ^= (uint64_t)b->w_to_move | (uint64_t)b->castling << 8 | (uint64_t)b->en_passant << 16;
    h *= 0x100000001b3ULL;
    return h ^ (h >> 32);
}

int orc_play_game(const orc_board* start, const orc_mcts_config* cfg, uint64_t seed, orc_eval_fn eval, void* user,
                  orc_game_sample* samples, int max_samples, int* result_white, int* plies) {
    orc_tree* t = orc_mcts_new(start, seed);
    orc_board board = *start;
    uint64_t hist[1024];
    int nh = 0, ns = 0, move_count = 0, early = 0, early_score = 0;
    hist[nh++] = pos_key(&board);
    const int max_plies = cfg->max_plies > 0 ? cfg->max_plies : 200;
    for (;;) {
        if (cfg->tier1_light) { /* Tier-1 pre-check, :241-283: hill in 3 or mate_search(board, 5, exhaustive 2) - the side to
                                   move wins at once; tier1_light = 1 runs both at depth 1 */
            const int deep = cfg->tier1_light >= 2;
            if ((cfg->koth && orc_koth_center_in_n(&board, deep ? 3 : 1, NULL) >= 0) ||
                orc_mate_search(&board, deep ? 5 : 1, deep ? 2 : 0, NULL, NULL) != 0) {
                early = 1;
                early_score = board.w_to_move ? 1 : -1;
                break;
            }
        }
        orc_mcts_search(t, cfg, eval, user);
        if (!orc_mcts_best_move(t)) break;
        orc_move mv[ORC_MAX_MOVES];
        uint32_t vis[ORC_MAX_MOVES];
        const int nc = orc_mcts_root_children(t, mv, vis, NULL);
        uint32_t total = 0;
        for (int i = 0; i < nc; ++i) total += vis[i];
        if (ns < max_samples) {
            orc_game_sample* s = &samples[ns];
            s->board = board;
            s->n = 0;
            if (total > 0)
                for (int i = 0; i < nc; ++i) {
                    s->idx[i] = (uint16_t)orc_move_to_index_stm(mv[i], board.w_to_move);
                    s->prob[i] = (float)vis[i] / (float)total;
                    s->n = i + 1;
                }
            s->material = leaf_material(&board, cfg); /* computed whatever enable_material says, :315-329 */
            s->w_to_move = board.w_to_move;
        }
        ++ns;
        /* proportional to visits-1 with probability explore_base^(move-1), else greedy (:343-351,628-650) */
        const int move_number = move_count / 2 + 1;
        orc_move chosen = 0;
        int greedy = 1;
        if (rng_uniform(&t->rng) < pow(cfg->explore_base, (double)(move_number - 1))) {
            uint32_t tot1 = 0;
            for (int i = 0; i < nc; ++i) tot1 += vis[i] ? vis[i] - 1 : 0;
            if (tot1 > 0) {
                const uint32_t thr = (uint32_t)(rng_next(&t->rng) % tot1);
                uint32_t cum = 0;
                chosen = mv[nc - 1];
                for (int i = 0; i < nc; ++i) {
                    cum += vis[i] ? vis[i] - 1 : 0;
                    if (cum > thr) { chosen = mv[i]; break; }
                }
                greedy = 0;
            }
        }
        if (greedy) { /* Iterator::max_by_key keeps the LAST maximum */
            uint32_t bv = 0;
            int have = 0;
            for (int i = 0; i < nc; ++i)
                if (!have || vis[i] >= bv) { bv = vis[i]; chosen = mv[i]; have = 1; }
        }
        orc_mcts_advance(t, chosen);
        orc_board nb;
        orc_make_move(&board, chosen, &nb);
        board = nb;
        if (board.halfmove == 0) nh = 0;
        if (nh < 1024) hist[nh++] = pos_key(&board);
        ++move_count;
        if (cfg->koth && ((board.pieces[K] | board.pieces[6 + K]) & CENTER)) break;
        int reps = 0;
        for (int i = 0; i < nh; ++i) reps += hist[i] == hist[nh - 1];
        if (reps >= 3) break;
        if (board.halfmove >= 100) break;
        if (move_count > max_plies) break;
        int mate, stale;
        mate_or_stalemate(&board, &mate, &stale);
        if (mate || stale) break;
    }
    int mate, stale, res = 0;
    mate_or_stalemate(&board, &mate, &stale);
    if (early) res = early_score;
    else if (cfg->koth && (board.pieces[K] & CENTER)) res = 1;
    else if (cfg->koth && (board.pieces[6 + K] & CENTER)) res = -1;
    else if (mate) res = board.w_to_move ? -1 : 1;
    const int n_out = ns < max_samples ? ns : max_samples;
    for (int i = 0; i < n_out; ++i) samples[i].value_target = (float)(samples[i].w_to_move ? res : -res);
    if (result_white) *result_white = res;
    if (plies) *plies = move_count;
    orc_mcts_free(t);
    return ns;
}
