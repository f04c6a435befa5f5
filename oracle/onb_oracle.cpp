// =====================================================================================
// oracle/onb_oracle.cpp  --  TEST INFRASTRUCTURE ONLY (never linked into / imported by the product).
//
// CPU restatement of the cyoq/onitama-alphazero self-play hot path, written in the same
// algorithmic shape as the Rust reference (heap move vectors, 25x25 bit scans, State clone per
// playout, per-tree node arena with child vectors, f64 PUCT). Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load this library, and only as the
// checker or as the timed CPU baseline.
//
// PARITY STATUS
//   * rules, move generation, make_move, terminal detection, expansion order: PINNED against the
//     reference's own known-answer tests (onitama-game/src/common/mod.rs:82-134,
//     onitama-game/src/game/state.rs:420-889, onitama-game/src/ai/mcts/mcts_arena.rs:403-457),
//     re-expressed in tests/test_oracle_golden.py.
//   * PUCT arena, encoder, self-play loop: the reference has NO tests for them and the Rust
//     workspace cannot be built in this image (no cargo/rustc) -> "parity unpinned": this file is a
//     line-by-line restatement; tests cross-check it against an independent pure-Python
//     restatement (tests/golden/gen_golden.py) and the SURVEY.md Appendix-A values.
//   * RNG: the reference uses rand::thread_rng (OS seeded, not reproducible). The counter-based RNG
//     below is this project's own definition (restated independently from include/onb.h).
//
// All file:line citations are relative to the reference repository root.
// =====================================================================================
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

namespace orc {

// ------------------------------------------------------------------ bit helpers
// onitama-game/src/common/mod.rs:2-4  -- square n <-> bit (31 - n), MSB first.
static inline uint32_t get_bit(uint32_t x, unsigned n) { return (x >> (31u - n)) & 1u; }
// onitama-game/src/common/mod.rs:29-43
static inline void set_bit(uint32_t& v, unsigned pos) { v |= 1u << (31u - pos); }
static inline void clear_bit(uint32_t& v, unsigned pos) { v &= ~(1u << (31u - pos)); }

enum Color : int { RED = 0, BLUE = 1 };                     // player_color.rs:6-9
enum Piece : int { PAWN = 0, KING = 1 };                    // piece.rs:6-9
enum MoveResult : int { CAPTURE = 0, RED_WIN = 1, BLUE_WIN = 2, IN_PROGRESS = 3 };  // move_result.rs:4-9
static inline bool is_win(int r) { return r == RED_WIN || r == BLUE_WIN; }          // move_result.rs:13-15
static inline int enemy(int c) { return c ^ 1; }

// ------------------------------------------------------------------ cards
// onitama-game/src/game/card.rs:17-463. positions = Red orientation, mirror = Blue orientation,
// colour stamp decides the first mover when the card is the neutral one (game_state.rs:36).
struct Card {
    uint32_t positions, mirror;
    int player_color;
    int index;
};
static const Card ORIGINAL_CARDS[16] = {
    {0x20004000u, 0x01000200u, BLUE, 0},   // Tiger     card.rs:17-40
    {0x0440A000u, 0x02811000u, RED, 1},    // Dragon    card.rs:42-65
    {0x02202000u, 0x02022000u, RED, 2},    // Frog      card.rs:67-90
    {0x00828000u, 0x00A08000u, RED, 3},    // Rabbit    card.rs:92-115
    {0x01220000u, 0x00224000u, BLUE, 4},   // Crab      card.rs:117-144
    {0x02940000u, 0x0014A000u, RED, 5},    // Elephant  card.rs:146-173
    {0x02142000u, 0x02142000u, BLUE, 6},   // Goose     card.rs:175-202
    {0x00948000u, 0x00948000u, RED, 7},    // Rooster   card.rs:204-231
    {0x0280A000u, 0x0280A000u, BLUE, 8},   // Monkey    card.rs:233-260
    {0x02804000u, 0x0100A000u, RED, 9},    // Mantis    card.rs:262-289
    {0x0100A000u, 0x02804000u, BLUE, 10},  // Crane = Mantis swapped   card.rs:291-318
    {0x01104000u, 0x01044000u, RED, 11},   // Horse     card.rs:320-347
    {0x01044000u, 0x01104000u, BLUE, 12},  // Ox = Horse swapped       card.rs:349-376
    {0x01140000u, 0x00144000u, RED, 13},   // Boar      card.rs:378-405
    {0x02048000u, 0x00902000u, BLUE, 14},  // Eel       card.rs:407-434
    {0x00902000u, 0x02048000u, RED, 15},   // Cobra = Eel swapped      card.rs:436-463
};

// card.rs:487-517
static const uint32_t FILE_A = 0x84210800u, FILE_E = 0x08421080u, FILE_AB = 0xC6318C00u, FILE_DE = 0x18C63180u;

static uint32_t ATTACK_MAPS[2][16][25];

// card.rs:562-604: centre-anchored pattern shifted by +-n, wrapped files masked by n % 5.
static void generate_attack_maps_for_card(uint32_t card, uint32_t out[25]) {
    for (int i = 0; i < 25; ++i) out[i] = 0;
    out[12] = card;
    for (int n = 1; n < 13; ++n) {
        uint32_t left = (card << n) & 0xFFFFFF80u;
        uint32_t right = (card >> n) & 0xFFFFFF80u;
        switch (n % 5) {
            case 1: left &= ~FILE_E; right &= ~FILE_A; break;
            case 2: left &= ~FILE_DE; right &= ~FILE_AB; break;
            case 3: left &= ~FILE_AB; right &= ~FILE_DE; break;
            case 4: left &= ~FILE_A; right &= ~FILE_E; break;
            default: break;
        }
        out[12 - n] = left;
        out[12 + n] = right;
    }
}
// card.rs:520-541
static void generate_attack_maps() {
    for (int player = 0; player < 2; ++player)
        for (int ci = 0; ci < 16; ++ci) {
            const Card& c = ORIGINAL_CARDS[ci];
            generate_attack_maps_for_card(player == BLUE ? c.mirror : c.positions, ATTACK_MAPS[player][c.index]);
        }
}
static struct Init { Init() { generate_attack_maps(); } } g_init;

// ------------------------------------------------------------------ deck / state
// deck.rs:14-18
static const int RED_CARD1 = 0, RED_CARD2 = 1, BLUE_CARD1 = 2, BLUE_CARD2 = 3, NEUTRAL = 4;
// state.rs:24-49
static const uint32_t RED_KING_SP = 0x00000200u, BLUE_KING_SP = 0x20000000u;
static const uint32_t BLUE_PAWNS_SP = 0xD8000000u, RED_PAWNS_SP = 0x00000D80u;
static const unsigned BLUE_TEMPLE = 2, RED_TEMPLE = 22;

struct Move {  // move.rs:21-25
    uint32_t from, to;
    int piece;
};
struct DoneMove {  // done_move.rs:4-7
    Move mov;
    unsigned used_card_idx;
};

struct State {  // state.rs:51-56 (deck holds card indices; Card is looked up in ORIGINAL_CARDS)
    uint8_t deck[5];
    uint32_t kings[2];
    uint32_t pawns[2];

    static State with_deck(const uint8_t d[5]) {  // state.rs:67-73
        State s;
        for (int i = 0; i < 5; ++i) s.deck[i] = d[i];
        s.kings[RED] = RED_KING_SP; s.kings[BLUE] = BLUE_KING_SP;
        s.pawns[RED] = RED_PAWNS_SP; s.pawns[BLUE] = BLUE_PAWNS_SP;
        return s;
    }
    void rotate(unsigned idx) { std::swap(deck[idx], deck[NEUTRAL]); }  // deck.rs:87-90

    int current_state() const {  // state.rs:120-134
        int r = IN_PROGRESS;
        if (kings[RED] == 0 || kings[BLUE] == RED_KING_SP) r = BLUE_WIN;
        else if (kings[BLUE] == 0 || kings[RED] == BLUE_KING_SP) r = RED_WIN;
        return r;
    }
    int pass(unsigned card_idx) { rotate(card_idx); return IN_PROGRESS; }  // state.rs:139-142

    // state.rs:145-202 -- no legality check.
    int make_move(const Move& mov, int player_color, unsigned used_card_idx) {
        unsigned from = mov.from, to = mov.to;
        int result = IN_PROGRESS;
        if (mov.piece == PAWN) clear_bit(pawns[player_color], from);
        else clear_bit(kings[player_color], from);
        int en = enemy(player_color);
        uint32_t enemy_pawn = get_bit(pawns[en], to), enemy_king = get_bit(kings[en], to);
        if (enemy_pawn == 1) {
            clear_bit(pawns[en], to);
            result = CAPTURE;
        } else if (enemy_king == 1) {
            clear_bit(kings[en], to);
            result = player_color == RED ? RED_WIN : BLUE_WIN;
        }
        if (mov.piece == PAWN) set_bit(pawns[player_color], to);
        else set_bit(kings[player_color], to);
        if (mov.piece == KING) {
            if (player_color == RED && to == BLUE_TEMPLE) result = RED_WIN;
            else if (player_color == BLUE && to == RED_TEMPLE) result = BLUE_WIN;
        }
        rotate(used_card_idx);
        return result;
    }

    // state.rs:323-378 -- from ascending, to ascending.
    std::vector<Move> generate_legal_moves(int player_color, int card_index) const {
        std::vector<Move> result;
        uint32_t p = pawns[player_color], k = kings[player_color];
        for (unsigned n = 0; n < 25; ++n) {
            uint32_t pawn_bit = get_bit(p, n), king_bit = get_bit(k, n);
            if (pawn_bit == 0 && king_bit == 0) continue;
            uint32_t attack_map = ATTACK_MAPS[player_color][card_index][n];
            uint32_t map;
            int figure;
            if (pawn_bit == 1) { map = ((attack_map | p) & ~p) & ~k; figure = PAWN; }
            else { map = ((attack_map | k) & ~k) & ~p; figure = KING; }
            for (unsigned i = 0; i < 25; ++i) {
                if (get_bit(map, i) == 0) continue;
                result.push_back(Move{n, i, figure});
            }
        }
        return result;
    }
    // state.rs:301-310 + deck.rs:48-53 -- slot ascending.
    std::vector<std::pair<unsigned, Move>> generate_all_legal_moves(int player_color) const {
        unsigned cards[2] = {player_color == RED ? 0u : 2u, player_color == RED ? 1u : 3u};
        std::vector<std::pair<unsigned, Move>> result;
        for (unsigned c : cards) {
            std::vector<Move> moves = generate_legal_moves(player_color, deck[c]);
            for (const Move& m : moves) result.emplace_back(c, m);
        }
        return result;
    }
};

// ------------------------------------------------------------------ counter RNG (project-defined)
static inline uint64_t mix64(uint64_t z) {
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27; z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return z;
}
static inline uint32_t rand_u32(uint64_t seed, uint64_t game, uint32_t step, uint32_t draw) {
    uint64_t h = mix64(seed + 0x9E3779B97F4A7C15ull * (game + 1));
    h = mix64(h ^ (((uint64_t)step << 32) | draw));
    return (uint32_t)(h >> 32);
}
static inline uint32_t rand_index(uint32_t r, uint32_t n) { return (uint32_t)(((uint64_t)r * n) >> 32); }
enum Draw : uint32_t { DRAW_MOVE = 0, DRAW_PASS = 1, DRAW_AGENT_SLOT = 2, DRAW_AGENT_MOVE = 3, DRAW_DEAL = 8 };

// first 5 of a Fisher-Yates shuffle of 0..15 (deck.rs:139-151 draws a uniformly random deal).
static void deal(uint64_t seed, uint64_t game, uint32_t epoch, uint8_t out[5]) {
    uint8_t ids[16];
    for (int i = 0; i < 16; ++i) ids[i] = (uint8_t)i;
    for (uint32_t i = 0; i < 5; ++i) {
        uint32_t j = i + rand_index(rand_u32(seed, game, epoch, DRAW_DEAL + i), 16 - i);
        std::swap(ids[i], ids[j]);
    }
    for (int i = 0; i < 5; ++i) out[i] = ids[i];
}

// ------------------------------------------------------------------ root exploration noise (train mode)
// The reference adds Dirichlet(0.03) noise at the root in train mode, drawn from thread_rng with a FRESH Dirichlet sample per
// uct() call (alphazero_mcts/mcts_arena.rs:186-202): not reproducible, statistical parity only. Project-defined restatement:
// the noise of child i in one uct() call is component i of a fresh Dirichlet(alpha; k) sample (rand_distr 0.4.3 normalises k
// independent Gamma(alpha) draws), i.e. a Beta(alpha, (k-1) alpha) variate, sampled from the counter RNG.
static inline float noise_uniform(uint64_t key, uint32_t step, uint32_t code) {
    const uint32_t r = (uint32_t)(mix64(key ^ (((uint64_t)step << 32) | code)) >> 32);
    return ((float)(r >> 8) + 0.5f) * (1.0f / 16777216.0f);
}
// Beta(alpha, (k-1) alpha) by Joehnk's method in log space, single precision (see the comment in csrc/onb_mcts.cu).
static float noise_beta(uint32_t k, float alpha, uint64_t key, uint32_t step, uint32_t sample) {
    const float inv_a = 1.0f / alpha, inv_b = 1.0f / ((float)(k - 1) * alpha);
    float lx = 0.f, ly = 0.f;
    for (uint32_t attempt = 0; attempt < 16; ++attempt) {
        const uint32_t code = 0x80000000u | (sample << 8) | (attempt << 2);
        lx = std::log(noise_uniform(key, step, code)) * inv_a;
        ly = std::log(noise_uniform(key, step, code | 1u)) * inv_b;
        if (std::exp(lx) + std::exp(ly) <= 1.0f) break;
    }
    return 1.0f / (1.0f + std::exp(ly - lx));
}

// ------------------------------------------------------------------ boundary structs (C layout)
}  // namespace orc

extern "C" {
// Same memory layout as onb_state (include/onb.h), restated here on purpose.
struct orc_state {
    uint32_t pawns[2];
    uint32_t kings[2];
    uint8_t cards[5];
    uint8_t side;    // side to move, 0 = Red, 1 = Blue
    uint8_t result;  // 0 in progress, 1 Red won, 2 Blue won
    uint8_t flags;   // bit0: at least one pass (zero-legal-move turn) happened since reset
};
}

namespace orc {

static State to_state(const orc_state& s) {
    State st;
    for (int i = 0; i < 5; ++i) st.deck[i] = s.cards[i];
    st.pawns[0] = s.pawns[0]; st.pawns[1] = s.pawns[1];
    st.kings[0] = s.kings[0]; st.kings[1] = s.kings[1];
    return st;
}
static void from_state(const State& st, orc_state& s) {
    for (int i = 0; i < 5; ++i) s.cards[i] = st.deck[i];
    s.pawns[0] = st.pawns[0]; s.pawns[1] = st.pawns[1];
    s.kings[0] = st.kings[0]; s.kings[1] = st.kings[1];
}
// action code: to | from<<5 | used_card_idx<<10 | piece<<12 | pass<<13
static inline uint16_t encode_action(unsigned card_idx, const Move& m) {
    return (uint16_t)(m.to | (m.from << 5) | (card_idx << 10) | ((unsigned)m.piece << 12));
}
static inline uint16_t encode_pass(unsigned card_idx) { return (uint16_t)((card_idx << 10) | (1u << 13)); }
static inline DoneMove decode_action(uint16_t a) {
    DoneMove d;
    d.mov.to = a & 31u; d.mov.from = (a >> 5) & 31u; d.used_card_idx = (a >> 10) & 3u; d.mov.piece = (a >> 12) & 1;
    return d;
}
static inline uint8_t result_code(int move_result) { return move_result == RED_WIN ? 1 : move_result == BLUE_WIN ? 2 : 0; }

static void new_game(orc_state& g, uint64_t seed, uint64_t game, uint32_t epoch, const uint8_t* fixed_deck) {
    uint8_t d[5];
    if (fixed_deck) memcpy(d, fixed_deck, 5); else deal(seed, game, epoch, d);
    State st = State::with_deck(d);
    from_state(st, g);
    g.side = (uint8_t)ORIGINAL_CARDS[d[NEUTRAL]].player_color;  // game_state.rs:34-41, train.rs:49
    g.result = 0;
    g.flags = 0;
}

// policy 0: mcts/mcts_arena.rs:190-241 ("simulate": uniform over all legal moves, pass if none)
// policy 1: ai/random.rs:12-43 (Random agent incl. its fabricated move and Blue card-slot bug)
static uint16_t choose_random_action(const State& st, int side, int policy, uint64_t seed, uint64_t game, uint32_t step) {
    if (policy == 0) {
        auto moves = st.generate_all_legal_moves(side);
        if (moves.empty()) {
            unsigned base = side == RED ? 0u : 2u;  // mcts_arena.rs:213-216
            return encode_pass(base + rand_index(rand_u32(seed, game, step, DRAW_PASS), 2));
        }
        auto& pick = moves[rand_index(rand_u32(seed, game, step, DRAW_MOVE), (uint32_t)moves.size())];
        return encode_action(pick.first, pick.second);
    }
    unsigned card_idx = rand_index(rand_u32(seed, game, step, DRAW_AGENT_SLOT), 2);  // random.rs:17
    unsigned slot = (side == RED ? 0u : 2u) + card_idx;                                // random.rs:16,18
    auto moves = st.generate_legal_moves(side, st.deck[slot]);
    Move mv;
    if (!moves.empty()) mv = moves[rand_index(rand_u32(seed, game, step, DRAW_AGENT_MOVE), (uint32_t)moves.size())];
    else mv = Move{0, 5, PAWN};                                                        // random.rs:26-34
    return encode_action(card_idx, mv);                                                // random.rs:39 (no +2 for Blue)
}

// one env transition on the boundary struct; returns MoveResult
static int apply_action(orc_state& g, uint16_t a) {
    State st = to_state(g);
    int side = g.side, r;
    if (a & (1u << 13)) { r = st.pass((a >> 10) & 3u); g.flags |= 1; }
    else { DoneMove d = decode_action(a); r = st.make_move(d.mov, side, d.used_card_idx); }
    from_state(st, g);
    g.side = (uint8_t)enemy(side);
    g.result = result_code(r);
    return r;
}

// alphazero-training/src/common.rs:26-80
static void encode_planes(const State& st, int player_color, float* out /*525*/) {
    for (int i = 0; i < 525; ++i) out[i] = 0.f;
    if (player_color == BLUE) for (int i = 0; i < 25; ++i) out[20 * 25 + i] = 1.f;
    const uint32_t boards[4] = {st.pawns[RED], st.kings[RED], st.pawns[BLUE], st.kings[BLUE]};
    for (int p = 0; p < 4; ++p)
        for (unsigned i = 0; i < 25; ++i) out[p * 25 + i] = (float)((boards[p] >> (31 - i)) & 1u);  // get_bit_array, common/mod.rs:69-75
    unsigned base = player_color == RED ? 0u : 2u;
    int c1 = st.deck[base], c2 = st.deck[base + 1];
    for (int i = 0; i < 25; ++i) out[(c1 + 4) * 25 + i] = 1.f;
    for (int i = 0; i < 25; ++i) out[(c2 + 4) * 25 + i] = 1.f;
}

// policy-shaped legal mask: word s = union of `to` squares of own hand slot s, MSB-first like the boards.
static void legal_mask(const State& st, int side, uint32_t out[2]) {
    out[0] = out[1] = 0;
    for (auto& cm : st.generate_all_legal_moves(side)) out[cm.first & 1u] |= 1u << (31u - cm.second.to);
}

// ------------------------------------------------------------------ perft
struct PerftOut { uint64_t* nodes; uint64_t* wins; uint64_t* zero; };
// nodes[d-1] = positions reached by exactly d plies (a line stops at a win), wins[d-1] = those that are wins,
// zero[d-1] = non-terminal positions at depth d-1 with no legal move (not expanded; the reference would panic).
static void perft_rec(const State& st, int side, int depth, int max_depth, PerftOut& o) {
    auto moves = st.generate_all_legal_moves(side);
    if (moves.empty()) { o.zero[depth] += 1; return; }
    for (auto& cm : moves) {
        State child = st;
        int r = child.make_move(cm.second, side, cm.first);
        o.nodes[depth] += 1;
        if (is_win(r)) { o.wins[depth] += 1; continue; }
        if (depth + 1 < max_depth) perft_rec(child, enemy(side), depth + 1, max_depth, o);
    }
}

// ------------------------------------------------------------------ AlphaZero PUCT arena
// alphazero_mcts/mcts_arena.rs:355-402
struct MctsNode {
    int64_t parent;  // -1 = None
    std::vector<uint32_t> children;
    uint32_t idx;
    bool has_mov;
    DoneMove mov;
    bool is_pass;  // project-defined extension for zero-legal-move nodes
    uint32_t visits;
    double reward, winrate;
    bool is_terminal, is_expanded;
    int player_color;
    double probability;
    void update(double r) { visits += 1; reward += r; winrate = reward / (double)visits; }  // :398-402
};

typedef void (*eval_fn)(const float* planes, float* policy50, float* value, void* user);

// f64::total_cmp key (Rust core): flips the magnitude bits of negatives so an i64 compare is a total order.
static inline int64_t total_key(double x) {
    int64_t b; memcpy(&b, &x, 8);
    b ^= (int64_t)((uint64_t)(b >> 63) >> 1);
    return b;
}

// alphazero_mcts/mod.rs:45-53
static double reward_fn(int move_result, int reward_color) {
    if (reward_color == RED) return move_result == RED_WIN ? 1. : move_result == BLUE_WIN ? -1. : 0.;
    return move_result == RED_WIN ? -1. : move_result == BLUE_WIN ? 1. : 0.;
}

struct EvaluationResult {
    std::vector<std::pair<unsigned, Move>> legal_moves;
    double value;
    double priors[2][25];
};

struct MctsArena {  // mcts_arena.rs:37-73
    State root_state;
    int root_color;
    double exploration_c;
    uint32_t max_playouts;
    std::vector<MctsNode> arena;
    uint32_t playouts = 0;
    eval_fn eval; void* user;
    bool pass_seen = false;
    bool train = false;          // root Dirichlet noise (mcts_arena.rs:186-202)
    double epsilon = 0.25, eta = 0.03;
    uint64_t noise_key = 0;
    uint64_t depth_sum = 0;      // instrumentation for DESIGN.md's d and k
    uint64_t expanded_children = 0, expansions = 0;

    static MctsNode make_node(int64_t parent, uint32_t idx, bool has_mov, DoneMove mov, int color, double prob) {
        MctsNode n; n.parent = parent; n.idx = idx; n.has_mov = has_mov; n.mov = mov; n.is_pass = false; n.visits = 0;
        n.reward = 0.; n.winrate = 0.; n.is_terminal = false; n.is_expanded = false; n.player_color = color; n.probability = prob;
        return n;
    }
    MctsArena(const State& s, int color, double c, uint32_t sims, eval_fn e, void* u)
        : root_state(s), root_color(color), exploration_c(c), max_playouts(sims), eval(e), user(u) {
        arena.push_back(make_node(-1, 0, false, DoneMove{}, color, 1.));  // :57-58
    }

    // mcts_arena.rs:183-223, eval mode (train-mode Dirichlet noise is thread_rng driven -> not restated)
    uint32_t select(const MctsNode& parent) const {
        const auto& children = parent.children;
        if (parent.parent < 0 && train && children.size() > 1) {
            // max_by folds left to right and evaluates uct() of BOTH operands afresh at every comparison, each call with a
            // new noise sample; the last maximal element wins.
            const uint32_t k = (uint32_t)children.size();
            auto uct_noisy = [&](const MctsNode& child, uint32_t sample) {
                const double noise = (double)noise_beta(k, (float)eta, noise_key, parent.visits, sample);
                return child.winrate + exploration_c * (child.probability * (1. - epsilon) + noise * epsilon) *
                                           (std::sqrt((double)parent.visits) / (double)(child.visits + 1));
            };
            uint32_t best = children[0];
            for (uint32_t i = 1; i < k; ++i) {
                const double ua = uct_noisy(arena[best], 2 * (i - 1)), ub = uct_noisy(arena[children[i]], 2 * (i - 1) + 1);
                if (total_key(ua) <= total_key(ub)) best = children[i];
            }
            return best;
        }
        auto uct = [&](const MctsNode& child) {
            return child.winrate + exploration_c * child.probability * (std::sqrt((double)parent.visits) / (double)(child.visits + 1));
        };
        // Iterator::max_by keeps the LAST maximal element.
        uint32_t best = children[0];
        for (size_t i = 1; i < children.size(); ++i) {
            if (total_key(uct(arena[children[i]])) >= total_key(uct(arena[best]))) best = children[i];
        }
        return best;
    }

    // mcts_arena.rs:267-310
    EvaluationResult evaluate(const State& st, int color) const {
        float planes[525], policy[50], value = 0.f;
        encode_planes(st, color, planes);
        eval(planes, policy, &value, user);
        EvaluationResult r;
        r.legal_moves = st.generate_all_legal_moves(color);
        for (int c = 0; c < 2; ++c) for (int i = 0; i < 25; ++i) r.priors[c][i] = 0.;
        for (auto& cm : r.legal_moves) {
            unsigned idx = cm.second.to;
            r.priors[cm.first & 1u][idx] = (double)policy[(cm.first & 1u) * 25 + idx];
        }
        for (int c = 0; c < 2; ++c) {
            double sum = 0.;
            for (int i = 0; i < 25; ++i) sum += r.priors[c][i];  // sequential f64 fold
            if (sum > 0.) for (int i = 0; i < 25; ++i) r.priors[c][i] /= sum;
        }
        r.value = (double)value;
        return r;
    }

    // mcts_arena.rs:231-260
    void expand(uint32_t parent, const EvaluationResult& ev) {
        int color = arena[parent].player_color;
        for (auto& cm : ev.legal_moves) {
            uint32_t idx = (uint32_t)arena.size();
            double prob = ev.priors[cm.first & 1u][cm.second.to];
            arena.push_back(make_node(parent, idx, true, DoneMove{cm.second, cm.first}, enemy(color), prob));
            arena[parent].children.push_back(idx);
        }
        if (ev.legal_moves.empty()) {
            // Project-defined: the reference would panic in the next select (mcts_arena.rs:213-220).
            // Two pass pseudo-children (one per own hand slot, prior 1/2), mirroring simulate's pass rule.
            pass_seen = true;
            unsigned base = color == RED ? 0u : 2u;
            for (unsigned s = 0; s < 2; ++s) {
                uint32_t idx = (uint32_t)arena.size();
                MctsNode n = make_node(parent, idx, true, DoneMove{Move{0, 0, PAWN}, base + s}, enemy(color), 0.5);
                n.is_pass = true;
                arena.push_back(n);
                arena[parent].children.push_back(idx);
            }
        }
        arena[parent].is_expanded = true;
        expansions += 1; expanded_children += arena[parent].children.size();
    }

    // mcts_arena.rs:312-323
    void back_propagate(uint32_t node_idx, double reward) {
        int64_t n = node_idx;
        for (;;) {
            arena[n].update(reward);
            if (arena[n].parent >= 0) { n = arena[n].parent; reward = -reward; } else break;
        }
    }

    // mcts_arena.rs:127-177
    void playout() {
        State st = root_state;  // clone per playout (:128)
        int color = root_color;
        uint32_t node_idx = 0;
        while (arena[node_idx].is_expanded && !arena[node_idx].is_terminal) {
            node_idx = select(arena[node_idx]);
            depth_sum += 1;
            if (arena[node_idx].has_mov) {
                int64_t parent = arena[node_idx].parent;
                const DoneMove& mv = arena[node_idx].mov;
                int r = arena[node_idx].is_pass ? st.pass(mv.used_card_idx)
                                                : st.make_move(mv.mov, arena[parent].player_color, mv.used_card_idx);
                color = enemy(color);
                if (is_win(r)) arena[node_idx].is_terminal = true;
            }
        }
        // Q11: the reference evaluates even terminal leaves and discards the result; skipped here.
        bool need_expand = !arena[node_idx].is_expanded && !arena[node_idx].is_terminal;
        EvaluationResult ev; ev.value = 0.;
        int result = st.current_state();
        if (need_expand || !is_win(result)) ev = evaluate(st, color);
        if (need_expand) expand(node_idx, ev);
        int64_t parent = arena[node_idx].parent >= 0 ? arena[node_idx].parent : 0;
        int reward_color = arena[parent].player_color;
        if (is_win(result)) back_propagate(node_idx, reward_fn(result, reward_color));
        else back_propagate(node_idx, ev.value);
    }

    // mcts_arena.rs:75-102 with search_time = infinity
    uint32_t search(float pi[50]) {
        while (playouts < max_playouts) { playout(); playouts += 1; }
        const auto& children = arena[0].children;
        // calculate_priors, mcts_arena.rs:104-124 (f32 tensor arithmetic)
        for (int i = 0; i < 50; ++i) pi[i] = 0.f;
        for (uint32_t c : children) {
            const MctsNode& ch = arena[c];
            pi[(ch.mov.used_card_idx & 1u) * 25 + (ch.is_pass ? 0 : ch.mov.mov.to)] += (float)ch.visits;
        }
        float sum = 0.f;
        for (int i = 0; i < 50; ++i) sum += pi[i];
        if ((double)sum > 0.) for (int i = 0; i < 50; ++i) pi[i] = pi[i] / sum;
        if (children.empty()) return 0;
        uint32_t best = children[0];
        for (size_t i = 1; i < children.size(); ++i) {
            double a = (double)arena[children[i]].visits / (double)arena[0].visits;
            double b = (double)arena[best].visits / (double)arena[0].visits;
            if (total_key(a) >= total_key(b)) best = children[i];
        }
        return best;
    }
};

static void uniform_eval(const float*, float* policy50, float* value, void*) {
    for (int i = 0; i < 50; ++i) policy50[i] = 1.0f / 50.0f;
    *value = 0.f;
}
// deterministic non-uniform evaluator used by the parity tests (must match tests/ and the CUDA test evaluator):
// policy[i] = softmax-free positive weights from a hash of the planes, value = small hash-derived number.
static void hash_eval(const float* planes, float* policy50, float* value, void*) {
    uint64_t h = 0x243F6A8885A308D3ull;
    for (int i = 0; i < 525; ++i) if (planes[i] != 0.f) h = mix64(h ^ (uint64_t)(i + 1));
    float tot = 0.f;
    for (int i = 0; i < 50; ++i) {
        uint32_t r = (uint32_t)(mix64(h + (uint64_t)i * 0x9E3779B97F4A7C15ull) >> 40);  // 24 bits
        policy50[i] = (float)(r + 1) * (1.0f / 16777216.0f);
        tot += policy50[i];
    }
    for (int i = 0; i < 50; ++i) policy50[i] = policy50[i] / tot;
    uint32_t rv = (uint32_t)(mix64(h ^ 0xA5A5A5A5A5A5A5A5ull) >> 40);
    *value = ((float)rv * (1.0f / 16777216.0f)) * 2.0f - 1.0f;
}

}  // namespace orc

// =====================================================================================
// C entry points (ctypes)
// =====================================================================================
extern "C" {

void orc_attack_maps(uint32_t* out800) { memcpy(out800, orc::ATTACK_MAPS, sizeof(orc::ATTACK_MAPS)); }
int orc_card_color(int card) { return orc::ORIGINAL_CARDS[card].player_color; }
uint32_t orc_rand_u32(uint64_t seed, uint64_t game, uint32_t step, uint32_t draw) { return orc::rand_u32(seed, game, step, draw); }
void orc_deal(uint64_t seed, uint64_t game, uint32_t epoch, uint8_t* out5) { orc::deal(seed, game, epoch, out5); }

void orc_new_games(orc_state* g, int64_t n, uint64_t game0, uint64_t seed, uint32_t epoch, const uint8_t* fixed_deck) {
    for (int64_t i = 0; i < n; ++i) orc::new_game(g[i], seed, game0 + (uint64_t)i, epoch, fixed_deck);
}

// all legal moves of `side` as action codes in reference order; returns count (<= 40)
int orc_gen_moves(const orc_state* g, int side, uint16_t* out) {
    orc::State st = orc::to_state(*g);
    int n = 0;
    for (auto& cm : st.generate_all_legal_moves(side)) out[n++] = orc::encode_action(cm.first, cm.second);
    return n;
}
// moves of one card index (not slot) as in State::generate_legal_moves(color, card)
int orc_gen_moves_card(const orc_state* g, int side, int card_index, uint16_t* out) {
    orc::State st = orc::to_state(*g);
    int n = 0;
    for (auto& m : st.generate_legal_moves(side, card_index)) out[n++] = orc::encode_action(0, m);
    return n;
}
int orc_make_move(orc_state* g, uint16_t action) { return orc::apply_action(*g, action); }
int orc_current_state(const orc_state* g) { return orc::to_state(*g).current_state(); }
void orc_legal_masks(const orc_state* g, int64_t n, uint32_t* out2n) {
    for (int64_t i = 0; i < n; ++i) orc::legal_mask(orc::to_state(g[i]), g[i].side, out2n + 2 * i);
}
void orc_encode(const orc_state* g, int64_t n, float* out525n) {
    for (int64_t i = 0; i < n; ++i) orc::encode_planes(orc::to_state(g[i]), g[i].side, out525n + 525 * i);
}

// lockstep env step with explicit actions. Finished games (result != 0) are left untouched.
void orc_env_step(orc_state* g, int64_t n, const uint16_t* actions) {
    for (int64_t i = 0; i < n; ++i) if (g[i].result == 0 && actions[i] != 0xFFFF) orc::apply_action(g[i], actions[i]);
}
// lockstep env step with the counter-RNG random policy; writes the chosen actions (0xFFFF for finished games).
// auto_reset: a game that ends at `step` is replaced by a fresh deal with epoch step+1.
void orc_env_step_random(orc_state* g, int64_t n, uint64_t game0, uint64_t seed, uint32_t step, int policy, int auto_reset,
                         const uint8_t* fixed_deck, uint16_t* actions_out) {
    for (int64_t i = 0; i < n; ++i) {
        if (g[i].result != 0) { if (actions_out) actions_out[i] = 0xFFFF; continue; }
        uint16_t a = orc::choose_random_action(orc::to_state(g[i]), g[i].side, policy, seed, game0 + (uint64_t)i, step);
        orc::apply_action(g[i], a);
        if (actions_out) actions_out[i] = a;
        if (auto_reset && g[i].result != 0) orc::new_game(g[i], seed, game0 + (uint64_t)i, step + 1, fixed_deck);
    }
}

// Play games [game0, game0+n) to terminal (or max_plies) with the random policy. Outputs per game: final state,
// plies, and a trace hash (fold of mix64 over the action codes). Returns total env steps.
int64_t orc_playout_games(orc_state* final_states, int64_t n, uint64_t game0, uint64_t seed, int policy, uint32_t max_plies,
                          const uint8_t* fixed_deck, uint32_t* plies_out, uint64_t* trace_out) {
    int64_t total = 0;
    for (int64_t i = 0; i < n; ++i) {
        orc_state g;
        orc::new_game(g, seed, game0 + (uint64_t)i, 0, fixed_deck);
        uint32_t ply = 0; uint64_t trace = 0;
        while (g.result == 0 && ply < max_plies) {
            uint16_t a = orc::choose_random_action(orc::to_state(g), g.side, policy, seed, game0 + (uint64_t)i, ply);
            orc::apply_action(g, a);
            trace = orc::mix64(trace ^ (uint64_t)a);
            ++ply;
        }
        final_states[i] = g;
        if (plies_out) plies_out[i] = ply;
        if (trace_out) trace_out[i] = trace;
        total += ply;
    }
    return total;
}

// perft from `g` with side g->side; arrays of length depth.
void orc_perft(const orc_state* g, int depth, uint64_t* nodes, uint64_t* wins, uint64_t* zero) {
    for (int d = 0; d < depth; ++d) nodes[d] = wins[d] = zero[d] = 0;
    orc::PerftOut o{nodes, wins, zero};
    if (orc::is_win(orc::to_state(*g).current_state())) return;
    orc::perft_rec(orc::to_state(*g), g->side, 0, depth, o);
}

struct orc_tree_dump {  // flat copies of the arena, length = n_nodes
    uint32_t* visits; double* reward; double* winrate; double* prior; uint16_t* action; int32_t* parent;
    uint32_t* first_child; uint32_t* n_child; uint8_t* flags;  // flags: bit0 expanded, bit1 terminal, bit2 pass child
};

// One PUCT search (eval mode, no wall clock). evaluator: 0 uniform, 1 hash_eval, 2 callback `cb`.
// Returns the number of nodes; out_* may be NULL. If dump != NULL and cap >= n_nodes the arena is copied out.
static double g_noise_eps = 0.25, g_noise_alpha = 0.03;
static uint64_t g_noise_seed = 0, g_noise_game0 = 0;
static int g_noise_on = 0;
// train-mode switch for the searches below (global: the oracle is single-purpose test code)
void orc_mcts_set_noise(int enabled, double epsilon, double alpha, uint64_t seed, uint64_t game0) {
    g_noise_on = enabled; g_noise_eps = epsilon; g_noise_alpha = alpha; g_noise_seed = seed; g_noise_game0 = game0;
}
static void apply_noise_cfg(orc::MctsArena& a, uint64_t tree) {
    a.train = g_noise_on != 0; a.epsilon = g_noise_eps; a.eta = g_noise_alpha;
    a.noise_key = orc::mix64(g_noise_seed + 0x9E3779B97F4A7C15ull * (g_noise_game0 + tree + 1));
}

int64_t orc_mcts_search(const orc_state* g, double c_puct, uint32_t sims, int evaluator, orc::eval_fn cb, void* user,
                        uint16_t* best_action, float* pi50, uint32_t* root_visits, double* root_q, int32_t* pass_seen,
                        double* mean_depth, double* mean_children, orc_tree_dump* dump, int64_t cap) {
    orc::eval_fn e = evaluator == 0 ? orc::uniform_eval : evaluator == 1 ? orc::hash_eval : cb;
    orc::MctsArena arena(orc::to_state(*g), g->side, c_puct, sims, e, user);
    apply_noise_cfg(arena, 0);
    float pi[50];
    uint32_t best = arena.search(pi);
    if (best_action) {
        if (arena.arena[0].children.empty()) *best_action = 0xFFFF;
        else {
            const orc::MctsNode& b = arena.arena[best];
            *best_action = b.is_pass ? orc::encode_pass(b.mov.used_card_idx) : orc::encode_action(b.mov.used_card_idx, b.mov.mov);
        }
    }
    if (pi50) memcpy(pi50, pi, sizeof(pi));
    if (root_visits) *root_visits = arena.arena[0].visits;
    if (root_q) *root_q = arena.arena[0].winrate;
    if (pass_seen) *pass_seen = arena.pass_seen ? 1 : 0;
    if (mean_depth) *mean_depth = sims ? (double)arena.depth_sum / (double)sims : 0.;
    if (mean_children) *mean_children = arena.expansions ? (double)arena.expanded_children / (double)arena.expansions : 0.;
    int64_t nn = (int64_t)arena.arena.size();
    if (dump && cap >= nn) {
        for (int64_t i = 0; i < nn; ++i) {
            const orc::MctsNode& nd = arena.arena[i];
            dump->visits[i] = nd.visits; dump->reward[i] = nd.reward; dump->winrate[i] = nd.winrate; dump->prior[i] = nd.probability;
            dump->action[i] = !nd.has_mov ? 0xFFFF : nd.is_pass ? orc::encode_pass(nd.mov.used_card_idx)
                                                                 : orc::encode_action(nd.mov.used_card_idx, nd.mov.mov);
            dump->parent[i] = (int32_t)nd.parent;
            dump->first_child[i] = nd.children.empty() ? 0u : nd.children[0];
            dump->n_child[i] = (uint32_t)nd.children.size();
            dump->flags[i] = (uint8_t)((nd.is_expanded ? 1 : 0) | (nd.is_terminal ? 2 : 0) | (nd.is_pass ? 4 : 0));
        }
    }
    return nn;
}

void orc_hash_eval(const float* planes, float* policy50, float* value) { orc::hash_eval(planes, policy50, value, nullptr); }

// Batched searches over many roots (multi-threaded; used for the timed CPU baseline and bulk parity).
// Outputs per tree: best action, root child visit vector (40 slots, zero padded), n_nodes, root_q.
void orc_mcts_search_batch(const orc_state* roots, int64_t n, double c_puct, uint32_t sims, int evaluator, int threads,
                           uint16_t* best_actions, uint32_t* child_visits40, int64_t* n_nodes, double* root_q, float* pi50,
                           int32_t* pass_seen) {
    std::atomic<int64_t> next(0);
    auto work = [&]() {
        for (;;) {
            int64_t i = next.fetch_add(1);
            if (i >= n) break;
            orc::eval_fn e = evaluator == 0 ? orc::uniform_eval : orc::hash_eval;
            orc::MctsArena arena(orc::to_state(roots[i]), roots[i].side, c_puct, sims, e, nullptr);
            apply_noise_cfg(arena, (uint64_t)i);
            float pi[50];
            uint32_t best = arena.search(pi);
            const auto& ch = arena.arena[0].children;
            if (best_actions) {
                if (ch.empty()) best_actions[i] = 0xFFFF;
                else {
                    const orc::MctsNode& b = arena.arena[best];
                    best_actions[i] = b.is_pass ? orc::encode_pass(b.mov.used_card_idx) : orc::encode_action(b.mov.used_card_idx, b.mov.mov);
                }
            }
            if (child_visits40) for (size_t k = 0; k < 40; ++k) child_visits40[i * 40 + k] = k < ch.size() ? arena.arena[ch[k]].visits : 0u;
            if (n_nodes) n_nodes[i] = (int64_t)arena.arena.size();
            if (root_q) root_q[i] = arena.arena[0].winrate;
            if (pi50) memcpy(pi50 + 50 * i, pi, sizeof(pi));
            if (pass_seen) pass_seen[i] = arena.pass_seen ? 1 : 0;
        }
    };
    if (threads <= 1) { work(); return; }
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) pool.emplace_back(work);
    for (auto& t : pool) t.join();
}

// ---- timed CPU baselines (ref-shaped code paths above), multi-threaded over games ----
// cfg 3: n games, `steps` lockstep random steps with auto-reset + legal mask + plane encode per step.
// Returns seconds. `sink` receives a checksum so the work cannot be optimised away.
// cfg 3 on a game array the caller keeps between calls: `steps` lockstep steps (step ids step0, step0+1, ...) of games
// [0, n) with global ids game0 + i: random action, apply, auto-reset, legal mask and (with_encode) the 21x5x5 planes.
// Returns seconds. This is what bench.py times per step for the CPU arm: the same 1 048 576 games, one step per timed step.
double orc_bench_env_steps(orc_state* g, int64_t n, uint64_t game0, uint32_t step0, uint32_t steps, uint64_t seed, int threads, int with_encode,
                           double* sink) {
    std::vector<double> sums((size_t)std::max(threads, 1), 0.);
    auto t0 = std::chrono::steady_clock::now();
    auto work = [&](int tid, int64_t lo, int64_t hi) {
        std::vector<float> planes(525);
        uint32_t mask[2];
        double acc = 0.;
        for (uint32_t s = step0; s < step0 + steps; ++s)
            for (int64_t i = lo; i < hi; ++i) {
                uint16_t a = orc::choose_random_action(orc::to_state(g[i]), g[i].side, 0, seed, game0 + (uint64_t)i, s);
                orc::apply_action(g[i], a);
                if (g[i].result != 0) orc::new_game(g[i], seed, game0 + (uint64_t)i, s + 1, nullptr);
                orc::legal_mask(orc::to_state(g[i]), g[i].side, mask);
                acc += mask[0] ^ mask[1];
                if (with_encode) { orc::encode_planes(orc::to_state(g[i]), g[i].side, planes.data()); acc += planes[(s * 7 + i) % 525]; }
            }
        sums[tid] = acc;
    };
    if (threads <= 1) work(0, 0, n);
    else {
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; ++t) pool.emplace_back(work, t, n * t / threads, n * (t + 1) / threads);
        for (auto& t : pool) t.join();
    }
    double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (sink) { double s = 0; for (double v : sums) s += v; *sink = s; }
    return dt;
}
double orc_bench_env(int64_t n, uint32_t steps, uint64_t seed, int threads, int with_encode, double* sink) {
    std::vector<orc_state> g((size_t)n);
    orc_new_games(g.data(), n, 0, seed, 0, nullptr);
    return orc_bench_env_steps(g.data(), n, 0, 0, steps, seed, threads, with_encode, sink);
}

// cfg 4: n trees (roots given), `sims` playouts each, uniform evaluator. Returns seconds.
double orc_bench_mcts(const orc_state* roots, int64_t n, double c_puct, uint32_t sims, int threads, double* sink) {
    std::vector<int64_t> nodes((size_t)n);
    auto t0 = std::chrono::steady_clock::now();
    orc_mcts_search_batch(roots, n, c_puct, sims, 0, threads, nullptr, nullptr, nodes.data(), nullptr, nullptr, nullptr);
    double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (sink) { double s = 0; for (int64_t v : nodes) s += (double)v; *sink = s; }
    return dt;
}

// cfg 2: perft over a list of deals, multi-threaded. Returns seconds; totals[d] summed over deals.
double orc_bench_perft(const uint8_t* decks5, int64_t n_decks, int depth, int threads, uint64_t* totals) {
    std::vector<std::vector<uint64_t>> part((size_t)std::max(threads, 1), std::vector<uint64_t>((size_t)depth, 0));
    std::atomic<int64_t> next(0);
    auto t0 = std::chrono::steady_clock::now();
    auto work = [&](int tid) {
        std::vector<uint64_t> nodes((size_t)depth), wins((size_t)depth), zero((size_t)depth);
        for (;;) {
            int64_t i = next.fetch_add(1);
            if (i >= n_decks) break;
            orc_state g;
            orc::new_game(g, 0, 0, 0, decks5 + 5 * i);
            orc_perft(&g, depth, nodes.data(), wins.data(), zero.data());
            for (int d = 0; d < depth; ++d) part[tid][d] += nodes[d];
        }
    };
    if (threads <= 1) work(0);
    else {
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; ++t) pool.emplace_back(work, t);
        for (auto& t : pool) t.join();
    }
    double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (int d = 0; d < depth; ++d) { totals[d] = 0; for (auto& p : part) totals[d] += p[d]; }
    return dt;
}

}  // extern "C"

// ---- policy/value network (harness restatement of alphazero-training/src/net.rs:9-232, eval mode) -----------------------------
// Plain f32 tensors, f64 accumulation inside each convolution / linear layer (a checker, not a timed path). Layer by layer as the
// reference: SmallBlock = conv3x3(pad 1) -> batch_norm (running statistics, eps 1e-5) (net.rs:9-38); ResNetBlock =
// relu(block2(relu(block1(x))) + x) (net.rs:40-66); value head conv1x1 -> bn -> relu -> linear 25->h -> relu -> linear h->1 -> tanh,
// policy head conv1x1 (2 channels) -> bn -> relu -> linear 50->50 -> softmax -> [2,25] (net.rs:118-232). Parameters are looked up
// by their VarStore names ('|' or '.' separators). Parity with the Rust/libtorch original is UNPINNED (no tch here); the restatement
// is pinned against the PyTorch twin (onitama_alphazero_b200/net.py, which loads the reference's .ot archives) by tests/test_net_cpu.py.
#include <map>
#include <string>
namespace orc {
struct NetParams {
    std::map<std::string, std::pair<const float*, int64_t>> t;
    const float* get(const std::string& name, int64_t numel) const {
        auto it = t.find(name);
        if (it == t.end() || it->second.second != numel) return nullptr;
        return it->second.first;
    }
};
// y[co][5][5] = sum_ci sum_taps w[co][ci][ky][kx] * x[ci][y+ky-1][x+kx-1] + b[co], then batch norm with running statistics
static bool conv_bn(const NetParams& P, const std::string& conv, const std::string& bn, int c_in, int c_out, int ks, const std::vector<float>& x,
                    std::vector<float>& y) {
    const float *w = P.get(conv + ".weight", (int64_t)c_out * c_in * ks * ks), *b = P.get(conv + ".bias", c_out), *g = P.get(bn + ".weight", c_out),
                *be = P.get(bn + ".bias", c_out), *mu = P.get(bn + ".running_mean", c_out), *var = P.get(bn + ".running_var", c_out);
    if (!w || !b || !g || !be || !mu || !var) return false;
    y.assign((size_t)c_out * 25, 0.f);
    const int r = ks / 2;
    for (int co = 0; co < c_out; ++co)
        for (int py = 0; py < 5; ++py)
            for (int px = 0; px < 5; ++px) {
                double acc = 0.0;
                for (int ci = 0; ci < c_in; ++ci)
                    for (int ky = 0; ky < ks; ++ky)
                        for (int kx = 0; kx < ks; ++kx) {
                            const int sy = py + ky - r, sx = px + kx - r;
                            if (sy < 0 || sy >= 5 || sx < 0 || sx >= 5) continue;
                            acc += (double)w[(((size_t)co * c_in + ci) * ks + ky) * ks + kx] * (double)x[(size_t)ci * 25 + sy * 5 + sx];
                        }
                const float conv_out = (float)(acc + (double)b[co]);
                y[(size_t)co * 25 + py * 5 + px] = (float)(((double)conv_out - (double)mu[co]) / std::sqrt((double)var[co] + 1e-5) * (double)g[co] + (double)be[co]);
            }
    return true;
}
static bool linear(const NetParams& P, const std::string& name, int n_in, int n_out, const std::vector<float>& x, std::vector<float>& y) {
    const float *w = P.get(name + ".weight", (int64_t)n_out * n_in), *b = P.get(name + ".bias", n_out);
    if (!w || !b) return false;
    y.assign((size_t)n_out, 0.f);
    for (int j = 0; j < n_out; ++j) {
        double acc = (double)b[j];
        for (int i = 0; i < n_in; ++i) acc += (double)w[(size_t)j * n_in + i] * (double)x[i];
        y[j] = (float)acc;
    }
    return true;
}
static void relu(std::vector<float>& v) { for (float& x : v) x = x > 0.f ? x : 0.f; }
}  // namespace orc

extern "C" {
// planes [n][21][5][5] -> policy [n][50] (softmax), value [n]. Returns 0, or -1 when a tensor is missing / has the wrong size.
int orc_net_forward(int32_t n_tensors, const char* const* names, const float* const* data, const int64_t* numel, const float* planes, int64_t n,
                    float* policy, float* value) {
    orc::NetParams P;
    int hidden = 0;
    for (int32_t i = 0; i < n_tensors; ++i) {
        std::string s(names[i]);
        for (char& ch : s) if (ch == '|') ch = '.';
        P.t[s] = {data[i], numel[i]};
        if (s == "bn1.weight") hidden = (int)numel[i];
    }
    if (hidden <= 0) return -1;
    int n_blocks = 0;
    while (P.t.count("resnet_" + std::to_string(n_blocks) + ".resnet_small_block1.small_block_conv.weight")) ++n_blocks;
    const int c_in = (int)(P.t.count("conv_init_1.weight") ? P.t["conv_init_1.weight"].second / (9 * hidden) : 0);
    if (c_in <= 0) return -1;
    for (int64_t s = 0; s < n; ++s) {
        std::vector<float> x(planes + s * c_in * 25, planes + (s + 1) * c_in * 25), y, h, t;
        if (!orc::conv_bn(P, "conv_init_1", "bn1", c_in, hidden, 3, x, y)) return -1;
        orc::relu(y);
        for (int b = 0; b < n_blocks; ++b) {
            const std::string blk = "resnet_" + std::to_string(b) + ".resnet_small_block";
            if (!orc::conv_bn(P, blk + "1.small_block_conv", blk + "1.small_block_bn", hidden, hidden, 3, y, h)) return -1;
            orc::relu(h);
            if (!orc::conv_bn(P, blk + "2.small_block_conv", blk + "2.small_block_bn", hidden, hidden, 3, h, t)) return -1;
            for (size_t i = 0; i < t.size(); ++i) t[i] += y[i];
            orc::relu(t);
            y.swap(t);
        }
        std::vector<float> v, v1, v2, p, logits;
        if (!orc::conv_bn(P, "vh_conv", "vh_bn", hidden, 1, 1, y, v)) return -1;
        orc::relu(v);
        if (!orc::linear(P, "vh_linear1", 25, hidden, v, v1)) return -1;
        orc::relu(v1);
        if (!orc::linear(P, "vh_linear2", hidden, 1, v1, v2)) return -1;
        value[s] = (float)std::tanh((double)v2[0]);
        if (!orc::conv_bn(P, "policy_conv", "policy_bn", hidden, 2, 1, y, p)) return -1;
        orc::relu(p);
        if (!orc::linear(P, "ph_linear2", 50, 50, p, logits)) return -1;
        double m = logits[0], sum = 0.0;
        for (float l : logits) m = std::max(m, (double)l);
        for (float l : logits) sum += std::exp((double)l - m);
        for (int j = 0; j < 50; ++j) policy[s * 50 + j] = (float)(std::exp((double)logits[j] - m) / sum);
    }
    return 0;
}
}  // extern "C"

// ---- plain UCT search with random rollouts (onitama-game/src/ai/mcts/mcts_arena.rs:16-264, the `Mcts` agent of mcts/mod.rs) ----
// f32 statistics as in the reference: reward sums of +-1 / 0 (exact), winrate = reward / visits, UCT = winrate + c * sqrt(ln(N_parent)
// / n) with Iterator::max_by(total_cmp) (the LAST maximal child wins; unvisited children score +inf). A node is expanded only once
// it has been visited more than min_node_visits times (:116-121); every playout ends with a uniformly random rollout from the leaf's
// position (`simulate`, :190-241, including its pass rule). Project-defined: rand::thread_rng is replaced by the counter RNG -- draw
// 16 + 2*ply (+1 for the pass slot) of step = playout index; rollouts are cut after 4096 plies (reward 0; the reference loops on);
// zero-legal-move expansions get two pass pseudo-children as in the PUCT arena (the reference would panic in the next select).
namespace orc {
struct UctNode {
    int64_t parent;
    std::vector<uint32_t> children;
    bool has_mov, is_pass;
    DoneMove mov;
    uint32_t visits;
    float reward, winrate;
    bool is_terminal, is_expanded;
    int player_color;
};
struct UctArena {
    State root_state;
    int root_color;
    uint32_t min_node_visits;
    float exploration_c;
    uint64_t seed, game;
    std::vector<UctNode> arena;
    uint32_t playouts = 0;
    bool pass_seen = false;
    uint64_t rollout_plies = 0;

    static UctNode make_node(int64_t parent, bool has_mov, DoneMove mov, int color) {
        UctNode n; n.parent = parent; n.has_mov = has_mov; n.is_pass = false; n.mov = mov; n.visits = 0; n.reward = 0.f; n.winrate = 0.f;
        n.is_terminal = false; n.is_expanded = false; n.player_color = color;
        return n;
    }
    UctArena(const State& s, int color, uint32_t min_visits, float c, uint64_t seed_, uint64_t game_)
        : root_state(s), root_color(color), min_node_visits(min_visits), exploration_c(c), seed(seed_), game(game_) {
        arena.push_back(make_node(-1, false, DoneMove{}, color));
    }
    static int32_t total_key32(float x) {  // f32::total_cmp as an integer key
        int32_t b; memcpy(&b, &x, 4);
        b ^= (int32_t)((uint32_t)(b >> 31) >> 1);
        return b;
    }
    uint32_t select(const UctNode& parent) const {  // :137-160
        const float lnp = std::log((float)parent.visits);
        auto uct = [&](const UctNode& child) { return child.winrate + exploration_c * std::sqrt(lnp / (float)child.visits); };
        uint32_t best = parent.children[0];
        for (size_t i = 1; i < parent.children.size(); ++i)
            if (total_key32(uct(arena[parent.children[i]])) >= total_key32(uct(arena[best]))) best = parent.children[i];
        return best;
    }
    void expand(uint32_t parent, const State& st) {  // :167-187
        const int color = arena[parent].player_color;
        auto moves = st.generate_all_legal_moves(color);
        for (auto& cm : moves) {
            const uint32_t idx = (uint32_t)arena.size();
            arena.push_back(make_node(parent, true, DoneMove{cm.second, cm.first}, enemy(color)));
            arena[parent].children.push_back(idx);
        }
        if (moves.empty()) {
            pass_seen = true;
            const unsigned base = color == RED ? 0u : 2u;
            for (unsigned s = 0; s < 2; ++s) {
                const uint32_t idx = (uint32_t)arena.size();
                UctNode n = make_node(parent, true, DoneMove{Move{0, 0, PAWN}, base + s}, enemy(color));
                n.is_pass = true;
                arena.push_back(n);
                arena[parent].children.push_back(idx);
            }
        }
        arena[parent].is_expanded = true;
    }
    float simulate(State st, int color, int reward_color, uint32_t playout) {  // :190-241
        int move_result = st.current_state();
        uint32_t ply = 0;
        while (!is_win(move_result)) {
            if (ply >= 4096u) return 0.f;
            auto moves = st.generate_all_legal_moves(color);
            if (moves.empty()) {
                const unsigned base = color == RED ? 0u : 2u;
                st.pass(base + rand_index(rand_u32(seed, game, playout, 16u + 2u * ply + 1u), 2));
                color = enemy(color);
                ++ply; ++rollout_plies;
                continue;
            }
            auto& pick = moves[rand_index(rand_u32(seed, game, playout, 16u + 2u * ply), (uint32_t)moves.size())];
            move_result = st.make_move(pick.second, color, pick.first);
            color = enemy(color);
            ++ply; ++rollout_plies;
        }
        return (float)reward_fn(move_result, reward_color);
    }
    void playout() {  // :87-131
        State st = root_state;
        int color = root_color;
        uint32_t node_idx = 0;
        while (arena[node_idx].is_expanded && !arena[node_idx].is_terminal) {
            node_idx = select(arena[node_idx]);
            const int64_t parent = arena[node_idx].parent;
            const DoneMove& mv = arena[node_idx].mov;
            const int r = arena[node_idx].is_pass ? st.pass(mv.used_card_idx) : st.make_move(mv.mov, arena[parent].player_color, mv.used_card_idx);
            color = enemy(color);
            if (is_win(r)) arena[node_idx].is_terminal = true;
        }
        if (!arena[node_idx].is_expanded && !arena[node_idx].is_terminal && arena[node_idx].visits > min_node_visits) expand(node_idx, st);
        const int64_t parent = arena[node_idx].parent >= 0 ? arena[node_idx].parent : 0;
        float reward = simulate(st, color, arena[parent].player_color, playouts);
        int64_t n = node_idx;  // back_propagate, :243-254
        for (;;) {
            UctNode& nd = arena[n];
            nd.visits += 1; nd.reward += reward; nd.winrate = nd.reward / (float)nd.visits;  // MctsNode::update, :300-304
            if (nd.parent >= 0) { n = nd.parent; reward = -reward; } else break;
        }
    }
    int64_t search(uint32_t max_playouts) {  // :55-84: the child with the most visits, last maximum wins (max_by_key)
        while (playouts < max_playouts) { playout(); playouts += 1; }
        const auto& ch = arena[0].children;
        if (ch.empty()) return -1;
        uint32_t best = ch[0];
        for (size_t i = 1; i < ch.size(); ++i) if (arena[ch[i]].visits >= arena[best].visits) best = ch[i];
        return best;
    }
};
}  // namespace orc

extern "C" {
// One search per root (threads > 1: roots in parallel). Outputs per tree: best action (0xFFFF if the root was never expanded),
// root child visits / reward sums (40 slots, zero padded), node count, winrate of the best child, pass flag.
void orc_uct_search_batch(const orc_state* roots, int64_t n, float exploration_c, uint32_t min_node_visits, uint32_t sims, uint64_t seed,
                          uint64_t game0, int threads, uint16_t* best_actions, uint32_t* child_visits40, int32_t* child_rewards40, int64_t* n_nodes,
                          float* best_winrate, int32_t* pass_seen, uint64_t* rollout_plies) {
    std::atomic<int64_t> next(0);
    std::atomic<uint64_t> plies(0);
    auto work = [&]() {
        for (;;) {
            const int64_t i = next.fetch_add(1);
            if (i >= n) break;
            orc::UctArena a(orc::to_state(roots[i]), roots[i].side, min_node_visits, exploration_c, seed, game0 + (uint64_t)i);
            const int64_t best = a.search(sims);
            const auto& ch = a.arena[0].children;
            if (best_actions) {
                if (best < 0) best_actions[i] = 0xFFFF;
                else {
                    const orc::UctNode& b = a.arena[(size_t)best];
                    best_actions[i] = b.is_pass ? orc::encode_pass(b.mov.used_card_idx) : orc::encode_action(b.mov.used_card_idx, b.mov.mov);
                }
            }
            for (size_t k = 0; k < 40; ++k) {
                if (child_visits40) child_visits40[i * 40 + k] = k < ch.size() ? a.arena[ch[k]].visits : 0u;
                if (child_rewards40) child_rewards40[i * 40 + k] = k < ch.size() ? (int32_t)a.arena[ch[k]].reward : 0;
            }
            if (n_nodes) n_nodes[i] = (int64_t)a.arena.size();
            if (best_winrate) best_winrate[i] = best < 0 ? 0.f : a.arena[(size_t)best].winrate;
            if (pass_seen) pass_seen[i] = a.pass_seen ? 1 : 0;
            plies.fetch_add(a.rollout_plies);
        }
    };
    if (threads <= 1) work();
    else {
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; ++t) pool.emplace_back(work);
        for (auto& t : pool) t.join();
    }
    if (rollout_plies) *rollout_plies = plies.load();
}
}  // extern "C"
