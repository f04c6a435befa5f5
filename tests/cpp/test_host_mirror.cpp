// The reference's own unit tests re-expressed against the C++ host mirror (include/onitama_b200.hpp); every rules call
// below executes on the GPU through libonb.so. Sources: onitama-game/src/game/state.rs:420-889,
// onitama-game/src/ai/mcts/mcts_arena.rs:403-457. Exit code 0 + "ALL OK" on success.
#include <cstdio>
#include <cmath>
#include "onitama_b200.hpp"

using namespace onitama;
static int failures = 0;
#define CHECK(cond) do { if (!(cond)) { std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond); ++failures; } } while (0)

static Move mv(uint32_t r0, uint32_t c0, uint32_t r1, uint32_t c1, PieceKind p) { return Move::from_2d({{{r0, c0}, {r1, c1}}}, p); }

static void create_all_legal_moves_for_red_in_starting_position() {  // state.rs:420-455
    State state = State::with_deck(Deck({CRAB, RABBIT, DRAGON, TIGER, FROG}));
    auto cards = state.deck.get_player_cards(PlayerColor::Red);
    std::vector<Move> crab_expected = {mv(4, 0, 3, 0, PieceKind::Pawn), mv(4, 1, 3, 1, PieceKind::Pawn), mv(4, 2, 3, 2, PieceKind::King),
                                       mv(4, 3, 3, 3, PieceKind::Pawn), mv(4, 4, 3, 4, PieceKind::Pawn)};
    auto crab_moves = state.generate_legal_moves(PlayerColor::Red, *cards[0]);
    std::sort(crab_moves.begin(), crab_moves.end()); std::sort(crab_expected.begin(), crab_expected.end());
    CHECK(crab_moves == crab_expected);
    std::vector<Move> rabbit_expected = {mv(4, 0, 3, 1, PieceKind::Pawn), mv(4, 1, 3, 2, PieceKind::Pawn), mv(4, 2, 3, 3, PieceKind::King),
                                         mv(4, 3, 3, 4, PieceKind::Pawn)};
    auto rabbit_moves = state.generate_legal_moves(PlayerColor::Red, *cards[1]);
    std::sort(rabbit_moves.begin(), rabbit_moves.end()); std::sort(rabbit_expected.begin(), rabbit_expected.end());
    CHECK(rabbit_moves == rabbit_expected);
}
static void create_all_legal_moves_for_blue_in_starting_position() {  // state.rs:457-492
    State state = State::with_deck(Deck({DRAGON, TIGER, CRAB, RABBIT, FROG}));
    auto cards = state.deck.get_player_cards(PlayerColor::Blue);
    std::vector<Move> crab_expected = {mv(0, 0, 1, 0, PieceKind::Pawn), mv(0, 1, 1, 1, PieceKind::Pawn), mv(0, 2, 1, 2, PieceKind::King),
                                       mv(0, 3, 1, 3, PieceKind::Pawn), mv(0, 4, 1, 4, PieceKind::Pawn)};
    auto crab_moves = state.generate_legal_moves(PlayerColor::Blue, *cards[0]);
    std::sort(crab_moves.begin(), crab_moves.end()); std::sort(crab_expected.begin(), crab_expected.end());
    CHECK(crab_moves == crab_expected);
    std::vector<Move> rabbit_expected = {mv(0, 1, 1, 0, PieceKind::Pawn), mv(0, 2, 1, 1, PieceKind::King), mv(0, 3, 1, 2, PieceKind::Pawn),
                                         mv(0, 4, 1, 3, PieceKind::Pawn)};
    auto rabbit_moves = state.generate_legal_moves(PlayerColor::Blue, *cards[1]);
    std::sort(rabbit_moves.begin(), rabbit_moves.end()); std::sort(rabbit_expected.begin(), rabbit_expected.end());
    CHECK(rabbit_moves == rabbit_expected);
}
static State base() { return State::with_deck(Deck({CRAB, RABBIT, DRAGON, TIGER, FROG})); }

static void make_move_tests() {  // state.rs:495-816
    {   // make_move_as_red
        State s = base(); Card crab = *s.deck.get_player_cards(PlayerColor::Red)[0];
        CHECK(s.make_move(Move{20, 15, PieceKind::Pawn}, PlayerColor::Red, 0) == MoveResult::InProgress);
        CHECK(get_bit(s.pawns[0], 15) == 1 && get_bit(s.pawns[0], 20) == 0 && s.deck.neutral_card() == crab);
    }
    {   // make_move_as_blue
        State s = base(); Card tiger = *s.deck.get_player_cards(PlayerColor::Blue)[1];
        CHECK(s.make_move(Move{1, 11, PieceKind::Pawn}, PlayerColor::Blue, 3) == MoveResult::InProgress);
        CHECK(get_bit(s.pawns[1], 11) == 1 && get_bit(s.pawns[1], 1) == 0 && s.deck.neutral_card() == tiger);
    }
    {   // capture_as_red
        State s = base(); s.pawns[1] = 0x58010000u; Card crab = *s.deck.get_player_cards(PlayerColor::Red)[0];
        CHECK(s.make_move(Move{20, 15, PieceKind::Pawn}, PlayerColor::Red, 0) == MoveResult::Capture);
        CHECK(get_bit(s.pawns[0], 15) == 1 && get_bit(s.pawns[0], 20) == 0 && s.deck.neutral_card() == crab && s.pawns[1] == 0x58000000u);
    }
    {   // capture_as_blue
        State s = base(); s.pawns[1] = 0x58200000u; Card tiger = *s.deck.get_player_cards(PlayerColor::Blue)[1];
        CHECK(s.make_move(Move{10, 20, PieceKind::Pawn}, PlayerColor::Blue, 3) == MoveResult::Capture);
        CHECK(get_bit(s.pawns[1], 20) == 1 && get_bit(s.pawns[1], 10) == 0 && s.deck.neutral_card() == tiger && s.pawns[0] == 0x00000580u);
    }
    {   // capture_win_as_red
        State s = base(); s.pawns[0] = 0x01000680u;
        CHECK(s.make_move(Move{7, 2, PieceKind::Pawn}, PlayerColor::Red, 0) == MoveResult::RedWin);
        CHECK(get_bit(s.pawns[0], 2) == 1 && get_bit(s.pawns[0], 7) == 0 && s.kings[1] == 0);
    }
    {   // capture_win_as_blue
        State s = base(); s.pawns[1] = 0x58080000u;
        CHECK(s.make_move(Move{12, 22, PieceKind::Pawn}, PlayerColor::Blue, 3) == MoveResult::BlueWin);
        CHECK(get_bit(s.pawns[1], 22) == 1 && get_bit(s.pawns[1], 12) == 0 && s.kings[0] == 0);
    }
    {   // king_in_temple_as_blue
        State s = base(); s.kings[1] = 0x00080000u; s.kings[0] = 0x00002000u;
        CHECK(s.make_move(Move{12, 22, PieceKind::King}, PlayerColor::Blue, 3) == MoveResult::BlueWin);
        CHECK(get_bit(s.kings[1], 22) == 1 && get_bit(s.kings[1], 12) == 0 && s.kings[0] > 0);
    }
    {   // king_in_temple_as_red
        State s = base(); s.kings[1] = 0x00080000u; s.kings[0] = 0x01000000u;
        CHECK(s.make_move(Move{7, 2, PieceKind::King}, PlayerColor::Red, 0) == MoveResult::RedWin);
        CHECK(get_bit(s.kings[0], 2) == 1 && get_bit(s.kings[0], 7) == 0 && s.kings[1] > 0);
    }
}
static void no_legal_moves() {  // state.rs:819-889
    State a = State::with_deck(Deck({DRAGON, TIGER, RABBIT, HORSE, FROG}));
    a.kings[1] = 67108864; a.kings[0] = 512; a.pawns[0] = 61568; a.pawns[1] = 3221225472u;
    CHECK(a.generate_legal_moves(PlayerColor::Blue, *a.deck.get_player_cards(PlayerColor::Blue)[0]).empty());
    State b = State::with_deck(Deck({DRAGON, RABBIT, TIGER, HORSE, FROG}));
    b.kings[1] = 131072; b.kings[0] = 16384; b.pawns[0] = 2148009984u; b.pawns[1] = 138416256;
    auto cards = b.deck.get_player_cards(PlayerColor::Blue);
    CHECK(b.generate_legal_moves(PlayerColor::Blue, *cards[0]).empty());
    CHECK(b.generate_legal_moves(PlayerColor::Blue, *cards[1]).empty());
    CHECK(b.generate_all_legal_moves(PlayerColor::Blue).empty());
}
static void expand_order() {  // ai/mcts/mcts_arena.rs:403-457: enumeration order after expand
    State s = State::with_deck(Deck({DRAGON, FROG, TIGER, RABBIT, HORSE}));
    const char* expected[] = {"Dragon a1-c2", "Dragon b1-d2", "Dragon c1-a2", "Dragon c1-e2", "Dragon d1-b2", "Dragon e1-c2",
                              "Frog b1-a2", "Frog c1-b2", "Frog d1-c2", "Frog e1-d2"};
    auto moves = s.generate_all_legal_moves(PlayerColor::Red);
    CHECK(moves.size() == 10);
    for (size_t i = 0; i < moves.size() && i < 10; ++i) {
        std::string str = std::string(s.deck.get_card(moves[i].first).name()) + " " + Move::convert_idx_to_notation(moves[i].second.from) + "-" +
                          Move::convert_idx_to_notation(moves[i].second.to);
        CHECK(str == expected[i]);
    }
}
static void search_and_drivers() {
    // deck Dragon,Frog,Tiger,Rabbit,Horse, c = sqrt(2), 400 playouts, uniform evaluator -> best = (slot 1, 24 -> 18)
    TrainingAlphaZeroMcts mcts;
    mcts.config.max_playouts = 400;
    State s = State::with_deck(Deck({DRAGON, FROG, TIGER, RABBIT, HORSE}));
    auto r = mcts.generate_move_tensor(s, PlayerColor::Red);
    CHECK((r.first == DoneMove{Move{24, 18, PieceKind::Pawn}, 1}));
    float sum = 0.f; for (float p : r.second) sum += p;
    CHECK(std::fabs(sum - 1.f) < 1e-5f);
    CHECK(r.second[25 + 18] == 46.f / 399.f);  // pi[slot 1][to 18] = visits / sum of child visits (399), f32 division
    // the same search with the network as a host-side black box (uniform policy, value 0) must be identical
    TrainingAlphaZeroMcts host = mcts;
    host.config.max_playouts = 60;
    TrainingAlphaZeroMcts dev = host;
    host.model = [](const float*, int64_t n, float* pol, float* val) { for (int64_t i = 0; i < n * 50; ++i) pol[i] = 1.0f / 50.0f; for (int64_t i = 0; i < n; ++i) val[i] = 0.f; };
    auto rh = host.generate_move_tensor(s, PlayerColor::Red), rd = dev.generate_move_tensor(s, PlayerColor::Red);
    CHECK(rh.first == rd.first && rh.second == rd.second);
    // self_play: sample invariants (train.rs:55-88)
    TrainConfig tc; tc.mcts_config.max_playouts = 32; tc.self_play_game_amnt = 8; tc.max_plies = 10; tc.seed = 3;
    TrainingAlphaZeroMcts sp; sp.config = tc.mcts_config;
    auto data = self_play(sp, tc);
    CHECK(!data.empty() && data.size() <= 8 * 12);
    for (auto& d : data) {
        float ps = 0.f; for (float p : d.pi) ps += p;
        CHECK(std::fabs(ps - 1.f) < 1e-5f);
        CHECK(d.z == 0.f || d.z == 1.f || d.z == -1.f);
        float side_plane = d.state[20 * 25];
        CHECK(side_plane == (d.player_color == PlayerColor::Blue ? 1.f : 0.f));
    }
    // fight: AlphaZero MCTS (uniform evaluator, 64 playouts) against the Random agent, colours alternate
    EvaluatorConfig ec; ec.game_amnt = 4; ec.seed = 11;
    auto az = std::make_unique<AlphaZeroMcts>(); az->config.max_playouts = 64;
    FightStatistics fs = fight(ec, std::move(az), std::make_unique<Random>(5));
    CHECK(fs.wins + fs.losses + fs.draws == 4 && fs.games_red == 2 && fs.games_blue == 2);
    CHECK(fs.wins >= fs.losses);
    // error behaviour: the reference panics, the mirror throws
    bool threw = false;
    try { Deck d; d.rotate(4); } catch (const Error&) { threw = true; }
    CHECK(threw);
}

// A ConvResNetConfig{64, 21, 1} VarStore with deterministic pseudo-random values (names of net.rs:118-213)
static std::vector<std::pair<std::string, std::vector<float>>> fake_var_store() {
    std::vector<std::pair<std::string, std::vector<float>>> vs;
    uint32_t rng = 12345u;
    auto uni = [&]() { rng = rng * 1664525u + 1013904223u; return (float)((rng >> 8) & 0xFFFF) / 65536.f - 0.5f; };
    auto add = [&](const std::string& name, size_t n, float scale, float offset) {
        std::vector<float> v(n);
        for (float& x : v) x = uni() * scale + offset;
        vs.emplace_back(name, std::move(v));
    };
    auto conv_bn = [&](const std::string& conv, const std::string& bn, size_t c_out, size_t c_in, size_t k) {
        add(conv + "|weight", c_out * c_in * k * k, 2.8f / std::sqrt((float)(c_in * k * k)), 0.f);
        add(conv + "|bias", c_out, 0.2f, 0.f);
        add(bn + "|weight", c_out, 0.4f, 1.f);
        add(bn + "|bias", c_out, 0.4f, 0.1f);
        add(bn + "|running_mean", c_out, 0.4f, 0.f);
        add(bn + "|running_var", c_out, 0.8f, 1.f);
    };
    conv_bn("conv_init_1", "bn1", 64, 21, 3);
    conv_bn("resnet_0|resnet_small_block1|small_block_conv", "resnet_0|resnet_small_block1|small_block_bn", 64, 64, 3);
    conv_bn("resnet_0|resnet_small_block2|small_block_conv", "resnet_0|resnet_small_block2|small_block_bn", 64, 64, 3);
    conv_bn("vh_conv", "vh_bn", 1, 64, 1);
    conv_bn("policy_conv", "policy_bn", 2, 64, 1);
    add("vh_linear1|weight", 64 * 25, 0.8f, 0.f); add("vh_linear1|bias", 64, 0.6f, 0.f);
    add("vh_linear2|weight", 64, 0.5f, 0.f); add("vh_linear2|bias", 1, 0.2f, 0.f);
    add("ph_linear2|weight", 2500, 0.6f, 0.f); add("ph_linear2|bias", 50, 0.6f, 0.f);
    return vs;
}

// TrainingAlphaZeroMcts::generate_move_tensor with the network inside the library (ONB_EVAL_NET) must build the same tree as the
// split-phase search that gets the same network through the host-evaluator hook (here: the kernel on a second one-game engine).
static void device_network_search() {
    const auto vs = fake_var_store();
    Engine::single(64).load_network(vs);
    Engine side(1, 2, 0, 0, false);
    side.load_network(vs);
    HostEvaluator through_side = [&](const float* planes, int64_t n, float* policy, float* value) {
        for (int64_t i = 0; i < n; ++i) {
            side.check(onb_write_buffer(side.ctx(), ONB_BUF_LEAF_PLANES, planes + i * 525, 2100));
            side.check(onb_net_forward(side.ctx(), ONB_BUF_LEAF_PLANES));
            side.check(onb_read_buffer(side.ctx(), ONB_BUF_POLICY, policy + i * 50, 200));
            side.check(onb_read_buffer(side.ctx(), ONB_BUF_VALUE, value + i, 4));
        }
    };
    AlphaZeroMctsConfig cfg;
    cfg.max_playouts = 48;
    cfg.exploration_c = 2.0;
    State s = State::with_deck(Deck({DRAGON, FROG, TIGER, RABBIT, HORSE}));
    TrainingAlphaZeroMcts on_device{cfg, ONB_EVAL_NET, nullptr}, via_host{cfg, ONB_EVAL_UNIFORM, through_side};
    auto a = on_device.generate_move_tensor(s, PlayerColor::Red);
    auto b = via_host.generate_move_tensor(s, PlayerColor::Red);
    CHECK(a.first.mov.from == b.first.mov.from && a.first.mov.to == b.first.mov.to && a.first.used_card_idx == b.first.used_card_idx);
    float sum = 0.f;
    bool same = true;
    for (int i = 0; i < 50; ++i) { sum += a.second[i]; same = same && a.second[i] == b.second[i]; }
    CHECK(same);
    CHECK(std::fabs(sum - 1.f) < 1e-5f);
    bool threw = false;  // a VarStore of the wrong width is rejected, not silently mis-read
    try { auto bad = vs; bad[0].second.resize(10); side.load_network(bad); } catch (const Error&) { threw = true; }
    CHECK(threw);
}

// ai/mcts/mcts_arena.rs:459-554: the reference's known-answer searches for the `Mcts` agent, through the mirror
static void plain_mcts_reference_tests() {
    {   // test_best_move_win: Blue must take the Red king with the Dragon card
        Deck deck({RABBIT, FROG, TIGER, DRAGON, HORSE});
        GameState gs = GameState::with_deck(deck);
        gs.state.kings[0] = from_2d_to_bitboard(1, 3);
        gs.curr_player_color = PlayerColor::Blue;
        Mcts m;  // defaults of mod.rs:21-30: sqrt(2), 5 visits, 5000 playouts
        auto r = m.generate_move(gs);
        CHECK((r.first == DoneMove{Move{1, 8, PieceKind::Pawn}, 3}));
        CHECK(r.second > 0.9);
    }
    {   // test_no_way_to_hide_for_blue: the only move that does not lose is Rabbit e5-c5
        Deck deck({OX, MONKEY, RABBIT, HORSE, DRAGON});
        GameState gs = GameState::with_deck(deck);
        gs.state.kings[1] = from_2d_to_bitboard(0, 4);
        gs.state.pawns[1] = 0;
        gs.state.pawns[0] = from_2d_to_bitboard(0, 3) | from_2d_to_bitboard(1, 4);
        gs.curr_player_color = PlayerColor::Blue;
        Mcts m; m.exploration_c = 2.0f;
        CHECK((m.generate_move(gs).first == DoneMove{Move{4, 2, PieceKind::King}, 2}));
    }
    {   // test_worst_case_capture_blue
        Deck deck({MONKEY, ROOSTER, GOOSE, MANTIS, HORSE});
        GameState gs = GameState::with_deck(deck);
        gs.state.kings[1] = from_2d_to_bitboard(1, 2);
        gs.state.pawns[0] = from_2d_to_bitboard(2, 3) | from_2d_to_bitboard(3, 2);
        gs.curr_player_color = PlayerColor::Blue;
        Mcts m; m.exploration_c = 1.0f; m.min_node_visits = 1;
        CHECK((m.generate_move(gs).first == DoneMove{Move{7, 2, PieceKind::King}, 3}));
    }
}

// self_play_continuous: the loop inside the library; sample invariants of train.rs:55-88 and the quota
static void native_self_play() {
    Engine e(16, 24, 5);
    TrainingAlphaZeroMcts sp;
    sp.config.max_playouts = 24;
    sp.config.exploration_c = 2.0;
    TrainConfig tc;
    tc.self_play_game_amnt = 40;
    tc.max_plies = 20;
    auto data = self_play_continuous(e, sp, tc);
    CHECK(!data.empty() && data.size() <= 16u * 22u * 40u);   // exactly 40 complete games of at most 22 plies
    size_t decided = 0;
    for (auto& d : data) {
        float ps = 0.f; for (float p : d.pi) ps += p;
        CHECK(std::fabs(ps - 1.f) < 1e-5f);
        CHECK(d.z == 0.f || d.z == 1.f || d.z == -1.f);
        decided += d.z != 0.f;
        CHECK(d.state[20 * 25] == (d.player_color == PlayerColor::Blue ? 1.f : 0.f));
    }
    CHECK(decided > 0);
    bool threw = false;  // a host evaluator cannot run inside the library
    sp.model = [](const float*, int64_t, float*, float*) {};
    try { self_play_continuous(e, sp, tc); } catch (const Error&) { threw = true; }
    CHECK(threw);
}

// state.rs:397-416 correct_display (host-only code)
static void correct_display() {
    const std::string expected =
        "\n---+---+---+---+---+---+\n 5 | b | b | B | b | b |\n---+---+---+---+---+---+\n 4 | . | . | . | . | . |\n---+---+---+---+---+---+\n"
        " 3 | . | . | . | . | . |\n---+---+---+---+---+---+\n 2 | . | . | . | . | . |\n---+---+---+---+---+---+\n 1 | r | r | R | r | r |\n"
        "---+---+---+---+---+---+\n   | a | b | c | d | e |\n";
    CHECK("\n" + State::with_deck(Deck()).display() + "\n" == expected);
}

// fight_device: the arena inside the library; PUCT (uniform evaluator) must beat the Random agent, plain UCT must beat it too
static void native_fight() {
    Engine e(24, 200, 9, 0, false);
    EvaluatorConfig ec;
    ec.game_amnt = 24;
    onb_agent puct{ONB_AGENT_PUCT, ONB_EVAL_UNIFORM, 0, 64, 2.0, 0, 0}, rnd{ONB_AGENT_RANDOM, 0, 0, 0, 0., 0, 0}, uct{ONB_AGENT_UCT, 0, 0, 200, 1.41421356, 5, 0};
    FightStatistics a = fight_device(e, ec, puct, rnd);
    CHECK(a.wins + a.losses + a.draws == 24 && a.games_red == 12 && a.games_blue == 12);
    CHECK(a.wins > a.losses);
    // FightStatistics incl. Elo come from the device fold (onb_fight_stats): the host fold of the same per-game results
    // (evaluator.rs:58-110) must agree -- counts exactly, ratings to rounding (device pow vs libm pow)
    FightStatistics host(800., 800.);
    for (size_t g = 0; g < 24; ++g) {
        const uint8_t r = e.last_fight_results[g];
        host.update(r == 1 ? MoveResult::RedWin : r == 2 ? MoveResult::BlueWin : MoveResult::InProgress, g % 2 == 0 ? PlayerColor::Red : PlayerColor::Blue);
    }
    CHECK(host.wins == a.wins && host.losses == a.losses && host.draws == a.draws && host.wins_red == a.wins_red && host.wins_blue == a.wins_blue);
    CHECK(std::fabs(host.rating_a - a.rating_a) < 1e-9 && std::fabs(host.rating_b - a.rating_b) < 1e-9 && a.rating_a > 800. && a.rating_b < 800.);
    CHECK(a.rating_change_history.size() == 24 && a.rating_change_history[0].before_a == 800.);
    for (size_t g = 0; g < 24; ++g) CHECK(std::fabs(host.rating_change_history[g].after_a - a.rating_change_history[g].after_a) < 1e-9);
    CHECK(std::fabs(host.winrate - a.winrate) < 1e-15);
    FightStatistics b = fight_device(e, ec, uct, rnd, 900., 700.);
    CHECK(b.wins + b.losses + b.draws == 24 && b.wins > b.losses && b.rating_change_history[0].before_a == 900.);
    bool threw = false;
    onb_agent bad = puct; bad.sims = 100000;   // more simulations than the engine was created for
    try { fight_device(e, ec, bad, rnd); } catch (const Error&) { threw = true; }
    CHECK(threw);
}

int main() {
    try {
        correct_display();
        create_all_legal_moves_for_red_in_starting_position();
        create_all_legal_moves_for_blue_in_starting_position();
        make_move_tests();
        no_legal_moves();
        expand_order();
        search_and_drivers();
        device_network_search();
        plain_mcts_reference_tests();
        native_self_play();
        native_fight();
    } catch (const Error& e) {
        std::printf("FAIL exception %d: %s\n", e.code, e.what());
        return 2;
    }
    if (failures) { std::printf("%d FAILURES\n", failures); return 1; }
    std::printf("ALL OK\n");
    return 0;
}
