// onb_internal.h -- context layout and launcher prototypes shared by the translation units of libonb.so.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include <string>

#include "../../include/onb.h"

namespace onb {

// One search-tree node: exactly one 32-byte sector. A node's children occupy CONTIGUOUS slots (the
// reference pushes them back to back into its arena, mcts_arena.rs:231-260), so a warp reads a whole
// children block with one coalesced request and every lane gets (N, W, P, header) of its child at once.
struct __align__(32) Node {
    double w;              // MctsNode::reward (sum of backed-up values)      mcts_arena.rs:366-367
    double p;              // MctsNode::probability                           mcts_arena.rs:372
    uint32_t n;            // MctsNode::visits                                mcts_arena.rs:364-365
    uint32_t first_child;  // index of child 0 in this tree's pool
    uint32_t parent;       // 0xFFFFFFFF for the root
    uint16_t action;       // move that leads here (onb_action)
    uint8_t n_child;
    uint8_t flags;         // bit0 expanded, bit1 terminal, bit2 pass pseudo-child
};
static_assert(sizeof(Node) == 32, "node must be one sector");
constexpr uint8_t kNodeExpanded = 1, kNodeTerminal = 2, kNodePass = 4;
constexpr uint8_t kTreePassSeen = 1, kTreeOverflow = 2;
constexpr int kMaxDepth = 32;  // path entries kept in registers (lane l <-> level l); deeper paths spill to the parent chain

struct Ctx {
    onb_config cfg;
    cudaStream_t stream;
    bool own_stream;
    int64_t n;
    // env
    uint4* d_states;
    uint32_t* d_masks;
    float* d_planes;
    uint16_t* d_actions;
    unsigned long long* d_stats;
    int32_t fixed_cards;  // -1: auto-reset deals from the RNG, else the nibble-packed deck every reset re-deals (onb_env_reset with one deck)
    // scratch for host<->device transfers of boundary structs
    onb_state* d_io_states;
    uint16_t* d_moves;  // [n][40]
    uint8_t* d_counts;  // [n]
    // mcts
    Node* d_nodes;          // [n][node_cap]
    uint32_t node_cap;
    uint32_t* d_tree_size;  // [n]
    uint8_t* d_tree_flags;  // [n]
    uint4* d_roots;         // [n] root states captured by mcts_begin
    uint32_t* d_leaf_node;  // [n] leaf reached by the last select
    uint4* d_leaf_state;    // [n]
    float* d_leaf_planes;   // [n][525]
    float* d_policy;        // [n][50]
    float* d_value;         // [n]
    float* d_pi;            // [n][50]
    uint16_t* d_best;       // [n]
    uint32_t* d_root_visits;
    double* d_root_q;
    uint32_t* d_child_visits;  // [n][40]
    double c_puct;
    int noise_on;           // train mode: Dirichlet-like exploration noise at the root (mcts_arena.rs:186-202)
    double noise_eps, noise_alpha;
    uint64_t noise_seed;
    uint32_t sims_target, sims_done;
    int mcts_phase;  // 0 idle, 1 begun/after expand, 2 after select
    // subset search (onb_fight / onb_self_play): the first n_act trees search the games listed in d_tree_game (device, ascending);
    // n_act == 0: every game is searched by the tree of the same index (the public onb_mcts_begin always selects this)
    int64_t n_act;
    const int32_t* d_tree_game;
    float* d_ln_table;  // logf(i) for the plain UCT search (filled by the host: the libm call behind Rust's f32::ln)
    uint32_t ln_cap;
    // policy/value network (onb_net.cu): weights in tensor-core operand layout, folded biases, head parameters
    struct NetSlot {
        float* w;     // conv taps in operand layout
        float* bias;  // folded biases
        float* head;  // head parameters
        int blocks, loaded, f16, x3;  // f16: f16 operands (else tf32); x3: split-operand f32-faithful mode (ONB_NET_F32)
        size_t pair_off;              // x3: byte offset in w of the taps in CTA-pair order (onb_net.cu, k_net_forward_x3p<true>)
    } net[2];         // two networks can be resident (an arena pits the new model against the previous one, evaluator.rs:355-399)
    void* sp_buf[32];          // grow-only buffers of onb_self_play (slots 0-15) and onb_fight (16-31), see onb_selfplay.cu
    size_t sp_cap[32];
    int fight_valid;           // an onb_fight has left its per-game results on the device (onb_fight_statistics)
    void* d_net_scratch;       // residual scratch of the three-CTAs-per-SM network kernel
    size_t net_scratch_bytes;
    int net_cur;      // slot used by onb_net_load / onb_net_forward / ONB_EVAL_NET (onb_net_select)
    int net_tf32;     // requested arithmetic for the next onb_net_load: ONB_NET_F16 (default) | ONB_NET_TF32 | ONB_NET_F32
    // grow-only device scratch reused across onb_perft calls (counters, cursor, two ping-pong frontiers): repeated
    // cudaMalloc/cudaFree of several hundred MB made the call time vary by +-50 %
    void* scratch[16];
    size_t scratch_cap[16];
    char err[512];
};

inline int64_t trees(const Ctx* c) { return c->n_act > 0 ? c->n_act : c->n; }

// env (onb_env.cu)
cudaError_t launch_env_reset(Ctx* c, const uint8_t* d_decks5, int64_t n_decks, uint32_t epoch);
cudaError_t launch_env_reset_where(Ctx* c, const uint8_t* d_mask, uint32_t epoch, unsigned long long* d_count);
cudaError_t launch_states_export(Ctx* c, onb_state* d_out, int64_t first, int64_t n);
cudaError_t launch_states_import(Ctx* c, const onb_state* d_in, int64_t first, int64_t n);
cudaError_t launch_legal_moves(Ctx* c);
cudaError_t launch_observe(Ctx* c, uint32_t out_flags);
cudaError_t launch_env_step(Ctx* c, int mode, uint32_t step, int auto_reset, uint32_t out_flags);
// the same on games [first, first + count) and a stream of the caller's choice (first % 64 == 0); done_bits: [count/32][2] or nullptr
cudaError_t launch_env_step_slice(Ctx* c, int mode, uint32_t step, int auto_reset, uint32_t out_flags, int64_t first, int64_t count,
                                  cudaStream_t stream, uint32_t* done_bits);
// perft (onb_perft.cu)
int32_t run_perft(Ctx* c, const onb_state* roots_host, int64_t n, int depth, uint64_t* nodes, uint64_t* wins, uint64_t* zero);
// mcts (onb_mcts.cu)
cudaError_t launch_selftest_div(Ctx* c, unsigned long long* d_mismatches);
cudaError_t launch_mcts_begin(Ctx* c);
cudaError_t launch_mcts_select(Ctx* c);
cudaError_t launch_mcts_expand_backup(Ctx* c);
cudaError_t launch_mcts_step(Ctx* c);  // expand_backup of this simulation + select of the next one, one launch
cudaError_t launch_mcts_eval(Ctx* c, int evaluator);
cudaError_t launch_mcts_run(Ctx* c, int evaluator, uint32_t sims);
cudaError_t launch_uct_run(Ctx* c, float exploration_c, uint32_t min_node_visits, uint32_t sims);
cudaError_t launch_mcts_finish(Ctx* c);
cudaError_t launch_mcts_play_best(Ctx* c, uint32_t out_flags);
cudaError_t launch_mcts_scatter_best(Ctx* c, uint16_t* dst_actions);  // dst[game of tree t] = best[t]

// network (onb_net.cu)
int32_t net_load(Ctx* c, int32_t n_tensors, const char* const* names, const float* const* data, const int64_t* numel, std::string& err);
cudaError_t launch_net_forward(Ctx* c, const float* planes, float* policy, float* value, int64_t count);  // count positions

// native self-play driver (onb_selfplay.cu)
int32_t run_self_play(Ctx* c, const onb_selfplay_config* cfg, onb_selfplay_result* out,
                      int32_t (*search)(Ctx*, const onb_selfplay_config*, const int32_t* d_games, int64_t m), char* err, size_t err_len);

int32_t run_fight(Ctx* c, const onb_agent* a, const onb_agent* b, const uint8_t* a_is_red_host, uint32_t max_plies, onb_fight_result* out,
                  int32_t (*move)(Ctx*, const onb_agent*, uint32_t ply, const int32_t* d_games, int64_t m, uint16_t* d_out), char* err, size_t err_len);
int32_t run_fight_statistics(Ctx* c, double rating_a, double rating_b, onb_fight_statistics* out, double* history_host, char* err, size_t err_len);

constexpr int kModeActions = 2;  // env step modes: 0 = ONB_POLICY_UNIFORM, 1 = ONB_POLICY_AGENT, 2 = explicit actions

}  // namespace onb
