// onb_selfplay.cu -- self_play (alphazero-training/src/train.rs:35-98) for all games of a context, natively: search, sample
// recording, move, game-over bookkeeping and the restart of finished slots all stay on the device; the host only reads one counter
// per ply to know when the requested number of games is complete. A worker of the reference plays its games one after the other
// (train.rs:218-245); here every slot starts its next game the moment the previous one is over, so all slots stay busy.
//
// Sample i = tick * n + slot (ply-major, every slot records every ply): planes = create_tensor_from_state of the searched position
// (train.rs:58), pi = the search's visit distribution, colour = side to move, serial = slot + n * (games already finished in the
// slot). When a game ends, z = reward(final result, sample colour) (train.rs:83-85, alphazero_mcts/mod.rs:45-53; 0 for a game cut
// at the ply cap) is written to the game's samples (the last `plies` samples of the slot: i, i - n, i - 2n, ...) and they become
// valid; samples of games still running when the quota is reached stay invalid.
#include <vector>

#include "onb_internal.h"
#include "onb_rules.cuh"

namespace onb {
namespace {

__global__ void __launch_bounds__(256) k_sp_record(const uint4* __restrict__ states, int64_t n, const int32_t* __restrict__ generation, int64_t base,
                                                   uint8_t* __restrict__ color, int64_t* __restrict__ serial, uint8_t* __restrict__ valid,
                                                   float* __restrict__ z) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const Game g = unpack(states[s]);
    color[base + s] = (uint8_t)g.side;
    serial[base + s] = s + n * (int64_t)generation[s];
    valid[base + s] = 0;
    z[base + s] = 0.f;
}

// after the move: close finished games (write z to their samples, mark them valid), restart their slots
__global__ void __launch_bounds__(256) k_sp_close(uint4* __restrict__ states, int64_t n, int32_t* __restrict__ generation, int32_t* __restrict__ plies,
                                                  int64_t base, uint32_t max_plies, const uint8_t* __restrict__ color, uint8_t* __restrict__ valid,
                                                  float* __restrict__ z, unsigned long long* __restrict__ done, int32_t fixed_cards, uint64_t seed,
                                                  uint64_t game0, uint32_t epoch) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const Game g = unpack(states[s]);
    const int32_t p = plies[s] + 1;
    // train.rs:74-79: the cap is checked after the move while max_plies counts down from 150: a game has at most max_plies + 2 plies
    const bool over = g.result != 0 || (uint32_t)p >= max_plies + 2u;
    if (!over) {
        plies[s] = p;
        return;
    }
    for (int32_t k = 0; k < p; ++k) {  // the game's samples: this ply's and the p - 1 before it
        const int64_t i = base + s - (int64_t)k * n;
        z[i] = g.result == 0 ? 0.f : ((g.result - 1u) == (uint32_t)color[i] ? 1.f : -1.f);
        valid[i] = 1;
    }
    states[s] = pack(start_game(fixed_cards >= 0 ? (uint32_t)fixed_cards : deal_cards(game_key(seed, game0 + (uint64_t)s), epoch)));
    generation[s] += 1;
    plies[s] = 0;
    atomicAdd(done, 1ull);
}

// arena: pick agent A's or agent B's action per game (agent A moves when the side to move is the colour it plays in that game)
__global__ void __launch_bounds__(256) k_fight_merge(const uint4* __restrict__ states, int64_t n, const uint8_t* __restrict__ a_is_red,
                                                     const uint16_t* __restrict__ act_a, const uint16_t* __restrict__ act_b, uint16_t* __restrict__ actions) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Game g = unpack(states[i]);
    const bool a_to_move = (g.side == 0u) == (a_is_red[i] != 0);
    actions[i] = g.result == 0 ? (a_to_move ? act_a[i] : act_b[i]) : (uint16_t)0xFFFFu;
}
__global__ void __launch_bounds__(256) k_count_live(const uint4* __restrict__ states, int64_t n, unsigned long long* __restrict__ live) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool alive = i < n && unpack(states[i]).result == 0;
    const unsigned b = __ballot_sync(0xFFFFFFFFu, alive);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(live, (unsigned long long)__popc(b));
}
__global__ void __launch_bounds__(256) k_fight_tally(const uint4* __restrict__ states, int64_t n, const uint8_t* __restrict__ a_is_red,
                                                     uint8_t* __restrict__ results, unsigned long long* __restrict__ tally) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t r = unpack(states[i]).result;
    results[i] = (uint8_t)r;
    if (r == 0) atomicAdd(&tally[2], 1ull);                                             // draw: the ply cap ended the game
    else atomicAdd(&tally[((r == 1u) == (a_is_red[i] != 0)) ? 0 : 1], 1ull);            // agent A wins when its colour won
}

cudaError_t grow(Ctx* c, int slot, size_t bytes, void** out) {
    if (bytes == 0) bytes = 16;
    if (c->sp_cap[slot] < bytes) {
        if (c->sp_buf[slot]) cudaFree(c->sp_buf[slot]);
        c->sp_buf[slot] = nullptr;
        c->sp_cap[slot] = 0;
        const cudaError_t e = cudaMalloc(&c->sp_buf[slot], bytes);
        if (e != cudaSuccess) return e;
        c->sp_cap[slot] = bytes;
    }
    *out = c->sp_buf[slot];
    return cudaSuccess;
}

}  // namespace

// Returns ONB_OK, or an ONB_E_* code with `err` filled. `evaluate_round` runs one search for all slots (begin / run / finish).
int32_t run_self_play(Ctx* c, const onb_selfplay_config* cfg, onb_selfplay_result* out, int32_t (*search)(Ctx*, const onb_selfplay_config*),
                      char* err, size_t err_len) {
    const int64_t n = c->n;
    const int64_t cap_ticks = cfg->sample_cap / n;
    if (cap_ticks < 1) {
        snprintf(err, err_len, "onb_self_play: sample_cap %lld is smaller than one ply of %lld games", (long long)cfg->sample_cap, (long long)n);
        return ONB_E_INVALID;
    }
    const size_t cap = (size_t)cap_ticks * (size_t)n;
    float *planes = nullptr, *pi = nullptr, *z = nullptr;
    uint8_t *color = nullptr, *valid = nullptr;
    int64_t *serial = nullptr, *idx = nullptr;
    int32_t *generation = nullptr, *plies = nullptr;
    unsigned long long* done = nullptr;
    cudaError_t e;
#define SP(call)                                                                       \
    do {                                                                               \
        e = (call);                                                                    \
        if (e != cudaSuccess) {                                                        \
            snprintf(err, err_len, "onb_self_play: %s: %s", #call, cudaGetErrorString(e)); \
            return e == cudaErrorMemoryAllocation ? ONB_E_NOMEM : ONB_E_CUDA;          \
        }                                                                              \
    } while (0)
    SP(grow(c, 0, cap * 2100, (void**)&planes));
    SP(grow(c, 1, cap * 200, (void**)&pi));
    SP(grow(c, 2, cap * 4, (void**)&z));
    SP(grow(c, 3, cap, (void**)&color));
    SP(grow(c, 4, cap, (void**)&valid));
    SP(grow(c, 5, cap * 8, (void**)&serial));
    SP(grow(c, 6, (size_t)n * 4, (void**)&generation));
    SP(grow(c, 7, (size_t)n * 4, (void**)&plies));
    SP(grow(c, 8, 8, (void**)&done));
    SP(cudaMemsetAsync(generation, 0, (size_t)n * 4, c->stream));
    SP(cudaMemsetAsync(plies, 0, (size_t)n * 4, c->stream));
    SP(cudaMemsetAsync(done, 0, 8, c->stream));
    c->fixed_cards = -1;                      // like onb_env_reset without decks: every game is dealt from the counter RNG
    SP(launch_env_reset(c, nullptr, 0, 0));
    const unsigned grid = (unsigned)((n + 255) / 256);
    unsigned long long finished = 0;
    int64_t tick = 0;
    int truncated = 0;
    while ((int64_t)finished < cfg->n_games) {
        if (tick >= cap_ticks) {
            truncated = 1;
            break;
        }
        const int64_t base = tick * n;
        SP(launch_observe(c, ONB_OUT_PLANES));  // planes of the position the search starts from
        SP(cudaMemcpyAsync(planes + (size_t)base * 525, c->d_planes, (size_t)n * 2100, cudaMemcpyDeviceToDevice, c->stream));
        k_sp_record<<<grid, 256, 0, c->stream>>>(c->d_states, n, generation, base, color, serial, valid, z);
        SP(cudaGetLastError());
        const int32_t rc = search(c, cfg);
        if (rc != ONB_OK) {
            snprintf(err, err_len, "%s", c->err);
            return rc;
        }
        SP(cudaMemcpyAsync(pi + (size_t)base * 50, c->d_pi, (size_t)n * 200, cudaMemcpyDeviceToDevice, c->stream));
        SP(launch_mcts_play_best(c, 0));
        k_sp_close<<<grid, 256, 0, c->stream>>>(c->d_states, n, generation, plies, base, cfg->max_plies, color, valid, z, done, c->fixed_cards,
                                                c->cfg.seed, c->cfg.game_id_base, (uint32_t)(tick + 1));
        SP(cudaGetLastError());
        SP(cudaMemcpyAsync(&finished, done, 8, cudaMemcpyDeviceToHost, c->stream));
        SP(cudaStreamSynchronize(c->stream));
        ++tick;
    }
    // index list of the valid samples (ascending): flags to the host, indices back
    const size_t total = (size_t)tick * (size_t)n;
    std::vector<uint8_t> flags(total);
    std::vector<int64_t> keep;
    if (total) SP(cudaMemcpy(flags.data(), valid, total, cudaMemcpyDeviceToHost));
    keep.reserve(total);
    for (size_t i = 0; i < total; ++i)
        if (flags[i]) keep.push_back((int64_t)i);
    SP(grow(c, 9, keep.size() * 8, (void**)&idx));
    if (!keep.empty()) SP(cudaMemcpy(idx, keep.data(), keep.size() * 8, cudaMemcpyHostToDevice));
#undef SP
    out->n_samples = (int64_t)total;
    out->n_valid = (int64_t)keep.size();
    out->n_games = (int64_t)finished;
    out->plies_run = tick;
    out->truncated = truncated;
    out->planes = planes;
    out->pi = pi;
    out->z = z;
    out->color = color;
    out->serial = serial;
    out->valid_idx = idx;
    c->mcts_phase = 0;
    return ONB_OK;
}

// fight (evaluator.rs:355-399) for all games of the context in lockstep: `move` leaves an agent's actions for every game in
// d_actions; the per-game choice, the step and the end-of-game test stay on the device (one 8-byte counter per ply is read back).
int32_t run_fight(Ctx* c, const onb_agent* a, const onb_agent* b, const uint8_t* a_is_red_host, uint32_t max_plies, onb_fight_result* out,
                  int32_t (*move)(Ctx*, const onb_agent*, uint32_t), char* err, size_t err_len) {
    const int64_t n = c->n;
    uint8_t *mask = nullptr, *results = nullptr;
    uint16_t *act_a = nullptr, *act_b = nullptr;
    unsigned long long* counters = nullptr;  // [0] live games of the current ply, [1..3] tally
    cudaError_t e;
#define FT(call)                                                                    \
    do {                                                                            \
        e = (call);                                                                 \
        if (e != cudaSuccess) {                                                     \
            snprintf(err, err_len, "onb_fight: %s: %s", #call, cudaGetErrorString(e)); \
            return e == cudaErrorMemoryAllocation ? ONB_E_NOMEM : ONB_E_CUDA;       \
        }                                                                           \
    } while (0)
    FT(grow(c, 3, (size_t)n, (void**)&mask));       // the arena reuses the self-play scratch slots (a context runs one driver at a time)
    FT(grow(c, 4, (size_t)n, (void**)&results));
    FT(grow(c, 6, (size_t)n * 2, (void**)&act_a));
    FT(grow(c, 7, (size_t)n * 2, (void**)&act_b));
    FT(grow(c, 8, 32, (void**)&counters));
    FT(cudaMemcpyAsync(mask, a_is_red_host, (size_t)n, cudaMemcpyHostToDevice, c->stream));
    const unsigned grid = (unsigned)((n + 255) / 256);
    int64_t plies_left = (int64_t)max_plies, tick = 0;
    for (;;) {
        unsigned long long live = 0;
        FT(cudaMemsetAsync(counters, 0, 8, c->stream));
        k_count_live<<<grid, 256, 0, c->stream>>>(c->d_states, n, counters);
        FT(cudaGetLastError());
        FT(cudaMemcpyAsync(&live, counters, 8, cudaMemcpyDeviceToHost, c->stream));
        FT(cudaStreamSynchronize(c->stream));
        if (live == 0) break;                       // evaluator.rs:366: every game is decided
        int32_t rc = move(c, a, (uint32_t)tick);
        if (rc != ONB_OK) { snprintf(err, err_len, "%s", c->err); return rc; }
        FT(cudaMemcpyAsync(act_a, c->d_actions, (size_t)n * 2, cudaMemcpyDeviceToDevice, c->stream));
        rc = move(c, b, (uint32_t)tick);
        if (rc != ONB_OK) { snprintf(err, err_len, "%s", c->err); return rc; }
        FT(cudaMemcpyAsync(act_b, c->d_actions, (size_t)n * 2, cudaMemcpyDeviceToDevice, c->stream));
        k_fight_merge<<<grid, 256, 0, c->stream>>>(c->d_states, n, mask, act_a, act_b, c->d_actions);
        FT(cudaGetLastError());
        FT(launch_env_step(c, kModeActions, 0, 0, 0));
        ++tick;
        if (plies_left < 0) break;                  // evaluator.rs:386-392: checked after the move, then decremented (max_plies + 2 plies)
        plies_left -= 1;
    }
    FT(cudaMemsetAsync(counters + 1, 0, 24, c->stream));
    k_fight_tally<<<grid, 256, 0, c->stream>>>(c->d_states, n, mask, results, counters + 1);
    FT(cudaGetLastError());
    unsigned long long tally[3] = {0, 0, 0};
    FT(cudaMemcpyAsync(tally, counters + 1, 24, cudaMemcpyDeviceToHost, c->stream));
    FT(cudaStreamSynchronize(c->stream));
#undef FT
    out->a_wins = (int64_t)tally[0];
    out->b_wins = (int64_t)tally[1];
    out->draws = (int64_t)tally[2];
    out->plies_run = tick;
    out->results = results;
    c->mcts_phase = 0;
    return ONB_OK;
}

}  // namespace onb
