// onb_env.cu -- lockstep batched Onitama dynamics: reset, legal moves/masks, step (explicit or random
// actions), terminal detection, auto-reset and 21x5x5 plane encoding, one thread per game for the rules
// and CTA-cooperative, fully coalesced 16-byte stores for the planes.
//
// Roofline: HBM. Algorithmic bytes per env step: 16 B state read + 16 B state write
// + 8 B legal mask + 2 100 B planes = 2 140 B (+2 B action when requested).
#include "onb_internal.h"
#include "onb_rules.cuh"

namespace onb {

constexpr int kTile = 128;  // games per CTA tile == threads per CTA
constexpr int kPlanes = 21;
constexpr int kPlaneFloats = 525;

// ------------------------------------------------------------------------------------------ reset
__global__ void __launch_bounds__(256) k_env_reset(uint4* __restrict__ states, int64_t n, const uint8_t* __restrict__ decks5,
                                                   int64_t n_decks, uint64_t seed, uint64_t game0, uint32_t epoch) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t cards;
    if (n_decks == 0) {
        cards = deal_cards(game_key(seed, game0 + (uint64_t)i), epoch);
    } else {
        const uint8_t* d = decks5 + (n_decks == 1 ? 0 : 5 * i);
        cards = (d[0] & 15u) | ((d[1] & 15u) << 4) | ((d[2] & 15u) << 8) | ((d[3] & 15u) << 12) | ((d[4] & 15u) << 16);
    }
    states[i] = pack(start_game(cards));
}

// re-deal selected games in place: mask[i] != 0, or (mask == nullptr) every game that is over
__global__ void __launch_bounds__(256) k_env_reset_where(uint4* __restrict__ states, int64_t n, const uint8_t* __restrict__ mask, int32_t fixed_cards,
                                                         uint64_t seed, uint64_t game0, uint32_t epoch, unsigned long long* __restrict__ count) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool hit = false;
    if (i < n) {
        hit = mask ? mask[i] != 0 : unpack(states[i]).result != 0;
        if (hit) states[i] = pack(start_game(fixed_cards >= 0 ? (uint32_t)fixed_cards : deal_cards(game_key(seed, game0 + (uint64_t)i), epoch)));
    }
    const unsigned b = __ballot_sync(0xFFFFFFFFu, hit);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(count, (unsigned long long)__popc(b));
}

// ------------------------------------------------------------------------------------------ boundary conversion
__global__ void __launch_bounds__(256) k_states_export(const uint4* __restrict__ states, uint32_t* __restrict__ out6, int64_t first, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Game g = unpack(states[first + i]);
    uint32_t* o = out6 + 6 * i;
    o[0] = __brev(g.pawn_r); o[1] = __brev(g.pawn_b); o[2] = __brev(g.king_r); o[3] = __brev(g.king_b);
    o[4] = card_at(g.cards, 0) | (card_at(g.cards, 1) << 8) | (card_at(g.cards, 2) << 16) | (card_at(g.cards, 3) << 24);
    o[5] = card_at(g.cards, 4) | (g.side << 8) | (g.result << 16) | (g.passed << 24);
}
__global__ void __launch_bounds__(256) k_states_import(uint4* __restrict__ states, const uint32_t* __restrict__ in6, int64_t first, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t* o = in6 + 6 * i;
    Game g;
    g.pawn_r = __brev(o[0]) & kAll25; g.pawn_b = __brev(o[1]) & kAll25; g.king_r = __brev(o[2]) & kAll25; g.king_b = __brev(o[3]) & kAll25;
    g.cards = (o[4] & 15u) | (((o[4] >> 8) & 15u) << 4) | (((o[4] >> 16) & 15u) << 8) | (((o[4] >> 24) & 15u) << 12) | ((o[5] & 15u) << 16);
    g.side = (o[5] >> 8) & 1u;
    g.result = (o[5] >> 16) & 3u;
    g.passed = (o[5] >> 24) & 1u;
    states[first + i] = pack(g);
}

// ------------------------------------------------------------------------------------------ full move lists (parity / perft-style checks)
__global__ void __launch_bounds__(kTile) k_legal_moves(const uint4* __restrict__ states, int64_t n, uint16_t* __restrict__ moves40,
                                                       uint8_t* __restrict__ counts) {
    __shared__ uint32_t s_att[800];
    load_attack_table_to_smem(s_att);
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Game g = unpack(states[i]);
    const uint32_t side = g.side;
    const uint32_t own_p = side ? g.pawn_b : g.pawn_r, own_k = side ? g.king_b : g.king_r, own = own_p | own_k;
    uint32_t cnt = 0;
    for (uint32_t s = 0; s < 2; ++s) {
        const uint32_t idx = side * 2u + s;
        const uint32_t* Ts = s_att + (side * 16u + card_at(g.cards, idx)) * 25u;
        uint32_t rem = own;
        while (rem) {
            const int f = __ffs(rem) - 1;
            rem &= rem - 1;
            uint32_t a = Ts[f] & ~own;
            while (a) {
                const uint32_t to = __ffs(a) - 1;
                a &= a - 1;
                if (cnt < 40) moves40[i * 40 + cnt] = (uint16_t)make_action(idx, (uint32_t)f, to, ((own_p >> f) & 1u) ^ 1u);
                ++cnt;
            }
        }
    }
    for (uint32_t k = cnt; k < 40; ++k) moves40[i * 40 + k] = 0xFFFFu;
    counts[i] = (uint8_t)(cnt > 255 ? 255 : cnt);
}

// ------------------------------------------------------------------------------------------ the step kernel
// MODE 0: ONB_POLICY_UNIFORM, 1: ONB_POLICY_AGENT, 2: explicit actions, 3: observe only (no transition)
__device__ __forceinline__ uint32_t choose_action(const uint32_t* T, const Game& g, int mode, uint64_t key, uint32_t step) {
    const uint32_t side = g.side;
    const MoveSummary s = summarize_moves(T, g, side);
    if (mode == 0) {  // ai/mcts/mcts_arena.rs:190-241
        const uint32_t total = s.n0 + s.n1;
        if (total == 0) return kPassBit | ((side * 2u + rand_index(rand_from_key(key, step, kDrawPass), 2u)) << 10);
        uint32_t r = rand_index(rand_from_key(key, step, kDrawMove), total);
        const uint32_t slot01 = r >= s.n0 ? 1u : 0u;
        r -= slot01 ? s.n0 : 0u;
        return nth_move_of_slot(T, g, side, slot01, r);
    }
    // ai/random.rs:12-43: slot first, then a move of that card; fabricated a5->a4 pawn move if the card has none;
    // used_card_idx is the 0/1 slot number even for Blue (random.rs:39)
    const uint32_t card_idx = rand_index(rand_from_key(key, step, kDrawAgentSlot), 2u);
    const uint32_t cnt = card_idx ? s.n1 : s.n0;
    uint32_t a;
    if (cnt) a = nth_move_of_slot(T, g, side, card_idx, rand_index(rand_from_key(key, step, kDrawAgentMove), cnt));
    else a = make_action(0, 0, 5, 0);
    return (a & ~(3u << 10)) | (card_idx << 10);
}

// GAMES games per CTA are stepped by the first GAMES threads (one thread per game); when planes are written, ALL
// THREADS (>= GAMES) of the CTA then stream the tile's GAMES x 525 floats. A small tile drained by many threads keeps the
// chip-wide write front compact, which is what HBM wants (tools/wbench.cu: 8-32 games per CTA reach the memset ceiling,
// 128 games per 128-thread CTA lose 10-15 %).
// plane-writing variant: 64 games per CTA on 640 threads, 3 CTAs per SM (<= 34 registers). Measured per 1 Mi-game step on B200:
// 512 thr x 4 CTAs 314.8 us, 384 x 5 317.0, 640 x 3 311.1, 320 x 6 319.8, 1024 x 2 382; without the min-blocks bound ptxas takes
// more registers and the kernel drops to 2 CTAs per SM (405 us).
#ifndef ONB_ENV_MINB
#define ONB_ENV_MINB 3
#endif
#ifndef ONB_ENV_THREADS
#define ONB_ENV_THREADS 640
#endif
template <int MODE, bool PLANES, int GAMES, int THREADS>
__global__ void __launch_bounds__(THREADS, (PLANES ? ONB_ENV_MINB : 1)) k_env_step(uint4* __restrict__ states, int64_t n, uint16_t* __restrict__ actions,
                                                       uint32_t* __restrict__ masks, float* __restrict__ planes,
                                                       unsigned long long* __restrict__ stats, uint64_t seed, uint64_t game0, uint32_t step,
                                                       int auto_reset, int32_t fixed_cards, uint32_t out_flags, int choose_only,
                                                       uint2* __restrict__ done_bits) {
    __shared__ __align__(16) uint32_t s_att[800];
    __shared__ uint32_t s_pl[PLANES ? GAMES * kPlanes + 1 : 1];
    __shared__ uint32_t s_stat[5];
    load_attack_table_to_smem(s_att);
    if (threadIdx.x < 5) s_stat[threadIdx.x] = 0;
    __syncthreads();

    const int64_t n_tiles = (n + GAMES - 1) / GAMES;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        if (threadIdx.x < GAMES) {  // warp-uniform: GAMES is a multiple of 32
            const int64_t i = tile * GAMES + threadIdx.x;
            const bool live = i < n;
            Game g{};
            bool stepped = false, passed = false;
            uint32_t res = 0;
            if (live) {
                g = unpack(states[i]);
                uint32_t a = 0xFFFFu;
                if (MODE == 2) {
                    a = actions[i];
                    // a host action naming a square above 24 would shift into the card / side / result bits of the packed state
                    // (the reference's layout only sets its 7 padding bits there): rejected, counted, game untouched
                    if (a != 0xFFFFu && !(a & kPassBit) && ((a & 31u) > 24u || ((a >> 5) & 31u) > 24u)) {
                        a = 0xFFFFu;
                        atomicAdd(&s_stat[4], 1u);
                    }
                }
                // explicit actions: ONB_ACTION_NONE leaves the game untouched (e.g. the best move of a tree whose root was decided)
                if (MODE != 3 && g.result == 0 && !(MODE == 2 && a == 0xFFFFu)) {
                    const uint64_t key = game_key(seed, game0 + (uint64_t)i);
                    if (MODE != 2) a = choose_action(s_att, g, MODE, key, step);
                    if (MODE != 2 && choose_only) {  // agents for the arena loop: pick, do not play
                        actions[i] = (uint16_t)a;
                    } else {
                        res = apply_move(g, a);
                        stepped = true;
                        passed = (a & kPassBit) != 0;
                        if (MODE != 2 && (out_flags & ONB_OUT_ACTIONS)) actions[i] = (uint16_t)a;
                        if (res && auto_reset) g = start_game(fixed_cards >= 0 ? (uint32_t)fixed_cards : deal_cards(key, step + 1u));
                        states[i] = pack(g);
                    }
                } else if (MODE != 3 && MODE != 2 && (out_flags & ONB_OUT_ACTIONS)) {
                    actions[i] = 0xFFFFu;
                }
                if (out_flags & ONB_OUT_MASKS) {
                    const MoveSummary s = summarize_moves(s_att, g, g.side);
                    reinterpret_cast<uint2*>(masks)[i] = make_uint2(__brev(s.m0), __brev(s.m1));
                }
            }
            if (MODE != 3) {
                const uint32_t b0 = __ballot_sync(0xFFFFFFFFu, stepped), b1 = __ballot_sync(0xFFFFFFFFu, res == 1),
                               b2 = __ballot_sync(0xFFFFFFFFu, res == 2), b3 = __ballot_sync(0xFFFFFFFFu, passed);
                if ((threadIdx.x & 31) == 0) {
                    // ONB_HOST_DONE: which of these 32 games were decided by this step (2 bits per game instead of a state read-back)
                    if (done_bits && tile * GAMES + threadIdx.x < n) done_bits[(tile * GAMES + threadIdx.x) >> 5] = make_uint2(b1, b2);
                    if (b0) atomicAdd(&s_stat[0], __popc(b0));
                    if (b1) atomicAdd(&s_stat[1], __popc(b1));
                    if (b2) atomicAdd(&s_stat[2], __popc(b2));
                    if (b3) atomicAdd(&s_stat[3], __popc(b3));
                }
            }
            if (PLANES) {  // stage the game's 21 plane words (create_tensor_from_state, common.rs:26-80)
                uint32_t* pl = s_pl + threadIdx.x * kPlanes;
                const uint32_t c0 = card_at(g.cards, g.side * 2u), c1 = card_at(g.cards, g.side * 2u + 1u);
                pl[0] = g.pawn_r; pl[1] = g.king_r; pl[2] = g.pawn_b; pl[3] = g.king_b;
#pragma unroll
                for (uint32_t c = 0; c < 16; ++c) pl[4 + c] = (c == c0 || c == c1) ? kAll25 : 0u;
                pl[20] = g.side ? kAll25 : 0u;
                if (threadIdx.x == 0) s_pl[GAMES * kPlanes] = 0;
            }
        }
        if (PLANES) {
            __syncthreads();
            // the tile's planes are one contiguous, 16-byte aligned run of GAMES x 525 floats: stream it as float4
            const int64_t cnt = (n - tile * GAMES) < GAMES ? (n - tile * GAMES) : GAMES;
            const uint32_t n_float = (uint32_t)cnt * kPlaneFloats;
            const uint32_t n_vec = n_float >> 2;
            float4* __restrict__ dst = reinterpret_cast<float4*>(planes + tile * (int64_t)GAMES * kPlaneFloats);
            for (uint32_t q = threadIdx.x; q < n_vec; q += THREADS) {
                const uint32_t e = q * 4u;
                const uint32_t G = e / 25u, r = e - G * 25u;
                const uint32_t v = (s_pl[G] >> r) | (s_pl[G + 1] << (25u - r));
                float4 f;
                f.x = __uint_as_float(0x3F800000u & (0u - (v & 1u)));
                f.y = __uint_as_float(0x3F800000u & (0u - ((v >> 1) & 1u)));
                f.z = __uint_as_float(0x3F800000u & (0u - ((v >> 2) & 1u)));
                f.w = __uint_as_float(0x3F800000u & (0u - ((v >> 3) & 1u)));
                __stcs(dst + q, f);
            }
            if (threadIdx.x < (n_float & 3u)) {  // ragged tail of the last tile
                const uint32_t e = n_vec * 4u + threadIdx.x;
                const uint32_t G = e / 25u, r = e - G * 25u;
                planes[tile * (int64_t)GAMES * kPlaneFloats + e] = (float)((s_pl[G] >> r) & 1u);
            }
            __syncthreads();
        }
    }
    if (MODE != 3) {
        __syncthreads();
        if (threadIdx.x < 5 && s_stat[threadIdx.x]) {
            const int slot = threadIdx.x == 0 ? ONB_STAT_STEPS : threadIdx.x == 1 ? ONB_STAT_RED_WINS : threadIdx.x == 2 ? ONB_STAT_BLUE_WINS
                           : threadIdx.x == 3 ? ONB_STAT_PASSES : ONB_STAT_BAD_ACTIONS;
            atomicAdd(&stats[slot], (unsigned long long)s_stat[threadIdx.x]);
            if (auto_reset && (threadIdx.x == 1 || threadIdx.x == 2)) atomicAdd(&stats[ONB_STAT_RESETS], (unsigned long long)s_stat[threadIdx.x]);
        }
    }
}

// ------------------------------------------------------------------------------------------ whole games in registers (BASELINE config 1)
// Every thread plays its game from the current state until it ends or `max_plies` more plies were played;
// step index = ply index, so the trajectory equals lockstep stepping with steps step0, step0+1, ...
__global__ void __launch_bounds__(kTile) k_env_playout(uint4* __restrict__ states, int64_t n, uint32_t* __restrict__ plies_out,
                                                       unsigned long long* __restrict__ trace_out, unsigned long long* __restrict__ stats,
                                                       uint64_t seed, uint64_t game0, uint32_t step0, uint32_t max_plies, int mode) {
    __shared__ uint32_t s_att[800];
    load_attack_table_to_smem(s_att);
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t ply = 0, passes = 0;
    uint32_t res = 0;
    if (i < n) {
        Game g = unpack(states[i]);
        const uint64_t key = game_key(seed, game0 + (uint64_t)i);
        uint64_t trace = 0;
        res = g.result;
        while (res == 0 && ply < max_plies) {
            const uint32_t a = choose_action(s_att, g, mode, key, step0 + ply);
            res = apply_move(g, a);
            passes += (a & kPassBit) ? 1u : 0u;
            trace = mix64(trace ^ (uint64_t)a);
            ++ply;
        }
        states[i] = pack(g);
        if (plies_out) plies_out[i] = ply;
        if (trace_out) trace_out[i] = trace;
    }
    // block-level statistics
    __shared__ unsigned long long s_acc[4];
    if (threadIdx.x < 4) s_acc[threadIdx.x] = 0;
    __syncthreads();
    uint32_t v0 = ply, v3 = passes;
    for (int o = 16; o; o >>= 1) { v0 += __shfl_xor_sync(0xFFFFFFFFu, v0, o); v3 += __shfl_xor_sync(0xFFFFFFFFu, v3, o); }
    const uint32_t b1 = __ballot_sync(0xFFFFFFFFu, ply > 0 && res == 1), b2 = __ballot_sync(0xFFFFFFFFu, ply > 0 && res == 2);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&s_acc[0], (unsigned long long)v0); atomicAdd(&s_acc[1], (unsigned long long)__popc(b1));
        atomicAdd(&s_acc[2], (unsigned long long)__popc(b2)); atomicAdd(&s_acc[3], (unsigned long long)v3);
    }
    __syncthreads();
    if (threadIdx.x < 4 && s_acc[threadIdx.x]) {
        const int slot = threadIdx.x == 0 ? ONB_STAT_STEPS : threadIdx.x == 1 ? ONB_STAT_RED_WINS : threadIdx.x == 2 ? ONB_STAT_BLUE_WINS : ONB_STAT_PASSES;
        atomicAdd(&stats[slot], s_acc[threadIdx.x]);
    }
}

// ------------------------------------------------------------------------------------------ launchers
// A slice [first, first + count) of the context's games on a stream of the caller's choice (onb_actor_*: sub-batches of one context
// stepped on their own streams); first must be a multiple of 64 so that the plane tiles stay 16-byte aligned and warps map to whole
// words of the done bits. The whole context on its own stream is the slice {0, n, c->stream, nullptr}.
struct StepSlice {
    int64_t first, count;
    cudaStream_t stream;
    uint2* done_bits;  // [count / 32] or nullptr
};
template <int MODE, bool PLANES, int GAMES, int THREADS>
static cudaError_t launch_step_shape(Ctx* c, const StepSlice& sl, uint32_t step, int auto_reset, int32_t fixed_cards, uint32_t out_flags,
                                     int choose_only) {
    // one tile per CTA (hardware block scheduling balances the load); grid-strided only for absurdly large n
    const int64_t tiles = (sl.count + GAMES - 1) / GAMES;
    const int64_t cap = (int64_t)1 << 30;
    const int grid = (int)(tiles < cap ? tiles : cap);
    if (grid == 0) return cudaSuccess;
    k_env_step<MODE, PLANES, GAMES, THREADS><<<grid, THREADS, 0, sl.stream>>>(
        c->d_states + sl.first, sl.count, c->d_actions + sl.first, c->d_masks + 2 * sl.first, c->d_planes ? c->d_planes + 525 * sl.first : nullptr,
        c->d_stats, c->cfg.seed, c->cfg.game_id_base + (uint64_t)sl.first, step, auto_reset, fixed_cards, out_flags, choose_only, sl.done_bits);
    return cudaGetLastError();
}
template <int MODE>
static cudaError_t launch_step_mode(Ctx* c, const StepSlice& sl, uint32_t step, int auto_reset, int32_t fixed_cards, uint32_t out_flags,
                                    int choose_only) {
    // rules-only variants: 128 games on 128 threads. With planes: 64 games stepped by 2 warps, then all threads of the CTA drain
    // the 134 KB tile (measured on B200, 1 Mi games: 128g/128t 347 us, 32g/256t 318 us, 64g/256t 321 us, 64g/512t 313 us,
    // 128g/512t 319 us, 64g/640t 311 us; a pure fill of the same buffer takes 295 us).
    if (!(out_flags & ONB_OUT_PLANES)) return launch_step_shape<MODE, false, 128, 128>(c, sl, step, auto_reset, fixed_cards, out_flags, choose_only);
    return launch_step_shape<MODE, true, 64, ONB_ENV_THREADS>(c, sl, step, auto_reset, fixed_cards, out_flags, choose_only);
}

cudaError_t launch_env_reset(Ctx* c, const uint8_t* d_decks5, int64_t n_decks, uint32_t epoch) {
    k_env_reset<<<(unsigned)((c->n + 255) / 256), 256, 0, c->stream>>>(c->d_states, c->n, d_decks5, n_decks, c->cfg.seed, c->cfg.game_id_base, epoch);
    return cudaGetLastError();
}
cudaError_t launch_env_reset_where(Ctx* c, const uint8_t* d_mask, uint32_t epoch, unsigned long long* d_count) {
    k_env_reset_where<<<(unsigned)((c->n + 255) / 256), 256, 0, c->stream>>>(c->d_states, c->n, d_mask, c->fixed_cards, c->cfg.seed, c->cfg.game_id_base,
                                                                             epoch, d_count);
    return cudaGetLastError();
}
cudaError_t launch_states_export(Ctx* c, onb_state* d_out, int64_t first, int64_t n) {
    k_states_export<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->d_states, reinterpret_cast<uint32_t*>(d_out), first, n);
    return cudaGetLastError();
}
cudaError_t launch_states_import(Ctx* c, const onb_state* d_in, int64_t first, int64_t n) {
    k_states_import<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->d_states, reinterpret_cast<const uint32_t*>(d_in), first, n);
    return cudaGetLastError();
}
cudaError_t launch_legal_moves(Ctx* c) {
    k_legal_moves<<<(unsigned)((c->n + kTile - 1) / kTile), kTile, 0, c->stream>>>(c->d_states, c->n, c->d_moves, c->d_counts);
    return cudaGetLastError();
}

static cudaError_t launch_env_step_on(Ctx* c, const StepSlice& sl, int mode, uint32_t step, int auto_reset, uint32_t out_flags) {
    const int32_t fixed = c->fixed_cards;
    switch (mode) {
        case 0: return launch_step_mode<0>(c, sl, step, auto_reset, fixed, out_flags, 0);
        case 1: return launch_step_mode<1>(c, sl, step, auto_reset, fixed, out_flags, 0);
        case 2: return launch_step_mode<2>(c, sl, step, auto_reset, fixed, out_flags, 0);
        case 4: return launch_step_mode<0>(c, sl, step, 0, fixed, 0, 1);
        case 5: return launch_step_mode<1>(c, sl, step, 0, fixed, 0, 1);
        default: return launch_step_mode<3>(c, sl, step, 0, fixed, out_flags, 0);
    }
}
cudaError_t launch_env_step(Ctx* c, int mode, uint32_t step, int auto_reset, uint32_t out_flags) {
    return launch_env_step_on(c, StepSlice{0, c->n, c->stream, nullptr}, mode, step, auto_reset, out_flags);
}
cudaError_t launch_env_step_slice(Ctx* c, int mode, uint32_t step, int auto_reset, uint32_t out_flags, int64_t first, int64_t count,
                                  cudaStream_t stream, uint32_t* done_bits) {
    return launch_env_step_on(c, StepSlice{first, count, stream, reinterpret_cast<uint2*>(done_bits)}, mode, step, auto_reset, out_flags);
}
cudaError_t launch_observe(Ctx* c, uint32_t out_flags) { return launch_env_step(c, 3, 0, 0, out_flags); }

cudaError_t launch_env_playout(Ctx* c, uint32_t* d_plies, unsigned long long* d_trace, uint32_t step0, uint32_t max_plies, int mode) {
    k_env_playout<<<(unsigned)((c->n + kTile - 1) / kTile), kTile, 0, c->stream>>>(c->d_states, c->n, d_plies, d_trace, c->d_stats, c->cfg.seed,
                                                                                  c->cfg.game_id_base, step0, max_plies, mode);
    return cudaGetLastError();
}

}  // namespace onb
