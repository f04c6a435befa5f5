// onb_replay.cu -- the trainer's side of the replay pipeline on the device: the data buffer of train.rs
// (`data_buffer = Vec::with_capacity(config.buffer_size)`, extended by every iteration's self-play, train.rs:213,241-245 -- the
// reference never trims it; here `capacity` bounds it and the newest samples overwrite the oldest) as a fixed-capacity ring in HBM, and `data_buffer.iter().choose_multiple(&mut rng, train_batch_size)` (train.rs:280-283) as a minibatch gather:
// batch_size DISTINCT samples, uniformly at random. The reference draws from thread_rng (not reproducible); here the choice is a
// keyed pseudo-random permutation of [0, size) -- a 4-round Feistel network over the next power of four, cycle-walked into range --
// so a (seed, size) pair always yields the same minibatch and no index table is ever materialised.
// 2 304 B per sample: planes [21][5][5] f32, pi [2][25] f32, z f32.
#include <new>

#include "onb_internal.h"
#include "onb_rules.cuh"

namespace onb {
namespace {

struct Replay {
    Ctx* c;
    int64_t capacity, size, head;  // head = next write position
    float *planes, *pi, *z;        // ring storage
    float *b_planes, *b_pi, *b_z;  // minibatch output (grow-only)
    int64_t b_cap;
};

// pseudo-random permutation of [0, 2^(2*half_bits)) keyed by `key`
__host__ __device__ __forceinline__ uint64_t feistel(uint64_t x, uint32_t half_bits, uint64_t key) {
    const uint64_t mask = (1ull << half_bits) - 1ull;
    uint64_t l = x >> half_bits, r = x & mask;
#pragma unroll
    for (uint32_t round = 0; round < 4; ++round) {
        const uint64_t f = mix64(r ^ key ^ (0x9E3779B97F4A7C15ull * (round + 1))) & mask;
        const uint64_t t = l ^ f;
        l = r;
        r = t;
    }
    return (l << half_bits) | r;
}
__host__ __device__ __forceinline__ uint64_t permuted_index(uint64_t i, uint64_t size, uint32_t half_bits, uint64_t key) {
    uint64_t x = feistel(i, half_bits, key);
    while (x >= size) x = feistel(x, half_bits, key);  // cycle walking: a permutation of [0, size), < 4 steps on average
    return x;
}

// sample j of the minibatch = ring slot perm(j); one thread per output float (coalesced writes, gathered reads of whole rows)
__global__ void __launch_bounds__(256) k_replay_gather(const float* __restrict__ src, float* __restrict__ dst, int64_t batch, int width, uint64_t size,
                                                       uint32_t half_bits, uint64_t key) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= batch * width) return;
    const int64_t j = e / width;
    const uint64_t slot = permuted_index((uint64_t)j, size, half_bits, key);
    dst[e] = src[(int64_t)slot * width + (e - j * width)];
}

int32_t rfail(Ctx* c, int32_t code, const char* msg, cudaError_t e = cudaSuccess) {
    if (e != cudaSuccess) snprintf(c->err, sizeof(c->err), "%s: %s", msg, cudaGetErrorString(e));
    else snprintf(c->err, sizeof(c->err), "%s", msg);
    return code;
}
struct ReplayDeviceGuard {
    int prev = -1, mine;
    explicit ReplayDeviceGuard(int dev) : mine(dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != mine) cudaSetDevice(mine);
    }
    ~ReplayDeviceGuard() {
        if (prev >= 0 && prev != mine) cudaSetDevice(prev);
    }
};

}  // namespace
}  // namespace onb

using namespace onb;

extern "C" {

int32_t onb_replay_create(onb_ctx* ctx, int64_t capacity, onb_replay** out) {
    if (!ctx || !out) return ONB_E_INVALID;
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    *out = nullptr;
    if (capacity <= 0) return rfail(c, ONB_E_INVALID, "onb_replay_create: capacity must be positive");
    ReplayDeviceGuard guard(c->cfg.device);
    Replay* r = new (std::nothrow) Replay();
    if (!r) return ONB_E_NOMEM;
    *r = Replay{c, capacity, 0, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0};
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&r->planes), (size_t)capacity * 2100);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&r->pi), (size_t)capacity * 200);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&r->z), (size_t)capacity * 4);
    if (e != cudaSuccess) {
        onb_replay_destroy(reinterpret_cast<onb_replay*>(r));
        return rfail(c, ONB_E_NOMEM, "onb_replay_create", e);
    }
    *out = reinterpret_cast<onb_replay*>(r);
    return ONB_OK;
}

int32_t onb_replay_destroy(onb_replay* rb) {
    if (!rb) return ONB_E_INVALID;
    Replay* r = reinterpret_cast<Replay*>(rb);
    ReplayDeviceGuard guard(r->c->cfg.device);
    cudaStreamSynchronize(r->c->stream);
    for (float* p : {r->planes, r->pi, r->z, r->b_planes, r->b_pi, r->b_z})
        if (p) cudaFree(p);
    delete r;
    return ONB_OK;
}

// append m samples held in DEVICE memory (onb_selfplay_pack / onb_gather_samples output); the newest overwrite the oldest
int32_t onb_replay_add(onb_replay* rb, const float* planes_dev, const float* pi_dev, const float* z_dev, int64_t m) {
    if (!rb) return ONB_E_INVALID;
    Replay* r = reinterpret_cast<Replay*>(rb);
    Ctx* c = r->c;
    if (m < 0 || (m > 0 && (!planes_dev || !pi_dev || !z_dev))) return rfail(c, ONB_E_INVALID, "onb_replay_add: bad arguments");
    ReplayDeviceGuard guard(c->cfg.device);
    if (m > r->capacity) {  // only the newest `capacity` samples can survive
        const int64_t skip = m - r->capacity;
        planes_dev += skip * 525; pi_dev += skip * 50; z_dev += skip;
        m = r->capacity;
    }
    int64_t done = 0;
    while (done < m) {  // at most two contiguous pieces (wrap-around)
        const int64_t run = (m - done) < (r->capacity - r->head) ? (m - done) : (r->capacity - r->head);
        cudaError_t e = cudaMemcpyAsync(r->planes + r->head * 525, planes_dev + done * 525, (size_t)run * 2100, cudaMemcpyDeviceToDevice, c->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(r->pi + r->head * 50, pi_dev + done * 50, (size_t)run * 200, cudaMemcpyDeviceToDevice, c->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(r->z + r->head, z_dev + done, (size_t)run * 4, cudaMemcpyDeviceToDevice, c->stream);
        if (e != cudaSuccess) return rfail(c, ONB_E_CUDA, "onb_replay_add", e);
        r->head = (r->head + run) % r->capacity;
        done += run;
    }
    r->size = r->size + m < r->capacity ? r->size + m : r->capacity;
    return ONB_OK;
}

int32_t onb_replay_size(onb_replay* rb, int64_t* size, int64_t* capacity) {
    if (!rb) return ONB_E_INVALID;
    Replay* r = reinterpret_cast<Replay*>(rb);
    if (size) *size = r->size;
    if (capacity) *capacity = r->capacity;
    return ONB_OK;
}

// min(batch, size) DISTINCT samples, uniformly at random (choose_multiple, train.rs:280-283), as three contiguous device arrays owned by
// the ring (valid until the next onb_replay_sample / onb_replay_destroy); asynchronous on the context's stream
int32_t onb_replay_sample(onb_replay* rb, int64_t batch, uint64_t seed, float** planes, float** pi, float** z, int64_t* n_out) {
    if (!rb) return ONB_E_INVALID;
    Replay* r = reinterpret_cast<Replay*>(rb);
    Ctx* c = r->c;
    if (batch < 0 || !planes || !pi || !z || !n_out) return rfail(c, ONB_E_INVALID, "onb_replay_sample: bad arguments");
    ReplayDeviceGuard guard(c->cfg.device);
    const int64_t b = batch < r->size ? batch : r->size;
    if (b > r->b_cap) {
        cudaStreamSynchronize(c->stream);
        for (float** p : {&r->b_planes, &r->b_pi, &r->b_z})
            if (*p) { cudaFree(*p); *p = nullptr; }
        r->b_cap = 0;
        cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&r->b_planes), (size_t)b * 2100);
        if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&r->b_pi), (size_t)b * 200);
        if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&r->b_z), (size_t)b * 4);
        if (e != cudaSuccess) return rfail(c, ONB_E_NOMEM, "onb_replay_sample", e);
        r->b_cap = b;
    }
    if (b > 0) {
        uint32_t half_bits = 1;
        while ((1ull << (2 * half_bits)) < (uint64_t)r->size) ++half_bits;
        const uint64_t key = mix64(seed ^ 0xD6E8FEB86659FD93ull);
        k_replay_gather<<<(unsigned)((b * 525 + 255) / 256), 256, 0, c->stream>>>(r->planes, r->b_planes, b, 525, (uint64_t)r->size, half_bits, key);
        k_replay_gather<<<(unsigned)((b * 50 + 255) / 256), 256, 0, c->stream>>>(r->pi, r->b_pi, b, 50, (uint64_t)r->size, half_bits, key);
        k_replay_gather<<<(unsigned)((b + 255) / 256), 256, 0, c->stream>>>(r->z, r->b_z, b, 1, (uint64_t)r->size, half_bits, key);
        const cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return rfail(c, ONB_E_CUDA, "onb_replay_sample", e);
    }
    *planes = r->b_planes; *pi = r->b_pi; *z = r->b_z; *n_out = b;
    return ONB_OK;
}

// the ring slots a (seed, size) pair selects, on the host: what onb_replay_sample gathers, for callers that index their own side data
int32_t onb_replay_indices(int64_t size, int64_t batch, uint64_t seed, int64_t* out) {
    if (size < 0 || batch < 0 || !out) return ONB_E_INVALID;
    const int64_t b = batch < size ? batch : size;
    uint32_t half_bits = 1;
    while ((1ull << (2 * half_bits)) < (uint64_t)size) ++half_bits;
    const uint64_t key = mix64(seed ^ 0xD6E8FEB86659FD93ull);
    for (int64_t j = 0; j < b; ++j) out[j] = (int64_t)permuted_index((uint64_t)j, (uint64_t)size, half_bits, key);
    return ONB_OK;
}

}  // extern "C"
