// onb_comm.cu -- the one collective of the path behind the C ABI: the end-of-iteration gather of replay samples
// (SelfPlayData = planes 2 100 B + pi 200 B + z 4 B, train.rs:27-33) from every GPU's self-play to the trainer GPU over
// NVLink / NVSwitch -- what the reference does with `join` + `extend` over its worker threads (train.rs:241-245).
//
// NCCL is bound at RUN time (dlopen of the libnccl.so.2 already in the process -- torch's -- or the system one): libonb.so has
// no link-time dependency on it and everything else works without NCCL. Games shard with no data-path collective; this is the
// only place two GPUs talk.
#include <dlfcn.h>
#include <nccl.h>  // types and prototypes only; no symbol is linked

#include <cstring>
#include <new>
#include <vector>

#include "onb_internal.h"

namespace onb {
namespace {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    char err[256] = "";
};
NcclApi g_nccl;

const NcclApi* nccl_api() {
    NcclApi& a = g_nccl;
    if (a.lib) return &a;
    // the instance that is already mapped (a torch process brings its own libnccl.so.2) must be the one we talk to
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        snprintf(a.err, sizeof(a.err), "NCCL is not available: %s", dlerror());
        return nullptr;
    }
#define ONB_SYM(field, name)                                                        \
    a.field = reinterpret_cast<decltype(a.field)>(dlsym(h, name));                  \
    if (!a.field) {                                                                 \
        snprintf(a.err, sizeof(a.err), "NCCL symbol %s is missing", name);          \
        return nullptr;                                                             \
    }
    ONB_SYM(GetUniqueId, "ncclGetUniqueId")
    ONB_SYM(CommInitRank, "ncclCommInitRank")
    ONB_SYM(CommDestroy, "ncclCommDestroy")
    ONB_SYM(AllGather, "ncclAllGather")
    ONB_SYM(Send, "ncclSend")
    ONB_SYM(Recv, "ncclRecv")
    ONB_SYM(GroupStart, "ncclGroupStart")
    ONB_SYM(GroupEnd, "ncclGroupEnd")
    ONB_SYM(GetErrorString, "ncclGetErrorString")
#undef ONB_SYM
    a.lib = h;
    return &a;
}

struct Comm {
    ncclComm_t comm;
    int n_ranks, rank;
    bool owned;
    long long* d_counts;  // [n_ranks][2] device scratch for the (count, capacity) exchange
    int device;
};

int32_t comm_fail(Ctx* c, int32_t code, const char* fmt, const char* a = "", const char* b = "") {
    if (c) snprintf(c->err, sizeof(c->err), fmt, a, b);
    return code;
}

// valid samples of a self-play result -> contiguous rows
__global__ void __launch_bounds__(256) k_pack_rows(const float* __restrict__ src, const int64_t* __restrict__ idx, float* __restrict__ dst, int64_t m,
                                                   int width) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= m * width) return;
    const int64_t r = e / width;
    dst[e] = src[idx[r] * width + (e - r * width)];
}

cudaError_t grow_slot(Ctx* c, int slot, size_t bytes, void** out) {
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
constexpr int kPackPlanes = 28, kPackPi = 29, kPackZ = 30;  // Ctx::sp_buf slots of the packed samples

}  // namespace
}  // namespace onb

using namespace onb;

extern "C" {

int32_t onb_comm_unique_id(uint8_t id_out[128]) {
    if (!id_out) return ONB_E_INVALID;
    const NcclApi* api = nccl_api();
    if (!api) return ONB_E_STATE;
    ncclUniqueId id;
    if (api->GetUniqueId(&id) != ncclSuccess) return ONB_E_CUDA;
    static_assert(sizeof(id) == 128, "ncclUniqueId");
    memcpy(id_out, &id, 128);
    return ONB_OK;
}

int32_t onb_comm_create(onb_ctx* ctx, int32_t n_ranks, int32_t rank, const uint8_t id[128], void* existing_nccl_comm, onb_comm** out) {
    if (!ctx || !out) return ONB_E_INVALID;
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    *out = nullptr;
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks || (!id && !existing_nccl_comm)) return comm_fail(c, ONB_E_INVALID, "onb_comm_create: bad arguments");
    const NcclApi* api = nccl_api();
    if (!api) return comm_fail(c, ONB_E_STATE, "onb_comm_create: %s", g_nccl.err);
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(c->cfg.device);
    Comm* k = new (std::nothrow) Comm();
    if (!k) return ONB_E_NOMEM;
    k->n_ranks = n_ranks;
    k->rank = rank;
    k->device = c->cfg.device;
    k->d_counts = nullptr;
    int32_t rc = ONB_OK;
    if (existing_nccl_comm) {  // adopt the caller's ncclComm_t (e.g. the trainer's); it stays the caller's to destroy
        k->comm = reinterpret_cast<ncclComm_t>(existing_nccl_comm);
        k->owned = false;
    } else {
        ncclUniqueId uid;
        memcpy(&uid, id, 128);
        const ncclResult_t r = api->CommInitRank(&k->comm, n_ranks, uid, rank);
        k->owned = true;
        if (r != ncclSuccess) rc = comm_fail(c, ONB_E_CUDA, "onb_comm_create: ncclCommInitRank: %s", api->GetErrorString(r));
    }
    if (rc == ONB_OK && cudaMalloc(reinterpret_cast<void**>(&k->d_counts), (size_t)n_ranks * 16) != cudaSuccess)
        rc = comm_fail(c, ONB_E_NOMEM, "onb_comm_create: device allocation failed");
    if (prev >= 0 && prev != c->cfg.device) cudaSetDevice(prev);
    if (rc != ONB_OK) {
        delete k;
        return rc;
    }
    *out = reinterpret_cast<onb_comm*>(k);
    return ONB_OK;
}

int32_t onb_comm_destroy(onb_comm* comm) {
    if (!comm) return ONB_E_INVALID;
    Comm* k = reinterpret_cast<Comm*>(comm);
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(k->device);
    if (k->d_counts) cudaFree(k->d_counts);
    const NcclApi* api = nccl_api();
    if (api && k->owned) api->CommDestroy(k->comm);
    if (prev >= 0 && prev != k->device) cudaSetDevice(prev);
    delete k;
    return ONB_OK;
}

// the valid samples of an onb_self_play result as three contiguous device arrays owned by the context
int32_t onb_selfplay_pack(onb_ctx* ctx, const onb_selfplay_result* res, float** planes, float** pi, float** z, int64_t* m_out) {
    if (!ctx) return ONB_E_INVALID;
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!res || !planes || !pi || !z || !m_out) return comm_fail(c, ONB_E_INVALID, "onb_selfplay_pack: null argument");
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(c->cfg.device);
    const int64_t m = res->n_valid;
    float *dp = nullptr, *dpi = nullptr, *dz = nullptr;
    cudaError_t e = grow_slot(c, kPackPlanes, (size_t)m * 2100, (void**)&dp);
    if (e == cudaSuccess) e = grow_slot(c, kPackPi, (size_t)m * 200, (void**)&dpi);
    if (e == cudaSuccess) e = grow_slot(c, kPackZ, (size_t)m * 4, (void**)&dz);
    if (e == cudaSuccess && m > 0) {
        k_pack_rows<<<(unsigned)((m * 525 + 255) / 256), 256, 0, c->stream>>>(res->planes, res->valid_idx, dp, m, 525);
        k_pack_rows<<<(unsigned)((m * 50 + 255) / 256), 256, 0, c->stream>>>(res->pi, res->valid_idx, dpi, m, 50);
        k_pack_rows<<<(unsigned)((m + 255) / 256), 256, 0, c->stream>>>(res->z, res->valid_idx, dz, m, 1);
        e = cudaGetLastError();
    }
    if (prev >= 0 && prev != c->cfg.device) cudaSetDevice(prev);
    if (e != cudaSuccess) return comm_fail(c, e == cudaErrorMemoryAllocation ? ONB_E_NOMEM : ONB_E_CUDA, "onb_selfplay_pack: %s", cudaGetErrorString(e));
    *planes = dp; *pi = dpi; *z = dz; *m_out = m;
    return ONB_OK;
}

// every rank learns every rank's (count, capacity): 16 bytes per rank, one all-gather + host read. Collective.
static int32_t exchange_counts(Ctx* c, Comm* k, const NcclApi* api, long long m_local, long long cap, std::vector<long long>& pairs) {
    pairs.assign((size_t)k->n_ranks * 2, 0);
    const long long mine[2] = {m_local, cap};
    cudaError_t e = cudaMemcpyAsync(k->d_counts + 2 * k->rank, mine, 16, cudaMemcpyHostToDevice, c->stream);
    if (e != cudaSuccess) return comm_fail(c, ONB_E_CUDA, "onb_gather: %s", cudaGetErrorString(e));
    const ncclResult_t r = api->AllGather(k->d_counts + 2 * k->rank, k->d_counts, 2, ncclInt64, k->comm, c->stream);
    if (r != ncclSuccess) return comm_fail(c, ONB_E_CUDA, "onb_gather: ncclAllGather: %s", api->GetErrorString(r));
    e = cudaMemcpyAsync(pairs.data(), k->d_counts, (size_t)k->n_ranks * 16, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) return comm_fail(c, ONB_E_CUDA, "onb_gather: %s", cudaGetErrorString(e));
    return ONB_OK;
}

int32_t onb_gather_counts(onb_ctx* ctx, onb_comm* comm, int64_t m_local, int64_t* counts_host, int64_t* total) {
    if (!ctx || !comm) return ONB_E_INVALID;
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    Comm* k = reinterpret_cast<Comm*>(comm);
    const NcclApi* api = nccl_api();
    if (!api) return comm_fail(c, ONB_E_STATE, "onb_gather_counts: %s", g_nccl.err);
    if (m_local < 0) return comm_fail(c, ONB_E_INVALID, "onb_gather_counts: negative count");
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(c->cfg.device);
    std::vector<long long> pairs;
    const int32_t rc = exchange_counts(c, k, api, (long long)m_local, 0, pairs);
    if (prev >= 0 && prev != c->cfg.device) cudaSetDevice(prev);
    if (rc != ONB_OK) return rc;
    long long sum = 0;
    for (int r = 0; r < k->n_ranks; ++r) {
        if (counts_host) counts_host[r] = pairs[2 * r];
        sum += pairs[2 * r];
    }
    if (total) *total = sum;
    return ONB_OK;
}

int32_t onb_gather_samples(onb_ctx* ctx, onb_comm* comm, int32_t dst_rank, const float* planes, const float* pi, const float* z, int64_t m_local,
                           float* out_planes, float* out_pi, float* out_z, int64_t out_cap, int64_t* counts_host, int64_t* total) {
    if (!ctx || !comm) return ONB_E_INVALID;
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    Comm* k = reinterpret_cast<Comm*>(comm);
    const NcclApi* api = nccl_api();
    if (!api) return comm_fail(c, ONB_E_STATE, "onb_gather_samples: %s", g_nccl.err);
    if (dst_rank < 0 || dst_rank >= k->n_ranks || m_local < 0 || (m_local > 0 && (!planes || !pi || !z)))
        return comm_fail(c, ONB_E_INVALID, "onb_gather_samples: bad arguments");
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(c->cfg.device);
    struct Restore {
        int prev, mine;
        ~Restore() { if (prev >= 0 && prev != mine) cudaSetDevice(prev); }
    } restore{prev, c->cfg.device};
#define NC(call)                                                                                           \
    do {                                                                                                   \
        const ncclResult_t r__ = (call);                                                                   \
        if (r__ != ncclSuccess) return comm_fail(c, ONB_E_CUDA, "onb_gather_samples: %s: %s", #call, api->GetErrorString(r__)); \
    } while (0)
#define CU(call)                                                                                           \
    do {                                                                                                   \
        const cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess) return comm_fail(c, ONB_E_CUDA, "onb_gather_samples: %s: %s", #call, cudaGetErrorString(e__)); \
    } while (0)
    // 1. every rank learns every count AND the destination's capacity, so that "too small" is decided identically everywhere
    //    BEFORE anything is sent (a sender must never be left waiting for a receive that is not coming)
    const bool usable = k->rank != dst_rank || (out_planes && out_pi && out_z);
    std::vector<long long> pairs;
    const int32_t rc = exchange_counts(c, k, api, (long long)m_local, usable ? (long long)out_cap : -1, pairs);
    if (rc != ONB_OK) return rc;
    long long sum = 0;
    for (int r = 0; r < k->n_ranks; ++r) {
        if (counts_host) counts_host[r] = pairs[2 * r];
        sum += pairs[2 * r];
    }
    if (total) *total = sum;
    if (sum > pairs[2 * dst_rank + 1]) return comm_fail(c, ONB_E_OVERFLOW, "onb_gather_samples: the destination buffers are too small (nothing was sent)");
    // 2. rank order concatenation on the destination: one grouped set of sends / receives, the local part is a device copy
    if (k->rank == dst_rank) {
        NC(api->GroupStart());
        long long off = 0;
        for (int r = 0; r < k->n_ranks; ++r) {
            const long long m = pairs[2 * r];
            if (m > 0 && r != dst_rank) {
                NC(api->Recv(out_planes + off * 525, (size_t)m * 525, ncclFloat, r, k->comm, c->stream));
                NC(api->Recv(out_pi + off * 50, (size_t)m * 50, ncclFloat, r, k->comm, c->stream));
                NC(api->Recv(out_z + off, (size_t)m, ncclFloat, r, k->comm, c->stream));
            }
            off += m;
        }
        NC(api->GroupEnd());
        off = 0;
        for (int r = 0; r < dst_rank; ++r) off += pairs[2 * r];
        if (m_local > 0) {
            CU(cudaMemcpyAsync(out_planes + off * 525, planes, (size_t)m_local * 2100, cudaMemcpyDeviceToDevice, c->stream));
            CU(cudaMemcpyAsync(out_pi + off * 50, pi, (size_t)m_local * 200, cudaMemcpyDeviceToDevice, c->stream));
            CU(cudaMemcpyAsync(out_z + off, z, (size_t)m_local * 4, cudaMemcpyDeviceToDevice, c->stream));
        }
    } else if (m_local > 0) {
        NC(api->GroupStart());
        NC(api->Send(planes, (size_t)m_local * 525, ncclFloat, dst_rank, k->comm, c->stream));
        NC(api->Send(pi, (size_t)m_local * 50, ncclFloat, dst_rank, k->comm, c->stream));
        NC(api->Send(z, (size_t)m_local, ncclFloat, dst_rank, k->comm, c->stream));
        NC(api->GroupEnd());
    }
#undef NC
#undef CU
    return ONB_OK;
}

}  // extern "C"
