// Write-bandwidth microbenchmarks for the plane writer (exploration; not part of the product).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

// A: elementwise-style: block writes blockDim*UNROLL float4 contiguous, one block per chunk
template <int UNROLL, bool CS>
__global__ void kA(float4* __restrict__ dst, size_t nvec) {
    size_t base = (size_t)blockIdx.x * blockDim.x * UNROLL + threadIdx.x;
    const float4 v = make_float4(1.f, 0.f, 1.f, 0.f);
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
        size_t i = base + (size_t)u * blockDim.x;
        if (i < nvec) { if (CS) __stcs(dst + i, v); else dst[i] = v; }
    }
}
// B: CTA streams a contiguous region of `per_cta` float4 in a loop, constant value
__global__ void kB(float4* __restrict__ dst, size_t nvec, uint32_t per_cta) {
    float4* d = dst + (size_t)blockIdx.x * per_cta;
    const float4 v = make_float4(1.f, 0.f, 1.f, 0.f);
    for (uint32_t q = threadIdx.x; q < per_cta; q += blockDim.x) __stcs(d + q, v);
}
// C: like B but values expanded from shared-memory plane words (the real writer loop)
__global__ void kC(float4* __restrict__ dst, size_t nvec, uint32_t per_cta) {
    extern __shared__ uint32_t s_pl[];
    const uint32_t nwords = per_cta * 4 / 25 + 2;
    for (uint32_t i = threadIdx.x; i < nwords; i += blockDim.x) s_pl[i] = (i * 2654435761u + blockIdx.x) & 0x1FFFFFFu;
    __syncthreads();
    float4* d = dst + (size_t)blockIdx.x * per_cta;
    for (uint32_t q = threadIdx.x; q < per_cta; q += blockDim.x) {
        const uint32_t e = q * 4u, G = e / 25u, r = e - G * 25u;
        const uint32_t v = (s_pl[G] >> r) | (s_pl[G + 1] << (25u - r));
        float4 f;
        f.x = __uint_as_float(0x3F800000u & (0u - (v & 1u)));
        f.y = __uint_as_float(0x3F800000u & (0u - ((v >> 1) & 1u)));
        f.z = __uint_as_float(0x3F800000u & (0u - ((v >> 2) & 1u)));
        f.w = __uint_as_float(0x3F800000u & (0u - ((v >> 3) & 1u)));
        __stcs(d + q, f);
    }
}
// D: writer where each WARP owns a contiguous region (per_warp float4), to test warp-contiguous streams
__global__ void kD(float4* __restrict__ dst, size_t nvec, uint32_t per_warp) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    float4* d = dst + (size_t)warp * per_warp;
    const float4 v = make_float4(1.f, 0.f, 1.f, 0.f);
    for (uint32_t q = lane; q < per_warp; q += 32) __stcs(d + q, v);
}

int main() {
    const size_t nfloat = (size_t)(1 << 20) * 525;
    const size_t nvec = nfloat / 4;
    float4* d;
    CK(cudaMalloc(&d, nvec * 16));
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    auto timeit = [&](const char* name, auto launch) {
        for (int i = 0; i < 3; ++i) launch();
        cudaDeviceSynchronize();
        float best = 1e9;
        for (int r = 0; r < 10; ++r) {
            cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
        }
        cudaError_t e = cudaGetLastError();
        printf("%-44s %8.1f us %8.0f GB/s %s\n", name, best * 1e3, nvec * 16 / best / 1e6, e == cudaSuccess ? "" : cudaGetErrorString(e));
    };
    timeit("memset", [&] { cudaMemsetAsync(d, 0, nvec * 16); });
    timeit("A unroll4 plain 128thr", [&] { kA<4, false><<<(unsigned)((nvec + 511) / 512), 128>>>(d, nvec); });
    timeit("A unroll4 cs 128thr", [&] { kA<4, true><<<(unsigned)((nvec + 511) / 512), 128>>>(d, nvec); });
    timeit("A unroll8 cs 256thr", [&] { kA<8, true><<<(unsigned)((nvec + 2047) / 2048), 256>>>(d, nvec); });
    timeit("A unroll1 cs 256thr", [&] { kA<1, true><<<(unsigned)((nvec + 255) / 256), 256>>>(d, nvec); });
    for (uint32_t games : {8u, 32u, 128u, 512u}) {
        uint32_t per = games * 525 / 4;
        char nm[64];
        snprintf(nm, 64, "B const loop, %u games/CTA 128thr", games);
        timeit(nm, [&] { kB<<<(unsigned)(nvec / per), 128>>>(d, nvec, per); });
        snprintf(nm, 64, "C expand loop, %u games/CTA 128thr", games);
        timeit(nm, [&] { kC<<<(unsigned)(nvec / per), 128, (per * 4 / 25 + 2) * 4>>>(d, nvec, per); });
    }
    timeit("B const loop, 128 games/CTA 256thr", [&] { kB<<<(unsigned)(nvec / 16800), 256>>>(d, nvec, 16800); });
    timeit("B const loop, 128 games/CTA 512thr", [&] { kB<<<(unsigned)(nvec / 16800), 512>>>(d, nvec, 16800); });
    timeit("C expand loop, 128 games/CTA 256thr", [&] { kC<<<(unsigned)(nvec / 16800), 256, (16800 * 4 / 25 + 2) * 4>>>(d, nvec, 16800); });
    timeit("D warp-contig 4 games/warp 128thr", [&] { kD<<<(unsigned)(nvec / 525 / 4), 128>>>(d, nvec, 525); });
    timeit("D warp-contig 32 games/warp 128thr", [&] { kD<<<(unsigned)(nvec / 4200 / 4), 128>>>(d, nvec, 4200); });
    return 0;
}
