// Probe: how fast can resident CTAs stream a 7 GB factor with the access shapes the panel sweeps could use?
//   A  8-byte loads, 32 per lane, in the m8n8k4 fragment shape from a column-major panel (what WideSweepKernel does): every warp
//      instruction touches 4 columns x 64 bytes;
//   B  16-byte loads, 16 per lane, each warp instruction 512 contiguous bytes (a fragment-major copy of the panels);
//   C  cp.async.bulk of whole 32 KB slabs into a shared-memory ring (mbarrier complete_tx), consumers read shared memory.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o stream_patterns stream_patterns.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int kSlabDoubles = 32 * 128; // 32 KB

__global__ void __launch_bounds__(128) PatternA(const double *__restrict__ base, size_t slabs, uint32_t ld, unsigned long long *ticket, double *sink) {
    __shared__ size_t s_id;
    const uint32_t t = threadIdx.x, lane = t & 31, q = t >> 5, fr = lane >> 2, fk = lane & 3;
    double acc = 0;
    for (;;) {
        __syncthreads();
        if (t == 0) s_id = atomicAdd(ticket, 1ull);
        __syncthreads();
        const size_t id = s_id;
        if (id >= slabs) break;
        // slab id: 32 consecutive rows of a panel with leading dimension ld; panels of 64 slabs
        const double *p = base + (id / 64) * (size_t(ld) * 128) + (id % 64) * 32 + 8 * q + fr;
        double v[32];
#pragma unroll
        for (int ks = 0; ks < 32; ++ks) v[ks] = p[size_t(4 * ks + fk) * ld];
#pragma unroll
        for (int ks = 0; ks < 32; ++ks) acc += v[ks];
    }
    if (acc == 1.2345) sink[0] = acc;
}

__global__ void __launch_bounds__(128) PatternB(const double *__restrict__ base, size_t slabs, unsigned long long *ticket, double *sink) {
    __shared__ size_t s_id;
    const uint32_t t = threadIdx.x, lane = t & 31, q = t >> 5;
    double acc = 0;
    for (;;) {
        __syncthreads();
        if (t == 0) s_id = atomicAdd(ticket, 1ull);
        __syncthreads();
        const size_t id = s_id;
        if (id >= slabs) break;
        const double2 *p = reinterpret_cast<const double2 *>(base + id * kSlabDoubles + q * 1024) + lane;
        double2 v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = p[32 * j];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc += v[j].x + v[j].y;
    }
    if (acc == 1.2345) sink[0] = acc;
}

__device__ __forceinline__ uint32_t SmemAddr(const void *p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void MbarInit(uint64_t *bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(SmemAddr(bar)), "r"(count)); }
__device__ __forceinline__ void MbarExpectTx(uint64_t *bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(SmemAddr(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void MbarArrive(uint64_t *bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(SmemAddr(bar)) : "memory"); }
__device__ __forceinline__ void MbarWait(uint64_t *bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nWAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE;\nbra WAIT;\nDONE:\n}" ::"r"(SmemAddr(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void BulkLoad(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(SmemAddr(dst)), "l"(src), "r"(bytes), "r"(SmemAddr(bar)) : "memory");
}

// 160 threads: warp 4 produces (claims runs of `run` slabs and copies them into the ring), warps 0-3 consume.
template<int Stages>
__global__ void __launch_bounds__(160) PatternC(const double *__restrict__ base, size_t slabs, uint32_t run, unsigned long long *ticket, double *sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    double *ring = reinterpret_cast<double *>(smem);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + size_t(Stages) * kSlabDoubles * 8), *empty = full + Stages;
    __shared__ size_t s_id;
    const uint32_t t = threadIdx.x, lane = t & 31, q = t >> 5;
    if (t == 0)
        for (int s = 0; s < Stages; ++s) MbarInit(full + s, 1), MbarInit(empty + s, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    uint32_t seq = 0; // slabs handled so far by this CTA (same count on both sides)
    double acc = 0;
    for (;;) {
        __syncthreads();
        if (t == 0) s_id = atomicAdd(ticket, 1ull);
        __syncthreads();
        const size_t first = s_id * run;
        if (first >= slabs) break;
        const uint32_t count = uint32_t(min(size_t(run), slabs - first));
        if (q == 4) {
            if (lane == 0)
                for (uint32_t i = 0; i < count; ++i) {
                    const uint32_t n = seq + i, stage = n % Stages, round = n / Stages;
                    if (round > 0) MbarWait(empty + stage, (round - 1) & 1);
                    MbarExpectTx(full + stage, kSlabDoubles * 8);
                    BulkLoad(ring + size_t(stage) * kSlabDoubles, base + (first + i) * kSlabDoubles, kSlabDoubles * 8, full + stage);
                }
        } else {
            for (uint32_t i = 0; i < count; ++i) {
                const uint32_t n = seq + i, stage = n % Stages, round = n / Stages;
                MbarWait(full + stage, round & 1);
                const double2 *p = reinterpret_cast<const double2 *>(ring + size_t(stage) * kSlabDoubles + q * 1024) + lane;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const double2 v = p[32 * j];
                    acc += v.x + v.y;
                }
                __syncwarp();
                if (lane == 0) MbarArrive(empty + stage);
            }
        }
        seq += count;
    }
    if (acc == 1.2345) sink[0] = acc;
}

int main() {
    const size_t slabs = size_t(7) << 15; // 7 GB / 32 KB = 229376 slabs
    const size_t doubles = slabs * kSlabDoubles + (size_t(1) << 24);
    double *buf, *sink;
    unsigned long long *ticket;
    CK(cudaMalloc(&buf, doubles * 8));
    CK(cudaMemset(buf, 0, doubles * 8));
    CK(cudaMalloc(&sink, 8));
    CK(cudaMalloc(&ticket, 8));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    auto time = [&](const char *name, auto &&launch) {
        float best = 1e9f;
        for (int rep = 0; rep < 4; ++rep) {
            CK(cudaMemset(ticket, 0, 8));
            CK(cudaEventRecord(e0));
            launch();
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            CK(cudaGetLastError());
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep) best = ms < best ? ms : best;
        }
        printf("%-64s %.3f ms  %.2f TB/s\n", name, best, slabs * 32768.0 / (best * 1e-3) / 1e12);
    };
    char name[128];
    for (int per_sm : {2, 4, 5, 8}) {
        // panel of 64 slabs x 128 columns: ld = 2048 + 3 rows (odd, like real panels)
        const uint32_t ld = 2051;
        const size_t slabs_a = (doubles / (size_t(ld) * 128)) * 64;
        snprintf(name, sizeof name, "A  8-byte fragment loads, column-major ld %u, %d CTAs/SM", ld, per_sm);
        float best = 1e9f;
        for (int rep = 0; rep < 4; ++rep) {
            CK(cudaMemset(ticket, 0, 8));
            CK(cudaEventRecord(e0));
            PatternA<<<148 * per_sm, 128>>>(buf, slabs_a, ld, ticket, sink);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            CK(cudaGetLastError());
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep) best = ms < best ? ms : best;
        }
        printf("%-64s %.3f ms  %.2f TB/s\n", name, best, slabs_a * 32768.0 / (best * 1e-3) / 1e12);
        snprintf(name, sizeof name, "B  16-byte coalesced loads (fragment-major), %d CTAs/SM", per_sm);
        time(name, [&] { PatternB<<<148 * per_sm, 128>>>(buf, slabs, ticket, sink); });
    }
    for (uint32_t run : {1u, 4u, 8u}) {
        CK(cudaFuncSetAttribute(PatternC<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 32768 + 64));
        snprintf(name, sizeof name, "C  cp.async.bulk 32 KB slabs, 3 stages, 2 CTAs/SM, runs of %u", run);
        time(name, [&] { PatternC<3><<<148 * 2, 160, 3 * 32768 + 64>>>(buf, slabs, run, ticket, sink); });
        CK(cudaFuncSetAttribute(PatternC<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 32768 + 64));
        snprintf(name, sizeof name, "C  cp.async.bulk 32 KB slabs, 2 stages, 3 CTAs/SM, runs of %u", run);
        time(name, [&] { PatternC<2><<<148 * 3, 160, 2 * 32768 + 64>>>(buf, slabs, run, ticket, sink); });
        CK(cudaFuncSetAttribute(PatternC<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * 32768 + 64));
        snprintf(name, sizeof name, "C  cp.async.bulk 32 KB slabs, 6 stages, 1 CTA/SM, runs of %u", run);
        time(name, [&] { PatternC<6><<<148, 160, 6 * 32768 + 64>>>(buf, slabs, run, ticket, sink); });
    }
    return 0;
}
