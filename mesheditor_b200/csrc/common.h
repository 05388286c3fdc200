// Shared host-side helpers: per-thread error text, CUDA error capture, RAII device buffers.
#pragma once

#include "../../include/me_modal.h"

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace me {

void SetLastError(const char *fmt, ...);
const char *LastError();

// Thrown inside the library, converted to an MeStatus at the C-ABI boundary (no exception crosses it).
struct Failure {
    MeStatus Status;
};

[[noreturn]] inline void Fail(MeStatus status, const char *fmt, ...) {
    char buffer[512];
    va_list args;
    va_start(args, fmt);
    vsnprintf(buffer, sizeof buffer, fmt, args);
    va_end(args);
    SetLastError("%s", buffer);
    throw Failure{status};
}

#define ME_CUDA(expr)                                                                                         \
    do {                                                                                                      \
        const cudaError_t me_err_ = (expr);                                                                   \
        if (me_err_ != cudaSuccess)                                                                           \
            ::me::Fail(me_err_ == cudaErrorMemoryAllocation ? ME_OUT_OF_MEMORY : ME_CUDA_ERROR, "%s failed: %s (%s:%d)", #expr, \
                       cudaGetErrorString(me_err_), __FILE__, __LINE__);                                      \
    } while (0)

// Runs `body`, mapping Failure / std::exception to a status code.
template<typename Body>
MeStatus Guard(Body &&body) {
    try {
        body();
        return ME_OK;
    } catch (const Failure &f) {
        return f.Status;
    } catch (const std::bad_alloc &) {
        SetLastError("host allocation failed");
        return ME_OUT_OF_MEMORY;
    } catch (const std::exception &e) {
        SetLastError("%s", e.what());
        return ME_CUDA_ERROR;
    }
}

// Device allocations come from the device's stream-ordered memory pool with an unbounded release threshold: a solve
// frees ~10 GB of panels and basis vectors when it returns, and handing those pages back to the driver (cudaFree) and
// mapping them again on the next call (cudaMalloc) costs far more than the assembly they bracket. Release() first waits
// for the device, which keeps the old cudaFree semantics (no kernel can still be using the block).
void *PoolAllocate(size_t bytes);
void PoolFree(void *ptr);

// Device buffer that only ever grows; contents are not preserved across a growth unless asked.
template<typename T>
struct DeviceBuffer {
    T *Ptr{nullptr};
    size_t Capacity{0};

    DeviceBuffer() = default;
    DeviceBuffer(const DeviceBuffer &) = delete;
    DeviceBuffer &operator=(const DeviceBuffer &) = delete;
    ~DeviceBuffer() { Release(); }

    void Release() {
        if (Ptr) PoolFree(Ptr);
        Ptr = nullptr;
        Capacity = 0;
    }
    void Reserve(size_t count) {
        if (count <= Capacity) return;
        Release();
        Ptr = static_cast<T *>(PoolAllocate(count * sizeof(T)));
        Capacity = count;
    }
    void Upload(const T *host, size_t count, cudaStream_t stream) {
        Reserve(count);
        if (count) ME_CUDA(cudaMemcpyAsync(Ptr, host, count * sizeof(T), cudaMemcpyHostToDevice, stream));
    }
    void Upload(const std::vector<T> &host, cudaStream_t stream) { Upload(host.data(), host.size(), stream); }
};

// Pinned host staging buffer (growth only).
template<typename T>
struct PinnedBuffer {
    T *Ptr{nullptr};
    size_t Capacity{0};
    PinnedBuffer() = default;
    PinnedBuffer(const PinnedBuffer &) = delete;
    PinnedBuffer &operator=(const PinnedBuffer &) = delete;
    ~PinnedBuffer() {
        if (Ptr) cudaFreeHost(Ptr);
    }
    void Reserve(size_t count) {
        if (count <= Capacity) return;
        if (Ptr) cudaFreeHost(Ptr);
        Ptr = nullptr;
        Capacity = 0;
        ME_CUDA(cudaMallocHost(reinterpret_cast<void **>(&Ptr), count * sizeof(T)));
        Capacity = count;
    }
};

} // namespace me
