// Tall-skinny dense kernels for the Lanczos basis. See dense.h. All are HBM-bound streams over V except TallGemm.
#include "dense.h"

namespace me {
namespace {
constexpr int kThreads = 256;
constexpr uint32_t kColsPerCta = 8;
constexpr uint32_t kSplits = 64;

__global__ void __launch_bounds__(kThreads) GemvTPartialKernel(const double *__restrict__ V, size_t n, uint32_t cols, const double *__restrict__ x, double *__restrict__ partial) {
    __shared__ double red[kThreads / 32][kColsPerCta];
    const uint32_t j0 = blockIdx.x * kColsPerCta, split = blockIdx.y;
    const size_t chunk = (n + kSplits - 1) / kSplits, begin = split * chunk, end = min(n, begin + chunk);
    double acc[kColsPerCta]{};
    for (size_t i = begin + threadIdx.x; i < end; i += kThreads) {
        const double xv = x[i];
#pragma unroll
        for (uint32_t c = 0; c < kColsPerCta; ++c)
            if (j0 + c < cols) acc[c] += V[i + size_t(j0 + c) * n] * xv;
    }
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (uint32_t c = 0; c < kColsPerCta; ++c) {
        double v = acc[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp][c] = v;
    }
    __syncthreads();
    if (threadIdx.x < kColsPerCta && j0 + threadIdx.x < cols) {
        double v = 0;
        for (int w = 0; w < kThreads / 32; ++w) v += red[w][threadIdx.x];
        partial[size_t(j0 + threadIdx.x) * kSplits + split] = v;
    }
}
__global__ void GemvTFinalKernel(const double *__restrict__ partial, uint32_t cols, double *__restrict__ out) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= cols) return;
    double v = 0;
    for (uint32_t s = 0; s < kSplits; ++s) v += partial[size_t(j) * kSplits + s];
    out[j] = v;
}

__global__ void __launch_bounds__(kThreads) GemvNSubKernel(const double *__restrict__ V, size_t n, uint32_t cols, const double *__restrict__ c, double *__restrict__ y) {
    extern __shared__ double cs[];
    for (uint32_t j = threadIdx.x; j < cols; j += kThreads) cs[j] = c[j];
    __syncthreads();
    const size_t i = size_t(blockIdx.x) * kThreads + threadIdx.x;
    if (i >= n) return;
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    uint32_t j = 0;
    for (; j + 4 <= cols; j += 4) {
        s0 += V[i + size_t(j) * n] * cs[j];
        s1 += V[i + size_t(j + 1) * n] * cs[j + 1];
        s2 += V[i + size_t(j + 2) * n] * cs[j + 2];
        s3 += V[i + size_t(j + 3) * n] * cs[j + 3];
    }
    for (; j < cols; ++j) s0 += V[i + size_t(j) * n] * cs[j];
    y[i] -= (s0 + s1) + (s2 + s3);
}

__global__ void AxpbyKernel(size_t n, double a, const double *__restrict__ x, double b, const double *__restrict__ y, double *__restrict__ out) {
    const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a * x[i] + (b != 0.0 ? b * y[i] : 0.0);
}

__device__ __forceinline__ void Dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
constexpr int kChunk = 16, kLd = 68;
__global__ void __launch_bounds__(128) TallGemmKernel(const double *__restrict__ V, size_t n, uint32_t m, const double *__restrict__ Q, uint32_t ldq, uint32_t cols_out, double *__restrict__ C) {
    __shared__ double As[kChunk * kLd], Bs[kChunk * kLd];
    const size_t row0 = size_t(blockIdx.x) * 64;
    const uint32_t col0 = blockIdx.y * 64;
    const uint32_t na = uint32_t(min(size_t(64), n - row0)), nb = min(64u, cols_out - col0);
    const uint32_t t = threadIdx.x, lane = t & 31, w = t >> 5, wm = w & 1, wn = w >> 1;
    double acc[4][4][2]{};
    for (uint32_t kc = 0; kc < m; kc += kChunk) {
        __syncthreads();
        for (uint32_t idx = t; idx < kChunk * 64; idx += 128) {
            const uint32_t r = idx & 63, c = idx >> 6;
            As[c * kLd + r] = (kc + c < m && r < na) ? V[row0 + r + size_t(kc + c) * n] : 0.0;
            const uint32_t kk = idx & 15, j = idx >> 4;
            Bs[kk * kLd + j] = (kc + kk < m && j < nb) ? Q[(kc + kk) + size_t(col0 + j) * ldq] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int ks = 0; ks < kChunk / 4; ++ks) {
            double a[4], b[4];
            const uint32_t kk = 4 * ks + (lane & 3);
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) a[mi] = As[kk * kLd + 32 * wm + 8 * mi + (lane >> 2)];
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) b[ni] = Bs[kk * kLd + 32 * wn + 8 * ni + (lane >> 2)];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) Dmma(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
        }
    }
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const uint32_t r = 32 * wm + 8 * mi + (lane >> 2), c = 32 * wn + 8 * ni + 2 * (lane & 3) + e;
                if (r < na && c < nb) C[row0 + r + size_t(col0 + c) * n] = acc[mi][ni][e];
            }
}
} // namespace

void GemvT(DenseWorkspace &ws, const double *V, size_t n, uint32_t cols, const double *x, double *out, cudaStream_t s) {
    if (cols == 0) return;
    ws.Partial.Reserve(size_t(cols + kColsPerCta) * kSplits);
    GemvTPartialKernel<<<dim3((cols + kColsPerCta - 1) / kColsPerCta, kSplits), kThreads, 0, s>>>(V, n, cols, x, ws.Partial.Ptr);
    GemvTFinalKernel<<<(cols + 127) / 128, 128, 0, s>>>(ws.Partial.Ptr, cols, out);
    ws.Launches += 2;
}
void GemvNSub(DenseWorkspace &ws, const double *V, size_t n, uint32_t cols, const double *c, double *y, cudaStream_t s) {
    if (cols == 0) return;
    GemvNSubKernel<<<uint32_t((n + kThreads - 1) / kThreads), kThreads, cols * sizeof(double), s>>>(V, n, cols, c, y);
    ws.Launches += 1;
}
void Axpby(DenseWorkspace &ws, size_t n, double a, const double *x, double b, const double *y, double *out, cudaStream_t s) {
    AxpbyKernel<<<uint32_t((n + kThreads - 1) / kThreads), kThreads, 0, s>>>(n, a, x, b, y, out);
    ws.Launches += 1;
}
void TallGemm(DenseWorkspace &ws, const double *V, size_t n, uint32_t m, const double *Q, uint32_t ldq, uint32_t cols_out, double *C, cudaStream_t s) {
    if (cols_out == 0) return;
    TallGemmKernel<<<dim3(uint32_t((n + 63) / 64), (cols_out + 63) / 64), 128, 0, s>>>(V, n, m, Q, ldq, cols_out, C);
    ws.Launches += 1;
}

} // namespace me
