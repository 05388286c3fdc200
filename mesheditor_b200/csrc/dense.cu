// Tall-skinny dense kernels for the Lanczos basis. See dense.h. All are HBM-bound streams over V except TallGemm.
#include "dense.h"

#include <algorithm>

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
__global__ void __launch_bounds__(128) TallGemmKernel(const double *__restrict__ V, size_t n, uint32_t m, const double *__restrict__ Q, uint32_t ldq, uint32_t cols_out, double *__restrict__ C, double alpha, double beta) {
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
                if (r < na && c < nb) {
                    double *dst = C + row0 + r + size_t(col0 + c) * n;
                    *dst = beta != 0.0 ? alpha * acc[mi][ni][e] + beta * *dst : alpha * acc[mi][ni][e];
                }
            }
}

// Narrow forms (at most 8 columns on the short side): the long operand streams from HBM exactly once, straight into
// m8n8k4 A-fragments (32 loads in flight per thread, no shared-memory staging, no barriers in the main loop).
constexpr int kNarrow = 8;

// C[n x 8] = alpha * V[n x m] * Q[m x cols (<= 8)] + beta * C. One warp per 32 rows.
__global__ void __launch_bounds__(128) TallGemmNarrowKernel(const double *__restrict__ V, size_t n, uint32_t m, const double *__restrict__ Q, uint32_t ldq, uint32_t cols, double *__restrict__ C, double alpha, double beta) {
    extern __shared__ double qs[]; // [m rounded up to 32][8]
    const uint32_t t = threadIdx.x, lane = t & 31, w = t >> 5, fr = lane >> 2, fk = lane & 3;
    const uint32_t mpad = (m + 31) & ~31u;
    for (uint32_t idx = t; idx < mpad * kNarrow; idx += 128) {
        const uint32_t k = idx / kNarrow, c = idx % kNarrow;
        qs[idx] = (k < m && c < cols) ? Q[k + size_t(c) * ldq] : 0.0;
    }
    __syncthreads();
    const size_t row0 = size_t(blockIdx.x) * 128 + 32 * w;
    if (row0 >= n) return;
    double acc[4][2]{};
    for (uint32_t kc = 0; kc < m; kc += 32) {
        double val[32];
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) {
                const size_t r = row0 + 8 * mi + fr;
                const uint32_t k = kc + 4 * ks + fk;
                val[ks * 4 + mi] = (r < n && k < m) ? V[r + size_t(k) * n] : 0.0;
            }
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
            const double b = qs[(kc + 4 * ks + fk) * kNarrow + fr];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) Dmma(acc[mi][0], acc[mi][1], val[ks * 4 + mi], b);
        }
    }
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const size_t r = row0 + 8 * mi + fr;
            const uint32_t c = 2 * fk + e;
            if (r < n && c < cols) {
                double *dst = C + r + size_t(c) * n;
                *dst = beta != 0.0 ? alpha * acc[mi][e] + beta * *dst : alpha * acc[mi][e];
            }
        }
}

// partial[split][a_pad x 8] of X[:, :a]^T Y[:, :c (<= 8)]: CTA = 32 columns of X x one row range; its 4 warps take
// alternate 32-row steps and are reduced in shared memory. HBM-bound (X is read once: 8 n a bytes).
//   * The sum over rows does not care in which order the tensor core sees them, so a thread's fragment elements are PAIRS of
//     consecutive rows (rows 2 fk, 2 fk + 1 of each 8-row group feed DMMA steps 2g and 2g + 1): every load is 16 bytes and a
//     quarter-warp covers 64 contiguous bytes of a column, half the load instructions of the one-row-per-step layout.
//   * The grid is ONE wave: (column groups x splits) <= the CTAs resident at once (4 per SM by registers). The first version
//     launched 64..512 splits whatever the width; at 320 columns that was 640 CTAs for 592 slots, i.e. a second, almost
//     empty wave, and the kernel ran at 0.28 of HBM peak (profiles/r02_solve.md).
constexpr uint32_t kGramCtasPerSm = 4;
__global__ void __launch_bounds__(128, kGramCtasPerSm) GramNarrowPartialKernel(const double *__restrict__ X, size_t n, uint32_t a, const double *__restrict__ Y, uint32_t c, double *__restrict__ partial, uint32_t a_pad, uint32_t splits) {
    __shared__ double part[4 * 32 * kNarrow];
    const uint32_t t = threadIdx.x, lane = t & 31, w = t >> 5, fr = lane >> 2, fk = lane & 3;
    const uint32_t col0 = blockIdx.x * 32, split = blockIdx.y;
    size_t chunk = (n + splits - 1) / splits;
    chunk = (chunk + 127) & ~size_t(127); // whole CTA iterations (4 warps x 4 groups of 8 rows)
    const size_t begin = split * chunk, end = min(n, begin + chunk);
    const bool even = (n & 1) == 0; // 16-byte loads need every column to start 16-byte aligned
    double acc[4][2]{};
    for (size_t r0 = begin + 32 * w; r0 < end; r0 += 128) {
        double2 val[16], b[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const size_t r = r0 + 8 * g + 2 * fk;
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) {
                const uint32_t col = col0 + 8 * mi + fr;
                const double *src = X + r + size_t(col) * n;
                double2 v = make_double2(0.0, 0.0);
                if (col < a) {
                    if (even && r + 1 < end) v = *reinterpret_cast<const double2 *>(src);
                    else {
                        if (r < end) v.x = src[0];
                        if (r + 1 < end) v.y = src[1];
                    }
                }
                val[g * 4 + mi] = v;
            }
            double2 y = make_double2(0.0, 0.0);
            if (fr < c) {
                const double *src = Y + r + size_t(fr) * n;
                if (even && r + 1 < end) y = *reinterpret_cast<const double2 *>(src);
                else {
                    if (r < end) y.x = src[0];
                    if (r + 1 < end) y.y = src[1];
                }
            }
            b[g] = y;
        }
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) {
                Dmma(acc[mi][0], acc[mi][1], val[g * 4 + mi].x, b[g].x);
                Dmma(acc[mi][0], acc[mi][1], val[g * 4 + mi].y, b[g].y);
            }
    }
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
        part[(w * 32 + 8 * mi + fr) * kNarrow + 2 * fk] = acc[mi][0];
        part[(w * 32 + 8 * mi + fr) * kNarrow + 2 * fk + 1] = acc[mi][1];
    }
    __syncthreads();
    for (uint32_t idx = t; idx < 32 * kNarrow; idx += 128)
        partial[(size_t(split) * a_pad + col0) * kNarrow + idx] = (part[idx] + part[256 + idx]) + (part[512 + idx] + part[768 + idx]);
}
// One warp per output entry: lanes sum every 32nd split, then a shuffle tree (fixed order, so still deterministic).
__global__ void GramNarrowFinalKernel(const double *__restrict__ partial, uint32_t a, uint32_t c, uint32_t a_pad, double *__restrict__ out, uint32_t ldo, uint32_t splits) {
    const uint32_t idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (idx >= a * kNarrow) return;
    const uint32_t col = idx / kNarrow, w = idx % kNarrow;
    if (w >= c) return;
    double v = 0;
    for (uint32_t s = lane; s < splits; s += 32) v += partial[(size_t(s) * a_pad + col) * kNarrow + w];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) out[col + size_t(w) * ldo] = v;
}

// The QL rotation history applied to one warp's 32 columns of the transposed eigenvector matrix (hosteig.cpp holds the host
// form). The chain is sequential in the rotations (tens of thousands of them), so everything off the chain is taken out of it:
// the records arrive in shared memory by bulk asynchronous copies two batches ahead (the lanes then read a record from one
// address, a broadcast whose address does not depend on the arithmetic), and since a sweep of the QL iteration walks down the
// rows - rotation (i, i+1) is followed by (i-1, i) - row i is carried in a register between the two: one row loaded and one
// stored per rotation.
constexpr uint32_t kRotationBatch = 512; // records per staged batch (12 KB)
__global__ void __launch_bounds__(32) ApplyRotationsKernel(double *__restrict__ zt, uint32_t m, const QlRotation *__restrict__ rotations, size_t count) {
    extern __shared__ __align__(16) double columns[]; // [m][32], then two batches of records
    QlRotation *staged = reinterpret_cast<QlRotation *>(columns + size_t(m) * 32);
    const uint32_t lane = threadIdx.x, col = blockIdx.x * 32 + lane;
    const bool valid = col < m;
    const size_t batches = (count + kRotationBatch - 1) / kRotationBatch;
    // 16-byte pieces of batch `b` into buffer b & 1 (a record is 24 bytes: 3 pieces per 2 records; the list is padded to whole batches)
    const auto stage = [&](size_t b) {
        if (b < batches) {
            const char *src = reinterpret_cast<const char *>(rotations + b * kRotationBatch);
            const uint32_t dst = uint32_t(__cvta_generic_to_shared(staged + (b & 1) * kRotationBatch));
            for (uint32_t piece = lane; piece < kRotationBatch * sizeof(QlRotation) / 16; piece += 32)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + piece * 16), "l"(src + size_t(piece) * 16));
        }
        asm volatile("cp.async.commit_group;");
    };
    stage(0);
    stage(1);
    for (uint32_t r = 0; r < m; ++r) columns[r * 32 + lane] = valid ? zt[size_t(r) * m + col] : 0.0;
    uint32_t held_row = 0xFFFFFFFFu; // the row whose current value sits in `held` instead of shared memory
    double held = 0.0;
    for (size_t b = 0; b < batches; ++b) {
        asm volatile("cp.async.wait_group 1;");
        __syncwarp();
        const QlRotation *batch = staged + (b & 1) * kRotationBatch;
        const uint32_t in_batch = uint32_t(min(size_t(kRotationBatch), count - b * kRotationBatch));
        const auto rotate = [&](const QlRotation &q) {
            const double c = q.C, s = q.S;
            const uint32_t row = q.Row;
            double upper; // row + 1
            if (held_row == row + 1) {
                upper = held;
            } else {
                if (held_row != 0xFFFFFFFFu) columns[held_row * 32 + lane] = held;
                upper = columns[(row + 1) * 32 + lane];
            }
            const double lower = columns[row * 32 + lane];
            columns[(row + 1) * 32 + lane] = s * lower + c * upper;
            held = c * lower - s * upper;
            held_row = row;
        };
        // Four records at a time into registers before any of their stores: the compiler cannot tell that the staged records and the
        // columns never overlap, and would otherwise load every record behind the store of the rotation before it.
        uint32_t i = 0;
        for (; i + 4 <= in_batch; i += 4) {
            const QlRotation q0 = batch[i], q1 = batch[i + 1], q2 = batch[i + 2], q3 = batch[i + 3];
            // A sweep of the QL iteration walks down the rows one at a time. On such a stretch the three lower rows are not touched
            // by the stores of the rotations before them (nor held in the register), so they are read ahead of the whole group, and
            // what is left on the chain of a rotation is its two multiply-adds on the carried row.
            const bool descending = q1.Row + 1 == q0.Row && q2.Row + 1 == q1.Row && q3.Row + 1 == q2.Row && held_row != q1.Row && held_row != q2.Row && held_row != q3.Row;
            if (descending) {
                const double l1 = columns[q1.Row * 32 + lane], l2 = columns[q2.Row * 32 + lane], l3 = columns[q3.Row * 32 + lane];
                rotate(q0); // leaves row q0.Row = q1.Row + 1 in `held`
                columns[(q1.Row + 1) * 32 + lane] = q1.S * l1 + q1.C * held;
                held = q1.C * l1 - q1.S * held;
                columns[(q2.Row + 1) * 32 + lane] = q2.S * l2 + q2.C * held;
                held = q2.C * l2 - q2.S * held;
                columns[(q3.Row + 1) * 32 + lane] = q3.S * l3 + q3.C * held;
                held = q3.C * l3 - q3.S * held;
                held_row = q3.Row;
            } else {
                rotate(q0), rotate(q1), rotate(q2), rotate(q3);
            }
        }
        for (; i < in_batch; ++i) rotate(batch[i]);
        __syncwarp();
        stage(b + 2);
    }
    if (held_row != 0xFFFFFFFFu) columns[held_row * 32 + lane] = held;
    if (valid)
        for (uint32_t r = 0; r < m; ++r) zt[size_t(r) * m + col] = columns[r * 32 + lane];
}

__global__ void GatherRowsKernel(const double *__restrict__ zt, uint32_t m, const uint32_t *__restrict__ rows, uint32_t k, double *__restrict__ out) {
    const uint32_t j = blockIdx.x;
    const double *src = zt + size_t(rows[j]) * m;
    for (uint32_t r = threadIdx.x; r < m; r += blockDim.x) out[r + size_t(j) * m] = src[r];
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
void TallGemm(DenseWorkspace &ws, const double *V, size_t n, uint32_t m, const double *Q, uint32_t ldq, uint32_t cols_out, double *C, cudaStream_t s, double alpha, double beta) {
    if (cols_out == 0) return;
    if (cols_out <= kNarrow && m <= 736) { // Q (m x 8) fits the default 48 KB of shared memory
        const uint32_t mpad = (m + 31) & ~31u;
        TallGemmNarrowKernel<<<uint32_t((n + 127) / 128), 128, size_t(mpad) * kNarrow * sizeof(double), s>>>(V, n, m, Q, ldq, cols_out, C, alpha, beta);
    } else {
        TallGemmKernel<<<dim3(uint32_t((n + 63) / 64), (cols_out + 63) / 64), 128, 0, s>>>(V, n, m, Q, ldq, cols_out, C, alpha, beta);
    }
    ws.Launches += 1;
}
void Gram(DenseWorkspace &ws, const double *X, size_t n, uint32_t a, const double *Y, uint32_t c, double *out, uint32_t ldo, cudaStream_t s) {
    if (a == 0 || c == 0) return;
    const uint32_t a_pad = (a + 31) & ~31u;
    const uint32_t groups = a_pad / 32;
    static const uint32_t resident = [] {
        int device = 0, sms = 148;
        if (cudaGetDevice(&device) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        return uint32_t(sms) * kGramCtasPerSm;
    }();
    // one wave of CTAs, each with at least one iteration of 128 rows
    const uint32_t splits = uint32_t(std::max<size_t>(1, std::min<size_t>(resident / groups, (n + 127) / 128)));
    ws.GramPartial.Reserve(size_t(splits) * a_pad * kNarrow);
    for (uint32_t c0 = 0; c0 < c; c0 += kNarrow) {
        const uint32_t cw = c - c0 < uint32_t(kNarrow) ? c - c0 : uint32_t(kNarrow);
        GramNarrowPartialKernel<<<dim3(groups, splits), 128, 0, s>>>(X, n, a, Y + size_t(c0) * n, cw, ws.GramPartial.Ptr, a_pad, splits);
        GramNarrowFinalKernel<<<(a * kNarrow * 32 + 255) / 256, 256, 0, s>>>(ws.GramPartial.Ptr, a, cw, a_pad, out + size_t(c0) * ldo, ldo, splits);
        ws.Launches += 2;
    }
}

void ApplyRotations(double *zt, uint32_t m, const QlRotation *rotations, size_t count, cudaStream_t s, uint32_t &launches) {
    if (m == 0 || count == 0) return;
    if (m > kMaxDeviceRotationOrder) Fail(ME_BAD_ARG, "internal: ApplyRotations takes matrices up to order %u (got %u)", kMaxDeviceRotationOrder, m);
    const size_t smem = size_t(m) * 32 * sizeof(double) + 2 * kRotationBatch * sizeof(QlRotation);
    static size_t configured = 0;
    if (smem > configured) {
        ME_CUDA(cudaFuncSetAttribute(ApplyRotationsKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        configured = smem;
    }
    ApplyRotationsKernel<<<(m + 31) / 32, 32, smem, s>>>(zt, m, rotations, count);
    ME_CUDA(cudaGetLastError());
    ++launches;
}

void GatherRows(const double *zt, uint32_t m, const uint32_t *rows, uint32_t k, double *out, cudaStream_t s, uint32_t &launches) {
    if (k == 0 || m == 0) return;
    GatherRowsKernel<<<k, 128, 0, s>>>(zt, m, rows, k, out);
    ME_CUDA(cudaGetLastError());
    ++launches;
}

namespace {
// 32 columns of Q per CTA (a lane each), the rows shared out over kBasisWarps warps (row k belongs to warp k mod kBasisWarps): every
// reflector is a partial dot product per warp, a sum of the partials, an update of the warp's own rows - two CTA barriers. (One
// warp per CTA, as in ApplyRotationsKernel, is bound by its own instruction stream: ~4 cycles an instruction, 2.3 ms at order 328.)
constexpr uint32_t kBasisWarps = 8;
__global__ void __launch_bounds__(32 * kBasisWarps) HouseholderBasisKernel(const double *__restrict__ u, const double *__restrict__ v, const double *__restrict__ h, uint32_t m, double *__restrict__ zt) {
    extern __shared__ double column[]; // [m][32]: entry k of a lane's column of Q; then [kBasisWarps][32] partial dot products
    double *partial = column + size_t(m) * 32;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, j = blockIdx.x * 32 + lane;
    for (uint32_t k = warp; k < m; k += kBasisWarps) column[k * 32 + lane] = k == j ? 1.0 : 0.0;
    __syncthreads();
    // Reflector i changes the leading i entries of the columns j < i; for j >= i those entries are still zero, the dot product with
    // them too, and the update a no-op: the lanes need no predicate, only a common first step.
    for (uint32_t i = blockIdx.x * 32 + 1; i < m; ++i) {
        if (h[i] == 0.0) continue; // (the same decision in every thread)
        const double *ui = u + size_t(i) * m, *vi = v + size_t(i) * m;
        double g0 = 0, g1 = 0;
        uint32_t k = warp;
        for (; k + kBasisWarps < i; k += 2 * kBasisWarps) g0 += ui[k] * column[k * 32 + lane], g1 += ui[k + kBasisWarps] * column[(k + kBasisWarps) * 32 + lane];
        if (k < i) g0 += ui[k] * column[k * 32 + lane];
        partial[warp * 32 + lane] = g0 + g1;
        __syncthreads();
        double g = 0;
#pragma unroll
        for (uint32_t w = 0; w < kBasisWarps; ++w) g += partial[w * 32 + lane];
        for (k = warp; k < i; k += kBasisWarps) column[k * 32 + lane] -= g * vi[k];
        __syncthreads();
    }
    if (j < m)
        for (uint32_t k = warp; k < m; k += kBasisWarps) zt[size_t(j) * m + k] = column[k * 32 + lane];
}
} // namespace

void HouseholderBasis(const double *u, const double *v, const double *h, uint32_t m, double *zt, cudaStream_t s, uint32_t &launches) {
    if (m == 0) return;
    if (m > kMaxDeviceRotationOrder) Fail(ME_BAD_ARG, "internal: HouseholderBasis takes matrices up to order %u (got %u)", kMaxDeviceRotationOrder, m);
    const size_t bytes = (size_t(m) + kBasisWarps) * 32 * sizeof(double);
    ME_CUDA(cudaFuncSetAttribute(HouseholderBasisKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes)));
    HouseholderBasisKernel<<<(m + 31) / 32, 32 * kBasisWarps, bytes, s>>>(u, v, h, m, zt);
    ME_CUDA(cudaGetLastError());
    ++launches;
}

namespace {
constexpr uint32_t kRandomStretch = 64; // consecutive values per thread
__global__ void SimpleRandomKernel(double *__restrict__ out, size_t count) {
    const size_t first = (size_t(blockIdx.x) * blockDim.x + threadIdx.x) * kRandomStretch;
    if (first >= count) return;
    constexpr unsigned long long m = 2147483647ull;
    unsigned long long x = 1, base = 16807ull, e = first + 1; // value i is 16807^(i + 1) mod m
    while (e) {
        if (e & 1) x = x * base % m;
        base = base * base % m;
        e >>= 1;
    }
    const size_t last = min(count, first + kRandomStretch);
    for (size_t i = first; i < last; ++i) {
        out[i] = double(x) / 2147483647.0 - 0.5;
        x = x * 16807ull % m;
    }
}
} // namespace

void FillSimpleRandom(double *out, size_t count, cudaStream_t s, uint32_t &launches) {
    if (count == 0) return;
    const size_t threads = (count + kRandomStretch - 1) / kRandomStretch;
    SimpleRandomKernel<<<uint32_t((threads + 127) / 128), 128, 0, s>>>(out, count);
    ME_CUDA(cudaGetLastError());
    ++launches;
}

} // namespace me
