// sm_100a kernels of the supernodal sparse Cholesky and its triangular solves. See cholesky.h / symbolic.h.
//
// Storage: every supernode S (k columns, m below-diagonal rows, both multiples of 3) owns a dense column-major panel
// of (k + m) x k doubles in HBM: the k x k diagonal block on top, the m x k rectangle below. The factorisation is
// right-looking and level-scheduled on the supernodal tree (levels = height above the leaves):
//   level l:  FactorDiagKernel   one CTA per supernode: Cholesky of the diagonal block in shared memory, fused with the
//                                inverse of the triangular factor (kept for the panel solve and the triangular solves);
//             PanelTrsmKernel    one CTA per 64-row tile: panel <- panel * Linv^T, a GEMM on the FP64 tensor cores;
//             SyrkScatterKernel  one CTA per 64x64 tile of the trailing update panel * panel^T (FP64 DMMA, accumulators
//                                in registers), subtracted straight into the ancestors' panels with FP64 atomics:
//                                no frontal/update matrices are ever materialised in HBM.
// The dense contractions are genuine GEMMs, so they run on the tensor cores: mma.sync.m8n8k4.f64 (DMMA) is the only
// FP64 tensor instruction sm_100a has (tcgen05.mma has no f64 kind; the m16n8k* f64 shapes lower to the same DMMA.8x8x4).
#include "cholesky.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>

namespace me {
namespace {

struct FactorView {
    const uint32_t *SuperFirst, *Rows, *NodeSuper;
    const uint64_t *RowPtr, *PanelOffset, *InvOffset;
    const uint32_t *SegTarget, *SegBegin, *SegEnd;
    double *L, *Linv, *LinvT, *Slabs; // Slabs: the below-diagonal rectangles again, in 32-row slabs of 8 x 8 tiles (Symbolic::SlabOffset), for the sweeps
    int *Fail;
    const uint64_t *SlabOffset;
};

// Position of element (row, col) of a supernode's rectangle inside its slab copy; kp = columns rounded up to 8.
__device__ __forceinline__ size_t SlabIndex(uint32_t row, uint32_t col, uint32_t kp) {
    return size_t(row >> 5) * (kSolveRows * kp) + (((row >> 3) & 3) * (kp >> 3) + (col >> 3)) * 64 + (row & 7) * 8 + (col & 7);
}

__device__ __forceinline__ uint32_t PanelColumns(const FactorView &v, uint32_t s) { return 3 * (v.SuperFirst[s + 1] - v.SuperFirst[s]); }
__device__ __forceinline__ uint32_t PanelRows(const FactorView &v, uint32_t s) { return 3 * uint32_t(v.RowPtr[s + 1] - v.RowPtr[s]); }

// Position of `node` in the below-diagonal node list of supernode s, or 0xFFFFFFFF.
__device__ __forceinline__ uint32_t FindRow(const FactorView &v, uint32_t s, uint32_t node) {
    const uint32_t *rows = v.Rows + v.RowPtr[s];
    uint32_t lo = 0, hi = uint32_t(v.RowPtr[s + 1] - v.RowPtr[s]);
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (rows[mid] < node) lo = mid + 1;
        else hi = mid;
    }
    return (lo < uint32_t(v.RowPtr[s + 1] - v.RowPtr[s]) && rows[lo] == node) ? lo : 0xFFFFFFFFu;
}

__device__ __forceinline__ void Dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// A = K - sigma*M scattered into the (zeroed) panels under the fill-reducing permutation. One thread per stored block.
__global__ void ScatterMatrixKernel(FactorView v, const uint32_t *__restrict__ blk_row, const uint32_t *__restrict__ blk_col, const double *__restrict__ kblk,
                                    const double *__restrict__ mblk, const uint32_t *__restrict__ inv_perm, uint32_t n_blocks, double sigma) {
    const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_blocks) return;
    const uint32_t pr = inv_perm[blk_row[u]], pc = inv_perm[blk_col[u]];
    const bool tr = pr < pc;
    const uint32_t R = tr ? pc : pr, C = tr ? pr : pc;
    const uint32_t s = v.NodeSuper[C], k = PanelColumns(v, s), ld = k + PanelRows(v, s);
    const uint32_t lc = 3 * (C - v.SuperFirst[s]);
    uint32_t lr;
    if (R < v.SuperFirst[s + 1]) lr = 3 * (R - v.SuperFirst[s]);
    else {
        const uint32_t pos = FindRow(v, s, R);
        if (pos == 0xFFFFFFFFu) {
            atomicExch(v.Fail, 2);
            return;
        }
        lr = k + 3 * pos;
    }
    double *panel = v.L + v.PanelOffset[s];
    const double shift = sigma * mblk[u];
#pragma unroll
    for (int p = 0; p < 3; ++p)
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const double a = kblk[size_t(9) * u + 3 * p + q] - (p == q ? shift : 0.0);
            panel[(lr + (tr ? q : p)) + size_t(lc + (tr ? p : q)) * ld] = a;
        }
}

// Cholesky of one diagonal block (k <= 128) and the inverse of its factor, in shared memory, one CTA per supernode. On the
// separator chains of the upper levels this kernel runs alone, a level at a time: it is all latency, and the first version (a
// column at a time over the whole block, the inverse accumulated alongside: 126 steps of ~2 us) was 44 ms of a 300 ms
// factorisation. This one is blocked by 16 columns:
//   Cholesky   per block: the 16 x 16 diagonal block a column at a time on 256 threads (two named barriers per column), the rows
//              below solved against it (a thread per row), then the rank-16 update of the trailing triangle in 4 x 4 register tiles;
//   inverse    in place, last block column first (the dtrtri order): the 16 x 16 diagonal blocks are inverted by a warp each,
//              then block column j becomes -X22 * L21 * X_jj, X22 the already inverted trailing triangle: two barriers per block column.
constexpr int kFactorThreads = 512;
constexpr uint32_t kFactorBlock = 16;
__global__ void __launch_bounds__(kFactorThreads) FactorDiagKernel(FactorView v, const uint32_t *__restrict__ level_supers) {
    extern __shared__ __align__(16) double sm[];
    const uint32_t s = level_supers[blockIdx.x];
    const uint32_t k = PanelColumns(v, s), ld = k + PanelRows(v, s), lds = (k + 1) & ~1u; // (even: the tiles are read 16 bytes at a time)
    double *panel = v.L + v.PanelOffset[s];
    double *scratch = sm + size_t(lds) * k;       // [112 x 16]: a block column's intermediate product
    double *rdiag = scratch + 112 * kFactorBlock; // [k]: reciprocals of the factor's diagonal
    const uint32_t t = threadIdx.x;
    auto A = [&](uint32_t r, uint32_t c) -> double & { return sm[r + size_t(c) * lds]; };
    for (uint32_t idx = t; idx < lds * k; idx += kFactorThreads) {
        const uint32_t r = idx % lds, c = idx / lds;
        sm[idx] = (r >= c && r < k) ? panel[r + size_t(c) * ld] : 0.0;
    }
    bool bad = false;
    for (uint32_t jb = 0; jb < k; jb += kFactorBlock) {
        const uint32_t nb = min(kFactorBlock, k - jb), below = jb + nb, rest = k - below;
        __syncthreads();
        if (t < 256) { // the diagonal block: thread (i, c) owns entry (jb + i, jb + c)
            const uint32_t i = t & 15, c = t >> 4;
            for (uint32_t j = 0; j < nb; ++j) {
                asm volatile("bar.sync 2, 256;" ::: "memory"); // the previous column's writes are visible
                double d = A(jb + j, jb + j);
                if (!(d > 0.0) || !isfinite(d)) {
                    bad = true;
                    d = 1.0;
                }
                const double sq = sqrt(d), inv = 1.0 / sq;
                const double li = (i < nb) ? A(jb + i, jb + j) * inv : 0.0, lc = (c < nb) ? A(jb + c, jb + j) * inv : 0.0;
                asm volatile("bar.sync 2, 256;" ::: "memory"); // every thread has read column j
                if (i < nb && c < nb) {
                    if (c == j) {
                        if (i == j) A(jb + j, jb + j) = sq, rdiag[jb + j] = inv;
                        else if (i > j) A(jb + i, jb + j) = li;
                    } else if (c > j && i >= c) {
                        A(jb + i, jb + c) -= li * lc;
                    }
                }
            }
        }
        __syncthreads();
        if (t < rest) { // rows below: x D^T = a, a thread per row
            const uint32_t r = below + t;
            double x[kFactorBlock];
#pragma unroll
            for (uint32_t c = 0; c < kFactorBlock; ++c) {
                if (c < nb) {
                    double sum = A(r, jb + c);
#pragma unroll
                    for (uint32_t u = 0; u < kFactorBlock; ++u)
                        if (u < c) sum -= x[u] * A(jb + c, jb + u);
                    x[c] = sum * rdiag[jb + c];
                }
            }
#pragma unroll
            for (uint32_t c = 0; c < kFactorBlock; ++c)
                if (c < nb) A(r, jb + c) = x[c];
        }
        __syncthreads();
        // trailing triangle -= P P^T over the block's columns, 4 x 4 tiles of (row, column) >= below
        const uint32_t tiles = (rest + 3) / 4;
        for (uint32_t tile = t; tile < tiles * tiles; tile += kFactorThreads) {
            const uint32_t ti = tile / tiles, tc = tile % tiles;
            if (tc > ti) continue;
            const uint32_t r0 = below + 4 * ti, c0 = below + 4 * tc;
            double acc[4][4]{};
            for (uint32_t u = 0; u < nb; ++u) {
                double a[4], b[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) a[e] = r0 + e < k ? A(r0 + e, jb + u) : 0.0, b[e] = c0 + e < k ? A(c0 + e, jb + u) : 0.0;
#pragma unroll
                for (int e = 0; e < 4; ++e)
#pragma unroll
                    for (int f = 0; f < 4; ++f) acc[e][f] += a[e] * b[f];
            }
#pragma unroll
            for (int e = 0; e < 4; ++e)
#pragma unroll
                for (int f = 0; f < 4; ++f)
                    if (r0 + e < k && c0 + f <= r0 + e) A(r0 + e, c0 + f) -= acc[e][f];
        }
    }
    __syncthreads();
    if (bad && t == 0) atomicExch(v.Fail, 1);
    for (uint32_t idx = t; idx < k * k; idx += kFactorThreads) {
        const uint32_t r = idx % k, c = idx / k;
        if (r >= c) panel[r + size_t(c) * ld] = A(r, c);
    }
    __syncthreads(); // the factor has been read out: it is overwritten from here on
    // Inverse in place. Diagonal blocks first: warp w inverts block w, lane c its column c by forward substitution.
    const uint32_t blocks = (k + kFactorBlock - 1) / kFactorBlock, warp = t >> 5, lane = t & 31;
    if (warp < blocks) {
        const uint32_t jb = warp * kFactorBlock, nb = min(kFactorBlock, k - jb), c = lane;
        double x[kFactorBlock];
#pragma unroll
        for (uint32_t i = 0; i < kFactorBlock; ++i) {
            x[i] = 0.0;
            if (c < nb && i < nb && i >= c) {
                double sum = i == c ? 1.0 : 0.0;
#pragma unroll
                for (uint32_t u = 0; u < kFactorBlock; ++u)
                    if (u < i && u >= c) sum -= A(jb + i, jb + u) * x[u];
                x[i] = sum * rdiag[jb + i];
            }
        }
        __syncwarp();
#pragma unroll
        for (uint32_t i = 0; i < kFactorBlock; ++i)
            if (c < nb && i < nb && i >= c) A(jb + i, jb + c) = x[i];
    }
    for (uint32_t j = blocks - 1; j-- > 0;) {
        const uint32_t jb = j * kFactorBlock, below = jb + kFactorBlock, rest = k - below; // (every block but the last is full)
        __syncthreads();
        // scratch[row][0..15] = X22[row, :] L21: thread (row, four columns)
        for (uint32_t item = t; item < rest * 4; item += kFactorThreads) {
            const uint32_t row = item % rest, quad = item / rest;
            double sum[4]{};
            for (uint32_t l = 0; l <= row; ++l) {
                const double xv = A(below + row, below + l);
#pragma unroll
                for (int e = 0; e < 4; ++e) sum[e] += xv * A(below + l, jb + 4 * quad + e);
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) scratch[row * kFactorBlock + 4 * quad + e] = sum[e];
        }
        __syncthreads();
        // block column j <- -scratch X_jj (X_jj lower triangular)
        for (uint32_t item = t; item < rest * 4; item += kFactorThreads) {
            const uint32_t row = item % rest, quad = item / rest;
            double sum[4]{};
#pragma unroll
            for (uint32_t u = 0; u < kFactorBlock; ++u) {
                const double tv = scratch[row * kFactorBlock + u];
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (u >= 4 * quad + e) sum[e] += tv * A(jb + u, jb + 4 * quad + e);
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) A(below + row, jb + 4 * quad + e) = -sum[e];
        }
    }
    __syncthreads();
    double *linv = v.Linv + v.InvOffset[s], *linvt = v.LinvT + v.InvOffset[s];
    for (uint32_t idx = t; idx < k * k; idx += kFactorThreads) {
        const uint32_t r = idx % k, c = idx / k;
        const double value = r >= c ? A(r, c) : 0.0;
        linv[r + size_t(c) * k] = value;
        linvt[c + size_t(r) * k] = value;
    }
}

// Tile geometry of the DMMA kernels: operands are staged in shared memory K-major, [kk][row] with a row stride of
// 4 (mod 16) doubles, which makes the m8n8k4 fragment loads (lane -> row lane/4, k lane%4) bank-conflict free.
constexpr int kChunk = 16;          // K columns staged per step
constexpr int kLdA = 64 + 4;        // 64-row operand tiles
constexpr int kLdB128 = 128 + 4;    // 128-row operand tile (the inverse factor in the panel solve)

// panel <- panel * Linv^T for one 64-row tile of the below-diagonal rectangle (TRSM as GEMM, C[64 x k]).
constexpr int kTrsmThreads = 256;
__global__ void __launch_bounds__(kTrsmThreads) PanelTrsmKernel(FactorView v, const PanelTile *__restrict__ tiles) {
    extern __shared__ double sm[];
    double *As = sm;                 // [128][kLdA]: the whole 64 x k tile, K-major
    double *Bs = sm + 128 * kLdA;    // [kChunk][kLdB128]
    const PanelTile tile = tiles[blockIdx.x];
    const uint32_t s = tile.Super, k = PanelColumns(v, s), m = PanelRows(v, s), ld = k + m;
    double *p0 = v.L + v.PanelOffset[s] + k;
    const double *linv = v.Linv + v.InvOffset[s];
    double *slabs = v.Slabs + v.SlabOffset[s];
    const uint32_t kp = (k + 7) & ~7u;
    const uint32_t row0 = tile.RowTile * kTile, nrows = min(kTile, m - row0);
    const uint32_t t = threadIdx.x, lane = t & 31, w = t >> 5, wm = w & 1, wn = w >> 1;
    for (uint32_t idx = t; idx < 128 * 64; idx += kTrsmThreads) {
        const uint32_t r = idx & 63, c = idx >> 6;
        As[c * kLdA + r] = (r < nrows && c < k) ? p0[row0 + r + size_t(c) * ld] : 0.0;
    }
    double acc[4][4][2]{};
    for (uint32_t kc = 0; kc < k; kc += kChunk) {
        __syncthreads();
        for (uint32_t idx = t; idx < kChunk * 128; idx += kTrsmThreads) {
            const uint32_t j = idx & 127, c = idx >> 7;
            Bs[c * kLdB128 + j] = (j < k && kc + c < k) ? linv[j + size_t(kc + c) * k] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int ks = 0; ks < kChunk / 4; ++ks) {
            double a[4], b[4];
            const uint32_t kk = 4 * ks + (lane & 3);
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) a[mi] = As[(kc + kk) * kLdA + 32 * wm + 8 * mi + (lane >> 2)];
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) b[ni] = Bs[kk * kLdB128 + 32 * wn + 8 * ni + (lane >> 2)];
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
                if (r < nrows && c < k) {
                    p0[row0 + r + size_t(c) * ld] = acc[mi][ni][e];
                    slabs[SlabIndex(row0 + r, c, kp)] = acc[mi][ni][e];
                }
            }
}

// One 64 x 64 tile of the trailing update of supernode S restricted to one target segment:
//   U = P[rowsA, :] * P[rowsB, :]^T   (rowsB are rows of S that are columns of the ancestor T)
// subtracted into T's panel at the rows/columns the global indices select.
constexpr int kSyrkThreads = 128;
__global__ void __launch_bounds__(kSyrkThreads) SyrkScatterKernel(FactorView v, const UpdateTile *__restrict__ tiles) {
    __shared__ double As[kChunk * kLdA], Bs[kChunk * kLdA];
    __shared__ uint32_t row_dest[kTile], col_dest[kTile];
    const UpdateTile tile = tiles[blockIdx.x];
    const uint32_t s = tile.Super, g = tile.Segment;
    const uint32_t k = PanelColumns(v, s), m = PanelRows(v, s), ld = k + m;
    const double *p0 = v.L + v.PanelOffset[s] + k;
    const uint32_t c0 = 3 * v.SegBegin[g], c1 = 3 * v.SegEnd[g];
    const uint32_t row_a = c0 + tile.RowTile * kTile, row_b = c0 + tile.ColTile * kTile;
    const uint32_t na = min(kTile, m - row_a), nb = min(kTile, c1 - row_b);
    const uint32_t target = v.SegTarget[g], kt = PanelColumns(v, target), ldt = kt + PanelRows(v, target);
    const uint32_t t = threadIdx.x, lane = t & 31, w = t >> 5, wm = w & 1, wn = w >> 1;
    {
        const uint32_t *rows = v.Rows + v.RowPtr[s];
        if (t < kTile) {
            if (t < na) {
                const uint32_t br = row_a + t, node = rows[br / 3], comp = br % 3;
                uint32_t lr;
                if (node < v.SuperFirst[target + 1]) lr = 3 * (node - v.SuperFirst[target]) + comp;
                else {
                    const uint32_t pos = FindRow(v, target, node);
                    if (pos == 0xFFFFFFFFu) atomicExch(v.Fail, 3);
                    lr = pos == 0xFFFFFFFFu ? 0 : kt + 3 * pos + comp;
                }
                row_dest[t] = lr;
            }
        } else {
            const uint32_t j = t - kTile;
            if (j < nb) {
                const uint32_t bc = row_b + j, node = rows[bc / 3], comp = bc % 3;
                col_dest[j] = (3 * (node - v.SuperFirst[target]) + comp) * ldt;
            }
        }
    }
    double acc[4][4][2]{};
    for (uint32_t kc = 0; kc < k; kc += kChunk) {
        __syncthreads();
        for (uint32_t idx = t; idx < kChunk * kTile; idx += kSyrkThreads) {
            const uint32_t r = idx & 63, c = idx >> 6;
            const bool kin = kc + c < k;
            As[c * kLdA + r] = (kin && r < na) ? p0[row_a + r + size_t(kc + c) * ld] : 0.0;
            Bs[c * kLdA + r] = (kin && r < nb) ? p0[row_b + r + size_t(kc + c) * ld] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int ks = 0; ks < kChunk / 4; ++ks) {
            double a[4], b[4];
            const uint32_t kk = 4 * ks + (lane & 3);
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) a[mi] = As[kk * kLdA + 32 * wm + 8 * mi + (lane >> 2)];
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) b[ni] = Bs[kk * kLdA + 32 * wn + 8 * ni + (lane >> 2)];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) Dmma(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
        }
    }
    double *lt = v.L + v.PanelOffset[target];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const uint32_t r = 32 * wm + 8 * mi + (lane >> 2), c = 32 * wn + 8 * ni + 2 * (lane & 3) + e;
                if (r < na && c < nb && row_a + r >= row_b + c) atomicAdd(lt + row_dest[r] + size_t(col_dest[c]), -acc[mi][ni][e]);
            }
}

// ------------------------------------------------------------------------------------------------ macro block inverses
// The panel sweeps solve up to SymbolicOptions::MacroPanels consecutive panels of a separator chain in one step (symbolic.h):
// W = inverse of the macro block's lower-triangular diagonal block, built here from the factor by block forward substitution,
//   W_jj = Linv_j,   W_ij = -Linv_i * sum_{j <= l < i} L_il W_lj   (i > j),
// one CTA per (block column j, 64-column slice): the columns of W are independent. Each W_ij goes to the forward row block of
// panel i and, transposed, to the backward row block of panel j. All products are 128 x 64 DMMA tiles.
struct MacroView {
    const uint32_t *SuperFirst;
    const uint64_t *RowPtr, *PanelOffset, *InvOffset, *MacroOffset, *MacroOffsetT;
    const double *L, *Linv;
    double *W, *WT;
};
constexpr int kMacroThreads = 256;

// acc[128 x 64] += A[M x K] B[K x N], both column-major in global memory, M, K <= 128, N <= 64.
__device__ __forceinline__ void MacroGemm(const double *__restrict__ A, uint32_t lda, uint32_t M, uint32_t K, const double *B, uint32_t ldb, uint32_t N, double *As, double *Bs,
                                          double (&acc)[4][4][2]) {
    const uint32_t t = threadIdx.x, lane = t & 31, w = t >> 5, wm = w & 3, wn = w >> 2;
    for (uint32_t kc = 0; kc < K; kc += kChunk) {
        __syncthreads();
        for (uint32_t idx = t; idx < kChunk * 128; idx += kMacroThreads) {
            const uint32_t r = idx & 127, c = idx >> 7;
            As[c * kLdB128 + r] = (r < M && kc + c < K) ? A[r + size_t(kc + c) * lda] : 0.0;
        }
        for (uint32_t idx = t; idx < kChunk * kTile; idx += kMacroThreads) {
            const uint32_t kk = idx & (kChunk - 1), n = idx / kChunk;
            Bs[kk * kLdA + n] = (kc + kk < K && n < N) ? B[kc + kk + size_t(n) * ldb] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int ks = 0; ks < kChunk / 4; ++ks) {
            double a[4], b[4];
            const uint32_t kk = 4 * ks + (lane & 3);
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) a[mi] = As[kk * kLdB128 + 32 * wm + 8 * mi + (lane >> 2)];
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) b[ni] = Bs[kk * kLdA + 32 * wn + 8 * ni + (lane >> 2)];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) Dmma(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
        }
    }
}

__global__ void __launch_bounds__(kMacroThreads) MacroInverseKernel(MacroView v, const Symbolic::MacroJob *__restrict__ jobs) {
    __shared__ double As[kChunk * kLdB128], Bs[kChunk * kLdA];
    const Symbolic::MacroJob job = jobs[blockIdx.x];
    const uint32_t t = threadIdx.x, lane = t & 31, w = t >> 5, wm = w & 3, wn = w >> 2;
    auto columns = [&](uint32_t s) { return 3 * (v.SuperFirst[s + 1] - v.SuperFirst[s]); };
    auto start = [&](uint32_t s) { return 3 * (v.SuperFirst[s] - v.SuperFirst[job.First]); }; // a panel's first column inside the macro block
    const uint32_t j = job.Column, kj = columns(j), off_j = start(j), n0 = job.Slice * kTile, nn = min(uint32_t(kTile), kj - n0);
    double *wb = v.WT + v.MacroOffsetT[j]; // backward row block of panel j: element (r, c) = W(off_j + c, off_j + r), leading dimension k_j
    {
        const double *linv = v.Linv + v.InvOffset[j];
        double *wf = v.W + v.MacroOffset[j] + size_t(off_j + n0) * kj;
        for (uint32_t idx = t; idx < kj * nn; idx += kMacroThreads) {
            const uint32_t r = idx % kj, c = idx / kj;
            const double value = linv[r + size_t(n0 + c) * kj];
            wf[r + size_t(c) * kj] = value;
            wb[(n0 + c) + size_t(r) * kj] = value;
        }
    }
    for (uint32_t i = j + 1; i <= job.Last; ++i) {
        const uint32_t ki = columns(i), off_i = start(i);
        double acc[4][4][2]{};
        for (uint32_t l = j; l < i; ++l) {
            const uint32_t kl = columns(l), off_l = start(l), ld_l = kl + 3 * uint32_t(v.RowPtr[l + 1] - v.RowPtr[l]);
            // the rows of panel i inside panel l's rectangle: the chain's next panels are the first rows of its list, in order
            const double *a = v.L + v.PanelOffset[l] + kl + (off_i - off_l - kl);
            const double *b = v.W + v.MacroOffset[l] + size_t(off_j + n0) * kl;
            MacroGemm(a, ld_l, ki, kl, b, kl, nn, As, Bs, acc);
        }
        double *wf = v.W + v.MacroOffset[i] + size_t(off_j + n0) * ki; // W_ij's slice: k_i x nn; holds the sum first
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const uint32_t r = 32 * wm + 8 * mi + (lane >> 2), c = 32 * wn + 8 * ni + 2 * (lane & 3) + e;
                    if (r < ki && c < nn) wf[r + size_t(c) * ki] = acc[mi][ni][e];
                    acc[mi][ni][e] = 0.0;
                }
        MacroGemm(v.Linv + v.InvOffset[i], ki, ki, ki, wf, ki, nn, As, Bs, acc);
        __syncthreads(); // every read of the sum is done
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const uint32_t r = 32 * wm + 8 * mi + (lane >> 2), c = 32 * wn + 8 * ni + 2 * (lane & 3) + e;
                    if (r < ki && c < nn) {
                        wf[r + size_t(c) * ki] = -acc[mi][ni][e];
                        wb[(n0 + c) + size_t(off_i - off_j + r) * kj] = -acc[mi][ni][e];
                    }
                }
    }
}

// ------------------------------------------------------------------------------------------------ triangular solves
__global__ void PermuteOutKernel(const double *__restrict__ w, const uint32_t *__restrict__ inv_perm, uint32_t n_nodes, double *__restrict__ x) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 3 * n_nodes) x[i] = w[3 * inv_perm[i / 3] + i % 3];
}

// Load shape shared by the solve tasks: a 64-row x k-column slab (k <= 128) is covered by 256 threads, and every thread
// issues its 32 loads before using any, so a task costs about one L2/HBM round trip instead of k dependent ones.
constexpr int kSolveThreads = 256;

// ------------------------------------------------------------------------------------------------ dataflow sweeps
// A level-synchronous sweep (one or two launches per level of the tree) pays a launch gap and a ramp-up/tail per kernel,
// ~360 times per sweep on the 1M-tet mesh, and makes every supernode of a level wait for the slowest one. Here a whole
// sweep is ONE launch of persistent CTAs:
//   * work is cut into uniform slabs of 32 rows x k columns (of a diagonal block's inverse, or of a panel) listed in a
//     topological order (level by level); CTAs take them from a ticket counter;
//   * a task waits only for its own inputs. A task's inputs always hold smaller tickets and every CTA is resident, so the
//     spin-waits cannot deadlock (they are bounded all the same: a sweep that stalls raises Fail instead of hanging);
//   * a task is a short chain of dependent L2 round trips, so what keeps HBM busy is the number of chains in flight and
//     how much of each chain overlaps the 32 KB matrix load: the CTAs are small (128 threads, every thread issues its 32
//     loads before using any), five share an SM, and the chain is kept short:
//       - solved entries are SELF-VALIDATING: `out` is pre-filled with a NaN sentinel, a diagonal slab just stores its
//         results, and a consumer polls the entries it needs until none is the sentinel — no fence, no flag, and the
//         poll is the load;
//       - only the many-to-one direction (panel slabs accumulating into a supernode's entries with FP64 atomics) keeps a
//         counter, and its publication (fence + increment) is deferred until the NEXT task's matrix loads are in
//         flight, so the fence's round trip overlaps them.
// Forward: `acc` holds the right-hand side that panel slabs update, diagonal slabs write y into `out`.
// Backward: `acc` holds y, panel slabs subtract P^T x from it, diagonal slabs write x into `out`.
__device__ __forceinline__ uint32_t Peek(const uint32_t *p) { return *reinterpret_cast<const volatile uint32_t *>(p); }

constexpr unsigned long long kUnsolved = 0xFFFBADC0FFEE0DD5ull; // quiet NaN with a payload arithmetic never produces
constexpr uint32_t kSpinLimit = 1u << 19;                        // ~0.3 s of polling: far beyond any legitimate wait

__device__ __forceinline__ unsigned long long LoadL2(const double *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.b64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void StoreL2(double *p, double v) {
    asm volatile("st.relaxed.gpu.global.b64 [%0], %1;" ::"l"(p), "l"(__double_as_longlong(v)) : "memory");
}
// A wait gives up when it has spun past the limit, or (checked now and then) once any other wait has: a stalled sweep
// drains quickly instead of timing out task by task.
__device__ __forceinline__ bool GiveUp(uint32_t spin, int *fail) {
    if (spin > kSpinLimit) {
        atomicExch(fail, 4);
        return true;
    }
    return (spin & 1023u) == 1023u && *reinterpret_cast<volatile int *>(fail) != 0;
}
// Polls one solved entry until it is published. Returns 0.0 (and raises Fail) if it never is.
__device__ __forceinline__ double AwaitSolved(const double *p, int *fail) {
    unsigned long long v = LoadL2(p);
    for (uint32_t spin = 0; v == kUnsolved; ++spin) {
        if (GiveUp(spin, fail)) return 0.0;
        __nanosleep(20);
        v = LoadL2(p);
    }
    return __longlong_as_double((long long)v);
}

struct SweepArgs {
    const SweepTask *Tasks;
    uint32_t NumTasks;
    const uint32_t *Links;            // forward panel slabs: the ancestors they update; backward (panel sweeps): the ancestors they read
    const uint32_t *LinkNeed;         // backward panel sweeps: diagonal slabs of each such ancestor
    uint32_t *Ticket, *Arrived;
    uint32_t *Solved;                 // panel sweeps: diagonal slabs of each supernode whose results are published
    const uint32_t *Rows;             // below-diagonal node lists of all supernodes
    const double *Diag, *Panel;       // Linv + L (forward) or Linv^T + the slab copy (backward; the panel sweeps: the slab copy both ways)
    double *Acc, *Out;
    int *Fail;
    const double *Macro;              // panel sweeps: row blocks of the macro blocks' inverses (forward) / of their transposes (backward)
    const uint32_t *ArriveNeed;       // panel sweeps: arrivals every supernode's entries wait for (what a macro slab polls for each panel it reads)
    unsigned long long *Trace;        // ME_SWEEP_TRACE: per task, %globaltimer when it was taken up, when its inputs were complete, when it ended
};
__device__ __forceinline__ unsigned long long GlobalTimer() {
    unsigned long long v;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v)::"memory"); // (the clobber keeps the read where it is written: between the barriers)
    return v;
}
constexpr int kSweepThreads = 128;
constexpr int kSweepCtasPerSm = 5;

template<bool Backward>
__global__ void __launch_bounds__(kSweepThreads, kSweepCtasPerSm) SweepKernel(SweepArgs a) {
    __shared__ double vec[128], part[kSweepThreads];
    __shared__ SweepTask s_task;
    __shared__ uint32_t s_id;
    const uint32_t t = threadIdx.x, r = t & 31, q = t >> 5;
    // Warp 0 keeps the NEXT ticket and its 64-byte descriptor (8 lanes x 8 bytes) in registers, fetched while the
    // current task runs, so that neither the ticket atomic nor the descriptor load sits on the per-task chain.
    uint32_t next_id = 0;
    uint64_t next_word = 0;
    auto prefetch_next = [&] {
        if (t < 32) {
            if (t == 0) next_id = atomicAdd(a.Ticket, 1u);
            next_id = __shfl_sync(0xffffffffu, next_id, 0);
            if (t < 8 && next_id < a.NumTasks) next_word = reinterpret_cast<const uint64_t *>(a.Tasks + next_id)[t];
        }
    };
    // Deferred publication of the previous panel slab's contributions: tail_any is CTA-uniform, tail_mine marks the
    // threads that own one counter increment each.
    bool tail_any = false, tail_mine = false;
    uint32_t tail_target = 0;
    auto publish_tail = [&] {
        if (tail_any) {
            __threadfence();
            __syncthreads();
            if (tail_mine) atomicAdd(a.Arrived + tail_target, 1u);
            tail_any = tail_mine = false;
        }
    };
    // Waits (thread 0) until every contribution to the supernode's entries of `acc` has been published.
    auto await_arrivals = [&](uint32_t super, uint32_t need) {
        if (t == 0) {
            for (uint32_t spin = 0; Peek(a.Arrived + super) < need; ++spin) {
                if (GiveUp(spin, a.Fail)) break;
                __nanosleep(20);
            }
            __threadfence();
        }
        __syncthreads();
    };
    prefetch_next();
    for (;;) {
        __syncthreads();
        if (t < 8) reinterpret_cast<uint64_t *>(&s_task)[t] = next_word;
        if (t == 0) s_id = next_id;
        __syncthreads();
        const uint32_t id = s_id;
        if (id >= a.NumTasks) {
            publish_tail();
            return;
        }
        const SweepTask task = s_task;
        const uint32_t k = task.K;
        double val[32];
        if (task.Kind == 0) {
            // out_S[slab rows] = T_S[slab rows, :] acc_S once every contribution to acc_S has arrived. The matrix slab
            // does not depend on anything: it is requested before the wait.
            const double *mat = a.Diag + task.Base;
            const uint32_t row = task.Row0 + r;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const uint32_t col = q + 4 * j;
                const bool in = row < k && col < k && (Backward ? col >= row : col <= row);
                val[j] = in ? mat[row + size_t(col) * k] : 0.0;
            }
            prefetch_next();
            publish_tail();
            await_arrivals(task.Super, task.Need);
            vec[t] = t < k ? __ldcg(a.Acc + task.VecOffset + t) : 0.0;
            __syncthreads();
            double sum = 0;
#pragma unroll
            for (int j = 0; j < 32; ++j) sum += val[j] * vec[q + 4 * j];
            part[t] = sum;
            __syncthreads();
            if (t < 32 && row < k) StoreL2(a.Out + task.VecOffset + row, (part[t] + part[32 + t]) + (part[64 + t] + part[96 + t]));
        } else if constexpr (!Backward) {
            // acc[slab rows] -= P_slab out_S once out_S is complete. Panel: column-major, leading dimension k + m.
            const uint32_t row = task.Row0 + r;
            const double *p0 = a.Panel + task.Base;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const uint32_t col = q + 4 * j;
                val[j] = (row < task.Limit && col < k) ? p0[row + size_t(col) * task.Ld] : 0.0;
            }
            uint32_t node = 0, link = 0;
            if (t < 32 && row < task.Limit) node = a.Rows[task.RowsBase + row / 3];
            if (t < task.LinkCount) link = a.Links[task.LinkBegin + t];
            prefetch_next();
            publish_tail();
            vec[t] = t < k ? AwaitSolved(a.Out + task.VecOffset + t, a.Fail) : 0.0;
            __syncthreads();
            double sum = 0;
#pragma unroll
            for (int j = 0; j < 32; ++j) sum += val[j] * vec[q + 4 * j];
            part[t] = sum;
            __syncthreads();
            if (t < 32 && row < task.Limit) atomicAdd(a.Acc + size_t(3) * node + row % 3, -((part[t] + part[32 + t]) + (part[64 + t] + part[96 + t])));
            tail_any = true, tail_mine = t < task.LinkCount, tail_target = link;
        } else {
            // acc_S -= P_slab^T out[slab rows] once the ancestors owning those rows are solved. Read from the transposed
            // panel copy ([row][column]): thread = column, so the sum over the slab's rows stays in one register.
            const double *pt = a.Panel + task.Base + (t >> 3) * 64 + (t & 7); // the task's slab of the slab copy: 8 x 8 tiles, [strip][column tile]
            const uint32_t strip = ((k + 7) >> 3) * 64;
#pragma unroll
            for (int j = 0; j < 32; ++j) val[j] = (task.Row0 + j < task.Limit && t < k) ? pt[(j >> 3) * strip + (j & 7) * 8] : 0.0;
            uint32_t node = 0;
            if (t < kSolveRows && task.Row0 + t < task.Limit) node = a.Rows[task.RowsBase + (task.Row0 + t) / 3];
            prefetch_next();
            publish_tail();
            if (t < kSolveRows) vec[t] = task.Row0 + t < task.Limit ? AwaitSolved(a.Out + size_t(3) * node + (task.Row0 + t) % 3, a.Fail) : 0.0;
            __syncthreads();
            double sum = 0;
#pragma unroll
            for (int j = 0; j < 32; ++j) sum += val[j] * vec[j];
            if (t < k) atomicAdd(a.Acc + task.VecOffset + t, -sum);
            tail_any = true, tail_mine = t == 0, tail_target = task.Super;
        }
    }
}

// ------------------------------------------------------------------------------------------------ panel sweeps
// The same dataflow sweep for a PANEL of kWide right-hand sides (solve_panel, CholeskyShiftInvert.cpp:55-62): one pass
// over the factor serves all of them, so the bytes per right-hand side drop by kWide while the sweep stays HBM-bound.
// Vectors are stored [permuted DOF][kWide] (one 64-byte row per DOF: a gathered/scattered row is one full sector pair).
// Every slab product is a small dense contraction (32 x k times k x 8, or k x 32 times 32 x 8) and runs on the FP64
// tensor cores: each warp owns a quarter of it (forward/diagonal: 8 of the 32 rows; backward: a quarter of the output
// columns), loads its 32 A-fragments straight from global memory in the m8n8k4 fragment layout (all 32 loads in flight
// before anything waits), and takes B from the shared k x 8 panel (backward: from the solved entries themselves).
constexpr int kWide = 8;

__device__ __forceinline__ void LoadL2x2(const double *p, unsigned long long &a, unsigned long long &b) {
    asm volatile("ld.relaxed.gpu.global.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
// Polls a pair of adjacent solved entries (16-byte aligned) until both are published.
__device__ __forceinline__ void AwaitSolved2(const double *p, int *fail, double &x, double &y) {
    unsigned long long a, b;
    LoadL2x2(p, a, b);
    for (uint32_t spin = 0; a == kUnsolved || b == kUnsolved; ++spin) {
        if (GiveUp(spin, fail)) {
            a = b = 0;
            break;
        }
        __nanosleep(20);
        LoadL2x2(p, a, b);
    }
    x = __longlong_as_double((long long)a), y = __longlong_as_double((long long)b);
}

// Handshakes of the panel sweeps (what the ncu capture of the first version showed the warps waiting on: polls issued one
// after the other, a full fence per task that also waited for the task's own matrix loads, the ticket atomic):
//   * arrivals are published with ONE release-increment per link by a warp that issues it right after the CTA barrier
//     that follows the contributions (st/red.release.gpu after __syncthreads is cumulative over the CTA's writes: the
//     split-K semaphore pattern), and awaited with ld.acquire.gpu polls: no __threadfence anywhere;
//   * results are published the same way: a diagonal slab stores its rows of `out`, and after the CTA barrier ONE
//     release-increment of Solved[supernode] announces them; a panel slab waits with ONE polling thread per counter it
//     depends on and then loads the solved entries once. (The first version made every solved entry self-validating - a NaN
//     sentinel polled by all 128 threads of every waiting CTA, 8 KB per CTA per poll round: with the 740 resident CTAs
//     mostly waiting on the upper levels' dependency chains, the polls alone loaded L2 with terabytes per second.)
//   * tickets are claimed two tasks ahead, so neither the atomic nor the descriptor load is ever waited for.
__device__ __forceinline__ void ArriveRelease(uint32_t *counter) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
}
__device__ __forceinline__ uint32_t PeekAcquire(const uint32_t *counter) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    return v;
}

// Third form of the kernel (the timelines written by ME_SWEEP_TRACE decided it): with every warp loading its own fragments
// from global memory, a CTA had its 32 KB of loads in flight for about half of its ~4 us per slab, and the 600-740 resident CTAs
// together streamed the factor at 2.5-3.4 TB/s wherever the tree was wide, working 85 % of the time. A probe of the same access
// shape without the arithmetic reaches 6 TB/s, and whole slabs copied by cp.async.bulk into a shared-memory ring reach 7.5 TB/s
// (scripts/probes/stream_patterns.cu). So the matrix now travels on its own:
//   * the below-diagonal rectangles are kept a second time as contiguous 32-row slabs of 8 x 8 tiles (FactorView::Slabs,
//     written by PanelTrsmKernel); a slab is ONE bulk copy;
//   * a PRODUCER warp claims the tickets, hands each task's descriptor to the consumers through a small queue in shared memory
//     and streams the task's slabs into a ring of kWideStages buffers (mbarrier complete_tx / per-stage empty barriers). The
//     matrix has no dependencies, so the producer runs ahead of the consumers, into the next tasks, as far as the ring allows:
//     while the consumers wait for a task's inputs, its slabs and the next tasks' are already arriving;
//   * four CONSUMER warps do what the warps of the second form did (forward: a warp owns the 8-row strip q of every slab and its
//     8 x 8 product goes straight to the atomics; backward: a warp owns 32 output columns and gathers the solved entries of a
//     slab's rows itself), reading fragments from shared memory: a forward tile is one conflict-free 16-byte load per lane
//     feeding two DMMAs (columns {0,2,4,6} and {1,3,5,7} of the tile: the contraction does not care about the order), a backward
//     half-tile is 256 contiguous bytes. They meet only at a task's handshakes (a named barrier of the 128 consumer threads).
// Diagonal / macro slabs still load their fragments from global memory (4 % of the bytes).
// Measured on the 1M-tet factor (scripts/gpu_sweep_ab.py, one process): 4.7 ms per panel application against 5.3-5.8 (first form) and
// 6.4 (second form). Tried on top of it and dropped: two stages with three CTAs per SM (5.1 ms), the producer holding three tickets
// (5.2-6.1 ms: see the producer loop), eight copies of the backward accumulator to spread the runs' atomics (7.3 ms: the diagonal
// slabs, which are every level's critical path, had eight times the entries to load); with the atomics skipped altogether a panel
// application is 4.6 ms, so they are not what is left. What is: the ~90 narrow upper levels, where a level is one diagonal slab
// and one round of single-slab tasks of ~2.5 us each (1.4 of the 4.7 ms for 8 % of the bytes).
constexpr int kWideThreads = 160, kWideConsumers = 128;
constexpr int kWideStages = 3, kWideQueue = 4;
constexpr uint32_t kSlabDoubles = kSolveRows * 128;
struct WideShared {
    double Ring[kWideStages][kSlabDoubles]; // slabs in flight / being consumed
    double Vec[128 * kWide];                // the task's k x 8 right-hand operand
    SweepTask Queue[kWideQueue];
    uint32_t QueueId[kWideQueue];           // the tasks' tickets (for the timeline)
    uint64_t Full[kWideStages], Empty[kWideStages], QueueFull[kWideQueue], QueueEmpty[kWideQueue];
};
constexpr uint32_t kTerminator = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t SmemAddr(const void *p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void MbarInit(uint64_t *bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(SmemAddr(bar)), "r"(count)); }
__device__ __forceinline__ void MbarExpectTx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(SmemAddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void MbarArrive(uint64_t *bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(SmemAddr(bar)) : "memory"); }
__device__ __forceinline__ void MbarWait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra MBAR_DONE;\n"
        "bra MBAR_WAIT;\n"
        "MBAR_DONE:\n"
        "}" ::"r"(SmemAddr(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void BulkLoad(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(SmemAddr(dst)), "l"(src), "r"(bytes), "r"(SmemAddr(bar))
                 : "memory");
}
__device__ __forceinline__ void ConsumerBarrier() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

template<bool Backward>
__global__ void __launch_bounds__(kWideThreads, 2) WideSweepKernel(SweepArgs a) {
    extern __shared__ __align__(128) unsigned char wide_shared[];
    WideShared &sh = *reinterpret_cast<WideShared *>(wide_shared);
    const uint32_t t = threadIdx.x, lane = t & 31, q = t >> 5, fr = lane >> 2, fk = lane & 3;
    if (t == 0) {
        for (int i = 0; i < kWideStages; ++i) MbarInit(sh.Full + i, 1), MbarInit(sh.Empty + i, 4);
        for (int i = 0; i < kWideQueue; ++i) MbarInit(sh.QueueFull + i, 1), MbarInit(sh.QueueEmpty + i, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (q == 4) {
        // ---------------------------------------------------------------------------------------------------- producer
        if (lane != 0) return;
        uint32_t issued = 0; // slabs sent into the ring so far (the consumers count the same slabs)
        // One ticket at a time, claimed when the previous task's slabs have all been sent: a version that kept three tickets in
        // hand (atomic and descriptor never waited for) was 10-30 % SLOWER - a ticket claimed early and served late holds up every
        // task that depends on it.
        for (uint32_t n = 0;; ++n) {
            const uint32_t id = atomicAdd(a.Ticket, 1u);
            const bool done = id >= a.NumTasks;
            uint4 w[4]{};
            if (!done) {
                const uint4 *src = reinterpret_cast<const uint4 *>(a.Tasks + id);
#pragma unroll
                for (int i = 0; i < 4; ++i) w[i] = __ldg(src + i);
            } else {
                w[0].z = kTerminator; // SweepTask::Kind
            }
            const uint32_t slot = n % kWideQueue, turn = n / kWideQueue;
            if (turn > 0) MbarWait(sh.QueueEmpty + slot, (turn - 1) & 1);
            uint4 *dst = reinterpret_cast<uint4 *>(sh.Queue + slot);
#pragma unroll
            for (int i = 0; i < 4; ++i) dst[i] = w[i];
            sh.QueueId[slot] = id;
            MbarArrive(sh.QueueFull + slot); // (release: the descriptor is visible to whoever sees the phase)
            if (done) return;
            if (w[0].z == 1) { // SweepTask: Base (w0.x, w0.y), Kind w0.z, Super w0.w, K w1.x, ..., Count w3.y
                const uint64_t base = uint64_t(w[0].x) | (uint64_t(w[0].y) << 32);
                const uint32_t slab = kSolveRows * ((w[1].x + 7) & ~7u), count = w[3].y;
                for (uint32_t i = 0; i < count; ++i, ++issued) {
                    const uint32_t stage = issued % kWideStages, round = issued / kWideStages;
                    if (round > 0) MbarWait(sh.Empty + stage, (round - 1) & 1);
                    MbarExpectTx(sh.Full + stage, slab * 8);
                    BulkLoad(sh.Ring[stage], a.Panel + base + size_t(i) * slab, slab * 8, sh.Full + stage);
                }
            }
        }
    }
    // ------------------------------------------------------------------------------------------------------- consumers
    double *vec = sh.Vec;
    uint32_t consumed = 0; // slabs taken out of the ring so far
    // C[8 x 8] += A-fragments val[ks] (the strip's rows, k-steps 4 ks .. 4 ks + 3) times vec, on four accumulator chains.
    auto contract_strip = [&](const double (&val)[32], uint32_t ksteps, double &c0, double &c1) {
        double acc[4][2]{};
#pragma unroll
        for (int ks = 0; ks < 32; ++ks)
            if (uint32_t(ks) < ksteps) Dmma(acc[ks & 3][0], acc[ks & 3][1], val[ks], vec[(4 * ks + fk) * kWide + fr]);
        c0 += (acc[0][0] + acc[1][0]) + (acc[2][0] + acc[3][0]);
        c1 += (acc[0][1] + acc[1][1]) + (acc[2][1] + acc[3][1]);
    };
    auto stamp = [&](uint32_t id, uint32_t slot) {
        if (a.Trace && t == 0) a.Trace[size_t(3) * id + slot] = GlobalTimer();
    };
    for (uint32_t n = 0;; ++n) {
        const uint32_t slot = n % kWideQueue;
        MbarWait(sh.QueueFull + slot, (n / kWideQueue) & 1);
        const SweepTask task = sh.Queue[slot];
        const uint32_t id = sh.QueueId[slot];
        __syncwarp();
        if (lane == 0) MbarArrive(sh.QueueEmpty + slot);
        if (task.Kind == kTerminator) return;
        stamp(id, 0);
        const uint32_t k = task.K;
        // Counter increments owed by this task, held by warp 3 (one counter per lane): arrivals of a panel run's contributions, or
        // the publication of a diagonal slab's results. Released after the consumers' barrier that ends the task.
        uint32_t *owed = nullptr;
        if (task.Kind != 1) {
            // A 32-row slab of a diagonal block's inverse (Kind 0: k columns) or of a macro block's inverse (Kind 2: the row block of
            // this panel, task.Limit columns over the entries of task.LinkCount consecutive panels, walked in chunks of 128 columns):
            //   forward   out_S[rows] = sum over the block's panels up to S of  W[rows of S, their columns] acc[their entries]
            //   backward  out_S[rows] = sum over the block's panels from S on of W^T[...]
            // Forward the own diagonal block is the LAST of the row block (it starts at column DiagColumn), backward the first.
            // Warp q computes rows Row0 + 8 q .. + 7 from fragments it loads itself.
            const bool macro = task.Kind == 2;
            const double *mat = (macro ? a.Macro : a.Diag) + task.Base;
            const uint32_t diag_col = Backward ? 0u : task.DiagColumn;
            const uint32_t in_base = task.VecOffset - diag_col;
            const uint32_t col_end = Backward ? task.Limit : min(task.Limit, diag_col + task.Row0 + kSolveRows);
            const uint32_t row = task.Row0 + 8 * q + fr;
            double val[32];
            auto load_chunk = [&](uint32_t c0) {
#pragma unroll
                for (int ks = 0; ks < 32; ++ks) {
                    const uint32_t col = c0 + 4 * ks + fk;
                    const bool in = row < k && col < col_end && (Backward ? col >= row : col <= diag_col + row);
                    val[ks] = in ? mat[row + size_t(col) * k] : 0.0;
                }
            };
            load_chunk(0);
            if (!macro) {
                if (t == 96) {
                    for (uint32_t spin = 0; PeekAcquire(a.Arrived + task.Super) < task.Need; ++spin) {
                        if (GiveUp(spin, a.Fail)) break;
                        __nanosleep(20);
                    }
                }
            } else if (q == 3 && lane < task.LinkCount) { // one lane per panel of the block whose entries this slab reads
                const uint32_t super = Backward ? task.Super + lane : task.Super - (task.LinkCount - 1) + lane;
                const uint32_t need = a.ArriveNeed[super];
                for (uint32_t spin = 0; PeekAcquire(a.Arrived + super) < need; ++spin) {
                    if (GiveUp(spin, a.Fail)) break;
                    __nanosleep(20);
                }
            }
            double c0 = 0, c1 = 0;
            for (uint32_t col0 = 0; col0 < col_end; col0 += 128) {
                ConsumerBarrier(); // the entries are complete (first chunk) / the previous chunk's operand has been consumed
                if (col0 == 0) stamp(id, 1);
                {
                    const double2 *src = reinterpret_cast<const double2 *>(a.Acc + (size_t(in_base) + col0 + t) * kWide);
                    double2 *dst = reinterpret_cast<double2 *>(vec + t * kWide);
#pragma unroll
                    for (int j = 0; j < kWide / 2; ++j) dst[j] = col0 + t < col_end ? __ldcg(src + j) : make_double2(0.0, 0.0);
                }
                ConsumerBarrier();
                contract_strip(val, (min(col_end - col0, 128u) + 3) >> 2, c0, c1);
                if (col0 + 128 < col_end) load_chunk(col0 + 128);
            }
            if (row < k) {
                double *dst = a.Out + (size_t(task.VecOffset) + row) * kWide + 2 * fk;
                StoreL2(dst, c0);
                StoreL2(dst + 1, c1);
            }
            if (t == 96) owed = a.Solved + task.Super;
        } else if constexpr (!Backward) {
            // A run of task.Count consecutive slabs of one panel: acc[rows] -= P[rows, :] out_S. The run shares the wait for out_S,
            // the k x 8 operand in shared memory and the arrivals it owes; warp q takes the 8-row strip q of every slab.
            const uint32_t tiles = (k + 7) >> 3, row_end = min(task.Limit, task.Row0 + task.Count * kSolveRows);
            auto node_of = [&](uint32_t row) { return row >= task.RowMin && row < row_end ? __ldg(a.Rows + task.RowsBase + row / 3) : 0u; };
            uint32_t row = task.Row0 + 8 * q + fr, node = node_of(row);
            if (q == 3 && lane < task.LinkCount) owed = a.Arrived + a.Links[task.LinkBegin + lane];
            if (t == 96) { // out_S is complete once its diagonal slabs have all published
                for (uint32_t spin = 0; PeekAcquire(a.Solved + task.Super) < task.Need; ++spin) {
                    if (GiveUp(spin, a.Fail)) break;
                    __nanosleep(40);
                }
            }
            ConsumerBarrier();
            stamp(id, 1);
            {   // operand in two planes, even and odd columns: the tile loads pair column 2 fk with 2 fk + 1
                const double2 *src = reinterpret_cast<const double2 *>(a.Out + (size_t(task.VecOffset) + t) * kWide);
                double2 *dst = reinterpret_cast<double2 *>(vec + (t & 1) * (64 * kWide) + (t >> 1) * kWide);
#pragma unroll
                for (int j = 0; j < kWide / 2; ++j) dst[j] = t < k ? __ldcg(src + j) : make_double2(0.0, 0.0);
            }
            ConsumerBarrier();
            for (uint32_t i = 0; i < task.Count; ++i, ++consumed, row += kSolveRows) {
                const uint32_t stage = consumed % kWideStages;
                MbarWait(sh.Full + stage, (consumed / kWideStages) & 1);
                const double2 *tile = reinterpret_cast<const double2 *>(sh.Ring[stage] + size_t(q) * tiles * 64) + lane;
                double acc[4][2]{};
#pragma unroll 2
                for (uint32_t j = 0; j < tiles; j += 2) { // two tiles per step, four accumulator chains
                    const bool two = j + 1 < tiles;
                    const double2 a0 = tile[32 * j], a1 = two ? tile[32 * j + 32] : make_double2(0.0, 0.0);
                    const double *b = vec + (4 * j + fk) * kWide + fr;
                    Dmma(acc[0][0], acc[0][1], a0.x, b[0]);
                    Dmma(acc[1][0], acc[1][1], a0.y, b[64 * kWide]);
                    Dmma(acc[2][0], acc[2][1], a1.x, b[4 * kWide]);
                    Dmma(acc[3][0], acc[3][1], a1.y, b[(64 + 4) * kWide]);
                }
                __syncwarp();
                if (lane == 0) MbarArrive(sh.Empty + stage);
                const uint32_t mine = node, at = row;
                if (i + 1 < task.Count) node = node_of(row + kSolveRows);
                if (at >= task.RowMin && at < row_end) {
                    double *dst = a.Acc + (size_t(3) * mine + at % 3) * kWide + 2 * fk;
                    atomicAdd(dst, -((acc[0][0] + acc[1][0]) + (acc[2][0] + acc[3][0])));
                    atomicAdd(dst + 1, -((acc[0][1] + acc[1][1]) + (acc[2][1] + acc[3][1])));
                }
            }
        } else {
            // A run of task.Count consecutive slabs: acc_S[k x 8] -= sum over the slabs of P_slab^T [k x 32] out[slab rows x 8].
            // Warp q owns output columns 32q .. 32q+31 of the supernode and keeps their sums in its DMMA accumulators across the
            // run: ONE set of FP64 atomics per run. The solved entries of a slab's rows are gathered by every warp for itself
            // (B fragment: row 4 ks + fk of the slab, right-hand side fr).
            const uint32_t tiles = (k + 7) >> 3, row_end = min(task.Limit, task.Row0 + task.Count * kSolveRows);
            auto entry_of = [&](uint32_t row) -> const double * {
                return row >= task.RowMin && row < row_end ? a.Out + (size_t(3) * __ldg(a.Rows + task.RowsBase + row / 3) + row % 3) * kWide + fr : nullptr;
            };
            const double *entry[8];
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) entry[ks] = entry_of(task.Row0 + 4 * ks + fk);
            if (t == 96) owed = a.Arrived + task.Super;
            if (q == 3) { // the ancestors owning the run's rows: polled by the lanes of one warp
                for (uint32_t i = lane; i < task.LinkCount; i += 32) {
                    const uint32_t *solved = a.Solved + a.Links[task.LinkBegin + i];
                    const uint32_t need = a.LinkNeed[task.LinkBegin + i];
                    for (uint32_t spin = 0; PeekAcquire(solved) < need; ++spin) {
                        if (GiveUp(spin, a.Fail)) break;
                        __nanosleep(40);
                    }
                }
            }
            ConsumerBarrier(); // the links are solved
            stamp(id, 1);
            double c[4][2]{};
            for (uint32_t i = 0; i < task.Count; ++i, ++consumed) {
                double b[8];
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) b[ks] = entry[ks] ? __ldcg(entry[ks]) : 0.0;
                if (i + 1 < task.Count) {
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) entry[ks] = entry_of(task.Row0 + (i + 1) * kSolveRows + 4 * ks + fk);
                }
                const uint32_t stage = consumed % kWideStages;
                MbarWait(sh.Full + stage, (consumed / kWideStages) & 1);
                const double *slab = sh.Ring[stage] + (4 * q) * 64 + fk * 8 + fr;
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
#pragma unroll
                    for (int mi = 0; mi < 4; ++mi)
                        if (4 * q + mi < tiles) Dmma(c[mi][0], c[mi][1], slab[((ks >> 1) * tiles + mi) * 64 + (ks & 1) * 32], b[ks]);
                __syncwarp();
                if (lane == 0) MbarArrive(sh.Empty + stage);
            }
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) {
                const uint32_t col = 32 * q + 8 * mi + fr;
                if (col < k) {
                    double *dst = a.Acc + (size_t(task.VecOffset) + col) * kWide + 2 * fk;
                    atomicAdd(dst, -c[mi][0]);
                    atomicAdd(dst + 1, -c[mi][1]);
                }
            }
        }
        ConsumerBarrier(); // every contribution of the task has been issued; the shared operand is free
        if (owed) ArriveRelease(owed);
        stamp(id, 2);
    }
}

// Panel of up to kWide right-hand sides, column-major n x width in natural DOF order -> acc[permuted DOF][kWide]
// (missing columns are zero).
__global__ void WideBeginKernel(const double *__restrict__ b, size_t n, uint32_t width, const uint32_t *__restrict__ inv_perm, double *__restrict__ acc) {
    const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; // natural DOF
    if (i >= n) return;
    const size_t dst = (size_t(3) * inv_perm[i / 3] + i % 3) * kWide;
#pragma unroll
    for (int w = 0; w < kWide; ++w) acc[dst + w] = uint32_t(w) < width ? b[i + size_t(w) * n] : 0.0;
}
__global__ void WidePermuteOutKernel(const double *__restrict__ work, size_t n, uint32_t width, const uint32_t *__restrict__ inv_perm, double *__restrict__ x) {
    const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t src = (size_t(3) * inv_perm[i / 3] + i % 3) * kWide;
    for (uint32_t w = 0; w < width; ++w) x[i + size_t(w) * n] = work[src + w];
}

// b (natural DOF order) -> acc under the fill-reducing permutation; `out` is marked unsolved.
__global__ void SweepBeginKernel(const double *__restrict__ b, const uint32_t *__restrict__ inv_perm, uint32_t n_nodes, double *__restrict__ acc, double *__restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 3 * n_nodes) {
        acc[3 * inv_perm[i / 3] + i % 3] = b[i];
        out[i] = __longlong_as_double((long long)kUnsolved);
    }
}
__global__ void MarkUnsolvedKernel(double *__restrict__ out, size_t n) {
    const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __longlong_as_double((long long)kUnsolved);
}

// ------------------------------------------------------------------------------------------------ FP64 rate probes
template<int Mode>
__global__ void Fp64RateKernel(double *out, int iters) {
    double a = 1.0 + threadIdx.x * 1e-9, b = 0.999999, c[8][2]{};
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if constexpr (Mode == 0) {
                c[u][0] = fma(a, b, c[u][0]);
                c[u][1] = fma(b, a, c[u][1]);
            } else {
                Dmma(c[u][0], c[u][1], a, b);
            }
        }
    }
    double sum = 0;
    for (int u = 0; u < 8; ++u) sum += c[u][0] + c[u][1];
    if (sum == 123.456) out[0] = sum;
}

double Seconds() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
inline uint32_t Blocks(uint64_t n, int threads) { return uint32_t((n + threads - 1) / threads); }
} // namespace

double MeasureFp64Rate(int mode, int iters) {
    double *out = nullptr;
    ME_CUDA(cudaMalloc(&out, 8));
    cudaEvent_t e0, e1;
    ME_CUDA(cudaEventCreate(&e0));
    ME_CUDA(cudaEventCreate(&e1));
    const int grid = 148 * 8, threads = 256, inner = 4096;
    auto launch = [&] {
        if (mode == 0) Fp64RateKernel<0><<<grid, threads>>>(out, inner);
        else Fp64RateKernel<1><<<grid, threads>>>(out, inner);
    };
    launch();
    ME_CUDA(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i) launch();
    ME_CUDA(cudaEventRecord(e1));
    ME_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    ME_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    // Per thread per inner iteration: mode 0 = 16 DFMA = 32 flop; mode 1 = 8 DMMA per warp = 8 * 512 flop per 32 threads = 128 flop per thread.
    const double flop_per_thread = double(inner) * (mode == 0 ? 32.0 : 128.0);
    return flop_per_thread * grid * threads * iters / (ms * 1e-3);
}

SparseCholesky::SparseCholesky(FemSystem &fem, const SymbolicOptions &opt) : Fem(fem) {
    const double t0 = Seconds();
    ME_CUDA(cudaSetDevice(fem.Device));
    std::vector<uint32_t> rowptr, col;
    std::vector<float> xyz;
    fem.CopyFullPattern(rowptr, col);
    fem.CopyNodeCoords(xyz);
    // The solve schedules keep being built on a host thread while the structures below are uploaded and (in the caller) the numeric
    // factorisation runs: UploadSchedules, at the first solve, waits for them.
    SymbolicOptions options = opt;
    if (const char *env = std::getenv("ME_MACRO_PANELS")) options.MacroPanels = uint32_t(std::max(1, std::atoi(env))); // (tuning aids)
    if (std::getenv("ME_MACRO_BACKWARD")) options.MacroBackward = true;
    AnalyseInto(Sym, fem.NodeCount, rowptr.data(), col.data(), xyz.data(), options, true);
    if (Sym.MaxPanelColumns > 128) Fail(ME_BAD_ARG, "internal: panel of %u columns", Sym.MaxPanelColumns);
    auto s = fem.Stream;
    DSuperFirst.Upload(Sym.SuperFirst, s);
    DRows.Upload(Sym.Rows, s);
    DNodeSuper.Upload(Sym.NodeSuper, s);
    DInvPerm.Upload(Sym.InvPerm, s);
    DPerm.Upload(Sym.Perm, s);
    DSegTarget.Upload(Sym.SegTarget, s);
    DSegBegin.Upload(Sym.SegBegin, s);
    DSegEnd.Upload(Sym.SegEnd, s);
    DLevelOrder.Upload(Sym.LevelOrder, s);
    DRowPtr.Upload(Sym.RowPtr, s);
    DPanelOffset.Upload(Sym.PanelOffset, s);
    DInvOffset.Upload(Sym.InvOffset, s);
    DSlabOffset.Upload(Sym.SlabOffset, s);
    DPanelTiles.Upload(Sym.PanelTiles, s);
    DUpdateTiles.Upload(Sym.UpdateTiles, s);
    DMacroOffset.Upload(Sym.MacroOffset, s);
    DMacroOffsetT.Upload(Sym.MacroOffsetT, s);
    DMacroJobs.Upload(Sym.MacroJobs, s);
    if (Sym.Rows.size() >= (uint64_t(1) << 32)) Fail(ME_BAD_ARG, "mesh too large: supernodal row lists exceed 32-bit indexing");
    L.Reserve(Sym.FactorNonZeros);
    Linv.Reserve(Sym.InvOffset[Sym.NumSuper]);
    LinvT.Reserve(Sym.InvOffset[Sym.NumSuper]);
    Slabs.Reserve(Sym.SlabOffset[Sym.NumSuper] + 1);
    ME_CUDA(cudaMemsetAsync(Slabs.Ptr, 0, Slabs.Capacity * sizeof(double), s)); // (the padding rows and columns are never written again)
    MacroW.Reserve(Sym.MacroOffset[Sym.NumSuper] + 1);
    MacroWT.Reserve(Sym.MacroOffsetT[Sym.NumSuper] + 1);
    Work.Reserve(fem.N);
    Work2.Reserve(fem.N);
    DFail.Reserve(1);
    DCounters.Reserve(size_t(4) * Sym.NumSuper + 2);
    {
        int device = 0, sms = 0, fwd = 0, bwd = 0;
        ME_CUDA(cudaGetDevice(&device));
        ME_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
        ME_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&fwd, SweepKernel<false>, kSweepThreads, 0));
        ME_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bwd, SweepKernel<true>, kSweepThreads, 0));
        if (fwd < 1 || bwd < 1) Fail(ME_CUDA_ERROR, "sweep kernels do not fit an SM");
        FwdGrid = uint32_t(sms * fwd), BwdGrid = uint32_t(sms * bwd); // every CTA resident: the spin-waits rely on it
        ME_CUDA(cudaFuncSetAttribute(WideSweepKernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sizeof(WideShared))));
        ME_CUDA(cudaFuncSetAttribute(WideSweepKernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sizeof(WideShared))));
        ME_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&fwd, WideSweepKernel<false>, kWideThreads, sizeof(WideShared)));
        ME_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bwd, WideSweepKernel<true>, kWideThreads, sizeof(WideShared)));
        if (fwd < 1 || bwd < 1) Fail(ME_CUDA_ERROR, "panel sweep kernels do not fit an SM");
        WideFwdGrid = uint32_t(sms * fwd), WideBwdGrid = uint32_t(sms * bwd);
    }
    for (auto &e : Ev) ME_CUDA(cudaEventCreate(&e));
    ME_CUDA(cudaFuncSetAttribute(FactorDiagKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (128 * 128 + 112 * 16 + 128) * 8));
    ME_CUDA(cudaFuncSetAttribute(PanelTrsmKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (128 * kLdA + kChunk * kLdB128) * 8));
    ME_CUDA(cudaStreamSynchronize(s));
    Stats.AnalyseSeconds = Seconds() - t0;
    Stats.FactorNonZeros = Sym.FactorNonZeros;
    Stats.FactorFlops = Sym.FactorFlops;
    Stats.Supernodes = Sym.NumSuper;
    Stats.Levels = Sym.NumLevels;
}

void SparseCholesky::UploadSchedules() {
    if (SchedulesUploaded) return;
    Sym.WaitSchedules();
    auto s = Fem.Stream;
    DFwdTasks.Upload(Sym.FwdTasks, s);
    DFwdLinks.Upload(Sym.FwdLinks, s);
    DBwdTasks.Upload(Sym.BwdTasks, s);
    DWideFwdTasks.Upload(Sym.WideFwdTasks, s);
    DWideBwdTasks.Upload(Sym.WideBwdTasks, s);
    DWideFwdLinks.Upload(Sym.WideFwdLinks, s);
    DWideBwdLinks.Upload(Sym.WideBwdLinks, s);
    DWideBwdLinkNeed.Upload(Sym.WideBwdLinkNeed, s);
    DWideFwdNeed.Upload(Sym.WideFwdNeed, s);
    DWideBwdNeed.Upload(Sym.WideBwdNeed, s);
    Stats.SweepLevels = Sym.SweepLevels;
    SchedulesUploaded = true;
}

SparseCholesky::~SparseCholesky() {
    cudaSetDevice(Fem.Device);
    for (auto &e : Ev)
        if (e) cudaEventDestroy(e);
}

void SparseCholesky::Factorize(double sigma) {
    ME_CUDA(cudaSetDevice(Fem.Device));
    auto s = Fem.Stream;
    FactorView v{DSuperFirst.Ptr, DRows.Ptr, DNodeSuper.Ptr, DRowPtr.Ptr, DPanelOffset.Ptr, DInvOffset.Ptr, DSegTarget.Ptr, DSegBegin.Ptr, DSegEnd.Ptr, L.Ptr, Linv.Ptr, LinvT.Ptr, Slabs.Ptr, DFail.Ptr, DSlabOffset.Ptr};
    ME_CUDA(cudaEventRecord(Ev[0], s));
    ME_CUDA(cudaMemsetAsync(DFail.Ptr, 0, sizeof(int), s));
    ME_CUDA(cudaMemsetAsync(L.Ptr, 0, Sym.FactorNonZeros * sizeof(double), s));
    ScatterMatrixKernel<<<Blocks(Fem.NumBlocks, 256), 256, 0, s>>>(v, Fem.BlkRow.Ptr, Fem.BlkCol.Ptr, Fem.KBlk.Ptr, Fem.MBlk.Ptr, DInvPerm.Ptr, Fem.NumBlocks, sigma);
    uint32_t launches = 1;
    const uint32_t kmax = Sym.MaxPanelColumns;
    const size_t diag_smem = (size_t((kmax + 1) & ~1u) * kmax + 112 * 16 + 128) * 8;
    const size_t trsm_smem = (128 * kLdA + kChunk * kLdB128) * 8;
    for (uint32_t l = 0; l < Sym.NumLevels; ++l) {
        const uint32_t n_super = Sym.LevelPtr[l + 1] - Sym.LevelPtr[l];
        FactorDiagKernel<<<n_super, kFactorThreads, diag_smem, s>>>(v, DLevelOrder.Ptr + Sym.LevelPtr[l]);
        ++launches;
        const uint64_t n_panel = Sym.PanelTilePtr[l + 1] - Sym.PanelTilePtr[l];
        if (n_panel) {
            PanelTrsmKernel<<<uint32_t(n_panel), kTrsmThreads, trsm_smem, s>>>(v, DPanelTiles.Ptr + Sym.PanelTilePtr[l]);
            ++launches;
        }
        const uint64_t n_update = Sym.UpdateTilePtr[l + 1] - Sym.UpdateTilePtr[l];
        if (n_update) {
            SyrkScatterKernel<<<uint32_t(n_update), kSyrkThreads, 0, s>>>(v, DUpdateTiles.Ptr + Sym.UpdateTilePtr[l]);
            ++launches;
        }
    }
    if (!Sym.MacroJobs.empty()) {
        const MacroView mv{DSuperFirst.Ptr, DRowPtr.Ptr, DPanelOffset.Ptr, DInvOffset.Ptr, DMacroOffset.Ptr, DMacroOffsetT.Ptr, L.Ptr, Linv.Ptr, MacroW.Ptr, MacroWT.Ptr};
        MacroInverseKernel<<<uint32_t(Sym.MacroJobs.size()), kMacroThreads, 0, s>>>(mv, DMacroJobs.Ptr);
        ++launches;
    }
    ME_CUDA(cudaEventRecord(Ev[1], s));
    int fail = 0;
    ME_CUDA(cudaMemcpyAsync(&fail, DFail.Ptr, sizeof(int), cudaMemcpyDeviceToHost, s));
    ME_CUDA(cudaStreamSynchronize(s));
    ME_CUDA(cudaGetLastError());
    ME_CUDA(cudaEventElapsedTime(&Stats.FactorMs, Ev[0], Ev[1]));
    Stats.KernelLaunches += launches;
    if (fail == 1) Fail(ME_FACTOR_FAILED, "Cholesky factorization of K - sigma*M failed: non-positive pivot (sigma = %g)", sigma);
    if (fail) Fail(ME_CUDA_ERROR, "internal: symbolic structure does not cover the matrix (code %d)", fail);
    Factored = true;
}

void SparseCholesky::Solve(const double *b, double *x, uint32_t width) {
    if (!Factored) Fail(ME_BAD_ARG, "Solve before Factorize");
    ME_CUDA(cudaSetDevice(Fem.Device));
    UploadSchedules();
    auto s = Fem.Stream;
    const uint32_t n = Fem.N, ns = Sym.NumSuper;
    FactorView v{DSuperFirst.Ptr, DRows.Ptr, DNodeSuper.Ptr, DRowPtr.Ptr, DPanelOffset.Ptr, DInvOffset.Ptr, DSegTarget.Ptr, DSegBegin.Ptr, DSegEnd.Ptr, L.Ptr, Linv.Ptr, LinvT.Ptr, Slabs.Ptr, DFail.Ptr, DSlabOffset.Ptr};
    if (width > 2 && Work.Capacity < size_t(n) * kWide) {
        ME_CUDA(cudaStreamSynchronize(s)); // growing a buffer releases the old block
        Work.Reserve(size_t(n) * kWide);
        Work2.Reserve(size_t(n) * kWide);
    }
    // [0, ns) forward arrivals, [ns, 2 ns) backward arrivals, the two ticket counters, then (panel sweeps) the published
    // diagonal slabs: [2 ns + 2, 3 ns + 2) forward, [3 ns + 2, 4 ns + 2) backward.
    uint32_t *counters = DCounters.Ptr;
    // Forward: Work accumulates the right-hand side, Work2 receives y. Backward: Work2 accumulates, Work receives x.
    const SweepArgs fwd{DFwdTasks.Ptr, uint32_t(Sym.FwdTasks.size()), DFwdLinks.Ptr, nullptr, counters + 2 * size_t(ns), counters, counters + 2 * size_t(ns) + 2, DRows.Ptr, Linv.Ptr, L.Ptr, Work.Ptr, Work2.Ptr, DFail.Ptr};
    const SweepArgs bwd{DBwdTasks.Ptr, uint32_t(Sym.BwdTasks.size()), nullptr, nullptr, counters + 2 * size_t(ns) + 1, counters + ns, counters + 3 * size_t(ns) + 2, DRows.Ptr, LinvT.Ptr, Slabs.Ptr, Work2.Ptr, Work.Ptr, DFail.Ptr};
    SweepArgs wide_fwd = fwd, wide_bwd = bwd;
    wide_fwd.Tasks = DWideFwdTasks.Ptr, wide_fwd.NumTasks = uint32_t(Sym.WideFwdTasks.size()), wide_fwd.Links = DWideFwdLinks.Ptr, wide_fwd.Panel = Slabs.Ptr;
    wide_bwd.Tasks = DWideBwdTasks.Ptr, wide_bwd.NumTasks = uint32_t(Sym.WideBwdTasks.size()), wide_bwd.Links = DWideBwdLinks.Ptr, wide_bwd.LinkNeed = DWideBwdLinkNeed.Ptr;
    wide_fwd.Macro = MacroW.Ptr, wide_fwd.ArriveNeed = DWideFwdNeed.Ptr;
    wide_bwd.Macro = MacroWT.Ptr, wide_bwd.ArriveNeed = DWideBwdNeed.Ptr;
    auto single = [&](const double *bi, double *xi) {
        SweepBeginKernel<<<Blocks(n, 256), 256, 0, s>>>(bi, DInvPerm.Ptr, Fem.NodeCount, Work.Ptr, Work2.Ptr);
        ME_CUDA(cudaMemsetAsync(counters, 0, (size_t(4) * ns + 2) * sizeof(uint32_t), s));
        SweepKernel<false><<<FwdGrid, kSweepThreads, 0, s>>>(fwd);
        MarkUnsolvedKernel<<<Blocks(n, 256), 256, 0, s>>>(Work.Ptr, n);
        SweepKernel<true><<<BwdGrid, kSweepThreads, 0, s>>>(bwd);
        PermuteOutKernel<<<Blocks(n, 256), 256, 0, s>>>(Work.Ptr, DInvPerm.Ptr, Fem.NodeCount, xi);
        Stats.KernelLaunches += 5;
    };
    // Panels go through the factor kWide columns per pass (WideSweepKernel); a remainder of one or two columns is
    // cheaper as single sweeps than as a mostly empty panel.
    uint32_t rhs = 0;
    while (rhs < width) {
        const uint32_t left = width - rhs;
        if (left <= 2) {
            single(b + size_t(rhs) * n, x + size_t(rhs) * n);
            ++rhs;
            continue;
        }
        const uint32_t w = std::min<uint32_t>(left, kWide);
        static bool trace_done = false; // one timeline per process: the first panel application of the first factor
        const char *trace_path = trace_done ? nullptr : std::getenv("ME_SWEEP_TRACE");
        DeviceBuffer<unsigned long long> trace;
        if (trace_path) {
            trace.Reserve(size_t(3) * (wide_fwd.NumTasks + wide_bwd.NumTasks));
            ME_CUDA(cudaMemsetAsync(trace.Ptr, 0, trace.Capacity * sizeof(unsigned long long), s));
            wide_fwd.Trace = trace.Ptr, wide_bwd.Trace = trace.Ptr + size_t(3) * wide_fwd.NumTasks;
        }
        WideBeginKernel<<<Blocks(n, 256), 256, 0, s>>>(b + size_t(rhs) * n, n, w, DInvPerm.Ptr, Work.Ptr);
        ME_CUDA(cudaMemsetAsync(counters, 0, (size_t(4) * ns + 2) * sizeof(uint32_t), s));
        WideSweepKernel<false><<<WideFwdGrid, kWideThreads, sizeof(WideShared), s>>>(wide_fwd);
        WideSweepKernel<true><<<WideBwdGrid, kWideThreads, sizeof(WideShared), s>>>(wide_bwd);
        WidePermuteOutKernel<<<Blocks(n, 256), 256, 0, s>>>(Work.Ptr, n, w, DInvPerm.Ptr, x + size_t(rhs) * n);
        if (trace_path) { // one panel application's timeline: task records followed by their three time stamps
            std::vector<unsigned long long> stamps(size_t(3) * (wide_fwd.NumTasks + wide_bwd.NumTasks));
            ME_CUDA(cudaMemcpyAsync(stamps.data(), trace.Ptr, stamps.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
            ME_CUDA(cudaStreamSynchronize(s));
            if (FILE *f = std::fopen(trace_path, "wb")) {
                const uint32_t header[4] = {wide_fwd.NumTasks, wide_bwd.NumTasks, ns, uint32_t(sizeof(SweepTask))};
                std::fwrite(header, sizeof(header), 1, f);
                std::fwrite(Sym.WideFwdTasks.data(), sizeof(SweepTask), Sym.WideFwdTasks.size(), f);
                std::fwrite(Sym.WideBwdTasks.data(), sizeof(SweepTask), Sym.WideBwdTasks.size(), f);
                std::fwrite(Sym.Level.data(), sizeof(uint32_t), ns, f);
                std::fwrite(Sym.MacroFirst.data(), sizeof(uint32_t), ns, f);
                std::fwrite(stamps.data(), sizeof(unsigned long long), stamps.size(), f);
                std::fclose(f);
            }
            wide_fwd.Trace = wide_bwd.Trace = nullptr;
            trace_done = true;
        }
        Stats.KernelLaunches += 4;
        rhs += w;
    }
    SolvesSinceCheck += width;
}

void SparseCholesky::CheckSolves() {
    if (!SolvesSinceCheck) return;
    int fail = 0;
    ME_CUDA(cudaMemcpyAsync(&fail, DFail.Ptr, sizeof(int), cudaMemcpyDeviceToHost, Fem.Stream));
    ME_CUDA(cudaStreamSynchronize(Fem.Stream));
    SolvesSinceCheck = 0;
    if (fail) Fail(ME_CUDA_ERROR, "internal: a triangular-solve sweep stalled (code %d)", fail);
}

} // namespace me
