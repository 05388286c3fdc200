// sm_100a kernels of the FEM assembly stage. See fem.h for the data layout.
//
// The stage is HBM/latency bound integer + FP64 work, so the design goal is "touch every byte once, coalesced":
//   * numbering and pattern are built with radix sorts of packed keys (CUB DeviceRadixSort is the plumbing; the keys,
//     the first-seen ranking that reproduces the reference's unordered_map numbering, and the head/scan logic are ours);
//   * the numeric assembly is a GATHER: the stable sort that yields the pattern also yields, per stored 3x3 block, the
//     list of (element, local pair) contributions in element order. One thread owns one block, evaluates its
//     contributions on the fly from the per-element gradients (13 doubles per element, L1/L2 resident) and writes the
//     block once. No atomics, no per-element staging buffer, and the summation order is the reference's triplet
//     insertion order (mesh2modes.cpp:295-320 + setFromTriplets), so values are deterministic run to run.
#include "fem.h"

#include <cub/cub.cuh>

#include <algorithm>
#include <array>
#include <cmath>
#include <map>

namespace me {

// ------------------------------------------------------------------------------------------------ element tables (host)
namespace {
// Polynomials in the four barycentric coordinates: exponent tuple -> coefficient.
using Poly = std::map<std::array<int, 4>, double>;

Poly Mul(const Poly &a, const Poly &b) {
    Poly out;
    for (const auto &[ea, ca] : a)
        for (const auto &[eb, cb] : b) out[{ea[0] + eb[0], ea[1] + eb[1], ea[2] + eb[2], ea[3] + eb[3]}] += ca * cb;
    return out;
}
// int l^e dV / V over a straight tet = 6 prod(e!) / (sum(e) + 3)!   (mesh2modes.cpp:188-196)
double UnitIntegral(const Poly &p) {
    static constexpr double F[]{1, 1, 2, 6, 24, 120, 720, 5040};
    double sum = 0;
    for (const auto &[e, c] : p) sum += c * 6 * F[e[0]] * F[e[1]] * F[e[2]] * F[e[3]] / F[e[0] + e[1] + e[2] + e[3] + 3];
    return sum;
}
std::array<int, 4> Unit(int i) { return {i == 0, i == 1, i == 2, i == 3}; }
constexpr int kEdge[6][2]{{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
} // namespace

ElementTables MakeElementTables(uint32_t order) {
    ElementTables t;
    t.Npe = order == 2 ? 10 : 4;
    std::vector<Poly> n(t.Npe);
    std::vector<std::array<Poly, 4>> dn(t.Npe);
    if (order == 2) {
        // corners N_i = l_i (2 l_i - 1), edges N_ij = 4 l_i l_j  (mesh2modes.cpp:203-222)
        for (int i = 0; i < 4; ++i) {
            auto u = Unit(i);
            n[i][{2 * u[0], 2 * u[1], 2 * u[2], 2 * u[3]}] = 2;
            n[i][u] = -1;
            dn[i][i][u] = 4;
            dn[i][i][{0, 0, 0, 0}] = -1;
        }
        for (int e = 0; e < 6; ++e) {
            const int i = kEdge[e][0], j = kEdge[e][1];
            auto ui = Unit(i), uj = Unit(j);
            n[4 + e][{ui[0] + uj[0], ui[1] + uj[1], ui[2] + uj[2], ui[3] + uj[3]}] = 4;
            dn[4 + e][i][uj] = 4;
            dn[4 + e][j][ui] = 4;
        }
    } else {
        for (int i = 0; i < 4; ++i) {
            n[i][Unit(i)] = 1;
            dn[i][i][{0, 0, 0, 0}] = 1;
        }
    }
    t.Mass.assign(size_t(t.Npe) * t.Npe, 0.0);
    t.Grad.assign(size_t(t.Npe) * 4 * t.Npe * 4, 0.0);
    for (uint32_t a = 0; a < t.Npe; ++a)
        for (uint32_t c = 0; c < t.Npe; ++c) {
            t.Mass[a * t.Npe + c] = UnitIntegral(Mul(n[a], n[c]));
            for (int k = 0; k < 4; ++k)
                for (int l = 0; l < 4; ++l)
                    if (!dn[a][k].empty() && !dn[c][l].empty()) t.Grad[((a * 4 + k) * t.Npe + c) * 4 + l] = UnitIntegral(Mul(dn[a][k], dn[c][l]));
        }
    return t;
}

// ------------------------------------------------------------------------------------------------ kernels
namespace {
constexpr int kThreads = 256;
inline uint32_t Blocks(uint64_t n, int threads = kThreads) { return uint32_t((n + threads - 1) / threads); }

struct D3 {
    double x, y, z;
};
__device__ inline D3 LoadPoint(const double *__restrict__ pts, uint32_t i) { return {pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]}; }
__device__ inline D3 Sub(D3 a, D3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ inline D3 Cross(D3 a, D3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
__device__ inline double Dot(D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// FilterDegenerate (mesh2modes.cpp:42-60).
__global__ void FilterKernel(const double *__restrict__ pts, const uint4 *__restrict__ tets, uint32_t n, uint8_t *__restrict__ keep) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const uint4 id = tets[t];
    const D3 p[4]{LoadPoint(pts, id.x), LoadPoint(pts, id.y), LoadPoint(pts, id.z), LoadPoint(pts, id.w)};
    const double det = fabs(Dot(Sub(p[1], p[0]), Cross(Sub(p[2], p[0]), Sub(p[3], p[0]))));
    double lmax_sq = 0;
    for (int i = 0; i < 4; ++i)
        for (int j = i + 1; j < 4; ++j) {
            const D3 d = Sub(p[i], p[j]);
            lmax_sq = fmax(lmax_sq, Dot(d, d));
        }
    keep[t] = det > 1e-12 * lmax_sq * sqrt(lmax_sq);
}

// ComputeElementBases (mesh2modes.cpp:137-165), same cofactor formula. basis is SoA: [13][n] = volume, Phig[4][3].
__global__ void BasisKernel(const double *__restrict__ pts, const uint4 *__restrict__ tets, uint32_t n, double *__restrict__ basis) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const uint4 id = tets[t];
    const double v[4][3]{{pts[3 * id.x], pts[3 * id.x + 1], pts[3 * id.x + 2]}, {pts[3 * id.y], pts[3 * id.y + 1], pts[3 * id.y + 2]},
                         {pts[3 * id.z], pts[3 * id.z + 1], pts[3 * id.z + 2]}, {pts[3 * id.w], pts[3 * id.w + 1], pts[3 * id.w + 2]}};
    const D3 a{v[0][0], v[0][1], v[0][2]}, b{v[1][0], v[1][1], v[1][2]}, c{v[2][0], v[2][1], v[2][2]}, d{v[3][0], v[3][1], v[3][2]};
    const double det = Dot(Sub(d, a), Cross(Sub(b, a), Sub(c, a)));
    basis[t] = fabs(det / 6);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            D3 col[2];
            int ni = 0;
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) {
                if (ii == i) continue;
                int nj = 0;
#pragma unroll
                for (int jj = 0; jj < 3; ++jj) {
                    if (jj == j) continue;
                    (ni == 0 ? col[nj].x : ni == 1 ? col[nj].y : col[nj].z) = v[ii][jj];
                    ++nj;
                }
                ++ni;
            }
            const D3 cr = Cross(col[0], col[1]);
            const double sign = (i + j) % 2 == 0 ? -1.0 : 1.0;
            basis[size_t(1 + 3 * i + j) * n + t] = sign * (cr.x + cr.y + cr.z) / det;
        }
    }
}

__constant__ uint8_t cEdge[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};

// Edge keys of BuildQuadMesh (mesh2modes.cpp:252-258), packed as min * V + max; value = element * 6 + edge slot.
__global__ void EdgeKeyKernel(const uint32_t *__restrict__ tets, uint32_t n_tets, uint32_t n_points, uint64_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_tets * 6) return;
    const uint32_t e = i / 6, s = i - 6 * e;
    const uint32_t a = tets[4 * e + cEdge[s][0]], b = tets[4 * e + cEdge[s][1]];
    keys[i] = uint64_t(min(a, b)) * n_points + max(a, b);
    vals[i] = i;
}

__global__ void HeadKernel(const uint64_t *__restrict__ keys, uint32_t n, uint32_t *__restrict__ heads) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) heads[i] = i == 0 || keys[i] != keys[i - 1];
}

// first[u] = smallest (element, slot) index of unique edge u: the stable sort keeps a run in ascending index order.
__global__ void FirstSeenKernel(const uint32_t *__restrict__ heads, const uint32_t *__restrict__ uid1, const uint32_t *__restrict__ vals, uint32_t n,
                                uint32_t *__restrict__ first, uint32_t *__restrict__ iota) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && heads[i]) {
        first[uid1[i] - 1] = vals[i];
        iota[uid1[i] - 1] = uid1[i] - 1;
    }
}
__global__ void RankKernel(const uint32_t *__restrict__ sorted_u, uint32_t n, uint32_t *__restrict__ rank) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) rank[sorted_u[j]] = j;
}
__global__ void MidsideKernel(const uint32_t *__restrict__ uid1, const uint32_t *__restrict__ vals, const uint32_t *__restrict__ rank, uint32_t n, uint32_t n_points,
                              uint32_t *__restrict__ elem_nodes) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t v = vals[i], e = v / 6, s = v - 6 * e;
    elem_nodes[10 * e + 4 + s] = n_points + rank[uid1[i] - 1];
}
__global__ void CornerKernel(const uint32_t *__restrict__ tets, uint32_t n_tets, uint32_t npe, uint32_t *__restrict__ elem_nodes) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_tets * 4) elem_nodes[npe * (i / 4) + (i & 3)] = tets[i];
}
__global__ void NodeCoordKernel(const double *__restrict__ pts, uint32_t n_points, const uint32_t *__restrict__ elem_nodes, uint32_t n_tets, uint32_t order, float *__restrict__ xyz) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_points * 3) xyz[i] = float(pts[i]);
    if (order == 2 && i < n_tets * 6) {
        const uint32_t e = i / 6, s = i - 6 * e;
        const uint32_t a = elem_nodes[10 * e + cEdge[s][0]], b = elem_nodes[10 * e + cEdge[s][1]], m = elem_nodes[10 * e + 4 + s];
        for (int k = 0; k < 3; ++k) xyz[3 * m + k] = float(0.5 * (pts[3 * a + k] + pts[3 * b + k]));
    }
}

// One key per (element, unordered local node pair): the lower-triangular block (max node, min node) it lands in
// (`if (row < col) continue`, mesh2modes.cpp:297). value = ((element * pairs + pair) << 1) | swapped.
__global__ void PairKeyKernel(const uint32_t *__restrict__ elem_nodes, uint32_t n_tets, uint32_t npe, uint32_t n_pairs, const uint8_t *__restrict__ pair_a,
                              const uint8_t *__restrict__ pair_c, uint32_t node_count, uint64_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_tets * n_pairs) return;
    const uint32_t e = i / n_pairs, p = i - e * n_pairs;
    uint32_t r = elem_nodes[npe * e + pair_a[p]], c = elem_nodes[npe * e + pair_c[p]];
    const uint32_t swapped = r < c;
    if (swapped) {
        const uint32_t t = r;
        r = c;
        c = t;
    }
    keys[i] = uint64_t(c) * node_count + r;
    vals[i] = (i << 1) | swapped;
}
__global__ void BlockKernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ heads, const uint32_t *__restrict__ uid1, uint32_t n, uint32_t node_count,
                            uint32_t *__restrict__ blk_row, uint32_t *__restrict__ blk_col, uint32_t *__restrict__ contrib_ptr, uint32_t *__restrict__ blk_col_ptr) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !heads[i]) return;
    const uint32_t u = uid1[i] - 1;
    const uint32_t c = uint32_t(keys[i] / node_count), r = uint32_t(keys[i] - uint64_t(c) * node_count);
    blk_row[u] = r;
    blk_col[u] = c;
    contrib_ptr[u] = i;
    if (r == c) blk_col_ptr[c] = u; // every node carries its diagonal block, and it sorts first in its column
}

struct AssembleArgs {
    const uint32_t *ContribPtr, *Contrib;
    const uint32_t *Order; // blocks sorted by their number of contributions: a warp's 32 blocks then run the same number of iterations
    const double *Basis;   // [13][n_tets]
    const double *TabMass; // [npe][npe]
    const double *TermW;   // [npe*npe][4]
    const uint8_t *TermKL, *TermCount, *PairA, *PairC;
    double *KBlk, *MBlk;
    uint32_t NumBlocks, NumTets, Npe, NumPairs;
    double Lambda, Mu, Density;
};

// One thread per stored block: K(r,c) and M(r,c) as the in-order sum of their element contributions
// (AssembleQuadratic, mesh2modes.cpp:286-320).
// (An ncu capture of the first form, blocks in pattern order: issue slots busy 59 % of the cycles with 10 of 32 lanes active per
// instruction - every warp holds a few diagonal blocks with 24 contributions among off-diagonal ones with 4 to 6. Hence the order.)
__global__ void __launch_bounds__(kThreads) AssembleKernel(AssembleArgs a) {
    const uint32_t thread = blockIdx.x * blockDim.x + threadIdx.x;
    if (thread >= a.NumBlocks) return;
    const uint32_t u = a.Order[thread];
    double k[3][3]{}, m = 0;
    const uint32_t end = a.ContribPtr[u + 1];
    for (uint32_t i = a.ContribPtr[u]; i < end; ++i) {
        const uint32_t v = a.Contrib[i];
        const uint32_t ep = v >> 1, e = ep / a.NumPairs, p = ep - e * a.NumPairs;
        uint32_t la = a.PairA[p], lc = a.PairC[p];
        if (v & 1) {
            const uint32_t t = la;
            la = lc;
            lc = t;
        }
        const double vol = a.Basis[e];
        const uint32_t pair = la * a.Npe + lc;
        m += a.Density * vol * a.TabMass[pair];
        double g[3][3]{};
        const uint32_t terms = a.TermCount[pair];
        for (uint32_t t = 0; t < terms; ++t) {
            const double w = a.TermW[4 * pair + t];
            const uint32_t kl = a.TermKL[4 * pair + t], kk = kl >> 2, ll = kl & 3;
            double pk[3], pl[3];
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                pk[d] = a.Basis[size_t(1 + 3 * kk + d) * a.NumTets + e];
                pl[d] = a.Basis[size_t(1 + 3 * ll + d) * a.NumTets + e];
            }
#pragma unroll
            for (int p3 = 0; p3 < 3; ++p3)
#pragma unroll
                for (int q = 0; q < 3; ++q) g[p3][q] += w * (pk[p3] * pl[q]);
        }
        const double trace = g[0][0] + g[1][1] + g[2][2];
#pragma unroll
        for (int p3 = 0; p3 < 3; ++p3)
#pragma unroll
            for (int q = 0; q < 3; ++q) k[p3][q] += vol * (a.Lambda * g[p3][q] + a.Mu * g[q][p3] + (p3 == q ? a.Mu * trace : 0.0));
    }
    double *out = a.KBlk + size_t(9) * u;
#pragma unroll
    for (int p3 = 0; p3 < 3; ++p3)
#pragma unroll
        for (int q = 0; q < 3; ++q) out[3 * p3 + q] = k[p3][q];
    a.MBlk[u] = m;
}

// Sort keys of the assembly order: a block's number of contributions (clamped to 16 bits), largest first.
__global__ void ContribCountKernel(const uint32_t *__restrict__ contrib_ptr, uint32_t n_blocks, uint32_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_blocks) return;
    keys[u] = 65535u - min(contrib_ptr[u + 1] - contrib_ptr[u], 65535u);
    vals[u] = u;
}

// Diagonal node blocks keep only q <= p in the reference (mesh2modes.cpp:315); mirror that triangle so the block is
// exactly symmetric for the mat-vec.
__global__ void SymmetriseDiagKernel(const uint32_t *__restrict__ blk_col_ptr, uint32_t node_count, double *__restrict__ kblk) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= node_count) return;
    double *b = kblk + size_t(9) * blk_col_ptr[c];
    b[1] = b[3];
    b[2] = b[6];
    b[5] = b[7];
}

__global__ void RowKeyKernel(const uint32_t *__restrict__ blk_row, uint32_t n, uint32_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u < n) {
        keys[u] = blk_row[u];
        vals[u] = u;
    }
}
__global__ void LowPtrKernel(const uint32_t *__restrict__ keys, uint32_t n, uint32_t *__restrict__ low_ptr) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && (i == 0 || keys[i] != keys[i - 1])) low_ptr[keys[i]] = i;
}
__global__ void FullRowPtrKernel(const uint32_t *__restrict__ low_ptr, const uint32_t *__restrict__ blk_col_ptr, uint32_t node_count, uint32_t *__restrict__ full_row_ptr) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r <= node_count) full_row_ptr[r] = low_ptr[r] + blk_col_ptr[r] - r;
}
// Full row r = blocks (r, c <= r) in ascending c (from the row-sorted list, the diagonal last), then the rest of CSC
// column r, i.e. blocks (r' > r, r) read transposed.
__global__ void FillFullKernel(const uint32_t *__restrict__ sorted_blk, const uint32_t *__restrict__ low_ptr, const uint32_t *__restrict__ blk_row, const uint32_t *__restrict__ blk_col,
                               const uint32_t *__restrict__ blk_col_ptr, const uint32_t *__restrict__ full_row_ptr, uint32_t n_blocks, uint32_t *__restrict__ full_col,
                               uint32_t *__restrict__ full_src) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_blocks) return;
    {
        const uint32_t u = sorted_blk[i], r = blk_row[u];
        const uint32_t dest = full_row_ptr[r] + (i - low_ptr[r]);
        full_col[dest] = blk_col[u];
        full_src[dest] = u << 1;
    }
    {
        const uint32_t u = i, c = blk_col[u];
        if (u != blk_col_ptr[c]) {
            const uint32_t low_count = low_ptr[c + 1] - low_ptr[c];
            const uint32_t dest = full_row_ptr[c] + low_count + (u - blk_col_ptr[c] - 1);
            full_col[dest] = blk_row[u];
            full_src[dest] = (u << 1) | 1;
        }
    }
}
__global__ void GatherFullKernel(const uint32_t *__restrict__ full_src, const double *__restrict__ kblk, const double *__restrict__ mblk, uint32_t n_full, double *__restrict__ kfull,
                                 double *__restrict__ mfull) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_full * 9) return;
    const uint32_t d = i / 9, e = i - 9 * d, p = e / 3, q = e - 3 * p;
    const uint32_t s = full_src[d], u = s >> 1;
    kfull[i] = kblk[size_t(9) * u + ((s & 1) ? 3 * q + p : 3 * p + q)];
    if (e == 0) mfull[d] = mblk[u];
}

// y = K x, K in full block-CSR with 3x3 blocks. A CTA owns kSpmvRows consecutive block rows and STREAMS their blocks:
// the value array of those rows is one contiguous range, so 256 threads read it scalar by scalar, fully coalesced and
// with every load of a pass issued before the first is used (a row has only ~130 scalars for P1: a warp per row would
// sit on three dependent round trips — row pointer, value + column, x — with a single load per lane in flight).
// Products go to shared memory; 3 * kSpmvRows threads then sum their own output component.
constexpr int kSpmvRows = 32;   // block rows per CTA
constexpr int kSpmvCap = 512;   // blocks staged per pass: 9 * 512 products = 36 KB of shared memory
constexpr int kSpmvUnroll = 6;
__global__ void __launch_bounds__(kThreads) SpmvBsr3Kernel(const uint32_t *__restrict__ row_ptr, const uint32_t *__restrict__ col, const double *__restrict__ val,
                                                           const double *__restrict__ x, double *__restrict__ y, uint32_t n_rows) {
    __shared__ double prod[9 * kSpmvCap];
    __shared__ uint32_t rp[kSpmvRows + 1];
    const uint32_t t = threadIdx.x, row0 = blockIdx.x * kSpmvRows, rows = min(uint32_t(kSpmvRows), n_rows - row0);
    if (t <= rows) rp[t] = row_ptr[row0 + t];
    __syncthreads();
    const uint32_t b_end = rp[rows];
    const uint32_t my_row = t / 3, my_p = t - 3 * my_row;
    double acc = 0;
    for (uint32_t base = rp[0]; base < b_end; base += kSpmvCap) {
        const uint32_t nb = min(uint32_t(kSpmvCap), b_end - base), ns = 9 * nb;
        const double *v = val + size_t(9) * base;
        const uint32_t *cb = col + base;
        for (uint32_t s0 = t; s0 < ns; s0 += kSpmvUnroll * kThreads) {
            double vv[kSpmvUnroll], xx[kSpmvUnroll];
            uint32_t cc[kSpmvUnroll];
#pragma unroll
            for (int u = 0; u < kSpmvUnroll; ++u) {
                const uint32_t s = s0 + u * kThreads;
                vv[u] = s < ns ? __ldcs(v + s) : 0.0;
                cc[u] = s < ns ? __ldg(cb + s / 9) : 0u;
            }
#pragma unroll
            for (int u = 0; u < kSpmvUnroll; ++u) {
                const uint32_t s = s0 + u * kThreads;
                xx[u] = s < ns ? __ldg(x + 3 * cc[u] + (s % 9) % 3) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < kSpmvUnroll; ++u) {
                const uint32_t s = s0 + u * kThreads;
                if (s < ns) prod[s] = vv[u] * xx[u];
            }
        }
        __syncthreads();
        if (my_row < rows) {
            const uint32_t lo = max(rp[my_row], base) - base, hi = min(rp[my_row + 1], base + nb);
            for (uint32_t b = lo; base + b < hi; ++b) {
                const double *pb = prod + 9 * b + 3 * my_p;
                acc += (pb[0] + pb[1]) + pb[2];
            }
        }
        __syncthreads();
    }
    if (my_row < rows) y[size_t(3) * row0 + t] = acc;
}

// y = M x, M = (node mass matrix) (x) I3 in full node CSR: 12 bytes per stored scalar serve three output components.
// Eight lanes per node row.
__global__ void __launch_bounds__(kThreads) SpmvMassKernel(const uint32_t *__restrict__ row_ptr, const uint32_t *__restrict__ col, const double *__restrict__ val,
                                                           const double *__restrict__ x, double *__restrict__ y, uint32_t n_rows) {
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x, row = gid >> 3, sub = gid & 7;
    const bool live = row < n_rows;
    double a0 = 0, a1 = 0, a2 = 0;
    if (live) {
        const uint32_t end = row_ptr[row + 1];
        for (uint32_t j = row_ptr[row] + sub; j < end; j += 8) {
            const double m = val[j];
            const double *xc = x + size_t(3) * col[j];
            a0 += m * xc[0];
            a1 += m * xc[1];
            a2 += m * xc[2];
        }
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, o);
        a1 += __shfl_xor_sync(0xffffffffu, a1, o);
        a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    }
    if (live && sub < 3) y[size_t(3) * row + sub] = sub == 0 ? a0 : sub == 1 ? a1 : a2;
}

// Y = M X for a column-major panel of W right-hand sides: the matrix (12 bytes per stored scalar) is read once for all
// W columns instead of once per column.
template<int W>
__global__ void __launch_bounds__(kThreads) SpmvMassPanelKernel(const uint32_t *__restrict__ row_ptr, const uint32_t *__restrict__ col, const double *__restrict__ val, const double *__restrict__ x,
                                                                double *__restrict__ y, uint32_t n_rows, size_t ld) {
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x, row = gid >> 3, sub = gid & 7;
    const bool live = row < n_rows;
    double a[W][3];
#pragma unroll
    for (int w = 0; w < W; ++w) a[w][0] = a[w][1] = a[w][2] = 0;
    if (live) {
        const uint32_t end = row_ptr[row + 1];
        for (uint32_t j = row_ptr[row] + sub; j < end; j += 8) {
            const double m = val[j];
            const double *xc = x + size_t(3) * col[j];
#pragma unroll
            for (int w = 0; w < W; ++w) {
                a[w][0] += m * xc[w * ld];
                a[w][1] += m * xc[w * ld + 1];
                a[w][2] += m * xc[w * ld + 2];
            }
        }
    }
#pragma unroll
    for (int w = 0; w < W; ++w)
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            a[w][0] += __shfl_xor_sync(0xffffffffu, a[w][0], o);
            a[w][1] += __shfl_xor_sync(0xffffffffu, a[w][1], o);
            a[w][2] += __shfl_xor_sync(0xffffffffu, a[w][2], o);
        }
    // Lane `sub` writes column `sub` (three consecutive doubles): with W = 8 every lane of the row group stores.
    if (live) {
#pragma unroll
        for (int w = 0; w < W; ++w)
            if (int(sub) == w) {
                double *out = y + w * ld + size_t(3) * row;
                out[0] = a[w][0], out[1] = a[w][1], out[2] = a[w][2];
            }
    }
}

// Scalar lower-triangular CSC exactly as Eigen lays it out after setFromTriplets (mesh2modes.cpp:322-325).
__global__ void ExportCscKernel(int which, const uint32_t *__restrict__ blk_row, const uint32_t *__restrict__ blk_col, const uint32_t *__restrict__ blk_col_ptr, const double *__restrict__ kblk,
                                const double *__restrict__ mblk, uint32_t n_blocks, uint64_t *__restrict__ colptr, uint32_t *__restrict__ rowidx, double *__restrict__ values) {
    const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_blocks) return;
    const uint32_t c = blk_col[u], r = blk_row[u], first = blk_col_ptr[c], nb = blk_col_ptr[c + 1] - first, j = u - first;
    if (which == 0) {
        uint64_t start[3];
        start[0] = uint64_t(9) * first - uint64_t(3) * c;
        start[1] = start[0] + 3 * nb;
        start[2] = start[1] + 3 * nb - 1;
        for (uint32_t q = 0; q < 3; ++q) {
            if (j == 0) {
                colptr[3 * c + q] = start[q];
                for (uint32_t p = q; p < 3; ++p) {
                    rowidx[start[q] + (p - q)] = 3 * r + p;
                    values[start[q] + (p - q)] = kblk[size_t(9) * u + 3 * p + q];
                }
            } else {
                for (uint32_t p = 0; p < 3; ++p) {
                    const uint64_t pos = start[q] + (3 - q) + 3 * (j - 1) + p;
                    rowidx[pos] = 3 * r + p;
                    values[pos] = kblk[size_t(9) * u + 3 * p + q];
                }
            }
        }
    } else {
        for (uint32_t k = 0; k < 3; ++k) {
            const uint64_t start = uint64_t(3) * first + uint64_t(k) * nb;
            if (j == 0) colptr[3 * c + k] = start;
            rowidx[start + j] = 3 * r + k;
            values[start + j] = mblk[u];
        }
    }
}

// --- deterministic first-fit colouring -----------------------------------------------------------------------
__global__ void IncidenceKeyKernel(const uint32_t *__restrict__ elem_nodes, uint32_t n, uint32_t npe, uint32_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        keys[i] = elem_nodes[i];
        vals[i] = i / npe;
    }
}
__global__ void NodePtrKernel(const uint32_t *__restrict__ keys, uint32_t n, uint32_t node_count, uint32_t *__restrict__ node_ptr) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (i == 0 || keys[i] != keys[i - 1]) node_ptr[keys[i]] = i;
    if (i == n - 1) node_ptr[node_count] = n;
}
constexpr int kColourWords = 8; // up to 256 colours
// An element takes its colour once every earlier element around each of its nodes has one: the coloured set at a node
// is always a prefix of its (ascending) incident-element list, so "the next uncoloured element at every node is me"
// is the readiness test, and two ready elements never share a node. This reproduces the sequential first-fit exactly.
__global__ void ColourRoundKernel(const uint32_t *__restrict__ elem_nodes, uint32_t n_tets, uint32_t npe, const uint32_t *__restrict__ node_ptr, const uint32_t *__restrict__ node_elems,
                                  volatile uint32_t *node_done, volatile uint32_t *node_mask, uint32_t *colour, uint32_t *__restrict__ n_coloured) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_tets || colour[e] != 0xFFFFFFFFu) return;
    uint32_t mask[kColourWords]{};
    for (uint32_t k = 0; k < npe; ++k) {
        const uint32_t n = elem_nodes[npe * e + k];
        if (node_elems[node_ptr[n] + node_done[n]] != e) return;
        __threadfence(); // the mask below is read after the counter that published it
        for (int w = 0; w < kColourWords; ++w) mask[w] |= node_mask[kColourWords * n + w];
    }
    uint32_t c = 0xFFFFFFFFu;
    for (int w = 0; w < kColourWords && c == 0xFFFFFFFFu; ++w)
        if (mask[w] != 0xFFFFFFFFu) c = 32 * w + __ffs(~mask[w]) - 1;
    if (c == 0xFFFFFFFFu) c = 32 * kColourWords - 1;
    // Only one element per node can be ready at a time, so these plain read-modify-writes have a single writer. A
    // successor that already sees the advanced counter also sees this element's colour bit (mask first, fence, counter).
    colour[e] = c;
    for (uint32_t k = 0; k < npe; ++k) {
        const uint32_t n = elem_nodes[npe * e + k];
        node_mask[kColourWords * n + (c >> 5)] = node_mask[kColourWords * n + (c >> 5)] | (1u << (c & 31));
        __threadfence();
        node_done[n] = node_done[n] + 1;
    }
    atomicAdd(n_coloured, 1u);
}

struct CubTemp {
    DeviceBuffer<uint8_t> Buf;
    void *Get(size_t bytes) {
        Buf.Reserve(bytes ? bytes : 1);
        return Buf.Ptr;
    }
};

int BitsFor(uint64_t max_value) {
    int bits = 1;
    while (bits < 64 && (max_value >> bits) != 0) ++bits;
    return bits;
}

template<typename K>
void SortPairs(CubTemp &temp, const K *keys_in, K *keys_out, const uint32_t *vals_in, uint32_t *vals_out, uint32_t n, int end_bit, cudaStream_t stream) {
    size_t bytes = 0;
    ME_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys_in, keys_out, vals_in, vals_out, int(n), 0, end_bit, stream));
    void *ptr = temp.Get(bytes);
    ME_CUDA(cub::DeviceRadixSort::SortPairs(ptr, bytes, keys_in, keys_out, vals_in, vals_out, int(n), 0, end_bit, stream));
}
void InclusiveSum(CubTemp &temp, const uint32_t *in, uint32_t *out, uint32_t n, cudaStream_t stream) {
    size_t bytes = 0;
    ME_CUDA(cub::DeviceScan::InclusiveSum(nullptr, bytes, in, out, int(n), stream));
    void *ptr = temp.Get(bytes);
    ME_CUDA(cub::DeviceScan::InclusiveSum(ptr, bytes, in, out, int(n), stream));
}
} // namespace

// ------------------------------------------------------------------------------------------------ host side
FemSystem::FemSystem(int device) : Device(device) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) Fail(ME_CUDA_ERROR, "no CUDA device: the assembly has no CPU fallback");
    if (device < 0 || device >= count) Fail(ME_BAD_ARG, "device %d out of range (%d visible)", device, count);
    ME_CUDA(cudaSetDevice(device));
    ME_CUDA(cudaStreamCreateWithFlags(&Stream, cudaStreamNonBlocking));
}
FemSystem::~FemSystem() {
    cudaSetDevice(Device);
    if (Stream) cudaStreamDestroy(Stream);
}

void FemSystem::Build(const double *points_xyz, uint32_t n_points, const uint32_t *tets, uint32_t n_tets, const Material &material, uint32_t order) {
    if (order != 1 && order != 2) Fail(ME_BAD_ARG, "element order must be 1 or 2");
    if (!points_xyz || !tets || n_points < 4 || n_tets == 0) Fail(ME_BAD_ARG, "empty tet mesh");
    if (uint64_t(n_tets) * 55 * 2 >= (uint64_t(1) << 32)) Fail(ME_BAD_ARG, "mesh too large for 32-bit contribution ids");
    for (uint64_t i = 0; i < uint64_t(n_tets) * 4; ++i)
        if (tets[i] >= n_points) Fail(ME_BAD_ARG, "tet %llu references point %u of %u", (unsigned long long)(i / 4), tets[i], n_points);
    ME_CUDA(cudaSetDevice(Device));
    Order = order;
    Npe = order == 2 ? 10 : 4;
    NumPairs = Npe * (Npe + 1) / 2;
    NumPoints = n_points;
    NumTetsIn = n_tets;
    Mat = material;
    KernelLaunches = 0;
    CubTemp temp;
    auto s = Stream;

    // Element tables -> device.
    const ElementTables tab = MakeElementTables(order);
    std::vector<double> term_w(size_t(Npe) * Npe * 4, 0.0);
    std::vector<uint8_t> term_kl(size_t(Npe) * Npe * 4, 0), term_count(size_t(Npe) * Npe, 0), pair_a, pair_c;
    for (uint32_t a = 0; a < Npe; ++a)
        for (uint32_t c = 0; c < Npe; ++c) {
            uint32_t count = 0;
            for (uint32_t k = 0; k < 4; ++k)      // same (k, l) visiting order as the reference's accumulation loop (:303-311)
                for (uint32_t l = 0; l < 4; ++l) {
                    const double w = tab.Grad[((a * 4 + k) * Npe + c) * 4 + l];
                    if (w == 0) continue;
                    if (count == 4) Fail(ME_BAD_ARG, "internal: more than 4 gradient terms");
                    term_w[4 * (a * Npe + c) + count] = w;
                    term_kl[4 * (a * Npe + c) + count] = uint8_t(4 * k + l);
                    ++count;
                }
            term_count[a * Npe + c] = uint8_t(count);
        }
    for (uint32_t a = 0; a < Npe; ++a)
        for (uint32_t c = 0; c <= a; ++c) {
            pair_a.push_back(uint8_t(a));
            pair_c.push_back(uint8_t(c));
        }
    TabMass.Upload(tab.Mass, s);
    TabTermW.Upload(term_w, s);
    TabTermKL.Upload(term_kl, s);
    TabTermCount.Upload(term_count, s);
    TabPairA.Upload(pair_a, s);
    TabPairC.Upload(pair_c, s);

    // Mesh -> device; FilterDegenerate.
    Points.Upload(points_xyz, size_t(n_points) * 3, s);
    DeviceBuffer<uint32_t> tets_in;
    tets_in.Upload(tets, size_t(n_tets) * 4, s);
    DeviceBuffer<uint8_t> keep;
    keep.Reserve(n_tets);
    FilterKernel<<<Blocks(n_tets), kThreads, 0, s>>>(Points.Ptr, reinterpret_cast<const uint4 *>(tets_in.Ptr), n_tets, keep.Ptr);
    Tets.Reserve(size_t(n_tets) * 4);
    DeviceBuffer<uint32_t> d_count;
    d_count.Reserve(4);
    {
        size_t bytes = 0;
        ME_CUDA(cub::DeviceSelect::Flagged(nullptr, bytes, reinterpret_cast<const uint4 *>(tets_in.Ptr), keep.Ptr, reinterpret_cast<uint4 *>(Tets.Ptr), d_count.Ptr, int(n_tets), s));
        void *ptr = temp.Get(bytes);
        ME_CUDA(cub::DeviceSelect::Flagged(ptr, bytes, reinterpret_cast<const uint4 *>(tets_in.Ptr), keep.Ptr, reinterpret_cast<uint4 *>(Tets.Ptr), d_count.Ptr, int(n_tets), s));
    }
    ME_CUDA(cudaMemcpyAsync(&NumTets, d_count.Ptr, 4, cudaMemcpyDeviceToHost, s));
    ME_CUDA(cudaStreamSynchronize(s));
    KernelLaunches += 2;
    if (NumTets == 0) Fail(ME_BAD_ARG, "every tet is degenerate");
    const uint32_t T = NumTets;

    Basis.Reserve(size_t(13) * T);
    BasisKernel<<<Blocks(T), kThreads, 0, s>>>(Points.Ptr, reinterpret_cast<const uint4 *>(Tets.Ptr), T, Basis.Ptr);
    ++KernelLaunches;

    // Node numbering (BuildQuadMesh).
    ElemNodes.Reserve(size_t(T) * Npe);
    CornerKernel<<<Blocks(uint64_t(T) * 4), kThreads, 0, s>>>(Tets.Ptr, T, Npe, ElemNodes.Ptr);
    ++KernelLaunches;
    NodeCount = n_points;
    DeviceBuffer<uint64_t> keys_a, keys_b;
    DeviceBuffer<uint32_t> vals_a, vals_b, heads, uid1, keys32_a, keys32_b, order_in;
    if (order == 2) {
        const uint32_t n = T * 6;
        keys_a.Reserve(n), keys_b.Reserve(n), vals_a.Reserve(n), vals_b.Reserve(n), heads.Reserve(n), uid1.Reserve(n);
        EdgeKeyKernel<<<Blocks(n), kThreads, 0, s>>>(Tets.Ptr, T, n_points, keys_a.Ptr, vals_a.Ptr);
        SortPairs(temp, keys_a.Ptr, keys_b.Ptr, vals_a.Ptr, vals_b.Ptr, n, BitsFor(uint64_t(n_points) * n_points), s);
        HeadKernel<<<Blocks(n), kThreads, 0, s>>>(keys_b.Ptr, n, heads.Ptr);
        InclusiveSum(temp, heads.Ptr, uid1.Ptr, n, s);
        uint32_t n_edges = 0;
        ME_CUDA(cudaMemcpyAsync(&n_edges, uid1.Ptr + (n - 1), 4, cudaMemcpyDeviceToHost, s));
        ME_CUDA(cudaStreamSynchronize(s));
        DeviceBuffer<uint32_t> first, iota, first_sorted, u_sorted, rank;
        first.Reserve(n_edges), iota.Reserve(n_edges), first_sorted.Reserve(n_edges), u_sorted.Reserve(n_edges), rank.Reserve(n_edges);
        FirstSeenKernel<<<Blocks(n), kThreads, 0, s>>>(heads.Ptr, uid1.Ptr, vals_b.Ptr, n, first.Ptr, iota.Ptr);
        SortPairs(temp, first.Ptr, first_sorted.Ptr, iota.Ptr, u_sorted.Ptr, n_edges, BitsFor(n), s);
        RankKernel<<<Blocks(n_edges), kThreads, 0, s>>>(u_sorted.Ptr, n_edges, rank.Ptr);
        MidsideKernel<<<Blocks(n), kThreads, 0, s>>>(uid1.Ptr, vals_b.Ptr, rank.Ptr, n, n_points, ElemNodes.Ptr);
        ME_CUDA(cudaStreamSynchronize(s)); // the temporaries above die here
        NodeCount = n_points + n_edges;
        KernelLaunches += 8;
    }
    N = 3 * NodeCount;

    // Symbolic pattern: lower block CSC + per-block contribution lists.
    {
        const uint32_t n = T * NumPairs;
        keys_a.Reserve(n), keys_b.Reserve(n), vals_a.Reserve(n), heads.Reserve(n), uid1.Reserve(n);
        Contrib.Reserve(n);
        PairKeyKernel<<<Blocks(n), kThreads, 0, s>>>(ElemNodes.Ptr, T, Npe, NumPairs, TabPairA.Ptr, TabPairC.Ptr, NodeCount, keys_a.Ptr, vals_a.Ptr);
        SortPairs(temp, keys_a.Ptr, keys_b.Ptr, vals_a.Ptr, Contrib.Ptr, n, BitsFor(uint64_t(NodeCount) * NodeCount), s);
        HeadKernel<<<Blocks(n), kThreads, 0, s>>>(keys_b.Ptr, n, heads.Ptr);
        InclusiveSum(temp, heads.Ptr, uid1.Ptr, n, s);
        ME_CUDA(cudaMemcpyAsync(&NumBlocks, uid1.Ptr + (n - 1), 4, cudaMemcpyDeviceToHost, s));
        ME_CUDA(cudaStreamSynchronize(s));
        BlkRow.Reserve(NumBlocks), BlkCol.Reserve(NumBlocks), ContribPtr.Reserve(size_t(NumBlocks) + 1), BlkColPtr.Reserve(size_t(NodeCount) + 1);
        BlockKernel<<<Blocks(n), kThreads, 0, s>>>(keys_b.Ptr, heads.Ptr, uid1.Ptr, n, NodeCount, BlkRow.Ptr, BlkCol.Ptr, ContribPtr.Ptr, BlkColPtr.Ptr);
        ME_CUDA(cudaMemcpyAsync(ContribPtr.Ptr + NumBlocks, &n, 4, cudaMemcpyHostToDevice, s));
        ME_CUDA(cudaMemcpyAsync(BlkColPtr.Ptr + NodeCount, &NumBlocks, 4, cudaMemcpyHostToDevice, s));
        ME_CUDA(cudaStreamSynchronize(s));
        // The order in which the assembly's threads take the blocks: by contribution count (two radix passes over NumBlocks keys).
        AssembleOrder.Reserve(NumBlocks);
        keys32_a.Reserve(NumBlocks), keys32_b.Reserve(NumBlocks), order_in.Reserve(NumBlocks);
        ContribCountKernel<<<Blocks(NumBlocks), kThreads, 0, s>>>(ContribPtr.Ptr, NumBlocks, keys32_a.Ptr, order_in.Ptr);
        SortPairs(temp, keys32_a.Ptr, keys32_b.Ptr, order_in.Ptr, AssembleOrder.Ptr, NumBlocks, 16, s);
        KernelLaunches += 7;
    }

    // Numeric assembly.
    KBlk.Reserve(size_t(9) * NumBlocks), MBlk.Reserve(NumBlocks);
    {
        AssembleArgs a{ContribPtr.Ptr, Contrib.Ptr, AssembleOrder.Ptr, Basis.Ptr, TabMass.Ptr, TabTermW.Ptr, TabTermKL.Ptr, TabTermCount.Ptr, TabPairA.Ptr, TabPairC.Ptr,
                       KBlk.Ptr, MBlk.Ptr, NumBlocks, T, Npe, NumPairs, material.Lambda(), material.Mu(), material.Density};
        cudaEvent_t e0, e1;
        ME_CUDA(cudaEventCreate(&e0));
        ME_CUDA(cudaEventCreate(&e1));
        ME_CUDA(cudaEventRecord(e0, s));
        AssembleKernel<<<Blocks(NumBlocks), kThreads, 0, s>>>(a);
        ME_CUDA(cudaEventRecord(e1, s));
        SymmetriseDiagKernel<<<Blocks(NodeCount), kThreads, 0, s>>>(BlkColPtr.Ptr, NodeCount, KBlk.Ptr);
        ME_CUDA(cudaStreamSynchronize(s));
        ME_CUDA(cudaEventElapsedTime(&AssembleKernelMs, e0, e1));
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        KernelLaunches += 2;
    }

    // Full symmetric block CSR for the mat-vecs.
    {
        const uint32_t nb = NumBlocks;
        DeviceBuffer<uint32_t> rk_a, rk_b, rv_a, sorted_blk, low_ptr;
        rk_a.Reserve(nb), rk_b.Reserve(nb), rv_a.Reserve(nb), sorted_blk.Reserve(nb), low_ptr.Reserve(size_t(NodeCount) + 1);
        RowKeyKernel<<<Blocks(nb), kThreads, 0, s>>>(BlkRow.Ptr, nb, rk_a.Ptr, rv_a.Ptr);
        SortPairs(temp, rk_a.Ptr, rk_b.Ptr, rv_a.Ptr, sorted_blk.Ptr, nb, BitsFor(NodeCount), s);
        LowPtrKernel<<<Blocks(nb), kThreads, 0, s>>>(rk_b.Ptr, nb, low_ptr.Ptr);
        ME_CUDA(cudaMemcpyAsync(low_ptr.Ptr + NodeCount, &nb, 4, cudaMemcpyHostToDevice, s));
        NumFullBlocks = 2 * nb - NodeCount;
        FullRowPtr.Reserve(size_t(NodeCount) + 1), FullCol.Reserve(NumFullBlocks), FullSrc.Reserve(NumFullBlocks);
        FullRowPtrKernel<<<Blocks(size_t(NodeCount) + 1), kThreads, 0, s>>>(low_ptr.Ptr, BlkColPtr.Ptr, NodeCount, FullRowPtr.Ptr);
        FillFullKernel<<<Blocks(nb), kThreads, 0, s>>>(sorted_blk.Ptr, low_ptr.Ptr, BlkRow.Ptr, BlkCol.Ptr, BlkColPtr.Ptr, FullRowPtr.Ptr, nb, FullCol.Ptr, FullSrc.Ptr);
        KFull.Reserve(size_t(9) * NumFullBlocks), MFull.Reserve(NumFullBlocks);
        GatherFullKernel<<<Blocks(uint64_t(9) * NumFullBlocks), kThreads, 0, s>>>(FullSrc.Ptr, KBlk.Ptr, MBlk.Ptr, NumFullBlocks, KFull.Ptr, MFull.Ptr);
        ME_CUDA(cudaStreamSynchronize(s));
        KernelLaunches += 6;
    }
    ME_CUDA(cudaGetLastError());
}

void FemSystem::SpmvK(const double *x, double *y) {
    SpmvBsr3Kernel<<<Blocks(NodeCount, kSpmvRows), kThreads, 0, Stream>>>(FullRowPtr.Ptr, FullCol.Ptr, KFull.Ptr, x, y, NodeCount);
    ++KernelLaunches;
}
void FemSystem::SpmvM(const double *x, double *y) {
    SpmvMassKernel<<<Blocks(uint64_t(NodeCount) * 8), kThreads, 0, Stream>>>(FullRowPtr.Ptr, FullCol.Ptr, MFull.Ptr, x, y, NodeCount);
    ++KernelLaunches;
}
void FemSystem::SpmvMPanel(const double *x, double *y, uint32_t width) {
    const uint32_t blocks = Blocks(uint64_t(NodeCount) * 8);
    uint32_t j = 0;
    for (; j + 8 <= width; j += 8, ++KernelLaunches)
        SpmvMassPanelKernel<8><<<blocks, kThreads, 0, Stream>>>(FullRowPtr.Ptr, FullCol.Ptr, MFull.Ptr, x + size_t(j) * N, y + size_t(j) * N, NodeCount, N);
    for (; j + 4 <= width; j += 4, ++KernelLaunches)
        SpmvMassPanelKernel<4><<<blocks, kThreads, 0, Stream>>>(FullRowPtr.Ptr, FullCol.Ptr, MFull.Ptr, x + size_t(j) * N, y + size_t(j) * N, NodeCount, N);
    for (; j < width; ++j) SpmvM(x + size_t(j) * N, y + size_t(j) * N);
}

void FemSystem::ExportCsc(int which, uint64_t *colptr, uint32_t *rowidx, double *values) {
    ME_CUDA(cudaSetDevice(Device));
    const uint64_t nnz = which == 0 ? ScalarNonZerosK() : ScalarNonZerosM();
    DeviceBuffer<uint64_t> d_colptr;
    DeviceBuffer<uint32_t> d_row;
    DeviceBuffer<double> d_val;
    d_colptr.Reserve(size_t(N) + 1), d_row.Reserve(nnz), d_val.Reserve(nnz);
    ExportCscKernel<<<Blocks(NumBlocks), kThreads, 0, Stream>>>(which, BlkRow.Ptr, BlkCol.Ptr, BlkColPtr.Ptr, KBlk.Ptr, MBlk.Ptr, NumBlocks, d_colptr.Ptr, d_row.Ptr, d_val.Ptr);
    ME_CUDA(cudaMemcpyAsync(d_colptr.Ptr + N, &nnz, 8, cudaMemcpyHostToDevice, Stream));
    ME_CUDA(cudaMemcpyAsync(colptr, d_colptr.Ptr, (size_t(N) + 1) * 8, cudaMemcpyDeviceToHost, Stream));
    ME_CUDA(cudaMemcpyAsync(rowidx, d_row.Ptr, nnz * 4, cudaMemcpyDeviceToHost, Stream));
    ME_CUDA(cudaMemcpyAsync(values, d_val.Ptr, nnz * 8, cudaMemcpyDeviceToHost, Stream));
    ME_CUDA(cudaStreamSynchronize(Stream));
}

void FemSystem::CopyElementNodes(uint32_t *out) {
    ME_CUDA(cudaSetDevice(Device));
    ME_CUDA(cudaMemcpyAsync(out, ElemNodes.Ptr, size_t(NumTets) * Npe * 4, cudaMemcpyDeviceToHost, Stream));
    ME_CUDA(cudaStreamSynchronize(Stream));
}
void FemSystem::CopyFullPattern(std::vector<uint32_t> &rowptr, std::vector<uint32_t> &col) {
    ME_CUDA(cudaSetDevice(Device));
    rowptr.resize(size_t(NodeCount) + 1);
    col.resize(NumFullBlocks);
    ME_CUDA(cudaMemcpyAsync(rowptr.data(), FullRowPtr.Ptr, rowptr.size() * 4, cudaMemcpyDeviceToHost, Stream));
    ME_CUDA(cudaMemcpyAsync(col.data(), FullCol.Ptr, col.size() * 4, cudaMemcpyDeviceToHost, Stream));
    ME_CUDA(cudaStreamSynchronize(Stream));
}
void FemSystem::CopyNodeCoords(std::vector<float> &xyz) {
    ME_CUDA(cudaSetDevice(Device));
    DeviceBuffer<float> d;
    d.Reserve(size_t(NodeCount) * 3);
    const uint64_t n = std::max<uint64_t>(uint64_t(NumPoints) * 3, uint64_t(NumTets) * 6);
    NodeCoordKernel<<<Blocks(n), kThreads, 0, Stream>>>(Points.Ptr, NumPoints, ElemNodes.Ptr, NumTets, Order, d.Ptr);
    xyz.resize(size_t(NodeCount) * 3);
    ME_CUDA(cudaMemcpyAsync(xyz.data(), d.Ptr, xyz.size() * 4, cudaMemcpyDeviceToHost, Stream));
    ME_CUDA(cudaStreamSynchronize(Stream));
}

void FemSystem::ColourElements(uint32_t *out, uint32_t *n_colours) {
    ME_CUDA(cudaSetDevice(Device));
    auto s = Stream;
    CubTemp temp;
    const uint32_t n = NumTets * Npe;
    DeviceBuffer<uint32_t> k_a, k_b, v_a, node_elems, node_ptr, node_done, node_mask, colour, counter;
    k_a.Reserve(n), k_b.Reserve(n), v_a.Reserve(n), node_elems.Reserve(n), node_ptr.Reserve(size_t(NodeCount) + 1), node_done.Reserve(NodeCount);
    node_mask.Reserve(size_t(NodeCount) * kColourWords), colour.Reserve(NumTets), counter.Reserve(1);
    IncidenceKeyKernel<<<Blocks(n), kThreads, 0, s>>>(ElemNodes.Ptr, n, Npe, k_a.Ptr, v_a.Ptr);
    SortPairs(temp, k_a.Ptr, k_b.Ptr, v_a.Ptr, node_elems.Ptr, n, BitsFor(NodeCount), s);
    NodePtrKernel<<<Blocks(n), kThreads, 0, s>>>(k_b.Ptr, n, NodeCount, node_ptr.Ptr);
    ME_CUDA(cudaMemsetAsync(node_done.Ptr, 0, size_t(NodeCount) * 4, s));
    ME_CUDA(cudaMemsetAsync(node_mask.Ptr, 0, size_t(NodeCount) * kColourWords * 4, s));
    ME_CUDA(cudaMemsetAsync(colour.Ptr, 0xFF, size_t(NumTets) * 4, s));
    ME_CUDA(cudaMemsetAsync(counter.Ptr, 0, 4, s));
    uint32_t done = 0, rounds = 0;
    while (done < NumTets) {
        for (int r = 0; r < 32; ++r) ColourRoundKernel<<<Blocks(NumTets), kThreads, 0, s>>>(ElemNodes.Ptr, NumTets, Npe, node_ptr.Ptr, node_elems.Ptr, node_done.Ptr, node_mask.Ptr, colour.Ptr, counter.Ptr);
        rounds += 32;
        const uint32_t before = done;
        ME_CUDA(cudaMemcpyAsync(&done, counter.Ptr, 4, cudaMemcpyDeviceToHost, s));
        ME_CUDA(cudaStreamSynchronize(s));
        if (done == before) Fail(ME_CUDA_ERROR, "colouring made no progress after %u rounds (%u of %u elements)", rounds, done, NumTets);
    }
    KernelLaunches += rounds + 2;
    ME_CUDA(cudaMemcpyAsync(out, colour.Ptr, size_t(NumTets) * 4, cudaMemcpyDeviceToHost, s));
    ME_CUDA(cudaStreamSynchronize(s));
    uint32_t top = 0;
    for (uint32_t e = 0; e < NumTets; ++e) top = std::max(top, out[e]);
    if (n_colours) *n_colours = top + 1;
}

} // namespace me
