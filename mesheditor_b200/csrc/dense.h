// Dense tall-skinny helpers for the Lanczos basis V (n x m, column-major, resident in HBM): the reorthogonalisation
// GEMVs, vector updates and the basis rotation GEMM. Reference call sites: lib/spectra/include/Spectra/LinAlg/Lanczos.h
// :139-182 (adjoint_product, f -= V*Vf), HermEigsBase.h:105-155 (compress_V), :447-470 (eigenvectors = V * ritz_vec).
#pragma once

#include "common.h"

#include <cstdint>

namespace me {

struct DenseWorkspace {
    DeviceBuffer<double> Partial; // per-split partial sums of the two-stage (deterministic) reductions
    DeviceBuffer<double> Coeff;   // small coefficient vectors uploaded per call
    DeviceBuffer<double> GramPartial; // per-split partial Gram blocks
    uint32_t Launches{0};
};

// out[j] = V[:, j] . x for j < cols (device out, length cols). Deterministic two-stage reduction.
void GemvT(DenseWorkspace &, const double *V, size_t n, uint32_t cols, const double *x, double *out, cudaStream_t);
// y -= V[:, :cols] * c (c: device, length cols)
void GemvNSub(DenseWorkspace &, const double *V, size_t n, uint32_t cols, const double *c, double *y, cudaStream_t);
// y = a*x + b*y ; when y_out != nullptr writes there instead of y
void Axpby(DenseWorkspace &, size_t n, double a, const double *x, double b, const double *y, double *out, cudaStream_t);
// C[n x cols_out] = alpha * V[n x m] * Q[m x cols_out] + beta * C (Q: device column-major, leading dimension ldq; C must
// not alias V). FP64 DMMA; at most 8 output columns take the streaming form that reads V exactly once.
void TallGemm(DenseWorkspace &, const double *V, size_t n, uint32_t m, const double *Q, uint32_t ldq, uint32_t cols_out, double *C, cudaStream_t, double alpha = 1.0, double beta = 0.0);
// out[a x c] (column-major, leading dimension ldo, device) = X[:, :a]^T Y[:, :c]: the projections of the block methods
// (mesh2modes.cpp:379-388 Kr/Mr/C; the block form of Lanczos.h:139-182). Deterministic: fixed-order two-stage reduction.
// X is read once per 8 columns of Y.
void Gram(DenseWorkspace &, const double *X, size_t n, uint32_t a, const double *Y, uint32_t c, double *out, uint32_t ldo, cudaStream_t);

// One plane rotation of the implicit QL iteration (hosteig.cpp): rows Row and Row + 1 of the transposed eigenvector matrix
// become (C z_i - S z_i1, S z_i + C z_i1).
struct QlRotation {
    double C, S;
    uint32_t Row, Pad;
};
// Largest matrix order ApplyRotations takes (one warp's 32 columns of all rows live in shared memory).
constexpr uint32_t kMaxDeviceRotationOrder = 768;
// zt (m x m, row-major, device): the whole rotation history applied in order, one thread per column, the column in shared
// memory and the row shared by two consecutive rotations carried in a register. The arithmetic per entry is the host's.
// `rotations` must be readable up to the next multiple of 512 records past `count` (staged in whole batches).
void ApplyRotations(double *zt, uint32_t m, const QlRotation *rotations, size_t count, cudaStream_t, uint32_t &launches);
// zt (m x m, row-major, device) <- the transposed orthogonal basis Q^T of a Householder tridiagonalisation, Q = H_1 H_2 .. H_{m-1},
// H_i = I - u_i u_i^T / h[i] over the leading i entries (lanczos.h HouseholderTridiagonal): `u` row-major with u_i[k] at [i * m + k],
// `v` row-major with u_i[k] / h[i] at [i * m + k] (k < i), `h` of length m (0: no reflection). One thread per column of Q, the column in
// shared memory, every reflector a dot product and an update of its leading entries: the 5 ms this accumulation takes on the host
// become ~1 ms beside the host's QL iteration. m <= kMaxDeviceRotationOrder.
void HouseholderBasis(const double *u, const double *v, const double *h, uint32_t m, double *zt, cudaStream_t, uint32_t &launches);
// out (m x k, column-major, device): column j = row rows[j] of zt (m x m, row-major): the picked eigenvectors, as TallGemm takes them.
void GatherRows(const double *zt, uint32_t m, const uint32_t *rows, uint32_t k, double *out, cudaStream_t, uint32_t &launches);

// out[i] = the (i + 1)-th value of Spectra's SimpleRandom (Park-Miller LCG x <- 16807 x mod 2^31 - 1 from seed 1) mapped to
// (-0.5, 0.5): the reference's start residual (lib/spectra/include/Spectra/Util/SimpleRandom.h), value for value. The generator
// is a power map, so every thread jumps to its own stretch by modular exponentiation (the 4.2M values of the 1M-tet start block
// took 37 ms of host loop, allocation and pageable copy per solve).
void FillSimpleRandom(double *out, size_t count, cudaStream_t, uint32_t &launches);

} // namespace me
