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
    uint32_t Launches{0};
};

// out[j] = V[:, j] . x for j < cols (device out, length cols). Deterministic two-stage reduction.
void GemvT(DenseWorkspace &, const double *V, size_t n, uint32_t cols, const double *x, double *out, cudaStream_t);
// y -= V[:, :cols] * c (c: device, length cols)
void GemvNSub(DenseWorkspace &, const double *V, size_t n, uint32_t cols, const double *c, double *y, cudaStream_t);
// y = a*x + b*y ; when y_out != nullptr writes there instead of y
void Axpby(DenseWorkspace &, size_t n, double a, const double *x, double b, const double *y, double *out, cudaStream_t);
// C[n x cols_out] = V[n x m] * Q[m x cols_out] (Q: device column-major, leading dimension ldq). FP64 DMMA.
void TallGemm(DenseWorkspace &, const double *V, size_t n, uint32_t m, const double *Q, uint32_t ldq, uint32_t cols_out, double *C, cudaStream_t);

} // namespace me
