// Shift-invert Lanczos for the lowest eigenpairs of K x = lambda M x, basis and all vector work on the device.
// Replaces Spectra::SymGEigsShiftSolver<CholeskyShiftInvert, SparseSymMatProd, ShiftInvert> as configured by the
// reference (src/audio/mesh2modes.cpp:485-487; lib/spectra/include/Spectra/HermEigsBase.h:105-390, LinAlg/Lanczos.h:62-187).
#pragma once

#include "cholesky.h"
#include "dense.h"
#include "fem.h"

#include <atomic>
#include <vector>

namespace me {

constexpr uint32_t kLanczosBlock = 8; // columns per operator application of the block form = width of one panel sweep

struct LanczosOutcome {
    std::vector<double> Eigenvalues; // ascending, nev of them when converged
    uint32_t OpApplications{0}, Restarts{0};
    bool Converged{false}, Cancelled{false};
    bool RankLost{false};            // block form only: a new block lost rank (the caller falls back to the single-vector form)
    double OpSolveMs{0};             // device time inside the shift-invert operator (M x + two triangular solves)
    uint32_t KernelLaunches{0};
};

// Dense symmetric eigen-decomposition on the host (Householder tridiagonalisation + implicit QL), for the projected
// matrix H (at most ncv x ncv). a: row-major n x n, overwritten with the eigenvectors (columns); d: eigenvalues, unsorted.
bool SymmetricEigen(uint32_t n, std::vector<double> &a, std::vector<double> &d);
// The same in two halves. Reduce: Householder tridiagonalisation + implicit QL on the tridiagonal entries; leaves in `a` the
// TRANSPOSED orthogonal basis of the tridiagonal form (row-major: row j is what becomes eigenvector j), in `d` the eigenvalues
// and in `rotations` the QL rotation history, each mixing rows Row and Row + 1 of `a`. Apply: the history applied to `a` and
// `a` transposed back (a[r * n + c] = component r of vector c), on host threads; dense.h ApplyRotations is the device form.
bool SymmetricEigenReduce(uint32_t n, std::vector<double> &a, std::vector<double> &d, std::vector<QlRotation> &rotations);
void SymmetricEigenApplyHost(uint32_t n, std::vector<double> &a, const std::vector<QlRotation> &rotations);
// The three stages of SymmetricEigenReduce on their own. Tridiagonal: leaves the diagonal in d, the off-diagonal in e (e[i] couples
// i and i + 1, e[n - 1] = 0) and the reflectors in `a` (step i reflects with I - u_i u_i^T / h[i], h[i] = 0 meaning none: u_i in
// row i left of the diagonal, u_i / h[i] in column i above it). BasisHost: `a` <- the transposed orthogonal basis accumulated from
// those reflectors (dense.h HouseholderBasis is the device form). Ql: the rotation history from d and e alone.
void HouseholderTridiagonal(uint32_t n, std::vector<double> &a, std::vector<double> &d, std::vector<double> &e, std::vector<double> &h);
void HouseholderBasisHost(uint32_t n, std::vector<double> &a, const std::vector<double> &h);
bool TridiagonalQl(uint32_t n, std::vector<double> &d, std::vector<double> &e, std::vector<QlRotation> &rotations);

class ShiftInvertLanczos {
public:
    ShiftInvertLanczos(FemSystem &fem, SparseCholesky &factor, double sigma) : Fem(fem), Factor(factor), Sigma(sigma) {}
    // On success Vectors holds the n x nev M-orthonormal eigenvectors (column-major, device).
    LanczosOutcome Compute(uint32_t nev, uint32_t ncv, double tol, uint32_t max_restarts, const volatile int *cancelled);
    // The block form (kLanczosBlock vectors per operator application, panel solves). Basis size BlockBasisSize(nev);
    // needs BlockBasisSize(nev) + kLanczosBlock <= n.
    LanczosOutcome ComputeBlock(uint32_t nev, double tol, uint32_t max_restarts, const volatile int *cancelled);
    static uint32_t BlockBasisSize(uint32_t nev);
    DeviceBuffer<double> Vectors;

private:
    void Op(const double *x, double *y); // y = (K - sigma M)^-1 M x
    void OpPanel(const double *x, double *y, uint32_t width); // the same for `width` columns (one panel solve)
    FemSystem &Fem;
    SparseCholesky &Factor;
    double Sigma;
    DenseWorkspace Ws;
    DeviceBuffer<double> Tmp;
    std::vector<cudaEvent_t> OpEvents;
    uint32_t Ops{0}, OpCalls{0};
};

// Warm re-solve (SubspaceIterate, src/audio/mesh2modes.cpp:339-428): block subspace iteration with Rayleigh-Ritz,
// prefix locking at relative eigenvalue change < tol, and deflation against the locked pairs.
struct SubspaceOutcome {
    std::vector<double> Eigenvalues; // ascending, nev of them; empty when not converged (the reference's failure convention)
    uint32_t Iterations{0}, OpApplications{0};
    bool Converged{false}, Cancelled{false};
    double OpSolveMs{0};
    uint32_t KernelLaunches{0};
};

class SubspaceIteration {
public:
    SubspaceIteration(FemSystem &fem, SparseCholesky &factor, double sigma) : Fem(fem), Factor(factor), Sigma(sigma) {}
    // seed: host floats, n x seed_cols column-major (the previous solve's basis); the leading min(seed_cols, p) panel
    // columns start from it, the rest from Gaussian noise. On success Vectors holds the n x nev M-orthonormal Ritz vectors.
    SubspaceOutcome Compute(uint32_t nev, uint32_t p, double tol, uint32_t max_iters, const float *seed, uint32_t seed_cols, const volatile int *cancelled);
    DeviceBuffer<double> Vectors;

private:
    FemSystem &Fem;
    SparseCholesky &Factor;
    double Sigma;
    DenseWorkspace Ws;
};

} // namespace me
