// Warm re-solve: block subspace iteration with Rayleigh-Ritz projection, prefix locking and deflation, seeded by a
// prior eigenbasis. Replaces SubspaceIterate (src/audio/mesh2modes.cpp:339-428) on the SolveReuse::SeedBasis path of
// mesh2modes (:459-478). The structure follows the reference step by step; only the carrier changes: every n x w object
// lives in HBM, the panel solve is SparseCholesky::Solve (one pass over the factor per 8 columns), the projections are
// the streaming DMMA Gram / tall-GEMM kernels of dense.cu, and only the w x w pencil (w <= nev + 15) visits the host.
#include "lanczos.h"

#include <algorithm>
#include <cmath>
#include <limits>
#include <numeric>
#include <random>

namespace me {
namespace {
__global__ void SeedCastKernel(const float *__restrict__ src, size_t count, double *__restrict__ dst) {
    const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < count) dst[i] = double(src[i]);
}

// Generalized symmetric-definite eigenproblem Kr q = theta Mr q of order w on the host (what the reference asks of
// Eigen::GeneralizedSelfAdjointEigenSolver, mesh2modes.cpp:398): Mr = L L^T, C = L^-1 Kr L^-T, C y = theta y,
// q = L^-T y. kr, mr: row-major w x w symmetric. theta ascending; q row-major w x w (columns = vectors, q^T Mr q = I).
bool GeneralizedEigen(uint32_t w, const std::vector<double> &kr, std::vector<double> mr, std::vector<double> &theta, std::vector<double> &q) {
    auto at = [w](std::vector<double> &m, uint32_t r, uint32_t c) -> double & { return m[size_t(r) * w + c]; };
    // Cholesky of Mr (lower, in place).
    for (uint32_t j = 0; j < w; ++j) {
        double d = at(mr, j, j);
        for (uint32_t k = 0; k < j; ++k) d -= at(mr, j, k) * at(mr, j, k);
        if (!(d > 0.0) || !std::isfinite(d)) return false;
        const double l = std::sqrt(d);
        at(mr, j, j) = l;
        for (uint32_t i = j + 1; i < w; ++i) {
            double v = at(mr, i, j);
            for (uint32_t k = 0; k < j; ++k) v -= at(mr, i, k) * at(mr, j, k);
            at(mr, i, j) = v / l;
        }
    }
    // C = L^-1 Kr L^-T: forward substitution on the rows, then on the columns.
    std::vector<double> c = kr;
    for (uint32_t col = 0; col < w; ++col)
        for (uint32_t i = 0; i < w; ++i) {
            double v = at(c, i, col);
            for (uint32_t k = 0; k < i; ++k) v -= at(mr, i, k) * at(c, k, col);
            at(c, i, col) = v / at(mr, i, i);
        }
    for (uint32_t row = 0; row < w; ++row)
        for (uint32_t j = 0; j < w; ++j) {
            double v = at(c, row, j);
            for (uint32_t k = 0; k < j; ++k) v -= at(mr, j, k) * at(c, row, k);
            at(c, row, j) = v / at(mr, j, j);
        }
    for (uint32_t i = 0; i < w; ++i)
        for (uint32_t j = i + 1; j < w; ++j) at(c, i, j) = at(c, j, i) = 0.5 * (at(c, i, j) + at(c, j, i));
    std::vector<double> d;
    if (!SymmetricEigen(w, c, d)) return false;
    std::vector<uint32_t> order(w);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return d[a] < d[b]; });
    theta.resize(w);
    q.assign(size_t(w) * w, 0.0);
    for (uint32_t j = 0; j < w; ++j) {
        theta[j] = d[order[j]];
        // back substitution L^T x = y
        for (int64_t i = int64_t(w) - 1; i >= 0; --i) {
            double v = c[size_t(i) * w + order[j]];
            for (uint32_t k = uint32_t(i) + 1; k < w; ++k) v -= at(mr, k, uint32_t(i)) * q[size_t(k) * w + j];
            q[size_t(i) * w + j] = v / at(mr, uint32_t(i), uint32_t(i));
        }
    }
    return true;
}
} // namespace

SubspaceOutcome SubspaceIteration::Compute(uint32_t nev, uint32_t p, double tol, uint32_t max_iters, const float *seed, uint32_t seed_cols, const volatile int *cancelled) {
    SubspaceOutcome out;
    const size_t n = Fem.N;
    if (nev < 1 || p < nev || p > n) Fail(ME_BAD_ARG, "subspace iteration sizes: need 1 <= nev <= p <= n (nev %u, p %u, n %zu)", nev, p, n);
    ME_CUDA(cudaSetDevice(Fem.Device));
    auto s = Fem.Stream;
    const uint32_t launches0 = Fem.KernelLaunches, f_launches0 = Factor.Stats.KernelLaunches;
    auto col = [&](double *base, uint32_t j) { return base + size_t(j) * n; };
    auto mass_product = [&](const double *x, double *y, uint32_t cols) {
        Fem.SpmvMPanel(x, y, cols);
    };

    DeviceBuffer<double> MX, Xbar, MXbar, XL, MXL, DKr, DMr, DC, DQ;
    MX.Reserve(n * p), Xbar.Reserve(n * p), MXbar.Reserve(n * p), MXL.Reserve(n * nev), Vectors.Reserve(n * nev);
    DKr.Reserve(size_t(p) * p), DMr.Reserve(size_t(p) * p), DC.Reserve(size_t(p) * p), DQ.Reserve(size_t(p) * p);
    double *xl = Vectors.Ptr; // the locked Ritz vectors are the result

    // Seed panel (mesh2modes.cpp:352-363): the warm basis (float) in the leading columns, Gaussian columns from
    // mt19937_64{20260710} behind them; the iteration carries M X.
    {
        const uint32_t seeded = std::min(seed_cols, p);
        if (seeded) {
            DeviceBuffer<float> staged;
            staged.Upload(seed, n * seeded, s);
            const size_t count = n * seeded;
            SeedCastKernel<<<uint32_t((count + 255) / 256), 256, 0, s>>>(staged.Ptr, count, Xbar.Ptr);
            ++out.KernelLaunches;
            ME_CUDA(cudaStreamSynchronize(s));
        }
        if (seeded < p) {
            std::vector<double> fill(n * (p - seeded));
            std::mt19937_64 rng{20260710};
            std::normal_distribution<double> gauss;
            for (auto &v : fill) v = gauss(rng);
            ME_CUDA(cudaMemcpyAsync(col(Xbar.Ptr, seeded), fill.data(), fill.size() * sizeof(double), cudaMemcpyHostToDevice, s));
            ME_CUDA(cudaStreamSynchronize(s));
        }
        mass_product(Xbar.Ptr, MX.Ptr, p);
    }

    std::vector<double> theta_locked(nev, 0.0), prev_lambda(nev, std::numeric_limits<double>::max());
    std::vector<double> kr, mr, cmat, theta, q, dscale, host(size_t(p) * p);
    uint32_t c = 0; // locked count
    cudaEvent_t ev0, ev1;
    ME_CUDA(cudaEventCreate(&ev0));
    ME_CUDA(cudaEventCreate(&ev1));
    auto download = [&](const double *dev, uint32_t rows, uint32_t cols, std::vector<double> &row_major) {
        // device column-major rows x cols -> host row-major
        ME_CUDA(cudaMemcpyAsync(host.data(), dev, size_t(rows) * cols * sizeof(double), cudaMemcpyDeviceToHost, s));
        ME_CUDA(cudaStreamSynchronize(s));
        row_major.resize(size_t(rows) * cols);
        for (uint32_t r = 0; r < rows; ++r)
            for (uint32_t k = 0; k < cols; ++k) row_major[size_t(r) * cols + k] = host[r + size_t(k) * rows];
    };

    for (uint32_t iter = 0; iter < max_iters; ++iter) {
        if (cancelled && *cancelled) {
            out.Cancelled = true;
            break;
        }
        const uint32_t w = p - c;
        // (K - sigma M) Xbar = M X over the whole panel.
        ME_CUDA(cudaEventRecord(ev0, s));
        Factor.Solve(MX.Ptr, Xbar.Ptr, w);
        ME_CUDA(cudaEventRecord(ev1, s));
        out.OpApplications += w;
        // Kr = Xbar^T M X ; M Xbar
        Gram(Ws, Xbar.Ptr, n, w, MX.Ptr, w, DKr.Ptr, w, s);
        mass_product(Xbar.Ptr, MXbar.Ptr, w);
        if (c > 0) {
            // C = XL^T M Xbar ; Xbar -= XL C ; M Xbar -= MXL C ; Kr -= C^T theta C (host, below)
            Gram(Ws, xl, n, c, MXbar.Ptr, w, DC.Ptr, c, s);
            TallGemm(Ws, xl, n, c, DC.Ptr, c, w, Xbar.Ptr, s, -1.0, 1.0);
            TallGemm(Ws, MXL.Ptr, n, c, DC.Ptr, c, w, MXbar.Ptr, s, -1.0, 1.0);
        }
        Gram(Ws, Xbar.Ptr, n, w, MXbar.Ptr, w, DMr.Ptr, w, s);
        download(DKr.Ptr, w, w, kr);
        download(DMr.Ptr, w, w, mr);
        {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, ev0, ev1) == cudaSuccess) out.OpSolveMs += ms;
        }
        if (c > 0) {
            download(DC.Ptr, c, w, cmat);
            for (uint32_t i = 0; i < w; ++i)
                for (uint32_t j = 0; j < w; ++j) {
                    double v = 0;
                    for (uint32_t l = 0; l < c; ++l) v += cmat[size_t(l) * w + i] * theta_locked[l] * cmat[size_t(l) * w + j];
                    kr[size_t(i) * w + j] -= v;
                }
        }
        // Symmetrise, scale the columns to unit M-norm, solve the small pencil.
        dscale.resize(w);
        for (uint32_t i = 0; i < w; ++i) dscale[i] = 1.0 / std::sqrt(mr[size_t(i) * w + i]);
        for (uint32_t i = 0; i < w; ++i)
            for (uint32_t j = i; j < w; ++j) {
                const double ks = 0.5 * (kr[size_t(i) * w + j] + kr[size_t(j) * w + i]) * dscale[i] * dscale[j];
                const double ms = 0.5 * (mr[size_t(i) * w + j] + mr[size_t(j) * w + i]) * dscale[i] * dscale[j];
                kr[size_t(i) * w + j] = kr[size_t(j) * w + i] = ks;
                mr[size_t(i) * w + j] = mr[size_t(j) * w + i] = ms;
            }
        if (!GeneralizedEigen(w, kr, mr, theta, q)) break; // the reference returns the empty result here (:399)
        for (uint32_t i = 0; i < w; ++i)
            for (uint32_t j = 0; j < w; ++j) q[size_t(i) * w + j] *= dscale[i];

        // Lock the leading prefix of active pairs whose eigenvalue settled (:404-416).
        uint32_t newly_locked = 0;
        for (uint32_t i = 0; i < w && c + i < nev; ++i) {
            const double lambda = theta[i] + Sigma;
            const double rel = std::abs(lambda - prev_lambda[c + i]) / std::max(std::abs(lambda), std::abs(Sigma));
            prev_lambda[c + i] = lambda;
            if (newly_locked == i && rel < tol) ++newly_locked;
        }
        // q to the device, column-major w x w.
        for (uint32_t i = 0; i < w; ++i)
            for (uint32_t j = 0; j < w; ++j) host[i + size_t(j) * w] = q[size_t(i) * w + j];
        ME_CUDA(cudaMemcpyAsync(DQ.Ptr, host.data(), size_t(w) * w * sizeof(double), cudaMemcpyHostToDevice, s));
        if (newly_locked > 0) {
            TallGemm(Ws, Xbar.Ptr, n, w, DQ.Ptr, w, newly_locked, col(xl, c), s);
            TallGemm(Ws, MXbar.Ptr, n, w, DQ.Ptr, w, newly_locked, col(MXL.Ptr, c), s);
            for (uint32_t i = 0; i < newly_locked; ++i) theta_locked[c + i] = theta[i];
            c += newly_locked;
        }
        out.Iterations = iter + 1;
        if (c >= nev) {
            out.Converged = true;
            break;
        }
        // Rotate the maintained M X onto the remaining active Ritz vectors.
        TallGemm(Ws, MXbar.Ptr, n, w, DQ.Ptr + size_t(newly_locked) * w, w, w - newly_locked, MX.Ptr, s);
        ME_CUDA(cudaStreamSynchronize(s)); // host staging buffer is reused next iteration
    }
    ME_CUDA(cudaStreamSynchronize(s));
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    Factor.CheckSolves();
    if (out.Converged) out.Eigenvalues.assign(prev_lambda.begin(), prev_lambda.end());
    out.KernelLaunches += (Fem.KernelLaunches - launches0) + (Factor.Stats.KernelLaunches - f_launches0) + Ws.Launches;
    return out;
}

} // namespace me
