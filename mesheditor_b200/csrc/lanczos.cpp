// Host driver of the shift-invert Lanczos iteration. See lanczos.h.
//
// The reference restarts implicitly with exact shifts (HermEigsBase.h:105-155). For a symmetric operator that is the
// same subspace as a THICK restart: keep the k wanted Ritz vectors Y = V S_k and the residual direction, after which
// the projected matrix is diag(theta) bordered by one coupling row beta * (last row of S_k) and tridiagonal from there
// on. The thick form is what is implemented: on a GPU it is one tall GEMM (V <- V S_k) instead of ncv - k sweeps of
// Givens rotations over V. Everything the reference's numerics depend on is kept: the M inner product, the full
// re-orthogonalisation test after every step with up to 5 corrections (Lanczos.h:139-182), the convergence test
// |last Ritz-vector row| * ||f|| < tol * max(eps^(2/3), |theta|) (HermEigsBase.h:158-175), the restart size rule
// (:178-202), theta -> 1/theta + sigma and the ascending final order (SymGEigsShiftSolver.h:170-176).
#include "lanczos.h"

#include <algorithm>
#include <functional>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <limits>
#include <numeric>

namespace me {

void ShiftInvertLanczos::Op(const double *x, double *y) {
    auto s = Fem.Stream;
    if (OpEvents.size() < size_t(2) * (OpCalls + 1)) {
        cudaEvent_t a, b;
        ME_CUDA(cudaEventCreate(&a));
        ME_CUDA(cudaEventCreate(&b));
        OpEvents.push_back(a);
        OpEvents.push_back(b);
    }
    ME_CUDA(cudaEventRecord(OpEvents[2 * OpCalls], s));
    Fem.SpmvM(x, Tmp.Ptr);
    Factor.Solve(Tmp.Ptr, y, 1);
    ME_CUDA(cudaEventRecord(OpEvents[2 * OpCalls + 1], s));
    ++OpCalls;
    ++Ops;
}

LanczosOutcome ShiftInvertLanczos::Compute(uint32_t nev, uint32_t ncv, double tol, uint32_t max_restarts, const volatile int *cancelled) {
    LanczosOutcome out;
    const size_t n = Fem.N;
    if (nev < 1 || nev > n - 1 || ncv <= nev || ncv > n) Fail(ME_BAD_ARG, "Lanczos sizes: need 1 <= nev < ncv <= n (nev %u, ncv %u, n %zu)", nev, ncv, n);
    ME_CUDA(cudaSetDevice(Fem.Device));
    auto s = Fem.Stream;
    const uint32_t m = ncv;
    const double eps = std::numeric_limits<double>::epsilon(), near0 = std::numeric_limits<double>::min() * 10.0;
    const double eps23 = std::pow(eps, 2.0 / 3.0), beta_thresh = eps * std::sqrt(double(n)), eps_sqrt = std::sqrt(eps);
    const uint32_t launches0 = Fem.KernelLaunches, f_launches0 = Factor.Stats.KernelLaunches;

    DeviceBuffer<double> Va, Vb, W, MW, DSmall;
    Va.Reserve(n * (m + 1)), Vb.Reserve(n * (m + 1)), W.Reserve(n), MW.Reserve(n), Tmp.Reserve(n), DSmall.Reserve(size_t(m + 2) * (m + 2));
    double *V = Va.Ptr, *V2 = Vb.Ptr;
    std::vector<double> small(m + 2), H(size_t(m) * m, 0.0);
    auto Hm = [&](uint32_t r, uint32_t c) -> double & { return H[size_t(r) * m + c]; };
    auto col = [&](double *base, uint32_t j) { return base + size_t(j) * n; };
    auto fetch = [&](uint32_t count) {
        ME_CUDA(cudaMemcpyAsync(small.data(), DSmall.Ptr, count * sizeof(double), cudaMemcpyDeviceToHost, s));
        ME_CUDA(cudaStreamSynchronize(s));
    };
    // <x, y>_M with y's M-product already in MW: GemvT over one column.
    auto mdot = [&](const double *x, const double *y) {
        Fem.SpmvM(y, MW.Ptr);
        GemvT(Ws, x, n, 1, MW.Ptr, DSmall.Ptr, s);
        fetch(1);
        return small[0];
    };

    // Initial residual: the reference's SimpleRandom (Park-Miller LCG, seed 0 -> 1), uniform in (-0.5, 0.5).
    FillSimpleRandom(W.Ptr, n, s, Ws.Launches);
    // v0 <- Op(r) / ||.||_M ; w = Op(v0); H00 = <v0, w>; f = w - H00 v0   (Arnoldi::init)
    Op(W.Ptr, col(V, 0));
    double vnorm = std::sqrt(std::max(0.0, mdot(col(V, 0), col(V, 0))));
    if (!(vnorm > near0)) Fail(ME_NOT_CONVERGED, "Lanczos: the start vector is in the null space of the operator");
    Axpby(Ws, n, 1.0 / vnorm, col(V, 0), 0.0, nullptr, col(V, 0), s);
    Op(col(V, 0), W.Ptr);
    Hm(0, 0) = mdot(col(V, 0), W.Ptr);
    Axpby(Ws, n, -Hm(0, 0), col(V, 0), 1.0, W.Ptr, col(V, 1), s);
    double beta = std::sqrt(std::max(0.0, mdot(col(V, 1), col(V, 1))));

    std::vector<double> coupling; // thick-restart border, length k
    uint32_t arrow_k = 0;         // the step that carries the border (0 = none)

    // Extends the factorisation from step `from` to step m. Column i of V is final for i < from; column `from` holds f.
    auto factorize_from = [&](uint32_t from) {
        for (uint32_t i = from; i < m; ++i) {
            bool restart = beta < near0;
            double *vi = col(V, i);
            if (!restart) {
                Axpby(Ws, n, 1.0 / beta, vi, 0.0, nullptr, vi, s);
                if (beta < eps_sqrt) {
                    const double viv = mdot(col(V, i - 1), vi);
                    restart = std::abs(viv) > eps_sqrt;
                }
            }
            if (restart) {
                // Invariant subspace: continue with a fresh direction orthogonal to V (Arnoldi::expand_basis).
                for (uint32_t attempt = 0; attempt < 5; ++attempt) {
                    std::vector<double> r(n);
                    uint64_t x = (2ull * i + 123ull * attempt) & 0x7FFFFFFFull;
                    if (x == 0) x = 1;
                    for (size_t q = 0; q < n; ++q) {
                        x = (x * 16807ull) % 2147483647ull;
                        r[q] = double(x) / 2147483647.0 - 0.5;
                    }
                    ME_CUDA(cudaMemcpyAsync(W.Ptr, r.data(), n * sizeof(double), cudaMemcpyHostToDevice, s));
                    ME_CUDA(cudaStreamSynchronize(s));
                    Op(W.Ptr, vi);
                    for (int pass = 0; pass < 3; ++pass) {
                        Fem.SpmvM(vi, MW.Ptr);
                        GemvT(Ws, V, n, i, MW.Ptr, DSmall.Ptr, s);
                        GemvNSub(Ws, V, n, i, DSmall.Ptr, vi, s);
                    }
                    beta = std::sqrt(std::max(0.0, mdot(vi, vi)));
                    if (beta > near0) break;
                }
                if (!(beta > near0)) Fail(ME_NOT_CONVERGED, "Lanczos: could not extend the basis past an invariant subspace");
                Axpby(Ws, n, 1.0 / beta, vi, 0.0, nullptr, vi, s);
            }
            const bool arrow = arrow_k != 0 && i == arrow_k;
            if (!arrow && i > 0) {
                Hm(i, i - 1) = restart ? 0.0 : beta;
                Hm(i - 1, i) = Hm(i, i - 1);
            }
            Op(vi, W.Ptr);
            if (arrow) {
                // w -= sum_j coupling_j y_j : the border of the thick restart (exact-shift restart in Spectra)
                for (uint32_t j = 0; j < i; ++j) small[j] = restart ? 0.0 : coupling[j];
                ME_CUDA(cudaMemcpyAsync(DSmall.Ptr, small.data(), i * sizeof(double), cudaMemcpyHostToDevice, s));
                GemvNSub(Ws, V, n, i, DSmall.Ptr, W.Ptr, s);
                for (uint32_t j = 0; j < i; ++j) Hm(i, j) = Hm(j, i) = restart ? 0.0 : coupling[j];
            } else if (!restart && i > 0) {
                Axpby(Ws, n, -Hm(i, i - 1), col(V, i - 1), 1.0, W.Ptr, W.Ptr, s);
            }
            Hm(i, i) = mdot(vi, W.Ptr);
            double *f = col(V, i + 1);
            Axpby(Ws, n, -Hm(i, i), vi, 1.0, W.Ptr, f, s);
            // beta = ||f||_M and the orthogonality check V^T M f in one pass: f is column i+1 of V.
            int count = 0;
            while (true) {
                Fem.SpmvM(f, MW.Ptr);
                GemvT(Ws, V, n, i + 2, MW.Ptr, DSmall.Ptr, s);
                fetch(i + 2);
                beta = std::sqrt(std::max(0.0, small[i + 1]));
                double ortho_err = 0;
                for (uint32_t j = 0; j <= i; ++j) ortho_err = std::max(ortho_err, std::abs(small[j]));
                if (!(count < 5 && ortho_err > eps * beta)) break;
                if (beta < beta_thresh) {
                    ME_CUDA(cudaMemsetAsync(f, 0, n * sizeof(double), s));
                    beta = 0;
                    break;
                }
                GemvNSub(Ws, V, n, i + 1, DSmall.Ptr, f, s);
                if (i > 0) {
                    Hm(i - 1, i) += small[i - 1];
                    Hm(i, i - 1) = Hm(i - 1, i);
                }
                Hm(i, i) += small[i];
                ++count;
            }
        }
    };

    std::vector<double> evec, theta, ritz_val(m), ritz_est(m);
    std::vector<uint32_t> order(m);
    auto retrieve_ritz = [&] {
        evec = H;
        if (!SymmetricEigen(m, evec, theta)) Fail(ME_NOT_CONVERGED, "Lanczos: projected eigenproblem did not converge");
        std::iota(order.begin(), order.end(), 0u);
        std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return std::abs(theta[a]) > std::abs(theta[b]); }); // LargestMagn
        for (uint32_t i = 0; i < m; ++i) {
            ritz_val[i] = theta[order[i]];
            ritz_est[i] = evec[size_t(m - 1) * m + order[i]];
        }
    };

    factorize_from(1);
    retrieve_ritz();
    uint32_t iter = 0, nconv = 0;
    for (; iter < max_restarts; ++iter) {
        nconv = 0;
        for (uint32_t i = 0; i < nev; ++i)
            if (std::abs(ritz_est[i]) * beta < tol * std::max(eps23, std::abs(ritz_val[i]))) ++nconv;
        if (nconv >= nev) break;
        if (cancelled && *cancelled) {
            out.Cancelled = true;
            break;
        }
        // nev_adjusted (HermEigsBase.h:178-202)
        uint32_t k = nev;
        for (uint32_t i = nev; i < m; ++i)
            if (std::abs(ritz_est[i]) < near0) ++k;
        k += std::min(nconv, (m - k) / 2);
        if (k == 1 && m >= 6) k = m / 2;
        else if (k == 1 && m > 2) k = 2;
        if (k > m - 1) k = m - 1;
        // Thick restart: V <- [V S_k, f/beta], H <- diag(theta_k) bordered by beta * (last row of S_k).
        std::vector<double> q(size_t(m) * k);
        for (uint32_t j = 0; j < k; ++j)
            for (uint32_t r = 0; r < m; ++r) q[r + size_t(j) * m] = evec[size_t(r) * m + order[j]];
        DeviceBuffer<double> &dq = DSmall;
        dq.Reserve(std::max<size_t>(size_t(m) * k, size_t(m + 2) * (m + 2)));
        ME_CUDA(cudaMemcpyAsync(dq.Ptr, q.data(), q.size() * sizeof(double), cudaMemcpyHostToDevice, s));
        TallGemm(Ws, V, n, m, dq.Ptr, m, k, V2, s);
        ME_CUDA(cudaMemcpyAsync(col(V2, k), col(V, m), n * sizeof(double), cudaMemcpyDeviceToDevice, s));
        ME_CUDA(cudaStreamSynchronize(s));
        std::swap(V, V2);
        std::fill(H.begin(), H.end(), 0.0);
        coupling.assign(k, 0.0);
        for (uint32_t j = 0; j < k; ++j) {
            Hm(j, j) = ritz_val[j];
            coupling[j] = beta * ritz_est[j];
        }
        arrow_k = k;
        factorize_from(k);
        retrieve_ritz();
    }
    out.Restarts = iter + 1;
    out.OpApplications = Ops;
    out.Converged = nconv >= nev;
    if (out.Converged) {
        // theta -> lambda = 1/theta + sigma, ascending (SymGEigsShiftSolver.h:170-176); X = V * S[:, wanted].
        std::vector<uint32_t> pick(nev);
        std::iota(pick.begin(), pick.end(), 0u);
        std::vector<double> lambda(nev);
        for (uint32_t i = 0; i < nev; ++i) lambda[i] = 1.0 / ritz_val[i] + Sigma;
        std::stable_sort(pick.begin(), pick.end(), [&](uint32_t a, uint32_t b) { return lambda[a] < lambda[b]; });
        out.Eigenvalues.resize(nev);
        std::vector<double> q(size_t(m) * nev);
        for (uint32_t j = 0; j < nev; ++j) {
            out.Eigenvalues[j] = lambda[pick[j]];
            for (uint32_t r = 0; r < m; ++r) q[r + size_t(j) * m] = evec[size_t(r) * m + order[pick[j]]];
        }
        DSmall.Reserve(std::max<size_t>(q.size(), size_t(m + 2) * (m + 2)));
        ME_CUDA(cudaMemcpyAsync(DSmall.Ptr, q.data(), q.size() * sizeof(double), cudaMemcpyHostToDevice, s));
        Vectors.Reserve(n * nev);
        TallGemm(Ws, V, n, m, DSmall.Ptr, m, nev, Vectors.Ptr, s);
    }
    ME_CUDA(cudaStreamSynchronize(s));
    Factor.CheckSolves();
    for (uint32_t i = 0; i < OpCalls; ++i) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, OpEvents[2 * i], OpEvents[2 * i + 1]) == cudaSuccess) out.OpSolveMs += ms;
    }
    for (auto e : OpEvents) cudaEventDestroy(e);
    OpEvents.clear();
    OpCalls = 0;
    out.KernelLaunches = (Fem.KernelLaunches - launches0) + (Factor.Stats.KernelLaunches - f_launches0) + Ws.Launches;
    return out;
}

// ------------------------------------------------------------------------------------------------ block form
// The same shift-invert Krylov iteration carried kLanczosBlock vectors at a time. One operator application is then a
// PANEL solve (SparseCholesky::Solve with width 8: one pass over the factor for the 8 columns), which is what makes it
// worth it on the device: the triangular solves are HBM-bound, so the bytes per Krylov vector drop ~5x. The structure is
// the block counterpart of what Compute does: M-inner product, full re-orthogonalisation of every new block against the
// whole basis (two classical Gram-Schmidt passes, as tall GEMMs), block M-orthonormalisation by Cholesky-QR applied
// twice, Rayleigh-Ritz on the projected matrix T = V^T M Op V (host, <= mcap x mcap), the reference's convergence test
// with the block residual ||R S_last|| in place of beta |s_last| (HermEigsBase.h:158-175), thick restart keeping the
// wanted Ritz vectors plus the residual block, theta -> 1/theta + sigma (SymGEigsShiftSolver.h:170-176).
namespace {
// In-place Cholesky G = L L^T of a b x b matrix (row-major, lower triangle used). False when not positive definite.
bool SmallCholesky(uint32_t b, std::vector<double> &g) {
    for (uint32_t j = 0; j < b; ++j) {
        double d = g[size_t(j) * b + j];
        for (uint32_t k = 0; k < j; ++k) d -= g[size_t(j) * b + k] * g[size_t(j) * b + k];
        if (!(d > 0.0) || !std::isfinite(d)) return false;
        const double l = std::sqrt(d);
        g[size_t(j) * b + j] = l;
        for (uint32_t i = j + 1; i < b; ++i) {
            double v = g[size_t(i) * b + j];
            for (uint32_t k = 0; k < j; ++k) v -= g[size_t(i) * b + k] * g[size_t(j) * b + k];
            g[size_t(i) * b + j] = v / l;
        }
        for (uint32_t c = j + 1; c < b; ++c) g[size_t(j) * b + c] = 0.0;
    }
    return true;
}
// X = L^-1 (row-major, lower) of a lower-triangular L.
std::vector<double> LowerInverse(uint32_t b, const std::vector<double> &l) {
    std::vector<double> x(size_t(b) * b, 0.0);
    for (uint32_t c = 0; c < b; ++c) {
        x[size_t(c) * b + c] = 1.0 / l[size_t(c) * b + c];
        for (uint32_t r = c + 1; r < b; ++r) {
            double v = 0;
            for (uint32_t k = c; k < r; ++k) v -= l[size_t(r) * b + k] * x[size_t(k) * b + c];
            x[size_t(r) * b + c] = v / l[size_t(r) * b + r];
        }
    }
    return x;
}
} // namespace

uint32_t ShiftInvertLanczos::BlockBasisSize(uint32_t nev) {
    uint32_t extra = std::max(96u, nev / 2); // a deeper basis than the reference's nev + 20: restarts (host eigensolves) halve, the extra columns cost HBM only
    if (const char *env = std::getenv("ME_LANCZOS_EXTRA")) extra = uint32_t(std::max(8, std::atoi(env)));
    return (nev + extra + kLanczosBlock - 1) / kLanczosBlock * kLanczosBlock;
}

void ShiftInvertLanczos::OpPanel(const double *x, double *y, uint32_t width) {
    auto s = Fem.Stream;
    const size_t n = Fem.N;
    if (OpEvents.size() < size_t(2) * (OpCalls + 1)) {
        cudaEvent_t a, b;
        ME_CUDA(cudaEventCreate(&a));
        ME_CUDA(cudaEventCreate(&b));
        OpEvents.push_back(a);
        OpEvents.push_back(b);
    }
    Tmp.Reserve(n * width);
    ME_CUDA(cudaEventRecord(OpEvents[2 * OpCalls], s));
    Fem.SpmvMPanel(x, Tmp.Ptr, width);
    Factor.Solve(Tmp.Ptr, y, width);
    ME_CUDA(cudaEventRecord(OpEvents[2 * OpCalls + 1], s));
    ++OpCalls;
    Ops += width;
}

LanczosOutcome ShiftInvertLanczos::ComputeBlock(uint32_t nev, double tol, uint32_t max_restarts, const volatile int *cancelled) {
    LanczosOutcome out;
    const size_t n = Fem.N;
    constexpr uint32_t b = kLanczosBlock;
    const uint32_t mcap = BlockBasisSize(nev);
    if (nev < 1 || size_t(mcap) + b > n) Fail(ME_BAD_ARG, "block Lanczos: basis of %u + %u columns does not fit n = %zu", mcap, b, n);
    ME_CUDA(cudaSetDevice(Fem.Device));
    auto s = Fem.Stream;
    const double eps = std::numeric_limits<double>::epsilon(), eps23 = std::pow(eps, 2.0 / 3.0);
    const uint32_t launches0 = Fem.KernelLaunches, f_launches0 = Factor.Stats.KernelLaunches;
    const uint32_t tcap = mcap + b;

    const bool prof = std::getenv("ME_PROFILE") != nullptr;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_enter = now();
    DeviceBuffer<double> Va, Vb, W, Wb, MW, H1, H2, Small;
    Va.Reserve(n * tcap), Vb.Reserve(n * tcap), W.Reserve(n * b), Wb.Reserve(n * b), MW.Reserve(n * b);
    H1.Reserve(size_t(tcap) * b), H2.Reserve(size_t(tcap) * b), Small.Reserve(std::max<size_t>(size_t(tcap) * tcap, 4 * b * b));
    double *V = Va.Ptr, *V2 = Vb.Ptr;
    auto col = [&](double *base, uint32_t j) { return base + size_t(j) * n; };
    std::vector<double> T(size_t(tcap) * tcap, 0.0), h1(size_t(tcap) * b), h2(size_t(tcap) * b), g(size_t(b) * b), small(size_t(b) * b);
    auto Tm = [&](uint32_t r, uint32_t c) -> double & { return T[size_t(r) * tcap + c]; };
    auto mass_product = [&](const double *x, double *y) {
        Fem.SpmvMPanel(x, y, b);
    };
    // One Cholesky-QR pass in the M inner product: dst = src * L^-T with src^T M src = L L^T. r (row-major b x b)
    // receives L^T. False when the block has lost rank.
    auto cholqr_pass = [&](const double *src, double *dst, std::vector<double> &r) {
        mass_product(src, MW.Ptr);
        Gram(Ws, src, n, b, MW.Ptr, b, Small.Ptr, b, s);
        ME_CUDA(cudaMemcpyAsync(small.data(), Small.Ptr, size_t(b) * b * sizeof(double), cudaMemcpyDeviceToHost, s));
        ME_CUDA(cudaStreamSynchronize(s));
        for (uint32_t i = 0; i < b; ++i)
            for (uint32_t j = 0; j < b; ++j) g[size_t(i) * b + j] = 0.5 * (small[i + size_t(j) * b] + small[j + size_t(i) * b]);
        if (!SmallCholesky(b, g)) return false;
        const std::vector<double> linv = LowerInverse(b, g);
        // Q = L^-T, column-major b x b: Q[i + j b] = Linv[j][i]
        for (uint32_t i = 0; i < b; ++i)
            for (uint32_t j = 0; j < b; ++j) small[i + size_t(j) * b] = linv[size_t(j) * b + i];
        ME_CUDA(cudaMemcpyAsync(Small.Ptr + size_t(2) * b * b, small.data(), size_t(b) * b * sizeof(double), cudaMemcpyHostToDevice, s));
        TallGemm(Ws, src, n, b, Small.Ptr + size_t(2) * b * b, b, b, dst, s);
        r.assign(size_t(b) * b, 0.0);
        for (uint32_t i = 0; i < b; ++i)
            for (uint32_t j = i; j < b; ++j) r[size_t(i) * b + j] = g[size_t(j) * b + i];
        return true;
    };
    // dst = M-orthonormal basis of span(src) (src is overwritten), R (row-major b x b upper) with src = dst R.
    std::vector<double> r1, r2, R(size_t(b) * b);
    auto orthonormalize = [&](double *src, double *scratch, double *dst) {
        if (!cholqr_pass(src, scratch, r1) || !cholqr_pass(scratch, dst, r2)) return false;
        for (uint32_t i = 0; i < b; ++i)
            for (uint32_t j = 0; j < b; ++j) {
                double v = 0;
                for (uint32_t k = 0; k < b; ++k) v += r2[size_t(i) * b + k] * r1[size_t(k) * b + j];
                R[size_t(i) * b + j] = v;
            }
        return true;
    };

    const double t_alloc = now();
    // Start block: Op applied to the reference's pseudo-random residual (SimpleRandom, seed 0 -> 1), b columns of it.
    FillSimpleRandom(Wb.Ptr, size_t(n) * b, s, Ws.Launches);
    OpPanel(Wb.Ptr, W.Ptr, b);
    if (!orthonormalize(W.Ptr, Wb.Ptr, col(V, 0))) Fail(ME_NOT_CONVERGED, "block Lanczos: the start block is rank deficient");

    // ME_PROFILE=1: synchronising section timers printed to stderr (diagnostics only; perturbs the overlap of host and device).
    double sec[5]{};
    if (prof) {
        cudaStreamSynchronize(s);
        fprintf(stderr, "[block lanczos] setup: allocations %.3f s, start block %.3f s\n", t_alloc - t_enter, now() - t_alloc);
    }
    double mark_t = now();
    auto mark = [&](int which) {
        if (!prof) return;
        cudaStreamSynchronize(s);
        const double t = now();
        sec[which] += t - mark_t;
        mark_t = t;
    };
    uint32_t m = 0; // expanded columns: T is m x m, the residual block sits in columns [m, m + b)
    std::vector<double> evec, theta, ritz_val, est, tail, q;
    std::vector<uint32_t> order;
    std::vector<QlRotation> rotations;
    DeviceBuffer<double> Zt;       // the Ritz vectors of the last decomposition, one per row (m x m)
    DeviceBuffer<QlRotation> Rot;
    DeviceBuffer<uint32_t> Pick;
    DeviceBuffer<double> TriU, TriV, TriH; // the reflectors of the tridiagonalisation (dense.h HouseholderBasis)
    std::vector<double> offdiag, reflector_h, evec_t;
    bool ritz_on_device = false;
    std::function<void(const std::vector<uint32_t> &)> pick_last; // Small <- picked vectors of the last decomposition
    uint32_t iter = 0, nconv = 0;
    bool broke = false;
    bool op_ready = false; // W already holds Op applied to the residual block (issued ahead of the host eigensolve below)
    for (;; ++iter) {
        while (m < mcap) {
            if (cancelled && *cancelled) {
                out.Cancelled = true;
                break;
            }
            const uint32_t cur = m + b;
            mark(4);
            if (!op_ready) OpPanel(col(V, m), W.Ptr, b);
            op_ready = false;
            mark(0);
            // Two Gram-Schmidt passes against the whole basis; the coefficients are the block column of T.
            mass_product(W.Ptr, MW.Ptr);
            Gram(Ws, V, n, cur, MW.Ptr, b, H1.Ptr, cur, s);
            TallGemm(Ws, V, n, cur, H1.Ptr, cur, b, W.Ptr, s, -1.0, 1.0);
            mass_product(W.Ptr, MW.Ptr);
            Gram(Ws, V, n, cur, MW.Ptr, b, H2.Ptr, cur, s);
            TallGemm(Ws, V, n, cur, H2.Ptr, cur, b, W.Ptr, s, -1.0, 1.0);
            ME_CUDA(cudaMemcpyAsync(h1.data(), H1.Ptr, size_t(cur) * b * sizeof(double), cudaMemcpyDeviceToHost, s));
            ME_CUDA(cudaMemcpyAsync(h2.data(), H2.Ptr, size_t(cur) * b * sizeof(double), cudaMemcpyDeviceToHost, s));
            mark(1);
            if (!orthonormalize(W.Ptr, Wb.Ptr, col(V, cur))) { // (synchronises: h1, h2 are on the host now)
                broke = true;
                break;
            }
            mark(2);
            for (uint32_t j = 0; j < b; ++j)
                for (uint32_t i = 0; i < cur; ++i) {
                    const double v = h1[i + size_t(j) * cur] + h2[i + size_t(j) * cur];
                    Tm(i, m + j) = v;
                    Tm(m + j, i) = v;
                }
            for (uint32_t i = 0; i < b; ++i) // the diagonal block was written twice above: make it exactly symmetric
                for (uint32_t j = i + 1; j < b; ++j) Tm(m + i, m + j) = Tm(m + j, m + i) = 0.5 * (Tm(m + i, m + j) + Tm(m + j, m + i));
            for (uint32_t i = 0; i < b; ++i)
                for (uint32_t j = 0; j < b; ++j) Tm(cur + i, m + j) = Tm(m + j, cur + i) = R[size_t(i) * b + j];
            m += b;
        }
        if (out.Cancelled || broke) break;
        // Rayleigh-Ritz on T[0:m, 0:m]. The device would idle through the host eigensolve (tens of milliseconds), and what it has
        // to do next if the iteration goes on is already known: a restart keeps the residual block as it is, and the first step
        // after it applies the operator to exactly that block. So the application is issued now, ahead of the decision.
        mark(4);
        if (iter < max_restarts) {
            OpPanel(col(V, m), W.Ptr, b);
            op_ready = true;
        }
        evec.assign(size_t(m) * m, 0.0);
        for (uint32_t i = 0; i < m; ++i)
            for (uint32_t j = 0; j < m; ++j) evec[size_t(i) * m + j] = 0.5 * (Tm(i, j) + Tm(j, i));
        // Tridiagonalisation and the QL iteration on the host; the accumulation of the QL rotations into the eigenvector matrix
        // (two thirds of the decomposition) on the device, where the matrix is wanted for the restart GEMM anyway. The host only
        // takes back the last b components of every vector (residual estimates, the border of the restarted T).
        const double t_ritz0 = now();
        ritz_on_device = m >= 96 && m <= kMaxDeviceRotationOrder && !std::getenv("ME_HOST_RITZ");
        // The orthogonal basis of the tridiagonal form is accumulated from the reflectors on the device as well (behind the operator
        // application already in the stream), while the host runs the QL iteration on the two diagonals.
        static const bool host_basis = std::getenv("ME_HOST_BASIS") != nullptr; // (A/B switch: the basis accumulated on host threads)
        const bool basis_on_device = ritz_on_device && !host_basis;
        if (basis_on_device) {
            HouseholderTridiagonal(m, evec, theta, offdiag, reflector_h);
            evec_t.resize(size_t(m) * m);
            for (uint32_t i = 1; i < m; ++i) // row i of the transposed copy: u_i / h_i, contiguous
                for (uint32_t k = 0; k < i; ++k) evec_t[size_t(i) * m + k] = evec[size_t(k) * m + i];
            const size_t mm = size_t(m) * m;
            TriU.Reserve(size_t(mcap) * mcap), TriV.Reserve(size_t(mcap) * mcap), TriH.Reserve(mcap), Zt.Reserve(size_t(mcap) * mcap);
            ME_CUDA(cudaMemcpyAsync(TriU.Ptr, evec.data(), mm * sizeof(double), cudaMemcpyHostToDevice, s));
            ME_CUDA(cudaMemcpyAsync(TriV.Ptr, evec_t.data(), mm * sizeof(double), cudaMemcpyHostToDevice, s));
            ME_CUDA(cudaMemcpyAsync(TriH.Ptr, reflector_h.data(), m * sizeof(double), cudaMemcpyHostToDevice, s));
            HouseholderBasis(TriU.Ptr, TriV.Ptr, TriH.Ptr, m, Zt.Ptr, s, Ws.Launches);
            if (!TridiagonalQl(m, theta, offdiag, rotations)) Fail(ME_NOT_CONVERGED, "block Lanczos: projected eigenproblem did not converge");
        } else if (!SymmetricEigenReduce(m, evec, theta, rotations)) {
            Fail(ME_NOT_CONVERGED, "block Lanczos: projected eigenproblem did not converge");
        }
        const double t_ritz1 = now();
        if (ritz_on_device) {
            Zt.Reserve(size_t(m) * m), Rot.Reserve((rotations.size() / 512 + 1) * 512), Pick.Reserve(m);
            if (!basis_on_device) ME_CUDA(cudaMemcpyAsync(Zt.Ptr, evec.data(), size_t(m) * m * sizeof(double), cudaMemcpyHostToDevice, s));
            ME_CUDA(cudaMemcpyAsync(Rot.Ptr, rotations.data(), rotations.size() * sizeof(QlRotation), cudaMemcpyHostToDevice, s));
            ApplyRotations(Zt.Ptr, m, Rot.Ptr, rotations.size(), s, Ws.Launches);
            tail.resize(size_t(m) * b);
            ME_CUDA(cudaMemcpy2DAsync(tail.data(), b * sizeof(double), Zt.Ptr + (m - b), m * sizeof(double), b * sizeof(double), m, cudaMemcpyDeviceToHost, s));
            ME_CUDA(cudaStreamSynchronize(s));
        } else {
            SymmetricEigenApplyHost(m, evec, rotations);
        }
        if (prof) fprintf(stderr, "[block lanczos] ritz m = %u: reduce %.1f ms (%zu rotations), apply %.1f ms (%s)\n", m, 1e3 * (t_ritz1 - t_ritz0), rotations.size(), 1e3 * (now() - t_ritz1), ritz_on_device ? "device" : "host");
        // component m - b + c of Ritz vector `vec`
        const auto tail_of = [&](uint32_t vec, uint32_t c) { return ritz_on_device ? tail[size_t(vec) * b + c] : evec[size_t(m - b + c) * m + vec]; };
        // Small <- the picked Ritz vectors as the columns of an m x count matrix
        const auto pick_vectors = [&](const std::vector<uint32_t> &vectors) {
            Small.Reserve(size_t(m) * vectors.size());
            if (ritz_on_device) {
                ME_CUDA(cudaMemcpyAsync(Pick.Ptr, vectors.data(), vectors.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
                GatherRows(Zt.Ptr, m, Pick.Ptr, uint32_t(vectors.size()), Small.Ptr, s, Ws.Launches);
            } else {
                q.resize(size_t(m) * vectors.size());
                for (size_t j = 0; j < vectors.size(); ++j)
                    for (uint32_t r = 0; r < m; ++r) q[r + j * m] = evec[size_t(r) * m + vectors[j]];
                ME_CUDA(cudaMemcpyAsync(Small.Ptr, q.data(), q.size() * sizeof(double), cudaMemcpyHostToDevice, s));
            }
        };
        pick_last = pick_vectors;
        order.resize(m);
        std::iota(order.begin(), order.end(), 0u);
        std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t c) { return std::abs(theta[a]) > std::abs(theta[c]); });
        ritz_val.resize(m), est.resize(m);
        for (uint32_t i = 0; i < m; ++i) {
            ritz_val[i] = theta[order[i]];
            double sq = 0;
            for (uint32_t r = 0; r < b; ++r) {
                double v = 0;
                for (uint32_t c = 0; c < b; ++c) v += Tm(m + r, m - b + c) * tail_of(order[i], c);
                sq += v * v;
            }
            est[i] = std::sqrt(sq);
        }
        nconv = 0;
        for (uint32_t i = 0; i < nev; ++i)
            if (est[i] < tol * std::max(eps23, std::abs(ritz_val[i]))) ++nconv;
        mark(3);
        if (nconv >= nev || iter >= max_restarts) break;
        // Thick restart: keep k Ritz vectors (the reference's rule, HermEigsBase.h:178-202, rounded so that whole
        // blocks fill the basis again) and the residual block.
        uint32_t k = nev + std::min(nconv, (m - nev) / 2);
        k = std::min(k, m - b);
        k = mcap - (mcap - k) / b * b;
        pick_vectors(std::vector<uint32_t>(order.begin(), order.begin() + k));
        TallGemm(Ws, V, n, m, Small.Ptr, m, k, V2, s);
        ME_CUDA(cudaMemcpyAsync(col(V2, k), col(V, m), n * b * sizeof(double), cudaMemcpyDeviceToDevice, s));
        ME_CUDA(cudaStreamSynchronize(s));
        std::swap(V, V2);
        std::vector<double> border(size_t(b) * k);
        for (uint32_t r = 0; r < b; ++r)
            for (uint32_t j = 0; j < k; ++j) {
                double v = 0;
                for (uint32_t c = 0; c < b; ++c) v += Tm(m + r, m - b + c) * tail_of(order[j], c);
                border[size_t(r) * k + j] = v;
            }
        std::fill(T.begin(), T.end(), 0.0);
        for (uint32_t j = 0; j < k; ++j) {
            Tm(j, j) = ritz_val[j];
            for (uint32_t r = 0; r < b; ++r) Tm(k + r, j) = Tm(j, k + r) = border[size_t(r) * k + j];
        }
        m = k;
    }
    mark(4);
    const double t_loop_end = now();
    if (prof) fprintf(stderr, "[block lanczos] op %.3f s, reorth %.3f s, cholqr %.3f s, ritz (host) %.3f s, restart+other %.3f s, %u panel ops, %u restarts\n", sec[0], sec[1], sec[2], sec[3], sec[4], OpCalls, iter);
    out.Restarts = iter + 1;
    out.OpApplications = Ops;
    out.Converged = !out.Cancelled && !broke && nconv >= nev;
    if (out.Converged) {
        std::vector<uint32_t> pick(nev);
        std::iota(pick.begin(), pick.end(), 0u);
        std::vector<double> lambda(nev);
        for (uint32_t i = 0; i < nev; ++i) lambda[i] = 1.0 / ritz_val[i] + Sigma;
        std::stable_sort(pick.begin(), pick.end(), [&](uint32_t a, uint32_t c) { return lambda[a] < lambda[c]; });
        out.Eigenvalues.resize(nev);
        std::vector<uint32_t> wanted(nev);
        for (uint32_t j = 0; j < nev; ++j) {
            out.Eigenvalues[j] = lambda[pick[j]];
            wanted[j] = order[pick[j]];
        }
        pick_last(wanted);
        // The Ritz vectors are written into the spare basis buffer, which then becomes `Vectors`: no 0.9 GB allocation at the
        // end of a solve whose 20 GB of buffers are all still alive (the stream-ordered pool sometimes took 0.2-0.7 s over it).
        TallGemm(Ws, V, n, m, Small.Ptr, m, nev, V2, s);
        DeviceBuffer<double> &spare = V2 == Va.Ptr ? Va : Vb;
        std::swap(Vectors.Ptr, spare.Ptr);
        std::swap(Vectors.Capacity, spare.Capacity);
    }
    double t_tail[4]{};
    t_tail[0] = now();
    ME_CUDA(cudaStreamSynchronize(s));
    t_tail[1] = now();
    Factor.CheckSolves();
    t_tail[2] = now();
    for (uint32_t i = 0; i < OpCalls; ++i) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, OpEvents[2 * i], OpEvents[2 * i + 1]) == cudaSuccess) out.OpSolveMs += ms;
    }
    for (auto e : OpEvents) cudaEventDestroy(e);
    OpEvents.clear();
    OpCalls = 0;
    t_tail[3] = now();
    if (prof) fprintf(stderr, "[block lanczos] tail: ritz vectors launch %.3f s, sync %.3f s, check %.3f s, events %.3f s\n", t_tail[0] - t_loop_end, t_tail[1] - t_tail[0], t_tail[2] - t_tail[1], t_tail[3] - t_tail[2]);
    out.RankLost = broke;
    out.KernelLaunches = (Fem.KernelLaunches - launches0) + (Factor.Stats.KernelLaunches - f_launches0) + Ws.Launches;
    if (prof) fprintf(stderr, "[block lanczos] extraction + checks %.3f s\n", now() - t_loop_end);
    return out;
}

} // namespace me
