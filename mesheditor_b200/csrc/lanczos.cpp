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
#include <cmath>
#include <limits>
#include <numeric>

namespace me {

bool SymmetricEigen(uint32_t n, std::vector<double> &a, std::vector<double> &d) {
    d.assign(n, 0.0);
    if (n == 0) return true;
    std::vector<double> e(n, 0.0), scratch(n, 0.0);
    auto A = [&](uint32_t r, uint32_t c) -> double & { return a[size_t(r) * n + c]; };
    // Householder reduction to tridiagonal form, accumulating the transformation.
    for (uint32_t i = n - 1; i >= 1; --i) {
        const uint32_t l = i - 1;
        double h = 0, scale = 0;
        if (l > 0) {
            for (uint32_t k = 0; k <= l; ++k) scale += std::abs(A(i, k));
            if (scale == 0.0) e[i] = A(i, l);
            else {
                for (uint32_t k = 0; k <= l; ++k) {
                    A(i, k) /= scale;
                    h += A(i, k) * A(i, k);
                }
                double f = A(i, l);
                double g = f >= 0 ? -std::sqrt(h) : std::sqrt(h);
                e[i] = scale * g;
                h -= f * g;
                A(i, l) = f - g;
                f = 0;
                // e = A u / h with A symmetric and only its lower triangle stored: both sweeps run along contiguous rows.
                for (uint32_t j = 0; j <= l; ++j) e[j] = 0;
                const double *u = &a[size_t(i) * n];
                for (uint32_t j = 0; j <= l; ++j) {
                    const double *row = &a[size_t(j) * n];
                    const double uj = u[j];
                    double dot = 0;
                    for (uint32_t k = 0; k < j; ++k) {
                        dot += row[k] * u[k];
                        e[k] += row[k] * uj;
                    }
                    e[j] += dot + row[j] * uj;
                }
                for (uint32_t j = 0; j <= l; ++j) {
                    A(j, i) = A(i, j) / h;
                    e[j] /= h;
                    f += e[j] * A(i, j);
                }
                const double hh = f / (h + h);
                for (uint32_t j = 0; j <= l; ++j) {
                    f = A(i, j);
                    e[j] = g = e[j] - hh * f;
                    for (uint32_t k = 0; k <= j; ++k) A(j, k) -= f * e[k] + g * A(i, k);
                }
            }
        } else e[i] = A(i, l);
        d[i] = h;
    }
    d[0] = 0;
    e[0] = 0;
    for (uint32_t i = 0; i < n; ++i) {
        if (d[i] != 0.0 && i > 0) {
            // g = row_i * Q[0..i, 0..i), then Q[0..i, 0..i) -= Q[0..i, i] * g: contiguous row sweeps.
            std::vector<double> &g = scratch;
            std::fill(g.begin(), g.begin() + i, 0.0);
            for (uint32_t k = 0; k < i; ++k) {
                const double aik = A(i, k);
                const double *row = &a[size_t(k) * n];
                for (uint32_t j = 0; j < i; ++j) g[j] += aik * row[j];
            }
            for (uint32_t k = 0; k < i; ++k) {
                const double aki = A(k, i);
                double *row = &a[size_t(k) * n];
                for (uint32_t j = 0; j < i; ++j) row[j] -= g[j] * aki;
            }
        }
        d[i] = A(i, i);
        A(i, i) = 1;
        for (uint32_t j = 0; j < i; ++j) A(j, i) = A(i, j) = 0;
    }
    // Implicit QL on the tridiagonal matrix. The rotations mix two eigenvector columns at a time: work on the transpose
    // so that they are contiguous rows (this loop is ~3 n^3 flops and sits between every two restarts).
    for (uint32_t r = 0; r < n; ++r)
        for (uint32_t c = r + 1; c < n; ++c) std::swap(a[size_t(r) * n + c], a[size_t(c) * n + r]);
    for (uint32_t i = 1; i < n; ++i) e[i - 1] = e[i];
    e[n - 1] = 0;
    const double eps = std::numeric_limits<double>::epsilon();
    for (uint32_t l = 0; l < n; ++l) {
        uint32_t iter = 0, m;
        do {
            for (m = l; m + 1 < n; ++m) {
                const double dd = std::abs(d[m]) + std::abs(d[m + 1]);
                if (std::abs(e[m]) <= eps * dd) break;
            }
            if (m != l) {
                if (++iter > 200) return false;
                double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
                double r = std::hypot(g, 1.0);
                g = d[m] - d[l] + e[l] / (g + std::copysign(r, g));
                double s = 1, c = 1, p = 0;
                int64_t i;
                for (i = int64_t(m) - 1; i >= int64_t(l); --i) {
                    double f = s * e[i];
                    const double b = c * e[i];
                    e[i + 1] = r = std::hypot(f, g);
                    if (r == 0.0) {
                        d[i + 1] -= p;
                        e[m] = 0;
                        break;
                    }
                    s = f / r;
                    c = g / r;
                    g = d[i + 1] - p;
                    r = (d[i] - g) * s + 2.0 * c * b;
                    d[i + 1] = g + (p = s * r);
                    g = c * r - b;
                    double *zi = &a[size_t(i) * n], *zi1 = &a[size_t(i + 1) * n];
                    for (uint32_t k = 0; k < n; ++k) {
                        const double fk = zi1[k];
                        zi1[k] = s * zi[k] + c * fk;
                        zi[k] = c * zi[k] - s * fk;
                    }
                }
                if (r == 0.0 && i >= int64_t(l)) continue;
                d[l] -= p;
                e[l] = g;
                e[m] = 0;
            }
        } while (m != l);
    }
    for (uint32_t r = 0; r < n; ++r)
        for (uint32_t c = r + 1; c < n; ++c) std::swap(a[size_t(r) * n + c], a[size_t(c) * n + r]);
    return true;
}

void ShiftInvertLanczos::Op(const double *x, double *y) {
    auto s = Fem.Stream;
    if (OpEvents.size() < size_t(2) * (Ops + 1)) {
        cudaEvent_t a, b;
        ME_CUDA(cudaEventCreate(&a));
        ME_CUDA(cudaEventCreate(&b));
        OpEvents.push_back(a);
        OpEvents.push_back(b);
    }
    ME_CUDA(cudaEventRecord(OpEvents[2 * Ops], s));
    Fem.SpmvM(x, Tmp.Ptr);
    Factor.Solve(Tmp.Ptr, y, 1);
    ME_CUDA(cudaEventRecord(OpEvents[2 * Ops + 1], s));
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
    {
        std::vector<double> r0(n);
        uint64_t x = 1;
        for (size_t i = 0; i < n; ++i) {
            x = (x * 16807ull) % 2147483647ull;
            r0[i] = double(x) / 2147483647.0 - 0.5;
        }
        ME_CUDA(cudaMemcpyAsync(W.Ptr, r0.data(), n * sizeof(double), cudaMemcpyHostToDevice, s));
        ME_CUDA(cudaStreamSynchronize(s));
    }
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
    for (uint32_t i = 0; i < Ops; ++i) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, OpEvents[2 * i], OpEvents[2 * i + 1]) == cudaSuccess) out.OpSolveMs += ms;
    }
    for (auto e : OpEvents) cudaEventDestroy(e);
    OpEvents.clear();
    out.KernelLaunches = (Fem.KernelLaunches - launches0) + (Factor.Stats.KernelLaunches - f_launches0) + Ws.Launches;
    return out;
}

} // namespace me
