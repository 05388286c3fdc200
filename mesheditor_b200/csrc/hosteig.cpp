// Dense symmetric eigensolver of the projected (Rayleigh-Ritz) problems, on the host: Householder tridiagonalisation and
// implicit QL with accumulated vectors (the job Spectra hands to its TridiagEigen, lib/spectra/include/Spectra/LinAlg/
// TridiagEigen.h, after HermEigsBase::retrieve_ritzpair). a: n x n row-major symmetric in, eigenvectors out (a[r * n + c] =
// component r of vector c); d: eigenvalues, in the order the QL iteration leaves them.
//
// It sits on the critical path between two restarts of the block Lanczos iteration (m = 320 for the 1M-tet solve: 22 ms a
// call, seven calls a solve, while the GPU idles). Two thirds of it is the accumulation of the QL rotations into the
// eigenvector matrix, and the rotations themselves depend only on the tridiagonal entries: the QL iteration is run on d and
// e alone, recording every rotation, and the whole history is then applied to the eigenvectors by a few threads, each on its
// own column range, with a single fork and join. The arithmetic of each entry is the sequential algorithm's, operation for
// operation (only disjoint ranges are handed out), so the result does not depend on the thread count.
//
// The block Lanczos iteration goes one step further (lanczos.cpp): it takes the two halves apart - SymmetricEigenReduce leaves
// the transposed Householder basis and the rotation history - and applies the history on the device (dense.cu ApplyRotations),
// where the eigenvector matrix is needed anyway for the restart GEMM.
#include "lanczos.h"

#include <algorithm>
#include <cmath>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <thread>
#include <vector>

namespace me {
namespace {

// [begin, end) of worker w's share of `count` items.
inline void Share(uint32_t count, uint32_t w, uint32_t workers, uint32_t &begin, uint32_t &end) {
    begin = uint32_t(uint64_t(count) * w / workers);
    end = uint32_t(uint64_t(count) * (w + 1) / workers);
}

uint32_t TeamSize(uint32_t n) {
    if (n < 96) return 1; // a sweep over a small matrix is shorter than a barrier
    static const uint32_t configured = [] {
        if (const char *env = std::getenv("ME_HOST_THREADS")) return uint32_t(std::max(1, std::atoi(env)));
        return std::min(8u, std::max(1u, std::thread::hardware_concurrency() / 2));
    }();
    return configured;
}

// body(worker, workers) on `workers` threads (the caller is worker 0): one fork and one join.
template<typename F>
void RunTeam(uint32_t workers, F &&body) {
    std::vector<std::thread> pool;
    for (uint32_t w = 1; w < workers; ++w) pool.emplace_back([&body, w, workers] { body(w, workers); });
    body(0u, workers);
    for (auto &t : pool) t.join();
}
// Barrier of a team that meets every few microseconds: the waiters spin (a futex round trip costs more than the work between two
// of these barriers).
struct SpinBarrier {
    explicit SpinBarrier(uint32_t team) : Team(team) {}
    void Wait() {
        const uint32_t phase = Phase.load(std::memory_order_acquire);
        if (Arrived.fetch_add(1, std::memory_order_acq_rel) + 1 == Team) {
            Arrived.store(0, std::memory_order_relaxed);
            Phase.store(phase + 1, std::memory_order_release);
        } else {
            // (Spins, but not forever: several solves may be in flight on one host, each with a team of its own, and a teammate that
            // has lost its core must get one back - with nothing but pauses here a batch of concurrent solves ran half as fast.)
            for (uint32_t spins = 0; Phase.load(std::memory_order_acquire) == phase; ++spins) {
                if (spins < 256) {
#if defined(__x86_64__) || defined(__i386__)
                    __builtin_ia32_pause();
#endif
                } else {
                    std::this_thread::yield();
                }
            }
        }
    }
    const uint32_t Team;
    alignas(64) std::atomic<uint32_t> Arrived{0};
    alignas(64) std::atomic<uint32_t> Phase{0};
};

// The Householder reduction of SymmetricEigenReduce on a team in lockstep: the two O(l^2) sweeps of a step (e = A u over the stored
// lower triangle, and the rank-2 update) are shared out by rows in stretches of equal area, the O(l) parts stay with worker 0,
// five spin barriers a step. The partial products of e are summed in worker order, so the result is a function of the team size
// (fixed per machine: TeamSize), not of the timing.
void HouseholderReduceTeam(uint32_t n, std::vector<double> &a, std::vector<double> &d, std::vector<double> &e, uint32_t workers) {
    auto A = [&](uint32_t r, uint32_t c) -> double & { return a[size_t(r) * n + c]; };
    SpinBarrier barrier(workers);
    std::vector<std::vector<double>> partial(workers, std::vector<double>(n, 0.0));
    struct Shared {
        bool Active{false};
        double H{0};
    } shared;
    RunTeam(workers, [&](uint32_t w, uint32_t count) {
        double *mine = partial[w].data();
        for (uint32_t i = n - 1; i >= 1; --i) {
            const uint32_t l = i - 1, rows = l + 1;
            if (w == 0) {
                double h = 0, scale = 0;
                shared.Active = false;
                if (l > 0) {
                    for (uint32_t k = 0; k <= l; ++k) scale += std::abs(A(i, k));
                    if (scale == 0.0) e[i] = A(i, l);
                    else {
                        for (uint32_t k = 0; k <= l; ++k) {
                            A(i, k) /= scale;
                            h += A(i, k) * A(i, k);
                        }
                        const double f = A(i, l);
                        const double g = f >= 0 ? -std::sqrt(h) : std::sqrt(h);
                        e[i] = scale * g;
                        h -= f * g;
                        A(i, l) = f - g;
                        shared.Active = true;
                    }
                } else e[i] = A(i, l);
                shared.H = h;
                d[i] = h;
            }
            barrier.Wait();
            if (!shared.Active) {
                barrier.Wait(); // (worker 0 must not rewrite `shared` before everyone has read it)
                continue;
            }
            const double h = shared.H;
            const double *u = &a[size_t(i) * n];
            // rows [j0, j1) of equal area: row j costs j
            const uint32_t j0 = uint32_t(std::sqrt(double(w) / count) * rows), j1 = w + 1 == count ? rows : uint32_t(std::sqrt(double(w + 1) / count) * rows);
            std::fill(mine, mine + rows, 0.0);
            for (uint32_t j = j0; j < j1; ++j) {
                const double *row = &a[size_t(j) * n];
                const double uj = u[j];
                double dot = 0;
                for (uint32_t k = 0; k < j; ++k) {
                    dot += row[k] * u[k];
                    mine[k] += row[k] * uj;
                }
                mine[j] += dot + row[j] * uj;
            }
            barrier.Wait();
            {   // e[k] = sum of the partial products, worker by worker, over this worker's share of k
                uint32_t k0, k1;
                Share(rows, w, count, k0, k1);
                for (uint32_t k = k0; k < k1; ++k) {
                    double sum = 0;
                    for (uint32_t t = 0; t < count; ++t) sum += partial[t][k];
                    e[k] = sum;
                }
            }
            barrier.Wait();
            if (w == 0) {
                double f = 0;
                for (uint32_t j = 0; j <= l; ++j) {
                    A(j, i) = A(i, j) / h;
                    e[j] /= h;
                    f += e[j] * A(i, j);
                }
                const double hh = f / (h + h);
                for (uint32_t j = 0; j <= l; ++j) e[j] -= hh * A(i, j);
            }
            barrier.Wait();
            for (uint32_t j = j0; j < j1; ++j) {
                const double fj = u[j], gj = e[j];
                double *row = &a[size_t(j) * n];
                for (uint32_t k = 0; k <= j; ++k) row[k] -= fj * e[k] + gj * u[k];
            }
            barrier.Wait();
        }
    });
}
} // namespace

void HouseholderTridiagonal(uint32_t n, std::vector<double> &a, std::vector<double> &d, std::vector<double> &e, std::vector<double> &h_out) {
    d.assign(n, 0.0);
    e.assign(n, 0.0);
    h_out.assign(n, 0.0);
    if (n == 0) return;
    auto A = [&](uint32_t r, uint32_t c) -> double & { return a[size_t(r) * n + c]; };
    const uint32_t workers = TeamSize(n);
    // Householder reduction to tridiagonal form. Its steps are tens of microseconds of O(n^2) work each: a team shares them out only
    // behind spin barriers (HouseholderReduceTeam; with fork-join or futex barriers the threads gained less than the barriers cost).
    static const bool team_reduce = std::getenv("ME_HOST_TRIDIAG_SERIAL") == nullptr;
    static const bool timing = std::getenv("ME_EIG_TIMING") != nullptr;
    const auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_begin = now();
    // (measured on the 16-core box at order 328: 5.0 ms on one thread, 2.3 on four, 2.2 on eight; below order ~256 a step is shorter
    // than its five barriers are worth)
    const uint32_t team = std::min(workers, 4u);
    const bool on_team = team_reduce && team > 1 && n >= 256;
    if (on_team) HouseholderReduceTeam(n, a, d, e, team);
    for (uint32_t i = n - 1; i >= 1 && !on_team; --i) {
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
                for (uint32_t j = 0; j <= l; ++j) e[j] -= hh * A(i, j);
                {
                    const double *ui = &a[size_t(i) * n];
                    for (uint32_t j = 0; j <= l; ++j) {
                        const double fj = ui[j], gj = e[j];
                        double *row = &a[size_t(j) * n];
                        for (uint32_t k = 0; k <= j; ++k) row[k] -= fj * e[k] + gj * ui[k];
                    }
                }
            }
        } else e[i] = A(i, l);
        d[i] = h;
    }
    d[0] = 0;
    e[0] = 0;
    // Step i reflects with I - u_i u_i^T / h_i (h_i = 0: no reflection): u_i is row i of `a` left of the diagonal, u_i / h_i column i above it.
    h_out = d;
    for (uint32_t i = 0; i < n; ++i) d[i] = A(i, i);
    for (uint32_t i = 1; i < n; ++i) e[i - 1] = e[i];
    e[n - 1] = 0;
    if (timing) fprintf(stderr, "[me] host eigensolver n = %u: tridiagonalisation %.2f ms\n", n, now() - t_begin);
}

void HouseholderBasisHost(uint32_t n, std::vector<double> &a, const std::vector<double> &h) {
    if (n == 0) return;
    const uint32_t workers = TeamSize(n);
    {
        // The orthogonal basis Q = H_1 H_2 .. accumulated in a matrix of its own, starting from the identity: step i reads the
        // i-th Householder vector (row i of `a` left of the diagonal, and its scaled copy in column i) and updates Q's leading
        // i x i block, g = u_i Q then Q -= (u_i / h) g. A worker owns a fixed range of Q's columns for the whole accumulation - its
        // part of g and of the update involve nobody else's - so the workers never meet until the end.
        std::vector<double> q(size_t(n) * n, 0.0);
        for (uint32_t i = 0; i < n; ++i) q[size_t(i) * n + i] = 1.0;
        RunTeam(workers, [&](uint32_t w, uint32_t count) {
            uint32_t c0, c1;
            Share(n, w, count, c0, c1);
            std::vector<double> g(n, 0.0);
            for (uint32_t i = 1; i < n; ++i) {
                const uint32_t j1 = std::min(c1, i);
                if (h[i] == 0.0 || c0 >= j1) continue;
                std::fill(g.begin() + c0, g.begin() + j1, 0.0);
                for (uint32_t k = 0; k < i; ++k) {
                    const double aik = a[size_t(i) * n + k];
                    const double *row = &q[size_t(k) * n];
                    for (uint32_t j = c0; j < j1; ++j) g[j] += aik * row[j];
                }
                for (uint32_t k = 0; k < i; ++k) {
                    const double aki = a[size_t(k) * n + i];
                    double *row = &q[size_t(k) * n];
                    for (uint32_t j = c0; j < j1; ++j) row[j] -= g[j] * aki;
                }
            }
        });
        a.swap(q);
    }
    // The rotations of the QL iteration mix two eigenvector columns at a time: work on the transpose so that they are contiguous rows.
    for (uint32_t r = 0; r < n; ++r)
        for (uint32_t c = r + 1; c < n; ++c) std::swap(a[size_t(r) * n + c], a[size_t(c) * n + r]);
}

bool SymmetricEigenReduce(uint32_t n, std::vector<double> &a, std::vector<double> &d, std::vector<QlRotation> &rotations) {
    std::vector<double> e, h;
    HouseholderTridiagonal(n, a, d, e, h);
    HouseholderBasisHost(n, a, h);
    return TridiagonalQl(n, d, e, rotations);
}

// Implicit QL on the tridiagonal matrix with diagonal d and off-diagonal e (e[i] couples i and i + 1, e[n - 1] = 0). The rotations
// depend only on d and e: every one is recorded ({C, S, Row} mixes rows Row and Row + 1 of the transposed eigenvector matrix) and
// applied elsewhere (SymmetricEigenApplyHost, dense.h ApplyRotations: ~3 n^3 flops).
bool TridiagonalQl(uint32_t n, std::vector<double> &d, std::vector<double> &e, std::vector<QlRotation> &rotations) {
    rotations.clear();
    if (n == 0) return true;
    const double eps = std::numeric_limits<double>::epsilon();
    rotations.reserve(size_t(n) * n + size_t(n) * n / 4);
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
                    rotations.push_back({c, s, uint32_t(i), 0u});
                }
                if (r == 0.0 && i >= int64_t(l)) continue;
                d[l] -= p;
                e[l] = g;
                e[m] = 0;
            }
        } while (m != l);
    }
    return true;
}

void SymmetricEigenApplyHost(uint32_t n, std::vector<double> &a, const std::vector<QlRotation> &rotations) {
    if (n == 0) return;
    // The rotation history, applied in order to every worker's own columns.
    const auto apply = [&](uint32_t w, uint32_t workers) {
        uint32_t k0, k1;
        Share(n, w, workers, k0, k1);
        if (k0 == k1) return;
        for (const QlRotation &q : rotations) {
            double *zi = &a[size_t(q.Row) * n], *zi1 = zi + n;
            for (uint32_t k = k0; k < k1; ++k) {
                const double fk = zi1[k];
                zi1[k] = q.S * zi[k] + q.C * fk;
                zi[k] = q.C * zi[k] - q.S * fk;
            }
        }
    };
    const uint32_t workers = TeamSize(n);
    std::vector<std::thread> threads;
    for (uint32_t w = 1; w < workers; ++w) threads.emplace_back(apply, w, workers);
    apply(0, workers);
    for (auto &t : threads) t.join();
    for (uint32_t r = 0; r < n; ++r)
        for (uint32_t c = r + 1; c < n; ++c) std::swap(a[size_t(r) * n + c], a[size_t(c) * n + r]);
}

bool SymmetricEigen(uint32_t n, std::vector<double> &a, std::vector<double> &d) {
    std::vector<QlRotation> rotations;
    if (!SymmetricEigenReduce(n, a, d, rotations)) return false;
    SymmetricEigenApplyHost(n, a, rotations);
    return true;
}

} // namespace me
