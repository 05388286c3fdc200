// Host probe: times SymmetricEigenReduce / SymmetricEigenApplyHost (hosteig.cpp) on a random symmetric matrix.
// Build: g++ -O3 -std=c++20 -Imesheditor_b200/csrc -Iinclude -I/usr/local/cuda/include scripts/probes/eig_harness.cpp mesheditor_b200/csrc/hosteig.cpp -o /tmp/eig_harness -lpthread
#include "lanczos.h"
#include <chrono>
#include <cstdio>
#include <random>
using namespace me;
int main(int argc, char **argv) {
    const uint32_t n = argc > 1 ? atoi(argv[1]) : 328;
    std::mt19937_64 rng(7);
    std::normal_distribution<double> g;
    std::vector<double> a0(size_t(n) * n);
    for (uint32_t i = 0; i < n; ++i)
        for (uint32_t j = 0; j <= i; ++j) a0[size_t(i) * n + j] = a0[size_t(j) * n + i] = g(rng) * (i == j ? 10.0 : 1.0);
    double best_reduce = 1e9, best_total = 1e9;
    std::vector<double> a, d;
    for (int rep = 0; rep < 7; ++rep) {
        a = a0;
        std::vector<QlRotation> rot;
        auto t0 = std::chrono::steady_clock::now();
        bool ok = SymmetricEigenReduce(n, a, d, rot);
        auto t1 = std::chrono::steady_clock::now();
        SymmetricEigenApplyHost(n, a, rot);
        auto t2 = std::chrono::steady_clock::now();
        if (!ok) return 1;
        best_reduce = std::min(best_reduce, std::chrono::duration<double, std::milli>(t1 - t0).count());
        best_total = std::min(best_total, std::chrono::duration<double, std::milli>(t2 - t0).count());
    }
    // residual: max |A z - lambda z| over vectors
    double worst = 0;
    for (uint32_t c = 0; c < n; ++c) {
        for (uint32_t r = 0; r < n; ++r) {
            double s = 0;
            for (uint32_t k = 0; k < n; ++k) s += a0[size_t(r) * n + k] * a[size_t(k) * n + c];
            worst = std::max(worst, std::abs(s - d[c] * a[size_t(r) * n + c]));
        }
    }
    printf("n %u reduce %.2f ms total %.2f ms residual %.2e\n", n, best_reduce, best_total, worst);
    return 0;
}
