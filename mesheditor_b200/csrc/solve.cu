// modal::mesh2modes on the device: the pipeline behind me_modal_solve, the host-only post-processing
// (PostprocessModes / RescaleModes / ComputeMassProperties) and the C ABI of the analysis path.
// Reference: src/audio/mesh2modes.cpp:605-658 (mesh2modes), :441-512 (ComputeModes), :515-603, :73-126.
#include "cholesky.h"
#include "common.h"
#include "result.h"
#include "fem.h"
#include "lanczos.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <unordered_map>

namespace me {
namespace {
double Seconds() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct Config {
    float MinModeFreq{20}, MaxModeFreq{16000};
    uint32_t NumModes{30}, NumFemModes{45};
    double Tolerance{1e-8}, WarmTolerance{1e-4};
    uint32_t MaxRestarts{100};
    bool HasFundamental{false};
    float FundamentalFreq{0};
    uint32_t ElementOrder{2};
    int Device{0};
};
Config FromC(const MeSolverConfig *c) {
    Config out;
    if (!c) return out;
    out.MinModeFreq = c->min_mode_freq, out.MaxModeFreq = c->max_mode_freq;
    out.NumModes = c->num_modes, out.NumFemModes = c->num_fem_modes;
    out.Tolerance = c->tolerance, out.WarmTolerance = c->warm_tolerance;
    out.MaxRestarts = c->max_restarts;
    out.HasFundamental = c->has_fundamental_freq != 0, out.FundamentalFreq = c->fundamental_freq;
    out.ElementOrder = c->element_order ? c->element_order : 2;
    out.Device = c->device;
    return out;
}
Material FromC(const MeMaterial *m) {
    if (!m) Fail(ME_BAD_ARG, "null material");
    return {m->density, m->young_modulus, m->poisson_ratio, m->alpha, m->beta};
}


// modal::PostprocessModes (mesh2modes.cpp:515-588). shapes: [point][eigenpair][3].
Modes Postprocess(const std::vector<double> &eigenvalues, const std::vector<float> &shapes, uint32_t n_points, float shape_scale, const Material &material, const Config &config,
                  std::vector<float> positions) {
    const uint32_t fem_n_modes = uint32_t(eigenvalues.size());
    std::vector<float> mode_freqs(fem_n_modes), mode_t60s(fem_n_modes);
    std::vector<double> omega_undamped(fem_n_modes);
    const double lambda_eps = std::pow(2 * M_PI * config.MinModeFreq, 2) * 1e-10;
    for (uint32_t mode = 0; mode < fem_n_modes; ++mode) omega_undamped[mode] = eigenvalues[mode] > lambda_eps ? std::sqrt(eigenvalues[mode]) : 0;
    const auto c_from_omega = [&](double omega) { return material.Alpha + material.Beta * (omega * omega); };
    const auto damped_hz = [&](double omega, double c) {
        const double omega_d_sq = omega * omega - 0.25 * c * c;
        return omega_d_sq > 0 ? std::sqrt(omega_d_sq) / (2 * M_PI) : 0;
    };
    uint32_t lowest = fem_n_modes;
    float lowest_freq_orig = 0;
    for (uint32_t mode = 0; mode < fem_n_modes; ++mode) {
        const double omega = omega_undamped[mode];
        if (omega <= 0) {
            mode_freqs[mode] = mode_t60s[mode] = 0.f;
            continue;
        }
        mode_freqs[mode] = float(damped_hz(omega, c_from_omega(omega)));
        if (lowest == fem_n_modes && mode_freqs[mode] >= config.MinModeFreq) {
            lowest = mode;
            lowest_freq_orig = mode_freqs[mode];
        }
    }
    if (lowest == fem_n_modes) return {};
    static const double ln_1000 = std::log(1000);
    const float freq_scale = config.HasFundamental ? config.FundamentalFreq / lowest_freq_orig : 1.f;
    for (uint32_t mode = lowest; mode < fem_n_modes; ++mode) {
        const double omega_s = omega_undamped[mode] * freq_scale;
        const double c = c_from_omega(omega_s);
        mode_freqs[mode] = float(damped_hz(omega_s, c));
        mode_t60s[mode] = c > 0 ? float((2 * ln_1000) / c) : 0.f;
    }
    const float max_mode_freq = config.MaxModeFreq * std::max(1.f, freq_scale);
    uint32_t highest = fem_n_modes;
    while (highest > lowest && mode_freqs[highest - 1] > max_mode_freq) --highest;
    const uint32_t n_modes = std::min({config.NumModes, fem_n_modes, highest - lowest});
    Modes out;
    out.Freqs.assign(mode_freqs.begin() + lowest, mode_freqs.begin() + lowest + n_modes);
    out.T60s.assign(mode_t60s.begin() + lowest, mode_t60s.begin() + lowest + n_modes);
    out.Shapes.resize(size_t(n_points) * n_modes * 3);
    for (uint32_t p = 0; p < n_points; ++p)
        for (uint32_t mode = 0; mode < n_modes; ++mode)
            for (int k = 0; k < 3; ++k) out.Shapes[(size_t(p) * n_modes + mode) * 3 + k] = shapes[(size_t(p) * fem_n_modes + mode + lowest) * 3 + k] * shape_scale;
    out.Positions = std::move(positions);
    out.OriginalFundamentalFreq = lowest_freq_orig;
    return out;
}

// Jacobi eigen-decomposition of a symmetric 3x3 (ascending eigenvalues, columns of `vec`).
void Eigen3(double a[3][3], double val[3], double vec[3][3]) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) vec[i][j] = i == j;
    for (int sweep = 0; sweep < 64; ++sweep) {
        const double off = std::abs(a[0][1]) + std::abs(a[0][2]) + std::abs(a[1][2]);
        if (off < 1e-300 || off <= 1e-18 * (std::abs(a[0][0]) + std::abs(a[1][1]) + std::abs(a[2][2]))) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (a[p][q] == 0) continue;
                const double theta = (a[q][q] - a[p][p]) / (2 * a[p][q]);
                const double t = std::copysign(1.0, theta) / (std::abs(theta) + std::sqrt(theta * theta + 1));
                const double c = 1 / std::sqrt(t * t + 1), s = t * c;
                for (int k = 0; k < 3; ++k) {
                    const double akp = a[k][p], akq = a[k][q];
                    a[k][p] = c * akp - s * akq;
                    a[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) {
                    const double apk = a[p][k], aqk = a[q][k];
                    a[p][k] = c * apk - s * aqk;
                    a[q][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    const double vkp = vec[k][p], vkq = vec[k][q];
                    vec[k][p] = c * vkp - s * vkq;
                    vec[k][q] = s * vkp + c * vkq;
                }
            }
    }
    int order[3]{0, 1, 2};
    std::sort(order, order + 3, [&](int x, int y) { return a[x][x] < a[y][y]; });
    double v2[3][3];
    for (int c = 0; c < 3; ++c) {
        val[c] = a[order[c]][order[c]];
        for (int r = 0; r < 3; ++r) v2[r][c] = vec[r][order[c]];
    }
    std::copy(&v2[0][0], &v2[0][0] + 9, &vec[0][0]);
}

// ComputeMassProperties (mesh2modes.cpp:73-126): lumped quarter volumes at the vertices.
MeMassProperties MassProperties(const double *points, uint32_t n_points, const uint32_t *tets, uint32_t n_tets, const std::vector<uint8_t> &keep, double density, const float scale[3],
                                double length_to_si) {
    MeMassProperties out{};
    out.inertia_orientation[0] = 1;
    const double inv[3]{1.0 / scale[0], 1.0 / scale[1], 1.0 / scale[2]};
    std::vector<double> pos(size_t(n_points) * 3), vol(n_points, 0.0);
    for (size_t i = 0; i < size_t(n_points) * 3; ++i) pos[i] = points[i] * inv[i % 3];
    constexpr double sixth = double(1.f / 6.f); // the reference's (1.f / 6.f) float constant (:68)
    for (uint32_t t = 0; t < n_tets; ++t) {
        if (!keep[t]) continue;
        const double *a = &pos[3 * tets[4 * t]], *b = &pos[3 * tets[4 * t + 1]], *c = &pos[3 * tets[4 * t + 2]], *d = &pos[3 * tets[4 * t + 3]];
        const double u[3]{b[0] - a[0], b[1] - a[1], b[2] - a[2]}, v[3]{c[0] - a[0], c[1] - a[1], c[2] - a[2]}, w[3]{d[0] - a[0], d[1] - a[1], d[2] - a[2]};
        const double cr[3]{u[1] * v[2] - v[1] * u[2], u[2] * v[0] - v[2] * u[0], u[0] * v[1] - v[0] * u[1]};
        const double det = w[0] * cr[0] + w[1] * cr[1] + w[2] * cr[2];
        const double quarter = sixth * std::fabs(det) * 0.25;
        for (int k = 0; k < 4; ++k) vol[tets[4 * t + k]] += quarter;
    }
    double total = 0, com[3]{};
    for (uint32_t i = 0; i < n_points; ++i) {
        total += vol[i];
        for (int k = 0; k < 3; ++k) com[k] += vol[i] * pos[3 * i + k];
    }
    if (total <= 0) return out;
    for (double &c : com) c /= total;
    double inertia[3][3]{};
    for (uint32_t i = 0; i < n_points; ++i) {
        const double r[3]{pos[3 * i] - com[0], pos[3 * i + 1] - com[1], pos[3 * i + 2] - com[2]};
        const double rr = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
        for (int k = 0; k < 3; ++k) inertia[k][k] += vol[i] * (rr - r[k] * r[k]);
        inertia[0][1] -= vol[i] * r[0] * r[1];
        inertia[0][2] -= vol[i] * r[0] * r[2];
        inertia[1][2] -= vol[i] * r[1] * r[2];
    }
    inertia[1][0] = inertia[0][1], inertia[2][0] = inertia[0][2], inertia[2][1] = inertia[1][2];
    const double s = length_to_si, s5 = density * s * s * s * s * s;
    for (auto &row : inertia)
        for (double &x : row) x *= s5;
    double evals[3], axes[3][3];
    Eigen3(inertia, evals, axes);
    const double det = axes[0][0] * (axes[1][1] * axes[2][2] - axes[1][2] * axes[2][1]) - axes[0][1] * (axes[1][0] * axes[2][2] - axes[1][2] * axes[2][0]) +
                       axes[0][2] * (axes[1][0] * axes[2][1] - axes[1][1] * axes[2][0]);
    if (det < 0)
        for (int r = 0; r < 3; ++r) axes[r][0] = -axes[r][0]; // proper rotation for the quaternion (:119)
    out.mass = density * total * s * s * s;
    for (int k = 0; k < 3; ++k) {
        out.center_of_mass[k] = float(com[k]);
        out.inertia_diagonal[k] = float(evals[k]);
    }
    // Rotation matrix (columns = principal axes) -> unit quaternion (w, x, y, z).
    const double m00 = axes[0][0], m11 = axes[1][1], m22 = axes[2][2], tr = m00 + m11 + m22;
    double q[4];
    if (tr > 0) {
        const double r = std::sqrt(1 + tr) * 2;
        q[0] = 0.25 * r, q[1] = (axes[2][1] - axes[1][2]) / r, q[2] = (axes[0][2] - axes[2][0]) / r, q[3] = (axes[1][0] - axes[0][1]) / r;
    } else if (m00 > m11 && m00 > m22) {
        const double r = std::sqrt(1 + m00 - m11 - m22) * 2;
        q[0] = (axes[2][1] - axes[1][2]) / r, q[1] = 0.25 * r, q[2] = (axes[0][1] + axes[1][0]) / r, q[3] = (axes[0][2] + axes[2][0]) / r;
    } else if (m11 > m22) {
        const double r = std::sqrt(1 + m11 - m00 - m22) * 2;
        q[0] = (axes[0][2] - axes[2][0]) / r, q[1] = (axes[0][1] + axes[1][0]) / r, q[2] = 0.25 * r, q[3] = (axes[1][2] + axes[2][1]) / r;
    } else {
        const double r = std::sqrt(1 + m22 - m00 - m11) * 2;
        q[0] = (axes[1][0] - axes[0][1]) / r, q[1] = (axes[0][2] + axes[2][0]) / r, q[2] = (axes[1][2] + axes[2][1]) / r, q[3] = 0.25 * r;
    }
    const double qn = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    for (int k = 0; k < 4; ++k) out.inertia_orientation[k] = float(q[k] / qn);
    return out;
}

// Nearest tet point of each excitation position; the first strict minimum wins like the reference's scan (:627-636).
__global__ void __launch_bounds__(256) NearestPointKernel(const double *__restrict__ pts, uint32_t n_points, const float *__restrict__ excite, uint32_t *__restrict__ nearest) {
    __shared__ double best_d[256];
    __shared__ uint32_t best_i[256];
    const uint32_t e = blockIdx.x;
    const double px = double(excite[3 * e]), py = double(excite[3 * e + 1]), pz = double(excite[3 * e + 2]);
    double best = 1.7976931348623157e308;
    uint32_t idx = 0xFFFFFFFFu;
    for (uint32_t v = threadIdx.x; v < n_points; v += 256) {
        const double dx = px - pts[3 * v], dy = py - pts[3 * v + 1], dz = pz - pts[3 * v + 2];
        const double d = dx * dx + dy * dy + dz * dz;
        if (d < best) {
            best = d;
            idx = v;
        }
    }
    best_d[threadIdx.x] = best;
    best_i[threadIdx.x] = idx;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            const double d2 = best_d[threadIdx.x + o];
            const uint32_t i2 = best_i[threadIdx.x + o];
            if (d2 < best_d[threadIdx.x] || (d2 == best_d[threadIdx.x] && i2 < best_i[threadIdx.x])) {
                best_d[threadIdx.x] = d2;
                best_i[threadIdx.x] = i2;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) nearest[e] = best_i[0] == 0xFFFFFFFFu ? 0 : best_i[0];
}
// shapes[p][mode][k] = float(X[3 * point_p + k, mode])  (mesh2modes.cpp:498-504)
__global__ void GatherShapesKernel(const double *__restrict__ X, size_t n, uint32_t n_modes, const uint32_t *__restrict__ points, uint32_t n_points, float *__restrict__ shapes) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_points * n_modes * 3) return;
    const uint32_t k = i % 3, mode = (i / 3) % n_modes, p = i / (3 * n_modes);
    shapes[i] = float(X[size_t(3) * points[p] + k + size_t(mode) * n]);
}
__global__ void CastBasisKernel(const double *__restrict__ X, size_t count, float *__restrict__ out) {
    const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < count) out[i] = float(X[i]);
}
} // namespace
} // namespace me

struct MeFemSystem {
    me::FemSystem Impl;
    explicit MeFemSystem(int device) : Impl(device) {}
};
struct MeFactor {
    me::SparseCholesky Impl;
    explicit MeFactor(me::FemSystem &fem) : Impl(fem) {}
};

namespace me {
namespace {
MeStatus SolveImpl(const double *points, uint32_t n_points, const uint32_t *tets, uint32_t n_tets, const MeMaterial *c_material, const float *excite, uint32_t n_excite,
                   const float baked_scale[3], const MeSolverConfig *c_config, const float *seed_basis, uint32_t seed_rows, uint32_t seed_cols, int keep_basis, MeJobMonitor *monitor,
                   MeModalResult &res) {
    const Config config = FromC(c_config);
    const Material material = FromC(c_material);
    const float unit_scale[3]{1, 1, 1};
    const float *scale = baked_scale ? baked_scale : unit_scale;
    if (n_excite && !excite) Fail(ME_BAD_ARG, "null excitation positions");
    auto &profile = res.Profile;
    auto cancelled = [&] { return monitor && monitor->cancelled; };
    auto progress = [&](float p) {
        if (monitor) monitor->progress = p;
    };

    FemSystem fem(config.Device);
    auto s = fem.Stream;
    // FilterDegenerate, BuildQuadMesh and AssembleQuadratic run on the device; mass properties need the filter's verdict.
    double t0 = Seconds();
    fem.Build(points, n_points, tets, n_tets, material, config.ElementOrder);
    profile.assemble = Seconds() - t0; // numbering + pattern + values (QuadMesh is folded in: one device pipeline)
    profile.quad_mesh = 0;
    profile.assemble_kernel_ms = fem.AssembleKernelMs;
    profile.tets_kept = fem.NumTets;

    t0 = Seconds();
    {
        // The same degeneracy test on the host picks the tets the lumped masses sum over (mesh2modes.cpp:606-609).
        std::vector<uint8_t> keep(n_tets, 1);
        if (fem.NumTets != n_tets) {
            for (uint32_t t = 0; t < n_tets; ++t) {
                const double *p[4]{points + 3 * tets[4 * t], points + 3 * tets[4 * t + 1], points + 3 * tets[4 * t + 2], points + 3 * tets[4 * t + 3]};
                const double r0[3]{p[1][0] - p[0][0], p[1][1] - p[0][1], p[1][2] - p[0][2]}, r1[3]{p[2][0] - p[0][0], p[2][1] - p[0][1], p[2][2] - p[0][2]},
                    r2[3]{p[3][0] - p[0][0], p[3][1] - p[0][1], p[3][2] - p[0][2]};
                const double cr[3]{r1[1] * r2[2] - r2[1] * r1[2], r1[2] * r2[0] - r2[2] * r1[0], r1[0] * r2[1] - r2[0] * r1[1]};
                const double det = std::abs(r0[0] * cr[0] + r0[1] * cr[1] + r0[2] * cr[2]);
                double lmax_sq = 0;
                for (int i = 0; i < 4; ++i)
                    for (int j = i + 1; j < 4; ++j) {
                        const double d[3]{p[i][0] - p[j][0], p[i][1] - p[j][1], p[i][2] - p[j][2]};
                        lmax_sq = std::max(lmax_sq, d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
                    }
                keep[t] = det > 1e-12 * lmax_sq * std::sqrt(lmax_sq);
            }
        }
        const double length_to_si = (double(scale[0]) + scale[1] + scale[2]) / 3.0;
        res.MassProps = MassProperties(points, n_points, tets, n_tets, keep, material.Density, scale, length_to_si);
    }
    profile.mass_props = Seconds() - t0;
    progress(0.1f);
    profile.dofs = fem.N;
    profile.stiffness_nonzeros = uint32_t(fem.ScalarNonZerosK());
    if (cancelled()) return ME_CANCELLED;

    // Excitation sampling (mesh2modes.cpp:620-645): nearest tet point on the device, first-hit dedup on the host.
    t0 = Seconds();
    std::vector<uint32_t> sample_points;
    std::vector<float> positions;
    res.SamplePointOfExcitation.assign(n_excite, 0);
    if (n_excite) {
        DeviceBuffer<float> d_ex;
        DeviceBuffer<uint32_t> d_near;
        d_ex.Upload(excite, size_t(n_excite) * 3, s);
        d_near.Reserve(n_excite);
        NearestPointKernel<<<n_excite, 256, 0, s>>>(fem.Points.Ptr, n_points, d_ex.Ptr, d_near.Ptr);
        std::vector<uint32_t> nearest(n_excite);
        ME_CUDA(cudaMemcpyAsync(nearest.data(), d_near.Ptr, size_t(n_excite) * 4, cudaMemcpyDeviceToHost, s));
        ME_CUDA(cudaStreamSynchronize(s));
        std::unordered_map<uint32_t, uint32_t> sample_point_at;
        for (uint32_t i = 0; i < n_excite; ++i) {
            const auto [entry, first] = sample_point_at.emplace(nearest[i], uint32_t(sample_points.size()));
            if (first) {
                sample_points.push_back(nearest[i]);
                for (int k = 0; k < 3; ++k) positions.push_back(float(points[size_t(3) * nearest[i] + k] * (1.0 / scale[k])));
            }
            res.SamplePointOfExcitation[i] = entry->second;
        }
    }
    profile.sample_excite = Seconds() - t0;
    res.PointCount = uint32_t(sample_points.size());
    res.ReachedComputeModes = true;

    // ComputeModes (mesh2modes.cpp:441-512).
    const uint32_t n = fem.N;
    const uint32_t nev = std::min(config.NumFemModes, n - 1);
    const uint32_t ncv = std::min(std::max(nev + 20, 20u), n);
    const double sigma = -std::pow(2 * M_PI * config.MinModeFreq, 2);
    if (cancelled()) return ME_CANCELLED;
    t0 = Seconds();
    SparseCholesky factor(fem);
    factor.Factorize(sigma);
    profile.factorize = Seconds() - t0;
    profile.analyse = factor.Stats.AnalyseSeconds;
    profile.factor_flops = factor.Stats.FactorFlops;
    profile.factor_nonzeros = factor.Stats.FactorNonZeros;
    profile.factor_device_ms = factor.Stats.FactorMs;
    profile.supernodes = factor.Stats.Supernodes;
    profile.levels = factor.Stats.Levels;
    progress(0.3f);
    if (cancelled()) return ME_CANCELLED;

    t0 = Seconds();
    const bool trace = std::getenv("ME_PROFILE") != nullptr;
    const volatile int *cancel_flag = monitor ? &monitor->cancelled : nullptr;
    // A basis solved over a different mesh cannot seed this solve: it falls back to the cold path (mesh2modes.cpp:459-464).
    const bool use_subspace = seed_basis != nullptr && seed_rows == n && seed_cols >= nev;
    ShiftInvertLanczos lanczos(fem, factor, sigma);
    SubspaceIteration subspace(fem, factor, sigma);
    const double *eigenvectors = nullptr; // device, n x nev column-major, M-orthonormal
    if (use_subspace) {
        // Warm path: subspace iteration re-converges the seed in a few block iterations (:466-472).
        const SubspaceOutcome warm = subspace.Compute(nev, std::min(nev + 15, n), config.WarmTolerance, config.MaxRestarts, seed_basis, seed_cols, cancel_flag);
        profile.iterate = Seconds() - t0;
        profile.op_solve = warm.OpSolveMs * 1e-3;
        profile.op_applications = warm.OpApplications;
        profile.restarts = warm.Iterations;
        profile.kernel_launches = fem.KernelLaunches + factor.Stats.KernelLaunches + warm.KernelLaunches;
        if (warm.Cancelled) return ME_CANCELLED;
        if (!warm.Converged) {
            SetLastError("warm subspace iteration did not converge in %u iterations", config.MaxRestarts);
            return ME_NOT_CONVERGED;
        }
        res.Eigenvalues = warm.Eigenvalues;
        eigenvectors = subspace.Vectors.Ptr;
    } else {
        // Block form (panel solves, 8 Krylov vectors per pass over the factor) whenever its larger basis is a small part of
        // the problem; the single-vector form otherwise (tiny meshes) and as the fallback if a block ever loses rank.
        // ME_LANCZOS=single|block overrides the choice (A/B measurements).
        bool use_block = size_t(4) * (ShiftInvertLanczos::BlockBasisSize(nev) + kLanczosBlock) <= n;
        if (const char *env = std::getenv("ME_LANCZOS")) {
            if (!std::strcmp(env, "single")) use_block = false;
            else if (!std::strcmp(env, "block")) use_block = size_t(ShiftInvertLanczos::BlockBasisSize(nev)) + kLanczosBlock <= n;
        }
        LanczosOutcome outcome;
        if (use_block) outcome = lanczos.ComputeBlock(nev, config.Tolerance, config.MaxRestarts, cancel_flag);
        if (!use_block || outcome.RankLost) outcome = lanczos.Compute(nev, ncv, config.Tolerance, config.MaxRestarts, cancel_flag);
        profile.iterate = Seconds() - t0;
        if (trace) fprintf(stderr, "[solve] iterate %.3f s (analyse %.3f, factorize %.3f, assemble %.3f)\n", profile.iterate, profile.analyse, profile.factorize, profile.assemble);
        profile.op_solve = outcome.OpSolveMs * 1e-3;
        profile.op_applications = outcome.OpApplications;
        profile.restarts = outcome.Restarts;
        profile.kernel_launches = fem.KernelLaunches + factor.Stats.KernelLaunches + outcome.KernelLaunches;
        if (outcome.Cancelled) return ME_CANCELLED;
        if (!outcome.Converged) {
            SetLastError("eigensolver did not converge in %u restarts", config.MaxRestarts);
            return ME_NOT_CONVERGED;
        }
        res.Eigenvalues = outcome.Eigenvalues;
        eigenvectors = lanczos.Vectors.Ptr;
    }
    progress(0.95f);

    // Extract: shapes at the sample points, optional basis, post-processing.
    t0 = Seconds();
    res.SummaryShapes.assign(size_t(res.PointCount) * nev * 3, 0.f);
    if (res.PointCount) {
        DeviceBuffer<uint32_t> d_pts;
        DeviceBuffer<float> d_shapes;
        d_pts.Upload(sample_points, s);
        d_shapes.Reserve(res.SummaryShapes.size());
        const uint32_t count = res.PointCount * nev * 3;
        GatherShapesKernel<<<(count + 255) / 256, 256, 0, s>>>(eigenvectors, n, nev, d_pts.Ptr, res.PointCount, d_shapes.Ptr);
        ME_CUDA(cudaMemcpyAsync(res.SummaryShapes.data(), d_shapes.Ptr, res.SummaryShapes.size() * 4, cudaMemcpyDeviceToHost, s));
        ME_CUDA(cudaStreamSynchronize(s));
    }
    if (keep_basis) {
        DeviceBuffer<float> d_basis;
        const size_t count = size_t(n) * nev;
        d_basis.Reserve(count);
        CastBasisKernel<<<uint32_t((count + 255) / 256), 256, 0, s>>>(eigenvectors, count, d_basis.Ptr);
        res.Basis.resize(count);
        ME_CUDA(cudaMemcpyAsync(res.Basis.data(), d_basis.Ptr, count * 4, cudaMemcpyDeviceToHost, s));
        ME_CUDA(cudaStreamSynchronize(s));
        res.BasisRows = n, res.BasisCols = nev;
    }
    res.Modes = Postprocess(res.Eigenvalues, res.SummaryShapes, res.PointCount, 1.f, material, config, positions);
    profile.extract = Seconds() - t0;
    progress(1.f);
    if (res.Modes.Freqs.empty()) {
        SetLastError("no eigenfrequency at or above %g Hz", double(config.MinModeFreq));
        return ME_NO_MODES;
    }
    return ME_OK;
}
} // namespace
} // namespace me

using me::Fail;
using me::Guard;

extern "C" {

void me_solver_config_default(MeSolverConfig *c) {
    if (!c) return;
    *c = MeSolverConfig{20.f, 16000.f, 30, 45, 1e-8, 1e-4, 100, 0, 0.f, 2, 0};
}

MeStatus me_modal_solve(const double *points_xyz, uint32_t n_points, const uint32_t *tets, uint32_t n_tets, const MeMaterial *material, const float *excite_xyz, uint32_t n_excite,
                        const float baked_scale[3], const MeSolverConfig *config, const float *seed_basis, uint32_t seed_rows, uint32_t seed_cols, int keep_basis,
                        MeJobMonitor *monitor, MeModalResult **out) {
    MeStatus inner = ME_OK;
    const MeStatus outer = Guard([&] {
        if (!out) Fail(ME_BAD_ARG, "null out pointer");
        *out = nullptr;
        auto res = std::make_unique<MeModalResult>();
        inner = me::SolveImpl(points_xyz, n_points, tets, n_tets, material, excite_xyz, n_excite, baked_scale, config, seed_basis, seed_rows, seed_cols, keep_basis, monitor, *res);
        if (inner == ME_CANCELLED || inner == ME_NOT_CONVERGED) {
            // A cancel seen right after assembly returns a default-constructed ModalResult (mesh2modes.cpp:616). Past that point the
            // failure is ComputeModes returning empty ModalModes (:462,479,490): mesh2modes still hands back the mass properties, the
            // profile and the excitation remap around them (:655-657), with an empty eigen summary and basis.
            auto empty = std::make_unique<MeModalResult>();
            empty->Profile = res->Profile;
            if (res->ReachedComputeModes) {
                empty->MassProps = res->MassProps;
                empty->SamplePointOfExcitation = std::move(res->SamplePointOfExcitation);
                empty->ReachedComputeModes = true;
            }
            res = std::move(empty);
            if (inner == ME_CANCELLED) me::SetLastError("cancelled");
        }
        *out = res.release();
    });
    return outer != ME_OK ? outer : inner;
}
void me_modal_result_free(MeModalResult *r) { delete r; }

uint32_t me_modal_result_mode_count(const MeModalResult *r) { return r ? uint32_t(r->Modes.Freqs.size()) : 0; }
uint32_t me_modal_result_point_count(const MeModalResult *r) { return r ? r->PointCount : 0; }
const float *me_modal_result_freqs(const MeModalResult *r) { return r ? r->Modes.Freqs.data() : nullptr; }
const float *me_modal_result_t60s(const MeModalResult *r) { return r ? r->Modes.T60s.data() : nullptr; }
const float *me_modal_result_shapes(const MeModalResult *r) { return r ? r->Modes.Shapes.data() : nullptr; }
const float *me_modal_result_positions(const MeModalResult *r) { return r ? r->Modes.Positions.data() : nullptr; }
float me_modal_result_original_fundamental(const MeModalResult *r) { return r ? r->Modes.OriginalFundamentalFreq : 0.f; }
const uint32_t *me_modal_result_sample_point_of_excitation(const MeModalResult *r, uint32_t *count) {
    if (count) *count = r ? uint32_t(r->SamplePointOfExcitation.size()) : 0;
    return r ? r->SamplePointOfExcitation.data() : nullptr;
}
uint32_t me_modal_result_eigenpair_count(const MeModalResult *r) { return r ? uint32_t(r->Eigenvalues.size()) : 0; }
const double *me_modal_result_eigenvalues(const MeModalResult *r) { return r ? r->Eigenvalues.data() : nullptr; }
const float *me_modal_result_summary_shapes(const MeModalResult *r) { return r ? r->SummaryShapes.data() : nullptr; }
MeStatus me_modal_result_mass_properties(const MeModalResult *r, MeMassProperties *out) {
    return Guard([&] {
        if (!r || !out) Fail(ME_BAD_ARG, "null argument");
        *out = r->MassProps;
    });
}
MeStatus me_modal_result_profile(const MeModalResult *r, MeSolveProfile *out) {
    return Guard([&] {
        if (!r || !out) Fail(ME_BAD_ARG, "null argument");
        *out = r->Profile;
    });
}
const float *me_modal_result_basis(const MeModalResult *r, uint32_t *rows, uint32_t *cols) {
    if (rows) *rows = r ? r->BasisRows : 0;
    if (cols) *cols = r ? r->BasisCols : 0;
    return r && !r->Basis.empty() ? r->Basis.data() : nullptr;
}

MeStatus me_postprocess_modes(const double *eigenvalues, uint32_t n_eigen, const float *shapes, uint32_t n_points, float shape_scale, const MeMaterial *material,
                              const MeSolverConfig *config, const float *positions, MeModalResult **out) {
    return Guard([&] {
        if (!out || (!eigenvalues && n_eigen) || (n_points && (!shapes || !positions))) Fail(ME_BAD_ARG, "null argument");
        auto res = std::make_unique<MeModalResult>();
        res->Eigenvalues.assign(eigenvalues, eigenvalues + n_eigen);
        res->SummaryShapes.assign(shapes, shapes + size_t(n_points) * n_eigen * 3);
        res->PointCount = n_points;
        res->Modes = me::Postprocess(res->Eigenvalues, res->SummaryShapes, n_points, shape_scale, me::FromC(material), me::FromC(config),
                                     std::vector<float>(positions, positions + size_t(n_points) * 3));
        *out = res.release();
    });
}
MeStatus me_rescale_modes(const MeModalResult *solved, const MeMaterial *solved_material, const MeMaterial *material, const MeSolverConfig *config, MeModalResult **out) {
    return Guard([&] {
        if (!solved || !solved_material || !material || !out) Fail(ME_BAD_ARG, "null argument");
        if (solved->Eigenvalues.empty() || material->poisson_ratio != solved_material->poisson_ratio)
            Fail(ME_BAD_ARG, "material edit is not exactly scalable (Poisson ratio differs or no eigenpairs)");
        const double rho_ratio = material->density / solved_material->density;
        const double eigenvalue_scale = (material->young_modulus / solved_material->young_modulus) / rho_ratio;
        if (solved->SummaryShapes.size() != size_t(solved->PointCount) * solved->Eigenvalues.size() * 3) Fail(ME_BAD_ARG, "inconsistent eigen summary (%zu shape floats for %u points x %zu eigenpairs)", solved->SummaryShapes.size(), solved->PointCount, solved->Eigenvalues.size());
        // RescaleModes leaves the ModalEigenSummary untouched (it stays the SOLVED eigenpairs with their SolvedMaterial): only a local
        // copy of the eigenvalues is scaled, so the result can be rescaled again or archived with the original solved material.
        auto scaled = solved->Eigenvalues;
        for (auto &v : scaled) v *= eigenvalue_scale;
        auto res = std::make_unique<MeModalResult>();
        res->Eigenvalues = solved->Eigenvalues;
        res->SummaryShapes = solved->SummaryShapes;
        res->PointCount = solved->PointCount;
        res->MassProps = solved->MassProps;
        res->SamplePointOfExcitation = solved->SamplePointOfExcitation;
        res->Modes = me::Postprocess(scaled, res->SummaryShapes, res->PointCount, float(1 / std::sqrt(rho_ratio)), me::FromC(material), me::FromC(config), solved->Modes.Positions);
        *out = res.release();
    });
}

MeStatus me_fem_assemble(const double *points_xyz, uint32_t n_points, const uint32_t *tets, uint32_t n_tets, const MeMaterial *material, uint32_t element_order, int device,
                         MeFemSystem **out) {
    return Guard([&] {
        if (!out) Fail(ME_BAD_ARG, "null out pointer");
        *out = nullptr;
        auto fem = std::make_unique<MeFemSystem>(device);
        fem->Impl.Build(points_xyz, n_points, tets, n_tets, me::FromC(material), element_order);
        *out = fem.release();
    });
}
void me_fem_free(MeFemSystem *f) { delete f; }
MeStatus me_fem_info(const MeFemSystem *f, MeFemInfo *out) {
    return Guard([&] {
        if (!f || !out) Fail(ME_BAD_ARG, "null argument");
        const auto &s = f->Impl;
        *out = MeFemInfo{s.NumTets, s.NodeCount, s.N, s.Npe, s.ScalarNonZerosK(), s.ScalarNonZerosM(), s.NumBlocks, s.NumFullBlocks, s.AssembleKernelMs, s.KernelLaunches};
    });
}
MeStatus me_fem_get_element_nodes(MeFemSystem *f, uint32_t *out) {
    return Guard([&] {
        if (!f || !out) Fail(ME_BAD_ARG, "null argument");
        f->Impl.CopyElementNodes(out);
    });
}
MeStatus me_fem_get_csc(MeFemSystem *f, int which, uint64_t *colptr, uint32_t *rowidx, double *values) {
    return Guard([&] {
        if (!f || !colptr || !rowidx || !values || which < 0 || which > 1) Fail(ME_BAD_ARG, "bad argument");
        f->Impl.ExportCsc(which, colptr, rowidx, values);
    });
}
MeStatus me_fem_colour_elements(MeFemSystem *f, uint32_t *colours, uint32_t *n_colours) {
    return Guard([&] {
        if (!f || !colours) Fail(ME_BAD_ARG, "null argument");
        f->Impl.ColourElements(colours, n_colours);
    });
}
MeStatus me_fem_spmv(MeFemSystem *f, int which, const double *x, double *y, uint32_t repeats, float *ms_per_product) {
    return Guard([&] {
        if (!f || !x || !y || which < 0 || which > 1) Fail(ME_BAD_ARG, "bad argument");
        auto &fem = f->Impl;
        ME_CUDA(cudaSetDevice(fem.Device));
        me::DeviceBuffer<double> dx, dy;
        dx.Upload(x, fem.N, fem.Stream);
        dy.Reserve(fem.N);
        cudaEvent_t e0, e1;
        ME_CUDA(cudaEventCreate(&e0));
        ME_CUDA(cudaEventCreate(&e1));
        auto product = [&] { which == 0 ? fem.SpmvK(dx.Ptr, dy.Ptr) : fem.SpmvM(dx.Ptr, dy.Ptr); };
        product(); // warm
        ME_CUDA(cudaEventRecord(e0, fem.Stream));
        for (uint32_t i = 0; i < std::max(1u, repeats); ++i) product();
        ME_CUDA(cudaEventRecord(e1, fem.Stream));
        ME_CUDA(cudaMemcpyAsync(y, dy.Ptr, size_t(fem.N) * 8, cudaMemcpyDeviceToHost, fem.Stream));
        ME_CUDA(cudaStreamSynchronize(fem.Stream));
        float ms = 0;
        ME_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (ms_per_product) *ms_per_product = ms / float(std::max(1u, repeats));
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
    });
}

MeStatus me_factor_create(MeFemSystem *f, double sigma, MeFactor **out) {
    return Guard([&] {
        if (!f || !out) Fail(ME_BAD_ARG, "null argument");
        *out = nullptr;
        auto factor = std::make_unique<MeFactor>(f->Impl);
        factor->Impl.Factorize(sigma);
        *out = factor.release();
    });
}
void me_factor_free(MeFactor *f) { delete f; }
MeStatus me_factor_solve(MeFactor *f, const double *b, double *x, uint32_t width) {
    return Guard([&] {
        if (!f || !b || !x || width == 0) Fail(ME_BAD_ARG, "bad argument");
        auto &c = f->Impl;
        const size_t count = size_t(c.Rows()) * width;
        me::DeviceBuffer<double> db;
        db.Upload(b, count, c.Stream());
        cudaEvent_t e0, e1;
        ME_CUDA(cudaEventCreate(&e0));
        ME_CUDA(cudaEventCreate(&e1));
        ME_CUDA(cudaEventRecord(e0, c.Stream()));
        c.Solve(db.Ptr, db.Ptr, width);
        ME_CUDA(cudaEventRecord(e1, c.Stream()));
        ME_CUDA(cudaMemcpyAsync(x, db.Ptr, count * 8, cudaMemcpyDeviceToHost, c.Stream()));
        ME_CUDA(cudaStreamSynchronize(c.Stream()));
        ME_CUDA(cudaEventElapsedTime(&c.Stats.LastSolveMs, e0, e1));
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        c.CheckSolves();
    });
}
MeStatus me_factor_info(MeFactor *f, MeFactorInfo *out) {
    return Guard([&] {
        if (!f || !out) Fail(ME_BAD_ARG, "null argument");
        const auto &st = f->Impl.Stats;
        *out = MeFactorInfo{st.AnalyseSeconds, st.FactorFlops, st.FactorMs, st.LastSolveMs, st.FactorNonZeros, st.Supernodes, st.Levels, f->Impl.Rows(), st.KernelLaunches};
    });
}

MeStatus me_symbolic_analyse(uint32_t node_count, const uint32_t *rowptr, const uint32_t *col, const float *xyz, uint32_t *perm_out, MeSymbolicInfo *out) {
    return Guard([&] {
        if (!rowptr || !col || !xyz || !out || node_count == 0) Fail(ME_BAD_ARG, "bad argument");
        const me::Symbolic sym = me::Analyse(node_count, rowptr, col, xyz);
        uint64_t violations = 0;
        std::vector<uint8_t> seen(node_count, 0);
        for (uint32_t v : sym.Perm) {
            if (v >= node_count || seen[v]) ++violations;
            else seen[v] = 1;
        }
        auto in_rows = [&](uint32_t s, uint32_t node) {
            const auto b = sym.Rows.begin() + sym.RowPtr[s], e = sym.Rows.begin() + sym.RowPtr[s + 1];
            return std::binary_search(b, e, node);
        };
        for (uint32_t v = 0; v < node_count; ++v)
            for (uint32_t j = rowptr[v]; j < rowptr[v + 1]; ++j) {
                const uint32_t a = sym.InvPerm[v], b = sym.InvPerm[col[j]], r = std::max(a, b), c = std::min(a, b), s = sym.NodeSuper[c];
                if (r >= sym.SuperFirst[s + 1] && !in_rows(s, r)) ++violations;
            }
        for (uint32_t s = 0; s < sym.NumSuper; ++s) {
            const uint32_t p = sym.Parent[s];
            if (sym.RowPtr[s + 1] > sym.RowPtr[s] && !std::is_sorted(sym.Rows.begin() + sym.RowPtr[s], sym.Rows.begin() + sym.RowPtr[s + 1])) ++violations;
            for (uint64_t j = sym.RowPtr[s]; j < sym.RowPtr[s + 1]; ++j) {
                const uint32_t u = sym.Rows[j];
                if (u < sym.SuperFirst[s + 1]) ++violations;
                if (p >= sym.NumSuper) ++violations; // a root has no below-diagonal structure
                else if (!(u >= sym.SuperFirst[p] && u < sym.SuperFirst[p + 1]) && !in_rows(p, u)) ++violations;
            }
            if (p < sym.NumSuper && sym.Level[p] <= sym.Level[s]) ++violations;
        }
        // The dataflow schedules of the triangular solves, replayed in ticket order: every task must find its inputs published by
        // tasks holding smaller tickets (the sweeps' spin-waits rely on it), with exactly the counts it waits for.
        const auto replay = [&](const std::vector<me::SweepTask> &tasks, const std::vector<uint32_t> &links, const std::vector<uint32_t> *link_need, const std::vector<uint32_t> *need,
                                bool backward) {
            std::vector<uint32_t> arrived(sym.NumSuper, 0), solved(sym.NumSuper, 0), slabs(sym.NumSuper, 0);
            for (uint32_t s = 0; s < sym.NumSuper; ++s) slabs[s] = (3 * (sym.SuperFirst[s + 1] - sym.SuperFirst[s]) + me::kSolveRows - 1) / me::kSolveRows;
            for (const me::SweepTask &t : tasks) {
                if (t.Super >= sym.NumSuper) {
                    ++violations;
                    continue;
                }
                if (t.Kind == 0) {
                    if (arrived[t.Super] != t.Need) ++violations;
                    ++solved[t.Super];
                } else if (t.Kind == 2) {
                    for (uint32_t i = 0; i < t.LinkCount; ++i) {
                        const uint32_t s = backward ? t.Super + i : t.Super - (t.LinkCount - 1) + i;
                        if (!need || s >= sym.NumSuper || sym.MacroFirst[s] != sym.MacroFirst[t.Super] || arrived[s] != (*need)[s] || (backward && !sym.MacroBackward)) ++violations;
                    }
                    ++solved[t.Super];
                } else if (!backward) {
                    if (solved[t.Super] != slabs[t.Super] || t.Need != slabs[t.Super]) ++violations;
                    for (uint32_t i = 0; i < t.LinkCount; ++i) ++arrived[links[t.LinkBegin + i]];
                } else {
                    for (uint32_t i = 0; i < t.LinkCount; ++i) {
                        const uint32_t a = links[t.LinkBegin + i];
                        if (solved[a] != slabs[a] || (link_need && (*link_need)[t.LinkBegin + i] != slabs[a])) ++violations;
                    }
                    ++arrived[t.Super];
                }
            }
            for (uint32_t s = 0; s < sym.NumSuper; ++s)
                if (solved[s] != slabs[s] || (need && arrived[s] != (*need)[s])) ++violations;
        };
        replay(sym.FwdTasks, sym.FwdLinks, nullptr, nullptr, false);
        replay(sym.BwdTasks, sym.BwdLinks, &sym.BwdLinkNeed, nullptr, true);
        replay(sym.WideFwdTasks, sym.WideFwdLinks, nullptr, &sym.WideFwdNeed, false);
        replay(sym.WideBwdTasks, sym.WideBwdLinks, &sym.WideBwdLinkNeed, &sym.WideBwdNeed, true);
        if (perm_out) std::copy(sym.Perm.begin(), sym.Perm.end(), perm_out);
        *out = MeSymbolicInfo{sym.NumSuper, sym.NumLevels, sym.MaxPanelColumns, sym.MaxPanelRows, sym.FactorNonZeros, sym.UpdateTiles.size(), sym.PanelTiles.size(), violations,
                              sym.FactorFlops, sym.OrderingSeconds, sym.StructureSeconds};
    });
}

MeStatus me_measure_fp64_rate(int device, int mode, int iters, double *flops_per_second) {
    return Guard([&] {
        if (!flops_per_second || mode < 0 || mode > 1 || iters < 1) Fail(ME_BAD_ARG, "bad argument");
        ME_CUDA(cudaSetDevice(device));
        *flops_per_second = me::MeasureFp64Rate(mode, iters);
    });
}
}
