// Drop-in replacement translation unit for src/audio/mesh2modes.cpp of khiner/MeshEditor: the same three functions
// (modal::mesh2modes, modal::PostprocessModes, modal::RescaleModes; src/audio/mesh2modes.h:77-88) implemented on top of
// libme_modal.so's C ABI (include/me_modal.h). Build the reference with this file INSTEAD of mesh2modes.cpp and
// CholeskyShiftInvert.cpp and link libme_modal.so. It is compiled inside the reference tree (it includes
// the reference's own headers, Eigen included, for the ModalResult types), so it is not built in this repository.
#include "mesh2modes.h"

#include "AcousticMaterialProperties.h"
#include "Job.h"
#include "mesh/TetMesh.h"

#include "me_modal.h"

#include <atomic>
#include <stdexcept>
#include <thread>

namespace {
MeMaterial ToC(const AcousticMaterialProperties &m) { return {m.Density, m.YoungModulus, m.PoissonRatio, m.Alpha, m.Beta}; }
MeSolverConfig ToC(const modal::SolverConfig &c) {
    MeSolverConfig out;
    me_solver_config_default(&out);
    out.min_mode_freq = c.MinModeFreq, out.max_mode_freq = c.MaxModeFreq;
    out.num_modes = c.NumModes, out.num_fem_modes = c.NumFemModes;
    out.tolerance = c.Tolerance, out.warm_tolerance = c.WarmTolerance, out.max_restarts = c.MaxRestarts;
    out.has_fundamental_freq = c.FundamentalFreq.has_value(), out.fundamental_freq = c.FundamentalFreq.value_or(0.f);
    out.element_order = 2; // the reference's 10-node elements
    return out;
}
ModalModes ModesOf(const MeModalResult *r) {
    ModalModes m;
    const uint32_t modes = me_modal_result_mode_count(r), points = me_modal_result_point_count(r);
    if (modes == 0) return m;
    m.Freqs.assign(me_modal_result_freqs(r), me_modal_result_freqs(r) + modes);
    m.T60s.assign(me_modal_result_t60s(r), me_modal_result_t60s(r) + modes);
    const float *shapes = me_modal_result_shapes(r), *pos = me_modal_result_positions(r);
    m.Shapes.assign(points, std::vector<vec3>(modes));
    for (uint32_t p = 0; p < points; ++p) {
        for (uint32_t k = 0; k < modes; ++k) m.Shapes[p][k] = {shapes[(size_t(p) * modes + k) * 3], shapes[(size_t(p) * modes + k) * 3 + 1], shapes[(size_t(p) * modes + k) * 3 + 2]};
        m.Positions.emplace_back(pos[3 * p], pos[3 * p + 1], pos[3 * p + 2]);
    }
    m.OriginalFundamentalFreq = me_modal_result_original_fundamental(r);
    return m;
}
struct ResultGuard {
    MeModalResult *R{};
    ~ResultGuard() { me_modal_result_free(R); }
};
} // namespace

modal::ModalResult modal::mesh2modes(const TetMesh &tets, const AcousticMaterialProperties &material, const std::vector<vec3> &excite_positions, vec3 baked_scale, SolverConfig config,
                                     SolveReuse reuse, JobMonitor *monitor) {
    const MeMaterial c_material = ToC(material);
    const MeSolverConfig c_config = ToC(config);
    const float scale[3]{baked_scale.x, baked_scale.y, baked_scale.z};
    // JobMonitor holds std::atomics; mirror it into the plain-C monitor from a watcher thread while the solve runs.
    // The C monitor is a plain struct the library polls through volatile reads; this side touches it through atomic_ref only.
    MeJobMonitor c_monitor{0.f, 0};
    std::atomic<bool> done{false};
    std::thread watcher;
    const auto mirror = [&] {
        if (monitor->Cancelled()) std::atomic_ref<int>(c_monitor.cancelled).store(1, std::memory_order_relaxed);
        monitor->Progress.store(std::atomic_ref<float>(c_monitor.progress).load(std::memory_order_relaxed), std::memory_order_relaxed);
    };
    if (monitor) watcher = std::thread([&] {
        while (!done.load(std::memory_order_acquire)) {
            mirror();
            std::this_thread::sleep_for(std::chrono::milliseconds(2));
        }
    });
    ResultGuard guard;
    const float *seed = reuse.SeedBasis ? reuse.SeedBasis->data() : nullptr; // Eigen::MatrixXf is column-major
    const MeStatus status = me_modal_solve(tets.Points.empty() ? nullptr : &tets.Points[0].x, uint32_t(tets.Points.size()), tets.Tets.empty() ? nullptr : tets.Tets[0].data(), uint32_t(tets.Tets.size()), &c_material,
                                           excite_positions.empty() ? nullptr : &excite_positions[0].x, uint32_t(excite_positions.size()), scale, &c_config, seed,
                                           seed ? uint32_t(reuse.SeedBasis->rows()) : 0, seed ? uint32_t(reuse.SeedBasis->cols()) : 0, reuse.KeepBasis, monitor ? &c_monitor : nullptr, &guard.R);
    done.store(true, std::memory_order_release);
    if (watcher.joinable()) watcher.join();
    if (monitor) mirror(); // the watcher may have slept through the last store: the final progress (1.0) is copied here
    if (status == ME_FACTOR_FAILED) throw std::runtime_error(me_last_error()); // CholeskyShiftInvert.cpp:44
    // ME_CANCELLED / ME_NOT_CONVERGED / ME_NO_MODES: the library hands back what the reference does — nothing at all for a cancel
    // seen right after assembly (mesh2modes.cpp:616), and empty Modes / Summary / Basis around the mass properties, profile and
    // excitation remap for a failure inside ComputeModes (:462,479,490 return from ComputeModes, :655-657 still build the result).
    if (status != ME_OK && status != ME_NO_MODES && status != ME_CANCELLED && status != ME_NOT_CONVERGED) throw std::runtime_error(me_last_error());
    if (!guard.R) return {};
    ModalResult out;
    const MeModalResult *r = guard.R;
    out.Modes = ModesOf(r);
    MeMassProperties mp;
    me_modal_result_mass_properties(r, &mp);
    out.MassProps = {mp.mass, {mp.center_of_mass[0], mp.center_of_mass[1], mp.center_of_mass[2]}, {mp.inertia_diagonal[0], mp.inertia_diagonal[1], mp.inertia_diagonal[2]},
                     glm::quat{mp.inertia_orientation[0], mp.inertia_orientation[1], mp.inertia_orientation[2], mp.inertia_orientation[3]}};
    MeSolveProfile p;
    me_modal_result_profile(r, &p);
    out.Profile = {p.mass_props, p.quad_mesh, p.assemble, p.sample_excite, p.factorize, p.iterate, p.op_solve, p.extract, p.dofs, p.stiffness_nonzeros, p.op_applications, p.restarts};
    const uint32_t pairs = me_modal_result_eigenpair_count(r), points = me_modal_result_point_count(r);
    out.Summary.Eigenvalues.assign(me_modal_result_eigenvalues(r), me_modal_result_eigenvalues(r) + pairs);
    const float *ss = me_modal_result_summary_shapes(r);
    out.Summary.Shapes.assign(points, std::vector<vec3>(pairs));
    for (uint32_t q = 0; q < points; ++q)
        for (uint32_t k = 0; k < pairs; ++k) out.Summary.Shapes[q][k] = {ss[(size_t(q) * pairs + k) * 3], ss[(size_t(q) * pairs + k) * 3 + 1], ss[(size_t(q) * pairs + k) * 3 + 2]};
    if (pairs) out.Summary.SolvedMaterial = material; // ComputeModes fills the summary only on success (:503-506)
    uint32_t rows = 0, cols = 0, count = 0;
    if (const float *basis = me_modal_result_basis(r, &rows, &cols)) out.Basis = Eigen::Map<const Eigen::MatrixXf>(basis, rows, cols);
    const uint32_t *remap = me_modal_result_sample_point_of_excitation(r, &count);
    out.SamplePointOfExcitation.assign(remap, remap + count);
    return out;
}

ModalModes modal::PostprocessModes(std::span<const double> eigenvalues, const std::vector<std::vector<vec3>> &shapes, float shape_scale, const AcousticMaterialProperties &material,
                                   const SolverConfig &config, std::vector<vec3> positions) {
    std::vector<float> flat;
    for (const auto &point : shapes)
        for (const auto &s : point) flat.insert(flat.end(), {s.x, s.y, s.z});
    const MeMaterial c_material = ToC(material);
    const MeSolverConfig c_config = ToC(config);
    ResultGuard guard;
    if (me_postprocess_modes(eigenvalues.data(), uint32_t(eigenvalues.size()), flat.data(), uint32_t(shapes.size()), shape_scale, &c_material, &c_config,
                             positions.empty() ? nullptr : &positions[0].x, &guard.R) != ME_OK)
        throw std::runtime_error(me_last_error());
    return ModesOf(guard.R);
}
// modal::RescaleModes stays the reference's own eight lines (mesh2modes.cpp:590-603): it only calls PostprocessModes.
