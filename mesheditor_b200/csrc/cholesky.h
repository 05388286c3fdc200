// Supernodal sparse Cholesky of A = K - sigma*M on the device, and the triangular solves that make the shift-invert
// operator y = A^-1 x. Replaces src/audio/CholeskyShiftInvert.{h,cpp} (set_shift :26-46, perform_op :48-53,
// solve_panel :55-62), which wraps Apple Accelerate's sparse Cholesky in the reference.
#pragma once

#include "common.h"
#include "fem.h"
#include "symbolic.h"

namespace me {

struct CholeskyStats {
    double AnalyseSeconds{0};      // host: copy pattern, ordering, structures, upload
    float FactorMs{0};             // device: scatter A + all levels (CUDA events)
    float UpdateKernelMs{0};       // device: the DMMA trailing-update kernels only
    float LastSolveMs{0};
    uint64_t FactorNonZeros{0};
    double FactorFlops{0};
    uint32_t Supernodes{0}, Levels{0};
    uint32_t SweepLevels{0};       // dependency levels of the panel sweeps (a macro block of chain panels counts once)
    uint32_t KernelLaunches{0};
};

class SparseCholesky {
public:
    // Symbolic analysis of the pattern of `fem` (host) and upload of the schedules.
    explicit SparseCholesky(FemSystem &fem, const SymbolicOptions & = {});
    ~SparseCholesky();
    SparseCholesky(const SparseCholesky &) = delete;
    SparseCholesky &operator=(const SparseCholesky &) = delete;

    // Numeric factorisation of K - sigma*M (set_shift). Fails with ME_FACTOR_FAILED when a pivot is not positive.
    void Factorize(double sigma);
    // x = A^-1 b for `width` right-hand sides stored column-major (n x width), device pointers, natural DOF order
    // (perform_op for width 1, solve_panel otherwise). b and x may alias.
    void Solve(const double *b, double *x, uint32_t width = 1);
    // Synchronises and throws if a sweep since the last check raised the stall flag (a sweep never hangs: its waits are bounded).
    void CheckSolves();

    uint32_t Rows() const { return Fem.N; }
    const Symbolic &Analysis() const { return Sym; }
    CholeskyStats Stats;
    cudaStream_t Stream() const { return Fem.Stream; }

private:
    FemSystem &Fem;
    Symbolic Sym;
    DeviceBuffer<uint32_t> DSuperFirst, DRows, DNodeSuper, DInvPerm, DPerm, DSegTarget, DSegBegin, DSegEnd, DLevelOrder;
    DeviceBuffer<uint64_t> DRowPtr, DPanelOffset, DInvOffset, DSlabOffset;
    DeviceBuffer<PanelTile> DPanelTiles;
    DeviceBuffer<SweepTask> DFwdTasks, DBwdTasks;
    DeviceBuffer<uint32_t> DFwdLinks, DCounters;
    DeviceBuffer<SweepTask> DWideFwdTasks, DWideBwdTasks; // panel sweeps: runs of slabs (Symbolic::WideFwdTasks)
    DeviceBuffer<uint32_t> DWideFwdLinks, DWideBwdLinks, DWideBwdLinkNeed, DWideFwdNeed, DWideBwdNeed;
    DeviceBuffer<uint64_t> DMacroOffset, DMacroOffsetT;
    DeviceBuffer<Symbolic::MacroJob> DMacroJobs;
    DeviceBuffer<double> MacroW, MacroWT; // explicit inverses of the macro blocks' diagonal blocks (symbolic.h), forward and backward row blocks
    uint32_t FwdGrid{0}, BwdGrid{0}, WideFwdGrid{0}, WideBwdGrid{0};
    DeviceBuffer<UpdateTile> DUpdateTiles;
    DeviceBuffer<double> L, Linv, LinvT, Slabs, Work, Work2;
    DeviceBuffer<int> DFail;
    cudaEvent_t Ev[4]{};
    bool Factored{false};
    bool SchedulesUploaded{false};
    void UploadSchedules(); // waits for the background construction of the solve schedules (symbolic.h) and uploads them
    uint32_t SolvesSinceCheck{0};
};

// FP64 issue-rate micro-benchmark: mode 0 = DFMA, 1 = DMMA m8n8k4. Returns flop/s.
double MeasureFp64Rate(int mode, int iters);

} // namespace me
