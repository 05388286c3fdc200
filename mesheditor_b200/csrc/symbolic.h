// Host-side symbolic analysis for the sparse Cholesky of K - sigma*M: fill-reducing ordering, supernodes, level
// schedule and the work lists the device kernels consume.
// Replaces the analysis half of Apple Accelerate's SparseFactor(SparseFactorizationCholesky, ...) used by the
// reference (src/audio/CholeskyShiftInvert.cpp:26-46). Works on the NODE graph (one vertex per 3 DOFs): the three
// DOFs of a node always share their structure, so every list below is a third of its scalar size.
#pragma once

#include <cstdint>
#include <thread>
#include <vector>

namespace me {

struct SymbolicOptions {
    uint32_t LeafNodes{40};   // dissection stops at subdomains of at most this many nodes (one dense leaf supernode)
    uint32_t PanelNodes{42};  // separators are split into chain panels of at most this many nodes (126 columns <= 128)
    uint32_t MacroPanels{4};  // panel sweeps: consecutive panels of a chain solved as one block through its explicit inverse (1 = off;
                              // ME_MACRO_PANELS). Measured with the slab ring at 1M tets, 8 panels: forward 2.09 -> 1.90 ms, backward 2.67 -> 2.91 ms,
    bool MacroBackward{false}; // so the backward sweep keeps its panels one by one unless ME_MACRO_BACKWARD is set (forward only, same box: 4.72 ms per
                               // panel application without, 4.57 with 8 panels, 4.48 with 4)
};

// One tile of the trailing update of supernode S into the panel of an ancestor (see cholesky.cu SyrkScatterKernel).
struct UpdateTile {
    uint32_t Super, Segment; // source supernode and which of its target segments
    uint16_t RowTile, ColTile; // tile coordinates inside the segment's trapezoid, in units of the kernel's tile edge
};
struct PanelTile {
    uint32_t Super, RowTile;
};
// One task of a triangular-solve sweep: a slab of kSolveRows rows of a diagonal block's inverse (Kind 0) or of a panel
// (Kind 1), with everything the kernel needs in one 64-byte record so that a task costs a single descriptor fetch.
constexpr uint32_t kSolveRows = 32;
constexpr uint32_t kWideRun = 8;     // longest run of slabs in one task of the panel sweeps
constexpr uint32_t kWideRunLinks = 32; // a forward run owes at most this many arrivals (one per lane of the publishing warp)
struct alignas(64) SweepTask {
    uint64_t Base;       // offset in doubles of the matrix block: into Linv / Linv^T (diag), L (forward panel of the single-vector sweeps), or of the
                         // task's first 32-row slab in the slab copy of the rectangles (Symbolic::SlabOffset; every other panel task)
    uint32_t Kind;       // 0 diagonal slab, 1 panel slab, 2 slab of a macro block's inverse (panel sweeps only)
    uint32_t Super;
    uint32_t K, Limit;   // columns of the supernode; row limit (k for diagonal slabs, m for panel slabs; macro slabs: columns of the inverse's row block)
    uint32_t Ld;         // leading dimension of the block
    uint32_t Row0;       // first row of the slab
    uint32_t VecOffset;  // 3 * SuperFirst[s]: where the supernode's own entries sit in the permuted vectors
    uint32_t RowsBase;   // RowPtr[s]: the supernode's below-diagonal node list
    uint32_t LinkBegin, LinkCount; // ancestors updated (forward panel) / read (backward panel); macro slabs: LinkCount = panels of the macro block whose entries it reads
    uint32_t Need;       // arrivals to wait for: diagonal slab = contributions to the supernode; forward panel = its diagonal slabs
    uint32_t Count;      // panel sweeps (WideTasks): consecutive kSolveRows-row slabs of the panel covered by this task (>= 1)
    uint32_t DiagColumn; // macro slabs, forward: column of the row block where the supernode's own diagonal block starts
    uint32_t RowMin;     // panel tasks: rows below this one belong to the supernode's own macro block and are left out (Row0 is slab-aligned)
};

struct Symbolic {
    uint32_t NodeCount{0}, NumSuper{0}, NumLevels{0};
    std::vector<uint32_t> Perm, InvPerm;      // Perm[new] = old node, InvPerm[old] = new
    std::vector<uint32_t> SuperFirst;         // [NumSuper+1] first node (new numbering) of each supernode; elimination order
    std::vector<uint32_t> NodeSuper;          // [NodeCount] supernode of each node (new numbering)
    std::vector<uint64_t> RowPtr;             // [NumSuper+1] into Rows
    std::vector<uint32_t> Rows;               // below-diagonal node structure of each supernode, ascending (new numbering)
    std::vector<uint32_t> Parent, Level;      // supernodal tree (Parent == NumSuper for roots); Level = height above the leaves
    std::vector<uint32_t> LevelPtr, LevelOrder; // supernodes grouped by level
    std::vector<uint64_t> PanelOffset;        // [NumSuper+1] offset in doubles of each dense panel, (3k + 3m) x 3k column-major
    std::vector<uint64_t> InvOffset;          // [NumSuper+1] offset in doubles of each inverted diagonal block (3k x 3k)
    // The below-diagonal rectangles once more in the layout the sweeps stream: per supernode ceil(rows / 32) slabs of 32 rows, each
    // ONE contiguous block of 32 x kp doubles (kp = columns rounded up to 8) made of 8 x 8 row-major tiles, [strip of 8 rows][tile of
    // 8 columns]: a slab is a single bulk copy into shared memory, and both sweeps read DMMA fragments from it without bank conflicts
    // (forward: an 8 x 8 tile is one 16-byte load per lane; backward: half a tile, 4 rows, is 256 contiguous bytes).
    std::vector<uint64_t> SlabOffset;         // [NumSuper+1] offset in doubles of each supernode's slabs
    std::vector<uint64_t> SegPtr;             // [NumSuper+1] into Seg*
    std::vector<uint32_t> SegTarget, SegBegin, SegEnd; // runs of Rows that are columns of one ancestor supernode
    // Work lists per level.
    std::vector<uint64_t> PanelTilePtr, UpdateTilePtr; // [NumLevels+1]
    std::vector<PanelTile> PanelTiles;        // 64-row tiles of the below-diagonal panels (TRSM, solves)
    std::vector<UpdateTile> UpdateTiles;
    // Dataflow schedules of the triangular solves (cholesky.cu SweepKernel): tasks in a topological order, taken by
    // persistent CTAs through a ticket counter and gated by per-supernode arrival counters instead of level barriers.
    // Every task is one kSolveRows-row slab: a slab of a diagonal block's inverse, or of a panel.
    std::vector<SweepTask> FwdTasks;          // levels ascending: the level's diagonal slabs, then its panel slabs
    std::vector<SweepTask> BwdTasks;          // levels descending: the level's panel slabs, then its diagonal slabs
    std::vector<uint32_t> FwdLinks;           // forward panel slabs: the ancestors they update
    std::vector<uint32_t> BwdLinks, BwdLinkNeed; // backward panel slabs: the ancestors they read, and how many diagonal slabs each has
    // The same two schedules for the 8-wide panel sweeps (WideSweepKernel), where a panel task is a RUN of up to kWideRun
    // consecutive slabs of one panel (SweepTask::Count): long runs on the wide lower levels of the tree, where a run amortises
    // the handshakes and (backward) the atomics; single slabs on the narrow upper levels, where the slabs of one panel are all
    // the parallelism there is.
    std::vector<SweepTask> WideFwdTasks, WideBwdTasks;
    std::vector<uint32_t> WideFwdLinks, WideBwdLinks, WideBwdLinkNeed;
    std::vector<uint32_t> WideFwdNeed, WideBwdNeed; // [NumSuper] contributions every supernode's entries wait for in the panel sweeps
    // Macro blocks of the panel sweeps. The panels of a separator chain depend on each other one after the other (solve the diagonal
    // block, update the rest of the chain, next panel): two dependent tasks per panel, which on the dense top separators IS the
    // sweep's critical path. Up to MacroPanels consecutive panels of a chain therefore form a macro block whose lower-triangular
    // diagonal block (the panels' diagonal blocks and the rows that couple them) is inverted explicitly after the numeric
    // factorisation (cholesky.cu MacroInverseKernel): all its panels are then solved in ONE step from the block's entries, and
    // their updates of the rows outside the block follow in one more.
    std::vector<uint32_t> MacroFirst, MacroLast; // [NumSuper] first and last supernode of the macro block a supernode belongs to
    std::vector<uint64_t> MacroOffset;        // [NumSuper+1] forward row block of each macro panel i: k_i x (k_0 + .. + k_i), column-major; empty for single supernodes
    std::vector<uint64_t> MacroOffsetT;       // [NumSuper+1] backward row block: k_i x (k_i + .. + k_last), element (r, c) = inverse(c, r) counted from the panel's own start
    struct MacroJob {
        uint32_t First, Last, Column, Slice;  // macro block, which of its panels' columns, which 64-column slice of those
    };
    std::vector<MacroJob> MacroJobs;
    uint32_t SweepLevels{0};                  // dependency levels of the forward panel sweep (macro blocks count once)
    bool MacroBackward{false};                // the backward panel sweep takes the macro blocks too (SymbolicOptions::MacroBackward)
    uint64_t FactorNonZeros{0};               // scalars stored in the panels
    double FactorFlops{0};
    uint32_t MaxPanelColumns{0}, MaxPanelRows{0};
    double OrderingSeconds{0}, StructureSeconds{0};

    // AnalyseInto may leave the four solve schedules above under construction on a background thread (they are first needed by
    // the first triangular solve, a numeric factorisation later): WaitSchedules joins it. The object must not move meanwhile.
    std::thread ScheduleThread;
    bool ScheduleFailed{false};
    void WaitSchedules();
    Symbolic() = default;
    Symbolic(Symbolic &&) = default;
    Symbolic &operator=(Symbolic &&) = default;
    ~Symbolic() {
        if (ScheduleThread.joinable()) ScheduleThread.join();
    }
};

constexpr uint32_t kTile = 64; // edge of the dense tiles the numeric kernels work on

// rowptr/col: full symmetric node adjacency (CSR, diagonal included or not); xyz: node coordinates for the
// geometric nested dissection.
Symbolic Analyse(uint32_t node_count, const uint32_t *rowptr, const uint32_t *col, const float *xyz, const SymbolicOptions & = {});
// The same into an object that stays where it is; with `schedules_in_background` the solve schedules are still being built when
// it returns (Symbolic::WaitSchedules).
void AnalyseInto(Symbolic &out, uint32_t node_count, const uint32_t *rowptr, const uint32_t *col, const float *xyz, const SymbolicOptions &, bool schedules_in_background);

} // namespace me
