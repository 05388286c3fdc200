// Host-side symbolic analysis for the sparse Cholesky of K - sigma*M: fill-reducing ordering, supernodes, level
// schedule and the work lists the device kernels consume.
// Replaces the analysis half of Apple Accelerate's SparseFactor(SparseFactorizationCholesky, ...) used by the
// reference (src/audio/CholeskyShiftInvert.cpp:26-46). Works on the NODE graph (one vertex per 3 DOFs): the three
// DOFs of a node always share their structure, so every list below is a third of its scalar size.
#pragma once

#include <cstdint>
#include <vector>

namespace me {

struct SymbolicOptions {
    uint32_t LeafNodes{40};   // dissection stops at subdomains of at most this many nodes (one dense leaf supernode)
    uint32_t PanelNodes{42};  // separators are split into chain panels of at most this many nodes (126 columns <= 128)
};

// One tile of the trailing update of supernode S into the panel of an ancestor (see cholesky.cu SyrkScatterKernel).
struct UpdateTile {
    uint32_t Super, Segment; // source supernode and which of its target segments
    uint16_t RowTile, ColTile; // tile coordinates inside the segment's trapezoid, in units of the kernel's tile edge
};
struct PanelTile {
    uint32_t Super, RowTile;
};
// A run of consecutive 64-row tiles of one panel handled by one CTA of the backward solve (fewer atomics per column).
struct PanelGroup {
    uint32_t Super, FirstTile, Tiles;
};
constexpr uint32_t kGroupTiles = 8;
constexpr uint32_t kDiagTask = 0xFFFFFFFFu;

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
    std::vector<uint64_t> SegPtr;             // [NumSuper+1] into Seg*
    std::vector<uint32_t> SegTarget, SegBegin, SegEnd; // runs of Rows that are columns of one ancestor supernode
    // Work lists per level.
    std::vector<uint64_t> PanelTilePtr, UpdateTilePtr; // [NumLevels+1]
    std::vector<PanelTile> PanelTiles;        // 64-row tiles of the below-diagonal panels (TRSM, solves)
    std::vector<uint64_t> PanelGroupPtr;      // [NumLevels+1]
    std::vector<PanelGroup> PanelGroups;
    std::vector<UpdateTile> UpdateTiles;
    // Dataflow schedules of the triangular solves (cholesky.cu SweepKernel): tasks in a topological order, taken by
    // persistent CTAs through a ticket counter and gated by per-supernode arrival counters instead of level barriers.
    std::vector<PanelTile> FwdTasks;          // per supernode ascending: {s, kDiagTask} then its row tiles
    std::vector<uint32_t> FwdTargetPtr, FwdTargets; // per task: ancestors whose right-hand side the tile updates
    std::vector<uint32_t> FwdExpected;        // [NumSuper] tiles that must arrive before the supernode's diagonal solve
    std::vector<PanelGroup> BwdTasks;         // per supernode descending: its tile groups, then {s, 0, 0} = diagonal solve
    std::vector<uint32_t> BwdDepPtr, BwdDeps; // per task: ancestors whose solution the group reads
    std::vector<uint32_t> BwdExpected;        // [NumSuper] groups that must arrive before the supernode's diagonal solve
    uint64_t FactorNonZeros{0};               // scalars stored in the panels
    double FactorFlops{0};
    uint32_t MaxPanelColumns{0}, MaxPanelRows{0};
    double OrderingSeconds{0}, StructureSeconds{0};
};

constexpr uint32_t kTile = 64; // edge of the dense tiles the numeric kernels work on

// rowptr/col: full symmetric node adjacency (CSR, diagonal included or not); xyz: node coordinates for the
// geometric nested dissection.
Symbolic Analyse(uint32_t node_count, const uint32_t *rowptr, const uint32_t *col, const float *xyz, const SymbolicOptions & = {});

} // namespace me
