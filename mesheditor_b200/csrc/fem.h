// Tetrahedral linear-elastic FEM assembly on the device: node numbering, symbolic pattern, element kernel, SpMV.
// Reference: src/audio/mesh2modes.cpp:42-60 (FilterDegenerate), :137-165 (ComputeElementBases), :209-237 (GetQuadBasis),
// :246-264 (BuildQuadMesh), :273-327 (AssembleQuadratic); lib/spectra/include/Spectra/MatOp/SparseSymMatProd.h:83-88 (M*x).
#pragma once

#include "common.h"

#include <cstdint>
#include <vector>

namespace me {

struct Material {
    double Density, Young, Poisson, Alpha, Beta;
    double Lambda() const { return (Poisson * Young) / ((1 + Poisson) * (1 - 2 * Poisson)); }
    double Mu() const { return Young / (2 * (1 + Poisson)); }
};

// Exact unit-volume integrals of the element's shape functions (P2: mesh2modes.cpp:203-237; P1: N_i = lambda_i).
struct ElementTables {
    uint32_t Npe;                 // nodes per element: 4 or 10
    std::vector<double> Mass;     // [Npe][Npe]
    std::vector<double> Grad;     // [Npe][4][Npe][4]
};
ElementTables MakeElementTables(uint32_t order);

// The assembled pencil (K, M) resident in HBM.
//
// Layout. Both matrices share one node-block pattern: block (r, c), r >= c, exists iff nodes r and c share an element.
// The lower triangle is stored block-CSC (== upper block-CSR): BlkColPtr[NodeCount+1], BlkRow[NumBlocks] ascending in
// each column, the diagonal block first. K keeps a 3x3 block per entry (row-major, KBlk[9*u + 3*p + q]); M is
// (node mass matrix) (x) I3, so it keeps one scalar per block (MBlk[u]). That is exactly the information in the
// reference's two Eigen lower-triangular CSC matrices (explicit zeros kept); ExportCsc() writes that scalar layout.
// For mat-vecs the symmetric pattern is expanded once to full block-CSR (FullRowPtr/FullCol) with values gathered
// into KFull/MFull, so that the SpMV kernels stream each stored value exactly once per product.
class FemSystem {
public:
    FemSystem(int device);
    ~FemSystem();
    FemSystem(const FemSystem &) = delete;
    FemSystem &operator=(const FemSystem &) = delete;

    // Uploads the mesh, drops degenerate tets, numbers the nodes, builds the pattern and assembles K and M.
    void Build(const double *points_xyz, uint32_t n_points, const uint32_t *tets, uint32_t n_tets, const Material &, uint32_t order);

    // y = K x / y = M x over N = 3*NodeCount scalars (device pointers, on Stream).
    void SpmvK(const double *x, double *y);
    void SpmvM(const double *x, double *y);
    // Y = M X for `width` column-major right-hand sides (leading dimension N): one pass over M per 8 columns.
    void SpmvMPanel(const double *x, double *y, uint32_t width);

    // Host copies for parity checks and for the host-side symbolic analysis.
    uint64_t ScalarNonZerosK() const { return uint64_t(9) * NumBlocks - uint64_t(3) * NodeCount; }
    uint64_t ScalarNonZerosM() const { return uint64_t(3) * NumBlocks; }
    // Eigen-layout lower CSC of K (which = 0) or M (which = 1): colptr[N+1], rowidx[nnz], values[nnz].
    void ExportCsc(int which, uint64_t *colptr, uint32_t *rowidx, double *values);
    void CopyElementNodes(uint32_t *out);            // [NumTets][Npe]
    void CopyFullPattern(std::vector<uint32_t> &rowptr, std::vector<uint32_t> &col);
    void CopyNodeCoords(std::vector<float> &xyz);    // [NodeCount][3]; midside nodes at edge midpoints
    // Deterministic first-fit element colouring in element order (oracle/modal.py greedy_colouring): out[NumTets].
    void ColourElements(uint32_t *out, uint32_t *n_colours);

    int Device;
    cudaStream_t Stream{nullptr};
    uint32_t Order{2}, Npe{10}, NumPairs{55};
    uint32_t NumPoints{0}, NumTets{0}, NumTetsIn{0}, NodeCount{0}, N{0}, NumBlocks{0}, NumFullBlocks{0};
    Material Mat{};

    DeviceBuffer<double> Points, Basis, KBlk, MBlk, KFull, MFull;
    DeviceBuffer<uint32_t> Tets, ElemNodes, BlkColPtr, BlkRow, BlkCol, ContribPtr, Contrib, AssembleOrder, FullRowPtr, FullCol, FullSrc;
    DeviceBuffer<double> TabMass;   // [Npe][Npe]
    DeviceBuffer<double> TabTermW;  // per ordered local pair (a, c): up to 4 gradient terms w * Phig[k] (x) Phig[l]
    DeviceBuffer<uint8_t> TabTermKL, TabTermCount, TabPairA, TabPairC;
    uint32_t KernelLaunches{0};
    float AssembleKernelMs{0}, SpmvMs{0};
};

} // namespace me
