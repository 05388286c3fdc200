// The modal model a solve returns, shared by the solver (solve.cu) and the model interchange (interchange.cpp).
#pragma once

#include "common.h"

#include <cstdint>
#include <vector>

namespace me {
// ModalModes (src/audio/ModalModes.h:7-20).
struct Modes {
    std::vector<float> Freqs, T60s;
    std::vector<float> Shapes;    // [point][mode][3]
    std::vector<float> Positions; // [point][3]
    float OriginalFundamentalFreq{0};
};
} // namespace me

// modal::ModalResult (src/audio/mesh2modes.h:52-62).
struct MeModalResult {
    me::Modes Modes;
    MeMassProperties MassProps{};
    MeSolveProfile Profile{};
    std::vector<double> Eigenvalues;
    std::vector<float> SummaryShapes; // [point][eigenpair][3]
    std::vector<uint32_t> SamplePointOfExcitation;
    std::vector<float> Basis;
    uint32_t BasisRows{0}, BasisCols{0};
    uint32_t PointCount{0};
    bool ReachedComputeModes{false}; // a failure past this point keeps mass properties, profile and the excitation remap (mesh2modes.cpp:655-657)
};
