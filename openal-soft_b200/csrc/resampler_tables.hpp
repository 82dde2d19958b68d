// resampler_tables.hpp — host-side generation of the resampler coefficient tables the
// CUDA mixer keeps in HBM/L2.  Same mathematics and f64 operation order as the
// reference's static-init tables (core/bsinc_tables.cpp:34-375, core/cubic_tables.cpp:22-128),
// so the fp32 tables are bit-identical (checked against the compiled reference in
// tests/test_tables.py).  Product code: no dependency on oracle/.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

namespace b200mix {

constexpr unsigned kBsincScales = 16;   // BSincScaleCount  core/bsinc_defs.h:8
constexpr unsigned kBsincPhases = 32;   // BSincPhaseCount  core/bsinc_defs.h:10
constexpr unsigned kCubicPhases = 32;   // CubicPhaseCount  core/cubic_defs.h:8
constexpr unsigned kMaxTaps = 48;       // MaxResamplerPadding

struct BsincTable {
    float scaleBase{}, scaleRange{};
    uint32_t m[kBsincScales]{};
    uint32_t filterOffset[kBsincScales]{};
    std::vector<float> tab;
};

// Result of BsincPrepare (alc/alu.cpp:140-165) for one step value.
struct BsincState {
    float sf{};
    uint32_t m{}, l{}, offset{};
};

BsincTable BuildBsincTable(double rejection, double order, double maxScale);
BsincState PrepareBsinc(const BsincTable &t, uint32_t increment);
// [32][8]: coeffs[4] then deltas[4] per phase.
std::vector<float> BuildSplineTable();
std::vector<float> BuildGaussianTable();
// gCubicTable: 513 floats (reverb modulation taps).
std::vector<float> BuildCubicFilter();

} // namespace b200mix
