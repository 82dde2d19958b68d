// hrtf_store.hpp — the loaded HRTF data set (shared by hrtf_store.cpp and b200mix.cu).
#pragma once
#include <cstdint>
#include <vector>

struct b200mix_hrtf {
    uint32_t sample_rate{}, ir_size{};
    struct Field { float distance; uint32_t ev_count; };
    struct Elev { uint32_t az_count, ir_offset; };
    std::vector<Field> fields;
    std::vector<Elev> elevs;
    std::vector<float> coeffs;      // [ir_count][ir_size][2]
    std::vector<uint8_t> delays;    // [ir_count][2], quarter samples
};
