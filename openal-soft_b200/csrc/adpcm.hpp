// adpcm.hpp — IMA4 / MS ADPCM block decoders (see adpcm.cpp).
#pragma once
#include <cstddef>
#include <cstdint>

namespace b200mix {
size_t AdpcmBlockBytes(bool msadpcm, uint32_t channels, uint32_t samplesPerBlock);
bool AdpcmBlockValid(bool msadpcm, uint32_t samplesPerBlock);
// dst: [blocks*samplesPerBlock][channels] int16, interleaved
void DecodeIMA4(const uint8_t *src, uint32_t channels, uint32_t samplesPerBlock, size_t blocks, int16_t *dst);
void DecodeMSADPCM(const uint8_t *src, uint32_t channels, uint32_t samplesPerBlock, size_t blocks, int16_t *dst);
}
