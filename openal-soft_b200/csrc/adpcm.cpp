// adpcm.cpp — IMA4 and MS ADPCM block decoders (host side, run once per alBufferData).
//
// The reference keeps these formats compressed and decodes them inside the mixer
// (LoadSamples<IMA4Data> / <MSADPCMData>, core/voice.cpp:289-484, tables :199-238).  Both
// decoders are pure integer recurrences over a block, so decoding the whole buffer to
// int16 at upload time yields exactly the samples the reference's mixer would see
// (it converts with /32768.0f, as FMT_I16 does here); the GPU then streams plain int16.
#include "adpcm.hpp"

#include <algorithm>

namespace b200mix {
namespace {

constexpr int kImaStep[89] = {
       7,    8,    9,   10,   11,   12,   13,   14,   16,   17,   19,
      21,   23,   25,   28,   31,   34,   37,   41,   45,   50,   55,
      60,   66,   73,   80,   88,   97,  107,  118,  130,  143,  157,
     173,  190,  209,  230,  253,  279,  307,  337,  371,  408,  449,
     494,  544,  598,  658,  724,  796,  876,  963, 1060, 1166, 1282,
    1411, 1552, 1707, 1878, 2066, 2272, 2499, 2749, 3024, 3327, 3660,
    4026, 4428, 4871, 5358, 5894, 6484, 7132, 7845, 8630, 9493,10442,
   11487,12635,13899,15289,16818,18500,20350,22358,24633,27086,29794,
   32767};
constexpr int kImaCodeword[16] = {1, 3, 5, 7, 9, 11, 13, 15, -1, -3, -5, -7, -9, -11, -13, -15};
constexpr int kImaIndexAdjust[16] = {-1, -1, -1, -1, 2, 4, 6, 8, -1, -1, -1, -1, 2, 4, 6, 8};
constexpr int kMsAdaption[16] = {230, 230, 230, 230, 307, 409, 512, 614, 768, 614, 512, 409, 307, 230, 230, 230};
constexpr int kMsCoeff[7][2] = {{256, 0}, {512, -256}, {0, 0}, {192, 64}, {240, 0}, {460, -208}, {392, -232}};

inline int s16le(const uint8_t *p) { return int(int16_t(uint16_t(p[0]) | (uint16_t(p[1]) << 8))); }

} // namespace

size_t AdpcmBlockBytes(bool msadpcm, uint32_t channels, uint32_t samplesPerBlock)
{
    return msadpcm ? (size_t(samplesPerBlock - 2u)/2u + 7u)*channels
                   : (size_t(samplesPerBlock - 1u)/2u + 4u)*channels;
}

bool AdpcmBlockValid(bool msadpcm, uint32_t samplesPerBlock)
{
    // al/buffer.cpp: IMA4 blocks hold 1 + a multiple of 8 samples, MSADPCM an even count >= 2
    if(msadpcm) return samplesPerBlock >= 2u && (samplesPerBlock & 1u) == 0u;
    return samplesPerBlock >= 1u && ((samplesPerBlock - 1u) & 7u) == 0u;
}

void DecodeIMA4(const uint8_t *src, uint32_t channels, uint32_t samplesPerBlock, size_t blocks,
    int16_t *dst)
{
    const size_t blockBytes = AdpcmBlockBytes(false, channels, samplesPerBlock);
    for(size_t b = 0;b < blocks;++b, src += blockBytes, dst += size_t(samplesPerBlock)*channels)
    {
        for(uint32_t c = 0;c < channels;++c)
        {
            int sample = s16le(src + c*4u);
            int idx = std::clamp(s16le(src + c*4u + 2u), 0, 88);
            const uint8_t *nib = src + (channels + c)*4u;
            dst[c] = int16_t(sample);
            for(uint32_t n = 0;n + 1u < samplesPerBlock;++n)
            {
                const uint32_t shift = (n & 1u)*4u;
                const uint32_t word = (n >> 1) & ~3u;
                const uint32_t byte = word*channels + ((n >> 1) & 3u);
                const uint32_t code = (nib[byte] >> shift) & 0xfu;
                sample += kImaCodeword[code]*kImaStep[idx]/8;
                sample = std::clamp(sample, -32768, 32767);
                idx = std::clamp(idx + kImaIndexAdjust[code], 0, 88);
                dst[size_t(n + 1u)*channels + c] = int16_t(sample);
            }
        }
    }
}

void DecodeMSADPCM(const uint8_t *src, uint32_t channels, uint32_t samplesPerBlock, size_t blocks,
    int16_t *dst)
{
    const size_t blockBytes = AdpcmBlockBytes(true, channels, samplesPerBlock);
    for(size_t b = 0;b < blocks;++b, src += blockBytes, dst += size_t(samplesPerBlock)*channels)
    {
        for(uint32_t c = 0;c < channels;++c)
        {
            const uint32_t pred = std::min<uint32_t>(src[c], 6u);
            int scale = s16le(src + channels + 2u*c);
            int h0 = s16le(src + 3u*channels + 2u*c), h1 = s16le(src + 5u*channels + 2u*c);
            const uint8_t *nib = src + 7u*channels;
            dst[c] = int16_t(h1);
            if(samplesPerBlock > 1u) dst[size_t(channels) + c] = int16_t(h0);
            uint32_t off = c;
            for(uint32_t n = 2;n < samplesPerBlock;++n, off += channels)
            {
                const uint32_t shift = ((off & 1u) ^ 1u)*4u;
                const int nval = int((nib[off >> 1] >> shift) & 0xfu);
                const int p = ((nval ^ 0x08) - 0x08)*scale;
                const int diff = (h0*kMsCoeff[pred][0] + h1*kMsCoeff[pred][1])/256;
                const int sample = std::clamp(p + diff, -32768, 32767);
                h1 = h0; h0 = sample;
                scale = std::max(kMsAdaption[nval]*scale/256, 16);
                dst[size_t(n)*channels + c] = int16_t(sample);
            }
        }
    }
}

} // namespace b200mix
