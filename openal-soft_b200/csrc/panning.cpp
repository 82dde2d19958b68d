// Host-side panning helpers of the parameter stage (no GPU): the ambisonic encoder
// coefficients of a direction and the per-channel mixing gains they give on a mix bus.
// Restates CalcDirectionCoeffs / CalcAmbiCoeffs (core/mixer.h:68-73, core/mixer.cpp:16-91,
// core/ambidefs.h:219-272) and ComputePanGains (core/mixer.cpp:93-102) operation for
// operation, so a host that does not keep the reference's ALU gets bit-identical
// Gains.Target arrays (tests/test_pan_params.py pins both against the compiled reference).
#include "../../include/b200mix.h"

#include <algorithm>
#include <cmath>

#include "param_math.hpp"

namespace {
struct HostMath {
    static float sqrt(float x) { return std::sqrt(x); }
    static float sin(float x) { return std::sin(x); }
    static float cos(float x) { return std::cos(x); }
};
} // namespace

extern "C" {

int b200mix_ambi_coeffs(const float dir[3], float spread, float coeffs[B200MIX_MAX_AMBI_CHANNELS])
{
    if(!dir || !coeffs) return B200MIX_ERR_INVALID;
    b200mix::pm::ambi_coeffs<HostMath>(dir, spread, coeffs);
    return B200MIX_OK;
}

int b200mix_pan_gains(uint32_t channels, const float *scale, const uint32_t *index,
    const float coeffs[B200MIX_MAX_AMBI_CHANNELS], float ingain, float *gains, uint32_t gains_len)
{
    if(!scale || !index || !coeffs || !gains || channels > gains_len) return B200MIX_ERR_INVALID;
    return b200mix::pm::pan_gains(channels, scale, index, coeffs, ingain, gains, gains_len)
        ? B200MIX_OK : B200MIX_ERR_INVALID;
}

// BiquadFilter::SetParams behind setParamsFromSlope (core/filters/biquad.h:61-62,92-97;
// biquad.cpp:48-129).  Host arithmetic only.
int b200mix_biquad_coeffs(uint32_t type, float f0norm, float gain, float slope, float coeffs[5])
{
    if(type > 5u || !coeffs || !(slope > 0.0f)) return B200MIX_ERR_INVALID;
    b200mix::pm::biquad_coeffs<HostMath>(type, f0norm, gain, slope, coeffs);
    return B200MIX_OK;
}

int b200mix_convolution_gains(uint32_t layout, uint32_t pairwise, float slot_gain, uint32_t channels,
    const float *scale, const uint32_t *index, float *gains, uint32_t gains_stride)
{
    // ConvolutionState::update for a non-ambisonic impulse response (alc/effects/convolution.cpp:
    // 541-620): every IR channel's output line is panned to its speaker position (MonoMap ... X71Map,
    // :144-196), the front pair stretched for pair-wise stereo devices (ScaleAzimuthFront, :552-576);
    // the LFE line gets no gains
    if(!scale || !index || !gains || gains_stride < channels) return B200MIX_ERR_INVALID;
    constexpr float sin30 = 0.5f, cos30 = 0.866025403785f;
    constexpr float sin45 = 1.41421356237309504880f*0.5f, cos45 = 1.41421356237309504880f*0.5f;
    constexpr float sin110 = 0.939692620786f, cos110 = -0.342020143326f;
    struct Chan { bool lfe; float pos[3]; };
    Chan chans[8];
    uint32_t n = 0;
    auto add = [&](float x, float y, float z, bool lfe = false) { chans[n++] = Chan{lfe, {x, y, z}}; };
    switch(layout)
    {
    case 1u: add(0.0f, 0.0f, -1.0f); break;
    case B200MIX_LAYOUT_STEREO: add(-sin30, 0.0f, -cos30); add(sin30, 0.0f, -cos30); break;
    case B200MIX_LAYOUT_REAR: add(-sin30, 0.0f, cos30); add(sin30, 0.0f, cos30); break;
    case B200MIX_LAYOUT_QUAD:
        add(-sin45, 0.0f, -cos45); add(sin45, 0.0f, -cos45); add(-sin45, 0.0f, cos45); add(sin45, 0.0f, cos45);
        break;
    case B200MIX_LAYOUT_X51:
        add(-sin30, 0.0f, -cos30); add(sin30, 0.0f, -cos30); add(0.0f, 0.0f, -1.0f); add(0.0f, 0.0f, 0.0f, true);
        add(-sin110, 0.0f, -cos110); add(sin110, 0.0f, -cos110);
        break;
    case B200MIX_LAYOUT_X61:
        add(-sin30, 0.0f, -cos30); add(sin30, 0.0f, -cos30); add(0.0f, 0.0f, -1.0f); add(0.0f, 0.0f, 0.0f, true);
        add(0.0f, 0.0f, 1.0f); add(-1.0f, 0.0f, 0.0f); add(1.0f, 0.0f, 0.0f);
        break;
    case B200MIX_LAYOUT_X71:
        add(-sin30, 0.0f, -cos30); add(sin30, 0.0f, -cos30); add(0.0f, 0.0f, -1.0f); add(0.0f, 0.0f, 0.0f, true);
        add(-sin30, 0.0f, cos30); add(sin30, 0.0f, cos30); add(-1.0f, 0.0f, 0.0f); add(1.0f, 0.0f, 0.0f);
        break;
    default: return B200MIX_ERR_INVALID;
    }
    for(uint32_t c = 0;c < n;++c)
    {
        float *g = gains + size_t(c)*gains_stride;
        for(uint32_t k = 0;k < gains_stride;++k) g[k] = 0.0f;
        if(chans[c].lfe) continue;
        float pos[3] = {chans[c].pos[0], chans[c].pos[1], chans[c].pos[2]};
        if(pairwise && pos[2] < 0.0f)
        {
            const float len2d = std::sqrt(pos[0]*pos[0] + pos[2]*pos[2]);
            float x = pos[0] / len2d;
            float z = -pos[2] / len2d;
            if(z > cos30)
            {
                x = x*3.0f - x*x*x*4.0f;
                z = z*z*z*4.0f - z*3.0f;
                pos[0] = x * len2d;
                pos[2] = -z * len2d;
            }
            else
            {
                pos[0] = std::copysign(len2d, pos[0]);
                pos[2] = 0.0f;
            }
        }
        float coeffs[B200MIX_MAX_AMBI_CHANNELS];
        b200mix_ambi_coeffs(pos, 0.0f, coeffs);
        if(int rc = b200mix_pan_gains(channels, scale, index, coeffs, slot_gain, g, channels)) return rc;
    }
    return int(n);
}

int b200mix_builtin_decoder(uint32_t layout, uint32_t hq_mode, uint32_t sample_rate, b200mix_builtin_decoder_out *out)
{
    // InitPanning with one of the built-in layouts (alc/panning.cpp:542-640,718-845): the decoder
    // rows times the per-order gains, transposed into BFormatDec's [input][output] gains
    // (core/bformatdec.cpp:28-58), the Dry mix's AmbiMap, and the dual-band crossover (400 Hz,
    // core/device.h:238) when decoder/hq-mode is on and the layout has LF rows.  All of them are
    // pantaphonic (2D).  Output channel indices follow the device formats' channel order
    // (FL FR [FC LFE] [BL BR | BC] [SL SR]).
    if(!out || out->struct_size != sizeof(*out) || !sample_rate) return B200MIX_ERR_INVALID;
    struct Row { uint32_t real_index; float c[5]; };
    struct Config { uint32_t order, real_channels, nrows; bool fuma; float orderHF[3], orderLF[3]; bool hasLF;
        const Row *hf, *lf; };
    static const Row mono[1] = {{0u, {1.0f}}};
    static const Row stereo[2] = {{0u, {5.00000000e-1f, 2.88675135e-1f, 5.52305643e-2f}},
                                  {1u, {5.00000000e-1f, -2.88675135e-1f, 5.52305643e-2f}}};
    static const Row quad[4] = {{2u, {2.50000000e-1f, 2.04124145e-1f, -2.04124145e-1f}},
                                {0u, {2.50000000e-1f, 2.04124145e-1f, 2.04124145e-1f}},
                                {1u, {2.50000000e-1f, -2.04124145e-1f, 2.04124145e-1f}},
                                {3u, {2.50000000e-1f, -2.04124145e-1f, -2.04124145e-1f}}};
    // SideLeft, FrontLeft, FrontCenter, FrontRight, SideRight
    static const Row x51hf[5] = {
        {4u, {5.67316000e-1f, 4.22920000e-1f, -3.15495000e-1f, -6.34490000e-2f, -2.92380000e-2f}},
        {0u, {3.68584000e-1f, 2.72349000e-1f, 3.21616000e-1f, 1.92645000e-1f, 4.82600000e-2f}},
        {2u, {1.83579000e-1f, 0.00000000e+0f, 1.99588000e-1f, 0.00000000e+0f, 9.62820000e-2f}},
        {1u, {3.68584000e-1f, -2.72349000e-1f, 3.21616000e-1f, -1.92645000e-1f, 4.82600000e-2f}},
        {5u, {5.67316000e-1f, -4.22920000e-1f, -3.15495000e-1f, 6.34490000e-2f, -2.92380000e-2f}}};
    static const Row x51lf[5] = {
        {4u, {4.90109850e-1f, 3.77305010e-1f, -3.73106990e-1f, -1.25914530e-1f, 1.45133000e-2f}},
        {0u, {1.49085730e-1f, 3.03561680e-1f, 1.53290060e-1f, 2.45112480e-1f, -1.50753130e-1f}},
        {2u, {1.37654920e-1f, 0.00000000e+0f, 4.49417940e-1f, 0.00000000e+0f, 2.57844070e-1f}},
        {1u, {1.49085730e-1f, -3.03561680e-1f, 1.53290060e-1f, -2.45112480e-1f, -1.50753130e-1f}},
        {5u, {4.90109850e-1f, -3.77305010e-1f, -3.73106990e-1f, 1.25914530e-1f, 1.45133000e-2f}}};
    // SideLeft, FrontLeft, FrontRight, SideRight, BackCenter
    static const Row x61[5] = {
        {5u, {2.04460341e-1f, 2.17177926e-1f, -4.39996780e-2f, -2.60790269e-2f, -6.87239792e-2f}},
        {0u, {1.58923161e-1f, 9.21772680e-2f, 1.59658796e-1f, 6.66278083e-2f, 3.84686854e-2f}},
        {1u, {1.58923161e-1f, -9.21772680e-2f, 1.59658796e-1f, -6.66278083e-2f, 3.84686854e-2f}},
        {6u, {2.04460341e-1f, -2.17177926e-1f, -4.39996780e-2f, 2.60790269e-2f, -6.87239792e-2f}},
        {4u, {2.50001688e-1f, 0.00000000e+0f, -2.50000094e-1f, 0.00000000e+0f, 6.05133395e-2f}}};
    // BackLeft, SideLeft, FrontLeft, FrontRight, SideRight, BackRight (HF and LF rows are equal)
    static const Row x71[6] = {
        {4u, {1.66666667e-1f, 9.62250449e-2f, -1.66666667e-1f, -1.49071198e-1f, 8.60662966e-2f}},
        {6u, {1.66666667e-1f, 1.92450090e-1f, 0.00000000e+0f, 0.00000000e+0f, -1.72132593e-1f}},
        {0u, {1.66666667e-1f, 9.62250449e-2f, 1.66666667e-1f, 1.49071198e-1f, 8.60662966e-2f}},
        {1u, {1.66666667e-1f, -9.62250449e-2f, 1.66666667e-1f, -1.49071198e-1f, 8.60662966e-2f}},
        {7u, {1.66666667e-1f, -1.92450090e-1f, 0.00000000e+0f, 0.00000000e+0f, -1.72132593e-1f}},
        {5u, {1.66666667e-1f, -9.62250449e-2f, -1.66666667e-1f, 1.49071198e-1f, 8.60662966e-2f}}};
    Config cfg;
    switch(layout)
    {
    case 0u: cfg = Config{0, 1, 1, false, {1.0f, 1.0f, 1.0f}, {1.0f, 1.0f, 1.0f}, false, mono, nullptr}; break;
    case 1u: cfg = Config{1, 2, 2, false, {1.0f, 1.0f, 1.0f}, {1.0f, 1.0f, 1.0f}, false, stereo, nullptr}; break;
    case 2u: cfg = Config{1, 4, 4, false, {1.41421356e+0f, 1.0f, 1.0f}, {1.0f, 1.0f, 1.0f}, true, quad, quad}; break;
    case 3u: cfg = Config{2, 6, 5, true, {1.0f, 1.0f, 1.0f}, {1.0f, 1.0f, 1.0f}, true, x51hf, x51lf}; break;
    case 4u: cfg = Config{2, 7, 5, false, {1.0f, 1.0f, 1.0f}, {1.0f, 1.0f, 1.0f}, false, x61, nullptr}; break;
    case 5u: cfg = Config{2, 8, 6, false, {1.41421356e+0f, 1.22474487e+0f, 7.07106781e-1f}, {1.0f, 1.0f, 1.0f}, true, x71, x71}; break;
    default: return B200MIX_ERR_INVALID;
    }
    const uint32_t ambicount = cfg.order*2u + 1u;                 // Ambi2DChannelsFromOrder
    static const uint32_t acn2d[5] = {0u, 1u, 3u, 4u, 8u};        // AmbiIndex::FromACN2D
    static const uint32_t order2d[5] = {0u, 1u, 1u, 2u, 2u};      // AmbiIndex::OrderFrom2DChannel
    // AmbiScale::FromFuMa at ACN 0,1,3,4,8 (core/ambidefs.h:84-92); N3D is all ones
    static const float fuma[5] = {1.414213562f, 1.732050808f, 1.732050808f, 1.936491673f, 1.936491673f};
    out->ambi_order = cfg.order; out->is_2d = 1u;
    out->dry_channels = ambicount; out->real_channels = cfg.real_channels;
    for(uint32_t k = 0;k < 5u;++k) { out->map_scale[k] = 0.0f; out->map_index[k] = 0u; }
    for(uint32_t k = 0;k < ambicount;++k)
    { out->map_scale[k] = 1.0f/(cfg.fuma ? fuma[k] : 1.0f); out->map_index[k] = acn2d[k]; }
    const bool dual = hq_mode && cfg.hasLF;
    out->dual_band = dual ? 1u : 0u;
    for(uint32_t k = 0;k < 5u*8u;++k) { out->gains_hf[k] = 0.0f; out->gains_lf[k] = 0.0f; }
    const uint32_t nr = cfg.real_channels;
    for(uint32_t r = 0;r < cfg.nrows;++r)
        for(uint32_t k = 0;k < ambicount;++k)
        {
            out->gains_hf[k*nr + cfg.hf[r].real_index] = cfg.hf[r].c[k] * cfg.orderHF[order2d[k]];
            if(dual) out->gains_lf[k*nr + cfg.lf[r].real_index] = cfg.lf[r].c[k] * cfg.orderLF[order2d[k]];
        }
    out->xover_coeff = 0.0f;
    if(dual)
    {
        const float f0norm = 400.0f / float(sample_rate);
        const float w = 3.14159265358979323846f*2.0f * std::min(f0norm, 0.49f);
        const float cw = std::cos(w);
        out->xover_coeff = cw > 1.1920928955078125e-7f ? (std::sin(w) - 1.0f) / cw : cw * -0.5f;
    }
    return B200MIX_OK;
}

} // extern "C"
