// efx_math.hpp — EffectState::deviceUpdate + update of the EFX effects behind b200mix_slot_efx:
// effect PROPERTIES -> the coefficients, tap offsets and gain targets process() consumes
// (alc/effects/{echo,modulator,equalizer,compressor,dedicated,distortion}.cpp).  Host arithmetic,
// the reference's own float expressions and <cmath> calls (param_math.hpp's biquad designs), so the
// values are bit-identical to the reference's; the GPU kernels (efx_kernels.cu) only run process().
// Shared with the test oracle (oracle/efx_oracle.cpp), which keeps its own process().
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>

#include "param_math.hpp"

namespace b200mix {

constexpr uint32_t kEfxMaxLines = 16;     // output lines / input channels handled per slot

// What update() leaves behind (parameters only; state lives with the process side).
struct EfxParams {
    uint32_t type, in_channels, lines;
    uint32_t fade_len;                    // MixSamples Counter of the output mix: 0 = samplesToDo, else min(n, fade_len)
    uint32_t snap_gains;                  // 1: the gain applies at once (compressor: no Current/Target fade)
    uint32_t line_on[kEfxMaxLines];       // line feeds an output channel (mTargetChannel != InvalidChannelIndex)
    float gains[kEfxMaxLines][32];        // target gains [line][output channel of the target mix]
    // echo
    uint32_t echo_tap[2], echo_len; float echo_filter[5], echo_feed;
    // modulator
    uint32_t mod_range_new;               // the range update() rescales mIndex with (before the square's rounding)
    uint32_t mod_range, mod_wave; float mod_scale, mod_hp[5];      // wave: 0 one, 1 sin, 2 saw, 3 square
    // equalizer
    float eq[4][5];
    // compressor
    uint32_t comp_enabled; float comp_attack, comp_release;
    // distortion
    float dist_edge, dist_lp[5], dist_bp[5];
    // chorus / flanger
    uint32_t cho_len;                     // delay buffer length per line (power of two)
    uint32_t cho_wave; int32_t cho_delay; float cho_depth, cho_feedback;
    uint32_t cho_lfo_range_new, cho_lfo_range, cho_lfo_disp, cho_rate_on; float cho_lfo_scale;
    // autowah
    float wah_attack, wah_release, wah_res_gain, wah_peak_gain, wah_freq_min, wah_bandwidth;
    // vocal morpher
    uint32_t vm_step, vm_wave;            // wave: 0 half, 1 sin, 2 triangle, 3 saw (Oscillate<>, vmorpher.cpp:73-96)
    float vm_coeff[2][4], vm_fgain[2][4]; // FormantFilter::mCoeff / mGain of vowel A and B
    uint32_t vm_target[kEfxMaxLines];     // mTargetChannel per wet channel (0xffffffff: none)
    float vm_tgain[kEfxMaxLines];         // mTargetGain
    // frequency shifter
    uint32_t fs_phase_step[4]; float fs_sign[4]; uint32_t fs_reset_phase[4];   // ProcessParams::mPhaseStep / mSign; Off: mPhase = 0
    // pitch shifter
    uint32_t ps_shift_i; float ps_shift;  // PshifterState::mPitchShiftI (16.16) / mPitchShift
};

namespace efx_detail {
struct HostMath {
    static float sqrt(float x) { return std::sqrt(x); }
    static float sin(float x) { return std::sin(x); }
    static float cos(float x) { return std::cos(x); }
    static float sinh(float x) { return std::sinh(x); }
    static float log(float x) { return std::log(x); }
};
// float2int (common/alnumeric.h:195-222): truncation, out-of-range values clamped
inline int f2i(float f)
{
    if(!(f < 2147483648.0f)) return 2147483647;
    if(!(f > -2147483648.0f)) return -2147483647 - 1;
    return static_cast<int>(f);
}
inline uint32_t next_pow2(uint32_t v)     // NextPowerOf2, common/alnumeric.h:121-135
{
    if(v > 0) { v--; v |= v>>1; v |= v>>2; v |= v>>4; v |= v>>8; v |= v>>16; }
    return v + 1u;
}
inline uint32_t f2u(float f) { return static_cast<uint32_t>(static_cast<int32_t>(f)); }   // float2uint
// MixParams::setAmbiMixParams (core/device.h:126-147): wet channel i -> the output channel carrying
// the same ambisonic index
inline void ambi_mix_params(const b200mix_efx_target &T, float gainbase, uint32_t max_lines, EfxParams &P)
{
    for(uint32_t i = 0;i < T.wet_channels && i < max_lines;++i)
    {
        P.line_on[i] = 0u;
        for(uint32_t j = 0;j < T.out_channels;++j)
            if(T.out_index[j] == T.wet_index[i])
            {
                P.line_on[i] = 1u;
                P.gains[i][j] = T.out_scale[j] * gainbase;
                break;
            }
    }
}
} // namespace efx_detail

// Returns B200MIX_OK / _ERR_INVALID / _ERR_UNSUPPORTED.
inline int efx_update(const b200mix_efx_props &E, const b200mix_efx_target &T, EfxParams &P)
{
    using namespace efx_detail;
    std::memset(&P, 0, sizeof(P));
    if(T.out_channels > 32u || T.wet_channels < 1u || !T.out_scale || !T.out_index || !T.wet_index || !T.sample_rate)
        return B200MIX_ERR_INVALID;
    if(T.wet_channels > kEfxMaxLines) return B200MIX_ERR_UNSUPPORTED;
    const float frequency = static_cast<float>(T.sample_rate);
    P.type = E.type; P.in_channels = T.wet_channels;
    switch(E.type)
    {
    case B200MIX_EFFECT_ECHO:
    {
        // EchoState::deviceUpdate / update (echo.cpp:82-131)
        P.echo_len = next_pow2(f2u(0.207f*frequency + 0.5f) + f2u(0.404f*frequency + 0.5f));
        P.echo_tap[0] = std::max(f2u(std::round(E.echo.delay*frequency)), 1u);
        P.echo_tap[1] = f2u(std::round(E.echo.lr_delay*frequency)) + P.echo_tap[0];
        const float gainhf = std::max(1.0f - E.echo.damping, 0.0625f);
        pm::biquad_coeffs<HostMath>(0u, 5000.0f/frequency, gainhf, 1.0f, P.echo_filter);
        P.echo_feed = E.echo.feedback;
        const float x = E.echo.spread, z = std::sqrt(1.0f - x*x);
        float c0[B200MIX_MAX_AMBI_CHANNELS], c1[B200MIX_MAX_AMBI_CHANNELS];
        pm::sh_n3d(x, 0.0f, z, c0);           // CalcAmbiCoeffs( x, 0, z, 0)
        pm::sh_n3d(-x, 0.0f, z, c1);          // CalcAmbiCoeffs(-x, 0, z, 0)
        P.lines = 2u; P.line_on[0] = P.line_on[1] = 1u;
        if(!pm::pan_gains(T.out_channels, T.out_scale, T.out_index, c0, T.slot_gain, P.gains[0], 32u)
            || !pm::pan_gains(T.out_channels, T.out_scale, T.out_index, c1, T.slot_gain, P.gains[1], 32u))
            return B200MIX_ERR_INVALID;
        break;
    }
    case B200MIX_EFFECT_MODULATOR:
    {
        // ModulatorState::update (modulator.cpp:101-155)
        const float samplesPerCycle = E.modulator.frequency > 0.0f ? frequency/E.modulator.frequency + 0.5f : 1.0f;
        const uint32_t range = static_cast<uint32_t>(std::clamp(samplesPerCycle, 1.0f, frequency));
        P.mod_range_new = range; P.mod_range = range;
        if(range == 1u) { P.mod_scale = 0.0f; P.mod_wave = 0u; }
        else if(E.modulator.waveform == 0u)
        { P.mod_scale = 3.14159265358979323846f*2.0f / static_cast<float>(range); P.mod_wave = 1u; }
        else if(E.modulator.waveform == 1u)
        { P.mod_scale = 2.0f / static_cast<float>(range-1u); P.mod_wave = 2u; }
        else if(E.modulator.waveform == 2u)
        {
            P.mod_range = (range+1u) & ~1u;
            P.mod_scale = 1.0f / static_cast<float>(P.mod_range-1u);
            P.mod_wave = 3u;
        }
        else return B200MIX_ERR_INVALID;
        const float f0norm = std::clamp(E.modulator.high_pass_cutoff / frequency, 1.0f/512.0f, 0.49f);
        pm::biquad_coeffs_bandwidth<HostMath>(4u, f0norm, 1.0f, 0.75f, P.mod_hp);
        P.lines = T.wet_channels; P.fade_len = 64u;
        ambi_mix_params(T, T.slot_gain, kEfxMaxLines, P);
        break;
    }
    case B200MIX_EFFECT_EQUALIZER:
    {
        // EqualizerState::update (equalizer.cpp:117-163)
        float gain = std::sqrt(E.equalizer.low_gain);
        pm::biquad_coeffs<HostMath>(1u, E.equalizer.low_cutoff/frequency, gain, 0.75f, P.eq[0]);
        gain = std::sqrt(E.equalizer.mid1_gain);
        pm::biquad_coeffs_bandwidth<HostMath>(2u, E.equalizer.mid1_center/frequency, gain, E.equalizer.mid1_width, P.eq[1]);
        gain = std::sqrt(E.equalizer.mid2_gain);
        pm::biquad_coeffs_bandwidth<HostMath>(2u, E.equalizer.mid2_center/frequency, gain, E.equalizer.mid2_width, P.eq[2]);
        gain = std::sqrt(E.equalizer.high_gain);
        pm::biquad_coeffs<HostMath>(0u, E.equalizer.high_cutoff/frequency, gain, 0.75f, P.eq[3]);
        P.lines = T.wet_channels;
        ambi_mix_params(T, T.slot_gain, kEfxMaxLines, P);
        break;
    }
    case B200MIX_EFFECT_COMPRESSOR:
    {
        // CompressorState::deviceUpdate / update (compressor.cpp:80-109)
        const float attackCount = frequency * 0.1f, releaseCount = frequency * 0.2f;
        P.comp_attack = std::pow(2.0f/0.5f, 1.0f/attackCount);
        P.comp_release = std::pow(0.5f/2.0f, 1.0f/releaseCount);
        P.comp_enabled = E.compressor.on_off ? 1u : 0u;
        P.lines = T.wet_channels; P.snap_gains = 1u;
        ambi_mix_params(T, T.slot_gain, kEfxMaxLines, P);
        break;
    }
    case B200MIX_EFFECT_DEDICATED:
    {
        // DedicatedState::update (dedicated.cpp:67-103)
        const float Gain = T.slot_gain * E.dedicated.gain;
        P.lines = 1u; P.line_on[0] = 1u;
        if(E.dedicated.target == 0u)
        {
            if(T.real_center != B200MIX_NO_SLOT) return B200MIX_ERR_UNSUPPORTED;
            float c[B200MIX_MAX_AMBI_CHANNELS];
            pm::sh_n3d(-0.0f, 0.0f, 1.0f, c);        // CalcDirectionCoeffs({0, 0, -1})
            if(!pm::pan_gains(T.out_channels, T.out_scale, T.out_index, c, Gain, P.gains[0], 32u))
                return B200MIX_ERR_INVALID;
        }
        else if(T.real_lfe != B200MIX_NO_SLOT) return B200MIX_ERR_UNSUPPORTED;
        break;
    }
    case B200MIX_EFFECT_DISTORTION:
    {
        // DistortionState::update (distortion.cpp:140-196), first-order devices (no up-sampler)
        if(T.device_ambi_order > 1u) return B200MIX_ERR_UNSUPPORTED;
        const float edge = std::min(std::sin(3.14159265358979323846f*0.5f * E.distortion.edge), 0.99f);
        P.dist_edge = 2.0f * edge / (1.0f-edge);
        float cutoff = E.distortion.lowpass_cutoff;
        float bandwidth = 0.746268656716f;
        pm::biquad_coeffs_bandwidth<HostMath>(3u, cutoff/frequency*0.25f, 1.0f, bandwidth, P.dist_lp);
        cutoff = E.distortion.eq_center;
        bandwidth = E.distortion.eq_bandwidth / (cutoff * 0.67f);
        pm::biquad_coeffs_bandwidth<HostMath>(5u, cutoff/frequency*0.25f, 1.0f, bandwidth, P.dist_bp);
        P.lines = 4u;
        ambi_mix_params(T, T.slot_gain*E.distortion.gain, 4u, P);
        break;
    }
    case B200MIX_EFFECT_CHORUS:
    {
        // ChorusState::deviceUpdate / update (chorus.cpp:132-232), first-order devices
        if(T.device_ambi_order > 1u) return B200MIX_ERR_UNSUPPORTED;
        P.cho_len = next_pow2(f2u(0.016f*2.0f*frequency) + 1u);           // max(ChorusMaxDelay, FlangerMaxDelay)
        constexpr int mindelay = 24 << 8;                                  // MaxResamplerEdge << sTableBits
        if(E.chorus.waveform > 1u) return B200MIX_ERR_INVALID;
        P.cho_wave = E.chorus.waveform;
        const float stepscale = frequency * 256.0f;
        P.cho_delay = std::max(f2i(std::round(E.chorus.delay * stepscale)), mindelay);
        P.cho_depth = std::min(static_cast<float>(P.cho_delay) * E.chorus.depth,
            static_cast<float>(P.cho_delay - mindelay));
        P.cho_feedback = E.chorus.feedback;
        if(!(E.chorus.rate > 0.0f))
        { P.cho_rate_on = 0u; P.cho_lfo_range = 1u; P.cho_lfo_range_new = 1u; P.cho_lfo_scale = 0.0f; P.cho_lfo_disp = 0u; }
        else
        {
            constexpr int range_limit = 2147483647/360 - 180;
            const float range = std::round(frequency / E.chorus.rate);
            const uint32_t lfo_range = f2u(std::min(range, float(range_limit)));
            P.cho_rate_on = 1u; P.cho_lfo_range = lfo_range; P.cho_lfo_range_new = lfo_range;
            P.cho_lfo_scale = (P.cho_wave == 1u ? 4.0f : 3.14159265358979323846f*2.0f) / static_cast<float>(lfo_range);
            int phase = E.chorus.phase;
            if(phase < 0) phase += 360;
            P.cho_lfo_disp = (lfo_range*static_cast<uint32_t>(phase) + 180u) / 360u;
        }
        P.lines = 4u;
        ambi_mix_params(T, T.slot_gain, 4u, P);
        break;
    }
    case B200MIX_EFFECT_AUTOWAH:
    {
        // AutowahState::update (autowah.cpp:110-134)
        const float ReleaseTime = std::clamp(E.autowah.release_time, 0.001f, 1.0f);
        P.wah_attack = std::exp(-1.0f / (E.autowah.attack_time*frequency));
        P.wah_release = std::exp(-1.0f / (ReleaseTime*frequency));
        P.wah_res_gain = std::sqrt(std::log10(E.autowah.resonance)*10.0f / 3.0f);
        P.wah_peak_gain = 1.0f - std::log10(E.autowah.peak_gain / 31621.0f);
        P.wah_freq_min = 20.0f / frequency;
        P.wah_bandwidth = (2500.0f-20.0f) / frequency;
        P.lines = T.wet_channels;
        ambi_mix_params(T, T.slot_gain, kEfxMaxLines, P);
        break;
    }
    case B200MIX_EFFECT_VMORPHER:
    {
        // VmorpherState::update (vmorpher.cpp:232-270)
        const float step = E.vmorpher.rate / frequency;
        P.vm_step = static_cast<uint32_t>(std::lrint(std::clamp(step*16777216.0f, 0.0f, 16777216.0f-1.0f)));   // fastf2u
        P.vm_wave = P.vm_step == 0u ? 0u : (E.vmorpher.waveform == 0u ? 1u : (E.vmorpher.waveform == 1u ? 2u : 3u));
        const float pitch[2] = {std::pow(2.0f, static_cast<float>(E.vmorpher.phoneme_a_coarse_tuning) / 12.0f),
                                std::pow(2.0f, static_cast<float>(E.vmorpher.phoneme_b_coarse_tuning) / 12.0f)};
        // getFiltersByPhoneme (vmorpher.cpp:169-225): soprano formants of A E I O U
        static const float kFreq[5][4] = {{800, 1150, 2900, 3900}, {350, 2000, 2800, 3600}, {270, 2140, 2950, 3900},
                                          {450, 800, 2830, 3800}, {325, 700, 2700, 3800}};
        static const float kGain[5][4] = {{1.000000f, 0.501187f, 0.025118f, 0.100000f}, {1.000000f, 0.100000f, 0.177827f, 0.009999f},
                                          {1.000000f, 0.251188f, 0.050118f, 0.050118f}, {1.000000f, 0.281838f, 0.079432f, 0.079432f},
                                          {1.000000f, 0.158489f, 0.017782f, 0.009999f}};
        const uint32_t ph[2] = {E.vmorpher.phoneme_a, E.vmorpher.phoneme_b};
        for(int v = 0;v < 2;++v)
            for(int f = 0;f < 4;++f)
            {
                if(ph[v] < 5u)
                {
                    const float f0norm = (kFreq[ph[v]][f] * pitch[v]) / frequency;
                    P.vm_coeff[v][f] = std::tan(3.14159265358979323846f * f0norm);
                    P.vm_fgain[v][f] = kGain[ph[v]][f];
                }
                else { P.vm_coeff[v][f] = 0.0f; P.vm_fgain[v][f] = 1.0f; }      // FormantFilter{}
            }
        // the kernel applies the per-chunk gain ramp itself (MixSamples once per 256 samples,
        // vmorpher.cpp:317-318): the generic output mix gets unit gains, set at once
        P.lines = T.wet_channels; P.snap_gains = 1u;
        for(uint32_t i = 0;i < T.wet_channels;++i)
        {
            P.vm_target[i] = 0xffffffffu; P.vm_tgain[i] = 0.0f;
            for(uint32_t j = 0;j < T.out_channels;++j)
                if(T.out_index[j] == T.wet_index[i])
                {
                    P.vm_target[i] = j; P.vm_tgain[i] = T.out_scale[j] * T.slot_gain;
                    P.line_on[i] = 1u; P.gains[i][j] = 1.0f;
                    break;
                }
        }
        break;
    }
    case B200MIX_EFFECT_FSHIFTER:
    {
        // FshifterState::update (fshifter.cpp:171-233); the up-sampler of higher-order devices (:150-168) is not built
        if(T.device_ambi_order > 1u) return B200MIX_ERR_UNSUPPORTED;
        const float step = E.fshifter.frequency / frequency;
        const uint32_t pstep = static_cast<uint32_t>(std::lrint(std::min(step, 1.0f) * 65536.0f));    // fastf2u
        const uint32_t dir[4] = {E.fshifter.left_direction, E.fshifter.left_direction,
                                 E.fshifter.right_direction, E.fshifter.right_direction};
        for(int c = 0;c < 4;++c)
        {
            P.fs_phase_step[c] = pstep; P.fs_sign[c] = 1.0f; P.fs_reset_phase[c] = 0u;
            if(dir[c] == 0u) P.fs_sign[c] = -1.0f;
            else if(dir[c] == 2u) { P.fs_phase_step[c] = 0u; P.fs_reset_phase[c] = 1u; }   // with phase 0 the sign is moot
        }
        P.lines = 4u;
        ambi_mix_params(T, T.slot_gain, 4u, P);
        break;
    }
    case B200MIX_EFFECT_PSHIFTER:
    {
        // PshifterState::update (pshifter.cpp:170-205); the up-sampler of devices above second order
        // (:150-167, :445-461) is not built
        if(T.device_ambi_order > 2u) return B200MIX_ERR_UNSUPPORTED;
        const int tune = E.pshifter.coarse_tune*100 + E.pshifter.fine_tune;
        const float pitch = std::pow(2.0f, static_cast<float>(tune) / 1200.0f);
        P.ps_shift_i = static_cast<uint32_t>(std::lrint(std::min(std::max(pitch, 0.5f), 2.0f) * 65536.0f));   // fastf2u
        P.ps_shift = static_cast<float>(P.ps_shift_i) * (1.0f/65536.0f);
        P.lines = std::min(T.wet_channels, 9u);                       // NumLines: second order
        ambi_mix_params(T, T.slot_gain, 9u, P);
        break;
    }
    default: return B200MIX_ERR_INVALID;
    }
    return B200MIX_OK;
}

} // namespace b200mix
