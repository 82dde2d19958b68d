// Host-side parameter stage of the EAX reverb (no GPU): from the effect's properties to the
// b200mix_reverb_params + output gains the mixer consumes.  Restates, operation for
// operation, ReverbState::deviceUpdate / allocLines (alc/effects/reverb.cpp:728-851) and
// ReverbState::update with its helpers (:853-1351), so a host without the reference's
// ReverbState gets bit-identical parameters (tests/test_reverb_params.py pins every field
// against the compiled reference through oracle/ref_reverb_tap.cpp).
#include "../../include/b200mix.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>

namespace {

constexpr int kLines = 4;
// line geometry, seconds (alc/effects/reverb.cpp:185-252)
constexpr float kEarlyTap[kLines] = {0.000000e+0f, 1.010676e-3f, 2.126553e-3f, 3.358580e-3f};
constexpr float kEarlyAllpass[kLines] = {4.854840e-4f, 5.360178e-4f, 5.918117e-4f, 6.534130e-4f};
constexpr float kEarlyLine[kLines] = {2.992520e-3f, 5.456575e-3f, 7.688329e-3f, 9.709681e-3f};
constexpr float kLateAllpass[kLines] = {8.091400e-4f, 1.019453e-3f, 1.407968e-3f, 1.618280e-3f};
constexpr float kLateLine[kLines] = {9.709681e-3f, 1.223343e-2f, 1.689561e-2f, 1.941936e-2f};

constexpr float kDecayGain = 0.001f;             // ReverbDecayGain, core/effects/base.h:22
constexpr float kMaxReflectionsDelay = 0.3f, kMaxLateReverbDelay = 0.1f;   // :24-25
constexpr float kMinDecayTime = 0.1f, kMaxDecayTime = 20.0f;               // reverb.cpp:54-55
constexpr float kMaxModulationTime = 4.0f, kDefaultModulationTime = 0.25f; // :56-57
constexpr float kMaxHFReference = 20000.0f;                                // :58
constexpr float kModDepthCoeff = 0.05f;                                    // :84
constexpr float kModFracOne = float(1 << 24);                              // MOD_FRACONE :60-61
constexpr float kDensityScale = 1000.0f;                                   // :139
constexpr float kSpeedOfSound = 343.3f;                                    // core/context.h:32
constexpr unsigned kBufferLine = 1024u, kMaxUpdate = 256u;
constexpr float kInvSqrt2 = static_cast<float>(1.0/1.41421356237309504880);
constexpr float kSqrt3 = 1.7320508075688772935f;

// std::reduce over four floats as libstdc++ evaluates it: init + ((a0 + a1) + (a2 + a3))
constexpr float reduce4(const float (&a)[kLines]) { return 0.0f + ((a[0] + a[1]) + (a[2] + a[3])); }
constexpr float kLateAllpassAverage = reduce4(kLateAllpass) / float(kLines);
constexpr float kLateDelayAverage = reduce4(kLateLine) / float(kLines) + kLateAllpassAverage;

// float2uint (common/alnumeric.h:223-241): truncation, negative -> 0
unsigned f2u(float f) { return f > 0.0f ? (f >= 4294967296.0f ? 0xffffffffu : unsigned(f)) : 0u; }
// fastf2u: current rounding mode (nearest even)
unsigned fastf2u(float f) { return unsigned(std::lrintf(f)); }

unsigned next_pow2(unsigned v)
{
    if(v > 0) { v--; v |= v>>1; v |= v>>2; v |= v>>4; v |= v>>8; v |= v>>16; }
    return v + 1;
}
// DelayLine*::calcLineLength (:273-286,313-319), per line
unsigned line_length(float length, float frequency, unsigned extra)
{ return next_pow2(f2u(std::ceil(length*frequency)) + extra); }

float delay_length_mult(float density) { return std::max(1.0f, std::cbrt(density*kDensityScale)); }   // :722-723
float decay_coeff(float length, float decayTime) { return std::pow(kDecayGain, length/decayTime); }   // :861-862
float decay_length(float coeff, float decayTime) { return std::log10(coeff) * decayTime / -3.0f; }    // :867-871
float lerpf(float a, float b, float mu) { return a + (b-a)*mu; }

// BandSplitter::init (core/filters/splitter.cpp:15-26)
float splitter_coeff(float f0norm)
{
    const float w = 3.14159265358979323846f*2.0f * std::min(f0norm, 0.49f);
    const float cw = std::cos(w);
    if(cw > 1.1920928955078125e-7f) return (std::sin(w) - 1.0f) / cw;
    return cw * -0.5f;
}

// HF order scales, core/ambidefs.cpp:37-56
constexpr float kHFScales[5][5] = {
    {4.000000000e+00f, 2.309401077e+00f, 1.192569588e+00f, 7.189495850e-01f, 4.784482742e-01f},
    {4.000000000e+00f, 2.309401077e+00f, 1.192569588e+00f, 7.189495850e-01f, 4.784482742e-01f},
    {2.981423970e+00f, 2.309401077e+00f, 1.192569588e+00f, 7.189495850e-01f, 4.784482742e-01f},
    {2.359168820e+00f, 2.031565936e+00f, 1.444598386e+00f, 7.189495850e-01f, 4.784482742e-01f},
    {1.947005434e+00f, 1.764337084e+00f, 1.424707344e+00f, 9.755104127e-01f, 4.784482742e-01f}};
constexpr float kHFScales2D[5][5] = {
    {2.236067977e+00f, 1.581138830e+00f, 9.128709292e-01f, 6.050756345e-01f, 4.370160244e-01f},
    {2.236067977e+00f, 1.581138830e+00f, 9.128709292e-01f, 6.050756345e-01f, 4.370160244e-01f},
    {1.825741858e+00f, 1.581138830e+00f, 9.128709292e-01f, 6.050756345e-01f, 4.370160244e-01f},
    {1.581138830e+00f, 1.460781803e+00f, 1.118033989e+00f, 6.050756345e-01f, 4.370160244e-01f},
    {1.414213562e+00f, 1.344997024e+00f, 1.144122806e+00f, 8.312538756e-01f, 4.370160244e-01f}};

// A-format to B-format for the early and late outputs (:104-122)
constexpr float kEarlyA2B[4][4] = {
    {0.5f,  0.5f,  0.5f,  0.5f}, {0.5f, -0.5f,  0.5f, -0.5f}, {0.5f, -0.5f, -0.5f,  0.5f}, {0.5f,  0.5f, -0.5f, -0.5f}};
constexpr float kLateA2B[4][4] = {
    {0.5f, 0.5f, 0.5f, 0.5f}, {kInvSqrt2, -kInvSqrt2, 0.0f, 0.0f}, {0.0f, 0.0f, -kInvSqrt2, kInvSqrt2}, {0.5f, 0.5f, -0.5f, -0.5f}};

// AmbiScale::FirstOrderUp (core/ambidefs.cpp:65-84,281-307): a first-order decode to the eight
// cube corners re-encoded at full order, products summed in double
void first_order_up(float up[4][B200MIX_MAX_AMBI_CHANNELS])
{
    const float s = 0.57735026918962576451f;      // inv_sqrt3f
    // corner k: signs of (y, z, x), the order of the decoder rows [W, Y, Z, X]
    static const int sign[8][3] = {{1,1,1},{1,1,-1},{-1,1,1},{-1,1,-1},{1,-1,1},{1,-1,-1},{-1,-1,1},{-1,-1,-1}};
    float enc[8][B200MIX_MAX_AMBI_CHANNELS];
    for(int k = 0;k < 8;++k)
    {
        // CalcAmbiCoeffs(y, z, x) == b200mix_ambi_coeffs({-y, z, -x}) (core/mixer.h:68-73)
        const float dir[3] = {-(sign[k][0]*s), sign[k][1]*s, -(sign[k][2]*s)};
        b200mix_ambi_coeffs(dir, 0.0f, enc[k]);
    }
    for(int i = 0;i < 4;++i)
        for(unsigned j = 0;j < B200MIX_MAX_AMBI_CHANNELS;++j)
        {
            double sum = 0.0;
            for(int k = 0;k < 8;++k)
            {
                const float dec = (i == 0) ? 0.125f : 0.125f*float(sign[k][i-1]);
                sum += double(dec) * enc[k][j];
            }
            up[i][j] = float(sum);
        }
}

// GetTransformFromVector (:1106-1148), transposed like the reference's
void transform_from_vector(const float vec[3], float m[4][4])
{
    float norm[3] = {vec[0], vec[1], vec[2]};
    float mag = std::sqrt(vec[0]*vec[0] + vec[1]*vec[1] + vec[2]*vec[2]);
    if(mag > 1.0f)
    {
        const float scale = kSqrt3 / mag;
        norm[0] *= -scale; norm[1] *= scale; norm[2] *= scale;
        mag = 1.0f;
    }
    else
    { norm[0] *= -kSqrt3; norm[1] *= kSqrt3; norm[2] *= kSqrt3; }
    const float r[4][4] = {
        {1.0f, norm[0], norm[1], norm[2]},
        {0.0f, 1.0f-mag, 0.0f, 0.0f},
        {0.0f, 0.0f, 1.0f-mag, 0.0f},
        {0.0f, 0.0f, 0.0f, 1.0f-mag}};
    std::memcpy(m, r, sizeof(r));
}

// get_coeffs of update3DPanning (:1161-1206): per output line the 25 encoder coefficients
void line_coeffs(const float a2b[4][4], const float matrix[4][4], bool upmix,
    const float up[4][B200MIX_MAX_AMBI_CHANNELS], float res[4][B200MIX_MAX_AMBI_CHANNELS])
{
    std::memset(res, 0, sizeof(float)*4*B200MIX_MAX_AMBI_CHANNELS);
    if(upmix)
    {
        for(int i = 0;i < 4;++i)
            for(int j = 0;j < 4;++j)
            {
                const float a = matrix[i][j];
                for(unsigned k = 0;k < B200MIX_MAX_AMBI_CHANNELS;++k) res[i][k] = a*up[j][k] + res[i][k];
            }
    }
    else
    {
        for(int i = 0;i < 4;++i)
            for(int j = 0;j < 4;++j)
            {
                const float a = a2b[j][i];
                for(int k = 0;k < 4;++k) res[i][k] = a*matrix[j][k] + res[i][k];
            }
    }
}

} // namespace

extern "C" {

int b200mix_reverb_full_update_needed(const b200mix_efx_reverb *prev, const b200mix_efx_reverb *next)
{
    if(!next) return B200MIX_ERR_INVALID;
    if(!prev) return 1;
    auto times = [](const b200mix_efx_reverb &p, float &lf, float &hf)
    {
        float hfRatio = p.decay_hf_ratio;
        if(p.decay_hf_limit && p.air_absorption_gain_hf < 1.0f)
            hfRatio = std::min(1.0f / kSpeedOfSound / decay_length(p.air_absorption_gain_hf, p.decay_time), hfRatio);
        lf = std::clamp(p.decay_time*p.decay_lf_ratio, kMinDecayTime, kMaxDecayTime);
        hf = std::clamp(p.decay_time*hfRatio, kMinDecayTime, kMaxDecayTime);
    };
    float lf0, hf0, lf1, hf1;
    times(*prev, lf0, hf0); times(*next, lf1, hf1);
    return (prev->density != next->density || prev->diffusion != next->diffusion
        || prev->decay_time != next->decay_time || hf0 != hf1 || lf0 != lf1
        || prev->modulation_time != next->modulation_time || prev->modulation_depth != next->modulation_depth
        || prev->hf_reference != next->hf_reference || prev->lf_reference != next->lf_reference) ? 1 : 0;
}

int b200mix_reverb_params_from_efx(const b200mix_efx_reverb *props, const b200mix_reverb_target *tgt,
    b200mix_reverb_params *out, float *gains)
{
    if(!props || !tgt || !out || props->struct_size != sizeof(*props) || tgt->struct_size != sizeof(*tgt)
        || tgt->sample_rate == 0 || tgt->device_ambi_order > 4u
        || tgt->out_channels > B200MIX_MAX_DRY_CHANNELS || (gains && (!tgt->out_scale || !tgt->out_index)))
        return B200MIX_ERR_INVALID;
    const b200mix_efx_reverb &P = *props;
    const float frequency = float(tgt->sample_rate);
    std::memset(out, 0, sizeof(*out));
    out->struct_size = sizeof(*out);

    // ---- allocLines (:728-820), per-line lengths
    {
        const float multiplier = delay_length_mult(1.0f);
        const float max_mod_delay = kMaxModulationTime*kModDepthCoeff / 2.0f;
        const unsigned late_vecap_extra = f2u(std::ceil(kLateAllpass[0] * multiplier * frequency));
        out->main_len = line_length(kMaxReflectionsDelay + kEarlyTap[3]*multiplier, frequency, kBufferLine);
        const float LateDiffAvg = (kLateLine[3] - kLateLine[0]) / float(kLines);
        out->late_in_len = line_length(kMaxLateReverbDelay + LateDiffAvg*multiplier, frequency, kBufferLine);
        out->early_ap_len = line_length(kEarlyAllpass[3] * multiplier, frequency, 0u);
        out->early_len = line_length(kEarlyLine[3] * multiplier, frequency, kMaxUpdate);
        out->late_ap_len = line_length(kLateAllpass[3] * multiplier, frequency, late_vecap_extra);
        out->late_len = line_length(kLateLine[3]*multiplier + max_mod_delay, frequency, 4u);
    }
    // ---- deviceUpdate (:834-850)
    out->upmix = tgt->device_ambi_order > 1u ? 1u : 0u;
    if(out->upmix)
    {
        const auto &tab = tgt->device_2d ? kHFScales2D : kHFScales;
        out->order_scale[0] = tab[1][0] / tab[tgt->device_ambi_order][0];
        out->order_scale[1] = tab[1][1] / tab[tgt->device_ambi_order][1];
    }
    else
        out->order_scale[0] = out->order_scale[1] = 1.0f;
    out->splitter_coeff = splitter_coeff(tgt->xover_freq / frequency);

    // ---- update (:1222-1351)
    float hfRatio = P.decay_hf_ratio;
    if(P.decay_hf_limit && P.air_absorption_gain_hf < 1.0f)
        hfRatio = std::min(1.0f / kSpeedOfSound / decay_length(P.air_absorption_gain_hf, P.decay_time), hfRatio);
    const float lfDecayTime = std::clamp(P.decay_time*P.decay_lf_ratio, kMinDecayTime, kMaxDecayTime);
    const float hfDecayTime = std::clamp(P.decay_time*hfRatio, kMinDecayTime, kMaxDecayTime);
    const float density_mult = delay_length_mult(P.density);

    // updateDelayLine (:1071-1098)
    out->early_tap_coeff = P.gain;
    for(int j = 0;j < kLines;++j)
    {
        out->early_tap[j] = f2u((kEarlyTap[j]*density_mult + P.reflections_delay) * frequency);
        float length = kLateLine[j] - kLateLine[0];
        length = length*(1.0f/float(kLines))*density_mult + P.late_reverb_delay;
        out->late_tap[j] = f2u(length * frequency);
    }

    // master filters (:1299-1309)
    const float hf0norm = std::min(P.hf_reference/frequency, 0.49f);
    const float lf0norm = std::min(P.lf_reference/frequency, 0.49f);
    if(b200mix_biquad_coeffs(0u /*HighShelf*/, hf0norm, P.gain_hf, 1.0f, out->filter_lp)
        || b200mix_biquad_coeffs(1u /*LowShelf*/, lf0norm, P.gain_lf, 1.0f, out->filter_hp))
        return B200MIX_ERR_INVALID;

    // EarlyReflections::updateLines (:944-965)
    out->early_ap_coeff = P.diffusion*P.diffusion * kInvSqrt2;
    for(int j = 0;j < kLines;++j)
    {
        out->early_ap_offset[j] = f2u(kEarlyAllpass[j] * density_mult * frequency);
        out->early_offset[j] = f2u(kEarlyLine[j] * density_mult * frequency);
    }
    out->early_coeff = decay_coeff(reduce4(kEarlyLine) / float(kLines) * density_mult, P.decay_time);

    // CalcMatrixCoeffs (:897-906)
    {
        const float t = P.diffusion * std::atan(kSqrt3);
        out->mix_x = std::cos(t);
        out->mix_y = std::sin(t) / kSqrt3;
    }

    // Modulation::updateModulator (:971-1002)
    out->mod_step = std::max(fastf2u(kModFracOne / (frequency * P.modulation_time)), 1u);
    if(P.modulation_time >= kDefaultModulationTime)
        out->mod_depth = kModDepthCoeff / 4.0f * kDefaultModulationTime * P.modulation_depth * frequency;
    else
        out->mod_depth = kModDepthCoeff / 4.0f * P.modulation_time * P.modulation_depth * frequency;

    // LateReverb::updateLines (:1005-1067)
    {
        const float nwf = frequency / kMaxHFReference;
        const float decayTimeWeighted = lf0norm*nwf*lfDecayTime + (hf0norm - lf0norm)*nwf*P.decay_time
            + (1.0f - hf0norm*nwf)*hfDecayTime;
        const float a = decay_coeff(kLateDelayAverage*density_mult, decayTimeWeighted);
        out->density_gain = std::sqrt(1.0f - a*a);
        out->late_ap_coeff = P.diffusion*P.diffusion * kInvSqrt2;
        float lengths[kLines];
        for(int j = 0;j < kLines;++j)
        {
            out->late_ap_offset[j] = f2u(kLateAllpass[j]*density_mult * frequency);
            lengths[j] = kLateLine[j] * density_mult;
            out->late_offset[j] = std::max(f2u(lengths[j]*frequency + 0.5f), 1u) - 1u;
        }
        const float moddepth = out->mod_depth/frequency;
        for(int j = 0;j < kLines;++j)
        {
            const float length = lerpf(kLateAllpass[j], kLateAllpassAverage, P.diffusion)*density_mult + moddepth
                + lengths[j];
            // T60Filter::calcCoeffs (:927-941)
            const float mfGain = decay_coeff(length, P.decay_time);
            const float lfGain = decay_coeff(length, lfDecayTime) / mfGain;
            const float hfGain = decay_coeff(length, hfDecayTime) / mfGain;
            out->t60_mid_gain[j] = mfGain;
            if(b200mix_biquad_coeffs(1u /*LowShelf*/, lf0norm, lfGain, 1.0f, out->t60_lf[j])
                || b200mix_biquad_coeffs(0u /*HighShelf*/, hf0norm, hfGain, 1.0f, out->t60_hf[j]))
                return B200MIX_ERR_INVALID;
        }
    }

    // fade length (:1328-1350)
    {
        const float decayBase = tgt->slot_gain * P.gain * P.late_reverb_gain;
        const float decayDiff = kDecayGain / std::max(decayBase, kDecayGain);
        const float diffTime = !(decayDiff < 1.0f) ? 0.0f : (std::log10(decayDiff)*(20.0f / -60.0f) * P.decay_time);
        const float decaySamples = (P.reflections_delay + P.late_reverb_delay + diffTime) * frequency;
        out->fade_samples = uint32_t(std::min(decaySamples, 100000.0f));
    }

    // update3DPanning (:1151-1220) -> gains[8][out_channels]
    if(gains)
    {
        float up[4][B200MIX_MAX_AMBI_CHANNELS] = {};
        if(out->upmix) first_order_up(up);
        const float gain = tgt->slot_gain * tgt->reverb_boost;
        float earlymat[4][4], latemat[4][4];
        transform_from_vector(P.reflections_pan, earlymat);
        transform_from_vector(P.late_reverb_pan, latemat);
        float coeffs[4][B200MIX_MAX_AMBI_CHANNELS];
        line_coeffs(kEarlyA2B, earlymat, out->upmix != 0, up, coeffs);
        for(int j = 0;j < kLines;++j)
            if(int rc = b200mix_pan_gains(tgt->out_channels, tgt->out_scale, tgt->out_index, coeffs[j],
                P.reflections_gain*gain, gains + size_t(j)*tgt->out_channels, tgt->out_channels)) return rc;
        line_coeffs(kLateA2B, latemat, out->upmix != 0, up, coeffs);
        for(int j = 0;j < kLines;++j)
            if(int rc = b200mix_pan_gains(tgt->out_channels, tgt->out_scale, tgt->out_index, coeffs[j],
                P.late_reverb_gain*gain, gains + size_t(4 + j)*tgt->out_channels, tgt->out_channels)) return rc;
    }
    return B200MIX_OK;
}

} // extern "C"
