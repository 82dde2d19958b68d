// Host-side source parameter stage (no GPU): from a source's properties and the listener to the
// mixing parameters of its voice — resampler step, listener-relative direction, distance, spread
// and the direct / per-send gain triplets.  Restates CalcAttnVoiceParams (alc/alu.cpp:1712-2010)
// operation for operation for point sources; the results go through the panning helpers
// (b200mix_ambi_coeffs / b200mix_pan_gains / b200mix_hrtf_get_coeffs, b200mix_pairwise_azimuth)
// and b200mix_biquad_coeffs exactly as CalcPanningAndFilters does (alc/alu.cpp:1519-1656), giving
// bit-identical voice descriptors (tests/test_source_params.py pins them against live voices of
// the compiled reference).  This is the CPU form of the GPU parameter stage SURVEY §8(f) ranks next.
#include "../../include/b200mix.h"

#include <algorithm>
#include <cmath>
#include <limits>
#include <thread>
#include <vector>

#include "param_math.hpp"

namespace {

using namespace b200mix::pm;

// <cmath>: the reference's own libm calls (see param_math.hpp)
struct HostMath {
    static float sqrt(float x) { return std::sqrt(x); }
    static float pow(float a, float b) { return std::pow(a, b); }
    static float acos(float x) { return std::acos(x); }
    static float asin(float x) { return std::asin(x); }
    static float atan2(float y, float x) { return std::atan2(y, x); }
    static float sin(float x) { return std::sin(x); }
    static float cos(float x) { return std::cos(x); }
    static float copysign(float a, float b) { return std::copysign(a, b); }
    static long lrint(float x) { return std::lrintf(x); }
    static float infinity() { return std::numeric_limits<float>::infinity(); }
};

constexpr float kEps = std::numeric_limits<float>::epsilon();

float normalize(Vec &a) { return b200mix::pm::normalize<HostMath>(a); }

// The filter block of CalcPanningAndFilters (alc/alu.cpp:1619-1656) for every path of one voice
int design_filters(const b200mix_source_props &P, uint32_t device_rate, uint32_t num_sends, const float *gainHF,
    const float *gainLF, const uint32_t *voice, b200mix_voice_filter *filters)
{
    for(uint32_t path = 0;path <= num_sends;++path)
    {
        if(voice) filters[path].voice = *voice;
        design_filter<HostMath>(P, device_rate, path, gainHF[path], gainLF[path], filters[path]);
    }
    return B200MIX_OK;
}

} // namespace

extern "C" {

int b200mix_calc_listener_params(const b200mix_listener_props *props, b200mix_listener_params *out)
{
    if(!props || !out || props->struct_size != sizeof(*props) || props->distance_model > ExponentClamped)
        return B200MIX_ERR_INVALID;
    // CalcContextParams (alc/alu.cpp:508-555)
    *out = b200mix_listener_params{};
    out->struct_size = sizeof(*out);
    for(int i = 0;i < 3;++i) out->position[i] = props->position[i];
    Vec N{{props->orient_at[0], props->orient_at[1], props->orient_at[2], 0.0f}};
    normalize(N);
    Vec V{{props->orient_up[0], props->orient_up[1], props->orient_up[2], 0.0f}};
    normalize(V);
    Vec U{{N.v[1]*V.v[2] - N.v[2]*V.v[1], N.v[2]*V.v[0] - N.v[0]*V.v[2], N.v[0]*V.v[1] - N.v[1]*V.v[0], 0.0f}};
    normalize(U);
    const float rot[16] = {
        U.v[0], V.v[0], -N.v[0], 0.0f,
        U.v[1], V.v[1], -N.v[1], 0.0f,
        U.v[2], V.v[2], -N.v[2], 0.0f,
        0.0f,   0.0f,   0.0f,    1.0f};
    for(int i = 0;i < 16;++i) out->matrix[i] = rot[i];
    const Vec vel = mul(rot, Vec{{props->velocity[0], props->velocity[1], props->velocity[2], 0.0f}});
    for(int i = 0;i < 3;++i) out->velocity[i] = vel.v[i];
    out->gain = props->gain * props->gain_boost;
    out->meters_per_unit = props->meters_per_unit;
    out->air_absorption_gain_hf = props->air_absorption_gain_hf;
    out->doppler_factor = props->doppler_factor;
    out->speed_of_sound = props->speed_of_sound * props->doppler_velocity;
    out->source_distance_model = props->source_distance_model ? 1u : 0u;
    out->distance_model = props->distance_model;
    return B200MIX_OK;
}

int b200mix_pairwise_azimuth(const float pos[3], float out[3])
{
    if(!pos || !out) return B200MIX_ERR_INVALID;
    pairwise_azimuth<HostMath>(pos, out);
    return B200MIX_OK;
}

int b200mix_calc_source_params(const b200mix_source_props *props, const b200mix_listener_params *lis,
    uint32_t num_sends, uint32_t buffer_rate, uint32_t device_rate, b200mix_source_result *out)
{
    if(!props || !lis || !out || props->struct_size != sizeof(*props) || lis->struct_size != sizeof(*lis)
        || num_sends > B200MIX_MAX_SENDS || !device_rate || props->distance_model > ExponentClamped
        || lis->distance_model > ExponentClamped)
        return B200MIX_ERR_INVALID;
    calc_source_params<HostMath>(*props, *lis, num_sends, buffer_rate, device_rate, *out);
    return B200MIX_OK;
}

int b200mix_calc_voice(const b200mix_source_props *props, const b200mix_listener_params *listener,
    const b200mix_voice_env *env, uint32_t buffer_rate, b200mix_voice_params *voice, float dir[4],
    float *dry_gains, float *send_gains, b200mix_voice_filter *filters)
{
    if(!props || !listener || !env || !voice || !filters || env->struct_size != sizeof(*env)
        || env->num_sends > B200MIX_MAX_SENDS || env->render_mode > 2u)
        return B200MIX_ERR_INVALID;
    b200mix_source_result r{};
    if(int rc = b200mix_calc_source_params(props, listener, env->num_sends, buffer_rate, env->device_rate, &r))
        return rc;
    voice->step = r.step;
    if(env->render_mode == 2u) { if(!dir) return B200MIX_ERR_INVALID; }
    else if(!dry_gains || !env->dry.scale || !env->dry.index) return B200MIX_ERR_INVALID;
    float hrtf_gain = 0.0f; bool is_hrtf = false;
    if(!calc_panning<HostMath>(*props, r, *env, &hrtf_gain, &is_hrtf, dir, dry_gains, send_gains))
        return B200MIX_ERR_INVALID;
    if(is_hrtf) { voice->hrtf_gain = hrtf_gain; voice->flags |= B200MIX_VF_HRTF; }
    else voice->flags &= ~uint32_t(B200MIX_VF_HRTF);
    float gainHF[1 + B200MIX_MAX_SENDS], gainLF[1 + B200MIX_MAX_SENDS];
    gainHF[0] = r.dry_gain_hf; gainLF[0] = r.dry_gain_lf;
    for(uint32_t i = 0;i < B200MIX_MAX_SENDS;++i) { gainHF[1+i] = r.wet_gain_hf[i]; gainLF[1+i] = r.wet_gain_lf[i]; }
    return design_filters(*props, env->device_rate, env->num_sends, gainHF, gainLF, &voice->voice, filters);
}

int b200mix_calc_voice_channels(const b200mix_source_props *props, const b200mix_listener_params *listener,
    const b200mix_voice_env *env, uint32_t buffer_rate, const b200mix_channel_setup *setup,
    uint32_t *step, float *hrtf_gains, float *dirs, float *dry_gains, float *send_gains,
    b200mix_voice_filter *filters)
{
    if(!props || !listener || !env || !setup || !step || !filters || env->struct_size != sizeof(*env)
        || props->struct_size != sizeof(*props) || listener->struct_size != sizeof(*listener)
        || setup->struct_size != sizeof(*setup) || env->num_sends > B200MIX_MAX_SENDS || env->render_mode > 2u
        || !env->device_rate)
        return B200MIX_ERR_INVALID;
    const b200mix_source_props &P = *props;

    // channel positions (alc/alu.cpp:892-897,1471-1517; StereoMap :1525-1528 with StereoPan :1553-1562)
    enum Kind { L, R, C, Lfe };
    struct Chan { Kind kind; float pos[3]; };
    constexpr float sin30 = 0.5f, cos30 = 0.866025403785f;
    constexpr float sin45 = 1.41421356237309504880f*0.5f, cos45 = 1.41421356237309504880f*0.5f;
    constexpr float sin110 = 0.939692620786f, cos110 = -0.342020143326f;
    Chan chans[8];
    uint32_t nch = 0;
    auto add = [&](Kind k, float x, float y, float z) { chans[nch++] = Chan{k, {x, y, z}}; };
    switch(setup->layout)
    {
    case B200MIX_LAYOUT_STEREO:
        add(L, -std::sin(setup->stereo_pan[0]), 0.0f, -std::cos(setup->stereo_pan[0]));
        add(R, -std::sin(setup->stereo_pan[1]), 0.0f, -std::cos(setup->stereo_pan[1]));
        break;
    case B200MIX_LAYOUT_REAR: add(L, -sin30, 0.0f, cos30); add(R, sin30, 0.0f, cos30); break;
    case B200MIX_LAYOUT_QUAD:
        add(L, -sin45, 0.0f, -cos45); add(R, sin45, 0.0f, -cos45); add(L, -sin45, 0.0f, cos45); add(R, sin45, 0.0f, cos45);
        break;
    case B200MIX_LAYOUT_X51:
        add(L, -sin30, 0.0f, -cos30); add(R, sin30, 0.0f, -cos30); add(C, 0.0f, 0.0f, -1.0f); add(Lfe, 0.0f, 0.0f, 0.0f);
        add(L, -sin110, 0.0f, -cos110); add(R, sin110, 0.0f, -cos110);
        break;
    case B200MIX_LAYOUT_X61:
        add(L, -sin30, 0.0f, -cos30); add(R, sin30, 0.0f, -cos30); add(C, 0.0f, 0.0f, -1.0f); add(Lfe, 0.0f, 0.0f, 0.0f);
        add(C, 0.0f, 0.0f, 1.0f); add(L, -1.0f, 0.0f, 0.0f); add(R, 1.0f, 0.0f, 0.0f);
        break;
    case B200MIX_LAYOUT_X71:
        add(L, -sin30, 0.0f, -cos30); add(R, sin30, 0.0f, -cos30); add(C, 0.0f, 0.0f, -1.0f); add(Lfe, 0.0f, 0.0f, 0.0f);
        add(L, -sin30, 0.0f, cos30); add(R, sin30, 0.0f, cos30); add(L, -1.0f, 0.0f, 0.0f); add(R, 1.0f, 0.0f, 0.0f);
        break;
    default: return B200MIX_ERR_INVALID;
    }

    float dryBase, wetBase[B200MIX_MAX_SENDS] = {};
    float gainHF[1 + B200MIX_MAX_SENDS], gainLF[1 + B200MIX_MAX_SENDS];
    b200mix_source_result attn{};
    bool warp = false;
    if(setup->spatialized)
    {
        // AL_SOURCE_SPATIALIZE_SOFT on a multi-channel source: CalcAttnVoiceParams (:1712-2010)
        if(int rc = b200mix_calc_source_params(props, listener, env->num_sends, buffer_rate, env->device_rate, &attn))
            return rc;
        *step = attn.step;
        dryBase = attn.dry_gain;
        gainHF[0] = attn.dry_gain_hf; gainLF[0] = attn.dry_gain_lf;
        for(uint32_t i = 0;i < env->num_sends;++i)
        { wetBase[i] = attn.wet_gain[i]; gainHF[1+i] = attn.wet_gain_hf[i]; gainLF[1+i] = attn.wet_gain_lf[i]; }
        warp = attn.distance > kEps;
    }
    else
    {
        // CalcNonAttnVoiceParams (alc/alu.cpp:1658-1710)
        const float pitch = float(buffer_rate) / float(env->device_rate) * P.pitch;
        if(pitch > float(kMaxPitch)) *step = kMaxPitch << kFracBits;
        else *step = std::max(uint32_t(std::lrintf(pitch * kFracOne)), 1u);
        const float mingain = std::min(P.min_gain, P.max_gain);
        const float srcgain = std::clamp(P.gain, mingain, P.max_gain);
        dryBase = std::min(kGainMixMax, srcgain * P.direct.gain * listener->gain);
        gainHF[0] = P.direct.gain_hf; gainLF[0] = P.direct.gain_lf;
        for(uint32_t i = 0;i < env->num_sends;++i)
        {
            wetBase[i] = std::min(kGainMixMax, srcgain * P.sends[i].gain * listener->gain);
            gainHF[1+i] = P.sends[i].gain_hf; gainLF[1+i] = P.sends[i].gain_lf;
        }
    }

    // GetPanGainSelector (:1078-1116)
    const float lgain = std::min(1.0f - setup->panning, 1.0f), rgain = std::min(1.0f + setup->panning, 1.0f);
    const float cgain = std::min(lgain, rgain);

    // the no-distance paths of CalcHrtfPanning (:1268-1310) and CalcNormalPanning (:1420-1466); a
    // multi-channel source has no spread of its own here (spreadmult = 0)
    const uint32_t nd = env->dry.channels;
    for(uint32_t c = 0;c < nch;++c)
    {
        const Chan &ch = chans[c];
        const float pangain = ch.kind == L ? lgain : ch.kind == R ? rgain : cgain;
        float *dg = dry_gains ? dry_gains + size_t(c)*nd : nullptr;
        float *sg = (send_gains && env->wet_stride) ? send_gains + size_t(c)*env->num_sends*env->wet_stride : nullptr;
        if(dg) for(uint32_t k = 0;k < nd;++k) dg[k] = 0.0f;
        if(sg) for(uint32_t k = 0;k < env->num_sends*env->wet_stride;++k) sg[k] = 0.0f;
        if(hrtf_gains) hrtf_gains[c] = 0.0f;
        if(dirs) { dirs[c*4+0] = 0.0f; dirs[c*4+1] = 0.0f; dirs[c*4+2] = std::numeric_limits<float>::infinity(); dirs[c*4+3] = 0.0f; }
        float coeffs[B200MIX_MAX_AMBI_CHANNELS];
        if(ch.kind == Lfe)
        {
            // LFE plays only where the Dry mix IS the output mix and has an LFE channel (:1441-1450)
            if(env->render_mode != 2u && dg && setup->lfe_dry_index < nd) dg[setup->lfe_dry_index] = dryBase*pangain;
            continue;
        }
        float cpos[3] = {ch.pos[0], ch.pos[1], ch.pos[2]};
        if(warp)
        {
            // a spatialized source pulls its channels toward its direction as the spread shrinks
            // (:1234-1250,1380-1396)
            const float a = 1.0f - (0.318309886183790671538f*0.5f)*attn.spread;
            for(int k = 0;k < 3;++k) cpos[k] = lerpf(ch.pos[k], attn.pos[k], a);
            if(const float len = std::sqrt(cpos[0]*cpos[0] + cpos[1]*cpos[1] + cpos[2]*cpos[2]); len < 1.0f)
            { cpos[0] /= len; cpos[1] /= len; cpos[2] /= len; }
        }
        if(env->render_mode == 2u)
        {
            if(!hrtf_gains || !dirs) return B200MIX_ERR_INVALID;
            dirs[c*4+0] = warp ? std::asin(std::clamp(cpos[1], -1.0f, 1.0f)) : std::asin(cpos[1]);
            dirs[c*4+1] = std::atan2(cpos[0], -cpos[2]);
            if(warp) dirs[c*4+2] = attn.distance;
            hrtf_gains[c] = dryBase * pangain;
            ambi_coeffs<HostMath>(cpos, 0.0f, coeffs);
        }
        else
        {
            if(!dg || !env->dry.scale || !env->dry.index) return B200MIX_ERR_INVALID;
            float pos[3] = {cpos[0], cpos[1], cpos[2]};
            if(env->render_mode == 1u && pos[2] < 0.0f)
            {
                // ScaleAzimuthFront3 (:642-673)
                const float len2d = std::sqrt(pos[0]*pos[0] + pos[2]*pos[2]);
                float z = -pos[2] / len2d;
                if(z > 0.866025403785f)
                {
                    float x = pos[0] / len2d;
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
            b200mix_ambi_coeffs(pos, 0.0f, coeffs);
            if(int rc = b200mix_pan_gains(nd, env->dry.scale, env->dry.index, coeffs, dryBase * pangain, dg, nd))
                return rc;
        }
        for(uint32_t i = 0;i < env->num_sends && sg;++i)
        {
            const b200mix_mix_map &w = env->wet[i];
            if(!P.sends[i].active || !w.channels) continue;
            if(w.channels > env->wet_stride || !w.scale || !w.index) return B200MIX_ERR_INVALID;
            if(int rc = b200mix_pan_gains(w.channels, w.scale, w.index, coeffs, wetBase[i] * pangain,
                sg + size_t(i)*env->wet_stride, w.channels)) return rc;
        }
    }

    // every channel shares channel 0's filters
    if(int rc = design_filters(P, env->device_rate, env->num_sends, gainHF, gainLF, nullptr, filters)) return rc;
    return int(nch);
}

namespace {

// Rotation of real spherical harmonics above first order: the Ivanic / Ruedenberg band recursion as
// the reference evaluates it (AmbiRotator and RotatorCoeffs, alc/alu.cpp:709-889).  `R` holds the
// first-order block on entry (ACN rows / columns 1..3 = m -1, 0, +1); band l (rows / columns
// l*l .. l*l + 2l) is built from band l-1 and the first-order block.  Products and sums are taken in
// the reference's order, so the matrix is the reference's bit for bit.
struct BandRotator {
    float (*R)[B200MIX_MAX_AMBI_CHANNELS];
    int l;                       // band being built
    unsigned prev;               // ACN index of band l-1's first row / column: (l-1)^2

    // one term of the recursion: first-order column i (-1, 0, +1) against band l-1's column a, for row n of band l
    float term(int i, int a, int n) const
    {
        const unsigned fc = unsigned(i + 2);
        const unsigned col = prev + unsigned((l - 1) + a);
        const unsigned lo = prev, hi = prev + unsigned(2*(l - 1));
        if(n == -l) return R[3][fc]*R[lo][col] + R[1][fc]*R[hi][col];
        if(n == l) return R[3][fc]*R[hi][col] - R[1][fc]*R[lo][col];
        return R[2][fc]*R[prev + unsigned((l - 1) + n)][col];
    }
    float u_term(int m, int n) const { return term(0, m, n); }
    float v_term(int m, int n) const
    {
        const float sqrt2 = 1.41421356237309504880f;
        if(m > 0)
        {
            const float a = term(1, m - 1, n), b = term(-1, 1 - m, n);
            return (m == 1) ? a*sqrt2 : (a - b);
        }
        const float a = term(1, m + 1, n), b = term(-1, -m - 1, n);
        return (m == -1) ? b*sqrt2 : (a + b);
    }
    float w_term(int m, int n) const
    {
        if(m > 0) return term(1, m + 1, n) + term(-1, -m - 1, n);
        return term(1, m - 1, n) - term(-1, 1 - m, n);
    }
};

void rotate_higher_orders(float (*R)[B200MIX_MAX_AMBI_CHANNELS], int order)
{
    for(int l = 2;l <= order;++l)
    {
        const BandRotator B{R, l, unsigned((l - 1)*(l - 1))};
        const unsigned base = unsigned(l*l);
        for(int n = -l;n <= l;++n)
        {
            // Table I of the paper (alc/alu.cpp:727-772): evaluated in double, stored as float
            const double denom = (n == l || n == -l) ? double((2*l)*(2*l - 1)) : double(l*l - n*n);
            for(int m = -l;m <= l;++m)
            {
                float u, v, w;
                if(m == 0)
                {
                    u = float(std::sqrt(l*l / denom));
                    v = float(std::sqrt((l - 1)*l / denom) * -1.0);
                    w = 0.0f;
                }
                else
                {
                    const int am = m < 0 ? -m : m;
                    u = float(std::sqrt((l*l - m*m) / denom));
                    v = float(std::sqrt((l + am - 1)*(l + am) / denom) * 0.5);
                    w = float(std::sqrt((l - am - 1)*(l - am) / denom) * -0.5);
                }
                float r = 0.0f;
                if(u != 0.0f) r += u * B.u_term(m, n);
                if(v != 0.0f) r += v * B.v_term(m, n);
                if(w != 0.0f) r += w * B.w_term(m, n);
                R[base + unsigned(n + l)][base + unsigned(m + l)] = r;
            }
        }
    }
}

// AmbiScale::FromFuMa / FromSN3D / FromN3D (core/ambidefs.h:49-124; FuMa is defined up to third order)
const float kAmbiScaleTab[3][B200MIX_MAX_AMBI_CHANNELS] = {
    {1.414213562f, 1.732050808f, 1.732050808f, 1.732050808f, 1.936491673f, 1.936491673f, 2.236067978f, 1.936491673f,
     1.936491673f, 2.091650066f, 1.972026594f, 2.231093404f, 2.645751311f, 2.231093404f, 1.972026594f, 2.091650066f,
     0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f},
    {1.000000000f, 1.732050808f, 1.732050808f, 1.732050808f, 2.236067978f, 2.236067978f, 2.236067978f, 2.236067978f,
     2.236067978f, 2.645751311f, 2.645751311f, 2.645751311f, 2.645751311f, 2.645751311f, 2.645751311f, 2.645751311f,
     3.0f, 3.0f, 3.0f, 3.0f, 3.0f, 3.0f, 3.0f, 3.0f, 3.0f},
    {1.0f, 1.0f, 1.0f, 1.0f, 1.0f, 1.0f, 1.0f, 1.0f, 1.0f, 1.0f, 1.0f, 1.0f, 1.0f, 1.0f, 1.0f, 1.0f,
     1.0f, 1.0f, 1.0f, 1.0f, 1.0f, 1.0f, 1.0f, 1.0f, 1.0f}};
// AmbiIndex::FromFuMa / FromFuMa2D / FromACN2D (core/ambidefs.h:143-190): buffer channel -> ACN
const unsigned char kFromFuMa[16] = {0, 3, 1, 2, 6, 7, 5, 8, 4, 12, 13, 11, 14, 10, 15, 9};
const unsigned char kFromFuMa2D[7] = {0, 3, 1, 8, 4, 15, 9};
const unsigned char kFromACN2D[9] = {0, 1, 3, 4, 8, 9, 15, 16, 24};

} // namespace

int b200mix_calc_voice_bformat(const b200mix_source_props *props, const b200mix_listener_params *listener,
    const b200mix_voice_env *env, uint32_t buffer_rate, const b200mix_bformat_setup *setup, uint32_t *step,
    float *dry_gains, float *send_gains, b200mix_voice_filter *filters)
{
    if(!props || !listener || !env || !setup || !step || !filters || !dry_gains
        || env->struct_size != sizeof(*env) || props->struct_size != sizeof(*props)
        || listener->struct_size != sizeof(*listener) || setup->struct_size != sizeof(*setup)
        || env->num_sends > B200MIX_MAX_SENDS || env->render_mode > 2u || !env->device_rate
        || setup->layout > 1u || setup->scaling > 2u || !env->dry.scale || !env->dry.index)
        return B200MIX_ERR_INVALID;
    const uint32_t srcOrder = setup->source_ambi_order ? setup->source_ambi_order : 1u;
    const uint32_t devOrder = setup->device_ambi_order;
    if(devOrder < 1u || devOrder > 4u || srcOrder > 4u || (setup->layout == 0u && srcOrder > 3u))
        return B200MIX_ERR_INVALID;
    // a device of higher order than the source (or a 2D bed on a 3D mix from second order on)
    // up-samples and band-splits the voice (alc/alu.cpp:1001-1036, core/voice.cpp:1082-1089,1362-1385)
    if(devOrder > srcOrder || (devOrder >= 2u && !setup->device_2d_mixing && setup->is_2d))
        return B200MIX_ERR_UNSUPPORTED;
    const b200mix_source_props &P = *props;

    // CalcNonAttnVoiceParams (alc/alu.cpp:1658-1710)
    const float pitch = float(buffer_rate) / float(env->device_rate) * P.pitch;
    if(pitch > float(kMaxPitch)) *step = kMaxPitch << kFracBits;
    else *step = std::max(uint32_t(std::lrintf(pitch * kFracOne)), 1u);
    const float mingain = std::min(P.min_gain, P.max_gain);
    const float srcgain = std::clamp(P.gain, mingain, P.max_gain);
    const float dryBase = std::min(kGainMixMax, srcgain * P.direct.gain * listener->gain);
    float wetBase[B200MIX_MAX_SENDS];
    for(uint32_t i = 0;i < env->num_sends;++i)
        wetBase[i] = std::min(kGainMixMax, srcgain * P.sends[i].gain * listener->gain);

    // CalcAmbisonicPanning with no distance: coverage 1 (:946-949)
    const float coverage = 1.0f;
    const float *scales = kAmbiScaleTab[setup->scaling];
    // the mixed channels in ACN terms: the buffer's leading channels up to the device's order
    // (Voice::prepare, core/voice.cpp:1246-1248) — (order+1)^2 of them, 2*order+1 for a 2D bed
    const uint32_t mixOrder = std::min(srcOrder, devOrder);
    const unsigned nch = setup->is_2d ? 2u*mixOrder + 1u : (mixOrder + 1u)*(mixOrder + 1u);
    unsigned index_map[B200MIX_MAX_AMBI_CHANNELS];
    for(unsigned c = 0;c < nch;++c)
        index_map[c] = setup->is_2d ? (setup->layout ? kFromACN2D[c] : kFromFuMa2D[c])
            : (setup->layout ? c : kFromFuMa[c]);

    // the panned W term (:951-957), then scaled by (1 - coverage) (:1049-1050)
    float pan[B200MIX_MAX_AMBI_CHANNELS];
    {
        const float front[3] = {0.0f, 0.0f, -1.0f};
        float pos[3] = {front[0], front[1], front[2]};
        if(env->render_mode == 1u) b200mix_pairwise_azimuth(front, pos);
        b200mix_ambi_coeffs(pos, 0.0f, pan);
        const float scale = (1.0f-coverage)*scales[0];
        for(unsigned k = 0;k < B200MIX_MAX_AMBI_CHANNELS;++k) pan[k] = pan[k] * scale;
    }

    // orientation -> first-order rotation (:976-999)
    Vec N{{P.orient_at[0], P.orient_at[1], P.orient_at[2], 0.0f}};
    normalize(N);
    Vec V{{P.orient_up[0], P.orient_up[1], P.orient_up[2], 0.0f}};
    normalize(V);
    if(!P.head_relative) { N = mul(listener->matrix, N); V = mul(listener->matrix, V); }
    Vec U{{N.v[1]*V.v[2] - N.v[2]*V.v[1], N.v[2]*V.v[0] - N.v[0]*V.v[2], N.v[0]*V.v[1] - N.v[1]*V.v[0], 0.0f}};
    normalize(U);
    // the first-order block by hand, the bands above it up to the device's order by recursion (:992-999);
    // bands the device does not mix stay zero
    float shrot[B200MIX_MAX_AMBI_CHANNELS][B200MIX_MAX_AMBI_CHANNELS] = {};
    shrot[0][0] = 1.0f;
    shrot[1][1] =  U.v[0]; shrot[1][2] = -U.v[1]; shrot[1][3] =  U.v[2];
    shrot[2][1] = -V.v[0]; shrot[2][2] =  V.v[1]; shrot[2][3] = -V.v[2];
    shrot[3][1] = -N.v[0]; shrot[3][2] =  N.v[1]; shrot[3][3] = -N.v[2];
    rotate_higher_orders(shrot, int(devOrder));

    const uint32_t nd = env->dry.channels;
    float coeffs[B200MIX_MAX_AMBI_CHANNELS];
    for(unsigned k = 0;k < B200MIX_MAX_AMBI_CHANNELS;++k) coeffs[k] = pan[k];
    for(unsigned c = 0;c < nch;++c)
    {
        const unsigned acn = index_map[c];
        const float scale = scales[acn] * coverage;
        for(unsigned k = 0;k < B200MIX_MAX_AMBI_CHANNELS;++k)
            coeffs[k] = shrot[acn][k]*scale + coeffs[k];
        if(int rc = b200mix_pan_gains(nd, env->dry.scale, env->dry.index, coeffs, dryBase,
            dry_gains + size_t(c)*nd, nd)) return rc;
        if(send_gains && env->wet_stride)
        {
            float *sg = send_gains + size_t(c)*env->num_sends*env->wet_stride;
            for(uint32_t k = 0;k < env->num_sends*env->wet_stride;++k) sg[k] = 0.0f;
            for(uint32_t i = 0;i < env->num_sends;++i)
            {
                const b200mix_mix_map &w = env->wet[i];
                if(!P.sends[i].active || !w.channels) continue;
                if(w.channels > env->wet_stride || !w.scale || !w.index) return B200MIX_ERR_INVALID;
                if(int rc = b200mix_pan_gains(w.channels, w.scale, w.index, coeffs, wetBase[i],
                    sg + size_t(i)*env->wet_stride, w.channels)) return rc;
            }
        }
        for(unsigned k = 0;k < B200MIX_MAX_AMBI_CHANNELS;++k) coeffs[k] = 0.0f;      // :1074
    }

    float gainHF[1 + B200MIX_MAX_SENDS], gainLF[1 + B200MIX_MAX_SENDS];
    gainHF[0] = P.direct.gain_hf; gainLF[0] = P.direct.gain_lf;
    for(uint32_t i = 0;i < B200MIX_MAX_SENDS;++i) { gainHF[1+i] = P.sends[i].gain_hf; gainLF[1+i] = P.sends[i].gain_lf; }
    if(int rc = design_filters(P, env->device_rate, env->num_sends, gainHF, gainLF, nullptr, filters)) return rc;
    return int(nch);
}

int b200mix_calc_voices(uint32_t n, const b200mix_source_props *props, const b200mix_listener_params *listener,
    const b200mix_voice_env *env, const uint32_t *buffer_rates, b200mix_voice_params *voices, float *dirs,
    float *dry_gains, float *send_gains, b200mix_voice_filter *filters, uint32_t threads)
{
    if(!n) return B200MIX_OK;
    if(!props || !listener || !env || !buffer_rates || !voices || !filters || env->struct_size != sizeof(*env))
        return B200MIX_ERR_INVALID;
    const size_t nd = env->dry.channels, ns = size_t(env->num_sends)*env->wet_stride;
    const size_t nf = 1u + env->num_sends;
    auto run = [&](uint32_t first, uint32_t last, int *status)
    {
        for(uint32_t i = first;i < last;++i)
        {
            const int rc = b200mix_calc_voice(&props[i], listener, env, buffer_rates[i], &voices[i],
                dirs ? dirs + size_t(i)*4 : nullptr, dry_gains ? dry_gains + size_t(i)*nd : nullptr,
                send_gains ? send_gains + size_t(i)*ns : nullptr, filters + size_t(i)*nf);
            if(rc && !*status) *status = rc;
        }
    };
    // sources are independent: split them evenly over the threads (the caller's included)
    const uint32_t nt = std::max(1u, std::min(threads ? threads : 1u, (n + 63u)/64u));
    std::vector<int> status(nt, 0);
    std::vector<std::thread> pool;
    const uint32_t per = (n + nt - 1u)/nt;
    for(uint32_t t = 1;t < nt;++t)
        pool.emplace_back(run, std::min(n, t*per), std::min(n, (t + 1u)*per), &status[t]);
    run(0u, std::min(n, per), &status[0]);
    for(auto &th : pool) th.join();
    for(int st : status) if(st) return st;
    return B200MIX_OK;
}

} // extern "C"
