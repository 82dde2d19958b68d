// param_math.hpp — the arithmetic of the parameter ("ALU") stage for point sources, ONE source
// text for the host helpers (csrc/source_params.cpp, csrc/panning.cpp, b200mix_biquad_coeffs) and
// for the GPU parameter kernel (csrc/param_kernels.cu): CalcAttnVoiceParams
// (alc/alu.cpp:1712-2010), CalcPanningAndFilters' point-source steps (:1196-1226,1318-1361,
// 1619-1656), CalcAmbiCoeffs (core/mixer.cpp:16-91), ComputePanGains (:93-102),
// BiquadFilter::SetParams (core/filters/biquad.cpp:48-129) and BsincPrepare (alc/alu.cpp:140-165),
// operation for operation.
//
// Every function is templated on a math policy M that supplies the libm calls.  HostMath is
// <cmath> (the reference's own calls, so the host helpers stay bit-identical to the reference —
// tests/test_source_params.py, test_pan_params.py, test_filter_params.py).  DeviceMath
// (param_kernels.cu) evaluates the same calls in double precision and rounds once, and that
// translation unit is compiled with -fmad=false (no FMA contraction, IEEE division and square
// root), so the device follows the same float operation sequence: results equal the host's
// except where a libm call of the host is not correctly rounded (<= 1 ulp, rare).
#pragma once
#include <cstdint>

#include "../../include/b200mix.h"

#if defined(__CUDACC__)
#define PM_HD __host__ __device__ __forceinline__
#else
#define PM_HD inline
#endif

namespace b200mix { namespace pm {

constexpr float kGainMixMax = 1000.0f;            // alc/alu.h:18
constexpr float kSpeedOfSound = 343.3f;           // core/context.h:32
constexpr float kReverbDecayGain = 0.001f;        // core/effects/base.h:22
constexpr unsigned kMaxPitch = 10u, kFracBits = 16u;
constexpr float kFracOne = 65536.0f;
constexpr float kPi = 3.14159265358979323846f;
constexpr float kEpsF = 1.1920928955078125e-07f;  // std::numeric_limits<float>::epsilon()

struct Vec { float v[4]; };

PM_HD float fminf_(float a, float b) { return b < a ? b : a; }          // std::min
PM_HD float fmaxf_(float a, float b) { return a < b ? b : a; }          // std::max
PM_HD float clampf_(float v, float lo, float hi) { return v < lo ? lo : (hi < v ? hi : v); }   // std::clamp
PM_HD float lerpf(float a, float b, float mu) { return a + (b-a)*mu; }
PM_HD float dot(const Vec &a, const Vec &b) { return a.v[0]*b.v[0] + a.v[1]*b.v[1] + a.v[2]*b.v[2]; }

// al::Vector::normalize (common/vecmat.h:51-65)
template<class M> PM_HD float normalize(Vec &a)
{
    const float length_sqr = a.v[0]*a.v[0] + a.v[1]*a.v[1] + a.v[2]*a.v[2];
    if(length_sqr > kEpsF)
    {
        const float length = M::sqrt(length_sqr);
        const float inv_length = 1.0f / length;
        a.v[0] *= inv_length; a.v[1] *= inv_length; a.v[2] *= inv_length;
        return length;
    }
    a.v[0] = a.v[1] = a.v[2] = 0.0f;
    return 0.0f;
}
// operator*(Matrix, Vector) (common/vecmat.h:113-120), m row-major
PM_HD Vec mul(const float *m, const Vec &x)
{
    Vec r;
    for(int c = 0;c < 4;++c)
        r.v[c] = x.v[0]*m[0*4+c] + x.v[1]*m[1*4+c] + x.v[2]*m[2*4+c] + x.v[3]*m[3*4+c];
    return r;
}

enum Model { Disable, Inverse, InverseClamped, Linear, LinearClamped, Exponent, ExponentClamped };

// BiquadFilter::SetParams behind setParamsFromSlope (core/filters/biquad.h:61-62,92-97;
// biquad.cpp:48-129).  type: 0 HighShelf, 1 LowShelf, 2 Peaking, 3 LowPass, 4 HighPass, 5 BandPass.
// BiquadFilter::SetParams (core/filters/biquad.cpp:48-129) with an explicit rcpQ.
template<class M> PM_HD void biquad_set_params(uint32_t type, float f0norm, float gain, float rcpQ, float coeffs[5]);
// rcpQFromBandwidth (core/filters/biquad.h:70-76)
template<class M> PM_HD float biquad_rcpq_from_bandwidth(float f0norm, float bandwidth)
{
    const float w0 = 3.14159265358979323846f*2.0f * f0norm;
    return 2.0f*M::sinh(M::log(2.0f)/2.0f*bandwidth*w0/M::sin(w0));
}
// setParamsFromBandwidth (core/filters/biquad.h:110-112)
template<class M> PM_HD void biquad_coeffs_bandwidth(uint32_t type, float f0norm, float gain, float bandwidth, float coeffs[5])
{ biquad_set_params<M>(type, f0norm, gain, biquad_rcpq_from_bandwidth<M>(f0norm, bandwidth), coeffs); }

template<class M> PM_HD void biquad_coeffs(uint32_t type, float f0norm, float gain, float slope, float coeffs[5])
{
    gain = fmaxf_(gain, 0.001f);
    const float rcpQ = M::sqrt((gain + 1.0f/gain)*(1.0f/slope - 1.0f) + 2.0f);
    biquad_set_params<M>(type, f0norm, gain, rcpQ, coeffs);
}

template<class M> PM_HD void biquad_set_params(uint32_t type, float f0norm, float gain, float rcpQ, float coeffs[5])
{
    gain = fmaxf_(gain, 0.00001f);
    const float w0 = 3.14159265358979323846f*2.0f * fminf_(f0norm, 0.49f);
    const float sin_w0 = M::sin(w0), cos_w0 = M::cos(w0);
    const float alpha = sin_w0/2.0f * rcpQ;
    float a[3] = {1.0f, 0.0f, 0.0f}, b[3] = {1.0f, 0.0f, 0.0f};
    const float gp = gain + 1.0f, gm = gain - 1.0f;
    if(type <= 1u)
    {
        // shelves: the low shelf is the high shelf with cos(w0) negated
        const float sg = 2.0f * M::sqrt(gain) * alpha;
        const float sgn = type == 0u ? 1.0f : -1.0f;
        const float cw = sgn*cos_w0;
        b[0] =            gain*(gp + gm*cw + sg);
        b[1] = sgn*-2.0f*gain*(gm + gp*cw);
        b[2] =            gain*(gp + gm*cw - sg);
        a[0] =                  gp - gm*cw + sg;
        a[1] = sgn*2.0f*       (gm - gp*cw);
        a[2] =                  gp - gm*cw - sg;
    }
    else if(type == 2u)
    {
        b[0] = 1.0f + alpha*gain; b[1] = -2.0f*cos_w0; b[2] = 1.0f - alpha*gain;
        a[0] = 1.0f + alpha/gain; a[1] = -2.0f*cos_w0; a[2] = 1.0f - alpha/gain;
    }
    else
    {
        a[0] = 1.0f + alpha; a[1] = -2.0f*cos_w0; a[2] = 1.0f - alpha;
        if(type == 3u) { b[0] = (1.0f - cos_w0)/2.0f; b[1] = 1.0f - cos_w0; b[2] = b[0]; }
        else if(type == 4u) { b[0] = (1.0f + cos_w0)/2.0f; b[1] = -(1.0f + cos_w0); b[2] = b[0]; }
        else { b[0] = alpha; b[1] = 0.0f; b[2] = -alpha; }
    }
    coeffs[0] = b[0]/a[0]; coeffs[1] = b[1]/a[0]; coeffs[2] = b[2]/a[0];
    coeffs[3] = a[1]/a[0]; coeffs[4] = a[2]/a[0];
}

// real N3D spherical harmonics up to fourth order, ACN order, ambisonic axes (x front, y left, z up)
PM_HD void sh_n3d(const float y, const float z, const float x, float c[B200MIX_MAX_AMBI_CHANNELS])
{
    const float xx = x*x, yy = y*y, zz = z*z;
    const float xy = x*y, yz = y*z, xz = x*z;
    const float x4 = xx*xx, y4 = yy*yy, x2y2 = xx*yy, z4 = zz*zz;
    constexpr float kSqrt3 = 1.7320508075688772935f;       // std::numbers::sqrt3_v<float>
    constexpr float kSqrt15 = 3.872983346e+00f, kSqrt5h = 1.118033989e+00f, kSqrt15h = 1.936491673e+00f;
    constexpr float k3a = 2.091650066e+00f, k3b = 1.024695076e+01f, k3c = 1.620185175e+00f;
    constexpr float k3d = 1.322875656e+00f, k3e = 5.123475383e+00f;
    constexpr float k4a = 8.874119675e+00f, k4b = 6.274950199e+00f, k4c = 3.354101966e+00f;
    constexpr float k4d = 2.371708245e+00f, k4e = 3.750000000e-01f, k4f = 1.677050983e+00f;
    constexpr float k4g = 2.218529919e+00f;

    c[0] = 1.0f;
    c[1] = kSqrt3 * y;
    c[2] = kSqrt3 * z;
    c[3] = kSqrt3 * x;

    c[4] = kSqrt15 * xy;
    c[5] = kSqrt15 * yz;
    c[6] = kSqrt5h * (3.0f*zz - 1.0f);
    c[7] = kSqrt15 * xz;
    c[8] = kSqrt15h * (xx - yy);

    c[9] = k3a * (y*(3.0f*xx - yy));
    c[10] = k3b * (z*xy);
    c[11] = k3c * (y*(5.0f*zz - 1.0f));
    c[12] = k3d * (z*(5.0f*zz - 3.0f));
    c[13] = k3c * (x*(5.0f*zz - 1.0f));
    c[14] = k3e * (z*(xx - yy));
    c[15] = k3a * (x*(xx - 3.0f*yy));

    c[16] = k4a * (xy*(xx - yy));
    c[17] = k4b * ((3.0f*xx - yy) * yz);
    c[18] = k4c * (xy * (7.0f*zz - 1.0f));
    c[19] = k4d * (yz * (7.0f*zz - 3.0f));
    c[20] = k4e * (35.0f*z4 - 30.0f*zz + 3.0f);
    c[21] = k4d * (xz * (7.0f*zz - 3.0f));
    c[22] = k4f * ((xx - yy) * (7.0f*zz - 1.0f));
    c[23] = k4b * ((xx - 3.0f*yy) * xz);
    c[24] = k4g * (x4 - 6.0f*x2y2 + y4);
}

// CalcDirectionCoeffs(dir, spread) (core/mixer.h:68-73 -> CalcAmbiCoeffs, core/mixer.cpp:16-91)
template<class M> PM_HD void ambi_coeffs(const float dir[3], float spread, float coeffs[B200MIX_MAX_AMBI_CHANNELS])
{
    // OpenAL to ambisonic axes: Y = -x, Z = y, X = -z (core/mixer.h:68-73)
    sh_n3d(-dir[0], dir[1], -dir[2], coeffs);
    if(spread > 0.0f)
    {
        // a spherical cap subtending `spread`, loudness-compensated zonal weights per order and
        // up to +3 dB for a full spread (core/mixer.cpp:21-88)
        const float ca = M::cos(spread * 0.5f);
        const float scale = M::sqrt(1.0f + 0.318309886183790671538f*0.5f*spread);   // inv_pi_v<float>
        const float caca = ca*ca;
        const float zh[5] = {
            scale,
            scale * 0.5f * (ca+1.0f),
            scale * 0.5f * ((ca+1.0f)*ca),
            scale * 0.125f * ((ca+1.0f)*(5.0f*caca - 1.0f)),
            scale * 0.125f * ((ca+1.0f)*(7.0f*caca - 3.0f)*ca)};
        for(unsigned i = 0;i < B200MIX_MAX_AMBI_CHANNELS;++i)
        {
            const unsigned order = i < 1u ? 0u : (i < 4u ? 1u : (i < 9u ? 2u : (i < 16u ? 3u : 4u)));
            coeffs[i] *= zh[order];
        }
    }
}

// ComputePanGains (core/mixer.cpp:93-102)
PM_HD bool pan_gains(uint32_t channels, const float *scale, const uint32_t *index,
    const float coeffs[B200MIX_MAX_AMBI_CHANNELS], float ingain, float *gains, uint32_t gains_len)
{
    for(uint32_t c = 0;c < channels;++c)
    {
        if(index[c] >= B200MIX_MAX_AMBI_CHANNELS) return false;
        gains[c] = scale[c] * coeffs[index[c]] * ingain;
    }
    for(uint32_t c = channels;c < gains_len;++c) gains[c] = 0.0f;
    return true;
}

// ScaleAzimuthFront3_2 (alc/alu.cpp:675-708): stretch front azimuths by 3/2 for the pairwise stereo mix
template<class M> PM_HD void pairwise_azimuth(const float pos[3], float out[3])
{
    float p[3] = {pos[0], pos[1], pos[2]};
    if(p[2] < 0.0f)
    {
        const float len2d = M::sqrt(p[0]*p[0] + p[2]*p[2]);
        float z = -p[2] / len2d;
        if(z > 0.5f)
        {
            float x = p[0] / len2d;
            x = M::copysign(M::sqrt((1.0f - z) * 0.5f), x);
            z = M::sqrt((1.0f + z) * 0.5f);
            x = x*3.0f - x*x*x*4.0f;
            z = z*z*z*4.0f - z*3.0f;
            p[0] = x * len2d;
            p[2] = -z * len2d;
        }
        else
        {
            p[0] = M::copysign(len2d, p[0]);
            p[2] = 0.0f;
        }
    }
    out[0] = p[0]; out[1] = p[1]; out[2] = p[2];
}

// CalcAttnVoiceParams (alc/alu.cpp:1712-2010) for a point source.
template<class M> PM_HD void calc_source_params(const b200mix_source_props &P, const b200mix_listener_params &lis,
    uint32_t num_sends, uint32_t buffer_rate, uint32_t device_rate, b200mix_source_result &out)
{
    float roomrolloff[B200MIX_MAX_SENDS] = {};
    for(uint32_t i = 0;i < num_sends;++i)
        if(P.sends[i].active) roomrolloff[i] = P.room_rolloff_factor + P.sends[i].slot_room_rolloff;

    // listener space (:1743-1760)
    Vec position{{P.position[0], P.position[1], P.position[2], 1.0f}};
    Vec velocity{{P.velocity[0], P.velocity[1], P.velocity[2], 0.0f}};
    Vec direction{{P.direction[0], P.direction[1], P.direction[2], 0.0f}};
    const Vec lvelocity{{lis.velocity[0], lis.velocity[1], lis.velocity[2], 0.0f}};
    if(!P.head_relative)
    {
        const Vec rel{{position.v[0] - lis.position[0], position.v[1] - lis.position[1],
            position.v[2] - lis.position[2], position.v[3] - 1.0f}};
        position = mul(lis.matrix, rel);
        velocity = mul(lis.matrix, velocity);
        direction = mul(lis.matrix, direction);
    }
    else
        for(int k = 0;k < 4;++k) velocity.v[k] += lvelocity.v[k];

    Vec tosource{{position.v[0], position.v[1], position.v[2], 0.0f}};
    const float distance = normalize<M>(tosource);
    const bool directional = normalize<M>(direction) > 0.0f;

    // distance attenuation (:1762-1852)
    const uint32_t model = lis.source_distance_model ? P.distance_model : lis.distance_model;
    float attenDistance = distance;
    if(model == InverseClamped || model == LinearClamped || model == ExponentClamped)
    {
        if(!(P.ref_distance <= P.max_distance)) attenDistance = P.ref_distance;
        else attenDistance = clampf_(distance, P.ref_distance, P.max_distance);
    }

    float dryBase = P.gain, dryHF = 1.0f, dryLF = 1.0f;
    float wetBase[B200MIX_MAX_SENDS], wetHF[B200MIX_MAX_SENDS], wetLF[B200MIX_MAX_SENDS];
    for(uint32_t i = 0;i < B200MIX_MAX_SENDS;++i) { wetBase[i] = P.gain; wetHF[i] = 1.0f; wetLF[i] = 1.0f; }

    float dryAttnBase = 1.0f;
    switch(model)
    {
    case Inverse: case InverseClamped:
        if(P.ref_distance > 0.0f)
        {
            {
                const float dist = lerpf(P.ref_distance, attenDistance, P.rolloff_factor);
                if(dist > 0.0f)
                {
                    dryAttnBase = P.ref_distance / dist;
                    dryBase *= dryAttnBase;
                }
            }
            for(uint32_t i = 0;i < num_sends;++i)
            {
                const float dist = lerpf(P.ref_distance, attenDistance, roomrolloff[i]);
                if(dist > 0.0f) wetBase[i] = wetBase[i] * (P.ref_distance / dist);
            }
        }
        break;
    case Linear: case LinearClamped:
        if(P.max_distance != P.ref_distance)
        {
            const float scale = (attenDistance-P.ref_distance) / (P.max_distance-P.ref_distance);
            dryAttnBase = fmaxf_(1.0f - scale*P.rolloff_factor, 0.0f);
            dryBase *= dryAttnBase;
            for(uint32_t i = 0;i < num_sends;++i)
                wetBase[i] = wetBase[i] * fmaxf_(1.0f - scale*roomrolloff[i], 0.0f);
        }
        break;
    case Exponent: case ExponentClamped:
        if(attenDistance > 0.0f && P.ref_distance > 0.0f)
        {
            const float dist_ratio = attenDistance / P.ref_distance;
            dryAttnBase = M::pow(dist_ratio, -P.rolloff_factor);
            dryBase *= dryAttnBase;
            for(uint32_t i = 0;i < num_sends;++i)
                wetBase[i] = wetBase[i] * M::pow(dist_ratio, -roomrolloff[i]);
        }
        break;
    default: break;
    }

    // sound cones (:1854-1883); ConeScale is 1 unless __ALSOFT_HALF_ANGLE_CONES is set (:92-104)
    float wetcone = 1.0f, wetconehf = 1.0f;
    if(directional && P.inner_angle < 360.0f)
    {
        constexpr float Rad2Deg = static_cast<float>(180.0 / 3.14159265358979323846);
        const float angle = Rad2Deg*2.0f * M::acos(-dot(direction, tosource)) * 1.0f;
        float conegain = 1.0f, conehf = 1.0f;
        if(angle >= P.outer_angle) { conegain = P.outer_gain; conehf = P.outer_gain_hf; }
        else if(angle >= P.inner_angle)
        {
            const float scale = (angle-P.inner_angle) / (P.outer_angle-P.inner_angle);
            conegain = lerpf(1.0f, P.outer_gain, scale);
            conehf = lerpf(1.0f, P.outer_gain_hf, scale);
        }
        dryBase *= conegain;
        if(P.dry_gain_hf_auto) dryHF *= conehf;
        if(P.wet_gain_auto) wetcone = conegain;
        if(P.wet_gain_hf_auto) wetconehf = conehf;
    }

    // gain limits and filters (:1885-1907)
    const float mingain = fminf_(P.min_gain, P.max_gain), maxgain = P.max_gain;
    dryBase = clampf_(dryBase, mingain, maxgain) * P.direct.gain;
    dryBase = fminf_(kGainMixMax, dryBase * lis.gain);
    dryHF = dryHF * P.direct.gain_hf;
    dryLF = P.direct.gain_lf;
    for(uint32_t i = 0;i < num_sends;++i)
    {
        const float gain = clampf_(wetBase[i]*wetcone, mingain, maxgain) * P.sends[i].gain;
        wetBase[i] = fminf_(kGainMixMax, gain * lis.gain);
        wetHF[i] = P.sends[i].gain_hf * wetconehf;
        wetLF[i] = P.sends[i].gain_lf;
    }

    // air absorption and initial send decay (:1909-1961)
    if(distance > P.ref_distance)
    {
        const float distance_units = (distance-P.ref_distance) * P.rolloff_factor;
        const float distance_meters = distance_units * lis.meters_per_unit;
        const float absorb = distance_meters * P.air_absorption_factor;
        if(absorb > kEpsF) dryHF *= M::pow(lis.air_absorption_gain_hf, absorb);
        for(uint32_t i = P.wet_gain_auto ? 0u : num_sends;i < num_sends;++i)
        {
            const b200mix_source_send &S = P.sends[i];
            if(!S.active || !(S.slot_decay_time > 0.0f)) continue;
            if(S.slot_air_absorption_gain_hf < 1.0f && absorb > kEpsF)
                wetHF[i] *= M::pow(S.slot_air_absorption_gain_hf, absorb);
            const float DecayDistance = S.slot_decay_time * kSpeedOfSound;
            const float fact = distance_meters / DecayDistance;
            const float gain = M::pow(kReverbDecayGain, fact)*(1.0f-dryAttnBase) + dryAttnBase;
            wetBase[i] *= gain;
        }
    }

    // doppler and the resampler step (:1964-2001)
    float pitch = P.pitch;
    {
        const float DopplerFactor = P.doppler_factor * lis.doppler_factor;
        if(DopplerFactor > 0.0f)
        {
            const float vss = dot(velocity, tosource) * -DopplerFactor;
            const float vls = dot(lvelocity, tosource) * -DopplerFactor;
            const float SpeedOfSound = lis.speed_of_sound;
            if(!(vls < SpeedOfSound)) pitch = 0.0f;
            else if(!(vss < SpeedOfSound)) pitch = M::infinity();
            else pitch *= (SpeedOfSound-vls) / (SpeedOfSound-vss);
        }
    }
    pitch *= float(buffer_rate) / float(device_rate);
    if(pitch > float(kMaxPitch)) out.step = kMaxPitch << kFracBits;
    else
    {
        const uint32_t st = uint32_t(M::lrint(pitch * kFracOne));
        out.step = st < 1u ? 1u : st;
    }

    // source radius (:2003-2007)
    float spread = 0.0f;
    if(P.radius > distance) spread = kPi*2.0f - distance/P.radius*kPi;
    else if(distance > 0.0f) spread = M::asin(P.radius/distance) * 2.0f;

    // XScale/YScale/ZScale are 1 unless the reverse-x/y/z compatibility options are set (:109-111)
    out.pos[0] = tosource.v[0]*1.0f; out.pos[1] = tosource.v[1]*1.0f; out.pos[2] = tosource.v[2]*1.0f;
    out.distance = distance; out.spread = spread;
    out.dry_gain = dryBase; out.dry_gain_hf = dryHF; out.dry_gain_lf = dryLF;
    for(uint32_t i = 0;i < B200MIX_MAX_SENDS;++i)
    {
        const bool on = i < num_sends;
        out.wet_gain[i] = on ? wetBase[i] : 0.0f;
        out.wet_gain_hf[i] = on ? wetHF[i] : 1.0f;
        out.wet_gain_lf[i] = on ? wetLF[i] : 1.0f;
    }
    // CalcHrtfPanning's direction for a point source (:1209-1214)
    out.hrtf_elevation = M::asin(clampf_(out.pos[1], -1.0f, 1.0f));
    out.hrtf_azimuth = M::atan2(out.pos[0], -out.pos[2]);
}

// The filter block of CalcPanningAndFilters (alc/alu.cpp:1619-1656) for one path (0 = direct,
// 1+s = send s): a high-shelf at HFReference and a low-shelf at LFReference with the path's
// HF / LF gains; the path's filter is active iff either gain differs from 1.
template<class M> PM_HD void design_filter(const b200mix_source_props &P, uint32_t device_rate, uint32_t path,
    float ghf, float glf, b200mix_voice_filter &f)
{
    const float inv_samplerate = 1.0f / float(device_rate);
    const float hfref = path ? P.sends[path-1].hf_reference : P.direct.hf_reference;
    const float lfref = path ? P.sends[path-1].lf_reference : P.direct.lf_reference;
    f.path = path;
    f.active = (ghf != 1.0f || glf != 1.0f) ? 1u : 0u;
    biquad_coeffs<M>(0u, hfref * inv_samplerate, ghf, 1.0f, f.lowpass);
    biquad_coeffs<M>(1u, lfref * inv_samplerate, glf, 1.0f, f.highpass);
}

// CalcPanningAndFilters for a point source (alc/alu.cpp:1196-1226,1318-1361; a source on the
// listener: :1268-1310,1420-1466) after calc_source_params: dir {elevation, azimuth, distance,
// spread} for the HRIR lookup (render_mode 2) or dry_gains[dry.channels], and
// send_gains[num_sends][wet_stride].  Returns false on a bad map.
template<class M> PM_HD bool calc_panning(const b200mix_source_props &P, const b200mix_source_result &r,
    const b200mix_voice_env &env, float *hrtf_gain, bool *is_hrtf, float dir[4], float *dry_gains,
    float *send_gains)
{
    float coeffs[B200MIX_MAX_AMBI_CHANNELS];
    if(r.distance > kEpsF)
    {
        float pos[3] = {r.pos[0], r.pos[1], r.pos[2]};
        if(env.render_mode == 2u)
        {
            dir[0] = r.hrtf_elevation; dir[1] = r.hrtf_azimuth; dir[2] = r.distance; dir[3] = r.spread;
            *hrtf_gain = r.dry_gain; *is_hrtf = true;
            ambi_coeffs<M>(r.pos, r.spread, coeffs);      // the sends' encoder coefficients
        }
        else
        {
            if(env.render_mode == 1u) pairwise_azimuth<M>(r.pos, pos);
            ambi_coeffs<M>(pos, r.spread, coeffs);        // shared by the dry mix and the sends
            if(!pan_gains(env.dry.channels, env.dry.scale, env.dry.index, coeffs, r.dry_gain,
                dry_gains, env.dry.channels)) return false;
            *is_hrtf = false;
            if(dir) { dir[0] = dir[1] = dir[2] = dir[3] = 0.0f; }
        }
    }
    else
    {
        // A source on the listener: the mono channel sits at the front-centre position of MonoMap
        // (:1471-1473, pan gain 1 with VoiceProps::Panning at its default 0), spread is all or nothing
        const float front[3] = {0.0f, 0.0f, -1.0f};
        if(env.render_mode == 2u)
        {
            dir[0] = M::asin(front[1]); dir[1] = M::atan2(front[0], -front[2]);
            dir[2] = M::infinity(); dir[3] = r.spread;
            *hrtf_gain = r.dry_gain; *is_hrtf = true;
        }
        else
        {
            // ScaleAzimuthFront3 leaves the front-centre direction where it is
            ambi_coeffs<M>(front, r.spread, coeffs);
            if(!pan_gains(env.dry.channels, env.dry.scale, env.dry.index, coeffs, r.dry_gain,
                dry_gains, env.dry.channels)) return false;
            *is_hrtf = false;
            if(dir) { dir[0] = dir[1] = dir[2] = dir[3] = 0.0f; }
        }
        ambi_coeffs<M>(front, r.spread, coeffs);
    }
    for(uint32_t s = 0;s < env.num_sends;++s)
    {
        const b200mix_mix_map &w = env.wet[s];
        if(!send_gains || !env.wet_stride) break;
        float *g = send_gains + size_t(s)*env.wet_stride;
        for(uint32_t c = 0;c < env.wet_stride;++c) g[c] = 0.0f;
        if(!P.sends[s].active || !w.channels) continue;
        if(w.channels > env.wet_stride || !w.scale || !w.index) return false;
        if(!pan_gains(w.channels, w.scale, w.index, coeffs, r.wet_gain[s], g, w.channels)) return false;
    }
    return true;
}

// BsincPrepare (alc/alu.cpp:140-165) from the table's per-scale metadata.
struct BsincMeta { float scaleBase, scaleRange; uint32_t m[16], filterOffset[16]; };
struct BsincPrep { float sf; uint32_t m, l, offset; };
template<class M> PM_HD BsincPrep prepare_bsinc(const BsincMeta &t, uint32_t increment)
{
    BsincPrep st;
    unsigned si = 16u - 1u;
    float sf = 0.0f;
    if(increment > 65536u)
    {
        sf = 65536.0f/float(increment) - t.scaleBase;
        sf = fmaxf_(0.0f, 16.0f*sf*t.scaleRange - 1.0f);
        si = static_cast<unsigned>(sf);
        sf -= float(si);
        // diagonally-symmetric curve reducing scale-transition ripple (alu.cpp:152-157)
        sf = 1.0f - M::sqrt(1.0f - sf*sf);
    }
    st.sf = sf;
    st.m = t.m[si];
    st.l = st.m/2u - 1u;
    st.offset = t.filterOffset[si];
    return st;
}

} } // namespace b200mix::pm
