// efx_oracle.cpp — CPU restatement of EffectState::process for the EFX effects the product runs
// on the GPU (echo, ring modulator, equalizer, compressor, dedicated, distortion, chorus / flanger,
// autowah, vocal morpher, frequency shifter, pitch shifter).  TEST
// INFRASTRUCTURE ONLY: linked into oracle/liboracle.so, used by tests/, smoke() and nothing else.
//
// Each process() below follows the reference line by line (file:line cited); the parameter side
// (deviceUpdate + update: coefficient designs, tap offsets, gain targets) is the shared
// restatement in openal-soft_b200/csrc/efx_math.hpp, whose values are pinned — together with
// these loops — by the golden vectors rendered by the compiled reference (tests/golden/efx_*).
// Compiled with -ffp-contract=off: plain mul/add like the reference's x86-64 build.
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstring>
#include <new>
#include <vector>

#include "efx_oracle.h"
#include "../openal-soft_b200/csrc/efx_math.hpp"

using b200mix::EfxParams;
namespace {
constexpr size_t LINE = 1024;
constexpr float kSilence = 0.00001f;      // GainSilenceThreshold, core/mixer/defs.h:28

struct Biquad { float z1{0.0f}, z2{0.0f}; };
// BiquadFilter::process, core/filters/biquad.cpp:175-200
void biquad_process(const float *c, Biquad &f, const float *src, float *dst, size_t n)
{
    float z1 = f.z1, z2 = f.z2;
    for(size_t i = 0;i < n;++i)
    {
        const float x = src[i];
        const float y = x*c[0] + z1;
        z1 = x*c[1] - y*c[3] + z2;
        z2 = x*c[2] - y*c[4];
        dst[i] = y;
    }
    f.z1 = z1; f.z2 = z2;
}
// BiquadFilter::dualProcess, core/filters/biquad.cpp:254-283
void dual_process(const float *c0, const float *c1, Biquad &f0, Biquad &f1, const float *src, float *dst, size_t n)
{
    float z01 = f0.z1, z02 = f0.z2, z11 = f1.z1, z12 = f1.z2;
    for(size_t i = 0;i < n;++i)
    {
        const float x0 = src[i];
        const float y0 = x0*c0[0] + z01;
        z01 = x0*c0[1] - y0*c0[3] + z02;
        z02 = x0*c0[2] - y0*c0[4];
        const float y1 = y0*c1[0] + z11;
        z11 = y0*c1[1] - y1*c1[3] + z12;
        z12 = y0*c1[2] - y1*c1[4];
        dst[i] = y1;
    }
    f0.z1 = z01; f0.z2 = z02; f1.z1 = z11; f1.z2 = z12;
}
// MixLine, core/mixer/mixer_c.cpp:150-186
void mix_line(const float *in, size_t n, float *dst, float &cur, float target, float delta, size_t fade_len, size_t counter)
{
    const float step = (target - cur) * delta;
    size_t pos = 0;
    if(std::fabs(step) > 1.1920929e-07f)
    {
        const float gain = cur;
        float step_count = 0.0f;
        for(;pos < fade_len;++pos) { dst[pos] += in[pos] * (gain + step*step_count); step_count += 1.0f; }
        if(fade_len < counter) { cur = gain + step*step_count; return; }
    }
    cur = target;
    if(!(std::fabs(target) > kSilence)) return;
    for(;pos < n;++pos) dst[pos] += in[pos]*target;
}
// Mix_C 1 -> many and 1 -> 1, mixer_c.cpp:247-268
void mix_many(const float *in, size_t n, float (*out)[LINE], size_t nout, float *cur, const float *tgt, size_t counter)
{
    const float delta = counter > 0 ? 1.0f/float(counter) : 0.0f;
    const size_t fade_len = std::min(counter, n);
    for(size_t c = 0;c < nout;++c) mix_line(in, n, out[c], cur[c], tgt[c], delta, fade_len, counter);
}
} // namespace

// complex_fft / complex_hilbert (common/alcomplex.cpp:110-215): radix-2 decimation in time in double.
// The reference multiplies its twiddles up recursively from tabulated angles; evaluating them
// directly differs in the last bits of a double, far below the float the effect outputs.
using cplx = std::complex<double>;
void fft_pow2(cplx *x, size_t n, double sign)
{
    for(size_t i = 1, j = 0;i < n;++i)
    {
        size_t bit = n >> 1;
        for(;j & bit;bit >>= 1) j ^= bit;
        j ^= bit;
        if(i < j) std::swap(x[i], x[j]);
    }
    for(size_t half = 1;half < n;half <<= 1)
        for(size_t j = 0;j < half;++j)
        {
            const double ang = 3.14159265358979323846 * double(j) / double(half);
            const cplx w{std::cos(ang), sign*std::sin(ang)};
            for(size_t k = j;k < n;k += half << 1)
            {
                const cplx t = x[k + half] * w;
                x[k + half] = x[k] - t;
                x[k] += t;
            }
        }
}
void hilbert_pow2(cplx *x, size_t n)
{
    fft_pow2(x, n, 1.0);
    const double inv = 1.0 / double(n);
    x[0] *= inv;
    for(size_t i = 1;i < n/2;++i) x[i] *= inv*2.0;
    x[n/2] *= inv;
    for(size_t i = n/2 + 1;i < n;++i) x[i] = cplx{};
    fft_pow2(x, n, -1.0);
}
// gHannWindow<1024>, common/hann_window.hpp:11-26
const float *hann1024()
{
    static float w[1024];
    static bool done = false;
    if(!done)
    {
        for(unsigned i = 0;i < 512;++i)
        {
            const double v = std::sin((i + 1.0) * (3.14159265358979323846 / 1025.0));
            w[i] = static_cast<float>(v * v); w[1023 - i] = w[i];
        }
        done = true;
    }
    return w;
}

struct oefx {
    EfxParams p{};
    float cur[b200mix::kEfxMaxLines][32]{};       // Current gains per line and output channel
    // echo
    std::vector<float> echo_buf; size_t echo_offset{0}; Biquad echo_f;
    // modulator
    uint32_t mod_index{0}, mod_range{1};
    Biquad chan[b200mix::kEfxMaxLines][4];
    // compressor
    float env{1.0f};
    // chorus
    std::vector<float> cho_buf; uint32_t cho_offset{0}, lfo_offset{0}, lfo_range{1};
    float cubic[513]{};
    // autowah
    float wah_env{0.0f};
    // frequency shifter (fshifter.cpp:92-118)
    size_t fs_count{0}, fs_pos{1024 - 256};
    std::vector<double> fs_in; std::vector<std::complex<double>> fs_outfifo, fs_accum, fs_outdata;
    uint32_t fs_phase[4]{};
    // pitch shifter (pshifter.cpp:84-118): mCount, mPos, per-channel mFIFO / mOutputAccum, mLastPhase, mSumPhase
    size_t ps_count{0}, ps_pos{1024 - 128};
    std::vector<float> ps_fifo, ps_accum, ps_last, ps_sum;
    // vocal morpher
    uint32_t vm_index{0}; float vm_cur[b200mix::kEfxMaxLines]{}; float vm_s[b200mix::kEfxMaxLines][2][4][2]{};
};

extern "C" void oracle_build_cubic_filter(float filter[513]);      // tables.c (gCubicTable)

extern "C" {

oefx *oefx_create(const b200mix_efx_props *props, const b200mix_efx_target *target, int *rc)
{
    auto *e = new(std::nothrow) oefx{};
    if(!e) { *rc = B200MIX_ERR_NOMEM; return nullptr; }
    *rc = b200mix::efx_update(*props, *target, e->p);
    if(*rc != B200MIX_OK) { delete e; return nullptr; }
    if(e->p.echo_len) e->echo_buf.assign(e->p.echo_len, 0.0f);
    if(e->p.cho_len) { e->cho_buf.assign(size_t(4)*e->p.cho_len, 0.0f); oracle_build_cubic_filter(e->cubic); }
    if(e->p.type == B200MIX_EFFECT_FSHIFTER)
    {
        e->fs_in.assign(4*1024, 0.0); e->fs_outfifo.assign(4*256, {}); e->fs_accum.assign(4*1024, {});
        e->fs_outdata.assign(4*1024, {});
    }
    if(e->p.type == B200MIX_EFFECT_PSHIFTER)
    {   // PshifterState::deviceUpdate, pshifter.cpp:131-145
        e->ps_fifo.assign(9*1024, 0.0f); e->ps_accum.assign(9*1024, 0.0f);
        e->ps_last.assign(513, 0.0f); e->ps_sum.assign(513, 0.0f);
    }
    e->lfo_range = e->p.cho_lfo_range ? e->p.cho_lfo_range : 1u;
    e->mod_range = e->p.mod_range ? e->p.mod_range : 1u;
    if(e->p.snap_gains) std::memcpy(e->cur, e->p.gains, sizeof(e->cur));
    return e;
}

int oefx_update(oefx *e, const b200mix_efx_props *props, const b200mix_efx_target *target)
{
    EfxParams P;
    if(int rc = b200mix::efx_update(*props, *target, P)) return rc;
    if(P.type != e->p.type || P.lines != e->p.lines || P.echo_len != e->p.echo_len || P.cho_len != e->p.cho_len)
        return B200MIX_ERR_INVALID;
    if(P.type == B200MIX_EFFECT_CHORUS)
    {   // mLfoOffset follows the LFO range, chorus.cpp:185-211
        e->lfo_offset = P.cho_rate_on ? e->lfo_offset * P.cho_lfo_range_new / e->lfo_range : 0u;
        e->lfo_range = P.cho_lfo_range;
    }
    if(P.type == B200MIX_EFFECT_MODULATOR)
    {   // mIndex rescale, modulator.cpp:117-118
        e->mod_index = uint32_t(uint64_t(e->mod_index) * P.mod_range_new / e->mod_range);
        e->mod_range = P.mod_range;
    }
    if(P.type == B200MIX_EFFECT_FSHIFTER)
        for(int c = 0;c < 4;++c) if(P.fs_reset_phase[c]) e->fs_phase[c] = 0u;      // fshifter.cpp:189-192,205-208
    if(P.type == B200MIX_EFFECT_VMORPHER)
        std::memset(e->vm_s, 0, sizeof(e->vm_s));       // update() installs new FormantFilters, vmorpher.cpp:252-260
    e->p = P;
    if(P.snap_gains) std::memcpy(e->cur, P.gains, sizeof(e->cur));
    return B200MIX_OK;
}

uint32_t oefx_type(const oefx *e) { return e->p.type; }
void oefx_free(oefx *e) { delete e; }

void oefx_process(oefx *e, size_t n, const float (*in)[1024], size_t nin_, float (*out)[1024], size_t nout)
{
    const EfxParams &P = e->p;
    const size_t nin = std::min<size_t>(nin_, P.in_channels);
    static thread_local float buf[LINE], tmp0[LINE], tmp1[LINE], mod[LINE];
    switch(P.type)
    {
    case B200MIX_EFFECT_DEDICATED:
        // dedicated.cpp:105-109
        mix_many(in[0], n, out, nout, e->cur[0], P.gains[0], n);
        break;
    case B200MIX_EFFECT_ECHO:
    {
        // echo.cpp:133-157
        const size_t mask = e->echo_buf.size() - 1;
        float *delaybuf = e->echo_buf.data();
        size_t offset = e->echo_offset;
        size_t tap1 = offset - P.echo_tap[0], tap2 = offset - P.echo_tap[1];
        for(size_t i = 0;i < n;++i)
        {
            offset &= mask; tap1 &= mask; tap2 &= mask;
            delaybuf[offset] = in[0][i];
            tmp0[i] = delaybuf[tap1++];
            tmp1[i] = delaybuf[tap2++];
            const float feedb = tmp1[i];
            const float y = feedb*P.echo_filter[0] + e->echo_f.z1;                 // processOne
            e->echo_f.z1 = feedb*P.echo_filter[1] - y*P.echo_filter[3] + e->echo_f.z2;
            e->echo_f.z2 = feedb*P.echo_filter[2] - y*P.echo_filter[4];
            delaybuf[offset++] += y * P.echo_feed;
        }
        e->echo_offset = offset;
        mix_many(tmp0, n, out, nout, e->cur[0], P.gains[0], n);
        mix_many(tmp1, n, out, nout, e->cur[1], P.gains[1], n);
        break;
    }
    case B200MIX_EFFECT_MODULATOR:
    {
        // modulator.cpp:157-199
        uint32_t index = e->mod_index;
        for(size_t i = 0;i < n;++i)
        {
            float m = 1.0f;
            if(P.mod_wave == 1u) m = std::sin(float(index) * P.mod_scale);
            else if(P.mod_wave == 2u) m = float(index)*P.mod_scale - 1.0f;
            else if(P.mod_wave == 3u) m = float(float(index)*P.mod_scale < 0.5f)*2.0f - 1.0f;
            mod[i] = m;
            if(++index == P.mod_range) index = 0;
        }
        e->mod_index = index;
        for(size_t c = 0;c < nin;++c)
        {
            if(!P.line_on[c]) continue;
            biquad_process(P.mod_hp, e->chan[c][0], in[c], buf, n);
            for(size_t i = 0;i < n;++i) buf[i] = buf[i] * mod[i];
            // MixSamples 1 -> 1 to the channel of the same ambisonic index, Counter = min(n, 64)
            for(size_t o = 0;o < nout;++o)
                if(P.gains[c][o] != 0.0f || e->cur[c][o] != 0.0f)
                {
                    const size_t counter = std::min<size_t>(n, 64);
                    mix_line(buf, n, out[o], e->cur[c][o], P.gains[c][o], 1.0f/float(counter), counter, counter);
                }
        }
        break;
    }
    case B200MIX_EFFECT_EQUALIZER:
        // equalizer.cpp:165-183
        for(size_t c = 0;c < nin;++c)
        {
            if(!P.line_on[c]) continue;
            dual_process(P.eq[0], P.eq[1], e->chan[c][0], e->chan[c][1], in[c], buf, n);
            dual_process(P.eq[2], P.eq[3], e->chan[c][2], e->chan[c][3], buf, buf, n);
            for(size_t o = 0;o < nout;++o)
                if(P.gains[c][o] != 0.0f || e->cur[c][o] != 0.0f)
                    mix_line(buf, n, out[o], e->cur[c][o], P.gains[c][o], 1.0f/float(n), n, n);
        }
        break;
    case B200MIX_EFFECT_COMPRESSOR:
    {
        // compressor.cpp:111-177
        float env = e->env;
        for(size_t i = 0;i < n;++i)
        {
            const float amplitude = P.comp_enabled ? std::clamp(std::fabs(in[0][i]), 0.5f, 2.0f) : 1.0f;
            if(amplitude > env) env = std::min(env*P.comp_attack, amplitude);
            else if(amplitude < env) env = std::max(env*P.comp_release, amplitude);
            buf[i] = 1.0f / env;
        }
        e->env = env;
        for(size_t c = 0;c < nin;++c)
        {
            if(!P.line_on[c]) continue;
            for(size_t o = 0;o < nout;++o)
            {
                const float gain = P.gains[c][o];
                if(std::fabs(gain) > kSilence)
                    for(size_t i = 0;i < n;++i) out[o][i] += in[c][i] * buf[i] * gain;
            }
        }
        break;
    }
    case B200MIX_EFFECT_DISTORTION:
    {
        // distortion.cpp:198-303 (first-order devices)
        static const float dc = static_cast<float>(0.25 / 1.7320508075688772935);
        static const float ec = static_cast<float>(0.5 * 1.7320508075688772935);
        static const float B2A[4][4] = {{0.25f, dc, dc, dc}, {0.25f, -dc, -dc, dc}, {0.25f, dc, -dc, -dc}, {0.25f, -dc, dc, -dc}};
        static const float A2B[4][4] = {{1.0f, 1.0f, 1.0f, 1.0f}, {ec, -ec, ec, -ec}, {ec, -ec, -ec, ec}, {ec, ec, -ec, -ec}};
        static thread_local float A[4][LINE], B[4][LINE];
        const size_t numInput = std::min<size_t>(nin, 4);
        for(size_t c = 0;c < 4;++c)
        {
            for(size_t i = 0;i < n;++i) A[c][i] = 0.0f;
            for(size_t k = 0;k < numInput;++k)
                for(size_t i = 0;i < n;++i) A[c][i] = A[c][i] + in[k][i]*B2A[c][k];
        }
        for(auto &row : B) std::fill_n(row, n, 0.0f);
        const float fc = P.dist_edge;
        for(size_t base = 0;base < n;)
        {
            const size_t todo = std::min<size_t>(LINE, (n-base)*4);
            for(size_t c = 0;c < 4;++c)
            {
                for(size_t i = 0;i < todo;++i) tmp0[i] = !(i&3) ? A[c][(i>>2)+base] * 4.0f : 0.0f;
                biquad_process(P.dist_lp, e->chan[c][0], tmp0, tmp1, todo);
                for(size_t i = 0;i < todo;++i)
                {
                    float smp = tmp1[i];
                    smp = ( 1.0f + fc) * smp/(1.0f + fc*std::fabs(smp));
                    smp = (-1.0f - fc) * smp/(1.0f + fc*std::fabs(smp));
                    smp = ( 1.0f + fc) * smp/(1.0f + fc*std::fabs(smp));
                    tmp0[i] = smp;
                }
                biquad_process(P.dist_bp, e->chan[c][1], tmp0, tmp1, todo);
                for(size_t k = 0;k < 4;++k)
                    for(size_t i = 0;i < (todo>>2);++i) B[k][base+i] += tmp1[i*4] * A2B[k][c];
            }
            base += todo >> 2;
        }
        for(size_t c = 0;c < 4;++c)
        {
            if(!P.line_on[c]) continue;
            for(size_t o = 0;o < nout;++o)
                if(P.gains[c][o] != 0.0f || e->cur[c][o] != 0.0f)
                    mix_line(B[c], n, out[o], e->cur[c][o], P.gains[c][o], 1.0f/float(n), n, n);
        }
        break;
    }
    case B200MIX_EFFECT_CHORUS:
    {
        // chorus.cpp:326-425 (first-order devices)
        static const float dc = static_cast<float>(0.25 / 1.7320508075688772935);
        static const float ec = static_cast<float>(0.5 * 1.7320508075688772935);
        static const float B2A[4][4] = {{0.25f, dc, dc, dc}, {0.25f, dc, -dc, -dc}, {0.25f, -dc, -dc, dc}, {0.25f, -dc, dc, -dc}};
        static const float A2B[4][4] = {{1.0f, 1.0f, 1.0f, 1.0f}, {ec, ec, -ec, -ec}, {ec, -ec, -ec, ec}, {ec, -ec, ec, -ec}};
        static thread_local float A[4][LINE], B[4][LINE];
        static thread_local uint32_t md[2][LINE];
        const size_t numInput = std::min<size_t>(nin, 4);
        for(size_t c = 0;c < 4;++c)
        {
            for(size_t i = 0;i < n;++i) A[c][i] = 0.0f;
            for(size_t k = 0;k < numInput;++k)
                for(size_t i = 0;i < n;++i) A[c][i] = A[c][i] + in[k][i]*B2A[c][k];
        }
        for(auto &row : B) std::fill_n(row, n, 0.0f);
        // calcTriangleDelays / calcSinusoidDelays, chorus.cpp:235-323
        auto gen = [&P](uint32_t offset) -> uint32_t {
            const float offset_norm = float(offset) * P.cho_lfo_scale;
            const float v = P.cho_wave == 1u ? (1.0f - std::fabs(2.0f - offset_norm)) * P.cho_depth
                                              : std::sin(offset_norm) * P.cho_depth;
            return uint32_t(int(std::lrintf(v)) + P.cho_delay);            // fastf2i = cvtss2si (current rounding mode)
        };
        const uint32_t range = e->lfo_range;
        uint32_t off = e->lfo_offset;
        for(size_t i = 0;i < n;++i) { md[0][i] = gen(off++); if(off == range) off = 0; }
        off = (e->lfo_offset + P.cho_lfo_disp) % range;
        for(size_t i = 0;i < n;++i) { md[1][i] = gen(off++); if(off == range) off = 0; }
        e->lfo_offset = uint32_t(e->lfo_offset + n) % range;
        const uint32_t bufmask = P.cho_len - 1u;
        const uint32_t avgdelay = (uint32_t(P.cho_delay) + 32768u) >> 16;
        for(size_t c = 0;c < 4;++c)
        {
            const uint32_t *moddelays = md[c < 2 ? 0 : 1];
            float *delaybuf = e->cho_buf.data() + size_t(c)*P.cho_len;
            uint32_t offset = e->cho_offset;
            for(size_t i = 0;i < n;++i)
            {
                delaybuf[offset&bufmask] = A[c][i];
                const uint32_t delay = offset - (moddelays[i] >> 8), phase = moddelays[i] & 255u;
                const float sample = delaybuf[(delay+1) & bufmask]*e->cubic[256 + phase] +
                    delaybuf[(delay  ) & bufmask]*e->cubic[phase] +
                    delaybuf[(delay-1) & bufmask]*e->cubic[256 - phase] +
                    delaybuf[(delay-2) & bufmask]*e->cubic[512 - phase];
                delaybuf[offset&bufmask] += delaybuf[(offset-avgdelay) & bufmask] * P.cho_feedback;
                ++offset;
                tmp0[i] = sample;
            }
            for(size_t k = 0;k < 4;++k)
                for(size_t i = 0;i < n;++i) B[k][i] = B[k][i] + tmp0[i]*A2B[k][c];
        }
        e->cho_offset += uint32_t(n);
        for(size_t c = 0;c < 4;++c)
        {
            if(!P.line_on[c]) continue;
            for(size_t o = 0;o < nout;++o)
                if(P.gains[c][o] != 0.0f || e->cur[c][o] != 0.0f)
                    mix_line(B[c], n, out[o], e->cur[c][o], P.gains[c][o], 1.0f/float(n), n, n);
        }
        break;
    }
    case B200MIX_EFFECT_AUTOWAH:
    {
        // autowah.cpp:136-205
        static thread_local float cosw[LINE], alpha[LINE];
        float env_delay = e->wah_env;
        for(size_t i = 0;i < n;++i)
        {
            const float sample = P.wah_peak_gain * std::fabs(in[0][i]);
            const float a = (sample > env_delay) ? P.wah_attack : P.wah_release;
            env_delay = sample + (env_delay - sample)*a;
            const float w0 = std::min(P.wah_bandwidth*env_delay + P.wah_freq_min, 0.46f) * (3.14159265358979323846f*2.0f);
            cosw[i] = std::cos(w0); alpha[i] = std::sin(w0)*(0.5f/5.0f);
        }
        e->wah_env = env_delay;
        for(size_t c = 0;c < nin;++c)
        {
            if(!P.line_on[c]) continue;
            float z1 = e->chan[c][0].z1, z2 = e->chan[c][0].z2;
            const float rg = P.wah_res_gain;
            for(size_t i = 0;i < n;++i)
            {
                const float b0 = 1.0f + alpha[i]*rg, b1 = -2.0f * cosw[i], b2 = 1.0f - alpha[i]*rg;
                const float a0 = 1.0f / (1.0f + alpha[i]/rg), a1 = -2.0f * cosw[i], a2 = 1.0f - alpha[i]/rg;
                const float input = in[c][i];
                const float output = input*(b0*a0) + z1;
                z1 = input*(b1*a0) - output*(a1*a0) + z2;
                z2 = input*(b2*a0) - output*(a2*a0);
                buf[i] = output;
            }
            e->chan[c][0].z1 = z1; e->chan[c][0].z2 = z2;
            for(size_t o = 0;o < nout;++o)
                if(P.gains[c][o] != 0.0f || e->cur[c][o] != 0.0f)
                    mix_line(buf, n, out[o], e->cur[c][o], P.gains[c][o], 1.0f/float(n), n, n);
        }
        break;
    }
    case B200MIX_EFFECT_VMORPHER:
    {
        // vmorpher.cpp:272-330
        static thread_local float lfo[256], bufA[256], bufB[256], blended[256];
        for(size_t base = 0;base < n;)
        {
            const size_t td = std::min<size_t>(256, n - base);
            {   // Oscillate<func>, vmorpher.cpp:86-96
                uint32_t index = e->vm_index;
                for(size_t i = 0;i < td;++i)
                {
                    index += P.vm_step; index &= 0xffffffu;
                    switch(P.vm_wave)
                    {
                    case 0: lfo[i] = 0.5f; break;
                    case 1: lfo[i] = std::sin(static_cast<float>(index) * (3.14159265358979323846f*2.0f / 16777216.0f))*0.5f + 0.5f; break;
                    case 2: lfo[i] = std::fabs(static_cast<float>(index)*(2.0f/16777216.0f) - 1.0f); break;
                    default: lfo[i] = static_cast<float>(index) / 16777216.0f; break;
                    }
                }
            }
            e->vm_index += uint32_t(P.vm_step * td);
            e->vm_index &= 0xffffffu;
            for(size_t c = 0;c < nin;++c)
            {
                const uint32_t outidx = P.vm_target[c];
                if(outidx == 0xffffffffu) continue;
                float *acc[2] = {bufA, bufB};
                for(int v = 0;v < 2;++v)
                {
                    std::fill_n(acc[v], td, 0.0f);
                    for(int f = 0;f < 4;++f)
                    {   // FormantFilter::process, vmorpher.cpp:106-140
                        const float g = P.vm_coeff[v][f], gain = P.vm_fgain[v][f];
                        const float h = 1.0f / (1.0f + (g*(1.0f/5.0f)) + (g*g));
                        const float coeff = (1.0f/5.0f) + g;
                        float s1 = e->vm_s[c][v][f][0], s2 = e->vm_s[c][v][f][1];
                        for(size_t i = 0;i < td;++i)
                        {
                            const float H = (in[c][base+i] - coeff*s1 - s2)*h;
                            const float B = g*H + s1;
                            const float L = g*B + s2;
                            s1 = g*H + B;
                            s2 = g*B + L;
                            acc[v][i] = acc[v][i] + B*gain;
                        }
                        e->vm_s[c][v][f][0] = s1; e->vm_s[c][v][f][1] = s2;
                    }
                }
                for(size_t i = 0;i < td;++i) blended[i] = bufA[i] + (bufB[i] - bufA[i])*lfo[i];
                const size_t counter = n - base;
                mix_line(blended, td, out[outidx] + base, e->vm_cur[c], P.vm_tgain[c], 1.0f/float(counter), std::min(counter, td), counter);
            }
            base += td;
        }
        break;
    }
    case B200MIX_EFFECT_FSHIFTER:
    {
        // fshifter.cpp:235-366 (first-order devices)
        static const float dc = static_cast<float>(0.25 / 1.7320508075688772935), ec = static_cast<float>(0.5 * 1.7320508075688772935);
        const float B2A[4][4] = {{0.25f, dc, dc, dc}, {0.25f, dc, -dc, -dc}, {0.25f, -dc, -dc, dc}, {0.25f, -dc, dc, -dc}};
        const float A2B[4][4] = {{1.0f, 1.0f, 1.0f, 1.0f}, {ec, ec, -ec, -ec}, {ec, -ec, -ec, ec}, {ec, -ec, ec, -ec}};
        static thread_local float bbuf[4][LINE];
        static thread_local cplx analytic[1024];
        const float *win = hann1024();
        for(int i = 0;i < 4;++i) std::fill_n(bbuf[i], n, 0.0f);
        const size_t numInput = std::min<size_t>(nin, 4);
        for(size_t base = 0;base < n;)
        {
            const size_t todo = std::min<size_t>(256 - e->fs_count, n - base);
            for(size_t c = 0;c < 4;++c)
            {
                double *infifo = e->fs_in.data() + c*1024 + e->fs_pos + e->fs_count;
                std::fill_n(infifo, todo, 0.0);
                for(size_t i = 0;i < numInput;++i)
                    for(size_t k = 0;k < todo;++k)
                        infifo[k] = infifo[k] + double(in[i][base+k])*double(B2A[c][i]);
                for(size_t k = 0;k < todo;++k)
                    e->fs_outdata[c*1024 + base + k] = e->fs_outfifo[c*256 + e->fs_count + k];
            }
            e->fs_count += todo; base += todo;
            if(e->fs_count < 256) break;
            e->fs_count = 0;
            e->fs_pos = (e->fs_pos + 256) & 1023;
            const size_t pos = e->fs_pos;
            for(size_t c = 0;c < 4;++c)
            {
                const double *fifo = e->fs_in.data() + c*1024;
                cplx *accum = e->fs_accum.data() + c*1024;
                for(size_t k = 0;k < 1024;++k) analytic[k] = cplx{fifo[(pos + k) & 1023] * double(win[k]), 0.0};
                hilbert_pow2(analytic, 1024);
                for(size_t k = 0;k < 1024;++k) analytic[k] = (2.0/4.0*double(win[k])) * analytic[k];
                for(size_t k = 0;k < 1024;++k) accum[(pos + k) & 1023] += analytic[k];
                for(size_t k = 0;k < 256;++k) { e->fs_outfifo[c*256 + k] = accum[pos + k]; accum[pos + k] = cplx{}; }
            }
        }
        for(size_t c = 0;c < 4;++c)
        {
            const double sign = double(P.fs_sign[c]);
            const uint32_t pstep = P.fs_phase_step[c];
            uint32_t pidx = e->fs_phase[c];
            for(size_t k = 0;k < n;++k)
            {
                const double phase = pidx * (3.14159265358979323846*2.0 / 65536.0);
                const cplx v = e->fs_outdata[c*1024 + k];
                buf[k] = static_cast<float>(v.real()*std::cos(phase) + v.imag()*std::sin(phase)*sign);
                pidx += pstep; pidx &= 0xffffu;
            }
            e->fs_phase[c] = pidx;
            for(size_t i = 0;i < 4;++i)
                for(size_t k = 0;k < n;++k) bbuf[i][k] = bbuf[i][k] + buf[k]*A2B[i][c];
        }
        for(size_t c = 0;c < 4;++c)
            if(P.line_on[c])
                for(size_t o = 0;o < nout;++o)
                    if(P.gains[c][o] != 0.0f || e->cur[c][o] != 0.0f)
                        mix_line(bbuf[c], n, out[o], e->cur[c][o], P.gains[c][o], 1.0f/float(n), n, n);
        break;
    }
    case B200MIX_EFFECT_PSHIFTER:
    {
        // PshifterState::process, pshifter.cpp:207-472 (devices up to second order).  The reference's
        // real FFT (pffft, single precision, ordered output) is evaluated here as a complex FFT in
        // double whose results are rounded to float: same transform, rounding differences only.
        constexpr size_t N = 1024, H = 512, STEP = 128;
        constexpr float pi = 3.14159265358979323846f, inv_pi = 0.318309886183790671538f;
        constexpr float expected_cycles = pi*2.0f / 8.0f;
        static thread_local float bbuf[9][LINE];
        static thread_local cplx X[N];
        static thread_local float fre[H+1], fim[H+1], smag[H+1], sfb[H+1];
        const float *win = hann1024();
        const size_t numInput = std::min<size_t>(nin, 9);
        const uint32_t psi = P.ps_shift_i; const float ps = P.ps_shift;
        auto f2i = [](float f) { return b200mix::efx_detail::f2i(f); };
        for(size_t base = 0;base < n;)
        {
            const size_t todo = std::min<size_t>(STEP - e->ps_count, n - base);
            for(size_t c = 0;c < numInput;++c)
            {
                float *fifo = e->ps_fifo.data() + c*N + e->ps_pos + e->ps_count;
                for(size_t k = 0;k < todo;++k) { bbuf[c][base + k] = fifo[k]; fifo[k] = in[c][base + k]; }
            }
            e->ps_count += todo; base += todo;
            if(e->ps_count < STEP) break;
            e->ps_count = 0;
            e->ps_pos = (e->ps_pos + STEP) & (N - 1);
            const size_t pos = e->ps_pos;
            for(size_t c = 0;c < numInput;++c)
            {
                float *fifo = e->ps_fifo.data() + c*N;
                float *accum = e->ps_accum.data() + c*N;
                for(size_t k = 0;k < N;++k) X[k] = cplx{double(fifo[(pos + k) & (N - 1)] * win[k]), 0.0};
                fft_pow2(X, N, -1.0);
                for(size_t k = 0;k <= H;++k) { fre[k] = float(X[k].real()); fim[k] = (k == 0 || k == H) ? 0.0f : float(X[k].imag()); }
                for(size_t k = 0;k <= H;++k) { smag[k] = 0.0f; sfb[k] = 0.0f; }
                if(c == 0)
                {
                    for(size_t k = 0;k <= H;++k)
                    {
                        const float magnitude = std::hypot(fre[k], fim[k]);        // std::abs(complex<float>)
                        const float phase = std::atan2(fim[k], fre[k]);            // std::arg
                        const float bin_offset = float(k & 7u);
                        float tmp = (phase - e->ps_last[k]) - bin_offset*expected_cycles;
                        e->ps_last[k] = phase;
                        tmp *= inv_pi;
                        const int qpd = f2i(tmp);
                        tmp -= float(qpd + (qpd%2));
                        tmp *= 0.5f*8.0f;
                        const float freqbin = float(k) + tmp;
                        const size_t j = (k*size_t(psi) + 32768u) >> 16;
                        if(j < H+1)
                        {
                            if(smag[j] < magnitude) sfb[j] = freqbin * ps;
                            smag[j] += magnitude;
                        }
                    }
                    for(size_t k = 0;k <= H;++k)
                    {
                        const float bin_offset = float(k & ~size_t{7});
                        float tmp = (sfb[k] - bin_offset) * expected_cycles;
                        tmp = (tmp + e->ps_sum[k]) * inv_pi;
                        const int qpd = f2i(tmp);
                        tmp -= float(qpd + (qpd%2));
                        e->ps_sum[k] = tmp * pi;
                        fre[k] = smag[k] * std::cos(e->ps_sum[k]);                 // std::polar
                        fim[k] = smag[k] * std::sin(e->ps_sum[k]);
                    }
                }
                else
                {
                    constexpr uint32_t bin_limit = ((uint32_t(H)+1u) << 16) - 32768u - 1u;
                    const size_t bin_count = std::min<size_t>(H+1, bin_limit/psi + 1u);
                    for(size_t k = 0;k < bin_count;++k)
                    {
                        const float magnitude = std::hypot(fre[k], fim[k]);
                        const float phasediff = std::atan2(fim[k], fre[k]) - e->ps_last[k];
                        const size_t j = (k*size_t(psi) + 32768u) >> 16;
                        if(smag[j] < magnitude) sfb[j] = phasediff;
                        smag[j] += magnitude;
                    }
                    for(size_t k = 0;k <= H;++k)
                    {
                        float tmp = e->ps_sum[k] + sfb[k];
                        tmp *= inv_pi;
                        const int qpd = f2i(tmp);
                        tmp -= float(qpd + (qpd%2));
                        const float phase = tmp * pi;
                        fre[k] = smag[k] * std::cos(phase);
                        fim[k] = smag[k] * std::sin(phase);
                    }
                }
                // the half-complex spectrum back to time: bins 0 and 512 contribute their real parts only
                X[0] = cplx{double(fre[0]), 0.0}; X[H] = cplx{double(fre[H]), 0.0};
                for(size_t k = 1;k < H;++k) { X[k] = cplx{double(fre[k]), double(fim[k])}; X[N - k] = std::conj(X[k]); }
                fft_pow2(X, N, 1.0);
                constexpr float scale = 3.0f / 8.0f / 1024.0f;
                for(size_t k = 0;k < N;++k)
                {
                    const float v = win[k]*float(X[k].real())*scale;
                    accum[(pos + k) & (N - 1)] += v;
                }
                for(size_t k = 0;k < STEP;++k) { fifo[pos + k] = accum[pos + k]; accum[pos + k] = 0.0f; }
            }
        }
        for(size_t c = 0;c < numInput;++c)
            if(P.line_on[c])
                for(size_t o = 0;o < nout;++o)
                    if(P.gains[c][o] != 0.0f || e->cur[c][o] != 0.0f)
                        mix_line(bbuf[c], n, out[o], e->cur[c][o], P.gains[c][o], 1.0f/float(n), n, n);
        break;
    }
    default: break;
    }
}

} // extern "C"
