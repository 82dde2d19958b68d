// hrtf_store.cpp — host-side HRTF data set: the MHR ("MinPHR03") loader and the
// bilinear HRIR lookup the parameter stage performs per moving voice
// (HrtfStore::getCoeffs, core/hrtf.cpp:192-260; loader core/hrtf_loader.cpp:583-721).
// This is the first step of SURVEY §8(f)#1 (parameter stage): it lets a host — and
// bench.py — derive b200mix_voice_params HRIRs from source directions with the
// reference's own data set (hrtf/Default HRTF.mhr) instead of synthetic filters.
#include "../../include/b200mix.h"
#include "hrtf_store.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <new>
#include <numbers>
#include <vector>

namespace {

struct Reader {
    const uint8_t *p; size_t n, pos{0}; bool ok{true};
    uint32_t le(unsigned bytes)
    {
        if(pos + bytes > n) { ok = false; return 0; }
        uint32_t v = 0;
        for(unsigned i = 0;i < bytes;++i) v |= uint32_t(p[pos+i]) << (8*i);
        pos += bytes;
        return v;
    }
    int32_t s24()
    {
        const uint32_t v = le(3);
        return int32_t((v ^ 0x800000u) - 0x800000u);
    }
};

constexpr float kPassthru = 0.70710678118654752440f;   // PassthruCoeff, core/hrtf.cpp:81

struct IdxBlend { uint32_t idx; float blend; };

// CalcEvIndex / CalcAzIndex, core/hrtf.cpp:162-181
IdxBlend EvIndex(uint32_t evcount, float ev)
{
    ev = (std::numbers::inv_pi_v<float>*ev + 0.5f) * float(evcount-1);
    const uint32_t idx = ev > 0.0f ? uint32_t(ev) : 0u;
    return {std::min(idx, evcount-1u), ev - float(idx)};
}
IdxBlend AzIndex(uint32_t azcount, float az)
{
    az = (std::numbers::inv_pi_v<float>*0.5f*az + 1.0f) * float(azcount);
    const uint32_t idx = az > 0.0f ? uint32_t(az) : 0u;
    return {idx % azcount, az - float(idx)};
}

} // namespace

extern "C" {

int b200mix_hrtf_load(const void *data, size_t bytes, b200mix_hrtf **out)
{
    if(!data || !out || bytes < 16) return B200MIX_ERR_INVALID;
    const auto *b = static_cast<const uint8_t*>(data);
    if(std::memcmp(b, "MinPHR03", 8) != 0) return B200MIX_ERR_UNSUPPORTED;
    Reader r{b, bytes, 8};
    auto *h = new(std::nothrow) b200mix_hrtf{};
    if(!h) return B200MIX_ERR_NOMEM;
    auto fail = [&](int code) { delete h; return code; };

    h->sample_rate = r.le(4);
    const uint32_t channelType = r.le(1);
    h->ir_size = r.le(1);
    const uint32_t fdCount = r.le(1);
    if(!r.ok || channelType > 1 || h->ir_size < 8 || h->ir_size > B200MIX_HRIR_LENGTH
        || fdCount < 1 || fdCount > 16) return fail(B200MIX_ERR_INVALID);
    try {
    for(uint32_t f = 0;f < fdCount;++f)
    {
        const uint32_t dist = r.le(2), evCount = r.le(1);
        if(!r.ok || evCount < 5 || evCount > 181) return fail(B200MIX_ERR_INVALID);
        h->fields.push_back({float(dist)/1000.0f, evCount});
        for(uint32_t e = 0;e < evCount;++e)
        {
            const uint32_t az = r.le(1);
            if(!r.ok || az < 1) return fail(B200MIX_ERR_INVALID);
            h->elevs.push_back({az, 0u});
        }
    }
    } catch(const std::bad_alloc&) { return fail(B200MIX_ERR_NOMEM); }
    uint32_t total = 0;
    for(auto &e : h->elevs) { e.ir_offset = total; total += e.az_count; }
    const uint32_t irs = h->ir_size;
    // the header's counts must be backed by data before anything is sized from them: 3 bytes per
    // tap and one delay byte per response (left ear only for channelType 0)
    {
        const size_t ears = channelType == 0 ? 1u : 2u;
        const size_t need = size_t(total)*ears*(size_t(irs)*3u + 1u);
        if(r.pos > bytes || bytes - r.pos < need) return fail(B200MIX_ERR_INVALID);
    }
    try
    {
        h->coeffs.assign(size_t(total)*irs*2, 0.0f);
        h->delays.assign(size_t(total)*2, 0);
    }
    catch(const std::bad_alloc&) { return fail(B200MIX_ERR_NOMEM); }
    if(channelType == 0)
    {
        for(uint32_t i = 0;i < total;++i)
            for(uint32_t j = 0;j < irs;++j)
                h->coeffs[(size_t(i)*irs + j)*2] = float(r.s24()) / 8388608.0f;
        for(uint32_t i = 0;i < total;++i) h->delays[i*2] = uint8_t(r.le(1));
        if(!r.ok) return fail(B200MIX_ERR_INVALID);
        // MirrorLeftHrirs, core/hrtf_loader.cpp:135-152
        for(const auto &e : h->elevs)
            for(uint32_t j = 0;j < e.az_count;++j)
            {
                const uint32_t l = e.ir_offset + j, rr = e.ir_offset + ((e.az_count - j) % e.az_count);
                for(uint32_t k = 0;k < irs;++k)
                    h->coeffs[(size_t(rr)*irs + k)*2 + 1] = h->coeffs[(size_t(l)*irs + k)*2];
                h->delays[rr*2 + 1] = h->delays[l*2];
            }
    }
    else
    {
        for(uint32_t i = 0;i < total;++i)
            for(uint32_t j = 0;j < irs*2;++j)
                h->coeffs[size_t(i)*irs*2 + j] = float(r.s24()) / 8388608.0f;
        for(uint32_t i = 0;i < total*2;++i) h->delays[i] = uint8_t(r.le(1));
        if(!r.ok) return fail(B200MIX_ERR_INVALID);
    }
    for(uint8_t dl : h->delays) if(dl > (63u<<2)) return fail(B200MIX_ERR_INVALID);
    *out = h;
    return B200MIX_OK;
}

void b200mix_hrtf_free(b200mix_hrtf *h) { delete h; }

int b200mix_hrtf_info(const b200mix_hrtf *h, uint32_t *sample_rate, uint32_t *ir_size,
    uint32_t *ir_count)
{
    if(!h) return B200MIX_ERR_INVALID;
    if(sample_rate) *sample_rate = h->sample_rate;
    if(ir_size) *ir_size = h->ir_size;
    if(ir_count) *ir_count = uint32_t(h->delays.size()/2);
    return B200MIX_OK;
}

int b200mix_hrtf_get_coeffs(const b200mix_hrtf *h, float elevation, float azimuth, float distance,
    float spread, float *coeffs, uint32_t delays[2])
{
    if(!h || !coeffs || !delays) return B200MIX_ERR_INVALID;
    const float dirfact = 1.0f - (std::numbers::inv_pi_v<float>/2.0f * spread);
    size_t ebase = 0, fi = 0;
    for(;fi+1 < h->fields.size();++fi)
    {
        if(distance >= h->fields[fi].distance) break;
        ebase += h->fields[fi].ev_count;
    }
    const uint32_t evCount = h->fields[fi].ev_count;
    const IdxBlend e0 = EvIndex(evCount, elevation);
    const uint32_t e1 = std::min(e0.idx+1u, evCount-1u);
    const auto &el0 = h->elevs[ebase + e0.idx];
    const auto &el1 = h->elevs[ebase + e1];
    const IdxBlend a0 = AzIndex(el0.az_count, azimuth), a1 = AzIndex(el1.az_count, azimuth);
    const uint32_t idx[4] = {el0.ir_offset + a0.idx, el0.ir_offset + ((a0.idx+1u) % el0.az_count),
        el1.ir_offset + a1.idx, el1.ir_offset + ((a1.idx+1u) % el1.az_count)};
    const float blend[4] = {(1.0f-e0.blend)*(1.0f-a0.blend)*dirfact, (1.0f-e0.blend)*a0.blend*dirfact,
        e0.blend*(1.0f-a1.blend)*dirfact, e0.blend*a1.blend*dirfact};
    for(int ear = 0;ear < 2;++ear)
    {
        const float dsum = float(h->delays[idx[0]*2+ear])*blend[0] + float(h->delays[idx[1]*2+ear])*blend[1]
            + float(h->delays[idx[2]*2+ear])*blend[2] + float(h->delays[idx[3]*2+ear])*blend[3];
        // fastf2u: round to nearest (even) under the default rounding mode
        delays[ear] = uint32_t(std::lrintf(dsum * 0.25f));
    }
    const uint32_t irs = h->ir_size;
    for(uint32_t k = 0;k < irs*2;++k) coeffs[k] = 0.0f;
    coeffs[0] = kPassthru * (1.0f-dirfact);
    coeffs[1] = kPassthru * (1.0f-dirfact);
    for(int c = 0;c < 4;++c)
    {
        const float *src = h->coeffs.data() + size_t(idx[c])*irs*2;
        const float mult = blend[c];
        for(uint32_t k = 0;k < irs*2;++k) coeffs[k] = src[k]*mult + coeffs[k];
    }
    return B200MIX_OK;
}

int b200mix_hrtf_build_decoder(const b200mix_hrtf *h, uint32_t ambi_order, uint32_t voice_ir_size,
    uint32_t *ir_size, float *coeffs, float hf_scale[4], float *splitter_coeff)
{
    // InitHrtfPanning's first-order set-up (alc/panning.cpp:847-1137: AmbiPoints1O, AmbiMatrix1O,
    // AmbiOrderHFGain1O, 700 Hz crossover) handed to DirectHrtfState::build (core/hrtf.cpp:265-366)
    if(!h || !ir_size || !coeffs || !hf_scale || !splitter_coeff || h->fields.empty()) return B200MIX_ERR_INVALID;
    if(ambi_order != 1u) return B200MIX_ERR_UNSUPPORTED;
    constexpr uint32_t kChannels = 4, kHrirLength = B200MIX_HRIR_LENGTH;
    constexpr float Deg180 = std::numbers::pi_v<float>, Deg_90 = Deg180 / 2.0f, Deg_45 = Deg_90 / 2.0f;
    constexpr float Deg135 = Deg_45 * 3.0f, Deg_35 = 6.154797087e-01f;
    const float pts[8][2] = {      // {elevation, azimuth}
        { Deg_35, -Deg_45}, { Deg_35, -Deg135}, { Deg_35, Deg_45}, { Deg_35, Deg135},
        {-Deg_35, -Deg_45}, {-Deg_35, -Deg135}, {-Deg_35, Deg_45}, {-Deg_35, Deg135}};
    static const float mtx[8][4] = {
        {0.125f,  0.125f,  0.125f,  0.125f}, {0.125f,  0.125f,  0.125f, -0.125f},
        {0.125f, -0.125f,  0.125f,  0.125f}, {0.125f, -0.125f,  0.125f, -0.125f},
        {0.125f,  0.125f, -0.125f,  0.125f}, {0.125f,  0.125f, -0.125f, -0.125f},
        {0.125f, -0.125f, -0.125f,  0.125f}, {0.125f, -0.125f, -0.125f, -0.125f}};
    const float orderHF[2] = {2.000000000e+00f, 1.154700538e+00f};

    // the channels' band splitter (BandSplitter::init, core/filters/splitter.cpp:15-26)
    {
        const double xover_norm = double(700.0f) / h->sample_rate;
        const float w = std::numbers::pi_v<float>*2.0f * std::min(float(xover_norm), 0.49f);
        const float cw = std::cos(w);
        *splitter_coeff = cw > std::numeric_limits<float>::epsilon() ? (std::sin(w) - 1.0f) / cw : cw * -0.5f;
    }
    hf_scale[0] = orderHF[0];
    hf_scale[1] = hf_scale[2] = hf_scale[3] = orderHF[1];

    // the HRIR nearest to every virtual speaker (:292-326)
    struct Impulse { const float *hrir; uint32_t ldelay, rdelay; };
    Impulse imp[8];
    uint32_t min_delay = B200MIX_HRTF_HISTORY * 4u, max_delay = 0u;
    const uint32_t evCount = h->fields[0].ev_count;
    for(int k = 0;k < 8;++k)
    {
        const IdxBlend elev0 = EvIndex(evCount, pts[k][0]);
        const uint32_t elev1_idx = std::min(elev0.idx + 1u, evCount - 1u);
        const uint32_t ir0offset = h->elevs[elev0.idx].ir_offset, ir1offset = h->elevs[elev1_idx].ir_offset;
        const IdxBlend az0 = AzIndex(h->elevs[elev0.idx].az_count, pts[k][1]);
        const IdxBlend az1 = AzIndex(h->elevs[elev1_idx].az_count, pts[k][1]);
        const uint32_t idx[4] = {
            ir0offset + az0.idx, ir0offset + ((az0.idx+1u) % h->elevs[elev0.idx].az_count),
            ir1offset + az1.idx, ir1offset + ((az1.idx+1u) % h->elevs[elev1_idx].az_count)};
        const uint32_t irOffset = idx[(elev0.blend >= 0.5f ? 2u : 0u) + (az1.blend >= 0.5f ? 1u : 0u)];
        imp[k] = Impulse{&h->coeffs[size_t(irOffset)*h->ir_size*2u], h->delays[irOffset*2u], h->delays[irOffset*2u + 1u]};
        min_delay = std::min(min_delay, std::min(imp[k].ldelay, imp[k].rdelay));
        max_delay = std::max(max_delay, std::max(imp[k].ldelay, imp[k].rdelay));
    }

    auto delay_round = [](uint32_t d) { return (d + 2u) >> 2; };      // HrirDelayFracHalf / FracBits
    std::vector<double> tmp(size_t(kChannels)*kHrirLength*2u, 0.0);
    max_delay = 0u;
    for(int k = 0;k < 8;++k)
    {
        const uint32_t base_delay = min_delay;                         // perHrirMin is false at first order
        const uint32_t ldelay = delay_round(imp[k].ldelay - base_delay);
        const uint32_t rdelay = delay_round(imp[k].rdelay - base_delay);
        max_delay = std::max(max_delay, std::max(imp[k].ldelay, imp[k].rdelay) - base_delay);
        for(uint32_t c = 0;c < kChannels;++c)
        {
            const double mult = mtx[k][c];
            double *res = &tmp[size_t(c)*kHrirLength*2u];
            for(uint32_t i = 0;i < h->ir_size && ldelay + i < kHrirLength;++i)
                res[(ldelay + i)*2u] = double(imp[k].hrir[i*2u])*mult + res[(ldelay + i)*2u];
            for(uint32_t i = 0;i < h->ir_size && rdelay + i < kHrirLength;++i)
                res[(rdelay + i)*2u + 1u] = double(imp[k].hrir[i*2u + 1u])*mult + res[(rdelay + i)*2u + 1u];
        }
    }
    const uint32_t max_length = std::min(delay_round(max_delay) + (voice_ir_size ? voice_ir_size : h->ir_size), kHrirLength);
    *ir_size = max_length;
    // [channels][max_length][2]
    for(uint32_t c = 0;c < kChannels;++c)
        for(uint32_t i = 0;i < max_length;++i)
        {
            coeffs[(size_t(c)*max_length + i)*2u] = float(tmp[(size_t(c)*kHrirLength + i)*2u]);
            coeffs[(size_t(c)*max_length + i)*2u + 1u] = float(tmp[(size_t(c)*kHrirLength + i)*2u + 1u]);
        }
    return int(kChannels);
}

} // extern "C"
