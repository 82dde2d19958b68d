/* ref_harness.cpp — TEST INFRASTRUCTURE ONLY (never linked or loaded by the product).
 *
 * A thin window into the UNMODIFIED reference (oracle/_ref/libopenal_ref.so,
 * built from /root/reference by oracle/refbuild/Makefile).  Compiled against
 * the reference's private headers where they lie; no reference source is copied.
 *
 * It does two jobs:
 *  1. "reference-side binding": converts the reference's live post-ALU objects
 *     (DeviceBase, Voice, BufferStorage) into the b200mix C-ABI structs of
 *     include/b200mix.h — exactly what a maintainer's seam in
 *     DeviceBase::renderSamples (alc/alu.cpp:2412) would do, see INTEGRATION.md.
 *  2. kernel-level taps: calls Resample_*_C/SSE, and dumps the static-init
 *     coefficient tables, so the oracle restatement (oracle/almix_oracle.c) and
 *     the CUDA tables can be pinned bit-for-bit.
 */
#include "config.h"

/* The only private members we need are plain floats/arrays; open them up for
 * reading instead of patching the reference. */
#include <algorithm>
#include <array>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <span>
#include <variant>
#include <vector>

/* pre-include their dependencies so only the two class bodies see the macros */
#include "opthelpers.h"
#include "core/ambidefs.h"
#include "core/bufferline.h"
#include "core/devformat.h"
#include <cmath>
#include <numbers>
#include "alnumeric.h"
#define class struct
#define private public
#define protected public
#include "core/filters/splitter.h"
#include "core/bformatdec.h"
#include "core/filters/biquad.h"
#undef protected
#undef private
#undef class

#include "AL/al.h"
#include "AL/alc.h"
#include "alc/context.hpp"
#include "alc/device.h"
#include "alc/alu.h"
#include "core/bsinc_tables.h"
#include "core/context.h"
#include "core/cubic_tables.h"
#include "core/device.h"
#include "core/effectslot.h"
#include "core/fpu_ctrl.h"
#include "core/hrtf.h"
#include "core/mixer/defs.h"
#include "core/mixer.h"
#include "core/mixer/hrtfdefs.h"
#include "core/voice.h"
#include "core/mastering.h"
#include "core/bs2b.h"
#include "core/front_stablizer.h"
#include "polyphase_resampler.h"
#include <limits>

#include "../include/b200mix.h"

namespace {
auto dev_of(ALCdevice *d) -> al::Device* { return static_cast<al::Device*>(d); }
auto ctx_of(ALCcontext *c) -> al::Context* { return static_cast<al::Context*>(c); }
} // namespace

extern "C" {

/* Extra per-voice mixer state the ABI does not carry (it is device-resident in
 * the product); used by tests to compare state evolution, and to seed it. */
struct refh_voice_state {
    int32_t  play_state;      /* Voice::State */
    uint32_t is_fading;
    int32_t  position;
    uint32_t position_frac;
    uint32_t buffer_frames;   /* VoiceBufferItem::mSampleLen */
    uint32_t buffer_type;     /* b200mix_sample_type or 0xffffffff */
    uint32_t buffer_channels; /* mFrameStep */
    uint32_t bsinc_m, bsinc_l;
    float    bsinc_sf;
    const void *buffer_data;  /* host pointer into the reference's BufferStorage */
    float    prev_samples[B200MIX_RESAMPLER_PADDING];
    float    hrtf_history[B200MIX_HRTF_HISTORY];
    float    old_coeffs[B200MIX_HRIR_LENGTH][2];
    uint32_t old_delay[2];
    float    old_gain;
    float    cur_dry_gains[B200MIX_MAX_DRY_CHANNELS];
    float    cur_send_gains[B200MIX_MAX_SENDS][B200MIX_MAX_WET_CHANNELS];
};

int refh_device_desc(ALCdevice *adev, b200mix_device_desc *out)
{
    auto *dev = dev_of(adev);
    std::memset(out, 0, sizeof(*out));
    out->struct_size = sizeof(*out);
    out->cuda_device = -1;
    out->sample_rate = dev->mSampleRate;
    out->dry_channels = static_cast<uint32_t>(dev->Dry.Buffer.size());
    out->real_channels = static_cast<uint32_t>(dev->RealOut.Buffer.size());
    out->num_sends = dev->NumAuxSends;
    out->ir_size = dev->mIrSize;
    out->wet_channels = 0; /* filled by refh_wet_channels once a slot exists */
    if(std::holds_alternative<AmbiDecPostProcess>(dev->mPostProcess)
        || std::holds_alternative<StablizerPostProcess>(dev->mPostProcess))
        out->post_process = B200MIX_POST_AMBIDEC;   /* + refh_front_stabilizer for the latter */
    else if(std::holds_alternative<HrtfPostProcess>(dev->mPostProcess))
        out->post_process = B200MIX_POST_HRTF;
    else if(std::holds_alternative<UhjPostProcess>(dev->mPostProcess))
        out->post_process = B200MIX_POST_UHJ;
    else if(std::holds_alternative<TsmePostProcess>(dev->mPostProcess))
        out->post_process = B200MIX_POST_TSME;
    else if(std::holds_alternative<std::monostate>(dev->mPostProcess))
        out->post_process = B200MIX_POST_NONE;
    else
        return -1;
    out->real_left = dev->RealOut.ChannelIndex[FrontLeft].c_val;
    out->real_right = dev->RealOut.ChannelIndex[FrontRight].c_val;
    /* RealOut aliases Dry when no decode is needed. */
    if(out->post_process == B200MIX_POST_NONE
        && dev->RealOut.Buffer.data() != dev->Dry.Buffer.data())
        return -2;
    return 0;
}

/* Returns the channel count; *ir_size receives DirectHrtfState::mIrSize (which can
 * exceed the device's per-voice mIrSize); arrays sized [channels][*ir_size][2],
 * [channels], [channels].  Call with NULL arrays first to learn the sizes. */
int refh_hrtf_decoder(ALCdevice *adev, uint32_t *ir_size, float *coeffs, float *hf_scale,
    float *splitter_coeff)
{
    auto *dev = dev_of(adev);
    auto *proc = std::get_if<HrtfPostProcess>(&dev->mPostProcess);
    if(!proc) return -1;
    auto &st = *proc->mHrtfState;
    const auto ir = st.mIrSize;
    *ir_size = ir;
    auto c = 0u;
    for(auto &chan : st.mChannels)
    {
        if(coeffs)
            for(auto j = 0u;j < ir;++j)
            {
                coeffs[(c*ir + j)*2 + 0] = chan.mCoeffs[j][0];
                coeffs[(c*ir + j)*2 + 1] = chan.mCoeffs[j][1];
            }
        if(hf_scale) hf_scale[c] = chan.mHfScale;
        if(splitter_coeff) splitter_coeff[c] = chan.mSplitter.mCoeff;
        ++c;
    }
    return static_cast<int>(c);
}

/* Returns in_channels; gains are [in][real_channels]. *dual = 1 if dual-band. */
int refh_ambi_decoder(ALCdevice *adev, float *gains_hf, float *gains_lf, float *xover_coeff,
    int *dual)
{
    auto *dev = dev_of(adev);
    BFormatDec *decp{nullptr};
    if(auto *proc = std::get_if<AmbiDecPostProcess>(&dev->mPostProcess)) decp = proc->mAmbiDecoder.get();
    else if(auto *sproc = std::get_if<StablizerPostProcess>(&dev->mPostProcess)) decp = sproc->mAmbiDecoder.get();
    if(!decp) return -1;
    auto &dec = *decp;
    const auto outs = dev->RealOut.Buffer.size();
    if(auto *sb = std::get_if<BFormatDec::SBandDecoderVector>(&dec.mChannelDec))
    {
        *dual = 0;
        for(auto i = 0_uz;i < sb->size();++i)
            for(auto o = 0_uz;o < outs;++o)
                gains_hf[i*outs + o] = (*sb)[i].mGains[o];
        return static_cast<int>(sb->size());
    }
    auto &db = std::get<BFormatDec::DBandDecoderVector>(dec.mChannelDec);
    *dual = 1;
    for(auto i = 0_uz;i < db.size();++i)
        for(auto o = 0_uz;o < outs;++o)
        {
            gains_hf[i*outs + o] = db[i].mGains[BFormatDec::sHFBand][o];
            gains_lf[i*outs + o] = db[i].mGains[BFormatDec::sLFBand][o];
        }
    *xover_coeff = db.empty() ? 0.0f : db[0].mXOver.mCoeff;
    return static_cast<int>(db.size());
}

/* DeviceBase::DitherDepth as the device open chose it (alc/alc.cpp:1690-1716). */
float refh_dither_depth(ALCdevice *adev) { return dev_of(adev)->DitherDepth; }

/* The device limiter as UpdateDeviceParams set it up: returns 0 when device->Limiter is null,
 * else fills the Compressor::Params of CreateDeviceLimiter (alc/alc.cpp:1079-1091) with the
 * threshold of alc/alc.cpp:1750-1770 (the same float expressions, evaluated here by the same
 * libm) — what a maintainer's binding hands to b200mix_set_limiter. */
int refh_limiter_desc(ALCdevice *adev, b200mix_limiter_desc *out)
{
    auto *device = dev_of(adev);
    if(!device->Limiter) return 0;
    auto thrshld = 1.0f;
    switch(device->FmtType)
    {
    case DevFmtByte: case DevFmtUByte: thrshld = 127.0f / 128.0f; break;
    case DevFmtShort: case DevFmtUShort: thrshld = 32767.0f / 32768.0f; break;
    case DevFmtInt: case DevFmtUInt: case DevFmtFloat: break;
    }
    if(device->DitherDepth > 0.0f)
        thrshld -= 1.0f / device->DitherDepth;
    *out = b200mix_limiter_desc{};
    out->struct_size = sizeof(*out);
    out->auto_flags = B200MIX_LIM_AUTO_KNEE | B200MIX_LIM_AUTO_ATTACK | B200MIX_LIM_AUTO_RELEASE
        | B200MIX_LIM_AUTO_POSTGAIN | B200MIX_LIM_AUTO_DECLIP;
    out->look_ahead_time = 0.001f; out->hold_time = 0.002f;
    out->pre_gain_db = 0.0f; out->post_gain_db = 0.0f;
    out->threshold_db = std::log10(thrshld) * 20.0f;
    out->ratio = std::numeric_limits<float>::infinity();
    out->knee_db = 0.0f; out->attack_time = 0.02f; out->release_time = 0.2f;
    return static_cast<int>(device->Limiter->getLookAhead()) + 1;
}

/* DeviceBase::ChannelDelays (core/device.h:85-100) as InitDistanceComp built it: per RealOut
 * channel the delay line length and gain.  Returns the channel count, 0 without distance comp. */
int refh_distance_comp(ALCdevice *adev, uint32_t *delays, float *gains)
{
    auto *device = dev_of(adev);
    if(!device->ChannelDelays) return 0;
    const auto n = device->RealOut.Buffer.size();
    for(size_t c{0};c < n;++c)
    {
        const auto &cd = device->ChannelDelays->mChannels[c];
        delays[c] = static_cast<uint32_t>(cd.Buffer.size());
        gains[c] = cd.Gain;
    }
    return static_cast<int>(n);
}

/* Kernel-level tap (SURVEY 8c ii): the reference's Bs2b::bs2b_processor, set up by set_params
 * and run over [left|right] in place; state (4 floats: history[0].lo/.hi, history[1].lo/.hi)
 * goes in and comes back so ragged runs can be chained; coef receives
 * {a0_lo, b1_lo, a0_hi, a1_hi, b1_hi}.  (The loopback device never selects Bs2bPostProcess,
 * alc/panning.cpp:1421, so the filter is pinned here.) */
void refh_bs2b_cross_feed(int level, int srate, float *left, float *right, int n, float *state, float *coef)
{
    Bs2b::bs2b_processor p{};
    p.set_params(level, srate);
    p.history[0].lo = state[0]; p.history[0].hi = state[1];
    p.history[1].lo = state[2]; p.history[1].hi = state[3];
    p.cross_feed(std::span{left, size_t(n)}, std::span{right, size_t(n)});
    state[0] = p.history[0].lo; state[1] = p.history[0].hi;
    state[2] = p.history[1].lo; state[3] = p.history[1].hi;
    coef[0] = p.a0_lo; coef[1] = p.b1_lo; coef[2] = p.a0_hi; coef[3] = p.a1_hi; coef[4] = p.b1_hi;
}

/* StablizerPostProcess (core/device.h:176-179): returns the FrontCenter RealOut index and the
 * band-splitter coefficient of FrontStablizer::MidFilter, or -1 when the device has none. */
int refh_front_stabilizer(ALCdevice *adev, float *splitter_coeff)
{
    auto *dev = dev_of(adev);
    auto *proc = std::get_if<StablizerPostProcess>(&dev->mPostProcess);
    if(!proc) return -1;
    *splitter_coeff = proc->mStablizer->MidFilter.mCoeff;
    return static_cast<int>(dev->RealOut.ChannelIndex[FrontCenter].c_val);
}

/* Kernel-level taps for the panning helpers: CalcDirectionCoeffs (core/mixer.h:68-73) and
 * ComputePanGains (core/mixer.cpp:93-102) on the device's Dry mix; refh_dry_ambi_map returns the
 * Dry.AmbiMap entries ({Scale, Index}, core/device.h:109-121) a binding passes to
 * b200mix_pan_gains. */
void refh_calc_direction_coeffs(const float *dir, float spread, float *out25)
{
    const auto c = CalcDirectionCoeffs(std::span<const float,3>{dir, 3}, spread);
    std::copy(c.begin(), c.end(), out25);
}

int refh_dry_ambi_map(ALCdevice *adev, float *scale, uint32_t *index)
{
    auto *dev = dev_of(adev);
    const auto n = dev->Dry.Buffer.size();
    for(size_t c{0};c < n;++c)
    {
        scale[c] = dev->Dry.AmbiMap[c].Scale;
        index[c] = static_cast<uint32_t>(dev->Dry.AmbiMap[c].Index);
    }
    return static_cast<int>(n);
}

void refh_dry_pan_gains(ALCdevice *adev, const float *coeffs25, float ingain, float *gains)
{
    auto *dev = dev_of(adev);
    std::array<float,MaxAmbiChannels> g{};
    ComputePanGains(&dev->Dry, std::span<const float,MaxAmbiChannels>{coeffs25, MaxAmbiChannels}, ingain, g);
    std::copy(g.begin(), g.end(), gains);
}

/* What ReverbState::deviceUpdate reads from the device (alc/effects/reverb.cpp:834-850). */
void refh_device_ambi(ALCdevice *adev, uint32_t *order, uint32_t *is2d, float *xover_freq)
{
    auto *dev = dev_of(adev);
    *order = dev->mAmbiOrder; *is2d = dev->m2DMixing ? 1u : 0u; *xover_freq = dev->mXOverFreq;
}

/* Kernel-level tap: PPhaseResampler (common/polyphase_resampler.h) as ConvolutionState::deviceUpdate
 * uses it — float samples widened to double, resampled, narrowed back with one rounding. */
void refh_pphase_resample(unsigned src_rate, unsigned dst_rate, const float *in, unsigned n_in,
    float *out, unsigned n_out)
{
    auto rs = PPhaseResampler{};
    rs.init(src_rate, dst_rate);
    auto din = std::vector<double>(in, in + n_in);
    auto dout = std::vector<double>(n_out);
    rs.process(din, dout);
    std::ranges::transform(dout, out, [](double d) { return static_cast<float>(d); });
}

/* Taps for the source parameter helper: ContextParams (core/context.h:67-84), the VoiceProps of
 * voice idx with the EffectSlotBase values CalcAttnVoiceParams reads through its send slots
 * (alc/alu.cpp:1712-1742,1925-1961), the device's render mode, and an active slot's Wet.AmbiMap. */
void refh_listener_params(ALCcontext *actx, b200mix_listener_params *out)
{
    auto &p = ctx_of(actx)->mParams;
    *out = b200mix_listener_params{};
    out->struct_size = sizeof(*out);
    for(size_t i{0};i < 3;++i) { out->position[i] = p.Position[i]; out->velocity[i] = p.Velocity[i]; }
    for(size_t r{0};r < 4;++r) for(size_t c{0};c < 4;++c) out->matrix[r*4 + c] = p.Matrix[r][c];
    out->gain = p.Gain; out->meters_per_unit = p.MetersPerUnit;
    out->air_absorption_gain_hf = p.AirAbsorptionGainHF; out->doppler_factor = p.DopplerFactor;
    out->speed_of_sound = p.SpeedOfSound;
    out->source_distance_model = p.SourceDistanceModel ? 1u : 0u;
    out->distance_model = static_cast<uint32_t>(p.mDistanceModel);
}

int refh_source_props(ALCcontext *actx, int idx, b200mix_source_props *out, uint32_t *buffer_rate)
{
    auto voices = ctx_of(actx)->getVoicesSpan();
    if(idx < 0 || size_t(idx) >= voices.size()) return -1;
    auto *voice = voices[size_t(idx)];
    auto const &P = voice->mProps;
    *out = b200mix_source_props{};
    out->struct_size = sizeof(*out);
    out->pitch = P.Pitch; out->gain = P.Gain; out->outer_gain = P.OuterGain;
    out->min_gain = P.MinGain; out->max_gain = P.MaxGain;
    out->inner_angle = P.InnerAngle; out->outer_angle = P.OuterAngle;
    out->ref_distance = P.RefDistance; out->max_distance = P.MaxDistance; out->rolloff_factor = P.RolloffFactor;
    for(size_t i{0};i < 3;++i)
    {
        out->position[i] = P.Position[i]; out->velocity[i] = P.Velocity[i]; out->direction[i] = P.Direction[i];
        out->orient_at[i] = P.OrientAt[i]; out->orient_up[i] = P.OrientUp[i];
    }
    out->head_relative = P.HeadRelative ? 1u : 0u;
    out->distance_model = static_cast<uint32_t>(P.mDistanceModel);
    out->dry_gain_hf_auto = P.DryGainHFAuto; out->wet_gain_auto = P.WetGainAuto;
    out->wet_gain_hf_auto = P.WetGainHFAuto; out->outer_gain_hf = P.OuterGainHF;
    out->air_absorption_factor = P.AirAbsorptionFactor; out->room_rolloff_factor = P.RoomRolloffFactor;
    out->doppler_factor = P.DopplerFactor; out->radius = P.Radius;
    out->direct.gain = P.Direct.Gain; out->direct.gain_hf = P.Direct.GainHF;
    out->direct.hf_reference = P.Direct.HFReference; out->direct.gain_lf = P.Direct.GainLF;
    out->direct.lf_reference = P.Direct.LFReference;
    for(size_t i{0};i < MaxSendCount;++i)
    {
        auto &S = out->sends[i];
        S.gain = P.Send[i].Gain; S.gain_hf = P.Send[i].GainHF; S.hf_reference = P.Send[i].HFReference;
        S.gain_lf = P.Send[i].GainLF; S.lf_reference = P.Send[i].LFReference;
        auto *slot = P.Send[i].Slot;
        S.active = (slot && slot->EffectType != EffectSlotType::None) ? 1u : 0u;
        if(S.active)
        {
            S.slot_room_rolloff = slot->RoomRolloff; S.slot_decay_time = slot->DecayTime;
            S.slot_air_absorption_gain_hf = slot->AirAbsorptionGainHF;
        }
    }
    *buffer_rate = voice->mFrequency;
    return 0;
}

int refh_device_render_mode(ALCdevice *adev) { return static_cast<int>(dev_of(adev)->mRenderMode); }

int refh_slot_ambi_map(ALCcontext *actx, int idx, float *scale, uint32_t *index)
{
    auto *arr = ctx_of(actx)->mActiveAuxSlots.load(std::memory_order_acquire);
    if(!arr || idx < 0 || size_t(idx) >= (arr->size()>>1)) return -1;
    auto &wet = (*arr)[size_t(idx)]->Wet;
    for(size_t c{0};c < wet.Buffer.size();++c)
    {
        scale[c] = wet.AmbiMap[c].Scale;
        index[c] = static_cast<uint32_t>(wet.AmbiMap[c].Index);
    }
    return static_cast<int>(wet.Buffer.size());
}

/* Which Voice::mChans[] entry refh_snapshot_voices reads (multi-channel sources: one
 * mixing channel per buffer channel, core/voice.h:236-257).  Default 0. */
static size_t g_snap_channel = 0;
void refh_set_snapshot_channel(int c) { g_snap_channel = c < 0 ? 0u : size_t(c); }

int refh_voice_count(ALCcontext *actx)
{ return static_cast<int>(ctx_of(actx)->getVoicesSpan().size()); }

/* Snapshot of every voice slot [0, count): ABI params + side arrays + state.
 * Any output pointer may be NULL.  dry_gains is [n][dry_channels], hrtf_coeffs
 * [n][ir_size][2], send_gains [n][num_sends][wet_channels]. */
int refh_snapshot_voices(ALCcontext *actx, b200mix_voice_params *params, float *hrtf_coeffs,
    float *dry_gains, float *send_gains, uint32_t wet_channels, refh_voice_state *state)
{
    auto *ctx = ctx_of(actx);
    auto *dev = static_cast<DeviceBase*>(ctx->mDevice);
    const auto voices = ctx->getVoicesSpan();
    const auto ir = dev->mIrSize;
    const auto cd = dev->Dry.Buffer.size();
    const auto ns = dev->NumAuxSends;
    auto slots = std::span<EffectSlotBase*>{};
    if(auto *arr = ctx->mActiveAuxSlots.load(std::memory_order_acquire))
    {
        auto all = std::span{*arr};
        slots = all.first(all.size()>>1);
    }

    auto n = 0u;
    for(auto *voice : voices)
    {
        auto &ch = voice->mChans[std::min(g_snap_channel, voice->mChans.size()-1)];
        const auto pstate = voice->mPlayState.load();
        auto *item = voice->mCurrentBuffer.load();
        auto *loop = voice->mLoopBuffer.load();
        if(params)
        {
            auto &p = params[n];
            std::memset(&p, 0, sizeof(p));
            p.voice = n;
            if(pstate == Voice::Playing) p.flags |= B200MIX_VF_PLAYING;
            else if(pstate == Voice::Stopping) p.flags |= B200MIX_VF_STOPPING;
            else p.flags |= B200MIX_VF_STOPPED;
            if(voice->mFlags.test(VoiceFlag::IsStatic)) p.flags |= B200MIX_VF_STATIC;
            if(loop) p.flags |= B200MIX_VF_LOOPING;
            if(voice->mFlags.test(VoiceFlag::HasHrtf)) p.flags |= B200MIX_VF_HRTF;
            p.buffer = 0;
            p.resampler = static_cast<uint32_t>(voice->mProps.mResampler);
            p.position = voice->mPosition.load();
            p.position_frac = voice->mPositionFrac.load();
            p.loop_start = item ? item->mLoopStart : 0u;
            p.loop_end = item ? item->mLoopEnd : 0u;
            p.step = voice->mStep;
            p.hrtf_delay[0] = ch.mDryParams.Hrtf.Target.Delay[0];
            p.hrtf_delay[1] = ch.mDryParams.Hrtf.Target.Delay[1];
            p.hrtf_gain = ch.mDryParams.Hrtf.Target.Gain;
            for(auto s = 0u;s < B200MIX_MAX_SENDS;++s)
            {
                p.send_slot[s] = B200MIX_NO_SLOT;
                if(s >= ns || voice->mSend[s].Buffer.empty()) continue;
                for(auto k = 0_uz;k < slots.size();++k)
                    if(slots[k]->Wet.Buffer.data() == voice->mSend[s].Buffer.data())
                        p.send_slot[s] = static_cast<uint32_t>(k);
            }
        }
        if(hrtf_coeffs)
            for(auto j = 0u;j < ir;++j)
            {
                hrtf_coeffs[(n*ir + j)*2 + 0] = ch.mDryParams.Hrtf.Target.Coeffs[j][0];
                hrtf_coeffs[(n*ir + j)*2 + 1] = ch.mDryParams.Hrtf.Target.Coeffs[j][1];
            }
        if(dry_gains)
            for(auto c = 0_uz;c < cd;++c)
                dry_gains[n*cd + c] = ch.mDryParams.Gains.Target[c];
        if(send_gains)
            for(auto s = 0u;s < ns;++s)
                for(auto c = 0u;c < wet_channels;++c)
                    send_gains[(n*ns + s)*wet_channels + c] = ch.mWetParams[s].Gains.Target[c];
        if(state)
        {
            auto &st = state[n];
            std::memset(&st, 0, sizeof(st));
            st.play_state = static_cast<int32_t>(pstate);
            st.is_fading = voice->mFlags.test(VoiceFlag::IsFading);
            st.position = voice->mPosition.load();
            st.position_frac = voice->mPositionFrac.load();
            st.buffer_type = 0xffffffffu;
            if(item)
            {
                st.buffer_frames = item->mSampleLen;
                st.buffer_channels = voice->mFrameStep;
                std::visit([&st]<typename T>(std::span<T> const &spl)
                {
                    st.buffer_data = spl.data();
                    if constexpr(std::is_same_v<T,u8>) st.buffer_type = B200MIX_FMT_U8;
                    else if constexpr(std::is_same_v<T,i16>) st.buffer_type = B200MIX_FMT_I16;
                    else if constexpr(std::is_same_v<T,i32>) st.buffer_type = B200MIX_FMT_I32;
                    else if constexpr(std::is_same_v<T,f32>) st.buffer_type = B200MIX_FMT_F32;
                    else if constexpr(std::is_same_v<T,f64>) st.buffer_type = B200MIX_FMT_F64;
                    else if constexpr(std::is_same_v<T,MulawSample>) st.buffer_type = B200MIX_FMT_MULAW;
                    else if constexpr(std::is_same_v<T,AlawSample>) st.buffer_type = B200MIX_FMT_ALAW;
                }, item->mSamples);
            }
            if(auto *bs = std::get_if<BsincState>(&voice->mResampleState))
            {
                st.bsinc_m = bs->m.c_val;
                st.bsinc_l = bs->l.c_val;
                st.bsinc_sf = bs->sf;
            }
            std::copy_n(voice->mPrevSamples[0].begin(), B200MIX_RESAMPLER_PADDING, st.prev_samples);
            std::copy_n(ch.mDryParams.Hrtf.History.begin(), B200MIX_HRTF_HISTORY, st.hrtf_history);
            for(auto j = 0u;j < B200MIX_HRIR_LENGTH;++j)
            {
                st.old_coeffs[j][0] = ch.mDryParams.Hrtf.Old.Coeffs[j][0];
                st.old_coeffs[j][1] = ch.mDryParams.Hrtf.Old.Coeffs[j][1];
            }
            st.old_delay[0] = ch.mDryParams.Hrtf.Old.Delay[0];
            st.old_delay[1] = ch.mDryParams.Hrtf.Old.Delay[1];
            st.old_gain = ch.mDryParams.Hrtf.Old.Gain;
            std::copy_n(ch.mDryParams.Gains.Current.begin(), B200MIX_MAX_DRY_CHANNELS,
                st.cur_dry_gains);
            for(auto s = 0u;s < B200MIX_MAX_SENDS;++s)
                std::copy_n(ch.mWetParams[s].Gains.Current.begin(), B200MIX_MAX_WET_CHANNELS,
                    st.cur_send_gains[s]);
        }
        ++n;
    }
    return static_cast<int>(n);
}

/* Active aux slots of the context (ContextBase::mActiveAuxSlots, core/context.h:142). */
int refh_slot_count(ALCcontext *actx)
{
    auto *arr = ctx_of(actx)->mActiveAuxSlots.load(std::memory_order_acquire);
    return arr ? static_cast<int>(arr->size()>>1) : 0;
}

/* Wet.Buffer.size() of active slot `idx` (aluInitEffectPanning, alc/panning.cpp:1441). */
int refh_slot_wet_channels(ALCcontext *actx, int idx)
{
    auto *arr = ctx_of(actx)->mActiveAuxSlots.load(std::memory_order_acquire);
    if(!arr || idx < 0 || size_t(idx) >= (arr->size()>>1)) return -1;
    return static_cast<int>((*arr)[size_t(idx)]->Wet.Buffer.size());
}

/* Wet mix of active slot idx after the last update: [wet_channels][1024]. */
int refh_get_wet(ALCcontext *actx, int idx, float *out)
{
    auto *arr = ctx_of(actx)->mActiveAuxSlots.load(std::memory_order_acquire);
    if(!arr || idx < 0 || size_t(idx) >= (arr->size()>>1)) return -1;
    auto &wet = (*arr)[size_t(idx)]->Wet.Buffer;
    for(auto c = 0_uz;c < wet.size();++c)
        std::copy_n(wet[c].begin(), BufferLineSize, out + c*BufferLineSize);
    return static_cast<int>(wet.size());
}

/* Per-voice send filter activity (Voice::mSend[s].FilterActive) and direct filter. */
int refh_voice_filters_active(ALCcontext *actx, int voice, int *direct, int *sends)
{
    auto voices = ctx_of(actx)->getVoicesSpan();
    if(voice < 0 || size_t(voice) >= voices.size()) return -1;
    *direct = voices[size_t(voice)]->mDirect.FilterActive;
    for(auto s = 0u;s < MaxSendCount;++s) sends[s] = voices[size_t(voice)]->mSend[s].FilterActive;
    return 0;
}

/* Direct/send filter targets of one voice as CalcVoiceParams left them
 * (alc/alu.cpp:1619-1656): per path p (0 = direct, 1+s = send s) the FilterActive flag,
 * LowPass/HighPass mTargetCoeffs {b0,b1,b2,a1,a2} and mCounter.
 * coeffs is [7][2][5], active [7], counters [7][2]. */
int refh_voice_filters(ALCcontext *actx, int voice, float *coeffs, int *active, int *counters)
{
    auto voices = ctx_of(actx)->getVoicesSpan();
    if(voice < 0 || size_t(voice) >= voices.size()) return -1;
    auto *v = voices[size_t(voice)];
    auto &ch = v->mChans[0];
    const auto put = [&](size_t p, BiquadInterpFilter &lp, BiquadInterpFilter &hp, bool act)
    {
        const auto one = [](float *o, const BiquadInterpFilter &f)
        {
            o[0] = f.mTargetCoeffs.mB0; o[1] = f.mTargetCoeffs.mB1; o[2] = f.mTargetCoeffs.mB2;
            o[3] = f.mTargetCoeffs.mA1; o[4] = f.mTargetCoeffs.mA2;
        };
        one(coeffs + (p*2 + 0)*5, lp);
        one(coeffs + (p*2 + 1)*5, hp);
        active[p] = act;
        counters[p*2 + 0] = lp.mCounter;
        counters[p*2 + 1] = hp.mCounter;
    };
    put(0, ch.mDryParams.LowPass, ch.mDryParams.HighPass, v->mDirect.FilterActive);
    for(auto s = 0_uz;s < MaxSendCount;++s)
        put(1 + s, ch.mWetParams[s].LowPass, ch.mWetParams[s].HighPass, v->mSend[s].FilterActive);
    return 0;
}

/* BiquadFilter::SetParams through setParamsFromSlope (core/filters/biquad.h:92-97,
 * biquad.cpp:48-129): type 0 = HighShelf, 1 = LowShelf, ... as enum BiquadType. */
void refh_biquad_coeffs(int type, float f0norm, float gain, float slope, float *out)
{
    auto f = BiquadFilter{};
    f.setParamsFromSlope(static_cast<BiquadType>(type), f0norm, gain, slope);
    out[0] = f.mCoeffs.mB0; out[1] = f.mCoeffs.mB1; out[2] = f.mCoeffs.mB2;
    out[3] = f.mCoeffs.mA1; out[4] = f.mCoeffs.mA2;
}

/* What ConvolutionState::update computes for a MONO impulse response
 * (alc/effects/convolution.cpp:541-600, MonoMap -> front centre): the pan gains of one
 * output line into the Dry mix, scaled by the slot gain.  out has dry_channels entries. */
void refh_mono_line_gains(ALCdevice *adev, float slot_gain, float *out)
{
    auto *dev = dev_of(adev);
    const auto pos = std::array{0.0f, 0.0f, -1.0f};
    const auto coeffs = CalcDirectionCoeffs(pos, 0.0f);
    auto gains = std::array<float, MaxAmbiChannels>{};
    ComputePanGains(&dev->Dry, coeffs, slot_gain, gains);
    for(auto c = 0_uz;c < dev->Dry.Buffer.size() && c < MaxAmbiChannels;++c)
        out[c] = gains[c];
}

/* HrtfAccumData as [1024+128][2]. */
/* As refh_mono_line_gains, but panned into active slot `target_idx`'s Wet mix (a slot whose
 * EffectSlotBase::Target is that slot, alc/alu.cpp:626-633).  out has wet_channels entries. */
int refh_mono_line_gains_slot(ALCcontext *actx, int target_idx, float slot_gain, float *out)
{
    auto *arr = ctx_of(actx)->mActiveAuxSlots.load(std::memory_order_acquire);
    if(!arr || target_idx < 0 || size_t(target_idx) >= (arr->size()>>1)) return -1;
    auto *tslot = (*arr)[size_t(target_idx)];
    const auto pos = std::array{0.0f, 0.0f, -1.0f};
    const auto coeffs = CalcDirectionCoeffs(pos, 0.0f);
    auto gains = std::array<float, MaxAmbiChannels>{};
    ComputePanGains(&tslot->Wet, coeffs, slot_gain, gains);
    for(auto c = 0_uz;c < tslot->Wet.Buffer.size() && c < MaxAmbiChannels;++c)
        out[c] = gains[c];
    return int(tslot->Wet.Buffer.size());
}

void refh_get_hrtf_accum(ALCdevice *adev, float *out)
{
    auto *dev = dev_of(adev);
    for(auto i = 0_uz;i < dev->HrtfAccumData.size();++i)
    {
        out[i*2 + 0] = dev->HrtfAccumData[i][0];
        out[i*2 + 1] = dev->HrtfAccumData[i][1];
    }
}

/* Dry mix of the last update, [dry_channels][1024]. */
void refh_get_dry(ALCdevice *adev, float *out)
{
    auto *dev = dev_of(adev);
    for(auto c = 0_uz;c < dev->Dry.Buffer.size();++c)
        std::copy_n(dev->Dry.Buffer[c].begin(), BufferLineSize, out + c*BufferLineSize);
}

/* ---- kernel-level taps --------------------------------------------------- */

/* Runs the reference resampler exactly as Voice::mix would call it:
 * src is mResampleData (position 0 at index MaxResamplerEdge=24).
 * simd=0 forces the *_C kernels, simd=1 uses PrepareResampler's CPU selection. */
int refh_resample(uint32_t resampler, int simd, uint32_t increment, uint32_t frac,
    const float *src, uint32_t src_len, float *dst, uint32_t dst_len)
{
    auto mixer_mode = FPUCtl{};
    auto state = InterpState{};
    auto func = PrepareResampler(static_cast<Resampler>(resampler), increment, &state);
    if(!simd)
    {
        switch(static_cast<Resampler>(resampler))
        {
        case Resampler::Point: func = Resample_Point_C; break;
        case Resampler::Linear: func = Resample_Linear_C; break;
        case Resampler::Spline: case Resampler::Gaussian: func = Resample_Cubic_C; break;
        case Resampler::BSinc12: case Resampler::BSinc24: case Resampler::BSinc48:
            func = (increment > MixerFracOne) ? Resample_BSinc_C : Resample_FastBSinc_C; break;
        default: func = Resample_FastBSinc_C; break;
        }
    }
    func(&state, std::span{src, src_len}, frac, increment, std::span{dst, dst_len});
    return 0;
}

/* BsincPrepare's result for one increment (alc/alu.cpp:140-165): sf, m, l and the
 * offset of the scale's sub-table.  The tables themselves are hidden symbols
 * (DECL_HIDDEN), so they are reached through PrepareResampler's state. */
namespace {
auto bsinc_state(uint32_t resampler, uint32_t increment) -> BsincState
{
    auto state = InterpState{};
    std::ignore = PrepareResampler(static_cast<Resampler>(resampler), increment, &state);
    return std::get<BsincState>(state);
}
auto bsinc_base(uint32_t resampler) -> const float*
{
    /* scale index 0 (offset 0) is selected by the largest down-sampling ratio */
    return bsinc_state(resampler, MaxPitch<<MixerFracBits).filter.data();
}
} // namespace

int refh_bsinc_state(uint32_t resampler, uint32_t increment, float *sf, uint32_t *m, uint32_t *l,
    uint32_t *offset)
{
    const auto st = bsinc_state(resampler, increment);
    *sf = st.sf; *m = st.m.c_val; *l = st.l.c_val;
    *offset = static_cast<uint32_t>(st.filter.data() - bsinc_base(resampler));
    return 0;
}

/* resampler = BSinc12/24/48 enum value.  Returns the float count of the table
 * (offset of the last scale + its size). */
int64_t refh_bsinc_table(uint32_t resampler, float *out, size_t max_floats)
{
    const auto *base = bsinc_base(resampler);
    const auto last = bsinc_state(resampler, MixerFracOne); /* si = 15 */
    const auto total = static_cast<size_t>(last.filter.data() - base)
        + size_t{last.m.c_val}*4u*BSincPhaseCount;
    if(out) std::copy_n(base, std::min(max_floats, total), out);
    return static_cast<int64_t>(total);
}

/* which: 0 = spline, 1 = gaussian.  out is [32][8] (coeffs[4], deltas[4]). */
void refh_cubic_table(int which, float *out)
{
    auto state = InterpState{};
    std::ignore = PrepareResampler(which ? Resampler::Gaussian : Resampler::Spline, MixerFracOne,
        &state);
    const auto t = std::get<CubicState>(state).filter;
    for(auto pi = 0u;pi < CubicPhaseCount;++pi)
        for(auto k = 0u;k < 4;++k)
        {
            out[pi*8 + k] = t[pi].mCoeffs[k];
            out[pi*8 + 4 + k] = t[pi].mDeltas[k];
        }
}

/* gCubicTable (used by the reverb's modulated taps): 513 floats. */
void refh_cubic_filter(float *out)
{ std::copy(gCubicTable.mFilter.begin(), gCubicTable.mFilter.end(), out); }

} // extern "C"
