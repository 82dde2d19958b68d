/* b200mix_seam.cpp — the reference-side binding of libb200mix.so: compiled INTO a patched
 * libopenal (integration/alu_seam.patch adds the three call sites in alc/alu.cpp) against the
 * reference's own headers.  Nothing here mixes: it converts the reference's live post-ALU
 * objects (DeviceBase, Voice, VoiceBufferItem) into the C ABI of include/b200mix.h, calls
 * b200mix_render, and writes the playback cursor back.
 *
 *   ALSOFT_B200MIX=1            enables the seam (else the library behaves as stock OpenAL Soft)
 *   ALSOFT_B200MIX_LIB=<path>   libb200mix.so (default: "libb200mix.so" on the loader path)
 *
 * Scope of this binding (v2): every output the reference renders to — HRTF, ambisonic decodes
 * (mono … 7.1, with the front stabilizer or BS2B), UHJ / TSME encoded stereo, B-Format —, static mono sources of any
 * PCM sample type, any resampler, moving sources (targets are re-sent when the ALU changed
 * them), source start / stop / loop / end of buffer, auxiliary sends into effect slots —
 * EAX / standard reverb and the EFX effects of b200mix_slot_efx (echo, ring modulator,
 * equalizer, compressor, dedicated, distortion, chorus / flanger, autowah, vocal morpher, frequency shifter, pitch shifter), slot gain, slot
 * targets (AL_SOFT_effect_target), property changes while playing, direct and send filters
 * (AL_DIRECT_FILTER / AL_AUXILIARY_SEND_FILTER low-, high- and band-pass), streaming sources
 * (alSourceQueueBuffers: queue advance, looping queues, buffer-completed events),
 * multi-channel sources (stereo … 7.1 buffers, first-order B-Format on first-order devices: one
 * device voice per mixing channel) and
 * convolution slots with mono … 7.1 impulse responses (any PCM type and rate).
 * Ambisonic / UHJ sources and impulse responses, NFC, direct channels and callback buffers are
 * not wired up here: the seam disconnects the device with
 * a message rather than mixing them wrong.
 */
#include "config.h"

#include <algorithm>
#include <array>
#include <atomic>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <span>
#include <unordered_map>
#include <variant>
#include <vector>

#include <dlfcn.h>

/* BandSplitter::mCoeff, BiquadInterpFilter::mTargetCoeffs and BFormatDec's matrices are private: a maintainer would add two
 * accessors; the out-of-tree build opens them up instead of touching more reference files. */
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
#include "core/filters/biquad.h"
#include "core/bformatdec.h"
#undef protected
#undef private
#undef class

#include "core/async_event.h"
#include "core/buffer_storage.h"
#include "core/fmt_traits.h"
#include "core/context.h"
#include "core/device.h"
#include "core/effectslot.h"
#include "core/effects/base.h"
#include "alc/effects/base.h"
#include "core/front_stablizer.h"
#include "core/bs2b.h"
#include "core/hrtf.h"
#include "core/uhjfilter.h"
#include "core/tsmefilter.hpp"
#include "core/logging.h"
#include "core/voice.h"
#include "ringbuffer.h"

#include "b200mix_seam.h"
#include "../include/b200mix.h"

namespace {

struct Api {
    void *lib{nullptr};
    decltype(&b200mix_create) create{};
    decltype(&b200mix_destroy) destroy{};
    decltype(&b200mix_last_error) last_error{};
    decltype(&b200mix_set_hrtf_decoder) set_hrtf_decoder{};
    decltype(&b200mix_set_ambi_decoder) set_ambi_decoder{};
    decltype(&b200mix_set_uhj_encoder) set_uhj_encoder{};
    decltype(&b200mix_set_bs2b) set_bs2b{};
    decltype(&b200mix_set_front_stabilizer) set_front_stabilizer{};
    decltype(&b200mix_buffer_data) buffer_data{};
    decltype(&b200mix_buffer_data_adpcm) buffer_data_adpcm{};
    decltype(&b200mix_voices_update) voices_update{};
    decltype(&b200mix_voices_filters) voices_filters{};
    decltype(&b200mix_voice_queue) voice_queue{};
    decltype(&b200mix_render) render{};
    decltype(&b200mix_slot_efx) slot_efx{};
    decltype(&b200mix_slot_convolution) slot_convolution{};
    decltype(&b200mix_convolution_gains) convolution_gains{};
    decltype(&b200mix_resample_ir) resample_ir{};
    decltype(&b200mix_resampled_ir_frames) resampled_ir_frames{};
    decltype(&b200mix_slot_reverb) slot_reverb{};
    decltype(&b200mix_slot_reverb_update) slot_reverb_update{};
    decltype(&b200mix_slot_output_gains) slot_output_gains{};
    decltype(&b200mix_slot_target) slot_target{};
    decltype(&b200mix_slot_disable) slot_disable{};
    decltype(&b200mix_reverb_params_from_efx) reverb_params_from_efx{};
    decltype(&b200mix_reverb_full_update_needed) reverb_full_update_needed{};
    bool ok{false};
};

Api &api()
{
    static Api a = [] {
        Api r;
        const char *on = std::getenv("ALSOFT_B200MIX");
        if(!on || on[0] != '1') return r;
        const char *path = std::getenv("ALSOFT_B200MIX_LIB");
        r.lib = dlopen(path ? path : "libb200mix.so", RTLD_NOW | RTLD_LOCAL);
        if(!r.lib) { ERR("b200mix: cannot load the mixer library: {}", dlerror()); return r; }
#define LOAD(n) r.n = reinterpret_cast<decltype(r.n)>(dlsym(r.lib, "b200mix_" #n))
        LOAD(create); LOAD(destroy); LOAD(last_error); LOAD(set_hrtf_decoder); LOAD(set_ambi_decoder);
        LOAD(set_uhj_encoder); LOAD(set_bs2b); LOAD(set_front_stabilizer);
        LOAD(buffer_data); LOAD(buffer_data_adpcm); LOAD(voices_update); LOAD(voices_filters); LOAD(voice_queue); LOAD(render);
        LOAD(slot_convolution); LOAD(convolution_gains); LOAD(resample_ir); LOAD(resampled_ir_frames);
        LOAD(slot_efx); LOAD(slot_reverb); LOAD(slot_reverb_update); LOAD(slot_output_gains); LOAD(slot_target);
        LOAD(slot_disable); LOAD(reverb_params_from_efx); LOAD(reverb_full_update_needed);
#undef LOAD
        r.ok = r.create && r.destroy && r.last_error && r.set_hrtf_decoder && r.set_ambi_decoder
            && r.set_uhj_encoder && r.set_bs2b && r.set_front_stabilizer
            && r.buffer_data && r.buffer_data_adpcm && r.voices_update && r.voices_filters && r.voice_queue && r.render && r.slot_efx && r.slot_reverb
            && r.slot_reverb_update && r.slot_output_gains && r.slot_target && r.slot_disable
            && r.reverb_params_from_efx && r.reverb_full_update_needed && r.slot_convolution
            && r.convolution_gains && r.resample_ir && r.resampled_ir_frames;
        if(!r.ok) ERR("b200mix: the mixer library lacks entry points of include/b200mix.h");
        return r;
    }();
    return a;
}

struct ChanCache {                   /* one mixing channel = one device voice: what was last sent */
    uint32_t id{0};
    b200mix_voice_params params{};
    std::vector<float> coeffs, dry, send;
    std::vector<b200mix_voice_filter> filt;      /* per path: what the device has (empty: never sent) */
    std::vector<b200mix_voice_filter> last;      /* per path: the reference's targets when this channel was last looked at */
};
struct VoiceCache {                  /* a Voice of the reference: resend only on change */
    unsigned source_id{0};
    bool live{false};
    bool parked{false};                          /* paused: the device voice is stopped but keeps its state */
    std::vector<ChanCache> ch;
    std::vector<const VoiceBufferItem*> queue;   /* streaming sources: the list the device walks */
    uint64_t seen{0};                            /* the update that last found this Voice in a context */
    bool queue_loops{false};                     /* ... and whether it wraps to its first item (mLoopBuffer) */
};

struct SlotCache {                   /* what was last installed for an effect slot */
    const EffectSlotBase *slot{nullptr};
    const EffectState *state{nullptr};   /* a new EffectState object = deviceUpdate (al/auxeffectslot.cpp initEffect) */
    bool live{false};
    uint32_t kind{0};                /* 0 none, 1 reverb, 2 b200mix_slot_efx, 3 convolution */
    uint32_t conv_layout{0};         /* kind 3: layout code of b200mix_convolution_gains, IR channels */
    uint32_t conv_channels{0};
    b200mix_efx_props efx{};
    b200mix_efx_reverb reverb{};
    float gain{0.0f};
    uint32_t target{B200MIX_NO_SLOT};
};

/* What a device reset changes (alcResetDeviceSOFT / reopen -> UpdateDeviceParams -> aluInitRenderer):
 * the mixer is re-created when any of it differs from what it was opened with. */
struct DeviceSig {
    uint32_t rate{0}, dry{0}, real{0}, ir{0}, sends{0}, order{0};
    size_t post{0};
    const void *post_state{nullptr}, *dry_buf{nullptr};
    bool operator==(const DeviceSig&) const = default;
};

struct Seam {
    DeviceSig sig{};
    b200mix_device *dev{nullptr};
    b200mix_device_desc desc{};
    bool failed{false};
    bool reset_pending{false};           /* b200seam_device_reset since the last update */
    bool filters_on{false};              /* some voice has had an active direct / send filter: targets are forwarded */
    struct BufferEntry { uint32_t id, frames; uint64_t hash; };
    std::unordered_map<const void*, BufferEntry> buffers;                     /* sample data -> device copy */
    uint32_t next_buffer{0};
    std::vector<VoiceCache> cache;
    std::unordered_map<const Voice*, uint32_t> cache_of;   /* Voice object -> its entry of `cache` */
    std::vector<uint32_t> cache_free, cidx;                /* unused entries; entry of vptr[n] this update */
    uint32_t next_cache{0};
    uint64_t update_no{0};
    std::vector<b200mix_voice_params> upd;
    std::vector<float> upd_coeffs, upd_dry;
    std::vector<b200mix_voice_result> results;
    std::vector<Voice*> vptr;
    std::vector<ContextBase*> vctx;
    std::vector<SlotCache> slots;
    std::unordered_map<const EffectSlotBase*, uint32_t> slot_ids;
    std::unordered_map<const void*, uint32_t> wet_ids;                        /* Wet.Buffer.data() -> slot id */
    std::vector<float> upd_send;
    std::vector<b200mix_voice_filter> upd_filt, upd_filt_first;
    std::vector<uint32_t> free_ids, qids;      /* device voice ids; scratch */
    uint32_t next_id{0};
    std::vector<const VoiceBufferItem*> qnow;
};

std::mutex g_lock;

/* Impulse responses by ConvolutionState object (b200seam_note_convolution): planar floats at the
 * buffer's rate, converted like LoadSamples (core/voice.cpp:271-287, core/fmt_traits.h:88-175). */
struct ConvIr { uint32_t layout{0}, rate{0}, frames{0}, channels{0}; std::vector<float> planar; };
std::mutex g_conv_lock;
std::unordered_map<const void*, ConvIr> g_conv;
std::unordered_map<const DeviceBase*, Seam> g_seams;

constexpr uint32_t kMaxVoices = 16384, kMaxBuffers = 16384, kMaxSlots = 64;   /* 64: alc/alc.cpp:3427 */

int sample_type_of(const SampleVariant &sv, const void **data, size_t *span_bytes)
{
    return std::visit([data,span_bytes]<typename T>(std::span<T> const &spl) -> int {
        *data = spl.data();
        *span_bytes = spl.size_bytes();
        if constexpr(std::is_same_v<T,u8>) return B200MIX_FMT_U8;
        else if constexpr(std::is_same_v<T,i16>) return B200MIX_FMT_I16;
        else if constexpr(std::is_same_v<T,i32>) return B200MIX_FMT_I32;
        else if constexpr(std::is_same_v<T,f32>) return B200MIX_FMT_F32;
        else if constexpr(std::is_same_v<T,f64>) return B200MIX_FMT_F64;
        else if constexpr(std::is_same_v<T,MulawSample>) return B200MIX_FMT_MULAW;
        else if constexpr(std::is_same_v<T,AlawSample>) return B200MIX_FMT_ALAW;
        else if constexpr(std::is_same_v<T,IMA4Data>) return B200MIX_FMT_IMA4;
        else if constexpr(std::is_same_v<T,MSADPCMData>) return B200MIX_FMT_MSADPCM;
        else return -1;
    }, sv);
}

bool fail(DeviceBase *device, Seam &S, const char *what)
{
    S.failed = true;
    const char *detail = (S.dev && api().last_error) ? api().last_error(S.dev) : "";
    ERR("b200mix: {} {}", what, detail);
    device->handleDisconnect("b200mix: {} {}", what, detail);
    return false;
}

DeviceSig sig_of(const DeviceBase *device)
{
    DeviceSig g;
    g.rate = device->mSampleRate; g.dry = uint32_t(device->Dry.Buffer.size());
    g.real = uint32_t(device->RealOut.Buffer.size()); g.ir = device->mIrSize;
    g.sends = device->NumAuxSends; g.order = device->mAmbiOrder;
    g.post = device->mPostProcess.index();
    if(auto *h = std::get_if<HrtfPostProcess>(&device->mPostProcess)) g.post_state = h->mHrtfState.get();
    else if(auto *a = std::get_if<AmbiDecPostProcess>(&device->mPostProcess)) g.post_state = a->mAmbiDecoder.get();
    else if(auto *u = std::get_if<UhjPostProcess>(&device->mPostProcess)) g.post_state = u->mUhjEncoder.get();
    else if(auto *t = std::get_if<TsmePostProcess>(&device->mPostProcess)) g.post_state = t->mTsmeEncoder.get();
    else if(auto *sp = std::get_if<StablizerPostProcess>(&device->mPostProcess)) g.post_state = sp->mAmbiDecoder.get();
    else if(auto *bp = std::get_if<Bs2bPostProcess>(&device->mPostProcess)) g.post_state = bp->mAmbiDecoder.get();
    g.dry_buf = device->Dry.Buffer.data();
    return g;
}

/* What aluInitRenderer decided -> b200mix_create + the post-process constants. */
bool open_device(DeviceBase *device, Seam &S)
{
    Api &A = api();
    b200mix_device_desc &d = S.desc;
    d = b200mix_device_desc{};
    d.struct_size = sizeof(d);
    d.cuda_device = -1;
    d.sample_rate = device->mSampleRate;
    d.dry_channels = static_cast<uint32_t>(device->Dry.Buffer.size());
    d.real_channels = static_cast<uint32_t>(device->RealOut.Buffer.size());
    d.ir_size = device->mIrSize;
    d.real_left = device->RealOut.ChannelIndex[FrontLeft].c_val;
    d.real_right = device->RealOut.ChannelIndex[FrontRight].c_val;
    d.max_voices = kMaxVoices; d.max_buffers = kMaxBuffers;
    d.num_sends = std::min<uint32_t>(device->NumAuxSends, B200MIX_MAX_SENDS);
    d.wet_channels = d.num_sends ? static_cast<uint32_t>(AmbiChannelsFromOrder(device->mAmbiOrder)) : 0u; /* aluInitEffectPanning */
    d.max_slots = d.num_sends ? kMaxSlots : 0u;
    if(std::holds_alternative<HrtfPostProcess>(device->mPostProcess)) d.post_process = B200MIX_POST_HRTF;
    else if(std::holds_alternative<AmbiDecPostProcess>(device->mPostProcess)
        || std::holds_alternative<StablizerPostProcess>(device->mPostProcess)
        || std::holds_alternative<Bs2bPostProcess>(device->mPostProcess)) d.post_process = B200MIX_POST_AMBIDEC;
    else if(std::holds_alternative<UhjPostProcess>(device->mPostProcess)) d.post_process = B200MIX_POST_UHJ;
    else if(std::holds_alternative<TsmePostProcess>(device->mPostProcess)) d.post_process = B200MIX_POST_TSME;
    else if(std::holds_alternative<std::monostate>(device->mPostProcess)) d.post_process = B200MIX_POST_NONE;
    else return fail(device, S, "this post-process is not wired into the seam yet");
    if(A.create(&d, &S.dev) != B200MIX_OK) return fail(device, S, "b200mix_create failed:");
    if(auto *proc = std::get_if<HrtfPostProcess>(&device->mPostProcess))
    {
        auto &st = *proc->mHrtfState;
        const auto ir = st.mIrSize;
        std::vector<float> coeffs(st.mChannels.size()*ir*2), hf(st.mChannels.size()), sc(st.mChannels.size());
        auto c = 0_uz;
        for(auto &chan : st.mChannels)
        {
            for(auto j = 0u;j < ir;++j)
            {
                coeffs[(c*ir + j)*2 + 0] = chan.mCoeffs[j][0];
                coeffs[(c*ir + j)*2 + 1] = chan.mCoeffs[j][1];
            }
            hf[c] = chan.mHfScale; sc[c] = chan.mSplitter.mCoeff;
            ++c;
        }
        if(A.set_hrtf_decoder(S.dev, static_cast<uint32_t>(c), ir, coeffs.data(), hf.data(), sc.data()) != B200MIX_OK)
            return fail(device, S, "b200mix_set_hrtf_decoder failed:");
    }
    const BFormatDec *bdec = nullptr;
    if(auto *aproc = std::get_if<AmbiDecPostProcess>(&device->mPostProcess)) bdec = aproc->mAmbiDecoder.get();
    else if(auto *sproc = std::get_if<StablizerPostProcess>(&device->mPostProcess)) bdec = sproc->mAmbiDecoder.get();
    else if(auto *bproc = std::get_if<Bs2bPostProcess>(&device->mPostProcess)) bdec = bproc->mAmbiDecoder.get();
    if(bdec)
    {
        auto &dec = *bdec;
        const auto outs = device->RealOut.Buffer.size();
        std::vector<float> ghf(d.dry_channels*outs), glf(d.dry_channels*outs);
        float xover = 0.0f; bool dual = false;
        if(auto *sb = std::get_if<BFormatDec::SBandDecoderVector>(&dec.mChannelDec))
        {
            for(auto i = 0_uz;i < sb->size();++i)
                for(auto o = 0_uz;o < outs;++o) ghf[i*outs + o] = (*sb)[i].mGains[o];
        }
        else
        {
            auto &db = std::get<BFormatDec::DBandDecoderVector>(dec.mChannelDec);
            dual = true;
            for(auto i = 0_uz;i < db.size();++i)
                for(auto o = 0_uz;o < outs;++o)
                {
                    ghf[i*outs + o] = db[i].mGains[BFormatDec::sHFBand][o];
                    glf[i*outs + o] = db[i].mGains[BFormatDec::sLFBand][o];
                }
            xover = db.empty() ? 0.0f : db[0].mXOver.mCoeff;
        }
        if(A.set_ambi_decoder(S.dev, d.dry_channels, ghf.data(), dual ? glf.data() : nullptr, xover) != B200MIX_OK)
            return fail(device, S, "b200mix_set_ambi_decoder failed:");
    }
    if(auto *sproc = std::get_if<StablizerPostProcess>(&device->mPostProcess))
    {   /* alc/alu.cpp:330-406: the decode, then the front image stabilizer */
        if(A.set_front_stabilizer(S.dev, device->RealOut.ChannelIndex[FrontCenter].c_val,
            sproc->mStablizer->MidFilter.mCoeff) != B200MIX_OK)
            return fail(device, S, "b200mix_set_front_stabilizer failed:");
    }
    if(auto *bproc = std::get_if<Bs2bPostProcess>(&device->mPostProcess))
    {   /* alc/alu.cpp:408-434: the decode, then the BS2B crossfeed */
        if(A.set_bs2b(S.dev, static_cast<uint32_t>(bproc->mBs2b->level)) != B200MIX_OK)
            return fail(device, S, "b200mix_set_bs2b failed:");
    }
    {   /* UhjEncodeQuality / TsmeEncodeQuality picked the encoder class (alc/alc.cpp:564-597) */
        uint32_t flen = 0u; bool enc = false;
        if(auto *uproc = std::get_if<UhjPostProcess>(&device->mPostProcess))
        {
            enc = true;
            if(dynamic_cast<UhjEncoder<256>*>(uproc->mUhjEncoder.get())) flen = 256u;
            else if(dynamic_cast<UhjEncoder<512>*>(uproc->mUhjEncoder.get())) flen = 512u;
        }
        else if(auto *tproc = std::get_if<TsmePostProcess>(&device->mPostProcess))
        {
            enc = true;
            if(dynamic_cast<TsmeEncoder<256>*>(tproc->mTsmeEncoder.get())) flen = 256u;
            else if(dynamic_cast<TsmeEncoder<512>*>(tproc->mTsmeEncoder.get())) flen = 512u;
        }
        if(enc && A.set_uhj_encoder(S.dev, flen, nullptr) != B200MIX_OK)
            return fail(device, S, "b200mix_set_uhj_encoder failed:");
    }
    S.cache.assign(kMaxVoices, VoiceCache{});
    S.slots.assign(kMaxSlots, SlotCache{});
    S.sig = sig_of(device);
    S.results.assign(kMaxVoices, b200mix_voice_result{});
    return true;
}


/* EffectProps (core/effects/base.h:62-178) -> the ABI's plain structs.  Returns the slot kind:
 * 0 none, 1 reverb, 2 b200mix_slot_efx, -1 not wired. */
int effect_of(const EffectSlotBase *slot, b200mix_efx_props &o, b200mix_efx_reverb &rv)
{
    std::memset(&o, 0, sizeof(o)); o.struct_size = sizeof(o);
    std::memset(&rv, 0, sizeof(rv)); rv.struct_size = sizeof(rv);
    const EffectProps &props = slot->mEffectProps;
    switch(slot->EffectType)
    {
    case EffectSlotType::None: return 0;
    case EffectSlotType::Reverb:
        if(auto *p = std::get_if<ReverbProps>(&props))
        {
            rv.density = p->Density; rv.diffusion = p->Diffusion; rv.gain = p->Gain; rv.gain_hf = p->GainHF;
            rv.gain_lf = p->GainLF; rv.decay_time = p->DecayTime; rv.decay_hf_ratio = p->DecayHFRatio;
            rv.decay_lf_ratio = p->DecayLFRatio; rv.reflections_gain = p->ReflectionsGain;
            rv.reflections_delay = p->ReflectionsDelay; rv.late_reverb_gain = p->LateReverbGain;
            rv.late_reverb_delay = p->LateReverbDelay;
            for(int k = 0;k < 3;++k) { rv.reflections_pan[k] = p->ReflectionsPan[size_t(k)]; rv.late_reverb_pan[k] = p->LateReverbPan[size_t(k)]; }
            rv.echo_time = p->EchoTime; rv.echo_depth = p->EchoDepth; rv.modulation_time = p->ModulationTime;
            rv.modulation_depth = p->ModulationDepth; rv.air_absorption_gain_hf = p->AirAbsorptionGainHF;
            rv.hf_reference = p->HFReference; rv.lf_reference = p->LFReference;
            rv.room_rolloff_factor = p->RoomRolloffFactor; rv.decay_hf_limit = p->DecayHFLimit ? 1u : 0u;
            return 1;
        }
        return -1;
    case EffectSlotType::Echo:
        if(auto *p = std::get_if<EchoProps>(&props))
        { o.type = B200MIX_EFFECT_ECHO; o.echo.delay = p->Delay; o.echo.lr_delay = p->LRDelay; o.echo.damping = p->Damping;
          o.echo.feedback = p->Feedback; o.echo.spread = p->Spread; return 2; }
        return -1;
    case EffectSlotType::RingModulator:
        if(auto *p = std::get_if<ModulatorProps>(&props))
        { o.type = B200MIX_EFFECT_MODULATOR; o.modulator.frequency = p->Frequency; o.modulator.high_pass_cutoff = p->HighPassCutoff;
          o.modulator.waveform = static_cast<uint32_t>(p->Waveform); return 2; }
        return -1;
    case EffectSlotType::Equalizer:
        if(auto *p = std::get_if<EqualizerProps>(&props))
        { o.type = B200MIX_EFFECT_EQUALIZER; auto &e = o.equalizer;
          e.low_cutoff = p->LowCutoff; e.low_gain = p->LowGain; e.mid1_center = p->Mid1Center; e.mid1_gain = p->Mid1Gain;
          e.mid1_width = p->Mid1Width; e.mid2_center = p->Mid2Center; e.mid2_gain = p->Mid2Gain; e.mid2_width = p->Mid2Width;
          e.high_cutoff = p->HighCutoff; e.high_gain = p->HighGain; return 2; }
        return -1;
    case EffectSlotType::Compressor:
        if(auto *p = std::get_if<CompressorProps>(&props))
        { o.type = B200MIX_EFFECT_COMPRESSOR; o.compressor.on_off = p->OnOff ? 1u : 0u; return 2; }
        return -1;
    case EffectSlotType::Dedicated:
        if(auto *p = std::get_if<DedicatedProps>(&props))
        { o.type = B200MIX_EFFECT_DEDICATED; o.dedicated.target = p->Target == DedicatedProps::Lfe ? 1u : 0u;
          o.dedicated.gain = p->Gain; return 2; }
        return -1;
    case EffectSlotType::Distortion:
        if(auto *p = std::get_if<DistortionProps>(&props))
        { o.type = B200MIX_EFFECT_DISTORTION; o.distortion.edge = p->Edge; o.distortion.gain = p->Gain;
          o.distortion.lowpass_cutoff = p->LowpassCutoff; o.distortion.eq_center = p->EQCenter;
          o.distortion.eq_bandwidth = p->EQBandwidth; return 2; }
        return -1;
    case EffectSlotType::Chorus:
    case EffectSlotType::Flanger:
        if(auto *p = std::get_if<ChorusProps>(&props))
        { o.type = B200MIX_EFFECT_CHORUS; o.chorus.waveform = static_cast<uint32_t>(p->Waveform); o.chorus.phase = p->Phase;
          o.chorus.rate = p->Rate; o.chorus.depth = p->Depth; o.chorus.feedback = p->Feedback; o.chorus.delay = p->Delay; return 2; }
        return -1;
    case EffectSlotType::Autowah:
        if(auto *p = std::get_if<AutowahProps>(&props))
        { o.type = B200MIX_EFFECT_AUTOWAH; o.autowah.attack_time = p->AttackTime; o.autowah.release_time = p->ReleaseTime;
          o.autowah.resonance = p->Resonance; o.autowah.peak_gain = p->PeakGain; return 2; }
        return -1;
    case EffectSlotType::VocalMorpher:
        if(auto *p = std::get_if<VmorpherProps>(&props))
        { o.type = B200MIX_EFFECT_VMORPHER; o.vmorpher.rate = p->Rate;
          o.vmorpher.phoneme_a = static_cast<uint32_t>(p->PhonemeA); o.vmorpher.phoneme_b = static_cast<uint32_t>(p->PhonemeB);
          o.vmorpher.phoneme_a_coarse_tuning = p->PhonemeACoarseTuning; o.vmorpher.phoneme_b_coarse_tuning = p->PhonemeBCoarseTuning;
          o.vmorpher.waveform = static_cast<uint32_t>(p->Waveform); return 2; }
        return -1;
    case EffectSlotType::FrequencyShifter:
        if(auto *p = std::get_if<FshifterProps>(&props))
        { o.type = B200MIX_EFFECT_FSHIFTER; o.fshifter.frequency = p->Frequency;
          o.fshifter.left_direction = static_cast<uint32_t>(p->LeftDirection);
          o.fshifter.right_direction = static_cast<uint32_t>(p->RightDirection); return 2; }
        return -1;
    case EffectSlotType::PitchShifter:
        if(auto *p = std::get_if<PshifterProps>(&props))
        { o.type = B200MIX_EFFECT_PSHIFTER; o.pshifter.coarse_tune = p->CoarseTune; o.pshifter.fine_tune = p->FineTune; return 2; }
        return -1;
    case EffectSlotType::Convolution: return 3;
    default: return -1;
    }
}

uint32_t slot_id_of(Seam &S, const EffectSlotBase *slot)
{
    if(auto it = S.slot_ids.find(slot); it != S.slot_ids.end()) return it->second;
    for(uint32_t i = 0;i < kMaxSlots;++i)
        if(!S.slots[i].slot) { S.slots[i] = SlotCache{}; S.slots[i].slot = slot; S.slot_ids[slot] = i; return i; }
    return B200MIX_NO_SLOT;
}

/* CalcEffectSlotParams' result (alc/alu.cpp:557-636) -> b200mix_slot_*: every active slot whose
 * effect object, properties, gain or target differ from what the device has is (re)installed. */
bool sync_slots(DeviceBase *device, Seam &S)
{
    Api &A = api();
    std::array<bool, kMaxSlots> seen{};
    S.wet_ids.clear();
    for(ContextBase *ctx : *device->mContexts.load(std::memory_order_acquire))
    {
        auto *arr = ctx->mActiveAuxSlots.load(std::memory_order_acquire);
        if(!arr) continue;
        /* the array's second half is ProcessContexts' sorting scratch (alc/alu.cpp:2187-2189) */
        const auto all = std::span{*arr};
        for(EffectSlotBase *slot : all.first(all.size()>>1))
        {
            if(!S.desc.max_slots) return fail(device, S, "effect slots on a device without auxiliary sends");
            const uint32_t id = slot_id_of(S, slot);
            if(id == B200MIX_NO_SLOT) return fail(device, S, "more effect slots than the seam's device was created for");
            seen[id] = true;
            if(slot->Wet.Buffer.size() != S.desc.wet_channels) return fail(device, S, "unexpected Wet buffer size");
            S.wet_ids[slot->Wet.Buffer.data()] = id;
        }
    }
    for(uint32_t id = 0;id < kMaxSlots;++id)
    {
        SlotCache &C = S.slots[id];
        if(!C.slot) continue;
        if(!seen[id])
        {   /* left the active list (deleted, or nothing plays into it any more) */
            if(C.live && C.kind && A.slot_disable(S.dev, id) != B200MIX_OK) return fail(device, S, "b200mix_slot_disable failed:");
            S.slot_ids.erase(C.slot);
            C = SlotCache{};
            continue;
        }
        const EffectSlotBase *slot = C.slot;
        b200mix_efx_props efx; b200mix_efx_reverb rv;
        const int kind = effect_of(slot, efx, rv);
        if(kind < 0) return fail(device, S, "this effect type is not wired into the seam yet");
        uint32_t target = B200MIX_NO_SLOT;
        if(slot->Target)
        {
            target = slot_id_of(S, slot->Target);
            if(target == B200MIX_NO_SLOT || !seen[target]) return fail(device, S, "effect slot target outside the active set");
        }
        const EffectState *state = slot->mEffectState.get();
        const bool fresh = !C.live || C.state != state || C.kind != uint32_t(kind);
        const bool changed = fresh || C.gain != slot->Gain || C.target != target
            || std::memcmp(&C.efx, &efx, sizeof(efx)) != 0 || std::memcmp(&C.reverb, &rv, sizeof(rv)) != 0;
        if(!changed) continue;

        if(C.target != target || fresh)
            if(A.slot_target(S.dev, id, target) != B200MIX_OK) return fail(device, S, "b200mix_slot_target failed:");
        /* EffectTarget (alc/alu.cpp:627-633): the target slot's Wet mix, or the Dry mix */
        const MixParams &out = slot->Target ? slot->Target->Wet : device->Dry;
        std::array<float, MaxAmbiChannels> oscale{};
        std::array<uint32_t, MaxAmbiChannels> oindex{}, windex{};
        const auto nout = static_cast<uint32_t>(out.Buffer.size());
        if(nout > MaxAmbiChannels) return fail(device, S, "effect target wider than an ambisonic mix");
        for(uint32_t c = 0;c < nout;++c) { oscale[c] = out.AmbiMap[c].Scale; oindex[c] = out.AmbiMap[c].Index; }
        for(uint32_t c = 0;c < S.desc.wet_channels;++c) windex[c] = slot->Wet.AmbiMap[c].Index;
        if(kind == 0)
        {
            if(C.live && C.kind && A.slot_disable(S.dev, id) != B200MIX_OK) return fail(device, S, "b200mix_slot_disable failed:");
        }
        else if(kind == 1)
        {
            b200mix_reverb_target t{};
            t.struct_size = sizeof(t);
            t.sample_rate = device->mSampleRate; t.device_ambi_order = device->mAmbiOrder;
            t.device_2d = device->m2DMixing ? 1u : 0u; t.xover_freq = device->mXOverFreq;
            t.slot_gain = slot->Gain; t.reverb_boost = ReverbBoost;
            t.out_channels = nout; t.out_scale = oscale.data(); t.out_index = oindex.data();
            b200mix_reverb_params rp{};
            rp.struct_size = sizeof(rp);
            std::vector<float> gains(size_t(8)*nout);
            if(A.reverb_params_from_efx(&rv, &t, &rp, gains.data()) != B200MIX_OK)
                return fail(device, S, "b200mix_reverb_params_from_efx failed");
            if(fresh)
            {
                if(C.live && C.kind && A.slot_disable(S.dev, id) != B200MIX_OK) return fail(device, S, "b200mix_slot_disable failed:");
                if(A.slot_reverb(S.dev, id, &rp) != B200MIX_OK) return fail(device, S, "b200mix_slot_reverb failed:");
            }
            else
            {
                const uint32_t full = uint32_t(A.reverb_full_update_needed(&C.reverb, &rv));
                if(A.slot_reverb_update(S.dev, id, &rp, full) != B200MIX_OK)
                    return fail(device, S, "b200mix_slot_reverb_update failed:");
            }
            if(A.slot_output_gains(S.dev, id, 8u, gains.data()) != B200MIX_OK)
                return fail(device, S, "b200mix_slot_output_gains failed:");
        }
        else if(kind == 3)
        {
            /* ConvolutionState::deviceUpdate + update (alc/effects/convolution.cpp:318-471,541-620),
             * plain channel layouts (mono ... 7.1 impulse responses) */
            if(fresh)
            {
                ConvIr ir;
                {
                    std::lock_guard<std::mutex> cg{g_conv_lock};
                    if(auto it = g_conv.find(state); it != g_conv.end()) ir = it->second;
                }
                if(C.live && C.kind && A.slot_disable(S.dev, id) != B200MIX_OK) return fail(device, S, "b200mix_slot_disable failed:");
                C.conv_layout = ir.layout; C.conv_channels = ir.frames ? ir.channels : 0u;
                if(ir.frames && !ir.layout) return fail(device, S, "this impulse response format is not wired into the seam yet");
                if(ir.frames)
                {
                    uint32_t frames = ir.frames;
                    std::vector<float> planar;
                    if(ir.rate != device->mSampleRate)
                    {
                        frames = uint32_t(A.resampled_ir_frames(ir.rate, device->mSampleRate, ir.frames));
                        planar.resize(size_t(frames)*ir.channels);
                        for(uint32_t c = 0;c < ir.channels;++c)
                            if(A.resample_ir(ir.rate, device->mSampleRate, ir.planar.data() + size_t(c)*ir.frames, ir.frames,
                                planar.data() + size_t(c)*frames, frames) != B200MIX_OK)
                                return fail(device, S, "b200mix_resample_ir failed");
                    }
                    else planar = std::move(ir.planar);
                    if(A.slot_convolution(S.dev, id, ir.channels, frames, planar.data()) != B200MIX_OK)
                        return fail(device, S, "b200mix_slot_convolution failed:");
                }
            }
            if(C.conv_channels)
            {
                std::vector<float> gains(size_t(C.conv_channels)*nout);
                const int rows = A.convolution_gains(C.conv_layout, device->mRenderMode == RenderMode::Pairwise ? 1u : 0u,
                    slot->Gain, nout, oscale.data(), oindex.data(), gains.data(), nout);
                if(rows != int(C.conv_channels)) return fail(device, S, "b200mix_convolution_gains failed");
                if(A.slot_output_gains(S.dev, id, C.conv_channels, gains.data()) != B200MIX_OK)
                    return fail(device, S, "b200mix_slot_output_gains failed:");
            }
        }
        else
        {
            b200mix_efx_target t{};
            t.struct_size = sizeof(t);
            t.sample_rate = device->mSampleRate; t.slot_gain = slot->Gain;
            t.out_channels = nout; t.out_scale = oscale.data(); t.out_index = oindex.data();
            t.wet_channels = S.desc.wet_channels; t.wet_index = windex.data();
            const auto rc = device->RealOut.ChannelIndex[FrontCenter].c_val, rl = device->RealOut.ChannelIndex[LFE].c_val;
            t.real_center = (slot->Target || rc == InvalidChannelIndex.c_val) ? B200MIX_NO_SLOT : rc;
            t.real_lfe = (slot->Target || rl == InvalidChannelIndex.c_val) ? B200MIX_NO_SLOT : rl;
            t.device_ambi_order = device->mAmbiOrder;
            if(fresh && C.live && C.kind && A.slot_disable(S.dev, id) != B200MIX_OK)
                return fail(device, S, "b200mix_slot_disable failed:");
            if(A.slot_efx(S.dev, id, &efx, &t) != B200MIX_OK) return fail(device, S, "b200mix_slot_efx failed:");
        }
        C.live = true; C.state = state; C.kind = (kind == 3 && !C.conv_channels) ? 0u : uint32_t(kind); C.efx = efx; C.reverb = rv;
        C.gain = slot->Gain; C.target = target;
    }
    return true;
}

} // namespace

bool b200seam_enabled(const DeviceBase*) noexcept { return api().ok; }

void b200seam_render(DeviceBase *device, unsigned samplesToDo) noexcept
{
    Api &A = api();
    std::lock_guard<std::mutex> guard{g_lock};
    Seam &S = g_seams[device];
    if(S.failed) return;
    if(S.dev && (S.reset_pending || !(S.sig == sig_of(device))))
    {
        /* the device was reset: a new mixer; every voice and slot is sent again from the
         * reference's objects (positions are theirs, histories start clean like Voice::prepare) */
        A.destroy(S.dev);
        S = Seam{};
    }
    if(!S.dev && !open_device(device, S)) return;
    const uint32_t ir = S.desc.ir_size, cd = S.desc.dry_channels;

    /* ---- the voices of every context, in mixing order ---- */
    S.vptr.clear(); S.vctx.clear();
    for(ContextBase *ctx : *device->mContexts.load(std::memory_order_acquire))
    {
        for(Voice *voice : ctx->getVoicesSpanAcquired()) { S.vptr.push_back(voice); S.vctx.push_back(ctx); }
    }
    if(!sync_slots(device, S)) return;
    if(S.vptr.size() > kMaxVoices) { fail(device, S, "more voices than the seam's device was created for"); return; }

    S.upd.clear(); S.upd_coeffs.clear(); S.upd_dry.clear(); S.upd_send.clear(); S.upd_filt.clear(); S.upd_filt_first.clear();
    const uint32_t ns = S.desc.num_sends, cw = S.desc.wet_channels;
    std::vector<float> sg(size_t(ns)*cw);
    auto push_stopped = [&](const ChanCache &cc)
    {   /* remove a device voice from the active set */
        b200mix_voice_params p = cc.params;
        p.voice = cc.id; p.flags = B200MIX_VF_STOPPED;
        S.upd.push_back(p);
        S.upd_coeffs.insert(S.upd_coeffs.end(), size_t(ir)*2, 0.0f);
        S.upd_dry.insert(S.upd_dry.end(), cd, 0.0f);
        S.upd_send.insert(S.upd_send.end(), size_t(ns)*cw, 0.0f);
    };
    auto release = [&](VoiceCache &C)
    {
        for(const ChanCache &cc : C.ch) S.free_ids.push_back(cc.id);
        const uint64_t seen = C.seen;
        C = VoiceCache{};
        C.seen = seen;
    };
    /* One upload per (data pointer, length): BufferStorage is immutable while attached to a source.
     * A deleted buffer's memory can come back with other content, so a voice that starts checks a
     * hash of the samples too (not every update: ~6 us per 96 KB). */
    auto content_hash = [](const void *data, size_t bytes) -> uint64_t
    {
        uint64_t h = 0xcbf29ce484222325ull ^ bytes;
        const auto *w = static_cast<const unsigned char*>(data);
        size_t i = 0;
        for(;i + 8 <= bytes;i += 8) { uint64_t v; std::memcpy(&v, w + i, 8); h = (h ^ v) * 0x100000001b3ull; h ^= h >> 29; }
        for(;i < bytes;++i) h = (h ^ w[i]) * 0x100000001b3ull;
        return h;
    };
    auto buffer_of = [&](const VoiceBufferItem *item, uint32_t channels, uint32_t *out, bool verify) -> bool
    {
        const void *data = nullptr;
        size_t span_bytes = 0;
        const int type = sample_type_of(item->mSamples, &data, &span_bytes);
        if(type < 0) return fail(device, S, "buffer format not wired into the seam yet");
        /* IMA4 / MSADPCM stay block-compressed in BufferStorage (the reference decodes them in the
         * mixer, core/voice.cpp:289-484): the library decodes them once at upload. */
        const bool adpcm = type == B200MIX_FMT_IMA4 || type == B200MIX_FMT_MSADPCM;
        if(adpcm && (!item->mBlockAlign || item->mSampleLen % item->mBlockAlign))
            return fail(device, S, "block-compressed buffer with a partial block");
        static const size_t sz[] = {1, 2, 4, 4, 8, 1, 1};
        const size_t bytes = adpcm ? span_bytes : size_t(item->mSampleLen)*channels*sz[type];
        auto upload = [&](uint32_t id) -> bool
        {
            if(adpcm)
                return A.buffer_data_adpcm(S.dev, id, uint32_t(type), channels, item->mBlockAlign,
                    item->mSampleLen / item->mBlockAlign, data, bytes) == B200MIX_OK;
            return A.buffer_data(S.dev, id, uint32_t(type), channels, item->mSampleLen, data, bytes) == B200MIX_OK;
        };
        auto it = S.buffers.find(data);
        const bool known = it != S.buffers.end() && it->second.frames == item->mSampleLen;
        uint64_t hash = 0;
        if(!known || verify) hash = content_hash(data, bytes);
        if(!known || (verify && it->second.hash != hash))
        {
            uint32_t id = it == S.buffers.end() ? S.next_buffer++ : it->second.id;
            if(id >= kMaxBuffers) return fail(device, S, "more buffers than the seam's device was created for");
            if(!upload(id))
            {
                /* the old copy is still playing somewhere: leave it, take a new id */
                id = S.next_buffer++;
                if(id >= kMaxBuffers || !upload(id))
                    return fail(device, S, "b200mix_buffer_data failed:");
            }
            it = S.buffers.insert_or_assign(data, Seam::BufferEntry{id, item->mSampleLen, hash}).first;
        }
        *out = it->second.id;
        return true;
    };
    /* the queue as the mixer will walk it: [current .. last] and, for a looping queue, the items
     * before the current one (mLoopBuffer is the queue's head; core/voice.cpp:1183-1196) */
    auto queue_of = [](const Voice *voice, std::vector<const VoiceBufferItem*> &out)
    {
        out.clear();
        const VoiceBufferItem *cur = voice->mCurrentBuffer.load(std::memory_order_relaxed);
        const VoiceBufferItem *loop = voice->mLoopBuffer.load(std::memory_order_relaxed);
        for(auto *it = cur;it && out.size() < B200MIX_MAX_QUEUE;it = it->mNext.load(std::memory_order_relaxed))
            out.push_back(it);
        for(auto *it = loop;it && it != cur && out.size() < B200MIX_MAX_QUEUE;it = it->mNext.load(std::memory_order_relaxed))
            out.push_back(it);
    };

    /* The cache entry of every Voice object, by address: a context's voice array grows in place
     * (ContextBase::allocVoices appends clusters) but the arrays of the contexts after it then
     * start later in this list, and contexts come and go. */
    ++S.update_no;
    S.cidx.assign(S.vptr.size(), UINT32_MAX);
    for(size_t n = 0;n < S.vptr.size();++n)
        if(auto it = S.cache_of.find(S.vptr[n]); it != S.cache_of.end())
        { S.cidx[n] = it->second; S.cache[it->second].seen = S.update_no; }
    for(auto it = S.cache_of.begin();it != S.cache_of.end();)
    {
        VoiceCache &C = S.cache[it->second];
        if(C.seen == S.update_no) { ++it; continue; }
        /* its context was destroyed (alcDestroyContext drops it from mContexts, alc/alc.cpp:3130-3160):
         * the reference stops mixing its voices there and then */
        if(C.live) { for(const ChanCache &cc : C.ch) push_stopped(cc); release(C); }
        S.cache_free.push_back(it->second);
        it = S.cache_of.erase(it);
    }
    for(size_t n = 0;n < S.vptr.size();++n)
    {
        if(S.cidx[n] != UINT32_MAX) continue;
        uint32_t idx;
        if(!S.cache_free.empty()) { idx = S.cache_free.back(); S.cache_free.pop_back(); }
        else idx = S.next_cache++;
        S.cache[idx] = VoiceCache{};
        S.cache[idx].seen = S.update_no;
        S.cache_of.emplace(S.vptr[n], idx);
        S.cidx[n] = idx;
    }

    /* Filters cross the ABI from the first update in which any playing voice has one (most
     * applications never attach any: the library then runs without filter records). */
    if(!S.filters_on)
        for(const Voice *voice : S.vptr)
        {
            const auto pstate = voice->mPlayState.load(std::memory_order_acquire);
            if(pstate != Voice::Playing && pstate != Voice::Stopping) continue;
            bool any = voice->mDirect.FilterActive;
            for(uint32_t snd = 0;snd < ns && !any;++snd) any = voice->mSend[snd].FilterActive;
            if(any) { S.filters_on = true; break; }
        }

    for(size_t n = 0;n < S.vptr.size();++n)
    {
        Voice *voice = S.vptr[n];
        VoiceCache &C = S.cache[S.cidx[n]];
        const auto pstate = voice->mPlayState.load(std::memory_order_acquire);
        const bool active = pstate == Voice::Playing || pstate == Voice::Stopping;
        if(!active)
        {
            const unsigned sid0 = voice->mSourceID.load(std::memory_order_relaxed);
            if(C.live && C.parked && sid0 != 0u && sid0 == C.source_id)
                continue;       /* paused: its (stopped) device voice waits for the resume */
            if(C.live)
            {   /* the host stopped it (alSourceStop / rewind): remove it from the active set */
                for(const ChanCache &cc : C.ch) push_stopped(cc);
                release(C);
            }
            continue;
        }
        if(pstate == Voice::Stopping && !C.live && !voice->mCurrentBuffer.load(std::memory_order_relaxed))
        {
            /* started and stopped (or its source deleted) between two updates: the parameter stage
             * never ran for it (alc/alu.cpp:2168-2172 skips voices without a source), its history is
             * the silence Voice::prepare left, and Voice::mix would fade that silence out
             * (core/voice.cpp:704-719) and leave the voice stopped */
            voice->mPlayState.store(Voice::Stopped, std::memory_order_release);
            continue;
        }
        C.parked = false;
        const bool mono = voice->mFmtChannels == FmtMono;
        const char *unwired = nullptr;
        if(voice->mFlags.test(VoiceFlag::IsCallback)) unwired = "callback buffers";
        else if(voice->mFlags.test(VoiceFlag::IsAmbisonic)) unwired = "up-sampled ambisonic sources";
        else if(voice->mFlags.test(VoiceFlag::HasNfc)) unwired = "near-field control filters";
        else if(voice->mDecoder || voice->mFmtChannels == FmtUHJ2 || voice->mFmtChannels == FmtSuperStereo)
            unwired = "UHJ / super-stereo sources";
        /* direct channels mix straight into RealOut (alc/alu.cpp:1592-1598); HRTF voices name RealOut too */
        else if(!voice->mFlags.test(VoiceFlag::HasHrtf) && !voice->mDirect.Buffer.empty()
            && voice->mDirect.Buffer.data() != device->Dry.Buffer.data())
            unwired = "direct-channel sources";
        if(unwired)
        {
            ERR("b200mix: {} are not wired into the seam yet", unwired);
            fail(device, S, "a source uses a feature outside the seam:");
            return;
        }
        const bool isStatic = voice->mFlags.test(VoiceFlag::IsStatic);
        const uint32_t nch = (mono && !voice->mDuplicateMono) ? 1u : static_cast<uint32_t>(voice->mChans.size());
        const uint32_t bufch = std::max(voice->mFrameStep, 1u);
        auto *item = voice->mCurrentBuffer.load(std::memory_order_relaxed);
        auto *loop = voice->mLoopBuffer.load(std::memory_order_relaxed);

        const unsigned sid = voice->mSourceID.load(std::memory_order_relaxed);
        /* a stop clears mSourceID while the voice fades out: that is not a new voice */
        const bool fresh = !C.live || (sid != 0u && C.source_id != sid) || C.ch.size() != nch;
        if(fresh)
        {
            if(C.live) { for(const ChanCache &cc : C.ch) push_stopped(cc); release(C); }
            C.ch.resize(nch);
            for(ChanCache &cc : C.ch)
            {
                if(!S.free_ids.empty()) { cc.id = S.free_ids.back(); S.free_ids.pop_back(); }
                else cc.id = S.next_id++;
                if(cc.id >= kMaxVoices) { fail(device, S, "more mixing channels than the seam's device was created for"); return; }
            }
        }

        uint32_t bufid = B200MIX_NO_BUFFER;
        if(!isStatic)
        {   /* streaming source: (re)send the list when it is not what the device walks */
            queue_of(voice, S.qnow);
            /* (a looping queue whose current item is its head lists the same items as the same
             * queue with looping just switched off: the loop point is part of the comparison) */
            if(fresh || S.qnow != C.queue || C.queue_loops != (loop != nullptr))
            {
                /* empty buffers in a queue are legal AL (the reference steps over them, core/voice.cpp:
                 * 1183-1196); the device list holds the items that have samples */
                S.qids.clear();
                for(const VoiceBufferItem *qi : S.qnow)
                {
                    if(qi->mSampleLen == 0u) continue;
                    uint32_t id = 0;
                    if(!buffer_of(qi, bufch, &id, true)) return;
                    S.qids.push_back(id);
                }
                for(const ChanCache &cc : C.ch)
                    if(A.voice_queue(S.dev, cc.id, uint32_t(S.qids.size()), S.qids.data(),
                        (loop && !S.qids.empty()) ? 0u : B200MIX_NO_LOOP) != B200MIX_OK)
                    { fail(device, S, "b200mix_voice_queue failed:"); return; }
                C.queue = S.qnow;
                C.queue_loops = loop != nullptr;
            }
            bool hasSamples = false;
            for(const VoiceBufferItem *qi : C.queue) hasSamples = hasSamples || qi->mSampleLen != 0u;
            bufid = (item && hasSamples) ? 0u : B200MIX_NO_BUFFER;
        }
        else if(item && !buffer_of(item, bufch, &bufid, fresh)) return;

        for(uint32_t c = 0;c < nch;++c)
        {
        auto &ch = voice->mChans[c];
        ChanCache &CC = C.ch[c];
        b200mix_voice_params p{};
        p.voice = CC.id;
        p.flags = (pstate == Voice::Playing ? B200MIX_VF_PLAYING : B200MIX_VF_STOPPING)
            | (isStatic ? B200MIX_VF_STATIC : 0u) | B200MIX_VF_CHANNEL(mono ? 0u : c);
        if(loop) p.flags |= B200MIX_VF_LOOPING;
        if(voice->mFlags.test(VoiceFlag::HasHrtf)) p.flags |= B200MIX_VF_HRTF;
        p.resampler = static_cast<uint32_t>(voice->mProps.mResampler);
        p.step = voice->mStep;
        p.hrtf_delay[0] = ch.mDryParams.Hrtf.Target.Delay[0];
        p.hrtf_delay[1] = ch.mDryParams.Hrtf.Target.Delay[1];
        p.hrtf_gain = ch.mDryParams.Hrtf.Target.Gain;
        for(auto &s : p.send_slot) s = B200MIX_NO_SLOT;
        /* auxiliary sends (core/voice.cpp:967-980): the slot by its Wet buffer, the gains the ALU set */
        std::fill(sg.begin(), sg.end(), 0.0f);
        for(uint32_t snd = 0;snd < ns;++snd)
        {
            const auto &tgt = voice->mSend[snd];
            if(tgt.Buffer.empty()) continue;
            auto wit = S.wet_ids.find(tgt.Buffer.data());
            if(wit == S.wet_ids.end())
            { fail(device, S, "send into a slot outside the active set"); return; }
            p.send_slot[snd] = wit->second;
            const float *wg = ch.mWetParams[snd].Gains.Target.data();
            std::copy_n(wg, cw, sg.begin() + size_t(snd)*cw);
        }
        p.buffer = bufid;                       /* B200MIX_NO_BUFFER: alSourceStop / rewind took it (alc/alu.cpp:2071) */
        if(item) { p.loop_start = item->mLoopStart; p.loop_end = item->mLoopEnd; }
        if(!isStatic && p.loop_end <= p.loop_start) p.flags &= ~uint32_t(B200MIX_VF_LOOPING);   /* a queue loops through its list, not through loop points */
        if(fresh)
        {
            /* Voice::prepare + the start offset the AL layer set (al/source.cpp) */
            p.flags |= B200MIX_VF_RESET;
            if(voice->mFlags.test(VoiceFlag::IsFading)) p.flags |= B200MIX_VF_FADING;
            p.position = voice->mPosition.load(std::memory_order_relaxed);
            p.position_frac = voice->mPositionFrac.load(std::memory_order_relaxed);
        }
        const float *co = &ch.mDryParams.Hrtf.Target.Coeffs[0][0];
        const float *dg = ch.mDryParams.Gains.Target.data();
        const bool hrtf = (p.flags & B200MIX_VF_HRTF) != 0;
        bool changed = fresh || std::memcmp(&CC.params, &p, sizeof(p)) != 0;
        if(!changed && hrtf && ir) changed = std::memcmp(CC.coeffs.data(), co, size_t(ir)*2*sizeof(float)) != 0;
        if(!changed && !hrtf) changed = std::memcmp(CC.dry.data(), dg, cd*sizeof(float)) != 0;
        if(!changed && !sg.empty()) changed = std::memcmp(CC.send.data(), sg.data(), sg.size()*sizeof(float)) != 0;
        if(changed)
        {
            S.upd.push_back(p);
            S.upd_coeffs.insert(S.upd_coeffs.end(), co, co + size_t(ir)*2);
            S.upd_dry.insert(S.upd_dry.end(), dg, dg + cd);
            S.upd_send.insert(S.upd_send.end(), sg.begin(), sg.end());
            CC.send = sg;
            b200mix_voice_params keep = p;
            keep.flags &= ~uint32_t(B200MIX_VF_RESET | B200MIX_VF_FADING);
            keep.position = 0; keep.position_frac = 0;
            CC.params = keep;
            CC.coeffs.assign(co, co + size_t(ir)*2);
            CC.dry.assign(dg, dg + cd);
        }
        /* direct / send filters: the targets the ALU's filter block left (alc/alu.cpp:1619-1656).
         * Every reference setParams call that moved a target is forwarded in the same update, so
         * the device's copy of BiquadInterpFilter::setParams' rule sees the same sequence. */
        {
            const uint32_t paths = 1u + ns;
            if(fresh) { CC.filt.clear(); CC.last.clear(); }
            auto entry = [&](uint32_t path, const BiquadInterpFilter &lp, const BiquadInterpFilter &hp, bool act)
            {
                b200mix_voice_filter f{};
                f.voice = p.voice; f.path = path; f.active = act ? 1u : 0u;
                f.lowpass[0] = lp.mTargetCoeffs.mB0; f.lowpass[1] = lp.mTargetCoeffs.mB1; f.lowpass[2] = lp.mTargetCoeffs.mB2;
                f.lowpass[3] = lp.mTargetCoeffs.mA1; f.lowpass[4] = lp.mTargetCoeffs.mA2;
                f.highpass[0] = hp.mTargetCoeffs.mB0; f.highpass[1] = hp.mTargetCoeffs.mB1; f.highpass[2] = hp.mTargetCoeffs.mB2;
                f.highpass[3] = hp.mTargetCoeffs.mA1; f.highpass[4] = hp.mTargetCoeffs.mA2;
                return f;
            };
            std::array<b200mix_voice_filter, 1u + B200MIX_MAX_SENDS> now{};
            now[0] = entry(0u, ch.mDryParams.LowPass, ch.mDryParams.HighPass, voice->mDirect.FilterActive);
            for(uint32_t snd = 0;snd < ns;++snd)
                now[1u + snd] = entry(1u + snd, ch.mWetParams[snd].LowPass, ch.mWetParams[snd].HighPass,
                    voice->mSend[snd].FilterActive && p.send_slot[snd] != B200MIX_NO_SLOT);
            if(!S.filters_on)
            {   /* no filter anywhere on this device yet: nothing crosses the ABI (the library then
                 * keeps no filter records at all); the targets are remembered for the day one does */
                CC.last.assign(now.begin(), now.begin() + paths);
            }
            else if(CC.filt.empty())
            {   /* a voice that starts: all its paths, active or not.  The reference designs the
                 * (identity) shelves of an unfiltered path too, and DoFilters' clear() keeps
                 * mCoeffs on them, so a filter attached later interpolates from THOSE coefficients
                 * (core/filters/biquad.cpp:131-149) — the device needs the same starting point */
                const bool played_unfiltered = !fresh && !CC.last.empty();
                if(played_unfiltered)
                    /* it played before the first filter of this device came up: the shelves the
                     * reference held for it go first (its records are still in their reset state and
                     * adopt them at once), then this update's targets interpolate from them */
                    S.upd_filt_first.insert(S.upd_filt_first.end(), CC.last.begin(), CC.last.end());
                CC.last.assign(now.begin(), now.begin() + paths);
                CC.filt.assign(now.begin(), now.begin() + paths);
                S.upd_filt.insert(S.upd_filt.end(), now.begin(), now.begin() + paths);
                /* ... and one that starts with an interpolation already pending (its source was
                 * paused, sought — which moves it to a new Voice — and given another filter before
                 * it resumed: the parameter stage runs for every voice that has a source,
                 * alc/alu.cpp:2168-2172): the coefficients it starts from go first */
                auto pending = [](const BiquadInterpFilter &f) { return f.mCounter > 0; };
                auto current = [&](uint32_t q, const BiquadInterpFilter &lp, const BiquadInterpFilter &hp)
                {
                    if(played_unfiltered || (!pending(lp) && !pending(hp))) return;
                    b200mix_voice_filter f = now[q];
                    f.lowpass[0] = lp.mCoeffs.mB0; f.lowpass[1] = lp.mCoeffs.mB1; f.lowpass[2] = lp.mCoeffs.mB2;
                    f.lowpass[3] = lp.mCoeffs.mA1; f.lowpass[4] = lp.mCoeffs.mA2;
                    f.highpass[0] = hp.mCoeffs.mB0; f.highpass[1] = hp.mCoeffs.mB1; f.highpass[2] = hp.mCoeffs.mB2;
                    f.highpass[3] = hp.mCoeffs.mA1; f.highpass[4] = hp.mCoeffs.mA2;
                    S.upd_filt_first.push_back(f);
                };
                current(0u, ch.mDryParams.LowPass, ch.mDryParams.HighPass);
                for(uint32_t snd = 0;snd < ns;++snd)
                    current(1u + snd, ch.mWetParams[snd].LowPass, ch.mWetParams[snd].HighPass);
            }
            else
            {
                for(uint32_t q = 0;q < paths;++q)
                    if(std::memcmp(&CC.filt[q], &now[q], sizeof(now[q])) != 0)
                    { CC.filt[q] = now[q]; S.upd_filt.push_back(now[q]); }
                CC.last.assign(now.begin(), now.begin() + paths);
            }
        }
        }   /* channels */
        C.live = true; C.source_id = sid;
    }
    if(!S.upd.empty()
        && A.voices_update(S.dev, uint32_t(S.upd.size()), S.upd.data(), ir ? S.upd_coeffs.data() : nullptr,
            S.upd_dry.data(), sg.empty() ? nullptr : S.upd_send.data()) != B200MIX_OK)
    { fail(device, S, "b200mix_voices_update failed:"); return; }

    /* two calls, in this order: the library applies one call's entries concurrently */
    if(!S.upd_filt_first.empty()
        && A.voices_filters(S.dev, uint32_t(S.upd_filt_first.size()), S.upd_filt_first.data()) != B200MIX_OK)
    { fail(device, S, "b200mix_voices_filters failed:"); return; }
    if(!S.upd_filt.empty()
        && A.voices_filters(S.dev, uint32_t(S.upd_filt.size()), S.upd_filt.data()) != B200MIX_OK)
    { fail(device, S, "b200mix_voices_filters failed:"); return; }

    /* ---- the update itself: RealOut comes back planar, where Limiter / Write<T> expect it ---- */
    std::array<float*, B200MIX_MAX_DRY_CHANNELS> outs{};
    for(size_t c = 0;c < device->RealOut.Buffer.size();++c) outs[c] = device->RealOut.Buffer[c].data();
    if(A.render(S.dev, samplesToDo, outs.data(), S.results.data()) != B200MIX_OK)
    { fail(device, S, "b200mix_render failed:"); return; }

    /* ---- cursor and play state back into the Voice objects (core/voice.cpp:1116-1232) ---- */
    for(size_t n = 0;n < S.vptr.size();++n)
    {
        Voice *voice = S.vptr[n];
        VoiceCache &C = S.cache[S.cidx[n]];
        if(!C.live) continue;
        ContextBase *ctx = S.vctx[n];
        const b200mix_voice_result &r = S.results[C.ch[0].id];       /* all channels move together */
        voice->mPosition.store(r.position, std::memory_order_relaxed);
        voice->mPositionFrac.store(r.position_frac, std::memory_order_relaxed);
        voice->mFlags.set(VoiceFlag::IsFading);
        const unsigned sid = voice->mSourceID.load(std::memory_order_relaxed);
        uint32_t itemsDone = 0;
        if(!voice->mFlags.test(VoiceFlag::IsStatic) && (r.position > 0 || r.buffers_done))
        {
            /* streaming source: the queue advance of core/voice.cpp:1183-1196 — the device counts
             * the items with samples it finished; empty items in between go with them */
            auto *it = voice->mCurrentBuffer.load(std::memory_order_relaxed);
            auto *lp = voice->mLoopBuffer.load(std::memory_order_relaxed);
            uint32_t real = r.buffers_done;
            while(it && itemsDone < 4u*B200MIX_MAX_QUEUE)
            {
                if(it->mSampleLen != 0u) { if(!real) break; --real; }
                it = it->mNext.load(std::memory_order_relaxed);
                if(!it) it = lp;
                ++itemsDone;
            }
            if(itemsDone) voice->mCurrentBuffer.store(it, std::memory_order_release);
        }
        if(itemsDone)
        {
            queue_of(voice, C.queue);            /* what the device now walks */
            if(sid && ctx->mEnabledEvts.load(std::memory_order_acquire).test(AsyncEnableBits::BufferCompleted))
            {
                auto *ring = ctx->mAsyncEvents.get();
                if(auto vec = ring->getWriteVector(); !vec[0].empty())
                {
                    auto &evt = InitAsyncEvent<AsyncBufferCompleteEvent>(vec[0].front());
                    evt.mId = sid;
                    evt.mCount = itemsDone;
                    ring->writeAdvance(1);
                }
            }
        }
        if(r.flags & B200MIX_VF_STOPPED)
        {
            /* the fade-out update of a Stopping voice (core/voice.cpp:1119-1123): only the state
             * changes.  Buffer and source id were cleared when it ran out or was stopped; a PAUSED
             * source keeps both, and its position, for the resume. */
            voice->mPlayState.store(Voice::Stopped, std::memory_order_release);
            if(voice->mSourceID.load(std::memory_order_relaxed) != 0u
                && voice->mCurrentBuffer.load(std::memory_order_relaxed) != nullptr)
            {
                /* alSourcePause: the voice stays the source's; a later alSourcePlay continues it with
                 * its resampler / HRTF history and gains (al/source.cpp, VChangeState::Play on a
                 * paused source) — the device voice keeps its record, only its cached flags change */
                C.parked = true;
                for(ChanCache &cc : C.ch)
                    cc.params.flags = (cc.params.flags & ~uint32_t(B200MIX_VF_PLAYING | B200MIX_VF_STOPPING)) | B200MIX_VF_STOPPED;
            }
            else release(C);
        }
        else if((r.flags & B200MIX_VF_STOPPING) && voice->mPlayState.load(std::memory_order_relaxed) == Voice::Playing)
        {
            /* ran out of data: the source reads as stopped from now on, the voice fades for
             * one more update (core/voice.cpp:1198-1232) */
            voice->mCurrentBuffer.store(nullptr, std::memory_order_release);
            voice->mLoopBuffer.store(nullptr, std::memory_order_relaxed);
            voice->mSourceID.store(0u, std::memory_order_release);
            voice->mPlayState.store(Voice::Stopping, std::memory_order_release);
            C.queue.clear();
            if(ctx->mEnabledEvts.load(std::memory_order_acquire).test(AsyncEnableBits::SourceState))
            {
                auto *ring = ctx->mAsyncEvents.get();
                if(auto vec = ring->getWriteVector(); !vec[0].empty())
                {
                    auto &evt = InitAsyncEvent<AsyncSourceStateEvent>(vec[0].front());
                    evt.mId = sid;
                    evt.mState = AsyncSrcState::Stop;
                    ring->writeAdvance(1);
                }
            }
        }
    }
}

void b200seam_note_convolution(const void *state, const BufferStorage *buffer) noexcept
{
    if(!api().ok) return;
    ConvIr ir;
    if(buffer && buffer->mSampleLen >= 1)
    {
        switch(buffer->mChannels)
        {
        case FmtMono: ir.layout = 1u; break;
        case FmtStereo: ir.layout = B200MIX_LAYOUT_STEREO; break;
        case FmtRear: ir.layout = B200MIX_LAYOUT_REAR; break;
        case FmtQuad: ir.layout = B200MIX_LAYOUT_QUAD; break;
        case FmtX51: ir.layout = B200MIX_LAYOUT_X51; break;
        case FmtX61: ir.layout = B200MIX_LAYOUT_X61; break;
        case FmtX71: ir.layout = B200MIX_LAYOUT_X71; break;
        default: ir.layout = 0u; break;          /* B-Format / UHJ responses: not wired */
        }
        ir.rate = buffer->mSampleRate; ir.frames = buffer->mSampleLen; ir.channels = buffer->channelsFromFmt();
        const bool ok = std::visit([&ir]<typename T>(std::span<T> const &spl) -> bool
        {
            if constexpr(std::is_same_v<T,IMA4Data> || std::is_same_v<T,MSADPCMData>) return false;
            else
            {
                if(spl.size() < size_t(ir.frames)*ir.channels) return false;
                ir.planar.resize(size_t(ir.frames)*ir.channels);
                for(uint32_t c = 0;c < ir.channels;++c)
                    for(uint32_t k = 0;k < ir.frames;++k)
                        ir.planar[size_t(c)*ir.frames + k] = SampleInfo<T>::to_float(spl[size_t(k)*ir.channels + c]);
                return true;
            }
        }, buffer->mData);
        if(!ok) { ir.layout = 0u; ir.planar.clear(); }
    }
    std::lock_guard<std::mutex> cg{g_conv_lock};
    g_conv[state] = std::move(ir);
}

void b200seam_device_closed(const DeviceBase *device) noexcept
{
    if(!api().ok) return;
    std::lock_guard<std::mutex> guard{g_lock};
    if(auto it = g_seams.find(device); it != g_seams.end())
    {
        if(it->second.dev) api().destroy(it->second.dev);
        g_seams.erase(it);
    }
}

void b200seam_device_reset(const DeviceBase *device) noexcept
{
    if(!api().ok) return;
    std::lock_guard<std::mutex> guard{g_lock};
    if(auto it = g_seams.find(device); it != g_seams.end())
    {
        it->second.reset_pending = true;
        it->second.failed = false;       /* ResetDeviceParams reconnects the device: the mixer gets another try */
    }
}
