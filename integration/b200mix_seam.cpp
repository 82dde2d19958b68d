/* b200mix_seam.cpp — the reference-side binding of libb200mix.so: compiled INTO a patched
 * libopenal (integration/alu_seam.patch adds the three call sites in alc/alu.cpp) against the
 * reference's own headers.  Nothing here mixes: it converts the reference's live post-ALU
 * objects (DeviceBase, Voice, VoiceBufferItem) into the C ABI of include/b200mix.h, calls
 * b200mix_render, and writes the playback cursor back.
 *
 *   ALSOFT_B200MIX=1            enables the seam (else the library behaves as stock OpenAL Soft)
 *   ALSOFT_B200MIX_LIB=<path>   libb200mix.so (default: "libb200mix.so" on the loader path)
 *
 * Scope of this binding (v1): HRTF and ambisonic-decode devices, static mono sources of any
 * PCM sample type, any resampler, moving sources (targets are re-sent when the ALU changed
 * them), source start / stop / loop.  Streaming queues, multi-channel sources, direct/send
 * filters and auxiliary effect slots are forwarded by the C ABI (b200mix_voice_queue,
 * B200MIX_VF_CHANNEL, b200mix_voices_filters, b200mix_slot_*) but not wired up here yet: the
 * seam disconnects the device with a message rather than mixing them wrong.
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

/* BandSplitter::mCoeff and BFormatDec's matrices are private: a maintainer would add two
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
#include "core/bformatdec.h"
#undef protected
#undef private
#undef class

#include "core/async_event.h"
#include "core/context.h"
#include "core/device.h"
#include "core/effectslot.h"
#include "core/hrtf.h"
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
    decltype(&b200mix_buffer_data) buffer_data{};
    decltype(&b200mix_voices_update) voices_update{};
    decltype(&b200mix_render) render{};
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
        LOAD(buffer_data); LOAD(voices_update); LOAD(render);
#undef LOAD
        r.ok = r.create && r.destroy && r.last_error && r.set_hrtf_decoder && r.set_ambi_decoder
            && r.buffer_data && r.voices_update && r.render;
        if(!r.ok) ERR("b200mix: the mixer library lacks entry points of include/b200mix.h");
        return r;
    }();
    return a;
}

struct VoiceCache {                  /* what was last sent for a voice: resend only on change */
    unsigned source_id{0};
    bool live{false};
    b200mix_voice_params params{};
    std::vector<float> coeffs, dry;
};

struct Seam {
    b200mix_device *dev{nullptr};
    b200mix_device_desc desc{};
    bool failed{false};
    std::unordered_map<const void*, std::pair<uint32_t, uint32_t>> buffers;   /* data -> (id, frames) */
    uint32_t next_buffer{0};
    std::vector<VoiceCache> cache;
    std::vector<b200mix_voice_params> upd;
    std::vector<float> upd_coeffs, upd_dry;
    std::vector<b200mix_voice_result> results;
    std::vector<Voice*> vptr;
    std::vector<ContextBase*> vctx;
};

std::mutex g_lock;
std::unordered_map<const DeviceBase*, Seam> g_seams;

constexpr uint32_t kMaxVoices = 16384, kMaxBuffers = 16384;

int sample_type_of(const SampleVariant &sv, const void **data)
{
    return std::visit([data]<typename T>(std::span<T> const &spl) -> int {
        *data = spl.data();
        if constexpr(std::is_same_v<T,u8>) return B200MIX_FMT_U8;
        else if constexpr(std::is_same_v<T,i16>) return B200MIX_FMT_I16;
        else if constexpr(std::is_same_v<T,i32>) return B200MIX_FMT_I32;
        else if constexpr(std::is_same_v<T,f32>) return B200MIX_FMT_F32;
        else if constexpr(std::is_same_v<T,f64>) return B200MIX_FMT_F64;
        else if constexpr(std::is_same_v<T,MulawSample>) return B200MIX_FMT_MULAW;
        else if constexpr(std::is_same_v<T,AlawSample>) return B200MIX_FMT_ALAW;
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
    if(std::holds_alternative<HrtfPostProcess>(device->mPostProcess)) d.post_process = B200MIX_POST_HRTF;
    else if(std::holds_alternative<AmbiDecPostProcess>(device->mPostProcess)) d.post_process = B200MIX_POST_AMBIDEC;
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
    else if(auto *aproc = std::get_if<AmbiDecPostProcess>(&device->mPostProcess))
    {
        auto &dec = *aproc->mAmbiDecoder;
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
    S.cache.assign(kMaxVoices, VoiceCache{});
    S.results.assign(kMaxVoices, b200mix_voice_result{});
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
    if(!S.dev && !open_device(device, S)) return;
    const uint32_t ir = S.desc.ir_size, cd = S.desc.dry_channels;

    /* ---- the voices of every context, in mixing order ---- */
    S.vptr.clear(); S.vctx.clear();
    for(ContextBase *ctx : *device->mContexts.load(std::memory_order_acquire))
    {
        if(auto *arr = ctx->mActiveAuxSlots.load(std::memory_order_acquire); arr && !arr->empty())
        { fail(device, S, "auxiliary effect slots are not wired into the seam yet"); return; }
        for(Voice *voice : ctx->getVoicesSpanAcquired()) { S.vptr.push_back(voice); S.vctx.push_back(ctx); }
    }
    if(S.vptr.size() > kMaxVoices) { fail(device, S, "more voices than the seam's device was created for"); return; }

    S.upd.clear(); S.upd_coeffs.clear(); S.upd_dry.clear();
    for(size_t n = 0;n < S.vptr.size();++n)
    {
        Voice *voice = S.vptr[n];
        VoiceCache &C = S.cache[n];
        const auto pstate = voice->mPlayState.load(std::memory_order_acquire);
        const bool active = pstate == Voice::Playing || pstate == Voice::Stopping;
        if(!active)
        {
            if(C.live)
            {   /* the host stopped it (alSourceStop / rewind): remove it from the active set */
                b200mix_voice_params p = C.params;
                p.flags = B200MIX_VF_STOPPED;
                S.upd.push_back(p);
                S.upd_coeffs.insert(S.upd_coeffs.end(), size_t(ir)*2, 0.0f);
                S.upd_dry.insert(S.upd_dry.end(), cd, 0.0f);
                C.live = false;
            }
            continue;
        }
        if(!voice->mFlags.test(VoiceFlag::IsStatic) || voice->mFlags.test(VoiceFlag::IsCallback)
            || voice->mFmtChannels != FmtMono || voice->mDuplicateMono || voice->mDirect.FilterActive)
        { fail(device, S, "streaming / multi-channel / filtered sources are not wired into the seam yet"); return; }
        auto *item = voice->mCurrentBuffer.load(std::memory_order_relaxed);
        auto *loop = voice->mLoopBuffer.load(std::memory_order_relaxed);
        auto &ch = voice->mChans[0];

        b200mix_voice_params p{};
        p.voice = static_cast<uint32_t>(n);
        p.flags = (pstate == Voice::Playing ? B200MIX_VF_PLAYING : B200MIX_VF_STOPPING) | B200MIX_VF_STATIC;
        if(loop) p.flags |= B200MIX_VF_LOOPING;
        if(voice->mFlags.test(VoiceFlag::HasHrtf)) p.flags |= B200MIX_VF_HRTF;
        p.resampler = static_cast<uint32_t>(voice->mProps.mResampler);
        p.step = voice->mStep;
        p.hrtf_delay[0] = ch.mDryParams.Hrtf.Target.Delay[0];
        p.hrtf_delay[1] = ch.mDryParams.Hrtf.Target.Delay[1];
        p.hrtf_gain = ch.mDryParams.Hrtf.Target.Gain;
        for(auto &s : p.send_slot) s = B200MIX_NO_SLOT;
        if(item)
        {
            const void *data = nullptr;
            const int type = sample_type_of(item->mSamples, &data);
            if(type < 0 || voice->mFrameStep != 1u) { fail(device, S, "buffer format not wired into the seam yet"); return; }
            auto it = S.buffers.find(data);
            if(it == S.buffers.end() || it->second.second != item->mSampleLen)
            {
                /* BufferStorage is immutable while attached: one upload per (data, length) */
                static const size_t sz[] = {1, 2, 4, 4, 8, 1, 1};
                const uint32_t id = it == S.buffers.end() ? S.next_buffer++ : it->second.first;
                if(id >= kMaxBuffers) { fail(device, S, "more buffers than the seam's device was created for"); return; }
                if(A.buffer_data(S.dev, id, uint32_t(type), 1u, item->mSampleLen, data,
                    size_t(item->mSampleLen)*sz[type]) != B200MIX_OK)
                { fail(device, S, "b200mix_buffer_data failed:"); return; }
                it = S.buffers.insert_or_assign(data, std::make_pair(id, item->mSampleLen)).first;
            }
            p.buffer = it->second.first;
            p.loop_start = item->mLoopStart; p.loop_end = item->mLoopEnd;
        }
        else p.buffer = B200MIX_NO_BUFFER;      /* alSourceStop / rewind took the buffer (alc/alu.cpp:2071) */
        const unsigned sid = voice->mSourceID.load(std::memory_order_relaxed);
        /* a stop clears mSourceID while the voice fades out: that is not a new voice */
        const bool fresh = !C.live || (sid != 0u && C.source_id != sid);
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
        bool changed = fresh || std::memcmp(&C.params, &p, sizeof(p)) != 0;
        if(!changed && hrtf && ir) changed = std::memcmp(C.coeffs.data(), co, size_t(ir)*2*sizeof(float)) != 0;
        if(!changed && !hrtf) changed = std::memcmp(C.dry.data(), dg, cd*sizeof(float)) != 0;
        if(changed)
        {
            S.upd.push_back(p);
            S.upd_coeffs.insert(S.upd_coeffs.end(), co, co + size_t(ir)*2);
            S.upd_dry.insert(S.upd_dry.end(), dg, dg + cd);
            b200mix_voice_params keep = p;
            keep.flags &= ~uint32_t(B200MIX_VF_RESET | B200MIX_VF_FADING);
            keep.position = 0; keep.position_frac = 0;
            C.params = keep;
            C.coeffs.assign(co, co + size_t(ir)*2);
            C.dry.assign(dg, dg + cd);
        }
        C.live = true; C.source_id = sid;
    }
    if(!S.upd.empty()
        && A.voices_update(S.dev, uint32_t(S.upd.size()), S.upd.data(), ir ? S.upd_coeffs.data() : nullptr,
            S.upd_dry.data(), nullptr) != B200MIX_OK)
    { fail(device, S, "b200mix_voices_update failed:"); return; }

    /* ---- the update itself: RealOut comes back planar, where Limiter / Write<T> expect it ---- */
    std::array<float*, B200MIX_MAX_DRY_CHANNELS> outs{};
    for(size_t c = 0;c < device->RealOut.Buffer.size();++c) outs[c] = device->RealOut.Buffer[c].data();
    if(A.render(S.dev, samplesToDo, outs.data(), S.results.data()) != B200MIX_OK)
    { fail(device, S, "b200mix_render failed:"); return; }

    /* ---- cursor and play state back into the Voice objects (core/voice.cpp:1116-1232) ---- */
    for(size_t n = 0;n < S.vptr.size();++n)
    {
        Voice *voice = S.vptr[n];
        VoiceCache &C = S.cache[n];
        if(!C.live) continue;
        const b200mix_voice_result &r = S.results[n];
        voice->mPosition.store(r.position, std::memory_order_relaxed);
        voice->mPositionFrac.store(r.position_frac, std::memory_order_relaxed);
        voice->mFlags.set(VoiceFlag::IsFading);
        if(r.flags & B200MIX_VF_STOPPED)
        {
            voice->mCurrentBuffer.store(nullptr, std::memory_order_relaxed);
            voice->mLoopBuffer.store(nullptr, std::memory_order_relaxed);
            voice->mSourceID.store(0u, std::memory_order_relaxed);
            voice->mPlayState.store(Voice::Stopped, std::memory_order_release);
            C.live = false;
        }
        else if((r.flags & B200MIX_VF_STOPPING) && voice->mPlayState.load(std::memory_order_relaxed) == Voice::Playing)
        {
            /* ran out of data: the source reads as stopped from now on, the voice fades for
             * one more update (core/voice.cpp:1198-1232) */
            const unsigned sid = voice->mSourceID.load(std::memory_order_relaxed);
            voice->mCurrentBuffer.store(nullptr, std::memory_order_release);
            voice->mLoopBuffer.store(nullptr, std::memory_order_relaxed);
            voice->mSourceID.store(0u, std::memory_order_release);
            voice->mPlayState.store(Voice::Stopping, std::memory_order_release);
            ContextBase *ctx = S.vctx[n];
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
