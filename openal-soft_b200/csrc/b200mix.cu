// b200mix.cu — host side of the C ABI in include/b200mix.h.
//
// Owns the device-resident mirrors of the reference's mixer state (voices, buffers,
// mix buffers, HRTF accumulator carry) and issues the per-update launch sequence:
//   [k_apply_updates]  k_mix_voices  k_reduce_rows  post-process  (D2H)
// on one CUDA stream.  No CPU mixing path exists: every entry point fails with
// B200MIX_ERR_CUDA when the CUDA runtime/device is unusable.
#include "../../include/b200mix.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include <dlfcn.h>

#include "mixer_kernels.cuh"
#include "effect_kernels.cuh"
#include "shard_kernels.cuh"
#include "panmix_tc.cuh"
#include "param_kernels.hpp"
#include "efx_kernels.hpp"
#include "resampler_tables.hpp"
#include "hrtf_store.hpp"
#include "adpcm.hpp"

using namespace b200mix;

namespace {

thread_local std::string g_create_error;

struct DevBuf {           // cudaMalloc'ed array with size bookkeeping
    void *ptr{nullptr}; size_t bytes{0};
};

} // namespace

struct b200mix_device {
    b200mix_device_desc desc{};
    int cuda_dev{0};
    int num_sms{0};
    cudaStream_t stream{nullptr};
    std::string error;
    uint64_t launches{0};

    // tables
    BsincTable bsinc[3];
    float *d_bsinc[3]{};
    float *d_cubic[2]{};

    // state
    VoiceRec *d_voices{nullptr};
    BufferRec *d_buffers{nullptr};
    std::vector<BufferRec> h_buffers;
    std::vector<uint32_t> h_vbuf;                 // static buffer an active voice plays (or NO_SLOT)
    std::vector<uint32_t> h_bufrefs;              // active static voices per buffer
    float2 *d_hrtf_tgt{nullptr}, *d_hrtf_old{nullptr};
    float *d_dry_cur{nullptr}, *d_dry_tgt{nullptr}, *d_send_cur{nullptr}, *d_send_tgt{nullptr};
    VoiceResult *d_results{nullptr};              // inside d_outblock
    b200mix_voice_result *h_results{nullptr};     // inside h_outblock (pinned)
    char *d_outblock{nullptr}, *h_outblock{nullptr}; size_t out_real_bytes{0};

    // mix buffers
    uint32_t dry_alloc_ch{0};
    float *d_dry{nullptr}, *d_real{nullptr}, *d_wet{nullptr};
    float *d_partial{nullptr}; size_t partial_floats{0};
    float *d_accum_sum{nullptr};                  // [2][kAccumLen]
    float *d_carry[2]{};                          // [2][kHrirLen] ping-pong
    int carry_idx{0};
    float *h_real{nullptr};                       // pinned [real][1024]

    // decoders
    uint32_t dec_channels{0}, dec_ir{0};
    float2 *d_dec_coef{nullptr}; float *d_dec_hfscale{nullptr}, *d_dec_state{nullptr};
    float *d_temp{nullptr}, *d_temp2{nullptr};
    uint32_t amb_in{0}; bool amb_dual{false};
    float *d_amb_hf{nullptr}, *d_amb_lf{nullptr}, *d_amb_state{nullptr};
    bool dry_active{false};
    float *d_uhj_state{nullptr}, *d_uhj_scratch{nullptr};
    uint32_t stab_center{B200MIX_NO_SLOT};                // StablizerPostProcess: FrontCenter index
    float stab_coeff{0.0f};
    float *d_stab_state{nullptr};                         // [0..2] MidFilter, [4+i] ChannelFilters[i].mApZ1
    uint32_t bs2b_level{0};                               // Bs2bPostProcess: 0 = off
    float *d_bs2b{nullptr};                               // [0..3] history, [4..8] coefficients
    uint32_t uhj_fir{0};                                  // 0 = IIR, 256 / 512 = UhjEncoder<N>
    float *d_uhj_fir_state{nullptr}, *d_uhj_fir_coef{nullptr};

    // update staging (pinned host + device)
    char *h_arena{nullptr}, *d_arena{nullptr};    // staging arena (see ensure_stage)
    VoiceUpdate *h_upd{nullptr};                  // == h_arena
    uint32_t stage_cap{0};
    cudaEvent_t stage_done{nullptr};
    bool stage_busy{false};

    // aux sends and effect slots
    std::vector<SlotRec> h_slots;            // host mirror (device pointers inside)
    SlotRec *d_slots{nullptr};
    std::vector<std::vector<void*>> slot_allocs;
    uint32_t active_slots{0};
    float *d_xscratch{nullptr};
    uint32_t *d_sendinfo{nullptr};
    // direct/send filters (allocated by the first b200mix_voices_filters)
    FilterRec *d_filt{nullptr};
    FilterUpdate *h_fupd{nullptr}, *d_fupd{nullptr};
    uint32_t fupd_cap{0};
    cudaEvent_t fstage_done{nullptr};
    bool fstage_busy{false};
    float *d_fscratch{nullptr};
    uint32_t fscratch_rows{0};
    float *d_dline{nullptr};                 // [max_voices][1024] filtered direct-path lines
    std::vector<uint8_t> h_dfilt;            // host mirror: direct filter active per voice
    std::vector<uint32_t> h_order2;          // active voices with an active direct filter
    uint32_t *d_order2{nullptr};
    uint32_t num_order2{0};
    bool order2_dirty{false};
    std::vector<uint32_t> h_send_slot;       // [max_voices][MAX_SENDS] host mirror
    std::vector<uint32_t> h_slot_start;
    std::vector<SendEntry> h_entries;
    uint32_t *d_slot_start{nullptr};
    SendEntry *d_entries{nullptr};
    uint32_t num_entries{0};
    bool sends_dirty{true};
    // EFX effect slots (b200mix_slot_efx): host mirrors + the per-slot views the kernel walks
    struct EfxHost { bool used{false}; EfxParams p{}; EfxDev *dev{nullptr}; uint32_t mod_index{0}, mod_range{1};
        uint32_t lfo_offset{0}, lfo_range{1}; };
    std::vector<EfxHost> efx;
    EfxSlotView *d_efx_views{nullptr};
    uint32_t efx_slots{0};
    uint32_t pshift_slots{0};                // EFX slots running the pitch shifter (k_efx_pshift)
    bool efx_ready{false};
    float2 *d_twiddle{nullptr};
    float *d_cubic_filter{nullptr};          // gCubicTable (reverb modulation taps)
    uint32_t reverb_slots{0};

    // attached HRTF data set (device-side HrtfStore::getCoeffs)
    float2 *d_st_fields{nullptr}; uint2 *d_st_elevs{nullptr}; float2 *d_st_coeffs{nullptr};
    uint8_t *d_st_delays{nullptr}; uint32_t st_num_fields{0}, st_ir{0};
    uint4 *d_qhdr{nullptr}; uint32_t *d_queue{nullptr};   // streaming queues (first b200mix_voice_queue)
    LimiterDev *d_limiter{nullptr};                      // DeviceBase::Limiter (b200mix_set_limiter)
    float *d_limiter_delay{nullptr};                     // Compressor::mDelay [real_channels][1024]
    uint32_t *d_dc_delay{nullptr}; float *d_dc_gain{nullptr}, *d_dc_buf{nullptr};   // DeviceBase::ChannelDelays
    void *d_outbuf{nullptr}, *h_outbuf{nullptr};          // interleaved output staging (render_interleaved)
    // reverb slots: host side of ReverbState's two-pipeline state machine
    struct RvHost {
        bool used{false};
        int cur{0}, state{4};                   // PipelineState: 1 StartFade, 2 Fading, 3 Cleanup, 4 Normal
        uint32_t fade[2]{1u, 1u};               // mFadeSampleCount per pipeline object
        uint32_t offset{0};                     // mOffset
        ReverbDev h[2];                         // host mirrors (parameters + pointers)
        ReverbDev *dev{nullptr};                // device array [2]
    };
    std::vector<RvHost> rv;
    std::vector<uint32_t> h_target;          // EffectSlotBase::Target per slot (NO_SLOT = Dry)
    uint32_t num_stages{1}; bool any_target{false};
    bool reverb_upmix{false};                // some reverb slot uses MixOutAmbiUp
    bool mid_render{false}; uint32_t mid_frames{0};   // between render_begin and render_end
    bool real_overwrite{false};
    // parked dry bus (kernel variants without register dry accumulators)
    std::vector<uint8_t> h_hrtf;             // host mirror: voice mixes through its own HRIR
    std::vector<SendEntry> h_dry_entries;
    SendEntry *d_dry_entries{nullptr};
    uint32_t *d_dry_slot_start{nullptr};
    uint32_t num_dry_entries{0};
    bool dry_entries_dirty{true};
    float *d_dry_partial{nullptr};           // [kDryChunksMax][cd][1024]
    float *d_dry_geff{nullptr};              // [max_voices][cd]
    float *d_send_geff{nullptr};             // [max_voices*num_sends][cw]
    float4 *d_dry_gramp{nullptr}, *d_send_gramp{nullptr};
    float *d_send_partial{nullptr};          // [send_chunks][max_slots][cw][1024]
    uint32_t send_partial_chunks{0}, max_slot_entries{0};
    bool profile{false};
    cudaEvent_t ev_mix0{nullptr}, ev_mix1{nullptr};
    bool ev_valid{false};
    // stage marks of the last update (profile >= 2): see b200mix_last_stage_ms
    static constexpr int kStages = 8;
    cudaEvent_t ev_stage[kStages + 1]{};
    bool stage_valid{false}; int profile_level{0};

    // mixing order (host mirror of which voices are configured active, and their cost)
    std::vector<uint8_t> h_active;
    std::vector<uint32_t> h_cost;
    std::vector<uint32_t> h_order;
    uint32_t *d_order{nullptr};
    uint32_t num_order{0};
    bool order_dirty{true};

    // GPU parameter stage (b200mix_sources_update): pinned input staging + device scratch
    char *h_src{nullptr}, *d_src{nullptr}; uint32_t src_cap{0};
    cudaEvent_t src_done{nullptr}; bool src_busy{false};
    bool dev_filters{false};                 // filter activity is decided on the device: order2 = order
    bool panmix_tc{false};                   // wide dry buses: pan-mix past the fades on the tensor cores
    bool mix_gather_only{false};             // B200MIX_MIX_GATHER=1: no TMA staging in k_mix_voices (A/B runs)

    // voice-sharded device set (b200mix_shard_*): transport 0 none, 1 peer stores, 2 NCCL
    struct Shard {
        uint32_t rank{0}, world{1}; int transport{0};
        char *own{nullptr}; size_t bytes{0};
        char *peer[kShardMaxWorld]{};
        size_t off_real{0}, off_wet{0}, real_floats{0}, wet_src_floats{0};
        uint32_t owned_max{0};
        uint32_t epoch{0};
        uint32_t *d_counters{nullptr};         // [0] real push, [1] real sum, [2] wet sum, [4..] wet push per owner
        uint32_t *h_status{nullptr};           // pinned copy of ShardCtl::status
        cudaEvent_t ev[4]{};                   // wet exchange begin/end, RealOut reduce begin/end
        bool ev_wet{false}, ev_real{false};
        // NCCL transport (dlopen'ed: the library carries no link-time NCCL dependency)
        void *nccl_lib{nullptr}; void *comm{nullptr};
        int (*reduce)(const void*, void*, size_t, int, int, int, void*, cudaStream_t){nullptr};
        int (*allreduce)(const void*, void*, size_t, int, int, void*, cudaStream_t){nullptr};
        int (*comm_destroy)(void*){nullptr};
    } shard;

    uint32_t ir_pad{0};
    uint32_t voice_hi{0};          // 1 + highest voice index ever configured
    // launch geometry (resolved at create)
    int mix_variant{0}; int mix_groups{2}; int mix_gs{64}; int mix_cdr{0};
    uint32_t reverb_seq{0};          // update counter of k_reverb_process' early/late hand-off (24 bits used)
    size_t mix_smem{0}; int mix_blocks_per_sm{1};
};

namespace {

#define CUDA_TRY(dev, expr) do { cudaError_t e_ = (expr); if(e_ != cudaSuccess) {            \
    (dev)->error = std::string(#expr) + ": " + cudaGetErrorString(e_); return B200MIX_ERR_CUDA; } } while(0)

template<typename T>
int dev_alloc(b200mix_device *d, T *&p, size_t count, bool zero = true)
{
    p = nullptr;
    if(count == 0) return B200MIX_OK;
    CUDA_TRY(d, cudaMalloc(reinterpret_cast<void**>(&p), count*sizeof(T)));
    if(zero) CUDA_TRY(d, cudaMemsetAsync(p, 0, count*sizeof(T), d->stream));
    return B200MIX_OK;
}

using MixKernel = void(*)(const MixParams);

struct Variant { MixKernel fn; int gs, groups, cdr; size_t smem; bool hrtf; };

template<int GS, int GROUPS, bool HRTF, int CDR, int OPT, int FP>
Variant make_variant()
{
    return Variant{k_mix_voices<GS, GROUPS, HRTF, CDR, OPT, FP>, GS, GROUPS, CDR,
        sizeof(GroupSmem<GS, OPT, FP>)*GROUPS, HRTF};
}

// 0: HRTF ir<=64, 1: HRTF ir<=128, 2: dry <=4 channels in registers,
// 3: wider dry mixes: resample + park, the dry bus is summed by k_send_mix
Variant get_variant(int idx)
{
    switch(idx)
    {
    case 0: return make_variant<64, 2, true, 0, 17, 64>();
    case 1: return make_variant<64, 2, true, 0, 19, 128>();
    case 2: return make_variant<64, 2, false, 4, 1, 8>();
    default: return make_variant<64, 2, false, 0, 1, 8>();
    }
}

constexpr uint32_t kDryChunksMax = 128;

// Storage of the parked dry bus (variants with CDR == 0 that meet a non-HRTF voice).
int ensure_dry_park(b200mix_device *d)
{
    const b200mix_device_desc &dd = d->desc;
    if(d->d_dry_entries) return B200MIX_OK;
    if(!d->d_xscratch)
        if(int rc = dev_alloc(d, d->d_xscratch, size_t(dd.max_voices)*kLine)) return rc;
    if(!d->d_sendinfo)
        if(int rc = dev_alloc(d, d->d_sendinfo, dd.max_voices)) return rc;
    if(int rc = dev_alloc(d, d->d_dry_entries, dd.max_voices)) return rc;
    if(dd.dry_channels > 4u && dd.dry_channels <= uint32_t(kPmN))
    {
        CUDA_TRY(d, cudaFuncSetAttribute(k_panmix_tc, cudaFuncAttributeMaxDynamicSharedMemorySize,
            kPmStages*kPmStageBytes + 1024));
        // B200MIX_PANMIX_SIMT=1 keeps the whole pan-mix on k_send_mix (A/B measurements only)
        const char *simt = std::getenv("B200MIX_PANMIX_SIMT");
        d->panmix_tc = !(simt && simt[0] == '1');
    }
    if(int rc = dev_alloc(d, d->d_dry_slot_start, 2)) return rc;
    if(int rc = dev_alloc(d, d->d_dry_partial, size_t(kDryChunksMax)*dd.dry_channels*kLine)) return rc;
    if(int rc = dev_alloc(d, d->d_dry_geff, size_t(dd.max_voices)*dd.dry_channels)) return rc;
    if(int rc = dev_alloc(d, d->d_dry_gramp, size_t(dd.max_voices)*dd.dry_channels)) return rc;
    return B200MIX_OK;
}

// One pinned + one device staging arena for b200mix_voices_update: a call packs
// [VoiceUpdate n][coefficients or directions][dry gains][send gains] back to back (16-byte
// aligned parts) and ships them with ONE host-to-device copy.
static size_t align16(size_t v) { return (v + 15u) & ~size_t(15); }

int ensure_stage(b200mix_device *d, uint32_t n)
{
    if(n <= d->stage_cap) return B200MIX_OK;
    const b200mix_device_desc &dd = d->desc;
    if(d->stage_cap)
    {
        cudaStreamSynchronize(d->stream);
        cudaFreeHost(d->h_arena); cudaFree(d->d_arena);
        d->h_arena = nullptr; d->d_arena = nullptr;
    }
    const uint32_t cap = std::max<uint32_t>(n, 256u);
    const size_t bytes = align16(size_t(cap)*sizeof(VoiceUpdate))
        + align16(size_t(cap)*std::max<size_t>(size_t(dd.ir_size)*2, 4)*sizeof(float))
        + align16(size_t(cap)*dd.dry_channels*sizeof(float))
        + align16(size_t(cap)*dd.num_sends*dd.wet_channels*sizeof(float)) + 64;
    CUDA_TRY(d, cudaMallocHost(reinterpret_cast<void**>(&d->h_arena), bytes));
    CUDA_TRY(d, cudaMalloc(reinterpret_cast<void**>(&d->d_arena), bytes));
    d->h_upd = reinterpret_cast<VoiceUpdate*>(d->h_arena);
    d->stage_cap = cap;
    return B200MIX_OK;
}

const BsincTable *bsinc_for(const b200mix_device *d, uint32_t resampler)
{
    if(resampler < B200MIX_RESAMPLER_FAST_BSINC12 || resampler > B200MIX_RESAMPLER_BSINC48)
        return nullptr;
    return &d->bsinc[(resampler - B200MIX_RESAMPLER_FAST_BSINC12) >> 1];
}

} // namespace

extern "C" {

static void shard_release(b200mix_device *d);
static int ensure_filters(b200mix_device *d);

uint32_t b200mix_version(void) { return (1u<<16) | 1u; }

const char *b200mix_last_error(const b200mix_device *dev)
{ return dev ? dev->error.c_str() : g_create_error.c_str(); }

int b200mix_create(const b200mix_device_desc *desc, b200mix_device **out)
{
    if(!desc || !out || desc->struct_size != sizeof(b200mix_device_desc))
    { g_create_error = "bad descriptor"; return B200MIX_ERR_INVALID; }
    if(desc->dry_channels > B200MIX_MAX_DRY_CHANNELS || desc->wet_channels > B200MIX_MAX_WET_CHANNELS
        || desc->num_sends > B200MIX_MAX_SENDS || desc->ir_size > B200MIX_HRIR_LENGTH
        || desc->real_channels > B200MIX_MAX_DRY_CHANNELS || desc->max_voices == 0)
    { g_create_error = "descriptor out of range"; return B200MIX_ERR_INVALID; }
    if((desc->post_process == B200MIX_POST_UHJ && desc->dry_channels < 3)
        || (desc->post_process == B200MIX_POST_TSME && desc->dry_channels < 4))
    { g_create_error = "UHJ post-process needs W,X,Y dry channels"; return B200MIX_ERR_INVALID; }
    if((desc->post_process == B200MIX_POST_HRTF || desc->post_process == B200MIX_POST_UHJ
        || desc->post_process == B200MIX_POST_TSME)
        && (desc->real_left >= desc->real_channels || desc->real_right >= desc->real_channels))
    { g_create_error = "real_left/real_right outside RealOut"; return B200MIX_ERR_INVALID; }

    auto *d = new(std::nothrow) b200mix_device{};
    if(!d) { g_create_error = "out of host memory"; return B200MIX_ERR_NOMEM; }
    d->desc = *desc;
    if(const char *g = std::getenv("B200MIX_MIX_GATHER")) d->mix_gather_only = g[0] == '1';
    auto fail = [&](int code) { g_create_error = d->error; b200mix_destroy(d); return code; };

    int count = 0;
    if(cudaGetDeviceCount(&count) != cudaSuccess || count < 1)
    { d->error = "no CUDA device: the b200mix mixer has no CPU path"; return fail(B200MIX_ERR_CUDA); }
    if(desc->cuda_device >= 0) d->cuda_dev = desc->cuda_device;
    else if(cudaGetDevice(&d->cuda_dev) != cudaSuccess) d->cuda_dev = 0;
    if(cudaSetDevice(d->cuda_dev) != cudaSuccess)
    { d->error = "cudaSetDevice failed"; return fail(B200MIX_ERR_CUDA); }
    cudaDeviceProp prop{};
    if(cudaGetDeviceProperties(&prop, d->cuda_dev) != cudaSuccess)
    { d->error = "cudaGetDeviceProperties failed"; return fail(B200MIX_ERR_CUDA); }
    d->num_sms = prop.multiProcessorCount;
    if(cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking) != cudaSuccess)
    { d->error = "cudaStreamCreate failed"; return fail(B200MIX_ERR_CUDA); }
    if(cudaEventCreateWithFlags(&d->stage_done, cudaEventDisableTiming) != cudaSuccess)
    { d->error = "cudaEventCreate failed"; return fail(B200MIX_ERR_CUDA); }

    auto run = [&]() -> int {
        // resampler tables (core/bsinc_tables.cpp:150-155)
        d->bsinc[0] = BuildBsincTable(60, 11, 2);
        d->bsinc[1] = BuildBsincTable(60, 23, 2);
        d->bsinc[2] = BuildBsincTable(80, 47, 1);
        for(int i = 0;i < 3;++i)
        {
            if(int rc = dev_alloc(d, d->d_bsinc[i], d->bsinc[i].tab.size(), false)) return rc;
            CUDA_TRY(d, cudaMemcpyAsync(d->d_bsinc[i], d->bsinc[i].tab.data(),
                d->bsinc[i].tab.size()*sizeof(float), cudaMemcpyHostToDevice, d->stream));
        }
        const std::vector<float> cubic[2] = {BuildSplineTable(), BuildGaussianTable()};
        for(int i = 0;i < 2;++i)
        {
            if(int rc = dev_alloc(d, d->d_cubic[i], cubic[i].size(), false)) return rc;
            CUDA_TRY(d, cudaMemcpyAsync(d->d_cubic[i], cubic[i].data(), cubic[i].size()*sizeof(float),
                cudaMemcpyHostToDevice, d->stream));
        }
        CUDA_TRY(d, cudaStreamSynchronize(d->stream));   // cubic[] goes out of scope

        const b200mix_device_desc &dd = d->desc;
        d->ir_pad = (dd.ir_size + 7u) & ~7u;
        if(int rc = dev_alloc(d, d->d_voices, dd.max_voices)) return rc;
        if(int rc = dev_alloc(d, d->d_buffers, std::max(dd.max_buffers, 1u))) return rc;
        d->h_buffers.assign(std::max(dd.max_buffers, 1u), BufferRec{});
        d->h_vbuf.assign(dd.max_voices, B200MIX_NO_SLOT);
        d->h_bufrefs.assign(std::max(dd.max_buffers, 1u), 0u);
        if(dd.ir_size)
        {
            if(int rc = dev_alloc(d, d->d_hrtf_tgt, size_t(dd.max_voices)*d->ir_pad)) return rc;
            if(int rc = dev_alloc(d, d->d_hrtf_old, size_t(dd.max_voices)*d->ir_pad)) return rc;
        }
        if(int rc = dev_alloc(d, d->d_dry_cur, size_t(dd.max_voices)*std::max(dd.dry_channels, 1u))) return rc;
        if(int rc = dev_alloc(d, d->d_dry_tgt, size_t(dd.max_voices)*std::max(dd.dry_channels, 1u))) return rc;
        if(dd.num_sends && dd.wet_channels)
        {
            const size_t per = size_t(dd.num_sends)*dd.wet_channels;
            if(int rc = dev_alloc(d, d->d_send_cur, dd.max_voices*per)) return rc;
            if(int rc = dev_alloc(d, d->d_send_tgt, dd.max_voices*per)) return rc;
        }
        if(int rc = dev_alloc(d, d->d_order, dd.max_voices)) return rc;
        d->h_active.assign(dd.max_voices, 0);
        d->h_cost.assign(dd.max_voices, 0);
        d->h_hrtf.assign(dd.max_voices, 0);

        // launch variant
        const bool hrtfDev = dd.ir_size > 0;
        d->mix_variant = hrtfDev ? (dd.ir_size <= 64 ? 0 : 1) : (dd.dry_channels <= 4 ? 2 : 3);
        const Variant var = get_variant(d->mix_variant);
        d->mix_gs = var.gs; d->mix_groups = var.groups; d->mix_cdr = var.cdr; d->mix_smem = var.smem;
        CUDA_TRY(d, cudaFuncSetAttribute(var.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, int(var.smem)));
        int perSm = 0;
        CUDA_TRY(d, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, var.fn, var.gs*var.groups, var.smem));
        d->mix_blocks_per_sm = std::max(perSm, 1);

        d->dry_alloc_ch = std::max<uint32_t>(std::max(dd.dry_channels, 1u), uint32_t(var.cdr));
        if(int rc = dev_alloc(d, d->d_dry, size_t(d->dry_alloc_ch)*kLine)) return rc;
        // RealOut and the voice results share one block (and one pinned mirror): a render that
        // returns both needs ONE device-to-host copy
        {
            const size_t realFloats = size_t(std::max(dd.real_channels, 1u))*kLine;
            static_assert(sizeof(VoiceResult) == 16 && sizeof(b200mix_voice_result) == 16, "result layout");
            const size_t blockBytes = realFloats*sizeof(float) + size_t(dd.max_voices)*sizeof(VoiceResult);
            char *blk = nullptr;
            if(int rc = dev_alloc(d, blk, blockBytes)) return rc;
            d->d_outblock = blk;
            d->d_results = reinterpret_cast<VoiceResult*>(blk + realFloats*sizeof(float));
            CUDA_TRY(d, cudaMallocHost(reinterpret_cast<void**>(&d->h_outblock), blockBytes));
            d->h_real = reinterpret_cast<float*>(d->h_outblock);
            d->h_results = reinterpret_cast<b200mix_voice_result*>(d->h_outblock + realFloats*sizeof(float));
            d->out_real_bytes = realFloats*sizeof(float);
            if(dd.post_process == B200MIX_POST_NONE) d->d_real = d->d_dry;
            else d->d_real = reinterpret_cast<float*>(blk);
        }
        if(dd.max_slots && dd.wet_channels)
            if(int rc = dev_alloc(d, d->d_wet, size_t(dd.max_slots)*dd.wet_channels*kLine)) return rc;
        const size_t maxRows = size_t(d->num_sms)*d->mix_blocks_per_sm;
        d->partial_floats = maxRows*(hrtfDev ? 2*kAccumLen : 0) + maxRows*size_t(var.cdr)*kLine;
        // two regions: the main pass and the deferred pass of voices with direct filters
        if(int rc = dev_alloc(d, d->d_partial, std::max<size_t>(2*d->partial_floats, 4))) return rc;
        if(int rc = dev_alloc(d, d->d_accum_sum, 2*kAccumLen)) return rc;
        if(int rc = dev_alloc(d, d->d_carry[0], 2*kHrirLen)) return rc;
        if(int rc = dev_alloc(d, d->d_carry[1], 2*kHrirLen)) return rc;
        if(int rc = dev_alloc(d, d->d_temp, size_t(std::max(dd.dry_channels, 1u))*kLine)) return rc;
        if(int rc = dev_alloc(d, d->d_temp2, size_t(std::max(dd.dry_channels, 1u))*kLine)) return rc;
        if(dd.max_slots && dd.wet_channels && dd.num_sends)
        {
            d->h_slots.assign(dd.max_slots, SlotRec{});
            d->slot_allocs.assign(dd.max_slots, {});
            if(int rc = dev_alloc(d, d->d_slots, dd.max_slots)) return rc;
            if(int rc = dev_alloc(d, d->d_xscratch, size_t(dd.max_voices)*kLine)) return rc;
            if(int rc = dev_alloc(d, d->d_sendinfo, dd.max_voices)) return rc;
            if(int rc = dev_alloc(d, d->d_send_geff, size_t(dd.max_voices)*dd.num_sends*dd.wet_channels)) return rc;
            if(int rc = dev_alloc(d, d->d_send_gramp, size_t(dd.max_voices)*dd.num_sends*dd.wet_channels)) return rc;
            d->h_send_slot.assign(size_t(dd.max_voices)*B200MIX_MAX_SENDS, B200MIX_NO_SLOT);
            if(int rc = dev_alloc(d, d->d_slot_start, dd.max_slots + 1)) return rc;
            if(int rc = dev_alloc(d, d->d_entries, size_t(dd.max_voices)*dd.num_sends)) return rc;
            std::vector<float2> tw(128);
            for(int k = 0;k < 128;++k)
            {
                const double a = -2.0*3.14159265358979323846*double(k)/256.0;
                tw[k] = make_float2(float(std::cos(a)), float(std::sin(a)));
            }
            const std::vector<float> cf = BuildCubicFilter();
            if(int rc = dev_alloc(d, d->d_cubic_filter, cf.size(), false)) return rc;
            CUDA_TRY(d, cudaMemcpyAsync(d->d_cubic_filter, cf.data(), cf.size()*sizeof(float),
                cudaMemcpyHostToDevice, d->stream));
            CUDA_TRY(d, cudaStreamSynchronize(d->stream));
            if(int rc = dev_alloc(d, d->d_twiddle, 128, false)) return rc;
            CUDA_TRY(d, cudaMemcpyAsync(d->d_twiddle, tw.data(), 128*sizeof(float2),
                cudaMemcpyHostToDevice, d->stream));
            CUDA_TRY(d, cudaStreamSynchronize(d->stream));
        }
        if(dd.post_process == B200MIX_POST_UHJ || dd.post_process == B200MIX_POST_TSME)
        {
            if(int rc = dev_alloc(d, d->d_uhj_state, 64)) return rc;
            if(int rc = dev_alloc(d, d->d_uhj_scratch, 5*1025)) return rc;
        }
        if(int rc = ensure_stage(d, std::min(dd.max_voices, 4096u))) return rc;
        CUDA_TRY(d, cudaStreamSynchronize(d->stream));
        return B200MIX_OK;
    };
    int rc;
    try { rc = run(); }
    catch(const std::exception &e) { d->error = e.what(); rc = B200MIX_ERR_NOMEM; }
    if(rc != B200MIX_OK) return fail(rc);
    *out = d;
    return B200MIX_OK;
}

void b200mix_destroy(b200mix_device *d)
{
    if(!d) return;
    if(d->stream) cudaStreamSynchronize(d->stream);
    shard_release(d);
    for(auto &b : d->h_buffers) if(b.data) cudaFree(const_cast<void*>(b.data));
    for(int i = 0;i < 3;++i) cudaFree(d->d_bsinc[i]);
    for(int i = 0;i < 2;++i) cudaFree(d->d_cubic[i]);
    cudaFree(d->d_voices); cudaFree(d->d_buffers); cudaFree(d->d_hrtf_tgt); cudaFree(d->d_hrtf_old);
    cudaFree(d->d_dry_cur); cudaFree(d->d_dry_tgt); cudaFree(d->d_send_cur); cudaFree(d->d_send_tgt);
    cudaFree(d->d_outblock); cudaFreeHost(d->h_outblock); cudaFree(d->d_order);
    cudaFree(d->d_dry); cudaFree(d->d_wet); cudaFree(d->d_partial); cudaFree(d->d_accum_sum);
    cudaFree(d->d_carry[0]); cudaFree(d->d_carry[1]);
    cudaFree(d->d_dec_coef); cudaFree(d->d_dec_hfscale); cudaFree(d->d_dec_state);
    cudaFree(d->d_temp); cudaFree(d->d_temp2);
    cudaFree(d->d_amb_hf); cudaFree(d->d_amb_lf); cudaFree(d->d_amb_state);
    cudaFree(d->d_uhj_state); cudaFree(d->d_uhj_scratch);
    cudaFree(d->d_uhj_fir_state); cudaFree(d->d_uhj_fir_coef); cudaFree(d->d_bs2b); cudaFree(d->d_stab_state);
    for(auto &v : d->slot_allocs) for(void *p : v) cudaFree(p);
    cudaFree(d->d_slots); cudaFree(d->d_xscratch); cudaFree(d->d_sendinfo);
    cudaFree(d->d_filt); cudaFree(d->d_fupd); cudaFree(d->d_fscratch);
    cudaFree(d->d_dline); cudaFree(d->d_order2); cudaFree(d->d_qhdr); cudaFree(d->d_queue);
    cudaFree(d->d_outbuf); if(d->h_outbuf) cudaFreeHost(d->h_outbuf);
    cudaFree(d->d_limiter); cudaFree(d->d_limiter_delay);
    cudaFree(d->d_dc_delay); cudaFree(d->d_dc_gain); cudaFree(d->d_dc_buf);
    cudaFree(d->d_st_fields); cudaFree(d->d_st_elevs); cudaFree(d->d_st_coeffs); cudaFree(d->d_st_delays);
    cudaFree(d->d_dry_entries); cudaFree(d->d_dry_slot_start); cudaFree(d->d_dry_partial);
    cudaFree(d->d_dry_geff); cudaFree(d->d_send_geff); cudaFree(d->d_send_partial);
    cudaFree(d->d_dry_gramp); cudaFree(d->d_send_gramp);
    if(d->h_fupd) cudaFreeHost(d->h_fupd);
    if(d->fstage_done) cudaEventDestroy(d->fstage_done);
    cudaFree(d->d_slot_start); cudaFree(d->d_entries); cudaFree(d->d_twiddle); cudaFree(d->d_cubic_filter);
    cudaFree(d->d_efx_views);
    cudaFreeHost(d->h_arena); cudaFree(d->d_arena);
    if(d->h_src) cudaFreeHost(d->h_src);
    cudaFree(d->d_src);
    if(d->src_done) cudaEventDestroy(d->src_done);
    if(d->stage_done) cudaEventDestroy(d->stage_done);
    if(d->ev_mix0) cudaEventDestroy(d->ev_mix0);
    if(d->ev_mix1) cudaEventDestroy(d->ev_mix1);
    for(cudaEvent_t e : d->ev_stage) if(e) cudaEventDestroy(e);
    if(d->stream) cudaStreamDestroy(d->stream);
    delete d;
}

int b200mix_set_hrtf_decoder(b200mix_device *d, uint32_t channels, uint32_t ir_size,
    const float *coeffs, const float *hf_scale, const float *splitter_coeff)
{
    if(!d || channels != d->desc.dry_channels || ir_size > B200MIX_HRIR_LENGTH || !coeffs
        || !hf_scale || !splitter_coeff || d->desc.post_process != B200MIX_POST_HRTF)
    { if(d) d->error = "set_hrtf_decoder: bad arguments"; return B200MIX_ERR_INVALID; }
    CUDA_TRY(d, cudaSetDevice(d->cuda_dev));
    cudaFree(d->d_dec_coef); cudaFree(d->d_dec_hfscale); cudaFree(d->d_dec_state);
    d->dec_channels = channels; d->dec_ir = ir_size;
    if(int rc = dev_alloc(d, d->d_dec_coef, size_t(channels)*ir_size, false)) return rc;
    if(int rc = dev_alloc(d, d->d_dec_hfscale, channels, false)) return rc;
    if(int rc = dev_alloc(d, d->d_dec_state, size_t(channels)*4)) return rc;
    std::vector<float> st(size_t(channels)*4, 0.0f);
    for(uint32_t c = 0;c < channels;++c) st[c*4] = splitter_coeff[c];
    CUDA_TRY(d, cudaMemcpyAsync(d->d_dec_coef, coeffs, size_t(channels)*ir_size*2*sizeof(float),
        cudaMemcpyHostToDevice, d->stream));
    CUDA_TRY(d, cudaMemcpyAsync(d->d_dec_hfscale, hf_scale, channels*sizeof(float),
        cudaMemcpyHostToDevice, d->stream));
    CUDA_TRY(d, cudaMemcpyAsync(d->d_dec_state, st.data(), st.size()*sizeof(float),
        cudaMemcpyHostToDevice, d->stream));
    CUDA_TRY(d, cudaStreamSynchronize(d->stream));
    return B200MIX_OK;
}

int b200mix_set_ambi_decoder(b200mix_device *d, uint32_t in_channels, const float *gains_hf,
    const float *gains_lf, float xover_coeff)
{
    if(!d || in_channels != d->desc.dry_channels || !gains_hf
        || d->desc.post_process != B200MIX_POST_AMBIDEC)
    { if(d) d->error = "set_ambi_decoder: bad arguments"; return B200MIX_ERR_INVALID; }
    CUDA_TRY(d, cudaSetDevice(d->cuda_dev));
    cudaFree(d->d_amb_hf); cudaFree(d->d_amb_lf); cudaFree(d->d_amb_state);
    d->d_amb_lf = nullptr;
    const size_t n = size_t(in_channels)*d->desc.real_channels;
    d->amb_in = in_channels; d->amb_dual = gains_lf != nullptr;
    if(int rc = dev_alloc(d, d->d_amb_hf, n, false)) return rc;
    CUDA_TRY(d, cudaMemcpyAsync(d->d_amb_hf, gains_hf, n*sizeof(float), cudaMemcpyHostToDevice, d->stream));
    if(gains_lf)
    {
        if(int rc = dev_alloc(d, d->d_amb_lf, n, false)) return rc;
        CUDA_TRY(d, cudaMemcpyAsync(d->d_amb_lf, gains_lf, n*sizeof(float), cudaMemcpyHostToDevice, d->stream));
    }
    std::vector<float> st(size_t(in_channels)*4, 0.0f);
    for(uint32_t c = 0;c < in_channels;++c) st[c*4] = xover_coeff;
    if(int rc = dev_alloc(d, d->d_amb_state, st.size(), false)) return rc;
    CUDA_TRY(d, cudaMemcpyAsync(d->d_amb_state, st.data(), st.size()*sizeof(float),
        cudaMemcpyHostToDevice, d->stream));
    CUDA_TRY(d, cudaStreamSynchronize(d->stream));
    return B200MIX_OK;
}

int b200mix_buffer_data(b200mix_device *d, uint32_t buffer, uint32_t sample_type, uint32_t channels,
    uint32_t frames, const void *data, size_t bytes)
{
    if(!d) return B200MIX_ERR_INVALID;
    static const size_t sz[] = {1, 2, 4, 4, 8, 1, 1};
    if(buffer >= d->desc.max_buffers || sample_type > B200MIX_FMT_ALAW || channels < 1 || !data)
    { d->error = "buffer_data: bad arguments"; return B200MIX_ERR_INVALID; }
    const size_t need = size_t(frames)*channels*sz[sample_type];
    if(bytes < need) { d->error = "buffer_data: short data"; return B200MIX_ERR_INVALID; }
    CUDA_TRY(d, cudaSetDevice(d->cuda_dev));
    BufferRec &h = d->h_buffers[buffer];
    if(h.data && d->h_bufrefs[buffer])
    { d->error = "buffer_data: the buffer is attached to an active voice (AL_INVALID_OPERATION)"; return B200MIX_ERR_INVALID; }
    if(h.data)
    {
        CUDA_TRY(d, cudaStreamSynchronize(d->stream));
        cudaFree(const_cast<void*>(h.data));
        h = BufferRec{};
    }
    void *p = nullptr;
    // +16 bytes so vector/tail reads past the last frame stay inside the allocation
    CUDA_TRY(d, cudaMalloc(&p, need + 16));
    CUDA_TRY(d, cudaMemcpyAsync(p, data, need, cudaMemcpyHostToDevice, d->stream));
    h.data = p; h.frames = frames; h.type = sample_type; h.channels = channels;
    CUDA_TRY(d, cudaMemcpyAsync(d->d_buffers + buffer, &h, sizeof(BufferRec), cudaMemcpyHostToDevice,
        d->stream));
    return B200MIX_OK;
}

int b200mix_buffer_data_adpcm(b200mix_device *d, uint32_t buffer, uint32_t sample_type,
    uint32_t channels, uint32_t samples_per_block, uint32_t blocks, const void *data, size_t bytes)
{
    if(!d) return B200MIX_ERR_INVALID;
    const bool ms = sample_type == B200MIX_FMT_MSADPCM;
    if((sample_type != B200MIX_FMT_IMA4 && !ms) || channels < 1 || channels > 2 || !data
        || !AdpcmBlockValid(ms, samples_per_block))
    { d->error = "buffer_data_adpcm: bad arguments"; return B200MIX_ERR_INVALID; }
    if(bytes < AdpcmBlockBytes(ms, channels, samples_per_block)*blocks
        || uint64_t(blocks)*samples_per_block > 0x7fffffffull)
    { d->error = "buffer_data_adpcm: short data"; return B200MIX_ERR_INVALID; }
    std::vector<int16_t> pcm(size_t(blocks)*samples_per_block*channels);
    if(ms) DecodeMSADPCM(static_cast<const uint8_t*>(data), channels, samples_per_block, blocks, pcm.data());
    else DecodeIMA4(static_cast<const uint8_t*>(data), channels, samples_per_block, blocks, pcm.data());
    const int rc = b200mix_buffer_data(d, buffer, B200MIX_FMT_I16, channels, blocks*samples_per_block,
        pcm.data(), pcm.size()*sizeof(int16_t));
    // the upload above reads pageable memory: make sure it is consumed before pcm goes away
    if(rc == B200MIX_OK) CUDA_TRY(d, cudaStreamSynchronize(d->stream));
    return rc;
}

int b200mix_buffer_free(b200mix_device *d, uint32_t buffer)
{
    if(!d || buffer >= d->desc.max_buffers) return B200MIX_ERR_INVALID;
    CUDA_TRY(d, cudaSetDevice(d->cuda_dev));
    BufferRec &h = d->h_buffers[buffer];
    if(h.data && d->h_bufrefs[buffer])
    { d->error = "buffer_free: the buffer is attached to an active voice; stop the voice first"; return B200MIX_ERR_INVALID; }
    if(h.data)
    {
        CUDA_TRY(d, cudaStreamSynchronize(d->stream));
        cudaFree(const_cast<void*>(h.data));
        h = BufferRec{};
        CUDA_TRY(d, cudaMemcpyAsync(d->d_buffers + buffer, &h, sizeof(BufferRec),
            cudaMemcpyHostToDevice, d->stream));
    }
    return B200MIX_OK;
}

// Processing stages (alc/alu.cpp:2211-2251: every slot before its target): stage = (longest
// chain length) - (hops from the slot to a slot that outputs to Dry); uploads stage/target of
// every slot record.
static int update_stages(b200mix_device *d)
{
    const uint32_t ns = uint32_t(d->h_slots.size());
    if(d->h_target.size() != ns) d->h_target.assign(ns, B200MIX_NO_SLOT);
    std::vector<uint32_t> depth(ns, 0);
    uint32_t maxd = 0; d->any_target = false;
    for(uint32_t sl = 0;sl < ns;++sl)
    {
        uint32_t hops = 0;
        for(uint32_t t = d->h_target[sl];t != B200MIX_NO_SLOT && hops <= ns;t = d->h_target[t]) ++hops;
        depth[sl] = hops;
        if(d->h_slots[sl].type) { maxd = std::max(maxd, hops); if(hops) d->any_target = true; }
    }
    d->num_stages = maxd + 1u;
    for(uint32_t sl = 0;sl < ns;++sl)
    {
        d->h_slots[sl].stage = d->h_slots[sl].type ? maxd - std::min(depth[sl], maxd) : 0u;
        d->h_slots[sl].target = d->h_target[sl];
    }
    if(ns && d->d_slots)
    {
        CUDA_TRY(d, cudaMemcpyAsync(d->d_slots, d->h_slots.data(), ns*sizeof(SlotRec), cudaMemcpyHostToDevice, d->stream));
        CUDA_TRY(d, cudaStreamSynchronize(d->stream));
    }
    if(ns && d->d_efx_views)
    {
        std::vector<EfxSlotView> views(ns, EfxSlotView{nullptr, nullptr, 0u, 0u});
        for(uint32_t sl = 0;sl < ns;++sl)
            if(d->h_slots[sl].type >= B200MIX_EFFECT_ECHO && sl < d->efx.size() && d->efx[sl].used)
                views[sl] = EfxSlotView{d->efx[sl].dev, d->h_slots[sl].lines, d->h_slots[sl].stage, 0u};
        CUDA_TRY(d, cudaMemcpyAsync(d->d_efx_views, views.data(), ns*sizeof(EfxSlotView), cudaMemcpyHostToDevice, d->stream));
        CUDA_TRY(d, cudaStreamSynchronize(d->stream));
    }
    return B200MIX_OK;
}

static void free_slot(b200mix_device *d, uint32_t slot)
{
    cudaStreamSynchronize(d->stream);
    for(void *p : d->slot_allocs[slot]) cudaFree(p);
    d->slot_allocs[slot].clear();
    if(d->h_slots[slot].type) --d->active_slots;
    if(d->h_slots[slot].type == B200MIX_EFFECT_REVERB) --d->reverb_slots;
    if(d->h_slots[slot].type >= B200MIX_EFFECT_ECHO)
    {
        --d->efx_slots;
        if(d->h_slots[slot].type == B200MIX_EFFECT_PSHIFTER) --d->pshift_slots;
        if(slot < d->efx.size()) d->efx[slot] = b200mix_device::EfxHost{};
    }
    if(slot < d->rv.size()) d->rv[slot].used = false;
    d->h_slots[slot] = SlotRec{};
    // the device's record goes with it: an install that fails half-way must not leave the old
    // record pointing at freed lines
    if(d->d_slots)
    {
        cudaMemcpyAsync(d->d_slots + slot, &d->h_slots[slot], sizeof(SlotRec), cudaMemcpyHostToDevice, d->stream);
        cudaStreamSynchronize(d->stream);
    }
}

int b200mix_slot_disable(b200mix_device *d, uint32_t slot)
{
    if(!d || slot >= d->h_slots.size()) { if(d) d->error = "slot_disable: bad slot"; return B200MIX_ERR_INVALID; }
    CUDA_TRY(d, cudaSetDevice(d->cuda_dev));
    free_slot(d, slot);
    return update_stages(d);
}

int b200mix_slot_convolution(b200mix_device *d, uint32_t slot, uint32_t ir_channels,
    uint32_t ir_frames, const float *ir)
{
    if(!d || slot >= d->h_slots.size() || !ir_channels || ir_channels > 16 || !ir_frames || !ir)
    { if(d) d->error = "slot_convolution: bad arguments (or the device has no sends/slots)"; return B200MIX_ERR_INVALID; }
    CUDA_TRY(d, cudaSetDevice(d->cuda_dev));
    free_slot(d, slot);
    CUDA_TRY(d, cudaFuncSetAttribute(k_conv_mac, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sizeof(ConvMacSmem))));
    SlotRec r{};
    r.type = B200MIX_EFFECT_CONVOLUTION; r.channels = ir_channels; r.frames = ir_frames;
    // mNumConvolveSegs (alc/effects/convolution.cpp:375-376)
    const uint32_t nseg = std::max<uint32_t>((ir_frames + kConvBlock - 1)/kConvBlock, 2u) - 1u;
    r.segs = nseg;
    auto alloc = [&](auto *&p, size_t count) -> int {
        if(int rc = dev_alloc(d, p, count)) return rc;
        d->slot_allocs[slot].push_back(p);
        return B200MIX_OK;
    };
    if(int rc = alloc(r.H, size_t(ir_channels)*nseg*kConvFft)) return rc;
    if(int rc = alloc(r.X, size_t(nseg + kConvMaxBlocks)*kConvFft)) return rc;
    if(int rc = alloc(r.head, size_t(ir_channels)*kConvBlock)) return rc;
    if(int rc = alloc(r.inbuf, kConvFft)) return rc;
    if(int rc = alloc(r.ov, size_t(ir_channels)*kConvFft)) return rc;
    if(int rc = alloc(r.yspec, size_t(ir_channels)*kConvMaxChunks*kConvMaxBlocks*kConvFft)) return rc;
    if(int rc = alloc(r.lines, size_t(ir_channels)*kLine)) return rc;
    if(int rc = alloc(r.gains, size_t(2)*ir_channels*32)) return rc;
    if(int rc = alloc(r.gtgt, size_t(ir_channels)*32)) return rc;

    // Filter spectra: segment s holds taps [128(s+1), 128(s+2)), zero padded to 256,
    // transformed in f64 and scaled by 1/256 (convolution.cpp:425-468); layout is ours.
    std::vector<float> H(size_t(ir_channels)*nseg*kConvFft, 0.0f), head(size_t(ir_channels)*kConvBlock, 0.0f);
    std::vector<double> cs(kConvFft), sn(kConvFft);
    for(int k = 0;k < kConvFft;++k)
    {
        cs[k] = std::cos(2.0*3.14159265358979323846*k/kConvFft);
        sn[k] = std::sin(2.0*3.14159265358979323846*k/kConvFft);
    }
    for(uint32_t c = 0;c < ir_channels;++c)
    {
        const float *h = ir + size_t(c)*ir_frames;
        for(uint32_t k = 0;k < uint32_t(kConvBlock) && k < ir_frames;++k) head[c*kConvBlock + k] = h[k];
        for(uint32_t sg = 0;sg < nseg;++sg)
        {
            const size_t base = size_t(kConvBlock)*(sg + 1);
            float *dst = H.data() + (size_t(c)*nseg + sg)*kConvFft;
            const uint32_t cnt = base < ir_frames ? std::min<uint32_t>(kConvBlock, uint32_t(ir_frames - base)) : 0u;
            for(int bin = 0;bin <= kConvBlock;++bin)
            {
                double re = 0.0, im = 0.0;
                for(uint32_t j = 0;j < cnt;++j)
                {
                    const int ph = int((uint64_t(bin)*j) % kConvFft);
                    re += double(h[base + j])*cs[ph];
                    im -= double(h[base + j])*sn[ph];
                }
                const double sc = 1.0/double(kConvFft);
                if(bin == 0) dst[0] = float(re*sc);
                else if(bin == kConvBlock) dst[1] = float(re*sc);
                else { dst[bin*2] = float(re*sc); dst[bin*2+1] = float(im*sc); }
            }
        }
    }
    CUDA_TRY(d, cudaMemcpyAsync(r.H, H.data(), H.size()*sizeof(float), cudaMemcpyHostToDevice, d->stream));
    CUDA_TRY(d, cudaMemcpyAsync(r.head, head.data(), head.size()*sizeof(float), cudaMemcpyHostToDevice, d->stream));
    d->h_slots[slot] = r;
    CUDA_TRY(d, cudaMemcpyAsync(d->d_slots + slot, &d->h_slots[slot], sizeof(SlotRec), cudaMemcpyHostToDevice, d->stream));
    CUDA_TRY(d, cudaStreamSynchronize(d->stream));
    ++d->active_slots;
    d->dry_active = true;
    return update_stages(d);
}

// b200mix_reverb_params -> the parameter part of a ReverbDev (state and pointers untouched)
static void reverb_fill_params(ReverbDev &h, const b200mix_reverb_params *p)
{
    std::memcpy(h.early_tap, p->early_tap, sizeof(h.early_tap)); h.early_tap_coeff = p->early_tap_coeff;
    std::memcpy(h.late_tap, p->late_tap, sizeof(h.late_tap));
    h.mix_x = p->mix_x; h.mix_y = p->mix_y;
    std::memcpy(h.filter_lp, p->filter_lp, sizeof(h.filter_lp));
    std::memcpy(h.filter_hp, p->filter_hp, sizeof(h.filter_hp));
    h.early_ap_coeff = p->early_ap_coeff;
    std::memcpy(h.early_ap_offset, p->early_ap_offset, sizeof(h.early_ap_offset));
    std::memcpy(h.early_offset, p->early_offset, sizeof(h.early_offset));
    h.early_coeff = p->early_coeff;
    std::memcpy(h.late_offset, p->late_offset, sizeof(h.late_offset));
    h.density_gain = p->density_gain;
    std::memcpy(h.t60_mid_gain, p->t60_mid_gain, sizeof(h.t60_mid_gain));
    std::memcpy(h.t60_hf, p->t60_hf, sizeof(h.t60_hf)); std::memcpy(h.t60_lf, p->t60_lf, sizeof(h.t60_lf));
    h.mod_step = p->mod_step; h.mod_depth = p->mod_depth; h.late_ap_coeff = p->late_ap_coeff;
    std::memcpy(h.late_ap_offset, p->late_ap_offset, sizeof(h.late_ap_offset));
    h.upmix = p->upmix ? 1u : 0u; h.order_scale[0] = p->order_scale[0]; h.order_scale[1] = p->order_scale[1];
    h.split_coeff = p->splitter_coeff;
}

static int reverb_check_params(b200mix_device *d, const b200mix_reverb_params *p)
{
    auto pow2 = [](uint32_t v) { return v >= 4u && !(v & (v-1u)); };
    if(!pow2(p->main_len) || !pow2(p->late_in_len) || !pow2(p->early_ap_len) || !pow2(p->early_len)
        || !pow2(p->late_ap_len) || !pow2(p->late_len) || !p->late_offset[0] || !p->late_ap_offset[0])
    { d->error = "slot_reverb: line lengths must be powers of two, feedback delays non-zero"; return B200MIX_ERR_INVALID; }
    for(int j = 0;j < 4;++j)
        if(!p->early_ap_offset[j] || p->late_ap_offset[j] < p->late_ap_offset[0])
        { d->error = "slot_reverb: all-pass delays must be non-zero, late all-pass sorted"; return B200MIX_ERR_INVALID; }
    return B200MIX_OK;
}

// the parameter prefix of ReverbDev (everything before the filter states)
static constexpr size_t kReverbParamBytes = offsetof(ReverbDev, z_lp);

// ReverbPipeline::clear (reverb.cpp:550-566) for one pipeline object: delay lines, filter and
// tap state, the parameters clear() resets, and the object's output gains.
static int reverb_clear_pipeline(b200mix_device *d, uint32_t slot, int obj)
{
    b200mix_device::RvHost &R = d->rv[slot];
    ReverbDev &h = R.h[obj];
    CUDA_TRY(d, cudaMemsetAsync(h.late_in, 0, size_t(4)*h.late_in_len*sizeof(float), d->stream));
    CUDA_TRY(d, cudaMemsetAsync(h.early_ap, 0, size_t(4)*h.early_ap_len*sizeof(float), d->stream));
    CUDA_TRY(d, cudaMemsetAsync(h.early_d, 0, size_t(4)*h.early_len*sizeof(float), d->stream));
    CUDA_TRY(d, cudaMemsetAsync(h.late_ap, 0, size_t(4)*h.late_ap_len*sizeof(float), d->stream));
    CUDA_TRY(d, cudaMemsetAsync(h.late_d, 0, size_t(4)*h.late_len*sizeof(float), d->stream));
    std::memset(h.early_tap, 0, sizeof(h.early_tap)); std::memset(h.late_tap, 0, sizeof(h.late_tap));
    h.early_tap_coeff = 0.0f; h.mod_step = 1u; h.mod_depth = 0.0f;
    std::memset(h.z_lp, 0, sizeof(h.z_lp)); std::memset(h.z_hp, 0, sizeof(h.z_hp));
    std::memset(h.z_t60hf, 0, sizeof(h.z_t60hf)); std::memset(h.z_t60lf, 0, sizeof(h.z_t60lf));
    std::memset(h.z_split, 0, sizeof(h.z_split));
    std::memset(h.early_tap_cur, 0, sizeof(h.early_tap_cur)); std::memset(h.late_tap_cur, 0, sizeof(h.late_tap_cur));
    h.early_coeff_cur = 0.0f; h.mod_index = 0u; h.offset = R.offset;
    CUDA_TRY(d, cudaMemcpyAsync(R.dev + obj, &h, sizeof(ReverbDev), cudaMemcpyHostToDevice, d->stream));
    const SlotRec &S = d->h_slots[slot];
    CUDA_TRY(d, cudaMemsetAsync(S.gtgt + size_t(obj)*8*32, 0, size_t(8)*32*sizeof(float), d->stream));
    for(int sel = 0;sel < 2;++sel)
        CUDA_TRY(d, cudaMemsetAsync(S.gains + (size_t(sel)*16 + size_t(obj)*8)*32, 0, size_t(8)*32*sizeof(float), d->stream));
    // the host mirrors were read by pageable-memory copies above: wait before they change again
    CUDA_TRY(d, cudaStreamSynchronize(d->stream));
    return B200MIX_OK;
}

int b200mix_slot_reverb(b200mix_device *d, uint32_t slot, const b200mix_reverb_params *p)
{
    if(!d || slot >= d->h_slots.size() || !p || p->struct_size != sizeof(*p))
    { if(d) d->error = "slot_reverb: bad arguments (or the device has no sends/slots)"; return B200MIX_ERR_INVALID; }
    if(int rc = reverb_check_params(d, p)) return rc;
    CUDA_TRY(d, cudaSetDevice(d->cuda_dev));
    free_slot(d, slot);
    if(d->rv.size() < d->h_slots.size()) d->rv.resize(d->h_slots.size());
    b200mix_device::RvHost &R = d->rv[slot];
    R = b200mix_device::RvHost{};
    SlotRec r{};
    r.type = B200MIX_EFFECT_REVERB; r.channels = 16;       // 2 pipeline objects x (4 early + 4 late) lines
    auto alloc = [&](auto *&ptr, size_t count) -> int {
        if(int rc = dev_alloc(d, ptr, count)) return rc;
        d->slot_allocs[slot].push_back(ptr);
        return B200MIX_OK;
    };
    float *main_d = nullptr;
    if(int rc = alloc(main_d, size_t(4)*p->main_len)) return rc;
    for(int obj = 0;obj < 2;++obj)
    {
        ReverbDev &h = R.h[obj];
        h = ReverbDev{};
        h.main_len = p->main_len; h.late_in_len = p->late_in_len; h.early_ap_len = p->early_ap_len;
        h.early_len = p->early_len; h.late_ap_len = p->late_ap_len; h.late_len = p->late_len;
        // a pipeline that has not been updated yet is in ReverbPipeline::clear()'s state
        reverb_fill_params(h, p);
        std::memset(h.early_tap, 0, sizeof(h.early_tap)); std::memset(h.late_tap, 0, sizeof(h.late_tap));
        h.early_tap_coeff = 0.0f; h.mod_step = 1u; h.mod_depth = 0.0f;
        h.main_d = main_d;
        if(int rc = alloc(h.late_in, size_t(4)*p->late_in_len)) return rc;
        if(int rc = alloc(h.early_ap, size_t(4)*p->early_ap_len)) return rc;
        if(int rc = alloc(h.early_d, size_t(4)*p->early_len)) return rc;
        if(int rc = alloc(h.late_ap, size_t(4)*p->late_ap_len)) return rc;
        if(int rc = alloc(h.late_d, size_t(4)*p->late_len)) return rc;
    }
    // deviceUpdate leaves DeviceClear; the first update is a full one: it switches to pipeline
    // object 1 and goes straight to Normal (reverb.cpp:1243-1280)
    R.used = true; R.cur = 1; R.state = 4; R.offset = 0;
    if(p->upmix)
    {
        d->reverb_upmix = true;
        CUDA_TRY(d, cudaFuncSetAttribute(k_reverb_upmix, cudaFuncAttributeMaxDynamicSharedMemorySize,
            int(8*kLine*sizeof(float))));
    }
    reverb_fill_params(R.h[1], p);
    R.fade[1] = p->fade_samples; R.fade[0] = 1u;
    if(int rc = alloc(R.dev, 2)) return rc;
    r.H = reinterpret_cast<float*>(R.dev);         // SlotRec::H carries the ReverbDev[2] block
    if(int rc = alloc(r.lines, size_t(16)*kLine)) return rc;
    if(int rc = alloc(r.gains, size_t(2)*16*32)) return rc;
    if(int rc = alloc(r.gtgt, size_t(16)*32)) return rc;
    r.rv_cur = 1u; r.rv_mask = 2u;
    CUDA_TRY(d, cudaMemcpyAsync(R.dev, R.h, 2*sizeof(ReverbDev), cudaMemcpyHostToDevice, d->stream));
    d->h_slots[slot] = r;
    CUDA_TRY(d, cudaMemcpyAsync(d->d_slots + slot, &d->h_slots[slot], sizeof(SlotRec), cudaMemcpyHostToDevice, d->stream));
    CUDA_TRY(d, cudaStreamSynchronize(d->stream));
    ++d->active_slots; ++d->reverb_slots;
    d->dry_active = true;
    return update_stages(d);
}

int b200mix_slot_efx(b200mix_device *d, uint32_t slot, const b200mix_efx_props *props,
    const b200mix_efx_target *target)
{
    if(!d || slot >= d->h_slots.size() || !props || !target || props->struct_size != sizeof(*props)
        || target->struct_size != sizeof(*target) || props->type < B200MIX_EFFECT_ECHO
        || props->type > B200MIX_EFFECT_PSHIFTER)
    { if(d) d->error = "slot_efx: bad arguments (or the device has no sends/slots)"; return B200MIX_ERR_INVALID; }
    const b200mix_device_desc &dd = d->desc;
    const bool toSlot = slot < d->h_target.size() && d->h_target[slot] != B200MIX_NO_SLOT;
    if(target->wet_channels != dd.wet_channels || target->out_channels != (toSlot ? dd.wet_channels : dd.dry_channels))
    { d->error = "slot_efx: the target maps do not match the device's wet / output mix"; return B200MIX_ERR_INVALID; }
    EfxParams P;
    if(int rc = efx_update(*props, *target, P))
    { d->error = rc == B200MIX_ERR_UNSUPPORTED ? "slot_efx: not supported in this configuration (see b200mix.h)"
        : "slot_efx: bad properties / maps"; return rc; }
    CUDA_TRY(d, cudaSetDevice(d->cuda_dev));
    if(!d->efx_ready) { CUDA_TRY(d, efx_kernels_init()); d->efx_ready = true; }
    if(d->efx.size() < d->h_slots.size()) d->efx.resize(d->h_slots.size());
    if(!d->d_efx_views)
        if(int rc = dev_alloc(d, d->d_efx_views, d->h_slots.size())) return rc;
    b200mix_device::EfxHost &H = d->efx[slot];
    const bool fresh = !H.used || d->h_slots[slot].type != props->type || H.p.lines != P.lines
        || H.p.echo_len != P.echo_len || H.p.cho_len != P.cho_len;
    if(P.type == B200MIX_EFFECT_AUTOWAH && P.lines > kEfxMaxLines - 2u)
    { d->error = "slot_efx: autowah handles up to 14 wet channels"; return B200MIX_ERR_UNSUPPORTED; }
    CUDA_TRY(d, cudaStreamSynchronize(d->stream));
    if(fresh)
    {
        // EffectState::deviceUpdate: new state, cleared
        free_slot(d, slot);
        SlotRec r{};
        r.type = props->type; r.channels = P.lines; r.fade_len = P.fade_len;
        auto alloc = [&](auto *&ptr, size_t count) -> int {
            if(int rc = dev_alloc(d, ptr, count)) return rc;
            d->slot_allocs[slot].push_back(ptr);
            return B200MIX_OK;
        };
        EfxDev *dev = nullptr;
        if(int rc = alloc(dev, 1)) return rc;
        if(int rc = alloc(r.lines, size_t(P.lines)*kLine)) return rc;
        if(int rc = alloc(r.gains, size_t(2)*P.lines*32)) return rc;
        if(int rc = alloc(r.gtgt, size_t(P.lines)*32)) return rc;
        EfxDev h{};
        h.p = P; h.comp_env = 1.0f;
        if(P.echo_len) { if(int rc = alloc(h.echo_buf, P.echo_len)) return rc; }
        if(P.cho_len) { if(int rc = alloc(h.cho_buf, size_t(4)*P.cho_len)) return rc; }
        if(P.type == B200MIX_EFFECT_FSHIFTER)
        {   // FshifterState::deviceUpdate (fshifter.cpp:140-148): cleared FIFOs, mPos = HilSize - HilStep
            if(int rc = alloc(h.fs_in, size_t(4)*1024)) return rc;
            if(int rc = alloc(h.fs_outfifo, size_t(4)*256)) return rc;
            if(int rc = alloc(h.fs_accum, size_t(4)*1024)) return rc;
            h.fs_count = 0u; h.fs_pos = 1024u - 256u;
        }
        if(P.type == B200MIX_EFFECT_PSHIFTER)
        {   // PshifterState::deviceUpdate (pshifter.cpp:131-145): cleared FIFOs and phases, mPos = StftSize - StftStep
            if(int rc = alloc(h.ps_fifo, size_t(9)*1024)) return rc;
            if(int rc = alloc(h.ps_accum, size_t(9)*1024)) return rc;
            if(int rc = alloc(h.ps_last, size_t(513))) return rc;
            if(int rc = alloc(h.ps_sum, size_t(513))) return rc;
            h.ps_count = 0u; h.ps_pos = 1024u - 128u;
        }
        r.H = reinterpret_cast<float*>(dev);
        CUDA_TRY(d, cudaMemcpyAsync(dev, &h, sizeof(h), cudaMemcpyHostToDevice, d->stream));
        CUDA_TRY(d, cudaStreamSynchronize(d->stream));
        H = b200mix_device::EfxHost{};
        H.used = true; H.dev = dev; H.mod_index = 0u; H.mod_range = P.mod_range ? P.mod_range : 1u;
        H.lfo_offset = 0u; H.lfo_range = P.cho_lfo_range ? P.cho_lfo_range : 1u;
        d->h_slots[slot] = r;
        ++d->active_slots; ++d->efx_slots;
        if(props->type == B200MIX_EFFECT_PSHIFTER) ++d->pshift_slots;
        d->dry_active = true;
    }
    else
    {
        // EffectState::update: new parameters, state kept.  The ring modulator rescales its
        // phase index to the new range (modulator.cpp:117-118); the host mirrors the index
        EfxDev hdr{};
        hdr.p = P;
        CUDA_TRY(d, cudaMemcpyAsync(H.dev, &hdr, sizeof(EfxParams), cudaMemcpyHostToDevice, d->stream));
        if(props->type == B200MIX_EFFECT_MODULATOR)
        {
            H.mod_index = uint32_t(uint64_t(H.mod_index) * P.mod_range_new / H.mod_range);
            H.mod_range = P.mod_range;
            CUDA_TRY(d, cudaMemcpyAsync(reinterpret_cast<char*>(H.dev) + offsetof(EfxDev, mod_index), &H.mod_index,
                sizeof(uint32_t), cudaMemcpyHostToDevice, d->stream));
        }
        if(props->type == B200MIX_EFFECT_CHORUS)
        {
            // mLfoOffset follows the LFO range (chorus.cpp:185-211)
            H.lfo_offset = P.cho_rate_on ? H.lfo_offset * P.cho_lfo_range_new / H.lfo_range : 0u;
            H.lfo_range = P.cho_lfo_range;
            CUDA_TRY(d, cudaMemcpyAsync(reinterpret_cast<char*>(H.dev) + offsetof(EfxDev, cho_lfo_offset), &H.lfo_offset,
                sizeof(uint32_t), cudaMemcpyHostToDevice, d->stream));
        }
        if(props->type == B200MIX_EFFECT_FSHIFTER)
        {   // a direction switched off zeroes that side's phase accumulators (fshifter.cpp:189-192,205-208)
            static const uint32_t zero = 0u;
            for(int c = 0;c < 4;++c)
                if(P.fs_reset_phase[c])
                    CUDA_TRY(d, cudaMemcpyAsync(reinterpret_cast<char*>(H.dev) + offsetof(EfxDev, fs_phase) + c*sizeof(uint32_t),
                        &zero, sizeof(uint32_t), cudaMemcpyHostToDevice, d->stream));
        }
        if(props->type == B200MIX_EFFECT_VMORPHER)
            // update() installs newly constructed formant filters: their histories restart at 0
            // (vmorpher.cpp:252-260)
            CUDA_TRY(d, cudaMemsetAsync(reinterpret_cast<char*>(H.dev) + offsetof(EfxDev, vm_s), 0,
                sizeof(EfxDev::vm_s), d->stream));
        d->h_slots[slot].fade_len = P.fade_len;
        CUDA_TRY(d, cudaStreamSynchronize(d->stream));
    }
    H.p = P;
    const SlotRec &S = d->h_slots[slot];
    CUDA_TRY(d, cudaMemcpyAsync(S.gtgt, P.gains, size_t(P.lines)*32*sizeof(float), cudaMemcpyHostToDevice, d->stream));
    if(P.snap_gains)
        for(int sel = 0;sel < 2;++sel)
            CUDA_TRY(d, cudaMemcpyAsync(S.gains + size_t(sel)*P.lines*32, P.gains, size_t(P.lines)*32*sizeof(float),
                cudaMemcpyHostToDevice, d->stream));
    CUDA_TRY(d, cudaStreamSynchronize(d->stream));
    return update_stages(d);
}

int b200mix_slot_target(b200mix_device *d, uint32_t slot, uint32_t target)
{
    if(!d || slot >= d->h_slots.size() || (target != B200MIX_NO_SLOT && target >= d->h_slots.size()))
    { if(d) d->error = "slot_target: slot out of range"; return B200MIX_ERR_INVALID; }
    if(d->h_target.size() != d->h_slots.size()) d->h_target.assign(d->h_slots.size(), B200MIX_NO_SLOT);
    uint32_t hops = 0;
    for(uint32_t t = target;t != B200MIX_NO_SLOT;t = d->h_target[t])
        if(t == slot || ++hops > d->h_slots.size())
        { d->error = "slot_target: the chain would loop"; return B200MIX_ERR_INVALID; }
    CUDA_TRY(d, cudaSetDevice(d->cuda_dev));
    CUDA_TRY(d, cudaStreamSynchronize(d->stream));
    d->h_target[slot] = target;
    return update_stages(d);
}

int b200mix_slot_reverb_update(b200mix_device *d, uint32_t slot, const b200mix_reverb_params *p,
    uint32_t full_update)
{
    if(!d || slot >= d->h_slots.size() || !p || p->struct_size != sizeof(*p)
        || d->h_slots[slot].type != B200MIX_EFFECT_REVERB)
    { if(d) d->error = "slot_reverb_update: no reverb installed on this slot / bad arguments"; return B200MIX_ERR_INVALID; }
    if(int rc = reverb_check_params(d, p)) return rc;
    b200mix_device::RvHost &R = d->rv[slot];
    const ReverbDev &h0 = R.h[0];
    if(p->main_len != h0.main_len || p->late_in_len != h0.late_in_len || p->early_ap_len != h0.early_ap_len
        || p->early_len != h0.early_len || p->late_ap_len != h0.late_ap_len || p->late_len != h0.late_len)
    { d->error = "slot_reverb_update: line lengths differ from the installed ones"; return B200MIX_ERR_INVALID; }
    CUDA_TRY(d, cudaSetDevice(d->cuda_dev));
    CUDA_TRY(d, cudaStreamSynchronize(d->stream));      // host mirrors are about to change
    if(full_update)
    {
        // reverb.cpp:1275-1279
        R.state = 1;
        R.cur ^= 1;
        const int old = R.cur ^ 1;
        R.h[old].early_tap_coeff = 0.0f;
        CUDA_TRY(d, cudaMemcpyAsync(reinterpret_cast<char*>(R.dev + old) + offsetof(ReverbDev, early_tap_coeff),
            &R.h[old].early_tap_coeff, sizeof(float), cudaMemcpyHostToDevice, d->stream));
        // the object coming back into use has not advanced mOffset while it was idle
        R.h[R.cur].offset = R.offset;
        CUDA_TRY(d, cudaMemcpyAsync(reinterpret_cast<char*>(R.dev + R.cur) + offsetof(ReverbDev, offset),
            &R.h[R.cur].offset, sizeof(uint32_t), cudaMemcpyHostToDevice, d->stream));
        d->h_slots[slot].rv_cur = uint32_t(R.cur);
    }
    reverb_fill_params(R.h[R.cur], p);
    R.fade[R.cur] = p->fade_samples;
    CUDA_TRY(d, cudaMemcpyAsync(R.dev + R.cur, &R.h[R.cur], kReverbParamBytes, cudaMemcpyHostToDevice, d->stream));
    CUDA_TRY(d, cudaStreamSynchronize(d->stream));
    return B200MIX_OK;
}

int b200mix_slot_output_gains(b200mix_device *d, uint32_t slot, uint32_t lines, const float *gains)
{
    if(!d || slot >= d->h_slots.size() || !d->h_slots[slot].type || !gains)
    { if(d) d->error = "slot_output_gains: bad arguments"; return B200MIX_ERR_INVALID; }
    const bool reverb = d->h_slots[slot].type == B200MIX_EFFECT_REVERB;
    if(lines != (reverb ? 8u : d->h_slots[slot].channels))
    { d->error = "slot_output_gains: wrong line count"; return B200MIX_ERR_INVALID; }
    CUDA_TRY(d, cudaSetDevice(d->cuda_dev));
    // gains address the slot's output target: the Dry mix or the target slot's Wet mix
    const bool toSlot = slot < d->h_target.size() && d->h_target[slot] != B200MIX_NO_SLOT;
    const uint32_t width = toSlot ? d->desc.wet_channels : d->desc.dry_channels;
    std::vector<float> g(size_t(lines)*32, 0.0f);
    for(uint32_t c = 0;c < lines;++c)
        for(uint32_t o = 0;o < width;++o)
            g[c*32 + o] = gains[c*width + o];
    // a reverb's gains are those of its CURRENT pipeline object (update3DPanning, reverb.cpp:1293-1296)
    float *dst = d->h_slots[slot].gtgt + (reverb ? size_t(d->rv[slot].cur)*8*32 : 0);
    CUDA_TRY(d, cudaMemcpyAsync(dst, g.data(), g.size()*sizeof(float), cudaMemcpyHostToDevice, d->stream));
    CUDA_TRY(d, cudaStreamSynchronize(d->stream));
    return B200MIX_OK;
}

int b200mix_hrtf_attach(b200mix_device *d, const b200mix_hrtf *h)
{
    if(!d || !h) return B200MIX_ERR_INVALID;
    if(h->ir_size > d->desc.ir_size)
    { d->error = "hrtf_attach: data set HRIRs are longer than the device's ir_size"; return B200MIX_ERR_INVALID; }
    CUDA_TRY(d, cudaSetDevice(d->cuda_dev));
    CUDA_TRY(d, cudaStreamSynchronize(d->stream));
    cudaFree(d->d_st_fields); cudaFree(d->d_st_elevs); cudaFree(d->d_st_coeffs); cudaFree(d->d_st_delays);
    d->d_st_fields = nullptr; d->d_st_elevs = nullptr; d->d_st_coeffs = nullptr; d->d_st_delays = nullptr;
    std::vector<float2> fields(h->fields.size());
    for(size_t i = 0;i < fields.size();++i)
    {
        uint32_t ev = h->fields[i].ev_count; float evf;
        std::memcpy(&evf, &ev, sizeof(evf));
        fields[i] = make_float2(h->fields[i].distance, evf);
    }
    std::vector<uint2> elevs(h->elevs.size());
    for(size_t i = 0;i < elevs.size();++i) elevs[i] = make_uint2(h->elevs[i].az_count, h->elevs[i].ir_offset);
    if(int rc = dev_alloc(d, d->d_st_fields, fields.size(), false)) return rc;
    if(int rc = dev_alloc(d, d->d_st_elevs, elevs.size(), false)) return rc;
    if(int rc = dev_alloc(d, d->d_st_coeffs, h->coeffs.size()/2, false)) return rc;
    if(int rc = dev_alloc(d, d->d_st_delays, h->delays.size(), false)) return rc;
    CUDA_TRY(d, cudaMemcpy(d->d_st_fields, fields.data(), fields.size()*sizeof(float2), cudaMemcpyHostToDevice));
    CUDA_TRY(d, cudaMemcpy(d->d_st_elevs, elevs.data(), elevs.size()*sizeof(uint2), cudaMemcpyHostToDevice));
    CUDA_TRY(d, cudaMemcpy(d->d_st_coeffs, h->coeffs.data(), h->coeffs.size()*sizeof(float), cudaMemcpyHostToDevice));
    CUDA_TRY(d, cudaMemcpy(d->d_st_delays, h->delays.data(), h->delays.size(), cudaMemcpyHostToDevice));
    d->st_num_fields = uint32_t(fields.size()); d->st_ir = h->ir_size;
    return B200MIX_OK;
}

static int voices_update_impl(b200mix_device *d, uint32_t n, const b200mix_voice_params *params,
    const float *hrtf_coeffs, const float *dirs, const float *dry_gains, const float *send_gains);

int b200mix_voices_update(b200mix_device *d, uint32_t n, const b200mix_voice_params *params,
    const float *hrtf_coeffs, const float *dry_gains, const float *send_gains)
{
    return voices_update_impl(d, n, params, hrtf_coeffs, nullptr, dry_gains, send_gains);
}

int b200mix_voices_update_dirs(b200mix_device *d, uint32_t n, const b200mix_voice_params *params,
    const float *dirs, const float *dry_gains, const float *send_gains)
{
    if(!d) return B200MIX_ERR_INVALID;
    if(!dirs) { d->error = "voices_update_dirs: null directions"; return B200MIX_ERR_INVALID; }
    if(!d->d_st_coeffs) { d->error = "voices_update_dirs: no HRTF data set attached"; return B200MIX_ERR_INVALID; }
    return voices_update_impl(d, n, params, nullptr, dirs, dry_gains, send_gains);
}

static int voices_update_impl(b200mix_device *d, uint32_t n, const b200mix_voice_params *params,
    const float *hrtf_coeffs, const float *dirs, const float *dry_gains, const float *send_gains)
{
    if(!d) return B200MIX_ERR_INVALID;
    if(n == 0) return B200MIX_OK;
    if(!params) { d->error = "voices_update: null params"; return B200MIX_ERR_INVALID; }
    const b200mix_device_desc &dd = d->desc;
    CUDA_TRY(d, cudaSetDevice(d->cuda_dev));
    if(d->stage_busy)
    {
        CUDA_TRY(d, cudaEventSynchronize(d->stage_done));
        d->stage_busy = false;
    }
    if(int rc = ensure_stage(d, n)) return rc;

    for(uint32_t i = 0;i < n;++i)
    {
        const b200mix_voice_params &p = params[i];
        const bool nobuf = p.buffer == B200MIX_NO_BUFFER;
        if(p.voice >= dd.max_voices || p.resampler > B200MIX_RESAMPLER_BSINC48
            || (!(p.flags & B200MIX_VF_STOPPED) && !nobuf && p.buffer >= dd.max_buffers))
        { d->error = "voices_update: voice/buffer/resampler out of range"; return B200MIX_ERR_INVALID; }
        if((p.flags & B200MIX_VF_LOOPING) && p.loop_end <= p.loop_start)
        { d->error = "voices_update: empty loop"; return B200MIX_ERR_INVALID; }
        // MaxPitch clamp of the parameter stage (alc/alu.cpp:1682-1685,1996-1999): CalculateBufferSize
        // relies on it
        if(p.step > (10u << 16))
        { d->error = "voices_update: step above MaxPitch<<16"; return B200MIX_ERR_INVALID; }
        if(!(p.flags & B200MIX_VF_STOPPED))
        {
            if(nobuf) { /* nothing to read */ }
            else if(p.flags & B200MIX_VF_STATIC)
            {
                const BufferRec &hb = d->h_buffers[p.buffer];
                if(!hb.data || !hb.frames)
                { d->error = "voices_update: static voice on a buffer without data"; return B200MIX_ERR_INVALID; }
                if((p.flags & B200MIX_VF_LOOPING) && p.loop_end > hb.frames)
                { d->error = "voices_update: loop end beyond the buffer"; return B200MIX_ERR_INVALID; }
            }
            else if(!d->d_qhdr)
            {
                // a streaming voice reads its queue: make sure the (empty) queue table exists
                if(int rc = dev_alloc(d, d->d_qhdr, dd.max_voices)) return rc;
                if(int rc = dev_alloc(d, d->d_queue, size_t(dd.max_voices)*kMaxQueue)) return rc;
            }
        }
        VoiceUpdate &u = d->h_upd[i];
        u.voice = p.voice; u.flags = p.flags; u.buffer = p.buffer; u.resampler = p.resampler;
        if(nobuf) { u.flags |= kUpNoBuffer; u.buffer = 0u; }
        u.position = p.position; u.position_frac = p.position_frac;
        u.loop_start = p.loop_start; u.loop_end = p.loop_end; u.step = p.step;
        u.bsinc_sf = 0.0f; u.bsinc_m = 0; u.bsinc_l = 0; u.bsinc_off = 0;
        if(const BsincTable *t = bsinc_for(d, p.resampler))
        {
            const BsincState st = PrepareBsinc(*t, p.step);
            u.bsinc_sf = st.sf; u.bsinc_m = st.m; u.bsinc_l = st.l; u.bsinc_off = st.offset;
        }
        u.delay0 = p.hrtf_delay[0]; u.delay1 = p.hrtf_delay[1]; u.gain = p.hrtf_gain;
        if(dirs) { u.delay0 = 0; u.delay1 = 0; }        // computed on the device
        if(u.delay0 >= B200MIX_HRTF_HISTORY || u.delay1 >= B200MIX_HRTF_HISTORY)
        { d->error = "voices_update: HRTF delay out of range"; return B200MIX_ERR_INVALID; }
        for(uint32_t s = 0;s < B200MIX_MAX_SENDS;++s)
        {
            u.send_slot[s] = (s < dd.num_sends) ? p.send_slot[s] : B200MIX_NO_SLOT;
            if(u.send_slot[s] != B200MIX_NO_SLOT && u.send_slot[s] >= dd.max_slots)
            { d->error = "voices_update: send slot out of range"; return B200MIX_ERR_INVALID; }
        }
        if(!d->h_send_slot.empty())
            for(uint32_t s2 = 0;s2 < B200MIX_MAX_SENDS;++s2)
            {
                uint32_t &m = d->h_send_slot[size_t(p.voice)*B200MIX_MAX_SENDS + s2];
                const uint32_t nv2 = (p.flags & B200MIX_VF_STOPPED) ? B200MIX_NO_SLOT : u.send_slot[s2];
                if(m != nv2) { m = nv2; d->sends_dirty = true; }
            }
        u.has_coeffs = (hrtf_coeffs != nullptr || (dirs != nullptr && (p.flags & B200MIX_VF_HRTF))) && dd.ir_size > 0;
        u.has_dry = dry_gains != nullptr;
        if(!(p.flags & B200MIX_VF_HRTF) && !(p.flags & B200MIX_VF_STOPPED))
        {
            d->dry_active = true;
            if(d->mix_cdr == 0)
                if(int rc = ensure_dry_park(d)) return rc;
        }
        {
            const uint8_t hv = (p.flags & B200MIX_VF_HRTF) ? 1 : 0;
            if(d->h_hrtf[p.voice] != hv) { d->h_hrtf[p.voice] = hv; d->dry_entries_dirty = true; }
        }
        {
            // mixing-order bookkeeping: membership and a cost key (resampler taps per output)
            const uint8_t act = (p.flags & B200MIX_VF_STOPPED) ? 0 : 1;
            uint32_t cost = (p.step == 65536u) ? 1u : (u.bsinc_m ? u.bsinc_m : (p.resampler >= 2u ? 4u : 2u));
            if(d->h_active[p.voice] != act || d->h_cost[p.voice] != cost) d->order_dirty = true;
            d->h_active[p.voice] = act; d->h_cost[p.voice] = cost;
        }
        d->voice_hi = std::max(d->voice_hi, p.voice + 1u);
        {
            const uint32_t nb = (!(p.flags & B200MIX_VF_STOPPED) && (p.flags & B200MIX_VF_STATIC) && !nobuf)
                ? p.buffer : B200MIX_NO_SLOT;
            uint32_t &ob = d->h_vbuf[p.voice];
            if(ob != nb)
            {
                if(ob != B200MIX_NO_SLOT) --d->h_bufrefs[ob];
                if(nb != B200MIX_NO_SLOT) ++d->h_bufrefs[nb];
                ob = nb;
            }
        }
        if((p.flags & B200MIX_VF_RESET) && !d->h_dfilt.empty() && d->h_dfilt[p.voice])
        { d->h_dfilt[p.voice] = 0; d->order2_dirty = true; }
    }
    ApplyParams A{};
    A.voices = d->d_voices; A.updates = reinterpret_cast<const VoiceUpdate*>(d->d_arena);
    size_t off = align16(size_t(n)*sizeof(VoiceUpdate));
    auto pack = [&](const float *src, size_t count) -> const float* {
        std::memcpy(d->h_arena + off, src, count*sizeof(float));
        const float *dev = reinterpret_cast<const float*>(d->d_arena + off);
        off += align16(count*sizeof(float));
        return dev;
    };
    if(dirs && dd.ir_size)
    {
        A.dirs = reinterpret_cast<const float4*>(pack(dirs, size_t(n)*4));
        A.st_fields = d->d_st_fields; A.st_elevs = d->d_st_elevs; A.st_coeffs = d->d_st_coeffs;
        A.st_delays = d->d_st_delays; A.st_num_fields = d->st_num_fields; A.st_ir = d->st_ir;
    }
    else if(hrtf_coeffs && dd.ir_size)
        A.coeffs = pack(hrtf_coeffs, size_t(n)*dd.ir_size*2);
    if(dry_gains && dd.dry_channels)
        A.dry = pack(dry_gains, size_t(n)*dd.dry_channels);
    if(send_gains && dd.num_sends && dd.wet_channels)
        A.send = pack(send_gains, size_t(n)*dd.num_sends*dd.wet_channels);
    CUDA_TRY(d, cudaMemcpyAsync(d->d_arena, d->h_arena, off, cudaMemcpyHostToDevice, d->stream));
    A.hrtf_tgt = d->d_hrtf_tgt; A.hrtf_old = d->d_hrtf_old;
    A.dry_cur = d->d_dry_cur; A.dry_tgt = d->d_dry_tgt;
    A.send_cur = d->d_send_cur; A.send_tgt = d->d_send_tgt;
    A.ir = dd.ir_size; A.ir_pad = d->ir_pad; A.cd = dd.dry_channels; A.cw = dd.wet_channels;
    A.num_sends = dd.num_sends;
    A.filt = d->d_filt; A.filt_paths = 1u + dd.num_sends;
    A.qhdr = d->d_qhdr;
    k_apply_updates<<<n, 64, 0, d->stream>>>(A);
    ++d->launches;
    CUDA_TRY(d, cudaGetLastError());
    CUDA_TRY(d, cudaEventRecord(d->stage_done, d->stream));
    d->stage_busy = true;
    return B200MIX_OK;
}

// Could any path of this source need a filter?  (All of these leave GainHF == GainLF == 1 exactly
// when false, alc/alu.cpp:1854-1961 — a cheap scan so that scenes without filters never allocate
// or run the filter stage.)
static bool source_may_filter(const b200mix_source_props &P, uint32_t num_sends)
{
    if(P.direct.gain_hf != 1.0f || P.direct.gain_lf != 1.0f || P.air_absorption_factor != 0.0f) return true;
    if(P.inner_angle < 360.0f && (P.outer_gain_hf != 1.0f)) return true;
    for(uint32_t s = 0;s < num_sends;++s)
        if(P.sends[s].gain_hf != 1.0f || P.sends[s].gain_lf != 1.0f
            || (P.sends[s].active && P.sends[s].slot_air_absorption_gain_hf < 1.0f)) return true;
    return false;
}

int b200mix_sources_update(b200mix_device *d, uint32_t n, const b200mix_source_voice *voices,
    const b200mix_source_props *props, const b200mix_listener_params *listener, const b200mix_voice_env *env)
{
    if(!d) return B200MIX_ERR_INVALID;
    if(n == 0) return B200MIX_OK;
    const b200mix_device_desc &dd = d->desc;
    if(!voices || !props || !listener || !env || env->struct_size != sizeof(*env)
        || listener->struct_size != sizeof(*listener) || env->render_mode > 2u
        || env->num_sends != dd.num_sends || !env->device_rate || listener->distance_model > 6u)
    { d->error = "sources_update: bad arguments (env->num_sends must equal the device's)"; return B200MIX_ERR_INVALID; }
    if(env->render_mode == 2u && (!dd.ir_size || !d->d_st_coeffs))
    { d->error = "sources_update: HRTF rendering needs an HRTF device and b200mix_hrtf_attach"; return B200MIX_ERR_INVALID; }
    if(env->render_mode != 2u && (env->dry.channels != dd.dry_channels || !env->dry.scale || !env->dry.index))
    { d->error = "sources_update: env->dry must describe the device's Dry mix"; return B200MIX_ERR_INVALID; }
    if(dd.num_sends && (env->wet_stride != dd.wet_channels))
    { d->error = "sources_update: env->wet_stride must equal the device's wet_channels"; return B200MIX_ERR_INVALID; }
    for(uint32_t s = 0;s < dd.num_sends;++s)
        if(env->wet[s].channels > dd.wet_channels || (env->wet[s].channels && (!env->wet[s].scale || !env->wet[s].index)))
        { d->error = "sources_update: bad wet map"; return B200MIX_ERR_INVALID; }
    CUDA_TRY(d, cudaSetDevice(d->cuda_dev));

    bool mayFilter = d->d_filt != nullptr;
    const bool hrtfMode = env->render_mode == 2u;
    for(uint32_t i = 0;i < n;++i)
    {
        const b200mix_source_voice &p = voices[i];
        const b200mix_source_props &P = props[i];
        if(P.struct_size != sizeof(P) || P.distance_model > 6u)
        { d->error = "sources_update: bad source props"; return B200MIX_ERR_INVALID; }
        if(p.voice >= dd.max_voices || p.resampler > B200MIX_RESAMPLER_BSINC48
            || (!(p.flags & B200MIX_VF_STOPPED) && p.buffer >= dd.max_buffers))
        { d->error = "sources_update: voice/buffer/resampler out of range"; return B200MIX_ERR_INVALID; }
        if((p.flags & B200MIX_VF_LOOPING) && p.loop_end <= p.loop_start)
        { d->error = "sources_update: empty loop"; return B200MIX_ERR_INVALID; }
        if(!(p.flags & B200MIX_VF_STOPPED))
        {
            if(p.flags & B200MIX_VF_STATIC)
            {
                const BufferRec &hb = d->h_buffers[p.buffer];
                if(!hb.data || !hb.frames)
                { d->error = "sources_update: static voice on a buffer without data"; return B200MIX_ERR_INVALID; }
                if((p.flags & B200MIX_VF_LOOPING) && p.loop_end > hb.frames)
                { d->error = "sources_update: loop end beyond the buffer"; return B200MIX_ERR_INVALID; }
            }
            else if(!d->d_qhdr)
            {
                if(int rc = dev_alloc(d, d->d_qhdr, dd.max_voices)) return rc;
                if(int rc = dev_alloc(d, d->d_queue, size_t(dd.max_voices)*kMaxQueue)) return rc;
            }
        }
        for(uint32_t s = 0;s < dd.num_sends;++s)
            if(p.send_slot[s] != B200MIX_NO_SLOT && p.send_slot[s] >= dd.max_slots)
            { d->error = "sources_update: send slot out of range"; return B200MIX_ERR_INVALID; }
        if(!mayFilter && source_may_filter(P, dd.num_sends)) mayFilter = true;
    }
    if(mayFilter)
    {
        if(int rc = ensure_filters(d)) return rc;
        if(!d->dev_filters) { d->dev_filters = true; d->order2_dirty = true; }
    }
    // host mirrors (as b200mix_voices_update keeps them); the step is not known here: the cost
    // key of the mixing order takes the resampler's widest filter
    for(uint32_t i = 0;i < n;++i)
    {
        const b200mix_source_voice &p = voices[i];
        const bool stopped = (p.flags & B200MIX_VF_STOPPED) != 0;
        if(!d->h_send_slot.empty())
            for(uint32_t s2 = 0;s2 < B200MIX_MAX_SENDS;++s2)
            {
                uint32_t &m = d->h_send_slot[size_t(p.voice)*B200MIX_MAX_SENDS + s2];
                const uint32_t nv2 = (stopped || s2 >= dd.num_sends) ? B200MIX_NO_SLOT : p.send_slot[s2];
                if(m != nv2) { m = nv2; d->sends_dirty = true; }
            }
        if(!hrtfMode && !stopped)
        {
            d->dry_active = true;
            if(d->mix_cdr == 0)
                if(int rc = ensure_dry_park(d)) return rc;
        }
        const uint8_t hv = hrtfMode ? 1 : 0;
        if(d->h_hrtf[p.voice] != hv) { d->h_hrtf[p.voice] = hv; d->dry_entries_dirty = true; }
        const uint8_t act = stopped ? 0 : 1;
        uint32_t cost = p.resampler >= 2u ? 4u : 2u;
        if(const BsincTable *t = bsinc_for(d, p.resampler)) cost = t->m[0];
        if(d->h_active[p.voice] != act || (!d->h_cost[p.voice] && act)) { d->order_dirty = true; d->h_cost[p.voice] = cost; }
        d->h_active[p.voice] = act;
        d->voice_hi = std::max(d->voice_hi, p.voice + 1u);
        const uint32_t nb = (!stopped && (p.flags & B200MIX_VF_STATIC)) ? p.buffer : B200MIX_NO_SLOT;
        uint32_t &ob = d->h_vbuf[p.voice];
        if(ob != nb)
        {
            if(ob != B200MIX_NO_SLOT) --d->h_bufrefs[ob];
            if(nb != B200MIX_NO_SLOT) ++d->h_bufrefs[nb];
            ob = nb;
        }
    }

    // staging: [voices n][props n] in, [VoiceUpdate n][dirs n][dry][send][hf/lf][FilterUpdate] scratch
    const uint32_t paths = 1u + dd.num_sends;
    const size_t inBytes = align16(size_t(n)*sizeof(b200mix_source_voice)) + align16(size_t(n)*sizeof(b200mix_source_props));
    if(d->src_busy) { CUDA_TRY(d, cudaEventSynchronize(d->src_done)); d->src_busy = false; }
    if(n > d->src_cap)
    {
        CUDA_TRY(d, cudaStreamSynchronize(d->stream));
        if(d->h_src) cudaFreeHost(d->h_src);
        cudaFree(d->d_src); d->h_src = nullptr; d->d_src = nullptr;
        const uint32_t cap = std::max(n, 2u*d->src_cap);
        const size_t in = align16(size_t(cap)*sizeof(b200mix_source_voice)) + align16(size_t(cap)*sizeof(b200mix_source_props));
        const size_t out = align16(size_t(cap)*sizeof(VoiceUpdate)) + align16(size_t(cap)*16)
            + align16(size_t(cap)*std::max(dd.dry_channels, 1u)*4) + align16(size_t(cap)*std::max(dd.num_sends*dd.wet_channels, 1u)*4)
            + align16(size_t(cap)*(1u + B200MIX_MAX_SENDS)*8) + align16(size_t(cap)*paths*sizeof(FilterUpdate));
        CUDA_TRY(d, cudaMallocHost(reinterpret_cast<void**>(&d->h_src), in));
        CUDA_TRY(d, cudaMalloc(reinterpret_cast<void**>(&d->d_src), in + out + 64));
        if(!d->src_done) CUDA_TRY(d, cudaEventCreateWithFlags(&d->src_done, cudaEventDisableTiming));
        d->src_cap = cap;
    }
    std::memcpy(d->h_src, voices, size_t(n)*sizeof(b200mix_source_voice));
    const size_t offProps = align16(size_t(n)*sizeof(b200mix_source_voice));
    std::memcpy(d->h_src + offProps, props, size_t(n)*sizeof(b200mix_source_props));
    CUDA_TRY(d, cudaMemcpyAsync(d->d_src, d->h_src, inBytes, cudaMemcpyHostToDevice, d->stream));
    CUDA_TRY(d, cudaEventRecord(d->src_done, d->stream));
    d->src_busy = true;

    CalcVoicesParams Q{};
    Q.voices = reinterpret_cast<const b200mix_source_voice*>(d->d_src);
    Q.props = reinterpret_cast<const b200mix_source_props*>(d->d_src + offProps);
    Q.n = n; Q.listener = *listener;
    Q.device_rate = env->device_rate; Q.num_sends = dd.num_sends; Q.render_mode = env->render_mode;
    Q.cd = dd.dry_channels; Q.cw = dd.wet_channels; Q.ir = dd.ir_size;
    if(!hrtfMode)
    {
        Q.dry_channels = env->dry.channels;
        for(uint32_t c = 0;c < env->dry.channels;++c) { Q.dry_scale[c] = env->dry.scale[c]; Q.dry_index[c] = env->dry.index[c]; }
    }
    for(uint32_t s = 0;s < dd.num_sends;++s)
    {
        Q.wet_channels[s] = env->wet[s].channels;
        for(uint32_t c = 0;c < env->wet[s].channels;++c) { Q.wet_scale[s][c] = env->wet[s].scale[c]; Q.wet_index[s][c] = env->wet[s].index[c]; }
    }
    for(int t = 0;t < 3;++t)
    {
        Q.bsinc[t].scaleBase = d->bsinc[t].scaleBase; Q.bsinc[t].scaleRange = d->bsinc[t].scaleRange;
        for(unsigned k = 0;k < kBsincScales;++k) { Q.bsinc[t].m[k] = d->bsinc[t].m[k]; Q.bsinc[t].filterOffset[k] = d->bsinc[t].filterOffset[k]; }
    }
    size_t off = inBytes;
    auto carve = [&](size_t bytes) { char *p = d->d_src + off; off += align16(bytes); return p; };
    Q.updates = reinterpret_cast<VoiceUpdate*>(carve(size_t(n)*sizeof(VoiceUpdate)));
    Q.dirs = reinterpret_cast<float4*>(carve(size_t(n)*16));
    Q.dry = reinterpret_cast<float*>(carve(size_t(n)*std::max(dd.dry_channels, 1u)*4));
    Q.send = (dd.num_sends && dd.wet_channels)
        ? reinterpret_cast<float*>(carve(size_t(n)*dd.num_sends*dd.wet_channels*4)) : nullptr;
    Q.gains_hflf = reinterpret_cast<float*>(carve(size_t(n)*(1u + B200MIX_MAX_SENDS)*8));
    Q.fupd = reinterpret_cast<FilterUpdate*>(carve(size_t(n)*paths*sizeof(FilterUpdate)));
    const bool filters = d->d_filt != nullptr;
    CUDA_TRY(d, launch_calc_voices(Q, filters, d->stream));
    d->launches += filters ? 2 : 1;

    ApplyParams A{};
    A.voices = d->d_voices; A.updates = Q.updates;
    if(hrtfMode)
    {
        A.dirs = Q.dirs;
        A.st_fields = d->d_st_fields; A.st_elevs = d->d_st_elevs; A.st_coeffs = d->d_st_coeffs;
        A.st_delays = d->d_st_delays; A.st_num_fields = d->st_num_fields; A.st_ir = d->st_ir;
    }
    else A.dry = Q.dry;
    A.send = Q.send;
    A.hrtf_tgt = d->d_hrtf_tgt; A.hrtf_old = d->d_hrtf_old;
    A.dry_cur = d->d_dry_cur; A.dry_tgt = d->d_dry_tgt;
    A.send_cur = d->d_send_cur; A.send_tgt = d->d_send_tgt;
    A.ir = dd.ir_size; A.ir_pad = d->ir_pad; A.cd = dd.dry_channels; A.cw = dd.wet_channels;
    A.num_sends = dd.num_sends;
    A.filt = d->d_filt; A.filt_paths = paths;
    A.qhdr = d->d_qhdr;
    k_apply_updates<<<n, 64, 0, d->stream>>>(A);
    ++d->launches;
    if(filters)
    {
        k_apply_filter_updates<<<(2u*n*paths + 127u)/128u, 128, 0, d->stream>>>(d->d_filt, paths, Q.fupd, n*paths);
        ++d->launches;
    }
    CUDA_TRY(d, cudaGetLastError());
    return B200MIX_OK;
}

int b200mix_get_voice_targets(b200mix_device *d, uint32_t voice, uint32_t *step, float bsinc[4],
    float *hrtf_gain, uint32_t hrtf_delay[2], float *hrtf_coeffs, float *dry_gains, float *send_gains,
    float *filters)
{
    if(!d || voice >= d->desc.max_voices) return B200MIX_ERR_INVALID;
    const b200mix_device_desc &dd = d->desc;
    CUDA_TRY(d, cudaSetDevice(d->cuda_dev));
    CUDA_TRY(d, cudaStreamSynchronize(d->stream));
    VoiceRec rec;
    CUDA_TRY(d, cudaMemcpy(&rec, d->d_voices + voice, offsetof(VoiceRec, prev), cudaMemcpyDeviceToHost));
    if(step) *step = rec.step;
    if(bsinc)
    {
        bsinc[0] = rec.bsinc_sf;
        std::memcpy(&bsinc[1], &rec.bsinc_m, 4); std::memcpy(&bsinc[2], &rec.bsinc_l, 4);
        std::memcpy(&bsinc[3], &rec.bsinc_off, 4);
    }
    if(hrtf_gain) *hrtf_gain = rec.tgt_gain;
    if(hrtf_delay) { hrtf_delay[0] = rec.tgt_delay0; hrtf_delay[1] = rec.tgt_delay1; }
    if(hrtf_coeffs && d->d_hrtf_tgt)
        CUDA_TRY(d, cudaMemcpy(hrtf_coeffs, d->d_hrtf_tgt + size_t(voice)*d->ir_pad, size_t(dd.ir_size)*8, cudaMemcpyDeviceToHost));
    if(dry_gains && dd.dry_channels)
        CUDA_TRY(d, cudaMemcpy(dry_gains, d->d_dry_tgt + size_t(voice)*dd.dry_channels, size_t(dd.dry_channels)*4, cudaMemcpyDeviceToHost));
    if(send_gains && d->d_send_tgt)
        CUDA_TRY(d, cudaMemcpy(send_gains, d->d_send_tgt + size_t(voice)*dd.num_sends*dd.wet_channels,
            size_t(dd.num_sends)*dd.wet_channels*4, cudaMemcpyDeviceToHost));
    if(filters)
    {
        const uint32_t paths = 1u + dd.num_sends;
        for(uint32_t pth = 0;pth < paths;++pth)
        {
            float *o = filters + size_t(pth)*11;
            for(int k = 0;k < 11;++k) o[k] = k == 1 ? 1.0f : (k == 6 ? 1.0f : 0.0f);
            if(!d->d_filt) continue;
            FilterRec fr;
            CUDA_TRY(d, cudaMemcpy(&fr, d->d_filt + size_t(voice)*paths + pth, sizeof(fr), cudaMemcpyDeviceToHost));
            o[0] = fr.active ? 1.0f : 0.0f;
            for(int k = 0;k < 5;++k) { o[1+k] = fr.tgt[0][k]; o[6+k] = fr.tgt[1][k]; }
        }
    }
    return B200MIX_OK;
}

int b200mix_voice_queue(b200mix_device *d, uint32_t voice, uint32_t count, const uint32_t *buffers,
    uint32_t loop_index)
{
    if(!d) return B200MIX_ERR_INVALID;
    const b200mix_device_desc &dd = d->desc;
    if(voice >= dd.max_voices || (count && !buffers))
    { d->error = "voice_queue: bad arguments"; return B200MIX_ERR_INVALID; }
    if(count > B200MIX_MAX_QUEUE)
    { d->error = "voice_queue: more than B200MIX_MAX_QUEUE items"; return B200MIX_ERR_UNSUPPORTED; }
    if(loop_index != B200MIX_NO_LOOP && loop_index >= count)
    { d->error = "voice_queue: loop index outside the list"; return B200MIX_ERR_INVALID; }
    QueueSet Q{};
    Q.voice = voice; Q.count = count; Q.loop = loop_index;
    for(uint32_t i = 0;i < count;++i)
    {
        if(buffers[i] >= dd.max_buffers || !d->h_buffers[buffers[i]].data)
        { d->error = "voice_queue: buffer id without data"; return B200MIX_ERR_INVALID; }
        Q.items[i] = buffers[i];
    }
    CUDA_TRY(d, cudaSetDevice(d->cuda_dev));
    if(!d->d_qhdr)
    {
        if(int rc = dev_alloc(d, d->d_qhdr, dd.max_voices)) return rc;
        if(int rc = dev_alloc(d, d->d_queue, size_t(dd.max_voices)*kMaxQueue)) return rc;
    }
    k_set_queue<<<1, 32, 0, d->stream>>>(d->d_voices, d->d_qhdr, d->d_queue, Q);
    ++d->launches;
    CUDA_TRY(d, cudaGetLastError());
    return B200MIX_OK;
}

// Filter state of every voice path (allocated by the first filter a host or the GPU parameter
// stage sets; a device that never sees one pays nothing).
static int ensure_filters(b200mix_device *d)
{
    if(d->d_filt) return B200MIX_OK;
    const b200mix_device_desc &dd = d->desc;
    const uint32_t paths = 1u + dd.num_sends;
    const size_t count = size_t(dd.max_voices)*paths;
    if(int rc = dev_alloc(d, d->d_filt, count, false)) return rc;
    k_filter_init<<<unsigned((count*32u + 255u)/256u), 256, 0, d->stream>>>(d->d_filt, count);
    ++d->launches;
    CUDA_TRY(d, cudaGetLastError());
    CUDA_TRY(d, cudaEventCreateWithFlags(&d->fstage_done, cudaEventDisableTiming));
    if(!d->d_xscratch)
        if(int rc = dev_alloc(d, d->d_xscratch, size_t(dd.max_voices)*kLine)) return rc;
    if(!d->d_sendinfo)
        if(int rc = dev_alloc(d, d->d_sendinfo, dd.max_voices)) return rc;
    if(int rc = dev_alloc(d, d->d_dline, size_t(dd.max_voices)*kLine)) return rc;
    if(int rc = dev_alloc(d, d->d_order2, dd.max_voices)) return rc;
    d->h_dfilt.assign(dd.max_voices, 0);
    return B200MIX_OK;
}

int b200mix_voices_filters(b200mix_device *d, uint32_t n, const b200mix_voice_filter *filters)
{
    static_assert(sizeof(FilterUpdate) == sizeof(b200mix_voice_filter), "FilterUpdate mirrors the ABI struct");
    if(!d) return B200MIX_ERR_INVALID;
    if(n == 0) return B200MIX_OK;
    if(!filters) { d->error = "voices_filters: null filters"; return B200MIX_ERR_INVALID; }
    const b200mix_device_desc &dd = d->desc;
    const uint32_t paths = 1u + dd.num_sends;
    for(uint32_t i = 0;i < n;++i)
        if(filters[i].voice >= dd.max_voices || filters[i].path >= paths)
        { d->error = "voices_filters: voice/path out of range"; return B200MIX_ERR_INVALID; }
    CUDA_TRY(d, cudaSetDevice(d->cuda_dev));
    if(int rc = ensure_filters(d)) return rc;
    for(uint32_t i = 0;i < n;++i)
        if(filters[i].path == 0)
        {
            const uint8_t act = filters[i].active ? 1 : 0;
            if(d->h_dfilt[filters[i].voice] != act) { d->h_dfilt[filters[i].voice] = act; d->order2_dirty = true; }
        }
    if(d->fstage_busy)
    {
        CUDA_TRY(d, cudaEventSynchronize(d->fstage_done));
        d->fstage_busy = false;
    }
    if(n > d->fupd_cap)
    {
        CUDA_TRY(d, cudaStreamSynchronize(d->stream));
        if(d->h_fupd) cudaFreeHost(d->h_fupd);
        cudaFree(d->d_fupd);
        d->h_fupd = nullptr; d->d_fupd = nullptr;
        const uint32_t cap = std::max(n, 2u*d->fupd_cap);
        CUDA_TRY(d, cudaMallocHost(reinterpret_cast<void**>(&d->h_fupd), size_t(cap)*sizeof(FilterUpdate)));
        CUDA_TRY(d, cudaMalloc(reinterpret_cast<void**>(&d->d_fupd), size_t(cap)*sizeof(FilterUpdate)));
        d->fupd_cap = cap;
    }
    std::memcpy(d->h_fupd, filters, size_t(n)*sizeof(FilterUpdate));
    CUDA_TRY(d, cudaMemcpyAsync(d->d_fupd, d->h_fupd, size_t(n)*sizeof(FilterUpdate),
        cudaMemcpyHostToDevice, d->stream));
    k_apply_filter_updates<<<(2u*n + 127u)/128u, 128, 0, d->stream>>>(d->d_filt, paths, d->d_fupd, n);
    ++d->launches;
    CUDA_TRY(d, cudaGetLastError());
    CUDA_TRY(d, cudaEventRecord(d->fstage_done, d->stream));
    d->fstage_busy = true;
    return B200MIX_OK;
}

static inline void stage_mark(b200mix_device *d, int i)
{ if(d->profile_level >= 2) cudaEventRecord(d->ev_stage[i], d->stream); }

// Phase A of an update: clear the mix buffers, mix every voice, reduce the partial rows and
// finish the aux sends -> the slots' Wet buffers are complete (alc/alu.cpp:2196-2206).
static int render_phase_a(b200mix_device *d, uint32_t frames, bool want_results, bool force_sends)
{
    const b200mix_device_desc &dd = d->desc;
    if(frames < 1 || frames > B200MIX_LINE_SIZE)
    { d->error = "render: frames out of range"; return B200MIX_ERR_INVALID; }
    if(dd.post_process == B200MIX_POST_HRTF && !d->dec_channels)
    { d->error = "render: HRTF decoder not set"; return B200MIX_ERR_INVALID; }
    if(dd.post_process == B200MIX_POST_AMBIDEC && !d->amb_in)
    { d->error = "render: ambisonic decoder not set"; return B200MIX_ERR_INVALID; }
    CUDA_TRY(d, cudaSetDevice(d->cuda_dev));

    stage_mark(d, 0);
    // clear MixBuffer (alc/alu.cpp:2417) and the wet buffers (alc/alu.cpp:2196-2198)
    // (the dry mix is left alone while nothing can write or read it: HRTF-only scenes; an
    // HRTF post-process whose RealOut is just L/R overwrites it instead of accumulating)
    if(d->dry_active || dd.post_process != B200MIX_POST_HRTF)
        CUDA_TRY(d, cudaMemsetAsync(d->d_dry, 0, size_t(d->dry_alloc_ch)*kLine*sizeof(float), d->stream));
    d->real_overwrite = dd.post_process == B200MIX_POST_HRTF && dd.real_channels == 2
        && dd.real_left != dd.real_right;
    if(d->d_real != d->d_dry && !d->real_overwrite)
        CUDA_TRY(d, cudaMemsetAsync(d->d_real, 0, size_t(dd.real_channels)*kLine*sizeof(float), d->stream));
    if(d->d_wet)
        CUDA_TRY(d, cudaMemsetAsync(d->d_wet, 0, size_t(dd.max_slots)*dd.wet_channels*kLine*sizeof(float), d->stream));

    if(d->order_dirty)
    {
        d->h_order.clear();
        for(uint32_t v = 0;v < d->voice_hi;++v) if(d->h_active[v]) d->h_order.push_back(v);
        std::stable_sort(d->h_order.begin(), d->h_order.end(),
            [d](uint32_t a, uint32_t b) { return d->h_cost[a] > d->h_cost[b]; });
        d->num_order = uint32_t(d->h_order.size());
        if(d->num_order)
        {
            // the vector may be reused before the copy completes: synchronise (rare path)
            CUDA_TRY(d, cudaMemcpyAsync(d->d_order, d->h_order.data(), d->num_order*sizeof(uint32_t),
                cudaMemcpyHostToDevice, d->stream));
            CUDA_TRY(d, cudaStreamSynchronize(d->stream));
        }
        d->order_dirty = false;
        d->order2_dirty = true;
        d->dry_entries_dirty = true;
    }
    if(d->d_filt && d->order2_dirty)
    {
        d->h_order2.clear();
        // with the GPU parameter stage the host does not know which direct filters are active:
        // the second pass then looks at every voice's kSiDeferred bit
        for(uint32_t v : d->h_order) if(d->dev_filters || d->h_dfilt[v]) d->h_order2.push_back(v);
        d->num_order2 = uint32_t(d->h_order2.size());
        if(d->num_order2)
        {
            CUDA_TRY(d, cudaMemcpyAsync(d->d_order2, d->h_order2.data(), d->num_order2*sizeof(uint32_t),
                cudaMemcpyHostToDevice, d->stream));
            CUDA_TRY(d, cudaStreamSynchronize(d->stream));
        }
        d->order2_dirty = false;
    }
    const Variant var = get_variant(d->mix_variant);
    const uint32_t nv = std::max(d->voice_hi, 1u);
    const uint32_t maxBlocks = uint32_t(d->num_sms*d->mix_blocks_per_sm);
    const uint32_t blocks = std::max(1u, std::min(maxBlocks, (d->num_order + var.groups - 1)/var.groups));
    const size_t rows = size_t(blocks);        // one partial row per CTA

    MixParams P{};
    P.voices = d->d_voices; P.buffers = d->d_buffers;
    P.hrtf_tgt = d->d_hrtf_tgt; P.hrtf_old = d->d_hrtf_old;
    P.dry_cur = d->d_dry_cur; P.dry_tgt = d->d_dry_tgt;
    P.send_cur = d->d_send_cur; P.send_tgt = d->d_send_tgt;
    P.dry = d->d_dry; P.wet = d->d_wet; P.partial = d->d_partial;
    P.results = want_results ? d->d_results : nullptr;
    for(int i = 0;i < 3;++i) P.bsinc_tab[i] = d->d_bsinc[i];
    for(int i = 0;i < 2;++i) P.cubic_tab[i] = d->d_cubic[i];
    P.max_voices = nv; P.frames = frames; P.ir = dd.ir_size; P.ir_pad = d->ir_pad;
    P.cd = dd.dry_channels; P.cw = dd.wet_channels; P.num_sends = dd.num_sends;
    P.max_buffers = dd.max_buffers;
    P.order = d->d_order; P.num_order = d->num_order;
    P.xscratch = d->d_xscratch; P.sendinfo = d->d_sendinfo;
    P.filt = d->d_filt; P.filt_paths = 1u + dd.num_sends;
    P.qhdr = d->d_qhdr; P.queue = d->d_queue;
    P.gather_only = d->mix_gather_only ? 1u : 0u;
    stage_mark(d, 1);
    if(d->profile) cudaEventRecord(d->ev_mix0, d->stream);
    var.fn<<<blocks, var.gs*var.groups, var.smem, d->stream>>>(P);
    if(d->profile) { cudaEventRecord(d->ev_mix1, d->stream); d->ev_valid = true; }
    ++d->launches;
    CUDA_TRY(d, cudaGetLastError());

    stage_mark(d, 2);
    // ---- voices with an active direct filter: filter the parked lines, then mix them ----
    size_t rows2 = 0;
    if(d->d_filt && d->num_order2)
    {
        FilterRunParams FP{};
        FP.filt = d->d_filt; FP.filt_paths = 1u + dd.num_sends; FP.sendinfo = d->d_sendinfo;
        FP.direct_order = d->d_order2; FP.num_direct = d->num_order2;
        FP.xscratch = d->d_xscratch; FP.dline = d->d_dline; FP.frames = frames;
        k_filters<<<(d->num_order2 + 31u)/32u, 32, 0, d->stream>>>(FP);
        ++d->launches;
        const uint32_t blocks2 = std::max(1u, std::min(maxBlocks, (d->num_order2 + var.groups - 1)/var.groups));
        rows2 = blocks2;
        MixParams P2 = P;
        P2.pass = 1u; P2.dline = d->d_dline;
        P2.order = d->d_order2; P2.num_order = d->num_order2;
        P2.partial = d->d_partial + d->partial_floats;
        P2.results = nullptr;
        var.fn<<<blocks2, var.gs*var.groups, var.smem, d->stream>>>(P2);
        ++d->launches;
        CUDA_TRY(d, cudaGetLastError());
    }

    stage_mark(d, 3);
    if(var.hrtf)
    {
        const uint32_t len = 2*kAccumLen;
        k_reduce_rows<<<(len/4 + kReduceCols - 1)/kReduceCols, 1024, 0, d->stream>>>(d->d_partial, uint32_t(rows), len,
            d->d_accum_sum, 0);
        ++d->launches;
        if(rows2)
        {
            k_reduce_rows<<<(len/4 + kReduceCols - 1)/kReduceCols, 1024, 0, d->stream>>>(d->d_partial + d->partial_floats,
                uint32_t(rows2), len, d->d_accum_sum, 1);
            ++d->launches;
        }
    }
    if(var.cdr > 0)
    {
        const uint32_t len = uint32_t(var.cdr)*kLine;
        const float *pd = d->d_partial + (var.hrtf ? rows*(2*kAccumLen) : 0);
        k_reduce_rows<<<(len/4 + kReduceCols - 1)/kReduceCols, 1024, 0, d->stream>>>(pd, uint32_t(rows), len, d->d_dry, 1);
        ++d->launches;
        if(rows2)
        {
            const float *pd2 = d->d_partial + d->partial_floats + (var.hrtf ? rows2*(2*kAccumLen) : 0);
            k_reduce_rows<<<(len/4 + kReduceCols - 1)/kReduceCols, 1024, 0, d->stream>>>(pd2, uint32_t(rows2), len, d->d_dry, 1);
            ++d->launches;
        }
    }
    CUDA_TRY(d, cudaGetLastError());

    stage_mark(d, 4);
    // ---- parked dry bus: non-HRTF voices of a variant without register accumulators ----
    if(var.cdr == 0 && d->d_dry_entries)
    {
        if(d->dry_entries_dirty)
        {
            d->h_dry_entries.clear();
            for(uint32_t v = 0;v < d->voice_hi;++v)
                if(d->h_active[v] && !d->h_hrtf[v]) d->h_dry_entries.push_back(SendEntry{v, 0u});
            d->num_dry_entries = uint32_t(d->h_dry_entries.size());
            const uint32_t ss[2] = {0u, d->num_dry_entries};
            CUDA_TRY(d, cudaMemcpyAsync(d->d_dry_slot_start, ss, sizeof(ss), cudaMemcpyHostToDevice, d->stream));
            if(d->num_dry_entries)
                CUDA_TRY(d, cudaMemcpyAsync(d->d_dry_entries, d->h_dry_entries.data(),
                    d->num_dry_entries*sizeof(SendEntry), cudaMemcpyHostToDevice, d->stream));
            CUDA_TRY(d, cudaStreamSynchronize(d->stream));
            d->dry_entries_dirty = false;
        }
        if(d->num_dry_entries)
        {
            SendMixParams DM{};
            DM.slot_start = d->d_dry_slot_start; DM.entries = d->d_dry_entries; DM.sendinfo = d->d_sendinfo;
            DM.xscratch = d->d_xscratch; DM.send_cur = d->d_dry_cur; DM.send_tgt = d->d_dry_tgt;
            DM.wet = d->d_dry; DM.frames = frames; DM.cw = dd.dry_channels; DM.num_sends = 1;
            DM.valid_bit = kSiDry; DM.dline = d->d_dline ? d->d_dline : d->d_xscratch;
            // a CTA's 8 warps share its entries evenly: chunks of 64 entries keep the first
            // (fading) tile's serial work per warp short
            const uint32_t chunks = std::max(1u, std::min(kDryChunksMax, (d->num_dry_entries + 63u)/64u));
            DM.chunks = chunks; DM.partial = d->d_dry_partial; DM.geff = d->d_dry_geff; DM.gramp = d->d_dry_gramp;
            {
                const uint32_t tot = d->num_dry_entries*dd.dry_channels;
                k_send_gains_prepare<<<(tot + 127)/128, 128, 0, d->stream>>>(DM, d->num_dry_entries);
                ++d->launches;
            }
            // Above 4 dry channels (third-order output) a full update's pan-mix past the gain fades
            // is a dense GEMM over the voices: samples 128..1023 go to the tensor cores
            // (k_panmix_tc), k_send_mix keeps the first tile with the fades
            const bool tc = d->panmix_tc && dd.dry_channels > 4u && dd.dry_channels <= uint32_t(kPmN)
                && frames == uint32_t(kLine) && chunks > 1u;
            const uint32_t tiles = tc ? 1u : (chunks > 1u ? uint32_t(kLine/128) : (frames + 127u)/128u);
            if(dd.dry_channels > 4u) k_send_mix<16><<<dim3(1, tiles, chunks), 256, 0, d->stream>>>(DM);
            else k_send_mix<4><<<dim3(1, tiles, chunks), 256, 0, d->stream>>>(DM);
            ++d->launches;
            if(tc)
            {
                PanMixTcParams TQ{d->d_dry_slot_start, d->d_dry_entries, d->d_sendinfo, d->d_xscratch,
                    d->d_dline, d->d_dry_geff, dd.dry_channels, chunks, d->d_dry_partial};
                k_panmix_tc<<<chunks, 128, kPmStages*kPmStageBytes + 1024, d->stream>>>(TQ);
                ++d->launches;
            }
            if(chunks > 1u)
            {
                const uint32_t len = dd.dry_channels*kLine;
                if(chunks <= 16u)
                    k_reduce_few<<<(len/4 + 255)/256, 256, 0, d->stream>>>(d->d_dry_partial, chunks, len, d->d_dry, 1);
                else
                k_reduce_rows<<<(len/4 + kReduceCols - 1)/kReduceCols, 1024, 0, d->stream>>>(
                    d->d_dry_partial, chunks, len, d->d_dry, 1);
                ++d->launches;
            }
            const uint32_t tot = d->num_dry_entries*dd.dry_channels;
            k_send_gains_update<<<(tot + 127)/128, 128, 0, d->stream>>>(DM, d->num_dry_entries);
            ++d->launches;
            CUDA_TRY(d, cudaGetLastError());
        }
    }

    stage_mark(d, 5);
    // ---- aux sends (core/voice.cpp:967-980) ----
    if((d->active_slots || force_sends) && d->d_wet && d->d_slot_start)
    {
        if(d->sends_dirty)
        {
            // CSR of (voice, send) pairs per slot, voices in index order (deterministic sums)
            d->h_slot_start.assign(dd.max_slots + 1, 0);
            d->h_entries.clear();
            for(uint32_t sl = 0;sl < dd.max_slots;++sl)
            {
                d->h_slot_start[sl] = uint32_t(d->h_entries.size());
                for(uint32_t v = 0;v < d->voice_hi;++v)
                    for(uint32_t s2 = 0;s2 < dd.num_sends;++s2)
                        if(d->h_send_slot[size_t(v)*B200MIX_MAX_SENDS + s2] == sl)
                            d->h_entries.push_back(SendEntry{v, s2});
            }
            d->h_slot_start[dd.max_slots] = uint32_t(d->h_entries.size());
            d->num_entries = uint32_t(d->h_entries.size());
            d->max_slot_entries = 0;
            for(uint32_t sl = 0;sl < dd.max_slots;++sl)
                d->max_slot_entries = std::max(d->max_slot_entries, d->h_slot_start[sl+1] - d->h_slot_start[sl]);
            CUDA_TRY(d, cudaMemcpyAsync(d->d_slot_start, d->h_slot_start.data(),
                (dd.max_slots + 1)*sizeof(uint32_t), cudaMemcpyHostToDevice, d->stream));
            if(d->num_entries)
                CUDA_TRY(d, cudaMemcpyAsync(d->d_entries, d->h_entries.data(),
                    d->num_entries*sizeof(SendEntry), cudaMemcpyHostToDevice, d->stream));
            CUDA_TRY(d, cudaStreamSynchronize(d->stream));
            d->sends_dirty = false;
        }
        SendMixParams SM{};
        SM.slot_start = d->d_slot_start; SM.entries = d->d_entries; SM.sendinfo = d->d_sendinfo;
        SM.xscratch = d->d_xscratch; SM.send_cur = d->d_send_cur; SM.send_tgt = d->d_send_tgt;
        SM.wet = d->d_wet; SM.frames = frames; SM.cw = dd.wet_channels; SM.num_sends = dd.num_sends;
        SM.valid_bit = kSiSend; SM.chunks = 1;
        if(d->d_filt && d->num_entries)
        {
            if(d->fscratch_rows < d->num_entries)
            {
                CUDA_TRY(d, cudaStreamSynchronize(d->stream));
                cudaFree(d->d_fscratch); d->d_fscratch = nullptr; d->fscratch_rows = 0;
                const uint32_t rows = std::max(d->num_entries, 64u);
                CUDA_TRY(d, cudaMalloc(reinterpret_cast<void**>(&d->d_fscratch), size_t(rows)*kLine*sizeof(float)));
                CUDA_TRY(d, cudaMemsetAsync(d->d_fscratch, 0, size_t(rows)*kLine*sizeof(float), d->stream));
                d->fscratch_rows = rows;
            }
            SM.filt = d->d_filt; SM.filt_paths = 1u + dd.num_sends; SM.fscratch = d->d_fscratch;
            FilterRunParams FP{};
            FP.filt = d->d_filt; FP.filt_paths = 1u + dd.num_sends; FP.sendinfo = d->d_sendinfo;
            FP.entries = d->d_entries; FP.num_entries = d->num_entries;
            FP.xscratch = d->d_xscratch; FP.fscratch = d->d_fscratch; FP.frames = frames;
            k_filters<<<(d->num_entries + 31u)/32u, 32, 0, d->stream>>>(FP);
            ++d->launches;
        }
        SM.geff = d->d_send_geff; SM.gramp = d->d_send_gramp;
        if(d->num_entries)
        {
            const uint32_t tot = d->num_entries*dd.wet_channels;
            k_send_gains_prepare<<<(tot + 127)/128, 128, 0, d->stream>>>(SM, d->num_entries);
            ++d->launches;
        }
        {
            // a CTA's 8 warps share its entries evenly: chunks of 128 entries per slot
            const uint32_t chunks = std::max(1u, std::min(16u, (d->max_slot_entries + 127u)/128u));
            if(chunks > 1u && d->send_partial_chunks < chunks)
            {
                CUDA_TRY(d, cudaStreamSynchronize(d->stream));
                cudaFree(d->d_send_partial); d->d_send_partial = nullptr; d->send_partial_chunks = 0;
                CUDA_TRY(d, cudaMalloc(reinterpret_cast<void**>(&d->d_send_partial),
                    size_t(chunks)*dd.max_slots*dd.wet_channels*kLine*sizeof(float)));
                d->send_partial_chunks = chunks;
            }
            SM.chunks = chunks; SM.partial = d->d_send_partial;
            const uint32_t tiles = chunks > 1u ? uint32_t(kLine/128) : (frames + 127u)/128u;
            if(dd.wet_channels > 4u) k_send_mix<16><<<dim3(dd.max_slots, tiles, chunks), 256, 0, d->stream>>>(SM);
            else k_send_mix<4><<<dim3(dd.max_slots, tiles, chunks), 256, 0, d->stream>>>(SM);
            ++d->launches;
            if(chunks > 1u)
            {
                const uint32_t len = dd.max_slots*dd.wet_channels*kLine;
                if(chunks <= 16u)
                    k_reduce_few<<<(len/4 + 255)/256, 256, 0, d->stream>>>(d->d_send_partial, chunks, len, d->d_wet, 1);
                else
                k_reduce_rows<<<(len/4 + kReduceCols - 1)/kReduceCols, 1024, 0, d->stream>>>(
                    d->d_send_partial, chunks, len, d->d_wet, 1);
                ++d->launches;
            }
        }
        if(d->num_entries)
        {
            const uint32_t tot = d->num_entries*dd.wet_channels;
            k_send_gains_update<<<(tot + 127)/128, 128, 0, d->stream>>>(SM, d->num_entries);
            ++d->launches;
        }
        CUDA_TRY(d, cudaGetLastError());
    }
    return B200MIX_OK;
}

// Phase B: run the effect slots on their Wet input, mix their output into Dry, post-process
// (alc/alu.cpp:2252-2256, 2439-2443).
static int render_phase_b(b200mix_device *d, uint32_t frames)
{
    const b200mix_device_desc &dd = d->desc;
    stage_mark(d, 6);
    if(d->active_slots)
    {
        ConvParams CP{};
        CP.slots = d->d_slots; CP.wet = d->d_wet; CP.twiddle = d->d_twiddle;
        CP.frames = frames; CP.cw = dd.wet_channels; CP.num_slots = dd.max_slots;
        uint32_t convCh = 1, convWork = 0, convSegs = 0;
        for(const SlotRec &sr : d->h_slots)
            if(sr.type == B200MIX_EFFECT_CONVOLUTION)
            { convCh = std::max(convCh, sr.channels); convWork += sr.channels; convSegs = std::max(convSegs, sr.segs); }
        // segment chunks of k_conv_mac: ~4 CTAs (of 128 threads) per SM over all convolution
        // slot-channels, at least 18 segments per chunk
        CP.chunks = 1u;
        if(convWork)
            CP.chunks = std::max(1u, std::min(std::min(uint32_t(kConvMaxChunks), (convSegs + 17u)/18u),
                (4u*uint32_t(d->num_sms) + convWork - 1u)/convWork));
        if(d->reverb_slots)
        {
            // ReverbState::process's pipeline state machine (reverb.cpp:1840-1878), host side
            for(uint32_t sl = 0;sl < d->rv.size();++sl)
            {
                b200mix_device::RvHost &R = d->rv[sl];
                if(!R.used || d->h_slots[sl].type != B200MIX_EFFECT_REVERB) continue;
                uint32_t mask = 1u << R.cur;
                if(R.state < 2) R.state = 2;                         // StartFade -> Fading
                if(R.state != 4)
                {
                    const int old = R.cur ^ 1;
                    if(R.state == 3)
                    {
                        if(int rc = reverb_clear_pipeline(d, sl, old)) return rc;
                        R.state = 4;
                    }
                    else
                    {
                        if(frames >= R.fade[old])
                        {
                            // final mix of the old pipeline: its gains fade to silence
                            CUDA_TRY(d, cudaMemsetAsync(d->h_slots[sl].gtgt + size_t(old)*8*32, 0,
                                size_t(8)*32*sizeof(float), d->stream));
                            R.fade[old] = 0; R.state = 3;
                        }
                        else R.fade[old] -= frames;
                        mask |= 1u << old;
                    }
                }
                R.offset += frames;
                if(d->h_slots[sl].rv_mask != mask || d->h_slots[sl].rv_cur != uint32_t(R.cur))
                {
                    d->h_slots[sl].rv_mask = mask; d->h_slots[sl].rv_cur = uint32_t(R.cur);
                    CUDA_TRY(d, cudaMemcpyAsync(d->d_slots + sl, &d->h_slots[sl], sizeof(SlotRec),
                        cudaMemcpyHostToDevice, d->stream));
                    CUDA_TRY(d, cudaStreamSynchronize(d->stream));
                }
            }
        }
        SlotMixParams SP{};
        SP.slots = d->d_slots; SP.dry = d->d_dry; SP.frames = frames; SP.cd = dd.dry_channels;
        SP.num_slots = dd.max_slots; SP.wet = d->d_wet; SP.cw = dd.wet_channels;
        // one pass per stage of the slot graph (a single pass unless slots target other slots)
        for(uint32_t st = 0;st < d->num_stages;++st)
        {
            if(d->reverb_slots)
            {
                ReverbParamsK RP{};
                RP.slots = d->d_slots; RP.wet = d->d_wet; RP.cubic = d->d_cubic_filter;
                RP.frames = frames; RP.cw = dd.wet_channels; RP.stage = st;
                RP.seq = ++d->reverb_seq;
                k_reverb_process<<<dim3(dd.max_slots, 2, 2), 128, 0, d->stream>>>(RP);
                k_reverb_commit<<<dd.max_slots, 2, 0, d->stream>>>(RP);
                d->launches += 2;
                if(d->reverb_upmix)
                {
                    k_reverb_upmix<<<dim3(dd.max_slots, 2), 256, 8*kLine*sizeof(float), d->stream>>>(RP);
                    ++d->launches;
                }
            }
            CP.stage = st; SP.stage = st;
            if(d->efx_slots)
            {
                EfxRunParams EQ{d->d_efx_views, d->d_wet, frames, dd.wet_channels, st, d->d_cubic_filter};
                CUDA_TRY(d, launch_efx_process(EQ, dd.max_slots, d->stream));
                ++d->launches;
                if(d->pshift_slots)
                {
                    CUDA_TRY(d, launch_efx_pshift(EQ, dd.max_slots, d->stream));
                    ++d->launches;
                }
            }
            if(convWork)
            {
                k_conv_input<<<dd.max_slots, 128, 0, d->stream>>>(CP);
                k_conv_mac<<<dim3(dd.max_slots, convCh, CP.chunks), 128, sizeof(ConvMacSmem), d->stream>>>(CP);
                k_conv_ifft<<<dim3(dd.max_slots, convCh, kConvMaxBlocks), 128, 0, d->stream>>>(CP);
                k_conv_output<<<dim3(dd.max_slots, convCh), 128, 0, d->stream>>>(CP);
                d->launches += 4;
            }
            k_slot_output_mix<<<dim3((frames + 127)/128, dd.dry_channels), 128, 0, d->stream>>>(SP);
            ++d->launches;
            if(d->any_target)
            {
                k_slot_target_mix<<<dim3((frames + 127)/128, dd.max_slots), 128, 0, d->stream>>>(SP);
                ++d->launches;
            }
        }
        k_slot_gains_commit<<<dd.max_slots, 64, 0, d->stream>>>(SP);
        ++d->launches;
        for(auto &eh : d->efx)
        {
            if(eh.used && eh.p.type == B200MIX_EFFECT_MODULATOR) eh.mod_index = (eh.mod_index + frames) % eh.mod_range;
            if(eh.used && eh.p.type == B200MIX_EFFECT_CHORUS) eh.lfo_offset = (eh.lfo_offset + frames) % eh.lfo_range;
        }
        CUDA_TRY(d, cudaGetLastError());
    }

    stage_mark(d, 7);
    switch(dd.post_process)
    {
    case B200MIX_POST_HRTF:
    {
        PostHrtfParams Q{};
        Q.accum_sum = d->d_accum_sum; Q.carry_in = d->d_carry[d->carry_idx];
        Q.carry_out = d->d_carry[d->carry_idx^1];
        Q.dry = d->d_dry; Q.real = d->d_real; Q.dec_coef = d->d_dec_coef;
        Q.dec_hfscale = d->d_dec_hfscale; Q.dec_state = d->d_dec_state; Q.temp = d->d_temp;
        Q.frames = frames; Q.cd = dd.dry_channels; Q.dec_ir = d->dec_ir;
        Q.real_left = dd.real_left; Q.real_right = dd.real_right; Q.dry_active = d->dry_active;
        if(d->dry_active)
        {
            k_post_hrtf_split<<<dd.dry_channels, 32, 0, d->stream>>>(Q);
            ++d->launches;
        }
        Q.overwrite = d->real_overwrite ? 1u : 0u;
        k_post_hrtf_mix<<<dim3((frames + kHrirLen + 127)/128, 2), 128, 0, d->stream>>>(Q);
        ++d->launches;
        d->carry_idx ^= 1;
        break;
    }
    case B200MIX_POST_AMBIDEC:
    {
        PostAmbiParams Q{};
        Q.dry = d->d_dry; Q.real = d->d_real; Q.gains_hf = d->d_amb_hf; Q.gains_lf = d->d_amb_lf;
        Q.split_state = d->d_amb_state; Q.temp_hf = d->d_temp; Q.temp_lf = d->d_temp2;
        Q.frames = frames; Q.cd = dd.dry_channels; Q.real_channels = dd.real_channels;
        Q.dual = d->amb_dual;
        if(d->amb_dual)
        {
            k_post_ambi_split<<<1, 32, 0, d->stream>>>(Q);
            ++d->launches;
        }
        const uint32_t total = dd.real_channels*frames;
        k_post_ambi_mix<<<(total + 127)/128, 128, 0, d->stream>>>(Q);
        ++d->launches;
        if(d->stab_center != B200MIX_NO_SLOT)
        {
            StabParams S{};
            S.real = d->d_real; S.state = d->d_stab_state; S.frames = frames;
            S.real_channels = dd.real_channels; S.lidx = dd.real_left; S.ridx = dd.real_right;
            S.cidx = d->stab_center; S.coeff = d->stab_coeff;
            const float halfPi = 3.14159265358979323846f*0.5f;
            S.mid_lf = std::cos(1.0f/3.0f * halfPi); S.mid_hf = std::cos(1.0f/4.0f * halfPi);
            S.center_lf = std::sin(1.0f/3.0f * halfPi); S.center_hf = std::sin(1.0f/4.0f * halfPi);
            const size_t smem = size_t(2u + dd.real_channels)*kLine*sizeof(float);
            k_post_stabilizer<<<1, 32u*dd.real_channels, smem, d->stream>>>(S);
            ++d->launches;
        }
        else if(d->bs2b_level)
        {
            // RealOut holds nothing but the decode here (no direct-channel voices), so the
            // copy-out / add-back of the direct signal around the filter (alc/alu.cpp:414-433)
            // has nothing to move
            Bs2bParams B{d->d_real, d->d_bs2b, d->d_bs2b + 4, frames, dd.real_left, dd.real_right};
            k_post_bs2b<<<1, 128, 0, d->stream>>>(B);
            ++d->launches;
        }
        break;
    }
    case B200MIX_POST_UHJ: case B200MIX_POST_TSME:
    {
        const MatrixEncSpec spec = dd.post_process == B200MIX_POST_TSME ? kTsmeEncSpec : kUhjEncSpec;
        PostUhjParams Q{};
        Q.dry = d->d_dry; Q.real = d->d_real; Q.state = d->d_uhj_state; Q.scratch = d->d_uhj_scratch;
        Q.frames = frames; Q.real_left = dd.real_left; Q.real_right = dd.real_right; Q.enc = spec;
        if(d->uhj_fir)
        {
            PostUhjFirParams F{d->d_dry, d->d_real, d->d_uhj_fir_state, d->d_uhj_fir_coef, frames,
                dd.real_left, dd.real_right, d->uhj_fir, spec};
            k_post_uhj_fir<<<1, 1024, 0, d->stream>>>(F);
        }
        else
            k_post_uhj<<<1, 1024, 0, d->stream>>>(Q);
        ++d->launches;
        break;
    }
    default: break;
    }
    stage_mark(d, 8);
    if(d->profile_level >= 2) d->stage_valid = true;
    CUDA_TRY(d, cudaGetLastError());
    return B200MIX_OK;
}

// The nonlinear output stage (limiter, speaker distance compensation): after the RealOut
// reduce of a sharded device set, on the root only (alc/alu.cpp:2446-2450).
static int render_output_stage(b200mix_device *d, uint32_t frames)
{
    const b200mix_device_desc &dd = d->desc;
    if(d->shard.transport && d->shard.rank != 0u) return B200MIX_OK;
    if(d->d_limiter)
    {
        // if(Limiter) Limiter->process(samplesToDo, RealOut.Buffer), alc/alu.cpp:2446
        LimiterParams LQ{d->d_limiter, d->d_real, d->d_limiter_delay, frames};
        k_limiter<<<1, 1024, 0, d->stream>>>(LQ);
        ++d->launches;
    }
    if(d->d_dc_delay)
    {
        // if(ChannelDelays) ApplyDistanceComp(RealOut.Buffer, ...), alc/alu.cpp:2449-2450
        DistCompParams DQ{d->d_real, d->d_dc_buf, d->d_dc_delay, d->d_dc_gain, frames};
        k_distance_comp<<<dd.real_channels, 1024, 0, d->stream>>>(DQ);
        ++d->launches;
    }
    CUDA_TRY(d, cudaGetLastError());
    return B200MIX_OK;
}

// ---- voice-sharded device sets: the two exchanges of an update -------------
// Wet reduce-scatter: every owner ends up with the summed send input of its slots.
static int shard_wet_exchange(b200mix_device *d)
{
    b200mix_device::Shard &S = d->shard;
    const b200mix_device_desc &dd = d->desc;
    if(!S.transport || !d->d_wet) return B200MIX_OK;
    if(d->profile) { cudaEventRecord(S.ev[0], d->stream); }
    const size_t wetFloats = size_t(dd.max_slots)*dd.wet_channels*kLine;
    if(S.transport == 2)
    {
        if(S.allreduce(d->d_wet, d->d_wet, wetFloats, 7 /*ncclFloat*/, 0 /*ncclSum*/, S.comm, d->stream) != 0)
        { d->error = "ncclAllReduce of the wet buffers failed"; return B200MIX_ERR_CUDA; }
    }
    else
    {
        const uint32_t slotFloats = dd.wet_channels*uint32_t(kLine);
        ShardPushParams P{};
        P.src = d->d_wet; P.rank = S.rank; P.world = S.world; P.epoch = S.epoch;
        P.wet = 1u; P.num_slots = dd.max_slots; P.slot_floats = slotFloats; P.owned_max = S.owned_max;
        P.own = reinterpret_cast<ShardCtl*>(S.own);
        for(uint32_t r = 0;r < S.world;++r) P.peer[r] = S.peer[r];
        P.off_data = S.off_wet; P.per_src_floats = S.wet_src_floats;
        P.counters = S.d_counters + 4;
        const uint32_t chunks = std::max(1u, std::min(32u, slotFloats*S.owned_max/(4u*256u*4u)));
        k_shard_push<<<dim3(chunks, S.world), 256, 0, d->stream>>>(P);
        ShardSumParams Q{};
        Q.dst = d->d_wet; Q.rank = S.rank; Q.world = S.world; Q.epoch = S.epoch;
        Q.wet = 1u; Q.num_slots = dd.max_slots; Q.slot_floats = slotFloats; Q.owned_max = S.owned_max;
        Q.own = P.own; for(uint32_t r = 0;r < S.world;++r) Q.peer[r] = S.peer[r];
        Q.off_data = S.off_wet; Q.per_src_floats = S.wet_src_floats; Q.counter = S.d_counters + 2;
        k_shard_sum<<<std::max(1u, std::min(64u, slotFloats*S.owned_max/(4u*256u*2u))), 256, 0, d->stream>>>(Q);
        d->launches += 2;
        CUDA_TRY(d, cudaGetLastError());
    }
    if(d->profile) { cudaEventRecord(S.ev[1], d->stream); S.ev_wet = true; }
    return B200MIX_OK;
}

// RealOut reduce onto rank 0.
static int shard_real_reduce(b200mix_device *d)
{
    b200mix_device::Shard &S = d->shard;
    const b200mix_device_desc &dd = d->desc;
    if(!S.transport) return B200MIX_OK;
    if(d->profile) { cudaEventRecord(S.ev[2], d->stream); }
    const uint32_t floats = dd.real_channels*uint32_t(kLine);
    if(S.transport == 2)
    {
        if(S.reduce(d->d_real, d->d_real, floats, 7, 0, 0, S.comm, d->stream) != 0)
        { d->error = "ncclReduce of RealOut failed"; return B200MIX_ERR_CUDA; }
    }
    else if(S.rank != 0u)
    {
        ShardPushParams P{};
        P.src = d->d_real; P.rank = S.rank; P.world = S.world; P.epoch = S.epoch; P.floats = floats;
        P.own = reinterpret_cast<ShardCtl*>(S.own);
        for(uint32_t r = 0;r < S.world;++r) P.peer[r] = S.peer[r];
        P.off_data = S.off_real; P.per_src_floats = S.real_floats; P.counters = S.d_counters;
        k_shard_push<<<dim3(std::max(1u, floats/(4u*256u*2u)), 1), 256, 0, d->stream>>>(P);
        ++d->launches;
    }
    else
    {
        ShardSumParams Q{};
        Q.dst = d->d_real; Q.rank = 0u; Q.world = S.world; Q.epoch = S.epoch; Q.floats = floats;
        Q.own = reinterpret_cast<ShardCtl*>(S.own);
        for(uint32_t r = 0;r < S.world;++r) Q.peer[r] = S.peer[r];
        Q.off_data = S.off_real; Q.per_src_floats = S.real_floats; Q.counter = S.d_counters + 1;
        k_shard_sum<<<std::max(1u, floats/(4u*256u*2u)), 256, 0, d->stream>>>(Q);
        ++d->launches;
    }
    CUDA_TRY(d, cudaGetLastError());
    if(d->profile) { cudaEventRecord(S.ev[3], d->stream); S.ev_real = true; }
    return B200MIX_OK;
}

static int render_launch(b200mix_device *d, uint32_t frames, bool want_results)
{
    if(d->mid_render) { d->error = "render: a render_begin is pending"; return B200MIX_ERR_INVALID; }
    const bool sharded = d->shard.transport != 0;
    if(sharded) ++d->shard.epoch;
    // a sharded set always finishes its sends: another rank may own the slots they feed
    if(int rc = render_phase_a(d, frames, want_results, sharded)) return rc;
    if(sharded) if(int rc = shard_wet_exchange(d)) return rc;
    if(int rc = render_phase_b(d, frames)) return rc;
    if(sharded) if(int rc = shard_real_reduce(d)) return rc;
    return render_output_stage(d, frames);
}

static int render_collect(b200mix_device *d, uint32_t frames, float *const *real_out,
    b200mix_voice_result *results)
{
    const b200mix_device_desc &dd = d->desc;
    const uint32_t nv = std::max(d->voice_hi, 1u);
    const bool contiguous = d->d_real == reinterpret_cast<float*>(d->d_outblock);
    if(real_out && results && contiguous)
        CUDA_TRY(d, cudaMemcpyAsync(d->h_outblock, d->d_outblock, d->out_real_bytes + size_t(nv)*sizeof(VoiceResult),
            cudaMemcpyDeviceToHost, d->stream));
    else
    {
        if(real_out)
            CUDA_TRY(d, cudaMemcpyAsync(d->h_real, d->d_real, size_t(dd.real_channels)*kLine*sizeof(float),
                cudaMemcpyDeviceToHost, d->stream));
        if(results)
            CUDA_TRY(d, cudaMemcpyAsync(d->h_results, d->d_results, size_t(nv)*sizeof(VoiceResult),
                cudaMemcpyDeviceToHost, d->stream));
    }
    if(d->shard.transport == 1)
        CUDA_TRY(d, cudaMemcpyAsync(d->shard.h_status, d->shard.own + offsetof(ShardCtl, status),
            sizeof(uint32_t), cudaMemcpyDeviceToHost, d->stream));
    CUDA_TRY(d, cudaStreamSynchronize(d->stream));
    d->stage_busy = false;
    if(d->shard.transport == 1 && *d->shard.h_status)
    { d->error = "render: a peer of the sharded device set did not answer in time"; return B200MIX_ERR_CUDA; }
    if(real_out)
        for(uint32_t c = 0;c < dd.real_channels;++c)
            if(real_out[c]) std::memcpy(real_out[c], d->h_real + size_t(c)*kLine, frames*sizeof(float));
    if(results)
    {
        std::memcpy(results, d->h_results, size_t(nv)*sizeof(b200mix_voice_result));
        for(uint32_t v = nv;v < dd.max_voices;++v)
            results[v] = b200mix_voice_result{0, 0u, B200MIX_VF_STOPPED, 0u};
    }
    return B200MIX_OK;
}

int b200mix_render(b200mix_device *d, uint32_t frames, float *const *real_out,
    b200mix_voice_result *results)
{
    if(!d) return B200MIX_ERR_INVALID;
    if(int rc = render_launch(d, frames, results != nullptr)) return rc;
    return render_collect(d, frames, real_out, results);
}

int b200mix_set_uhj_encoder(b200mix_device *d, uint32_t filter_length, uint32_t *delay)
{
    if(!d) return B200MIX_ERR_INVALID;
    const b200mix_device_desc &dd = d->desc;
    if(d->mid_render || (dd.post_process != B200MIX_POST_UHJ && dd.post_process != B200MIX_POST_TSME)
        || dd.dry_channels < 3
        || (filter_length != 0 && filter_length != 256 && filter_length != 512))
    { d->error = "set_uhj_encoder: needs a UHJ device and a length of 0, 256 or 512"; return B200MIX_ERR_INVALID; }
    CUDA_TRY(d, cudaStreamSynchronize(d->stream));
    CUDA_TRY(d, cudaMemset(d->d_uhj_state, 0, 64*sizeof(float)));
    if(filter_length)
    {
        if(!d->d_uhj_fir_state)
        {
            if(int rc = dev_alloc(d, d->d_uhj_fir_state, kUhjFirStateFloats)) return rc;
            if(int rc = dev_alloc(d, d->d_uhj_fir_coef, 256)) return rc;
            CUDA_TRY(d, cudaStreamSynchronize(d->stream));
        }
        CUDA_TRY(d, cudaMemset(d->d_uhj_fir_state, 0, kUhjFirStateFloats*sizeof(float)));
        // SegmentedFilter's desired response (core/allpass_conv.hpp:56-75): Blackman-Nuttall
        // windowed 2/(pi k) at the odd taps
        const uint32_t half = filter_length/2u;
        std::vector<float> coef(256, 0.0f);
        const double pi = 3.14159265358979323846;
        for(uint32_t i = 0;i < half;++i)
        {
            const int k = int(half) - int(i*2u + 1u);
            const double w = 2.0*pi/double(half - 1u) * double(i);
            const double window = 0.3635819 - 0.4891775*std::cos(w) + 0.1365995*std::cos(2.0*w)
                - 0.0106411*std::cos(3.0*w);
            coef[i] = float(window * 2.0 / (pi * double(k)));
        }
        CUDA_TRY(d, cudaMemcpy(d->d_uhj_fir_coef, coef.data(), 256*sizeof(float), cudaMemcpyHostToDevice));
    }
    d->uhj_fir = filter_length;
    if(delay) *delay = filter_length ? filter_length/2u + 128u : 1u;
    return B200MIX_OK;
}

int b200mix_set_front_stabilizer(b200mix_device *d, uint32_t center_channel, float splitter_coeff)
{
    if(!d) return B200MIX_ERR_INVALID;
    const b200mix_device_desc &dd = d->desc;
    if(d->mid_render) { d->error = "set_front_stabilizer: a render_begin is pending"; return B200MIX_ERR_INVALID; }
    CUDA_TRY(d, cudaStreamSynchronize(d->stream));
    if(center_channel == B200MIX_NO_SLOT) { d->stab_center = B200MIX_NO_SLOT; return B200MIX_OK; }
    if(dd.post_process != B200MIX_POST_AMBIDEC || center_channel >= dd.real_channels
        || dd.real_left == dd.real_right || dd.real_left >= dd.real_channels || dd.real_right >= dd.real_channels
        || center_channel == dd.real_left || center_channel == dd.real_right || dd.real_channels > 32u)
    { d->error = "set_front_stabilizer: needs an ambisonic-decode device with left, right and centre outputs"; return B200MIX_ERR_INVALID; }
    if(!d->d_stab_state)
    {
        if(int rc = dev_alloc(d, d->d_stab_state, 4 + 32)) return rc;
        CUDA_TRY(d, cudaStreamSynchronize(d->stream));
    }
    CUDA_TRY(d, cudaMemset(d->d_stab_state, 0, (4 + 32)*sizeof(float)));
    CUDA_TRY(d, cudaFuncSetAttribute(k_post_stabilizer, cudaFuncAttributeMaxDynamicSharedMemorySize,
        int(size_t(2u + dd.real_channels)*kLine*sizeof(float))));
    d->stab_center = center_channel; d->stab_coeff = splitter_coeff;
    return B200MIX_OK;
}

int b200mix_set_bs2b(b200mix_device *d, uint32_t level)
{
    if(!d) return B200MIX_ERR_INVALID;
    const b200mix_device_desc &dd = d->desc;
    if(d->mid_render || level > 6 || dd.post_process != B200MIX_POST_AMBIDEC || dd.real_left == dd.real_right
        || dd.real_left >= dd.real_channels || dd.real_right >= dd.real_channels)
    { d->error = "set_bs2b: needs a stereo ambisonic-decode device and a level of 0..6"; return B200MIX_ERR_INVALID; }
    CUDA_TRY(d, cudaStreamSynchronize(d->stream));
    if(!d->d_bs2b)
    {
        if(int rc = dev_alloc(d, d->d_bs2b, 16)) return rc;
        CUDA_TRY(d, cudaStreamSynchronize(d->stream));
    }
    float h[9] = {};
    if(level)
    {
        // init(), core/bs2b.cpp:41-91 (same float expressions, host libm)
        static const float tab[6][4] = {
            {360.0f,  501.0f, 0.398107170553497f, 0.205671765275719f},
            {500.0f,  711.0f, 0.459726988530872f, 0.228208484414988f},
            {700.0f, 1021.0f, 0.530884444230988f, 0.250105790667544f},
            {360.0f,  494.0f, 0.316227766016838f, 0.168236228897329f},
            {500.0f,  689.0f, 0.354813389233575f, 0.187169483835901f},
            {700.0f,  975.0f, 0.398107170553497f, 0.205671765275719f}};
        const float Fc_lo = tab[level-1][0], Fc_hi = tab[level-1][1];
        const float G_lo = tab[level-1][2], G_hi = tab[level-1][3];
        const float pi = 3.14159265358979323846f;
        const float g = 1.0f / (1.0f - G_hi + G_lo);
        float x = std::exp(-pi*2.0f*Fc_lo/float(dd.sample_rate));
        h[4+1] = x; h[4+0] = G_lo * (1.0f - x) * g;
        x = std::exp(-pi*2.0f*Fc_hi/float(dd.sample_rate));
        h[4+4] = x; h[4+2] = (1.0f - G_hi * (1.0f - x)) * g; h[4+3] = -x * g;
    }
    CUDA_TRY(d, cudaMemcpy(d->d_bs2b, h, sizeof(h), cudaMemcpyHostToDevice));
    d->bs2b_level = level;
    return B200MIX_OK;
}

int b200mix_set_distance_comp(b200mix_device *d, uint32_t channels, const uint32_t *delays, const float *gains)
{
    if(!d) return B200MIX_ERR_INVALID;
    if(d->mid_render) { d->error = "set_distance_comp: a render_begin is pending"; return B200MIX_ERR_INVALID; }
    const b200mix_device_desc &dd = d->desc;
    CUDA_TRY(d, cudaStreamSynchronize(d->stream));
    cudaFree(d->d_dc_delay); cudaFree(d->d_dc_gain); cudaFree(d->d_dc_buf);
    d->d_dc_delay = nullptr; d->d_dc_gain = nullptr; d->d_dc_buf = nullptr;
    if(!channels) return B200MIX_OK;
    if(channels > dd.real_channels || !delays || !gains)
    { d->error = "set_distance_comp: bad arguments"; return B200MIX_ERR_INVALID; }
    std::vector<uint32_t> hd(dd.real_channels, 0u);
    std::vector<float> hg(dd.real_channels, 1.0f);
    for(uint32_t c = 0;c < channels;++c)
    {
        if(delays[c] >= kLine) { d->error = "set_distance_comp: delay >= 1024"; return B200MIX_ERR_INVALID; }
        hd[c] = delays[c]; hg[c] = gains[c];
    }
    if(int rc = dev_alloc(d, d->d_dc_delay, dd.real_channels)) return rc;
    if(int rc = dev_alloc(d, d->d_dc_gain, dd.real_channels)) return rc;
    if(int rc = dev_alloc(d, d->d_dc_buf, size_t(dd.real_channels)*kLine)) return rc;
    CUDA_TRY(d, cudaStreamSynchronize(d->stream));
    CUDA_TRY(d, cudaMemcpy(d->d_dc_delay, hd.data(), hd.size()*sizeof(uint32_t), cudaMemcpyHostToDevice));
    CUDA_TRY(d, cudaMemcpy(d->d_dc_gain, hg.data(), hg.size()*sizeof(float), cudaMemcpyHostToDevice));
    return B200MIX_OK;
}

// Compressor::Create (core/mastering.cpp:108-166): the same float/double expressions, by the
// host's libm, so the derived constants are the reference's bit for bit.
int b200mix_set_limiter(b200mix_device *d, const b200mix_limiter_desc *p, uint32_t *look_ahead)
{
    if(!d) return B200MIX_ERR_INVALID;
    if(look_ahead) *look_ahead = 0;
    if(d->mid_render) { d->error = "set_limiter: a render_begin is pending"; return B200MIX_ERR_INVALID; }
    const b200mix_device_desc &dd = d->desc;
    if(!p)
    {
        CUDA_TRY(d, cudaStreamSynchronize(d->stream));
        cudaFree(d->d_limiter); d->d_limiter = nullptr;
        cudaFree(d->d_limiter_delay); d->d_limiter_delay = nullptr;
        return B200MIX_OK;
    }
    if(p->struct_size != sizeof(*p)) { d->error = "set_limiter: struct_size"; return B200MIX_ERR_INVALID; }
    const float rate = float(dd.sample_rate);
    auto clampf = [](float v, float lo, float hi) { return v < lo ? lo : (hi < v ? hi : v); };
    LimiterDev h{};
    h.look_ahead = uint32_t(clampf(std::round(p->look_ahead_time*rate), 0.0f, float(kLine) - 1.0f));
    const uint32_t hold = uint32_t(clampf(std::round(p->hold_time*rate), 0.0f, float(kLine) - 1.0f));
    h.flags = p->auto_flags & 31u;
    if(!(h.flags & B200MIX_LIM_AUTO_POSTGAIN)) h.flags &= ~uint32_t(B200MIX_LIM_AUTO_DECLIP);
    h.num_chans = dd.real_channels;
    h.pre_gain = std::pow(10.0f, p->pre_gain_db / 20.0f);
    h.post_gain = float(std::log(10.0)/20.0 * double(p->post_gain_db));
    h.threshold = float(std::log(10.0)/20.0 * double(p->threshold_db));
    h.slope = 1.0f/std::max(1.0f, p->ratio) - 1.0f;
    h.knee = float(std::max(0.0, std::log(10.0)/20.0 * double(p->knee_db)));
    h.attack = std::max(1.0f, p->attack_time * rate);
    h.release = std::max(1.0f, p->release_time * rate);
    if(h.flags & B200MIX_LIM_AUTO_KNEE) h.slope = -1.0f;
    // the hold needs a look-ahead and more than one sample (:141-153)
    h.hold = (h.look_ahead > 0 && hold > 1) ? hold : 0u;
    h.crest_coeff = std::exp(-1.0f / (0.200f * rate));
    h.gain_estimate = h.threshold * -0.5f * h.slope;
    h.adapt_coeff = std::exp(-1.0f / (2.0f * rate));
    for(float &v : h.hold_hist) v = -INFINITY;
    if(!d->d_limiter)
    {
        if(int rc = dev_alloc(d, d->d_limiter, 1)) return rc;
        if(int rc = dev_alloc(d, d->d_limiter_delay, size_t(std::max(dd.real_channels, 1u))*kLine)) return rc;
    }
    CUDA_TRY(d, cudaStreamSynchronize(d->stream));
    CUDA_TRY(d, cudaMemcpy(d->d_limiter, &h, sizeof(h), cudaMemcpyHostToDevice));
    CUDA_TRY(d, cudaMemset(d->d_limiter_delay, 0, size_t(std::max(dd.real_channels, 1u))*kLine*sizeof(float)));
    if(look_ahead) *look_ahead = h.look_ahead;
    return B200MIX_OK;
}

static uint32_t lcg_skip_host(uint32_t x, uint64_t k)
{
    uint32_t a = 96314165u, c = 907633515u, accA = 1u, accC = 0u;
    while(k)
    {
        if(k & 1u) { accA = accA*a; accC = accC*a + c; }
        c = c*a + c; a = a*a;
        k >>= 1;
    }
    return accA*x + accC;
}

int b200mix_render_interleaved(b200mix_device *d, uint32_t frames, void *out, uint32_t out_type,
    uint32_t frame_step, float dither_depth, uint32_t *dither_seed, b200mix_voice_result *results)
{
    if(!d) return B200MIX_ERR_INVALID;
    const b200mix_device_desc &dd = d->desc;
    if(!out || out_type > B200MIX_OUT_F32 || frame_step < dd.real_channels || frame_step > 64u
        || (dither_depth > 0.0f && !dither_seed))
    { d->error = "render_interleaved: bad arguments"; return B200MIX_ERR_INVALID; }
    if(int rc = render_launch(d, frames, results != nullptr)) return rc;
    static const size_t sz[] = {1, 1, 2, 2, 4, 4, 4};
    if(!d->d_outbuf)
    {
        CUDA_TRY(d, cudaMalloc(&d->d_outbuf, size_t(kLine)*64*4));
        CUDA_TRY(d, cudaMallocHost(&d->h_outbuf, size_t(kLine)*64*4));
    }
    OutputParams Q{};
    Q.real = d->d_real; Q.out = d->d_outbuf; Q.frames = frames; Q.channels = dd.real_channels;
    Q.frame_step = frame_step; Q.out_type = out_type; Q.dither_depth = dither_depth;
    Q.seed = dither_seed ? *dither_seed : 0u;
    const uint32_t total = frames*frame_step;
    k_output_write<<<(total + 255)/256, 256, 0, d->stream>>>(Q);
    ++d->launches;
    CUDA_TRY(d, cudaGetLastError());
    const size_t bytes = size_t(total)*sz[out_type];
    CUDA_TRY(d, cudaMemcpyAsync(d->h_outbuf, d->d_outbuf, bytes, cudaMemcpyDeviceToHost, d->stream));
    if(int rc = render_collect(d, frames, nullptr, results)) return rc;     // synchronises the stream
    std::memcpy(out, d->h_outbuf, bytes);
    if(dither_depth > 0.0f) *dither_seed = lcg_skip_host(*dither_seed, uint64_t(2)*dd.real_channels*frames);
    return B200MIX_OK;
}

int b200mix_render_begin(b200mix_device *d, uint32_t frames, float **wet_dev, size_t *wet_floats)
{
    if(!d) return B200MIX_ERR_INVALID;
    if(d->mid_render) { d->error = "render_begin: already begun"; return B200MIX_ERR_INVALID; }
    if(d->shard.transport)
    { d->error = "render_begin: a sharded device set exchanges its wet buffers itself — use b200mix_render"; return B200MIX_ERR_INVALID; }
    if(int rc = render_phase_a(d, frames, true, true)) return rc;
    d->mid_render = true; d->mid_frames = frames;
    if(wet_dev) *wet_dev = d->d_wet;
    if(wet_floats) *wet_floats = d->d_wet ? size_t(d->desc.max_slots)*d->desc.wet_channels*kLine : 0;
    return B200MIX_OK;
}

int b200mix_render_end(b200mix_device *d, float *const *real_out, b200mix_voice_result *results,
    const float **real_out_dev)
{
    if(!d) return B200MIX_ERR_INVALID;
    if(!d->mid_render) { d->error = "render_end: no render_begin pending"; return B200MIX_ERR_INVALID; }
    d->mid_render = false;
    if(int rc = render_phase_b(d, d->mid_frames)) return rc;
    if(int rc = render_output_stage(d, d->mid_frames)) return rc;
    if(real_out_dev) *real_out_dev = d->d_real;
    if(!real_out && !results) return B200MIX_OK;
    return render_collect(d, d->mid_frames, real_out, results);
}

int b200mix_render_device(b200mix_device *d, uint32_t frames, const float **real_out_dev)
{
    if(!d) return B200MIX_ERR_INVALID;
    if(int rc = render_launch(d, frames, false)) return rc;
    if(real_out_dev) *real_out_dev = d->d_real;
    return B200MIX_OK;
}

// ---- voice-sharded device sets ---------------------------------------------------------
static void shard_release(b200mix_device *d)
{
    b200mix_device::Shard &S = d->shard;
    if(S.transport == 1)
        for(uint32_t r = 0;r < S.world;++r)
            if(r != S.rank && S.peer[r]) cudaIpcCloseMemHandle(S.peer[r]);
    if(S.comm && S.comm_destroy) S.comm_destroy(S.comm);
    if(S.nccl_lib) dlclose(S.nccl_lib);
    cudaFree(S.own); cudaFree(S.d_counters);
    if(S.h_status) cudaFreeHost(S.h_status);
    for(cudaEvent_t e : S.ev) if(e) cudaEventDestroy(e);
    S = b200mix_device::Shard{};
}

static int shard_common(b200mix_device *d, uint32_t rank, uint32_t world)
{
    if(world < 1u || world > kShardMaxWorld || rank >= world)
    { d->error = "shard: rank/world out of range (world <= 16)"; return B200MIX_ERR_INVALID; }
    if(d->mid_render) { d->error = "shard: a render_begin is pending"; return B200MIX_ERR_INVALID; }
    CUDA_TRY(d, cudaSetDevice(d->cuda_dev));
    CUDA_TRY(d, cudaStreamSynchronize(d->stream));
    shard_release(d);
    d->shard.rank = rank; d->shard.world = world;
    for(cudaEvent_t &e : d->shard.ev) CUDA_TRY(d, cudaEventCreate(&e));
    return B200MIX_OK;
}

int b200mix_shard_init(b200mix_device *d, uint32_t rank, uint32_t world, void *handle_out)
{
    if(!d) return B200MIX_ERR_INVALID;
    if(!handle_out) { d->error = "shard_init: null handle"; return B200MIX_ERR_INVALID; }
    static_assert(sizeof(cudaIpcMemHandle_t) == B200MIX_SHARD_HANDLE_BYTES, "IPC handle size");
    if(int rc = shard_common(d, rank, world)) return rc;
    b200mix_device::Shard &S = d->shard;
    const b200mix_device_desc &dd = d->desc;
    S.real_floats = size_t(std::max(dd.real_channels, 1u))*kLine;
    S.owned_max = dd.max_slots ? (dd.max_slots + world - 1u)/world : 0u;
    S.wet_src_floats = size_t(S.owned_max)*dd.wet_channels*kLine;
    S.off_real = sizeof(ShardCtl);
    S.off_wet = S.off_real + size_t(2)*world*S.real_floats*sizeof(float);
    S.bytes = S.off_wet + size_t(2)*world*S.wet_src_floats*sizeof(float);
    CUDA_TRY(d, cudaMalloc(reinterpret_cast<void**>(&S.own), S.bytes));
    CUDA_TRY(d, cudaMemset(S.own, 0, S.bytes));
    if(int rc = dev_alloc(d, S.d_counters, 4 + kShardMaxWorld)) return rc;
    CUDA_TRY(d, cudaMallocHost(reinterpret_cast<void**>(&S.h_status), sizeof(uint32_t)));
    *S.h_status = 0u;
    CUDA_TRY(d, cudaStreamSynchronize(d->stream));
    cudaIpcMemHandle_t h;
    CUDA_TRY(d, cudaIpcGetMemHandle(&h, S.own));
    std::memcpy(handle_out, &h, sizeof(h));
    return B200MIX_OK;
}

int b200mix_shard_connect(b200mix_device *d, const void *handles)
{
    if(!d) return B200MIX_ERR_INVALID;
    b200mix_device::Shard &S = d->shard;
    if(!handles || !S.own || S.transport)
    { d->error = "shard_connect: call b200mix_shard_init first (once)"; return B200MIX_ERR_INVALID; }
    CUDA_TRY(d, cudaSetDevice(d->cuda_dev));
    for(uint32_t r = 0;r < S.world;++r)
    {
        if(r == S.rank) { S.peer[r] = S.own; continue; }
        cudaIpcMemHandle_t h;
        std::memcpy(&h, static_cast<const char*>(handles) + size_t(r)*sizeof(h), sizeof(h));
        void *p = nullptr;
        CUDA_TRY(d, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        S.peer[r] = static_cast<char*>(p);
    }
    S.transport = 1; S.epoch = 0;
    return B200MIX_OK;
}

namespace { struct NcclId { char internal[128]; }; }

static void *nccl_open(std::string &err)
{
    void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if(!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if(!lib) err = std::string("NCCL not found: ") + dlerror();
    return lib;
}

int b200mix_shard_nccl_id(void *id_out)
{
    if(!id_out) return B200MIX_ERR_INVALID;
    std::string err;
    void *lib = nccl_open(err);
    if(!lib) { g_create_error = err; return B200MIX_ERR_UNSUPPORTED; }
    auto get = reinterpret_cast<int(*)(NcclId*)>(dlsym(lib, "ncclGetUniqueId"));
    NcclId id{};
    const int rc = get ? get(&id) : 1;
    if(rc == 0) std::memcpy(id_out, &id, sizeof(id));
    dlclose(lib);
    return rc == 0 ? B200MIX_OK : B200MIX_ERR_CUDA;
}

int b200mix_shard_nccl(b200mix_device *d, uint32_t rank, uint32_t world, const void *nccl_id)
{
    if(!d) return B200MIX_ERR_INVALID;
    if(!nccl_id) { d->error = "shard_nccl: null id"; return B200MIX_ERR_INVALID; }
    if(int rc = shard_common(d, rank, world)) return rc;
    b200mix_device::Shard &S = d->shard;
    S.nccl_lib = nccl_open(d->error);
    if(!S.nccl_lib) return B200MIX_ERR_UNSUPPORTED;
    auto init = reinterpret_cast<int(*)(void**, int, NcclId, int)>(dlsym(S.nccl_lib, "ncclCommInitRank"));
    S.reduce = reinterpret_cast<decltype(S.reduce)>(dlsym(S.nccl_lib, "ncclReduce"));
    S.allreduce = reinterpret_cast<decltype(S.allreduce)>(dlsym(S.nccl_lib, "ncclAllReduce"));
    S.comm_destroy = reinterpret_cast<decltype(S.comm_destroy)>(dlsym(S.nccl_lib, "ncclCommDestroy"));
    if(!init || !S.reduce || !S.allreduce || !S.comm_destroy)
    { d->error = "shard_nccl: NCCL symbols missing"; return B200MIX_ERR_UNSUPPORTED; }
    NcclId id; std::memcpy(&id, nccl_id, sizeof(id));
    if(init(&S.comm, int(world), id, int(rank)) != 0)
    { d->error = "ncclCommInitRank failed"; S.comm = nullptr; return B200MIX_ERR_CUDA; }
    S.transport = 2; S.epoch = 0;
    return B200MIX_OK;
}

int b200mix_shard_last_us(b200mix_device *d, float *wet_us, float *real_us)
{
    if(!d || !d->shard.transport) return B200MIX_ERR_INVALID;
    b200mix_device::Shard &S = d->shard;
    float ms = 0.0f;
    if(wet_us)
    {
        *wet_us = -1.0f;
        if(S.ev_wet && cudaEventSynchronize(S.ev[1]) == cudaSuccess
            && cudaEventElapsedTime(&ms, S.ev[0], S.ev[1]) == cudaSuccess) *wet_us = ms*1000.0f;
    }
    if(real_us)
    {
        *real_us = -1.0f;
        if(S.ev_real && cudaEventSynchronize(S.ev[3]) == cudaSuccess
            && cudaEventElapsedTime(&ms, S.ev[2], S.ev[3]) == cudaSuccess) *real_us = ms*1000.0f;
    }
    return B200MIX_OK;
}

int b200mix_get_dry(b200mix_device *d, float *dry)
{
    if(!d || !dry) return B200MIX_ERR_INVALID;
    CUDA_TRY(d, cudaSetDevice(d->cuda_dev));
    CUDA_TRY(d, cudaStreamSynchronize(d->stream));
    CUDA_TRY(d, cudaMemcpy(dry, d->d_dry, size_t(d->desc.dry_channels)*kLine*sizeof(float),
        cudaMemcpyDeviceToHost));
    return B200MIX_OK;
}

int64_t b200mix_get_resampler_table(b200mix_device *d, uint32_t which, float *out, size_t max_floats)
{
    if(!d) return -1;
    if(cudaSetDevice(d->cuda_dev) != cudaSuccess) return -1;
    const float *src = nullptr; size_t n = 0;
    if(which == B200MIX_RESAMPLER_SPLINE) { src = d->d_cubic[0]; n = 256; }
    else if(which == B200MIX_RESAMPLER_GAUSSIAN) { src = d->d_cubic[1]; n = 256; }
    else if(bsinc_for(d, which))
    {
        const int i = int(which - B200MIX_RESAMPLER_FAST_BSINC12) >> 1;
        src = d->d_bsinc[i]; n = d->bsinc[i].tab.size();
    }
    else return -1;
    if(out)
    {
        cudaStreamSynchronize(d->stream);
        if(cudaMemcpy(out, src, std::min(n, max_floats)*sizeof(float), cudaMemcpyDeviceToHost)
            != cudaSuccess) return -1;
    }
    return int64_t(n);
}

// Taps per output sample the resampler of a voice with this step runs (BsincPrepare's m for
// the bsinc family, alc/alu.cpp:140-165; 4 for the cubic family, 2 linear, 1 point; 0 for the
// pitch-1.0 copy) and whether it is the full BSinc form (scale interpolation, mixer_c.cpp:84-105).
int b200mix_resampler_taps(b200mix_device *d, uint32_t resampler, uint32_t step, uint32_t *full)
{
    if(!d || resampler > B200MIX_RESAMPLER_BSINC48) return B200MIX_ERR_INVALID;
    if(full) *full = 0u;
    if(const BsincTable *t = bsinc_for(d, resampler))
    {
        const BsincState st = PrepareBsinc(*t, step);
        if(full) *full = (step > 65536u && (resampler & 1u)) ? 1u : 0u;
        return int(st.m);
    }
    return resampler >= 2u ? 4 : (resampler == 1u ? 2 : 1);
}

int b200mix_profile(b200mix_device *d, int enable)
{
    if(!d) return B200MIX_ERR_INVALID;
    CUDA_TRY(d, cudaSetDevice(d->cuda_dev));
    if(enable && !d->ev_mix0)
    {
        CUDA_TRY(d, cudaEventCreate(&d->ev_mix0));
        CUDA_TRY(d, cudaEventCreate(&d->ev_mix1));
    }
    if(enable >= 2 && !d->ev_stage[0])
        for(cudaEvent_t &e : d->ev_stage) CUDA_TRY(d, cudaEventCreate(&e));
    d->profile = enable != 0;
    d->profile_level = enable;
    d->ev_valid = false; d->stage_valid = false;
    return B200MIX_OK;
}

int b200mix_last_stage_ms(b200mix_device *d, float *ms, uint32_t count)
{
    if(!d || !ms || !d->stage_valid) return B200MIX_ERR_INVALID;
    if(cudaEventSynchronize(d->ev_stage[b200mix_device::kStages]) != cudaSuccess) return B200MIX_ERR_CUDA;
    for(uint32_t i = 0;i < count && i < uint32_t(b200mix_device::kStages);++i)
        if(cudaEventElapsedTime(&ms[i], d->ev_stage[i], d->ev_stage[i+1]) != cudaSuccess) return B200MIX_ERR_CUDA;
    return int(b200mix_device::kStages);
}

float b200mix_last_mix_kernel_ms(b200mix_device *d)
{
    if(!d || !d->ev_valid) return -1.0f;
    if(cudaEventSynchronize(d->ev_mix1) != cudaSuccess) return -1.0f;
    float ms = -1.0f;
    if(cudaEventElapsedTime(&ms, d->ev_mix0, d->ev_mix1) != cudaSuccess) return -1.0f;
    return ms;
}

uint64_t b200mix_launch_count(const b200mix_device *d) { return d ? d->launches : 0; }
void *b200mix_stream(b200mix_device *d) { return d ? static_cast<void*>(d->stream) : nullptr; }

} // extern "C"
