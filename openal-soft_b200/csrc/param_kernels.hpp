// param_kernels.hpp — interface between the library's host side (b200mix.cu) and the GPU
// parameter stage (param_kernels.cu, compiled separately with -fmad=false).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "param_math.hpp"
#include "voice_structs.hpp"

namespace b200mix {

struct CalcVoicesParams {
    const b200mix_source_voice *voices;       // [n] what the host still decides per voice
    const b200mix_source_props *props;        // [n] VoiceProps as the application set them
    uint32_t n;
    b200mix_listener_params listener;         // ContextParams
    uint32_t device_rate, num_sends, render_mode, cd, cw, ir;
    // DeviceBase::Dry.AmbiMap and the sends' slot Wet.AmbiMap
    uint32_t dry_channels; float dry_scale[B200MIX_MAX_DRY_CHANNELS]; uint32_t dry_index[B200MIX_MAX_DRY_CHANNELS];
    uint32_t wet_channels[B200MIX_MAX_SENDS];
    float wet_scale[B200MIX_MAX_SENDS][B200MIX_MAX_WET_CHANNELS];
    uint32_t wet_index[B200MIX_MAX_SENDS][B200MIX_MAX_WET_CHANNELS];
    pm::BsincMeta bsinc[3];                   // bsinc12 / 24 / 48 scale tables (BsincPrepare)
    // results, laid out as b200mix_voices_update_dirs / b200mix_voices_filters stage them
    VoiceUpdate *updates;                     // [n]
    float4 *dirs;                             // [n] {elevation, azimuth, distance, spread}
    float *dry;                               // [n][cd]
    float *send;                              // [n][num_sends][cw]
    float *gains_hflf;                        // [n][1 + MAX_SENDS][2] per-path {GainHF, GainLF}
    FilterUpdate *fupd;                       // [n][1 + num_sends]
};

// k_calc_voices (+ k_design_filters when `filters`) on `stream`.
cudaError_t launch_calc_voices(const CalcVoicesParams &Q, bool filters, cudaStream_t stream);

} // namespace b200mix
