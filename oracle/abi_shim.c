/* TEST INFRASTRUCTURE ONLY.  The oracle behind the b200mix_* entry-point names that
 * integration/b200mix_seam.cpp binds, so the seam's host logic (voice snapshots, change
 * detection, cursor write-back) can be exercised on a machine without a GPU:
 * tests/test_seam_cpu.py points ALSOFT_B200MIX_LIB at this library and compares the patched
 * reference's output with the stock reference's.  Never shipped, never loaded by the product. */
#include "almix_oracle.h"

#define EXPORT __attribute__((visibility("default")))

EXPORT int b200mix_create(const b200mix_device_desc *desc, b200mix_device **out)
{ return oracle_create(desc, (oracle_device**)out); }
EXPORT void b200mix_destroy(b200mix_device *dev) { oracle_destroy((oracle_device*)dev); }
EXPORT const char *b200mix_last_error(const b200mix_device *dev) { (void)dev; return "(oracle shim)"; }
EXPORT int b200mix_set_hrtf_decoder(b200mix_device *dev, uint32_t channels, uint32_t ir_size,
    const float *coeffs, const float *hf_scale, const float *splitter_coeff)
{ return oracle_set_hrtf_decoder((oracle_device*)dev, channels, ir_size, coeffs, hf_scale, splitter_coeff); }
EXPORT int b200mix_set_ambi_decoder(b200mix_device *dev, uint32_t in_channels, const float *gains_hf,
    const float *gains_lf, float splitter_coeff)
{ return oracle_set_ambi_decoder((oracle_device*)dev, in_channels, gains_hf, gains_lf, splitter_coeff); }
EXPORT int b200mix_buffer_data(b200mix_device *dev, uint32_t buffer, uint32_t sample_type, uint32_t channels,
    uint32_t frames, const void *data, size_t bytes)
{ return oracle_buffer_data((oracle_device*)dev, buffer, sample_type, channels, frames, data, bytes); }
EXPORT int b200mix_buffer_data_adpcm(b200mix_device *dev, uint32_t buffer, uint32_t sample_type, uint32_t channels,
    uint32_t samples_per_block, uint32_t blocks, const void *data, size_t bytes)
{ return oracle_buffer_data_adpcm((oracle_device*)dev, buffer, sample_type, channels, samples_per_block, blocks, data, bytes); }
EXPORT int b200mix_voices_update(b200mix_device *dev, uint32_t n, const b200mix_voice_params *params,
    const float *hrtf_coeffs, const float *dry_gains, const float *send_gains)
{ return oracle_voices_update((oracle_device*)dev, n, params, hrtf_coeffs, dry_gains, send_gains); }
EXPORT int b200mix_render(b200mix_device *dev, uint32_t frames, float *const *real_out, b200mix_voice_result *results)
{ return oracle_render((oracle_device*)dev, frames, real_out, results); }

/* ---- effect slots (seam v2) ---- */
EXPORT int b200mix_slot_efx(b200mix_device *dev, uint32_t slot, const b200mix_efx_props *props, const b200mix_efx_target *target)
{ return oracle_slot_efx((oracle_device*)dev, slot, props, target); }
EXPORT int b200mix_slot_reverb(b200mix_device *dev, uint32_t slot, const b200mix_reverb_params *params)
{ return oracle_slot_reverb((oracle_device*)dev, slot, params); }
EXPORT int b200mix_slot_reverb_update(b200mix_device *dev, uint32_t slot, const b200mix_reverb_params *params, uint32_t full_update)
{ return oracle_slot_reverb_update((oracle_device*)dev, slot, params, full_update); }
EXPORT int b200mix_slot_output_gains(b200mix_device *dev, uint32_t slot, uint32_t lines, const float *gains)
{ return oracle_slot_output_gains((oracle_device*)dev, slot, lines, gains); }
EXPORT int b200mix_slot_target(b200mix_device *dev, uint32_t slot, uint32_t target)
{ return oracle_slot_target((oracle_device*)dev, slot, target); }
EXPORT int b200mix_slot_disable(b200mix_device *dev, uint32_t slot)
{ return oracle_slot_disable((oracle_device*)dev, slot); }

/* The reverb's parameter stage is host code of the product (csrc/reverb_params.cpp, pinned
 * bit-exact against the reference by tests/test_reverb_params.py): forwarded, not restated. */
#include <dlfcn.h>
#include <stdlib.h>
static void *host_lib(void)
{
    static void *lib;
    if(!lib)
    {
        const char *path = getenv("B200MIX_HOST_LIB");
        lib = dlopen(path ? path : "libb200mix.so", RTLD_NOW | RTLD_LOCAL);
    }
    return lib;
}
EXPORT int b200mix_reverb_params_from_efx(const b200mix_efx_reverb *props, const b200mix_reverb_target *target,
    struct b200mix_reverb_params *params, float *gains)
{
    typedef int (*fn_t)(const b200mix_efx_reverb*, const b200mix_reverb_target*, struct b200mix_reverb_params*, float*);
    fn_t fn = host_lib() ? (fn_t)dlsym(host_lib(), "b200mix_reverb_params_from_efx") : NULL;
    return fn ? fn(props, target, params, gains) : B200MIX_ERR_INVALID;
}
EXPORT int b200mix_reverb_full_update_needed(const b200mix_efx_reverb *prev, const b200mix_efx_reverb *next)
{
    typedef int (*fn_t)(const b200mix_efx_reverb*, const b200mix_efx_reverb*);
    fn_t fn = host_lib() ? (fn_t)dlsym(host_lib(), "b200mix_reverb_full_update_needed") : NULL;
    return fn ? fn(prev, next) : 1;
}
EXPORT int b200mix_voices_filters(b200mix_device *dev, uint32_t n, const b200mix_voice_filter *filters)
{ return oracle_voices_filters((oracle_device*)dev, n, filters); }
EXPORT int b200mix_voice_queue(b200mix_device *dev, uint32_t voice, uint32_t count, const uint32_t *buffers, uint32_t loop_index)
{ return oracle_voice_queue((oracle_device*)dev, voice, count, buffers, loop_index); }

/* convolution slots: the device call goes to the oracle, the host helpers to the product library */
EXPORT int b200mix_slot_convolution(b200mix_device *dev, uint32_t slot, uint32_t ir_channels, uint32_t ir_frames, const float *ir)
{ return oracle_slot_convolution((oracle_device*)dev, slot, ir_channels, ir_frames, ir); }
EXPORT int b200mix_convolution_gains(uint32_t layout, uint32_t pairwise, float slot_gain, uint32_t channels, const float *scale,
    const uint32_t *index, float *gains, uint32_t gains_stride)
{
    typedef int (*fn_t)(uint32_t, uint32_t, float, uint32_t, const float*, const uint32_t*, float*, uint32_t);
    fn_t fn = host_lib() ? (fn_t)dlsym(host_lib(), "b200mix_convolution_gains") : NULL;
    return fn ? fn(layout, pairwise, slot_gain, channels, scale, index, gains, gains_stride) : B200MIX_ERR_INVALID;
}
EXPORT int64_t b200mix_resampled_ir_frames(uint32_t src_rate, uint32_t dst_rate, uint32_t frames)
{
    typedef int64_t (*fn_t)(uint32_t, uint32_t, uint32_t);
    fn_t fn = host_lib() ? (fn_t)dlsym(host_lib(), "b200mix_resampled_ir_frames") : NULL;
    return fn ? fn(src_rate, dst_rate, frames) : -1;
}
EXPORT int b200mix_resample_ir(uint32_t src_rate, uint32_t dst_rate, const float *in, uint32_t in_frames, float *out, uint32_t out_frames)
{
    typedef int (*fn_t)(uint32_t, uint32_t, const float*, uint32_t, float*, uint32_t);
    fn_t fn = host_lib() ? (fn_t)dlsym(host_lib(), "b200mix_resample_ir") : NULL;
    return fn ? fn(src_rate, dst_rate, in, in_frames, out, out_frames) : B200MIX_ERR_INVALID;
}

/* the other post-process variants */
EXPORT int b200mix_set_uhj_encoder(b200mix_device *dev, uint32_t filter_length, uint32_t *delay)
{ return oracle_set_uhj_encoder((oracle_device*)dev, filter_length, delay); }
EXPORT int b200mix_set_bs2b(b200mix_device *dev, uint32_t level)
{ return oracle_set_bs2b((oracle_device*)dev, level); }
EXPORT int b200mix_set_front_stabilizer(b200mix_device *dev, uint32_t center_channel, float splitter_coeff)
{ return oracle_set_front_stabilizer((oracle_device*)dev, center_channel, splitter_coeff); }
