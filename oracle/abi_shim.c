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
EXPORT int b200mix_voices_update(b200mix_device *dev, uint32_t n, const b200mix_voice_params *params,
    const float *hrtf_coeffs, const float *dry_gains, const float *send_gains)
{ return oracle_voices_update((oracle_device*)dev, n, params, hrtf_coeffs, dry_gains, send_gains); }
EXPORT int b200mix_render(b200mix_device *dev, uint32_t frames, float *const *real_out, b200mix_voice_result *results)
{ return oracle_render((oracle_device*)dev, frames, real_out, results); }
