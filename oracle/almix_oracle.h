/* almix_oracle.h — TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, scalar, single thread) of the reference's per-update
 * mixing hot path, behind the SAME call sequence as the product's C ABI
 * (include/b200mix.h) with the prefix oracle_ instead of b200mix_, so a test
 * drives both with identical inputs and compares outputs.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load this.  The product (openal-soft_b200/) never does.
 *
 * Parity pinning: the reference ships no golden vectors, so this
 * restatement is pinned against the compiled reference itself
 * (oracle/_ref/libopenal_ref.so, built by oracle/refbuild/Makefile): tables
 * bit-for-bit, resamplers bit-for-bit vs the reference's C kernels, whole
 * updates vs alcRenderSamplesSOFT — see tests/test_oracle_*.py and the
 * committed fixtures under tests/golden/.
 */
#ifndef ALMIX_ORACLE_H
#define ALMIX_ORACLE_H

#include "../include/b200mix.h"

#ifdef __cplusplus
extern "C" {
#endif

#define ORACLE_BSINC_SCALES 16u  /* BSincScaleCount core/bsinc_defs.h:8 */
#define ORACLE_BSINC_PHASES 32u  /* BSincPhaseCount core/bsinc_defs.h:10 */
#define ORACLE_CUBIC_PHASES 32u  /* CubicPhaseCount core/cubic_defs.h:8 */
#define ORACLE_MAX_TAPS     48u  /* MaxResamplerPadding */

typedef struct oracle_bsinc_table {
    float scaleBase, scaleRange;
    uint32_t m[ORACLE_BSINC_SCALES];
    uint32_t filterOffset[ORACLE_BSINC_SCALES];
    const float *tab;
    size_t total;
} oracle_bsinc_table;

typedef struct oracle_bsinc_state {
    float sf; uint32_t m, l; const float *filter;
} oracle_bsinc_state;

int  oracle_build_bsinc(oracle_bsinc_table *t, double rejection, double order, double maxScale);
void oracle_build_gaussian(float tab[ORACLE_CUBIC_PHASES][8]);
void oracle_build_spline(float tab[ORACLE_CUBIC_PHASES][8]);
void oracle_build_cubic_filter(float filter[513]);
void oracle_bsinc_prepare(const oracle_bsinc_table *t, uint32_t increment, oracle_bsinc_state *st);

/* table access for tests: which = enum b200mix_resampler; returns float count */
int64_t oracle_get_resampler_table(uint32_t which, float *out, size_t max_floats);
int oracle_get_bsinc_state(uint32_t which, uint32_t increment, float *sf, uint32_t *m, uint32_t *l,
    uint32_t *offset);

/* Resample_*_C (core/mixer/mixer_c.cpp:190-221): src is mResampleData
 * (position 0 at index 24). */
int oracle_resample(uint32_t resampler, uint32_t increment, uint32_t frac, const float *src,
    float *dst, uint32_t dst_len);

/* ---- same surface as include/b200mix.h ---------------------------------- */
typedef struct oracle_device oracle_device;
int  oracle_create(const b200mix_device_desc *desc, oracle_device **out);
void oracle_destroy(oracle_device *dev);
int  oracle_set_hrtf_decoder(oracle_device *dev, uint32_t channels, uint32_t ir_size,
    const float *coeffs, const float *hf_scale, const float *splitter_coeff);
int  oracle_set_ambi_decoder(oracle_device *dev, uint32_t in_channels, const float *gains_hf,
    const float *gains_lf, float xover_coeff);
int  oracle_buffer_data(oracle_device *dev, uint32_t buffer, uint32_t sample_type,
    uint32_t channels, uint32_t frames, const void *data, size_t bytes);
int  oracle_buffer_data_adpcm(oracle_device *dev, uint32_t buffer, uint32_t sample_type,
    uint32_t channels, uint32_t samples_per_block, uint32_t blocks, const void *data, size_t bytes);
int  oracle_buffer_free(oracle_device *dev, uint32_t buffer);
int  oracle_voices_update(oracle_device *dev, uint32_t n, const b200mix_voice_params *params,
    const float *hrtf_coeffs, const float *dry_gains, const float *send_gains);
int  oracle_voice_queue(oracle_device *dev, uint32_t voice, uint32_t count, const uint32_t *buffers,
    uint32_t loop_index);
int  oracle_voices_filters(oracle_device *dev, uint32_t n, const b200mix_voice_filter *filters);
int  oracle_biquad_coeffs(uint32_t type, float f0norm, float gain, float slope, float coeffs[5]);
int  oracle_render(oracle_device *dev, uint32_t frames, float *const *real_out,
    b200mix_voice_result *results);
int  oracle_render_begin(oracle_device *dev, uint32_t frames, float **wet_host, size_t *wet_floats);
int  oracle_render_end(oracle_device *dev, float *const *real_out, b200mix_voice_result *results,
    const float **real_out_host);
int  oracle_render_interleaved(oracle_device *dev, uint32_t frames, void *out, uint32_t out_type,
    uint32_t frame_step, float dither_depth, uint32_t *dither_seed, b200mix_voice_result *results);
int  oracle_slot_convolution(oracle_device *dev, uint32_t slot, uint32_t ir_channels,
    uint32_t ir_frames, const float *ir);
int  oracle_slot_output_gains(oracle_device *dev, uint32_t slot, uint32_t lines, const float *gains);
int  oracle_slot_reverb(oracle_device *dev, uint32_t slot, const b200mix_reverb_params *params);
int  oracle_slot_reverb_update(oracle_device *dev, uint32_t slot, const b200mix_reverb_params *params,
    uint32_t full_update);
int  oracle_slot_disable(oracle_device *dev, uint32_t slot);
int  oracle_slot_target(oracle_device *dev, uint32_t slot, uint32_t target);
int  oracle_slot_efx(oracle_device *dev, uint32_t slot, const b200mix_efx_props *props,
    const b200mix_efx_target *target);
int  oracle_set_distance_comp(oracle_device *dev, uint32_t channels, const uint32_t *delays, const float *gains);
int  oracle_set_uhj_encoder(oracle_device *dev, uint32_t filter_length, uint32_t *delay);
int  oracle_set_front_stabilizer(oracle_device *dev, uint32_t center_channel, float splitter_coeff);
int  oracle_set_bs2b(oracle_device *dev, uint32_t level);
int  oracle_bs2b_coeffs(uint32_t level, uint32_t srate, float out[5]);
void oracle_bs2b_cross_feed(const float coef[5], float hist[2][2], float *left, float *right, size_t n);
int  oracle_set_limiter(oracle_device *dev, const b200mix_limiter_desc *desc, uint32_t *look_ahead);
int  oracle_get_dry(oracle_device *dev, float *dry);
/* test-only: wet mix of one slot [wet_channels][1024] */
int  oracle_get_wet(oracle_device *dev, uint32_t slot, float *wet);
/* test-only: device-wide HRTF accumulator [1024+128][2] */
int  oracle_get_hrtf_accum(oracle_device *dev, float *accum);

#ifdef __cplusplus
}
#endif
#endif
