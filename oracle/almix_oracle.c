/* oracle/almix_oracle.c — TEST INFRASTRUCTURE ONLY (see almix_oracle.h).
 *
 * Scalar C restatement of one OpenAL Soft mix update.  Each function cites the
 * reference code it follows (paths relative to the reference root).  The order
 * of floating-point operations follows the reference's C kernels
 * (core/mixer/mixer_c.cpp) so that results are bit-identical to the reference
 * built with disable-cpu-exts=all, and within fp32 rounding of its SSE kernels.
 * Build with -ffp-contract=off (no FMA contraction; the reference's x86-64
 * build has none either).
 */
#include "almix_oracle.h"
#include "reverb_oracle.h"
#include "limiter_oracle.h"
#include "efx_oracle.h"

#define ORACLE_PI 3.14159265358979323846   /* std::numbers::pi */

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define LINE   B200MIX_LINE_SIZE
#define HRIR   B200MIX_HRIR_LENGTH
#define HIST   B200MIX_HRTF_HISTORY
#define EDGE   24u                      /* MaxResamplerEdge     core/resampler_limits.h:10 */
#define PAD    B200MIX_RESAMPLER_PADDING/* MaxResamplerPadding  core/resampler_limits.h:8 */
#define FRAC_BITS 16
#define FRAC_ONE  (1u<<FRAC_BITS)
#define FRAC_MASK (FRAC_ONE-1u)
#define DECODER_MAX_PADDING 256u        /* DecoderBase::sMaxPadding core/decoderbase.hpp */
#define RESBUF (LINE + DECODER_MAX_PADDING + PAD) /* DeviceBase::mResampleData core/device.h:282 */
#define SILENCE_THRESHOLD 0.00001f      /* GainSilenceThreshold core/mixer/defs.h:28 */

/* ---- tables (static-init in the reference) ------------------------------- */
static oracle_bsinc_table g_bsinc12, g_bsinc24, g_bsinc48;
static float g_spline[ORACLE_CUBIC_PHASES][8], g_gaussian[ORACLE_CUBIC_PHASES][8];
static int g_tables_ready;

static int ensure_tables(void)
{
    if(g_tables_ready) return 0;
    /* core/bsinc_tables.cpp:150-155 */
    if(oracle_build_bsinc(&g_bsinc12, 60, 11, 2)) return -1;
    if(oracle_build_bsinc(&g_bsinc24, 60, 23, 2)) return -1;
    if(oracle_build_bsinc(&g_bsinc48, 80, 47, 1)) return -1;
    oracle_build_spline(g_spline);
    oracle_build_gaussian(g_gaussian);
    g_tables_ready = 1;
    return 0;
}

static const oracle_bsinc_table *bsinc_for(uint32_t which)
{
    switch(which)
    {
    case B200MIX_RESAMPLER_FAST_BSINC12: case B200MIX_RESAMPLER_BSINC12: return &g_bsinc12;
    case B200MIX_RESAMPLER_FAST_BSINC24: case B200MIX_RESAMPLER_BSINC24: return &g_bsinc24;
    case B200MIX_RESAMPLER_FAST_BSINC48: case B200MIX_RESAMPLER_BSINC48: return &g_bsinc48;
    }
    return NULL;
}

int64_t oracle_get_resampler_table(uint32_t which, float *out, size_t max_floats)
{
    if(ensure_tables()) return -1;
    const float *src; size_t n;
    if(which == B200MIX_RESAMPLER_SPLINE) { src = &g_spline[0][0]; n = ORACLE_CUBIC_PHASES*8; }
    else if(which == B200MIX_RESAMPLER_GAUSSIAN) { src = &g_gaussian[0][0]; n = ORACLE_CUBIC_PHASES*8; }
    else
    {
        const oracle_bsinc_table *t = bsinc_for(which);
        if(!t) return -1;
        src = t->tab; n = t->total;
    }
    if(out) memcpy(out, src, sizeof(float)*(n < max_floats ? n : max_floats));
    return (int64_t)n;
}

int oracle_get_bsinc_state(uint32_t which, uint32_t increment, float *sf, uint32_t *m, uint32_t *l,
    uint32_t *offset)
{
    if(ensure_tables()) return -1;
    const oracle_bsinc_table *t = bsinc_for(which);
    if(!t) return -1;
    oracle_bsinc_state st;
    oracle_bsinc_prepare(t, increment, &st);
    *sf = st.sf; *m = st.m; *l = st.l; *offset = (uint32_t)(st.filter - t->tab);
    return 0;
}

/* ---- resamplers: core/mixer/mixer_c.cpp:39-137,190-221 ------------------- */
static float lerpf(float a, float b, float mu) { return a + (b-a)*mu; } /* alnumeric.h:115 */

int oracle_resample(uint32_t resampler, uint32_t increment, uint32_t frac, const float *src,
    float *dst, uint32_t dst_len)
{
    if(ensure_tables()) return -1;
    size_t pos = 0;
    switch(resampler)
    {
    case B200MIX_RESAMPLER_POINT: {
        const float *vals = src + EDGE;
        for(uint32_t i = 0;i < dst_len;++i)
        {
            dst[i] = vals[pos];
            frac += increment; pos += frac>>FRAC_BITS; frac &= FRAC_MASK;
        }
        return 0; }
    case B200MIX_RESAMPLER_LINEAR: {
        const float *vals = src + EDGE;
        for(uint32_t i = 0;i < dst_len;++i)
        {
            dst[i] = lerpf(vals[pos+0], vals[pos+1], (float)frac*(1.0f/FRAC_ONE));
            frac += increment; pos += frac>>FRAC_BITS; frac &= FRAC_MASK;
        }
        return 0; }
    case B200MIX_RESAMPLER_SPLINE:
    case B200MIX_RESAMPLER_GAUSSIAN: {
        /* do_cubic, mixer_c.cpp:48-61; CubicPhaseDiffBits = 16-5 = 11 */
        const float (*tab)[8] = (resampler == B200MIX_RESAMPLER_SPLINE) ? g_spline : g_gaussian;
        const float *vals = src + EDGE-1;
        for(uint32_t i = 0;i < dst_len;++i)
        {
            const unsigned pi = frac>>11;
            const float pf = (float)(frac&2047u)*(1.0f/2048.0f);
            const float *fil = tab[pi], *phd = tab[pi]+4;
            dst[i] = (fil[0] + pf*phd[0])*vals[pos+0] + (fil[1] + pf*phd[1])*vals[pos+1]
                + (fil[2] + pf*phd[2])*vals[pos+2] + (fil[3] + pf*phd[3])*vals[pos+3];
            frac += increment; pos += frac>>FRAC_BITS; frac &= FRAC_MASK;
        }
        return 0; }
    default: break;
    }
    const oracle_bsinc_table *t = bsinc_for(resampler);
    if(!t) return -1;
    oracle_bsinc_state st;
    oracle_bsinc_prepare(t, increment, &st);
    const size_t m = st.m;
    const float *vals = src + EDGE - st.l;
    /* SelectResampler, alc/alu.cpp:203-235: full BSinc only when down-sampling with
     * a non-"fast" kind; otherwise FastBSinc (phase interpolation only). */
    const int full = (increment > FRAC_ONE) && (resampler == B200MIX_RESAMPLER_BSINC12
        || resampler == B200MIX_RESAMPLER_BSINC24 || resampler == B200MIX_RESAMPLER_BSINC48);
    for(uint32_t i = 0;i < dst_len;++i)
    {
        const unsigned pi = frac>>11;
        const float pf = (float)(frac&2047u)*(1.0f/2048.0f);
        const float *fil = st.filter + 2*pi*m;
        const float *phd = fil + m;
        float r = 0.0f;
        if(full)
        {
            /* do_bsinc, mixer_c.cpp:84-105 */
            const float *scd = fil + ORACLE_BSINC_PHASES*2*m;
            const float *spd = scd + m;
            for(size_t j = 0;j < m;++j)
                r += (fil[j] + st.sf*scd[j] + pf*(phd[j] + st.sf*spd[j])) * vals[pos+j];
        }
        else
        {
            /* do_fastbsinc, mixer_c.cpp:63-82 */
            for(size_t j = 0;j < m;++j)
                r += (fil[j] + pf*phd[j]) * vals[pos+j];
        }
        dst[i] = r;
        frac += increment; pos += frac>>FRAC_BITS; frac &= FRAC_MASK;
    }
    return 0;
}

/* ---- device state --------------------------------------------------------- */
typedef struct { uint32_t type, channels, frames; void *data; } obuffer;

/* BiquadInterpFilter (core/filters/biquad.h:137-198): coefficient sets are
 * {b0,b1,b2,a1,a2}; z = {mZ1,mZ2}; counter = mCounter. */
typedef struct { float cur[5], tgt[5], z[2]; int counter; } obiquad;

static void obiquad_reset(obiquad *f)      /* BiquadInterpFilter::reset, biquad.h:144-150 */
{
    static const float ident[5] = {1.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    memcpy(f->cur, ident, sizeof(ident)); memcpy(f->tgt, ident, sizeof(ident));
    f->z[0] = f->z[1] = 0.0f;
    f->counter = -1;
}
static void obiquad_clear(obiquad *f)      /* BiquadInterpFilter::clear, biquad.h:152-157 */
{
    f->z[0] = f->z[1] = 0.0f;
    memcpy(f->cur, f->tgt, sizeof(f->cur));
    f->counter = 0;
}

typedef struct {
    int state;                 /* 0 stopped, 1 playing, 2 stopping (Voice::State) */
    uint32_t flags;            /* STATIC / LOOPING / HRTF */
    int fading;                /* VoiceFlag::IsFading */
    int have_buffer;           /* mCurrentBuffer != nullptr */
    uint32_t buffer, resampler;
    int32_t pos; uint32_t frac;
    uint32_t loop_start, loop_end, step;
    float prev[PAD];                          /* mPrevSamples[0] */
    float hist[HIST];                         /* Hrtf.History */
    float tgt_coef[HRIR][2], old_coef[HRIR][2];
    uint32_t tgt_delay[2], old_delay[2];
    float tgt_gain, old_gain;
    float dry_cur[B200MIX_MAX_DRY_CHANNELS], dry_tgt[B200MIX_MAX_DRY_CHANNELS];
    uint32_t send_slot[B200MIX_MAX_SENDS];
    float send_cur[B200MIX_MAX_SENDS][B200MIX_MAX_WET_CHANNELS];
    float send_tgt[B200MIX_MAX_SENDS][B200MIX_MAX_WET_CHANNELS];
    /* streaming queue (VoiceBufferItem list, core/voice.h:84-99): buffer ids from the
     * current item on; q_loop = index playback continues at after the last item */
    uint32_t q_count, q_head, q_loop;
    uint32_t q_items[B200MIX_MAX_QUEUE];
    uint32_t buffers_done;
    /* DirectParams/SendParams LowPass+HighPass, path 0 = direct, 1+s = send s */
    struct { obiquad lp, hp; int active; } filt[1 + B200MIX_MAX_SENDS];
} ovoice;

typedef struct { float coeff, lp_z1, lp_z2, ap_z1; } osplitter; /* core/filters/splitter.h */

/* One aux slot with a ConvolutionState (alc/effects/convolution.cpp:254-300).
 * RESTATEMENT NOTE: the reference evaluates the convolution as a 128-tap time-domain head
 * plus 256-point FFT partitions (pffft).  Both compute the linear convolution
 * y[n] = sum_k h[k] x[n-k]; this oracle evaluates that definition directly with a double
 * accumulator (no FFT), so it is not bit-exact with pffft's float butterflies — it is
 * pinned against the compiled reference to <= 1e-6 (tests/test_oracle_vs_ref.py,
 * tests/golden conv scene). */
typedef struct {
    uint32_t type, channels, frames;
    float *ir;                 /* [channels][frames] */
    float *hist;               /* input history ring, frames+1024 long (newest at hist_pos-1) */
    uint32_t hist_pos;
    float *cur, *tgt;          /* output mix gains [channels][MAX_DRY] */
    oreverb *reverb;           /* type == B200MIX_EFFECT_REVERB */
    oefx *efx;                 /* type >= B200MIX_EFFECT_ECHO (efx_oracle.cpp) */
    uint32_t target;           /* EffectSlotBase::Target as a slot index, B200MIX_NO_SLOT = the Dry mix */
} oslot;

struct oracle_device {
    b200mix_device_desc desc;
    int filters_seen;          /* b200mix_voices_filters has been called: filter records exist (see there) */
    obuffer *buffers;
    ovoice *voices;
    float (*dry)[LINE];        /* Dry.Buffer */
    float (*real)[LINE];       /* RealOut.Buffer (== dry when POST_NONE) */
    float (*wet)[LINE];        /* [slot*wet_channels + c] */
    oslot *slots;
    float accum[LINE+HRIR][2]; /* HrtfAccumData core/device.h:288 */
    /* HRTF decoder (DirectHrtfState) */
    uint32_t dec_channels, dec_ir;
    float (*dec_coef)[HRIR][2];
    float *dec_hfscale;
    osplitter *dec_split;
    /* ambi decoder (BFormatDec) */
    uint32_t amb_in; int amb_dual;
    float *amb_hf, *amb_lf;    /* [in][real] */
    osplitter *amb_split;
    /* UHJ IIR encoder state (UhjEncoderIIR, core/uhjfilter.h) */
    float uhj_f1wx[4][2], uhj_f2wx[4][2], uhj_f1y[4][2], uhj_f1d[2][4][2];
    float uhj_delay_wx, uhj_delay_y, uhj_delay_d[2];
    float uhj_s[LINE+1], uhj_d[LINE+1], uhj_wx[LINE], uhj_t[LINE+1];
    /* UHJ FIR encoder (UhjEncoder<N>, core/uhjfilter.h:27-70): N = 0 selects the IIR one */
    uint32_t uhj_fir;                   /* 0, 256 or 512 */
    double uhj_fir_h[512];              /* the phase-shift response (odd taps only are non-zero) */
    float uhj_wxhist[512+128];          /* the last N+127 values of -0.171 W + 0.208 X */
    float uhj_in_delay[4][512/2+128];   /* W, X, Y (and Z) delayed by sFilterDelay = N/2 + 128 */
    float uhj_out_delay[2][512/2+128];  /* mDirectDelay */
    /* scratch */
    float resample_data[RESBUF];
    float samples[LINE];
    float filtered[LINE];      /* DeviceBase::FilteredData core/device.h */
    uint32_t mid_frames;       /* between oracle_render_begin and oracle_render_end */
    olimiter *limiter;         /* DeviceBase::Limiter */
    /* FrontStablizer (core/front_stablizer.h): MidFilter + one all-pass state per RealOut channel */
    uint32_t stab_center;      /* B200MIX_NO_SLOT = off */
    osplitter stab_mid;
    float stab_ap_z1[B200MIX_MAX_DRY_CHANNELS*2];
    /* Bs2b::bs2b_processor (core/bs2b.h:50-89): level 0 = off */
    uint32_t bs2b_level;
    float bs2b_a0_lo, bs2b_b1_lo, bs2b_a0_hi, bs2b_a1_hi, bs2b_b1_hi;
    float bs2b_hist[2][2];     /* history[ch]{lo, hi} */
    uint32_t *dc_delay;        /* DeviceBase::ChannelDelays: Buffer.size() per RealOut channel */
    float *dc_gain;
    float (*dc_buf)[LINE];
    float hrtf_samples[LINE+HIST];
    float temp[LINE], temp2[LINE];
};

int oracle_create(const b200mix_device_desc *desc, oracle_device **out)
{
    if(!desc || !out || desc->dry_channels > B200MIX_MAX_DRY_CHANNELS
        || desc->wet_channels > B200MIX_MAX_WET_CHANNELS || desc->num_sends > B200MIX_MAX_SENDS)
        return B200MIX_ERR_INVALID;
    if(ensure_tables()) return B200MIX_ERR_NOMEM;
    oracle_device *d = calloc(1, sizeof(*d));
    if(!d) return B200MIX_ERR_NOMEM;
    d->desc = *desc;
    d->stab_center = B200MIX_NO_SLOT;
    d->buffers = calloc(desc->max_buffers ? desc->max_buffers : 1, sizeof(obuffer));
    d->voices = calloc(desc->max_voices ? desc->max_voices : 1, sizeof(ovoice));
    d->dry = calloc(desc->dry_channels ? desc->dry_channels : 1, sizeof(float[LINE]));
    if(desc->post_process == B200MIX_POST_NONE) d->real = d->dry;
    else d->real = calloc(desc->real_channels ? desc->real_channels : 1, sizeof(float[LINE]));
    size_t nwet = (size_t)desc->max_slots*desc->wet_channels;
    d->wet = calloc(nwet ? nwet : 1, sizeof(float[LINE]));
    d->slots = calloc(desc->max_slots ? desc->max_slots : 1, sizeof(oslot));
    if(d->slots) for(uint32_t i = 0;i < (desc->max_slots ? desc->max_slots : 1);++i) d->slots[i].target = B200MIX_NO_SLOT;
    *out = d;
    return B200MIX_OK;
}

void oracle_destroy(oracle_device *d)
{
    if(!d) return;
    for(uint32_t i = 0;i < d->desc.max_buffers;++i) free(d->buffers[i].data);
    free(d->buffers); free(d->voices);
    if(d->real != d->dry) free(d->real);
    for(uint32_t i = 0;i < d->desc.max_slots;++i)
    {
        free(d->slots[i].ir); free(d->slots[i].hist); free(d->slots[i].cur); free(d->slots[i].tgt);
        oreverb_destroy(d->slots[i].reverb);
    }
    free(d->slots);
    olimiter_destroy(d->limiter);
    free(d->dc_delay); free(d->dc_gain); free(d->dc_buf);
    free(d->dry); free(d->wet);
    free(d->dec_coef); free(d->dec_hfscale); free(d->dec_split);
    free(d->amb_hf); free(d->amb_lf); free(d->amb_split);
    free(d);
}

int oracle_set_hrtf_decoder(oracle_device *d, uint32_t channels, uint32_t ir_size,
    const float *coeffs, const float *hf_scale, const float *splitter_coeff)
{
    if(channels != d->desc.dry_channels || ir_size > HRIR) return B200MIX_ERR_INVALID;
    d->dec_channels = channels; d->dec_ir = ir_size;
    d->dec_coef = calloc(channels, sizeof(float[HRIR][2]));
    d->dec_hfscale = calloc(channels, sizeof(float));
    d->dec_split = calloc(channels, sizeof(osplitter));
    for(uint32_t c = 0;c < channels;++c)
    {
        for(uint32_t j = 0;j < ir_size;++j)
        {
            d->dec_coef[c][j][0] = coeffs[(c*ir_size + j)*2 + 0];
            d->dec_coef[c][j][1] = coeffs[(c*ir_size + j)*2 + 1];
        }
        d->dec_hfscale[c] = hf_scale[c];
        d->dec_split[c].coeff = splitter_coeff[c];
    }
    return B200MIX_OK;
}

int oracle_set_ambi_decoder(oracle_device *d, uint32_t in_channels, const float *gains_hf,
    const float *gains_lf, float xover_coeff)
{
    if(in_channels != d->desc.dry_channels) return B200MIX_ERR_INVALID;
    const size_t n = (size_t)in_channels*d->desc.real_channels;
    d->amb_in = in_channels; d->amb_dual = gains_lf != NULL;
    d->amb_hf = malloc(n*sizeof(float)); memcpy(d->amb_hf, gains_hf, n*sizeof(float));
    if(gains_lf) { d->amb_lf = malloc(n*sizeof(float)); memcpy(d->amb_lf, gains_lf, n*sizeof(float)); }
    d->amb_split = calloc(in_channels, sizeof(osplitter));
    for(uint32_t c = 0;c < in_channels;++c) d->amb_split[c].coeff = xover_coeff;
    return B200MIX_OK;
}

/* CreateDeviceLimiter / device->Limiter = nullptr (alc/alc.cpp:1079-1091,1508,1771-1774) */
int oracle_set_limiter(oracle_device *d, const b200mix_limiter_desc *desc, uint32_t *look_ahead)
{
    olimiter_destroy(d->limiter); d->limiter = NULL;
    if(look_ahead) *look_ahead = 0;
    if(!desc) return B200MIX_OK;
    if(desc->struct_size != sizeof(*desc)) return B200MIX_ERR_INVALID;
    d->limiter = olimiter_create(desc, d->desc.real_channels, (float)d->desc.sample_rate);
    if(!d->limiter) return B200MIX_ERR_NOMEM;
    if(look_ahead) *look_ahead = olimiter_look_ahead(d->limiter);
    return B200MIX_OK;
}

/* UhjEncodeQuality (alc/alc.cpp:564-574): 0 = UhjEncoderIIR, 256/512 = UhjEncoder<N>; resets the
 * encoder state like a device reset does.  *delay = EncoderBase::getDelay(). */
int oracle_set_uhj_encoder(oracle_device *d, uint32_t filter_length, uint32_t *delay)
{
    if(filter_length != 0 && filter_length != 256 && filter_length != 512) return B200MIX_ERR_INVALID;
    if(d->desc.post_process != B200MIX_POST_UHJ && d->desc.post_process != B200MIX_POST_TSME)
        return B200MIX_ERR_INVALID;
    d->uhj_fir = filter_length;
    memset(d->uhj_f1wx, 0, sizeof(d->uhj_f1wx)); memset(d->uhj_f2wx, 0, sizeof(d->uhj_f2wx));
    memset(d->uhj_f1y, 0, sizeof(d->uhj_f1y)); memset(d->uhj_f1d, 0, sizeof(d->uhj_f1d));
    d->uhj_delay_wx = d->uhj_delay_y = d->uhj_delay_d[0] = d->uhj_delay_d[1] = 0.0f;
    memset(d->uhj_wxhist, 0, sizeof(d->uhj_wxhist));
    memset(d->uhj_in_delay, 0, sizeof(d->uhj_in_delay));
    memset(d->uhj_out_delay, 0, sizeof(d->uhj_out_delay));
    if(filter_length)
    {
        /* SegmentedFilter's desired response, core/allpass_conv.hpp:56-75 */
        const size_t N = filter_length, half = N/2;
        memset(d->uhj_fir_h, 0, sizeof(d->uhj_fir_h));
        for(size_t i = 0;i < half;++i)
        {
            const int k = (int)half - (int)(i*2 + 1);
            const double w = 2.0*ORACLE_PI/(double)(half-1) * (double)i;
            const double window = 0.3635819 - 0.4891775*cos(w) + 0.1365995*cos(2.0*w)
                - 0.0106411*cos(3.0*w);
            const double pk = ORACLE_PI * (double)k;
            d->uhj_fir_h[i*2 + 1] = window * 2.0 / pk;
        }
    }
    if(delay) *delay = filter_length ? filter_length/2 + 128 : 1;
    return B200MIX_OK;
}

/* CreateStablizer, alc/panning.cpp:160-172 */
int oracle_set_front_stabilizer(oracle_device *d, uint32_t center_channel, float splitter_coeff)
{
    const b200mix_device_desc *dd = &d->desc;
    memset(&d->stab_mid, 0, sizeof(d->stab_mid)); memset(d->stab_ap_z1, 0, sizeof(d->stab_ap_z1));
    d->stab_center = B200MIX_NO_SLOT;
    if(center_channel == B200MIX_NO_SLOT) return B200MIX_OK;
    if(dd->post_process != B200MIX_POST_AMBIDEC || center_channel >= dd->real_channels
        || dd->real_left == dd->real_right || center_channel == dd->real_left
        || center_channel == dd->real_right || dd->real_channels > B200MIX_MAX_DRY_CHANNELS*2)
        return B200MIX_ERR_INVALID;
    d->stab_center = center_channel;
    d->stab_mid.coeff = splitter_coeff;
    return B200MIX_OK;
}

/* bs2b->set_params(cf_level, rate), alc/panning.cpp:1426-1427 */
int oracle_set_bs2b(oracle_device *d, uint32_t level)
{
    if(level > 6 || d->desc.post_process != B200MIX_POST_AMBIDEC
        || d->desc.real_left == d->desc.real_right) return B200MIX_ERR_INVALID;
    memset(d->bs2b_hist, 0, sizeof(d->bs2b_hist));
    d->bs2b_level = level;
    if(level)
    {
        float c[5];
        oracle_bs2b_coeffs(level, d->desc.sample_rate, c);
        d->bs2b_a0_lo = c[0]; d->bs2b_b1_lo = c[1]; d->bs2b_a0_hi = c[2]; d->bs2b_a1_hi = c[3]; d->bs2b_b1_hi = c[4];
    }
    return B200MIX_OK;
}

/* InitDistanceComp's result (alc/panning.cpp:301-371) */
int oracle_set_distance_comp(oracle_device *d, uint32_t channels, const uint32_t *delays, const float *gains)
{
    free(d->dc_delay); free(d->dc_gain); free(d->dc_buf);
    d->dc_delay = NULL; d->dc_gain = NULL; d->dc_buf = NULL;
    if(!channels) return B200MIX_OK;
    if(channels > d->desc.real_channels || !delays || !gains) return B200MIX_ERR_INVALID;
    for(uint32_t c = 0;c < channels;++c) if(delays[c] >= LINE) return B200MIX_ERR_INVALID;
    const uint32_t rc = d->desc.real_channels;
    d->dc_delay = calloc(rc, sizeof(uint32_t)); d->dc_gain = calloc(rc, sizeof(float));
    d->dc_buf = calloc(rc, sizeof(float[LINE]));
    if(!d->dc_delay || !d->dc_gain || !d->dc_buf) return B200MIX_ERR_NOMEM;
    for(uint32_t c = 0;c < channels;++c) { d->dc_delay[c] = delays[c]; d->dc_gain[c] = gains[c]; }
    return B200MIX_OK;
}

int oracle_slot_disable(oracle_device *d, uint32_t slot)
{
    if(slot >= d->desc.max_slots) return B200MIX_ERR_INVALID;
    oslot *s = &d->slots[slot];
    free(s->ir); free(s->hist); free(s->cur); free(s->tgt);
    oreverb_destroy(s->reverb);
    if(s->efx) oefx_free(s->efx);
    const uint32_t target = s->target;       /* the target belongs to the slot, not to its effect */
    memset(s, 0, sizeof(*s));
    s->target = target;
    return B200MIX_OK;
}

int oracle_slot_reverb(oracle_device *d, uint32_t slot, const b200mix_reverb_params *params)
{
    if(slot >= d->desc.max_slots || !params || params->struct_size != sizeof(*params))
        return B200MIX_ERR_INVALID;
    oracle_slot_disable(d, slot);
    oslot *s = &d->slots[slot];
    s->reverb = oreverb_create(params);
    if(!s->reverb) return B200MIX_ERR_NOMEM;
    s->type = B200MIX_EFFECT_REVERB; s->channels = 8;
    return B200MIX_OK;
}

/* The other EFX effects: deviceUpdate + update on the first call / a change of type, update()
 * afterwards (efx_oracle.cpp). */
int oracle_slot_efx(oracle_device *d, uint32_t slot, const b200mix_efx_props *props,
    const b200mix_efx_target *target)
{
    if(slot >= d->desc.max_slots || !props || !target || props->struct_size != sizeof(*props)
        || target->struct_size != sizeof(*target))
        return B200MIX_ERR_INVALID;
    oslot *s = &d->slots[slot];
    if(s->efx && s->type == props->type && oefx_update(s->efx, props, target) == B200MIX_OK)
        return B200MIX_OK;
    oracle_slot_disable(d, slot);
    int rc = B200MIX_OK;
    s->efx = oefx_create(props, target, &rc);
    if(!s->efx) return rc;
    s->type = props->type; s->channels = 1;
    return B200MIX_OK;
}

/* EffectSlotBase::Target (AL_EFFECTSLOT_TARGET_SOFT): the slot's output feeds `target`'s Wet
 * buffer instead of the Dry mix.  Chains must be acyclic (al/auxeffectslot.cpp rejects loops). */
int oracle_slot_target(oracle_device *d, uint32_t slot, uint32_t target)
{
    if(slot >= d->desc.max_slots || (target != B200MIX_NO_SLOT && target >= d->desc.max_slots))
        return B200MIX_ERR_INVALID;
    for(uint32_t t = target, hops = 0;t != B200MIX_NO_SLOT;t = d->slots[t].target)
        if(t == slot || ++hops > d->desc.max_slots) return B200MIX_ERR_INVALID;
    d->slots[slot].target = target;
    return B200MIX_OK;
}

int oracle_slot_reverb_update(oracle_device *d, uint32_t slot, const b200mix_reverb_params *params,
    uint32_t full_update)
{
    if(slot >= d->desc.max_slots || !params || params->struct_size != sizeof(*params))
        return B200MIX_ERR_INVALID;
    oslot *s = &d->slots[slot];
    if(s->type != B200MIX_EFFECT_REVERB) return B200MIX_ERR_INVALID;
    oreverb_update(s->reverb, params, full_update != 0);
    return B200MIX_OK;
}

int oracle_slot_convolution(oracle_device *d, uint32_t slot, uint32_t ir_channels,
    uint32_t ir_frames, const float *ir)
{
    if(slot >= d->desc.max_slots || !ir_channels || !ir_frames || !ir) return B200MIX_ERR_INVALID;
    oracle_slot_disable(d, slot);
    oslot *s = &d->slots[slot];
    s->type = B200MIX_EFFECT_CONVOLUTION; s->channels = ir_channels; s->frames = ir_frames;
    s->ir = malloc(sizeof(float)*(size_t)ir_channels*ir_frames);
    memcpy(s->ir, ir, sizeof(float)*(size_t)ir_channels*ir_frames);
    s->hist = calloc((size_t)ir_frames + LINE, sizeof(float));
    s->cur = calloc((size_t)ir_channels*B200MIX_MAX_DRY_CHANNELS, sizeof(float));
    s->tgt = calloc((size_t)ir_channels*B200MIX_MAX_DRY_CHANNELS, sizeof(float));
    return B200MIX_OK;
}

int oracle_slot_output_gains(oracle_device *d, uint32_t slot, uint32_t lines, const float *gains)
{
    if(slot >= d->desc.max_slots || !d->slots[slot].type || lines != d->slots[slot].channels)
        return B200MIX_ERR_INVALID;
    oslot *s = &d->slots[slot];
    /* gains address the slot's output target: the Dry mix or the target slot's Wet mix */
    const uint32_t width = s->target != B200MIX_NO_SLOT ? d->desc.wet_channels : d->desc.dry_channels;
    if(s->type == B200MIX_EFFECT_REVERB)
    {
        oreverb_set_gains(s->reverb, gains, width);
        return B200MIX_OK;
    }
    for(uint32_t c = 0;c < lines;++c)
        for(uint32_t o = 0;o < width;++o)
            s->tgt[c*B200MIX_MAX_DRY_CHANNELS + o] = gains[c*width + o];
    return B200MIX_OK;
}

static size_t sample_bytes(uint32_t type)
{
    switch(type)
    {
    case B200MIX_FMT_U8: case B200MIX_FMT_MULAW: case B200MIX_FMT_ALAW: return 1;
    case B200MIX_FMT_I16: return 2;
    case B200MIX_FMT_I32: case B200MIX_FMT_F32: return 4;
    case B200MIX_FMT_F64: return 8;
    }
    return 0;
}

int oracle_buffer_data(oracle_device *d, uint32_t buffer, uint32_t type, uint32_t channels,
    uint32_t frames, const void *data, size_t bytes)
{
    if(buffer >= d->desc.max_buffers || !sample_bytes(type) || channels < 1) return B200MIX_ERR_INVALID;
    if(bytes < (size_t)frames*channels*sample_bytes(type)) return B200MIX_ERR_INVALID;
    obuffer *b = &d->buffers[buffer];
    free(b->data);
    b->data = malloc(bytes ? bytes : 1);
    memcpy(b->data, data, bytes);
    b->type = type; b->channels = channels; b->frames = frames;
    return B200MIX_OK;
}

/* LoadSamples<IMA4Data> / LoadSamples<MSADPCMData> (core/voice.cpp:289-484; tables :199-238)
 * evaluated for whole blocks at upload time: the decoders are integer recurrences, so the
 * int16 results are exactly the values the reference's mixer converts with /32768.0f. */
static const int ima_step[89] = {
    7,8,9,10,11,12,13,14,16,17,19,21,23,25,28,31,34,37,41,45,50,55,60,66,73,80,88,97,107,118,130,143,
    157,173,190,209,230,253,279,307,337,371,408,449,494,544,598,658,724,796,876,963,1060,1166,1282,
    1411,1552,1707,1878,2066,2272,2499,2749,3024,3327,3660,4026,4428,4871,5358,5894,6484,7132,7845,
    8630,9493,10442,11487,12635,13899,15289,16818,18500,20350,22358,24633,27086,29794,32767};
static const int ima_codeword[16] = {1,3,5,7,9,11,13,15,-1,-3,-5,-7,-9,-11,-13,-15};
static const int ima_adjust[16] = {-1,-1,-1,-1,2,4,6,8,-1,-1,-1,-1,2,4,6,8};
static const int ms_adaption[16] = {230,230,230,230,307,409,512,614,768,614,512,409,307,230,230,230};
static const int ms_coeff[7][2] = {{256,0},{512,-256},{0,0},{192,64},{240,0},{460,-208},{392,-232}};
static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static int rd_s16(const uint8_t *p) { return (int)(int16_t)((uint16_t)p[0] | ((uint16_t)p[1] << 8)); }

int oracle_buffer_data_adpcm(oracle_device *d, uint32_t buffer, uint32_t type, uint32_t channels,
    uint32_t spb, uint32_t blocks, const void *data, size_t bytes)
{
    const int ms = type == B200MIX_FMT_MSADPCM;
    if((type != B200MIX_FMT_IMA4 && !ms) || channels < 1 || channels > 2 || !data) return B200MIX_ERR_INVALID;
    if(ms ? (spb < 2 || (spb & 1)) : (spb < 1 || ((spb-1) & 7))) return B200MIX_ERR_INVALID;
    const size_t blockBytes = ms ? ((size_t)(spb-2)/2 + 7)*channels : ((size_t)(spb-1)/2 + 4)*channels;
    if(bytes < blockBytes*blocks) return B200MIX_ERR_INVALID;
    int16_t *pcm = malloc(sizeof(int16_t)*(size_t)blocks*spb*channels);
    if(!pcm) return B200MIX_ERR_NOMEM;
    const uint8_t *src = data;
    int16_t *dst = pcm;
    for(uint32_t b = 0;b < blocks;++b, src += blockBytes, dst += (size_t)spb*channels)
        for(uint32_t c = 0;c < channels;++c)
        {
            if(!ms)
            {
                /* :309-351 */
                int sample = rd_s16(src + c*4), idx = clampi(rd_s16(src + c*4 + 2), 0, 88);
                const uint8_t *nib = src + (channels + c)*4;
                dst[c] = (int16_t)sample;
                for(uint32_t n = 0;n + 1 < spb;++n)
                {
                    const uint32_t shift = (n&1)*4, word = (n>>1) & ~3u;
                    const uint32_t code = (nib[word*channels + ((n>>1)&3)] >> shift) & 0xf;
                    sample = clampi(sample + ima_codeword[code]*ima_step[idx]/8, -32768, 32767);
                    idx = clampi(idx + ima_adjust[code], 0, 88);
                    dst[(size_t)(n+1)*channels + c] = (int16_t)sample;
                }
            }
            else
            {
                /* :400-462 */
                const uint32_t pred = src[c] < 6 ? src[c] : 6;
                int scale = rd_s16(src + channels + 2*c);
                int h0 = rd_s16(src + 3*channels + 2*c), h1 = rd_s16(src + 5*channels + 2*c);
                const uint8_t *nib = src + 7*channels;
                dst[c] = (int16_t)h1;
                dst[(size_t)channels + c] = (int16_t)h0;
                uint32_t off = c;
                for(uint32_t n = 2;n < spb;++n, off += channels)
                {
                    const int nval = (nib[off>>1] >> (((off&1)^1)*4)) & 0xf;
                    const int p = ((nval^0x08) - 0x08)*scale;
                    const int diff = (h0*ms_coeff[pred][0] + h1*ms_coeff[pred][1])/256;
                    const int sample = clampi(p + diff, -32768, 32767);
                    h1 = h0; h0 = sample;
                    scale = ms_adaption[nval]*scale/256; if(scale < 16) scale = 16;
                    dst[(size_t)n*channels + c] = (int16_t)sample;
                }
            }
        }
    const int rc = oracle_buffer_data(d, buffer, B200MIX_FMT_I16, channels, blocks*spb, pcm,
        sizeof(int16_t)*(size_t)blocks*spb*channels);
    free(pcm);
    return rc;
}

int oracle_buffer_free(oracle_device *d, uint32_t buffer)
{
    if(buffer >= d->desc.max_buffers) return B200MIX_ERR_INVALID;
    free(d->buffers[buffer].data);
    memset(&d->buffers[buffer], 0, sizeof(obuffer));
    return B200MIX_OK;
}

int oracle_voices_update(oracle_device *d, uint32_t n, const b200mix_voice_params *params,
    const float *hrtf_coeffs, const float *dry_gains, const float *send_gains)
{
    const uint32_t ir = d->desc.ir_size, cd = d->desc.dry_channels;
    const uint32_t ns = d->desc.num_sends, cw = d->desc.wet_channels;
    for(uint32_t i = 0;i < n;++i)
    {
        const b200mix_voice_params *p = &params[i];
        if(p->voice >= d->desc.max_voices) return B200MIX_ERR_INVALID;
        ovoice *v = &d->voices[p->voice];
        if(p->flags & B200MIX_VF_RESET)
        {
            /* Voice::prepare + InitVoice: core/voice.cpp:1235-1400, al/source.cpp:639-669 */
            /* the buffer list is the source's, not the voice's: it survives a restart */
            const uint32_t qc = v->q_count, ql = v->q_loop;
            uint32_t qi[B200MIX_MAX_QUEUE];
            memcpy(qi, v->q_items, sizeof(qi));
            memset(v, 0, sizeof(*v));
            v->q_count = qc; v->q_loop = ql; v->q_head = 0;
            memcpy(v->q_items, qi, sizeof(qi));
            for(uint32_t f = 0;f < 1 + B200MIX_MAX_SENDS;++f)
            { obiquad_reset(&v->filt[f].lp); obiquad_reset(&v->filt[f].hp); }
            v->pos = p->position; v->frac = p->position_frac;
            v->fading = (p->flags & B200MIX_VF_FADING) != 0;
            v->have_buffer = 1;
        }
        if(p->flags & B200MIX_VF_STOPPED) v->state = 0;
        else if(p->flags & B200MIX_VF_STOPPING) v->state = 2;
        else if(p->flags & B200MIX_VF_PLAYING) v->state = 1;
        v->flags = p->flags & (B200MIX_VF_STATIC|B200MIX_VF_LOOPING|B200MIX_VF_HRTF|B200MIX_VF_CHANNEL(0xff));
        v->buffer = p->buffer; v->resampler = p->resampler;
        if(p->buffer == B200MIX_NO_BUFFER) { v->buffer = 0; v->have_buffer = 0; }   /* alc/alu.cpp:2071 */
        v->loop_start = p->loop_start; v->loop_end = p->loop_end; v->step = p->step;
        v->tgt_delay[0] = p->hrtf_delay[0]; v->tgt_delay[1] = p->hrtf_delay[1];
        v->tgt_gain = p->hrtf_gain;
        memcpy(v->send_slot, p->send_slot, sizeof(v->send_slot));
        if(hrtf_coeffs)
            for(uint32_t j = 0;j < ir;++j)
            {
                v->tgt_coef[j][0] = hrtf_coeffs[((size_t)i*ir + j)*2 + 0];
                v->tgt_coef[j][1] = hrtf_coeffs[((size_t)i*ir + j)*2 + 1];
            }
        if(dry_gains)
            for(uint32_t c = 0;c < cd;++c) v->dry_tgt[c] = dry_gains[(size_t)i*cd + c];
        if(send_gains)
            for(uint32_t s = 0;s < ns;++s)
                for(uint32_t c = 0;c < cw;++c)
                    v->send_tgt[s][c] = send_gains[((size_t)i*ns + s)*cw + c];
    }
    return B200MIX_OK;
}

int oracle_voice_queue(oracle_device *d, uint32_t voice, uint32_t count, const uint32_t *buffers,
    uint32_t loop_index)
{
    if(voice >= d->desc.max_voices || count > B200MIX_MAX_QUEUE || (count && !buffers))
        return B200MIX_ERR_INVALID;
    if(loop_index != B200MIX_NO_LOOP && loop_index >= count) return B200MIX_ERR_INVALID;
    ovoice *v = &d->voices[voice];
    for(uint32_t i = 0;i < count;++i)
    {
        if(buffers[i] >= d->desc.max_buffers) return B200MIX_ERR_INVALID;
        v->q_items[i] = buffers[i];
    }
    v->q_count = count; v->q_head = 0; v->q_loop = loop_index;
    v->have_buffer = count > 0;
    return B200MIX_OK;
}

static void load_samples(float *dst, size_t count, const obuffer *b, size_t offset);

/* next item of the queue: mNext, or the loop item past the end (core/voice.cpp:563-565) */
static uint32_t queue_next(const ovoice *v, uint32_t item)
{
    if(item + 1 < v->q_count) return item + 1;
    return v->q_loop;            /* B200MIX_NO_LOOP ends the list */
}

/* LoadBufferQueue, core/voice.cpp:546-595 */
static void load_buffer_queue(const oracle_device *d, const ovoice *v, size_t dataPosInt, float *dst,
    size_t count)
{
    float lastSample = 0.0f;
    uint32_t item = v->q_head;
    while(item != B200MIX_NO_LOOP && count > 0)
    {
        const obuffer *b = &d->buffers[v->q_items[item]];
        if(dataPosInt >= b->frames)
        {
            dataPosInt -= b->frames;
            item = queue_next(v, item);
            continue;
        }
        size_t remaining = b->frames - dataPosInt;
        if(remaining > count) remaining = count;
        load_samples(dst, remaining, b, dataPosInt);
        lastSample = dst[remaining-1];
        dst += remaining; count -= remaining;
        if(!count) break;
        dataPosInt = 0;
        item = queue_next(v, item);
    }
    for(size_t i = 0;i < count;++i) dst[i] = lastSample;
}

/* BiquadInterpFilter::setParams after SetParams filled mTargetCoeffs
 * (core/filters/biquad.cpp:36-43,123-147): check_set's 1/64 hysteresis decides whether
 * the 8-step interpolation starts. */
static void obiquad_set_target(obiquad *f, const float tgt[5])
{
    int is_diff = 0;
    for(int k = 0;k < 5;++k)
    {
        is_diff |= !(fabsf(tgt[k] - f->tgt[k]) <= 0.015625f);
        f->tgt[k] = tgt[k];
    }
    if(!is_diff)
    {
        if(f->counter <= 0) { f->counter = 0; memcpy(f->cur, f->tgt, sizeof(f->cur)); }
    }
    else if(f->counter >= 0)
        f->counter = 8*32;                         /* InterpSteps*SamplesPerStep */
    else
    { f->counter = 0; memcpy(f->cur, f->tgt, sizeof(f->cur)); }
}

int oracle_voices_filters(oracle_device *d, uint32_t n, const b200mix_voice_filter *filters)
{
    if(!d->filters_seen && n)
    {
        /* The library allocates its filter records on the first call and starts every one of them
         * in BiquadInterpFilter::reset's state (mCounter = -1: the first target applies at once),
         * also for voices that have been playing unfiltered.  A host that forwards filters only
         * once one is in use therefore first sends the shelves the reference held so far, then the
         * new targets (integration/b200mix_seam.cpp). */
        for(uint32_t v = 0;v < d->desc.max_voices;++v)
            for(uint32_t f = 0;f < 1 + B200MIX_MAX_SENDS;++f)
            { obiquad_reset(&d->voices[v].filt[f].lp); obiquad_reset(&d->voices[v].filt[f].hp); d->voices[v].filt[f].active = 0; }
        d->filters_seen = 1;
    }
    for(uint32_t i = 0;i < n;++i)
    {
        const b200mix_voice_filter *q = &filters[i];
        if(q->voice >= d->desc.max_voices || q->path > d->desc.num_sends) return B200MIX_ERR_INVALID;
        ovoice *v = &d->voices[q->voice];
        v->filt[q->path].active = q->active != 0;
        obiquad_set_target(&v->filt[q->path].lp, q->lowpass);
        obiquad_set_target(&v->filt[q->path].hp, q->highpass);
    }
    return B200MIX_OK;
}

/* BiquadFilter::SetParams behind setParamsFromSlope (core/filters/biquad.h:92-97 with
 * rcpQFromSlope :61-62; biquad.cpp:48-129). */
int oracle_biquad_coeffs(uint32_t type, float f0norm, float gain, float slope, float coeffs[5])
{
    if(type > 5u) return B200MIX_ERR_INVALID;
    gain = fmaxf(gain, 0.001f);
    const float rcpQ = sqrtf((gain + 1.0f/gain)*(1.0f/slope - 1.0f) + 2.0f);
    gain = fmaxf(gain, 0.00001f);
    const float w0 = 3.14159265358979323846f*2.0f * fminf(f0norm, 0.49f);
    const float sin_w0 = sinf(w0), cos_w0 = cosf(w0);
    const float alpha = sin_w0/2.0f * rcpQ;
    float a[3] = {1.0f, 0.0f, 0.0f}, b[3] = {1.0f, 0.0f, 0.0f}, sg;
    switch(type)
    {
    case 0: /* HighShelf */
        sg = 2.0f * sqrtf(gain) * alpha;
        b[0] =       gain*((gain+1.0f) + (gain-1.0f)*cos_w0 + sg);
        b[1] = -2.0f*gain*((gain-1.0f) + (gain+1.0f)*cos_w0     );
        b[2] =       gain*((gain+1.0f) + (gain-1.0f)*cos_w0 - sg);
        a[0] =             (gain+1.0f) - (gain-1.0f)*cos_w0 + sg;
        a[1] =  2.0f*     ((gain-1.0f) - (gain+1.0f)*cos_w0     );
        a[2] =             (gain+1.0f) - (gain-1.0f)*cos_w0 - sg;
        break;
    case 1: /* LowShelf */
        sg = 2.0f * sqrtf(gain) * alpha;
        b[0] =       gain*((gain+1.0f) - (gain-1.0f)*cos_w0 + sg);
        b[1] =  2.0f*gain*((gain-1.0f) - (gain+1.0f)*cos_w0     );
        b[2] =       gain*((gain+1.0f) - (gain-1.0f)*cos_w0 - sg);
        a[0] =             (gain+1.0f) + (gain-1.0f)*cos_w0 + sg;
        a[1] = -2.0f*     ((gain-1.0f) + (gain+1.0f)*cos_w0     );
        a[2] =             (gain+1.0f) + (gain-1.0f)*cos_w0 - sg;
        break;
    case 2: /* Peaking */
        b[0] = 1.0f + alpha*gain; b[1] = -2.0f*cos_w0; b[2] = 1.0f - alpha*gain;
        a[0] = 1.0f + alpha/gain; a[1] = -2.0f*cos_w0; a[2] = 1.0f - alpha/gain;
        break;
    case 3: /* LowPass */
        b[0] = (1.0f - cos_w0)/2.0f; b[1] = 1.0f - cos_w0; b[2] = (1.0f - cos_w0)/2.0f;
        a[0] = 1.0f + alpha; a[1] = -2.0f*cos_w0; a[2] = 1.0f - alpha;
        break;
    case 4: /* HighPass */
        b[0] = (1.0f + cos_w0)/2.0f; b[1] = -(1.0f + cos_w0); b[2] = (1.0f + cos_w0)/2.0f;
        a[0] = 1.0f + alpha; a[1] = -2.0f*cos_w0; a[2] = 1.0f - alpha;
        break;
    default: /* BandPass */
        b[0] = alpha; b[1] = 0.0f; b[2] = -alpha;
        a[0] = 1.0f + alpha; a[1] = -2.0f*cos_w0; a[2] = 1.0f - alpha;
        break;
    }
    coeffs[0] = b[0]/a[0]; coeffs[1] = b[1]/a[0]; coeffs[2] = b[2]/a[0];
    coeffs[3] = a[1]/a[0]; coeffs[4] = a[2]/a[0];
    return B200MIX_OK;
}

/* BiquadFilter::dualProcess, core/filters/biquad.cpp:254-281 (transposed direct form II,
 * f0 then f1 per sample). */
static void obiquad_dual_run(obiquad *f0, obiquad *f1, const float *src, float *dst, size_t n)
{
    float z01 = f0->z[0], z02 = f0->z[1], z11 = f1->z[0], z12 = f1->z[1];
    const float *c0 = f0->cur, *c1 = f1->cur;
    for(size_t i = 0;i < n;++i)
    {
        const float x0 = src[i];
        const float y0 = x0*c0[0] + z01;
        z01 = x0*c0[1] - y0*c0[3] + z02;
        z02 = x0*c0[2] - y0*c0[4];
        const float x1 = y0;
        const float y1 = x1*c1[0] + z11;
        z11 = x1*c1[1] - y1*c1[3] + z12;
        z12 = x1*c1[2] - y1*c1[4];
        dst[i] = y1;
    }
    f0->z[0] = z01; f0->z[1] = z02; f1->z[0] = z11; f1->z[1] = z12;
}

/* BiquadInterpFilter::dualProcess, core/filters/biquad.cpp:283-343. */
static void obiquad_dual_interp(obiquad *f0, obiquad *f1, const float *src, float *dst, size_t n)
{
    const int maxcounter = f0->counter > f1->counter ? f0->counter : f1->counter;
    if(maxcounter > 0)
    {
        int counter = maxcounter / 32;
        size_t steprem = (size_t)(32 - (maxcounter & 31));
        while(counter > 0)
        {
            const size_t td = steprem < n ? steprem : n;
            obiquad_dual_run(f0, f1, src, dst, td);
            steprem -= td;
            if(steprem)
            {
                steprem = 32 - steprem;
                f0->counter = (counter*32) | (int)steprem;
                f1->counter = f0->counter;
                return;
            }
            src += td; dst += td; n -= td;
            steprem = 32;
            --counter;
            if(!counter)
            {
                f0->counter = 0; memcpy(f0->cur, f0->tgt, sizeof(f0->cur));
                f1->counter = 0; memcpy(f1->cur, f1->tgt, sizeof(f1->cur));
                break;
            }
            const float a = 1.0f / (float)(counter+1);
            for(int k = 0;k < 5;++k)
            {
                f0->cur[k] = lerpf(f0->cur[k], f0->tgt[k], a);
                f1->cur[k] = lerpf(f1->cur[k], f1->tgt[k], a);
            }
            if(!n)
            {
                f0->counter = counter*32;
                f1->counter = f0->counter;
                return;
            }
        }
    }
    obiquad_dual_run(f0, f1, src, dst, n);
}

/* DoFilters, core/voice.cpp:255-268 */
static const float *do_filters(const oracle_device *d, ovoice *v, uint32_t path, float *dst, const float *src, size_t n)
{
    if(v->filt[path].active)
    {
        obiquad_dual_interp(&v->filt[path].lp, &v->filt[path].hp, src, dst, n);
        return dst;
    }
    /* ABI lifecycle, as the library has it: until the first b200mix_voices_filters call there are
     * no filter records to clear */
    if(d->filters_seen)
    {
        obiquad_clear(&v->filt[path].lp);
        obiquad_clear(&v->filt[path].hp);
    }
    return src;
}

/* ITU-T G.711 expansion; equals muLawDecompressionTable / aLawDecompressionTable
 * (core/fmt_traits.h:12-81), computed instead of tabulated. */
static int mulaw_decode(uint8_t b)
{
    const unsigned u = (~b) & 0xffu;
    const int s = (int)((((u & 0x0fu)<<3) + 0x84u) << ((u>>4)&7u)) - 0x84;
    return (u & 0x80u) ? -s : s;
}
static int alaw_decode(uint8_t b)
{
    const unsigned a = b ^ 0x55u;
    const unsigned e = (a>>4)&7u, m = a & 0x0fu;
    const int s = (e == 0) ? (int)((m<<4) + 8u) : (int)(((m<<4) + 0x108u) << (e-1u));
    return (a & 0x80u) ? s : -s;
}

/* ---- sample loading: core/voice.cpp:271-288 (LoadSamples), fmt_traits.h:88-161 */
static float to_float(const obuffer *b, size_t idx)
{
    switch(b->type)
    {
    case B200MIX_FMT_U8: return ((float)((const uint8_t*)b->data)[idx]-128.0f) * (1.0f/128.0f);
    case B200MIX_FMT_I16: return (float)((const int16_t*)b->data)[idx] * (1.0f/32768.0f);
    case B200MIX_FMT_I32: return (float)((const int32_t*)b->data)[idx] * (1.0f/2147483648.0f);
    case B200MIX_FMT_F32: return ((const float*)b->data)[idx];
    case B200MIX_FMT_F64: return (float)((const double*)b->data)[idx];
    case B200MIX_FMT_MULAW: return (float)mulaw_decode(((const uint8_t*)b->data)[idx]) * (1.0f/32768.0f);
    case B200MIX_FMT_ALAW: return (float)alaw_decode(((const uint8_t*)b->data)[idx]) * (1.0f/32768.0f);
    }
    return 0.0f;
}
static size_t g_src_channel;   /* srcChannel of the voice being loaded (single-threaded oracle) */
static void load_samples(float *dst, size_t count, const obuffer *b, size_t offset)
{
    const size_t ch = g_src_channel < b->channels ? g_src_channel : 0;
    for(size_t i = 0;i < count;++i)
        dst[i] = to_float(b, (offset+i)*b->channels + ch);
}

/* LoadBufferStatic, core/voice.cpp:500-544 */
static void load_buffer_static(const obuffer *b, int looping, uint32_t loop_start,
    uint32_t loop_end, size_t dataPosInt, float *dst, size_t count)
{
    if(!looping)
    {
        float lastSample = 0.0f;
        if(b->frames > dataPosInt)
        {
            size_t remaining = b->frames - dataPosInt;
            if(remaining > count) remaining = count;
            load_samples(dst, remaining, b, dataPosInt);
            lastSample = dst[remaining-1];
            dst += remaining; count -= remaining;
        }
        for(size_t i = 0;i < count;++i) dst[i] = lastSample;
    }
    else
    {
        const size_t loopStart = loop_start, loopEnd = loop_end;
        const size_t intPos = (dataPosInt < loopEnd) ? dataPosInt
            : (((dataPosInt-loopStart)%(loopEnd-loopStart)) + loopStart);
        size_t remaining = loopEnd-intPos;
        if(remaining > count) remaining = count;
        load_samples(dst, remaining, b, intPos);
        dst += remaining; count -= remaining;
        const size_t loopSize = loopEnd - loopStart;
        while(count > 0)
        {
            const size_t toFill = (count < loopSize) ? count : loopSize;
            load_samples(dst, toFill, b, loopStart);
            dst += toFill; count -= toFill;
        }
    }
}

/* CalculateBufferSize, core/voice.cpp:601-640 */
static void calc_buffer_size(uint32_t fracPos, uint32_t increment, uint32_t dstRemaining,
    uint32_t *dst, uint32_t *src)
{
    const uint32_t SrcSizeMax = RESBUF - EDGE;
    const uint32_t ext = increment <= FRAC_ONE;
    const uint64_t srcSize64 = (((uint64_t)(dstRemaining - ext)*increment + fracPos) >> FRAC_BITS)
        + ext + EDGE;
    if(srcSize64 <= SrcSizeMax) { *dst = dstRemaining; *src = (uint32_t)srcSize64; return; }
    const uint64_t dstSize64 = (((uint64_t)(SrcSizeMax - EDGE)<<FRAC_BITS) - fracPos) / increment;
    if(dstSize64 < dstRemaining) { *dst = (uint32_t)dstSize64 & ~3u; *src = SrcSizeMax; return; }
    *dst = dstRemaining; *src = SrcSizeMax;
}

static int32_t add_sat_i32(int32_t a, int32_t b)
{
    int64_t r = (int64_t)a + b;
    if(r > INT32_MAX) r = INT32_MAX;
    if(r < INT32_MIN) r = INT32_MIN;
    return (int32_t)r;
}

/* LoadResampledSamples for one mono static voice, core/voice.cpp:642-822 */
static void load_resampled(oracle_device *d, ovoice *v, int vstate, int looping,
    uint32_t samplesToMix)
{
    float *rd = d->resample_data;
    float *srcBuffer = rd + EDGE;
    memcpy(rd, v->prev, sizeof(v->prev));
    int32_t intPos = v->pos;
    uint32_t fracPos = v->frac;
    const uint32_t increment = v->step;
    const int is_queue = !(v->flags & B200MIX_VF_STATIC);
    const obuffer *buf = v->have_buffer ? &d->buffers[is_queue ? v->q_items[v->q_head] : v->buffer] : NULL;
    g_src_channel = (v->flags >> 16) & 0xffu;

    for(uint32_t loaded = 0;loaded < samplesToMix;)
    {
        uint32_t dstn, srcn;
        calc_buffer_size(fracPos, increment, samplesToMix-loaded, &dstn, &srcn);

        uint32_t srcSampleDelay = 0;
        int silent = 0;
        if(intPos < 0)
        {
            srcSampleDelay = (uint32_t)(-intPos);
            if(srcSampleDelay >= srcn)
            {
                memset(d->samples+loaded, 0, sizeof(float)*dstn);
                memset(srcBuffer, 0, sizeof(float)*srcn);
                loaded += dstn;
                if(loaded < samplesToMix)
                {
                    fracPos += dstn*increment;
                    const uint32_t srcOffset = fracPos >> FRAC_BITS;
                    fracPos &= FRAC_MASK;
                    intPos = add_sat_i32(intPos, (int32_t)srcOffset);
                }
                silent = 1;
            }
            else
                memset(srcBuffer, 0, sizeof(float)*srcSampleDelay);
        }
        if(silent) continue;

        if(!buf)
        {
            /* voice ended prematurely: hold the sample closest to 0, voice.cpp:704-719 */
            const uint32_t avail = (srcn < EDGE) ? srcn : EDGE;
            const uint32_t tofill = (srcn > EDGE) ? srcn : EDGE;
            uint32_t best = 0;
            for(uint32_t i = 1;i < avail;++i)
                if(fabsf(srcBuffer[i]) < fabsf(srcBuffer[best])) best = i;
            for(uint32_t i = best+1;i < tofill;++i) srcBuffer[i] = srcBuffer[best];
        }
        else
        {
            const uint32_t uintPos = (intPos < 0) ? 0u : (uint32_t)intPos;
            if(is_queue)
                load_buffer_queue(d, v, uintPos, srcBuffer+srcSampleDelay, srcn-srcSampleDelay);
            else
                load_buffer_static(buf, looping, v->loop_start, v->loop_end, uintPos,
                    srcBuffer+srcSampleDelay, srcn-srcSampleDelay);
        }

        if(increment == FRAC_ONE && fracPos == 0)
            memcpy(d->samples+loaded, srcBuffer, sizeof(float)*dstn); /* bypass, :764-766 */
        else
            oracle_resample(v->resampler, increment, fracPos, rd, d->samples+loaded, dstn);

        if(vstate == 1)
        {
            const uint32_t loadEnd = loaded + dstn;
            if(samplesToMix > loaded && samplesToMix <= loadEnd)
            {
                const size_t dstOffset = samplesToMix - loaded;
                const size_t srcOffset = (dstOffset*increment + fracPos) >> FRAC_BITS;
                memmove(v->prev, rd+srcOffset, sizeof(v->prev));
            }
        }

        loaded += dstn;
        if(loaded < samplesToMix)
        {
            fracPos += dstn*increment;
            const uint32_t srcOffset = fracPos >> FRAC_BITS;
            fracPos &= FRAC_MASK;
            if(intPos < 0) intPos += (int32_t)srcOffset;
            else intPos = add_sat_i32(intPos, (int32_t)srcOffset);
            memmove(rd, rd+srcOffset, sizeof(float)*PAD);
        }
    }
}

/* MixLine + Mix_C, core/mixer/mixer_c.cpp:150-186,247-258 */
static void mix_line(const float *in, size_t n, float *dst, float *cur, float target, float delta,
    size_t fade_len, size_t counter)
{
    const float step = (target - *cur) * delta;
    size_t pos = 0;
    if(fabsf(step) > 1.1920929e-07f)
    {
        const float gain = *cur;
        float step_count = 0.0f;
        for(;pos < fade_len;++pos)
        {
            dst[pos] += in[pos] * (gain + step*step_count);
            step_count += 1.0f;
        }
        if(fade_len < counter)
        {
            *cur = gain + step*step_count;
            return;
        }
    }
    *cur = target;
    if(!(fabsf(target) > SILENCE_THRESHOLD))
        return;
    for(;pos < n;++pos)
        dst[pos] += in[pos]*target;
}

static void mix_samples(const float *in, size_t n, float (*out)[LINE], size_t nchan, float *cur,
    const float *tgt, size_t counter)
{
    const float delta = (counter > 0) ? 1.0f/(float)counter : 0.0f;
    const size_t fade_len = (counter < n) ? counter : n;
    for(size_t c = 0;c < nchan;++c)
        mix_line(in, n, out[c], &cur[c], tgt[c], delta, fade_len, counter);
}

/* ApplyCoeffs, mixer_c.cpp:139-148 */
static void apply_coeffs(float (*values)[2], size_t ir, const float (*coeffs)[2], float left,
    float right)
{
    for(size_t c = 0;c < ir;++c)
    {
        values[c][0] = values[c][0] + coeffs[c][0]*left;
        values[c][1] = values[c][1] + coeffs[c][1]*right;
    }
}

/* MixHrtfBase, core/mixer/hrtfbase.h:17-42 */
static void mix_hrtf(const float *in, float (*accum)[2], size_t ir, const float (*coeffs)[2],
    const uint32_t delay[2], float gain, float gainstep, size_t todo)
{
    size_t ldelay = HIST - delay[0], rdelay = HIST - delay[1];
    float stepcount = 0.0f;
    for(size_t i = 0;i < todo;++i)
    {
        const float g = gain + gainstep*stepcount;
        const float left = in[ldelay++] * g;
        const float right = in[rdelay++] * g;
        apply_coeffs(accum+i, ir, coeffs, left, right);
        stepcount += 1.0f;
    }
}

/* MixHrtfBlendBase, core/mixer/hrtfbase.h:44-89 */
static void mix_hrtf_blend(const float *in, float (*accum)[2], size_t ir,
    const float (*oldc)[2], const uint32_t olddelay[2], float oldgain,
    const float (*newc)[2], const uint32_t newdelay[2], float newGainStep, size_t todo)
{
    const float oldGainStep = oldgain / (float)todo;
    if(oldgain > SILENCE_THRESHOLD)
    {
        size_t ldelay = HIST - olddelay[0], rdelay = HIST - olddelay[1];
        float stepcount = (float)todo;
        for(size_t i = 0;i < todo;++i)
        {
            const float g = oldGainStep*stepcount;
            const float left = in[ldelay++] * g;
            const float right = in[rdelay++] * g;
            apply_coeffs(accum+i, ir, oldc, left, right);
            stepcount -= 1.0f;
        }
    }
    if(newGainStep*(float)todo > SILENCE_THRESHOLD)
    {
        size_t ldelay = HIST+1 - newdelay[0], rdelay = HIST+1 - newdelay[1];
        float stepcount = 1.0f;
        for(size_t i = 1;i < todo;++i)
        {
            const float g = newGainStep*stepcount;
            const float left = in[ldelay++] * g;
            const float right = in[rdelay++] * g;
            apply_coeffs(accum+i, ir, newc, left, right);
            stepcount += 1.0f;
        }
    }
}

/* DoHrtfMix, core/voice.cpp:827-902 (outPos == 0) */
static void do_hrtf_mix(oracle_device *d, ovoice *v, const float *samples, size_t n,
    float targetGain, size_t counter, int isPlaying)
{
    const size_t ir = d->desc.ir_size;
    float *hs = d->hrtf_samples;
    memcpy(hs, v->hist, sizeof(v->hist));
    memcpy(hs+HIST, samples, sizeof(float)*n);
    if(isPlaying)
        memcpy(v->hist, hs+n, sizeof(v->hist));

    size_t fademix = 0, outPos = 0;
    if(counter)
    {
        fademix = (n < counter) ? n : counter;
        float gain = targetGain;
        if(counter > fademix)
        {
            const float a = (float)fademix / (float)counter;
            gain = lerpf(v->old_gain, targetGain, a);
        }
        mix_hrtf_blend(hs, d->accum+outPos, ir, (const float(*)[2])v->old_coef, v->old_delay,
            v->old_gain, (const float(*)[2])v->tgt_coef, v->tgt_delay, gain/(float)fademix,
            fademix);
        memcpy(v->old_coef, v->tgt_coef, sizeof(v->old_coef));
        v->old_delay[0] = v->tgt_delay[0]; v->old_delay[1] = v->tgt_delay[1];
        v->old_gain = gain;
        outPos += fademix;
    }
    if(fademix < n)
    {
        const size_t todo = n - fademix;
        float gain = targetGain;
        if(counter > n)
        {
            const float a = (float)todo / (float)(counter-fademix);
            gain = lerpf(v->old_gain, targetGain, a);
        }
        mix_hrtf(hs+fademix, d->accum+outPos, ir, (const float(*)[2])v->tgt_coef, v->tgt_delay,
            v->old_gain, (gain - v->old_gain) / (float)todo, todo);
        v->old_gain = gain;
    }
}

/* Voice::mix for a mono static voice, core/voice.cpp:988-1233 */
static void voice_mix(oracle_device *d, ovoice *v, uint32_t n, b200mix_voice_result *res)
{
    static const float silent[B200MIX_MAX_DRY_CHANNELS];
    const int vstate = v->state;
    const uint32_t increment = v->step;
    if(increment < 1)
    {
        if(vstate == 2) v->state = 0;
        return;
    }
    const int is_queue = !(v->flags & B200MIX_VF_STATIC);
    const obuffer *buf = v->have_buffer ? &d->buffers[is_queue ? v->q_items[v->q_head] : v->buffer] : NULL;
    int looping = (v->flags & B200MIX_VF_LOOPING) != 0;
    if((v->flags & B200MIX_VF_STATIC) && looping && buf)
    {
        if(v->pos >= 0 && (uint32_t)v->pos >= v->loop_end) looping = 0; /* :1015-1019 */
    }

    load_resampled(d, v, vstate, looping, n);

    const size_t counter = v->fading ? (n < 64u ? n : 64u) : 0u;
    const uint32_t cd = d->desc.dry_channels, ns = d->desc.num_sends, cw = d->desc.wet_channels;
    if(!counter)
    {
        /* :1094-1112 */
        if(!(v->flags & B200MIX_VF_HRTF))
            memcpy(v->dry_cur, v->dry_tgt, sizeof(v->dry_cur));
        else
        {
            memcpy(v->old_coef, v->tgt_coef, sizeof(v->old_coef));
            v->old_delay[0] = v->tgt_delay[0]; v->old_delay[1] = v->tgt_delay[1];
            v->old_gain = v->tgt_gain;
        }
        for(uint32_t s = 0;s < ns;++s)
            if(v->send_slot[s] != B200MIX_NO_SLOT)
                memcpy(v->send_cur[s], v->send_tgt[s], sizeof(v->send_cur[s]));
    }

    /* DoMix, :934-984 */
    const float *samples = do_filters(d, v, 0, d->filtered, d->samples, n);
    if(v->flags & B200MIX_VF_HRTF)
    {
        const float targetGain = v->tgt_gain * (float)(vstate == 1);
        do_hrtf_mix(d, v, samples, n, targetGain, counter, vstate == 1);
    }
    else
    {
        const float *tg = (vstate == 1) ? v->dry_tgt : silent;
        mix_samples(samples, n, d->dry, cd, v->dry_cur, tg, counter);
    }
    for(uint32_t s = 0;s < ns;++s)
    {
        if(v->send_slot[s] == B200MIX_NO_SLOT) continue;
        samples = do_filters(d, v, 1 + s, d->filtered, d->samples, n);
        const float *tg = (vstate == 1) ? v->send_tgt[s] : silent;
        mix_samples(samples, n, d->wet + (size_t)v->send_slot[s]*cw, cw, v->send_cur[s], tg,
            counter);
    }

    v->fading = 1;
    if(vstate == 2)
    {
        v->state = 0;
        return;
    }

    /* position update, :1126-1153 */
    uint32_t frac = v->frac + increment*n;
    const uint32_t samplesDone = frac >> FRAC_BITS;
    int32_t pos = add_sat_i32(v->pos, (int32_t)samplesDone);
    frac &= FRAC_MASK;
    if(buf && pos > 0 && is_queue)
    {
        /* streaming source, core/voice.cpp:1183-1196 */
        uint32_t item = v->q_head;
        while(item != B200MIX_NO_LOOP)
        {
            const uint32_t len = d->buffers[v->q_items[item]].frames;
            if(len > (uint32_t)pos) break;
            pos -= (int32_t)len;
            ++v->buffers_done;
            item = queue_next(v, item);
        }
        if(item == B200MIX_NO_LOOP) v->have_buffer = 0;
        else v->q_head = item;
    }
    else if(buf && pos > 0)
    {
        if(looping)
        {
            uint32_t up = (uint32_t)pos;
            if(up >= v->loop_end)
            {
                up = ((up-v->loop_start)%(v->loop_end-v->loop_start)) + v->loop_start;
                pos = (int32_t)up;
            }
        }
        else if((uint32_t)pos >= buf->frames)
            v->have_buffer = 0;
    }
    v->pos = pos; v->frac = frac;
    if(!v->have_buffer)
        v->state = 2; /* Stopping: fade out on the next update, :1224-1232 */
    (void)res;
}

/* BandSplitter::processHfScale(input, output, hfscale), core/filters/splitter.cpp:64-95.
 * NOTE the reference's quirk at :79: lp_z1 = lp_y0 + d0*lp_coeff in this overload. */
static void splitter_hfscale(osplitter *s, const float *in, float *out, size_t n, float hfscale)
{
    const float ap_coeff = s->coeff;
    const float lp_coeff = s->coeff*0.5f + 0.5f;
    float lp_z1 = s->lp_z1, lp_z2 = s->lp_z2, ap_z1 = s->ap_z1;
    for(size_t i = 0;i < n;++i)
    {
        const float x = in[i];
        const float d0 = (x - lp_z1) * lp_coeff;
        const float lp_y0 = lp_z1 + d0;
        lp_z1 = lp_y0 + d0*lp_coeff;
        const float d1 = (lp_y0 - lp_z2) * lp_coeff;
        const float lp_y1 = lp_z2 + d1;
        lp_z2 = lp_y1 + d1;
        const float ap_y = x*ap_coeff + ap_z1;
        ap_z1 = x - ap_y*ap_coeff;
        out[i] = (ap_y-lp_y1)*hfscale + lp_y1;
    }
    s->lp_z1 = lp_z1; s->lp_z2 = lp_z2; s->ap_z1 = ap_z1;
}

/* BandSplitter::process, core/filters/splitter.cpp:28-62 */
static void splitter_process(osplitter *s, const float *in, float *hp, float *lp, size_t n)
{
    const float ap_coeff = s->coeff;
    const float lp_coeff = s->coeff*0.5f + 0.5f;
    float lp_z1 = s->lp_z1, lp_z2 = s->lp_z2, ap_z1 = s->ap_z1;
    for(size_t i = 0;i < n;++i)
    {
        const float x = in[i];
        const float d0 = (x - lp_z1) * lp_coeff;
        const float lp_y0 = lp_z1 + d0;
        lp_z1 = lp_y0 + d0;
        const float d1 = (lp_y0 - lp_z2) * lp_coeff;
        const float lp_y1 = lp_z2 + d1;
        lp_z2 = lp_y1 + d1;
        lp[i] = lp_y1;
        const float ap_y = x*ap_coeff + ap_z1;
        ap_z1 = x - ap_y*ap_coeff;
        hp[i] = ap_y - lp_y1;
    }
    s->lp_z1 = lp_z1; s->lp_z2 = lp_z2; s->ap_z1 = ap_z1;
}

/* MixDirectHrtfBase, core/mixer/hrtfbase.h:91-133 */
static void post_hrtf(oracle_device *d, size_t n)
{
    for(uint32_t c = 0;c < d->dec_channels;++c)
    {
        splitter_hfscale(&d->dec_split[c], d->dry[c], d->temp, n, d->dec_hfscale[c]);
        for(size_t i = 0;i < n;++i)
            apply_coeffs(d->accum+i, d->dec_ir, (const float(*)[2])d->dec_coef[c], d->temp[i],
                d->temp[i]);
    }
    float *left = d->real[d->desc.real_left], *right = d->real[d->desc.real_right];
    for(size_t i = 0;i < n;++i)
    {
        left[i] += d->accum[i][0];
        right[i] += d->accum[i][1];
    }
    memmove(d->accum, d->accum+n, sizeof(float[2])*HRIR);
    memset(d->accum+HRIR, 0, sizeof(float[2])*n);
}

/* BFormatDec::process, core/bformatdec.cpp:60-97 */
static void post_ambidec(oracle_device *d, size_t n)
{
    const uint32_t outs = d->desc.real_channels;
    for(uint32_t c = 0;c < d->amb_in;++c)
    {
        if(d->amb_dual)
        {
            splitter_process(&d->amb_split[c], d->dry[c], d->temp, d->temp2, n);
            float *g = d->amb_hf + (size_t)c*outs;
            float tmp[B200MIX_MAX_DRY_CHANNELS];
            memcpy(tmp, g, sizeof(float)*outs);
            mix_samples(d->temp, n, d->real, outs, tmp, g, 0);
            g = d->amb_lf + (size_t)c*outs;
            memcpy(tmp, g, sizeof(float)*outs);
            mix_samples(d->temp2, n, d->real, outs, tmp, g, 0);
        }
        else
        {
            float *g = d->amb_hf + (size_t)c*outs;
            float tmp[B200MIX_MAX_DRY_CHANNELS];
            memcpy(tmp, g, sizeof(float)*outs);
            mix_samples(d->dry[c], n, d->real, outs, tmp, g, 0);
        }
    }
}

/* BandSplitter::processAllPass, core/filters/splitter.cpp:163-174 */
static void splitter_allpass(float coeff, float *z1p, float *samples, size_t n)
{
    float z1 = *z1p;
    for(size_t i = 0;i < n;++i)
    {
        const float x = samples[i];
        const float y = x*coeff + z1;
        z1 = x - y*coeff;
        samples[i] = y;
    }
    *z1p = z1;
}

/* DeviceBase::Process(StablizerPostProcess), alc/alu.cpp:330-406 */
static void post_stabilizer(oracle_device *d, size_t n)
{
    const uint32_t lidx = d->desc.real_left, ridx = d->desc.real_right, cidx = d->stab_center;
    float *leftout = d->real[lidx], *rightout = d->real[ridx];
    float mid[LINE], side[LINE], tmp[LINE], midhf[LINE], midlf[LINE];
    for(size_t i = 0;i < n;++i) { mid[i] = leftout[i] + rightout[i]; side[i] = leftout[i] - rightout[i]; }
    memset(leftout, 0, sizeof(float)*n); memset(rightout, 0, sizeof(float)*n);

    post_ambidec(d, n);

    for(size_t i = 0;i < n;++i) side[i] += leftout[i] - rightout[i];
    for(size_t i = 0;i < n;++i) tmp[i] = leftout[i] + rightout[i];
    splitter_process(&d->stab_mid, tmp, midhf, midlf, n);

    for(uint32_t i = 0;i < d->desc.real_channels;++i)
    {
        float *buf = (i == lidx) ? mid : (i == ridx) ? side : d->real[i];
        splitter_allpass(d->stab_mid.coeff, &d->stab_ap_z1[i], buf, n);
    }

    const float half_pi = 3.14159265358979323846f*0.5f;
    const float mid_lf = cosf(1.0f/3.0f * half_pi), mid_hf = cosf(1.0f/4.0f * half_pi);
    const float center_lf = sinf(1.0f/3.0f * half_pi), center_hf = sinf(1.0f/4.0f * half_pi);
    float *centerout = d->real[cidx];
    for(size_t i = 0;i < n;++i)
    {
        const float m = midlf[i]*mid_lf + midhf[i]*mid_hf + mid[i];
        const float c = midlf[i]*center_lf + midhf[i]*center_hf;
        const float s = side[i];
        leftout[i] = (m + s) * 0.5f;
        rightout[i] = (m - s) * 0.5f;
        centerout[i] += c * 0.5f;
    }
}

/* init(), core/bs2b.cpp:41-91 */
int oracle_bs2b_coeffs(uint32_t level, uint32_t srate, float out[5])
{
    static const float tab[6][4] = {
        {360.0f,  501.0f, 0.398107170553497f, 0.205671765275719f},
        {500.0f,  711.0f, 0.459726988530872f, 0.228208484414988f},
        {700.0f, 1021.0f, 0.530884444230988f, 0.250105790667544f},
        {360.0f,  494.0f, 0.316227766016838f, 0.168236228897329f},
        {500.0f,  689.0f, 0.354813389233575f, 0.187169483835901f},
        {700.0f,  975.0f, 0.398107170553497f, 0.205671765275719f}};
    if(level < 1 || level > 6 || srate < 1) return B200MIX_ERR_INVALID;
    const float Fc_lo = tab[level-1][0], Fc_hi = tab[level-1][1];
    const float G_lo = tab[level-1][2], G_hi = tab[level-1][3];
    const float pi = 3.14159265358979323846f;
    const float g = 1.0f / (1.0f - G_hi + G_lo);
    float x = expf(-pi*2.0f*Fc_lo/(float)srate);
    out[1] = x;                              /* b1_lo */
    out[0] = G_lo * (1.0f - x) * g;          /* a0_lo */
    x = expf(-pi*2.0f*Fc_hi/(float)srate);
    out[4] = x;                              /* b1_hi */
    out[2] = (1.0f - G_hi * (1.0f - x)) * g; /* a0_hi */
    out[3] = -x * g;                         /* a1_hi */
    return B200MIX_OK;
}

/* bs2b_processor::cross_feed, core/bs2b.cpp:104-163 (the 128-sample blocking only bounds a
 * temporary; the recurrences run straight through).  coef = {a0_lo,b1_lo,a0_hi,a1_hi,b1_hi},
 * hist = history[2]{lo,hi}. */
void oracle_bs2b_cross_feed(const float coef[5], float hist[2][2], float *left, float *right, size_t n)
{
    const float a0lo = coef[0], b1lo = coef[1], a0hi = coef[2], a1hi = coef[3], b1hi = coef[4];
    float lz_lo = hist[0][0], lz_hi = hist[0][1], rz_lo = hist[1][0], rz_hi = hist[1][1];
    for(size_t i = 0;i < n;++i)
    {
        float x = left[i];
        const float l0 = a0hi*x + lz_hi;
        lz_hi = a1hi*x + b1hi*l0;
        const float l1 = a0lo*x + lz_lo;
        lz_lo = b1lo*l1;
        x = right[i];
        const float r0 = a0lo*x + rz_lo;
        rz_lo = b1lo*r0;
        const float r1 = a0hi*x + rz_hi;
        rz_hi = a1hi*x + b1hi*r1;
        left[i] = l0 + r0;
        right[i] = l1 + r1;
    }
    hist[0][0] = lz_lo; hist[0][1] = lz_hi; hist[1][0] = rz_lo; hist[1][1] = rz_hi;
}

/* DeviceBase::Process(Bs2bPostProcess), alc/alu.cpp:408-434 */
static void post_bs2b(oracle_device *d, size_t n)
{
    float *left = d->real[d->desc.real_left], *right = d->real[d->desc.real_right];
    float ldirect[LINE], rdirect[LINE];
    memcpy(ldirect, left, sizeof(float)*n); memcpy(rdirect, right, sizeof(float)*n);
    memset(left, 0, sizeof(float)*n); memset(right, 0, sizeof(float)*n);
    post_ambidec(d, n);
    const float coef[5] = {d->bs2b_a0_lo, d->bs2b_b1_lo, d->bs2b_a0_hi, d->bs2b_a1_hi, d->bs2b_b1_hi};
    oracle_bs2b_cross_feed(coef, d->bs2b_hist, left, right, n);
    for(size_t i = 0;i < n;++i) { left[i] += ldirect[i]; right[i] += rdirect[i]; }
}

/* process(AllPassFilter&...), core/allpass_iir.hpp:53-70 */
static void allpass_process(float st[4][2], const float coeffs[4], const float *src, float *dst,
    size_t n)
{
    for(size_t k = 0;k < n;++k)
    {
        float x = src[k];
        for(int i = 0;i < 4;++i)
        {
            const float y = x*coeffs[i] + st[i][0];
            st[i][0] = st[i][1];
            st[i][1] = y*coeffs[i] - x;
            x = y;
        }
        dst[k] = x;
    }
}

/* The two stereo matrix encoders share their structure (S from W, X [and Z]; D = j(W, X) + Y;
 * Left = S + D, Right = S - D) and differ in constants and in which Dry channels they read:
 * UhjEncoder* (core/uhjfilter.cpp:59-70; Dry 0,1,2 = W,X,Y, alc/alu.cpp:307-311) and TsmeEncoder*
 * (core/tsmefilter.cpp:156-163,289-309; Dry 0,1,2,3 = W,Y,Z,X, alc/alu.cpp:322-327). */
typedef struct { int w, x, y, z; float sw, sx, sz, dw, dx, dy; } oencspec;
static const oencspec kUhjSpec = {0, 1, 2, -1, 0.4698463f, 0.0757602682546f, 0.0f,
    -0.17101005f, 0.208149636675f, 0.267586995182f};
static const oencspec kTsmeSpec = {0, 3, 1, 2, 0.288397341271f, 0.166565447888f, 0.187684284734f,
    0.444008050325f, -0.256439256487f, 0.333238912931f};
static const oencspec *enc_spec(const oracle_device *d)
{ return d->desc.post_process == B200MIX_POST_TSME ? &kTsmeSpec : &kUhjSpec; }

/* UhjEncoderIIR::encode, core/uhjfilter.cpp:231-283 (DeviceBase::Process(UhjPostProcess),
 * alc/alu.cpp:300-312) */
static void post_uhj(oracle_device *d, size_t n)
{
    static const float F1[4] = {0.479400865589f, 0.876218493539f, 0.976597589508f, 0.997499255936f};
    static const float F2[4] = {0.161758498368f, 0.733028932341f, 0.945349700329f, 0.990599156684f};
    const oencspec *e = enc_spec(d);
    const float *w = d->dry[e->w], *x = d->dry[e->x], *y = d->dry[e->y];
    float *left = d->real[d->desc.real_left], *right = d->real[d->desc.real_right];

    for(size_t i = 0;i < n;++i) d->temp[i] = e->sw*w[i] + e->sx*x[i];
    if(e->z >= 0)
    {
        const float *z = d->dry[e->z];
        for(size_t i = 0;i < n;++i) d->temp[i] = d->temp[i] + e->sz*z[i];
    }
    allpass_process(d->uhj_f1wx, F1, d->temp, d->uhj_s+1, n);
    d->uhj_s[0] = d->uhj_delay_wx; d->uhj_delay_wx = d->uhj_s[n];

    for(size_t i = 0;i < n;++i) d->temp[i] = e->dw*w[i] + e->dx*x[i];
    allpass_process(d->uhj_f2wx, F2, d->temp, d->uhj_wx, n);

    allpass_process(d->uhj_f1y, F1, y, d->uhj_d+1, n);
    d->uhj_d[0] = d->uhj_delay_y; d->uhj_delay_y = d->uhj_d[n];
    for(size_t i = 0;i < n;++i) d->uhj_d[i] = d->uhj_wx[i] + e->dy*d->uhj_d[i];

    allpass_process(d->uhj_f1d[0], F1, left, d->uhj_t+1, n);
    d->uhj_t[0] = d->uhj_delay_d[0]; d->uhj_delay_d[0] = d->uhj_t[n];
    for(size_t i = 0;i < n;++i) left[i] = d->uhj_s[i] + d->uhj_d[i] + d->uhj_t[i];

    allpass_process(d->uhj_f1d[1], F1, right, d->uhj_t+1, n);
    d->uhj_t[0] = d->uhj_delay_d[1]; d->uhj_delay_d[1] = d->uhj_t[n];
    for(size_t i = 0;i < n;++i) right[i] = d->uhj_s[i] - d->uhj_d[i] + d->uhj_t[i];
}

/* A FIFO of `len` samples in front of a line: [delay | inout] -> inout, the tail stays behind
 * (the rotate/swap_ranges pairs of core/uhjfilter.cpp:174-193). */
static void fifo_delay(float *dl, size_t len, float *inout, size_t n)
{
    float tmp[LINE];
    if(n >= len)
    {
        memcpy(tmp, inout + (n-len), sizeof(float)*len);
        memmove(inout + len, inout, sizeof(float)*(n-len));
        memcpy(inout, dl, sizeof(float)*len);
        memcpy(dl, tmp, sizeof(float)*len);
    }
    else
    {
        memcpy(tmp, inout, sizeof(float)*n);
        memcpy(inout, dl, sizeof(float)*n);
        memmove(dl, dl + n, sizeof(float)*(len-n));
        memcpy(dl + (len-n), tmp, sizeof(float)*n);
    }
}

/* UhjEncoder<N>::encode, core/uhjfilter.cpp:83-205.  The reference applies the wide-band +90
 * degree shift j() by segmented FFT overlap-add (core/allpass_conv.hpp:42-103): a linear
 * convolution with the N-tap response built at :61-75, delivered one 128-sample segment late.
 * Restated by its definition — the direct convolution, accumulated in double — so it agrees
 * with the reference within float rounding, not bitwise (like the convolution effect).  The
 * other signals are delayed by sFilterDelay = N/2 + 128 to line up with it. */
static void post_uhj_fir(oracle_device *d, size_t n)
{
    const size_t N = d->uhj_fir, seg = 128, delay = N/2 + seg, hist = N + seg - 1;
    float *left = d->real[d->desc.real_left], *right = d->real[d->desc.real_right];
    const oencspec *e = enc_spec(d);
    float w[LINE], x[LINE], y[LINE], z[LINE], ext[512+128+LINE];
    memcpy(w, d->dry[e->w], sizeof(float)*n); memcpy(x, d->dry[e->x], sizeof(float)*n);
    memcpy(y, d->dry[e->y], sizeof(float)*n);
    if(e->z >= 0) memcpy(z, d->dry[e->z], sizeof(float)*n);

    /* j(-0.17101005*W + 0.208149636675*X) of the NON-delayed input (:112-114) */
    memcpy(ext, d->uhj_wxhist, sizeof(float)*hist);
    for(size_t i = 0;i < n;++i) ext[hist + i] = e->dw*w[i] + e->dx*x[i];
    for(size_t t = 0;t < n;++t)
    {
        double acc = 0.0;
        for(size_t k = 1;k < N;k += 2) acc += d->uhj_fir_h[k] * (double)ext[hist + t - seg - k];
        d->uhj_wx[t] = (float)acc;
    }
    memmove(d->uhj_wxhist, ext + n, sizeof(float)*hist);

    fifo_delay(d->uhj_in_delay[0], delay, w, n);
    fifo_delay(d->uhj_in_delay[1], delay, x, n);
    fifo_delay(d->uhj_in_delay[2], delay, y, n);
    if(e->z >= 0) fifo_delay(d->uhj_in_delay[3], delay, z, n);
    fifo_delay(d->uhj_out_delay[0], delay, left, n);
    fifo_delay(d->uhj_out_delay[1], delay, right, n);
    for(size_t i = 0;i < n;++i)
    {
        float S = e->sw*w[i] + e->sx*x[i];
        if(e->z >= 0) S = S + e->sz*z[i];
        const float D = d->uhj_wx[i] + e->dy*y[i];
        left[i] += S + D;
        right[i] += S - D;
    }
}

/* MixSamples(line, Dry, Current, Target, Counter = n): ReverbState::MixOutPlain */
/* mOutTarget of the slot being processed (alc/alu.cpp:626-633): the Dry mix or the target
 * slot's Wet buffer */
static float (*g_out_buf)[LINE];
static size_t g_out_channels;

static void reverb_mix_cb(void *ctx, const float *in, size_t n, float *cur, const float *tgt)
{
    (void)ctx;
    mix_samples(in, n, g_out_buf, g_out_channels, cur, tgt, n);
}

/* ConvolutionState::process + NormalMix (alc/effects/convolution.cpp:623-714,298-304):
 * out_c[i] = sum_k ir_c[k] * x[i-k] over the slot's whole input history, then
 * MixSamples(out_c, Dry, Current, Target, Counter = samplesToDo). */
static void slot_convolution_process(oracle_device *d, oslot *s, const float *in, size_t n)
{
    const uint32_t L = s->frames;
    const uint32_t R = L + LINE;        /* ring: every output of this update still sees L taps */
    /* append the new input to the history ring */
    for(size_t i = 0;i < n;++i)
    {
        s->hist[s->hist_pos] = in[i];
        s->hist_pos = (s->hist_pos+1u) % R;
    }
    for(uint32_t c = 0;c < s->channels;++c)
    {
        const float *h = s->ir + (size_t)c*L;
        for(size_t i = 0;i < n;++i)
        {
            /* newest sample of output i sits (n-1-i) behind the ring head */
            uint32_t p = (s->hist_pos + R - 1u - (uint32_t)(n-1-i)) % R;
            double acc = 0.0;
            for(uint32_t k = 0;k < L;++k)
            {
                acc += (double)h[k] * (double)s->hist[p];
                p = p ? p-1u : R-1u;
            }
            d->temp[i] = (float)acc;
        }
        mix_samples(d->temp, n, g_out_buf, g_out_channels, s->cur + c*B200MIX_MAX_DRY_CHANNELS,
            s->tgt + c*B200MIX_MAX_DRY_CHANNELS, n);
    }
}

/* DeviceBase::renderSamples(unsigned) + ProcessContexts, alc/alu.cpp:2412-2459,2177-2273 */
/* First half of an update: clear, voice loop (alc/alu.cpp:2196-2206).  *wet_host is the
 * oracle's own wet storage [max_slots][wet_channels][1024] (a HOST pointer here). */
int oracle_render_begin(oracle_device *d, uint32_t frames, float **wet_host, size_t *wet_floats)
{
    if(frames < 1 || frames > LINE) return B200MIX_ERR_INVALID;
    const b200mix_device_desc *dd = &d->desc;
    memset(d->dry, 0, sizeof(float[LINE])*dd->dry_channels);
    if(d->real != d->dry) memset(d->real, 0, sizeof(float[LINE])*dd->real_channels);
    memset(d->wet, 0, sizeof(float[LINE])*(size_t)dd->max_slots*dd->wet_channels);

    for(uint32_t i = 0;i < dd->max_voices;++i)
    {
        ovoice *v = &d->voices[i];
        v->buffers_done = 0;
        if(v->state == 1 || v->state == 2)
            voice_mix(d, v, frames, NULL);
    }
    d->mid_frames = frames;
    if(wet_host) *wet_host = &d->wet[0][0];
    if(wet_floats) *wet_floats = (size_t)dd->max_slots*dd->wet_channels*LINE;
    return B200MIX_OK;
}

/* Second half: slot loop and post-process (alc/alu.cpp:2252-2256, 2439-2443). */
int oracle_render_end(oracle_device *d, float *const *real_out, b200mix_voice_result *results,
    const float **real_out_host)
{
    const b200mix_device_desc *dd = &d->desc;
    const uint32_t frames = d->mid_frames;
    if(frames < 1) return B200MIX_ERR_INVALID;
    d->mid_frames = 0;
    /* EffectState::process for every slot (alc/alu.cpp:2252-2256); slots here have no
     * slot targets, so each mixes straight into Dry (mOutTarget, alc/alu.cpp:626-633). */
    /* Slot order (alc/alu.cpp:2211-2251): slots without a target last, in their own order;
     * before them the slots targeting those, and so on — a slot always runs before its target.
     * sorted[] is filled from the back exactly like the reference's partition passes. */
    uint32_t sorted[256], nsl = 0, split;
    {
        uint32_t act[256], na = 0;
        for(uint32_t si = 0;si < dd->max_slots && na < 256;++si) if(d->slots[si].type) act[na++] = si;
        nsl = na;
        /* partition_copy over the REVERSED list: targeted-away slots to the front (in reverse
         * order), the rest to the back in original order */
        uint32_t front = 0, back = na;
        for(uint32_t k = na;k-- > 0;)
        {
            if(d->slots[act[k]].target != B200MIX_NO_SLOT) sorted[front++] = act[k];
            else sorted[--back] = act[k];
        }
        /* partition_copy writes the no-target ones through a reverse iterator while walking the
         * source backwards: they end up in original order */
        split = front;
        uint32_t next_target = na;
        while(split > 1)
        {
            if(next_target == split) break;
            --next_target;
            /* std::partition(begin, split, not_next): elements NOT targeting sorted[next_target]
             * first; the unstable order of std::partition only matters for slots of one level,
             * which are independent of each other except for float summation order into a
             * shared target — we keep the relative order (stable) */
            uint32_t tmp[256], a = 0, b = 0, hold[256];
            for(uint32_t k = 0;k < split;++k)
            {
                if(d->slots[sorted[k]].target != sorted[next_target]) tmp[a++] = sorted[k];
                else hold[b++] = sorted[k];
            }
            for(uint32_t k = 0;k < b;++k) tmp[a+k] = hold[k];
            memcpy(sorted, tmp, sizeof(uint32_t)*split);
            split = a;
        }
    }
    for(uint32_t k = 0;k < nsl;++k)
    {
        const uint32_t si = sorted[k];
        const uint32_t tg = d->slots[si].target;
        if(tg != B200MIX_NO_SLOT)
        { g_out_buf = d->wet + (size_t)tg*dd->wet_channels; g_out_channels = dd->wet_channels; }
        else
        { g_out_buf = d->dry; g_out_channels = dd->dry_channels; }
        if(d->slots[si].type == B200MIX_EFFECT_CONVOLUTION)
            slot_convolution_process(d, &d->slots[si], d->wet[(size_t)si*dd->wet_channels], frames);
        else if(d->slots[si].type == B200MIX_EFFECT_REVERB)
            oreverb_process(d->slots[si].reverb, frames,
                (const float(*)[LINE])d->wet[(size_t)si*dd->wet_channels], dd->wet_channels,
                reverb_mix_cb, d);
        else if(d->slots[si].efx)
            oefx_process(d->slots[si].efx, frames, (const float(*)[LINE])d->wet[(size_t)si*dd->wet_channels],
                dd->wet_channels, g_out_buf, g_out_channels);
    }

    switch(dd->post_process)
    {
    case B200MIX_POST_HRTF: if(d->dec_channels) post_hrtf(d, frames); break;
    case B200MIX_POST_AMBIDEC:
        if(d->amb_in)
        {
            if(d->stab_center != B200MIX_NO_SLOT) post_stabilizer(d, frames);
            else if(d->bs2b_level) post_bs2b(d, frames);
            else post_ambidec(d, frames);
        }
        break;
    case B200MIX_POST_UHJ: case B200MIX_POST_TSME:
        if(dd->dry_channels >= (dd->post_process == B200MIX_POST_TSME ? 4u : 3u))
        { if(d->uhj_fir) post_uhj_fir(d, frames); else post_uhj(d, frames); }
        break;
    case B200MIX_POST_NONE: break;
    default: return B200MIX_ERR_UNSUPPORTED;
    }

    /* if(Limiter) Limiter->process(samplesToDo, RealOut.Buffer), alc/alu.cpp:2446 */
    if(d->limiter) olimiter_process(d->limiter, frames, d->real);

    /* if(ChannelDelays) ApplyDistanceComp(...), alc/alu.cpp:2449-2450 and :2276-2307: a FIFO of
     * `base` samples per channel (the rotate/swap pair), then the gain on what comes out */
    if(d->dc_delay)
        for(uint32_t c = 0;c < dd->real_channels;++c)
        {
            const uint32_t base = d->dc_delay[c];
            if(base < 1) continue;
            float *buf = d->real[c], *dl = d->dc_buf[c], tmp[LINE];
            if(frames >= base)
            {
                memcpy(tmp, buf + (frames-base), sizeof(float)*base);
                memmove(buf + base, buf, sizeof(float)*(frames-base));
                memcpy(buf, dl, sizeof(float)*base);
                memcpy(dl, tmp, sizeof(float)*base);
            }
            else
            {
                memcpy(tmp, buf, sizeof(float)*frames);
                memcpy(buf, dl, sizeof(float)*frames);
                memmove(dl, dl + frames, sizeof(float)*(base-frames));
                memcpy(dl + (base-frames), tmp, sizeof(float)*frames);
            }
            for(uint32_t i = 0;i < frames;++i) buf[i] = buf[i]*d->dc_gain[c];
        }

    if(real_out_host) *real_out_host = &d->real[0][0];
    if(real_out)
        for(uint32_t c = 0;c < dd->real_channels;++c)
            if(real_out[c]) memcpy(real_out[c], d->real[c], sizeof(float)*frames);
    if(results)
        for(uint32_t i = 0;i < dd->max_voices;++i)
        {
            const ovoice *v = &d->voices[i];
            results[i].position = v->pos; results[i].position_frac = v->frac;
            results[i].flags = (v->state == 1) ? B200MIX_VF_PLAYING
                : (v->state == 2) ? B200MIX_VF_STOPPING : B200MIX_VF_STOPPED;
            results[i].buffers_done = d->voices[i].buffers_done;
        }
    return B200MIX_OK;
}

int oracle_render(oracle_device *d, uint32_t frames, float *const *real_out,
    b200mix_voice_result *results)
{
    const int rc = oracle_render_begin(d, frames, NULL, NULL);
    if(rc) return rc;
    return oracle_render_end(d, real_out, results, NULL);
}

/* ApplyDither (alc/alu.cpp:2309-2333) + Write<T> (alc/alu.cpp:2362-2390) after a normal update;
 * the limiter, when installed, has already run in oracle_render_end.  lrintf rounds to nearest even like fastf2i / fast_roundf. */
int oracle_render_interleaved(oracle_device *d, uint32_t frames, void *out, uint32_t out_type,
    uint32_t frame_step, float dither_depth, uint32_t *dither_seed, b200mix_voice_result *results)
{
    const b200mix_device_desc *dd = &d->desc;
    if(!out || out_type > B200MIX_OUT_F32 || frame_step < dd->real_channels) return B200MIX_ERR_INVALID;
    const int rc = oracle_render(d, frames, NULL, results);
    if(rc) return rc;
    if(dither_depth > 0.0f)
    {
        const double invRNGRange = 1.0 / 4294967295.0;
        const float invscale = 1.0f / dither_depth;
        uint32_t seed = *dither_seed;
        for(uint32_t c = 0;c < dd->real_channels;++c)
            for(uint32_t i = 0;i < frames;++i)
            {
                float val = d->real[c][i] * dither_depth;
                seed = seed*96314165u + 907633515u; const uint32_t rng0 = seed;
                seed = seed*96314165u + 907633515u; const uint32_t rng1 = seed;
                val += (float)(rng0*invRNGRange - rng1*invRNGRange);
                d->real[c][i] = (float)lrintf(val) * invscale;
            }
        *dither_seed = seed;
    }
    for(uint32_t i = 0;i < frames;++i)
        for(uint32_t c = 0;c < frame_step;++c)
        {
            const float val = c < dd->real_channels ? d->real[c][i] : 0.0f;
            const size_t idx = (size_t)i*frame_step + c;
            switch(out_type)
            {
            case B200MIX_OUT_I8: case B200MIX_OUT_U8:
            {
                int v = (int)lrintf(fminf(fmaxf(val*128.0f, -128.0f), 127.0f));
                ((uint8_t*)out)[idx] = (uint8_t)(out_type == B200MIX_OUT_U8 ? v + 128 : v);
                break;
            }
            case B200MIX_OUT_I16: case B200MIX_OUT_U16:
            {
                int v = (int)lrintf(fminf(fmaxf(val*32768.0f, -32768.0f), 32767.0f));
                ((uint16_t*)out)[idx] = (uint16_t)(out_type == B200MIX_OUT_U16 ? v + 32768 : v);
                break;
            }
            case B200MIX_OUT_I32: case B200MIX_OUT_U32:
            {
                const int32_t v = (int32_t)lrintf(fminf(fmaxf(val*2147483648.0f, -2147483648.0f), 2147483520.0f));
                ((uint32_t*)out)[idx] = out_type == B200MIX_OUT_U32 ? (uint32_t)v + 2147483648u : (uint32_t)v;
                break;
            }
            default: ((float*)out)[idx] = val; break;
            }
        }
    return B200MIX_OK;
}

int oracle_get_dry(oracle_device *d, float *dry)
{
    memcpy(dry, d->dry, sizeof(float[LINE])*d->desc.dry_channels);
    return B200MIX_OK;
}

int oracle_get_wet(oracle_device *d, uint32_t slot, float *wet)
{
    if(slot >= d->desc.max_slots) return B200MIX_ERR_INVALID;
    memcpy(wet, d->wet[(size_t)slot*d->desc.wet_channels], sizeof(float[LINE])*d->desc.wet_channels);
    return B200MIX_OK;
}

int oracle_get_hrtf_accum(oracle_device *d, float *accum)
{
    memcpy(accum, d->accum, sizeof(d->accum));
    return B200MIX_OK;
}
