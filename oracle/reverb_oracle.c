/* oracle/reverb_oracle.c — TEST INFRASTRUCTURE ONLY.
 * Scalar C restatement of ReverbState::process (alc/effects/reverb.cpp:1813-1882) with its two
 * pipelines and their cross-fade state machine (update :1222-1351 as far as it switches
 * pipelines, ReverbPipeline::clear :550-566) and everything below it: processEarly :1558-1660,
 * processLate :1696-1811, Allpass4::process :1508-1538, VecAllpass::process :1452-1503,
 * VectorPartialScatter :1396-1405, Modulation::calcDelays :1662-1681, DelayLineU
 * :281-365, DualBiquad (core/filters/biquad.cpp:254-283), MixOutPlain :637-656.
 * Parameters are the post-update values of b200mix_reverb_params.  Operation order
 * follows the reference so results are bit-identical to its C build. */
#include "almix_oracle.h"
#include "reverb_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define NL 4
#define MAXUPD 256u                   /* MAX_UPDATE_SAMPLES reverb.cpp:68 */
#define MOD_FRACBITS 24
#define MOD_FRACONE (1u<<MOD_FRACBITS)
#define MOD_FRACMASK (MOD_FRACONE-1u)

/* B2A, reverb.cpp:91-97 */
static const float B2A[NL][NL] = {
    { 0.5f,  0.5f,  0.5f,  0.5f }, { 0.5f, -0.5f, -0.5f,  0.5f },
    { 0.5f,  0.5f, -0.5f, -0.5f }, { 0.5f, -0.5f,  0.5f, -0.5f } };

typedef struct { float b0, b1, b2, a1, a2, z1, z2; } obiquad;

/* ReverbPipeline (reverb.cpp:505-566): everything that exists once per pipeline */
typedef struct opipe {
    b200mix_reverb_params p;
    float *late_in, *early_ap, *early_d, *late_ap, *late_d;  /* [4][len] (late_ap interleaved) */
    obiquad lp[NL], hp[NL], t60hf[NL], t60lf[NL];
    uint32_t early_tap_cur[NL], late_tap_cur[NL];
    float early_coeff_cur;
    uint32_t mod_index;
    size_t fade_samples;              /* mFadeSampleCount */
    struct { float lp_z1, lp_z2, ap_z1; } split[2][NL];   /* mAmbiSplitter (MixOutAmbiUp) */
    float gcur[8][B200MIX_MAX_DRY_CHANNELS], gtgt[8][B200MIX_MAX_DRY_CHANNELS];
} opipe;

enum { ST_DEVICE_CLEAR, ST_START_FADE, ST_FADING, ST_CLEANUP, ST_NORMAL };   /* reverb.cpp:589-595 */

struct oreverb {
    opipe pipes[2];
    int cur, state;                   /* mCurrentPipeline, mPipelineState */
    float *main_d; uint32_t main_len; /* mMainDelay, shared */
    size_t offset;
    float cubic[513];                 /* gCubicTable */
    float temp[NL][MAXUPD];
    float early_out[NL][B200MIX_LINE_SIZE], late_out[NL][B200MIX_LINE_SIZE];
    unsigned moddelays[MAXUPD];
};

static void set_biquad(obiquad *f, const float c[5])   /* coefficients only: z survives */
{ f->b0 = c[0]; f->b1 = c[1]; f->b2 = c[2]; f->a1 = c[3]; f->a2 = c[4]; }

/* ReverbPipeline::clear, reverb.cpp:550-566 (+ EarlyReflections/LateReverb/Modulation::clear) */
static void pipe_clear(opipe *q)
{
    const b200mix_reverb_params *p = &q->p;
    memset(q->late_in, 0, sizeof(float)*NL*p->late_in_len);
    memset(q->early_ap, 0, sizeof(float)*NL*p->early_ap_len);
    memset(q->early_d, 0, sizeof(float)*NL*p->early_len);
    memset(q->late_ap, 0, sizeof(float)*NL*p->late_ap_len);
    memset(q->late_d, 0, sizeof(float)*NL*p->late_len);
    for(int j = 0;j < NL;++j)
    {
        q->lp[j].z1 = q->lp[j].z2 = q->hp[j].z1 = q->hp[j].z2 = 0.0f;
        q->t60hf[j].z1 = q->t60hf[j].z2 = q->t60lf[j].z1 = q->t60lf[j].z2 = 0.0f;
        q->early_tap_cur[j] = 0; q->late_tap_cur[j] = 0;
        q->p.early_tap[j] = 0; q->p.late_tap[j] = 0;
    }
    q->early_coeff_cur = 0.0f; q->p.early_tap_coeff = 0.0f;
    q->mod_index = 0; q->p.mod_step = 1; q->p.mod_depth = 0.0f;
    memset(q->gcur, 0, sizeof(q->gcur)); memset(q->gtgt, 0, sizeof(q->gtgt));
    memset(q->split, 0, sizeof(q->split));
}

/* ReverbState::update as far as the mixer sees it (reverb.cpp:1222-1351): `p` holds the
 * post-update values of the pipeline that is current AFTER the update. */
void oreverb_update(oreverb *r, const b200mix_reverb_params *p, int full_update)
{
    if(full_update || r->state == ST_DEVICE_CLEAR)
    {
        r->state = (r->state != ST_DEVICE_CLEAR) ? ST_START_FADE : ST_NORMAL;
        r->cur = !r->cur;
        r->pipes[!r->cur].p.early_tap_coeff = 0.0f;      /* oldpipeline.mEarlyDelayCoeff[1] = 0 */
    }
    opipe *q = &r->pipes[r->cur];
    const b200mix_reverb_params keep = q->p;
    q->p = *p;
    /* line lengths are fixed by deviceUpdate */
    q->p.main_len = keep.main_len; q->p.late_in_len = keep.late_in_len; q->p.early_ap_len = keep.early_ap_len;
    q->p.early_len = keep.early_len; q->p.late_ap_len = keep.late_ap_len; q->p.late_len = keep.late_len;
    for(int j = 0;j < NL;++j)
    {
        set_biquad(&q->lp[j], p->filter_lp); set_biquad(&q->hp[j], p->filter_hp);
        set_biquad(&q->t60hf[j], p->t60_hf[j]); set_biquad(&q->t60lf[j], p->t60_lf[j]);
    }
    q->fade_samples = p->fade_samples;
}

oreverb *oreverb_create(const b200mix_reverb_params *p)
{
    oreverb *r = calloc(1, sizeof(*r));
    if(!r) return NULL;
    r->main_len = p->main_len;
    r->main_d = calloc((size_t)NL*p->main_len, sizeof(float));
    for(int k = 0;k < 2;++k)
    {
        opipe *q = &r->pipes[k];
        q->p = *p;
        q->late_in = calloc((size_t)NL*p->late_in_len, sizeof(float));
        q->early_ap = calloc((size_t)NL*p->early_ap_len, sizeof(float));
        q->early_d = calloc((size_t)NL*p->early_len, sizeof(float));
        q->late_ap = calloc((size_t)NL*p->late_ap_len, sizeof(float));
        q->late_d = calloc((size_t)NL*p->late_len, sizeof(float));
        pipe_clear(q);
        q->fade_samples = 1;
    }
    oracle_build_cubic_filter(r->cubic);
    r->state = ST_DEVICE_CLEAR; r->cur = 0;
    oreverb_update(r, p, 1);          /* the first update after deviceUpdate is always a full one */
    return r;
}

void oreverb_destroy(oreverb *r)
{
    if(!r) return;
    for(int k = 0;k < 2;++k)
    {
        opipe *q = &r->pipes[k];
        free(q->late_in); free(q->early_ap); free(q->early_d); free(q->late_ap); free(q->late_d);
    }
    free(r->main_d); free(r);
}

/* update3DPanning's result for the CURRENT pipeline: 8 lines x cd target gains */
void oreverb_set_gains(oreverb *r, const float *gains, uint32_t cd)
{
    opipe *q = &r->pipes[r->cur];
    for(int l = 0;l < 8;++l)
        for(uint32_t c = 0;c < cd;++c) q->gtgt[l][c] = gains[l*cd + c];
}

static float lerpf_(float a, float b, float mu) { return a + (b-a)*mu; }

/* BiquadFilter::dualProcess, core/filters/biquad.cpp:254-283 */
static void dual_biquad(obiquad *f0, obiquad *f1, const float *src, float *dst, size_t n)
{
    float z01 = f0->z1, z02 = f0->z2, z11 = f1->z1, z12 = f1->z2;
    for(size_t i = 0;i < n;++i)
    {
        const float x0 = src[i];
        const float y0 = x0*f0->b0 + z01;
        z01 = x0*f0->b1 - y0*f0->a1 + z02;
        z02 = x0*f0->b2 - y0*f0->a2;
        const float x1 = y0;
        const float y1 = x1*f1->b0 + z11;
        z11 = x1*f1->b1 - y1*f1->a1 + z12;
        z12 = x1*f1->b2 - y1*f1->a2;
        dst[i] = y1;
    }
    f0->z1 = z01; f0->z2 = z02; f1->z1 = z11; f1->z2 = z12;
}

/* DelayLineU::write, reverb.cpp:324-338 */
static void line_write(float *buf, size_t len, size_t offset, size_t c, const float *in, size_t n)
{
    float *line = buf + c*len;
    for(size_t i = 0;i < n;++i) line[(offset+i) & (len-1)] = in[i];
}

/* VectorPartialScatter, reverb.cpp:1396-1405 */
static void scatter4(const float in[4], float x, float y, float out[4])
{
    out[0] = x*in[0] + y*(          in[1] + -in[2] + in[3]);
    out[1] = x*in[1] + y*(-in[0]          +  in[2] + in[3]);
    out[2] = x*in[2] + y*( in[0] + -in[1]          + in[3]);
    out[3] = x*in[3] + y*(-in[0] + -in[1] + -in[2]        );
}

/* Allpass4::process, reverb.cpp:1508-1538 */
static void allpass4(opipe *q, float samples[NL][MAXUPD], size_t offset, size_t todo)
{
    const size_t len = q->p.early_ap_len;
    const float c = q->p.early_ap_coeff;
    for(int j = 0;j < NL;++j)
    {
        float *buf = q->early_ap + (size_t)j*len;
        size_t dst = offset, vap = offset - q->p.early_ap_offset[j];
        for(size_t i = 0;i < todo;++i)
        {
            const float x = samples[j][i];
            const float y = buf[(vap++) & (len-1)] - c*x;
            buf[(dst++) & (len-1)] = x + c*y;
            samples[j][i] = y;
        }
    }
}

/* VecAllpass::process, reverb.cpp:1452-1503 (interleaved delay: index*4 + line) */
static void vec_allpass(opipe *q, float samples[NL][MAXUPD], size_t offset, float xc, float yc,
    size_t todo)
{
    const size_t mask = q->p.late_ap_len - 1;
    float *buf = q->late_ap;
    const float fc = q->p.late_ap_coeff;
    for(size_t base = 0;base < todo;)
    {
        size_t vap[NL];
        size_t maxoff;
        for(int c = 0;c < NL;++c) vap[c] = (offset - q->p.late_ap_offset[c]) & mask;
        offset &= mask;
        maxoff = offset;
        for(int c = 0;c < NL;++c) if(vap[c] > maxoff) maxoff = vap[c];
        size_t td = q->p.late_ap_offset[0];
        if(mask+1 - maxoff < td) td = mask+1 - maxoff;
        if(todo - base < td) td = todo - base;
        for(int c = 0;c < NL;++c)
        {
            size_t out_off = vap[c], in_off = offset;
            for(size_t i = 0;i < td;++i)
            {
                const float input = samples[c][base+i];
                const float out = buf[(out_off++)*NL + c] - fc*input;
                buf[(in_off++)*NL + c] = input + fc*out;
                samples[c][base+i] = out;
            }
        }
        for(size_t j = 0;j < td;++j)
        {
            float *d = buf + (offset+j)*NL, f[4];
            scatter4(d, xc, yc, f);
            d[0] = f[0]; d[1] = f[1]; d[2] = f[2]; d[3] = f[3];
        }
        offset += td; base += td;
    }
}

/* ReverbPipeline::processEarly, reverb.cpp:1558-1660 */
static void process_early(oreverb *r, opipe *q, size_t offset, size_t n)
{
    const b200mix_reverb_params *p = &q->p;
    for(size_t base = 0;base < n;)
    {
        const size_t todo = (n-base < MAXUPD) ? n-base : MAXUPD;
        const float fadeStep = 1.0f / (float)todo;
        const float c0 = q->early_coeff_cur, c1 = p->early_tap_coeff;
        q->early_coeff_cur = c1;
        for(int j = 0;j < NL;++j)
        {
            const float *input = r->main_d + (size_t)j*r->main_len;
            size_t t0 = offset - q->early_tap_cur[j], t1 = offset - p->early_tap[j];
            q->early_tap_cur[j] = p->early_tap[j];
            float fadeCount = 0.0f;
            for(size_t i = 0;i < todo;++i)
            {
                const float in0 = input[(t0++) & (r->main_len-1)];
                const float in1 = input[(t1++) & (r->main_len-1)];
                r->temp[j][i] = lerpf_(in0*c0, in1*c1, fadeStep*fadeCount);
                fadeCount += 1.0f;
            }
            dual_biquad(&q->lp[j], &q->hp[j], r->temp[j], r->temp[j], todo);
        }
        allpass4(q, r->temp, offset, todo);

        /* writeReflected, reverb.cpp:340-365 */
        for(size_t i = 0;i < todo;++i)
        {
            const float s0 = r->temp[0][i], s1 = r->temp[1][i], s2 = r->temp[2][i], s3 = r->temp[3][i];
            const size_t o = (offset+i) & (p->early_len-1);
            q->early_d[0*p->early_len + o] = (s0      - s1 - s2 - s3) * 0.5f;
            q->early_d[1*p->early_len + o] = (s1 - s0      - s2 - s3) * 0.5f;
            q->early_d[2*p->early_len + o] = (s2 - s0 - s1      - s3) * 0.5f;
            q->early_d[3*p->early_len + o] = (s3 - s0 - s1 - s2     ) * 0.5f;
        }
        for(int j = 0;j < NL;++j)
        {
            const float *dl = q->early_d + (size_t)j*p->early_len;
            size_t tap = offset - p->early_offset[j];
            for(size_t i = 0;i < todo;++i)
                r->early_out[j][base+i] = dl[(tap++) & (p->early_len-1)]*p->early_coeff + r->temp[j][i];
        }
        /* VectorScatter, reverb.cpp:1408-1423, then feed the late input line */
        for(size_t i = 0;i < todo;++i)
        {
            const float in[4] = {r->temp[0][i], r->temp[1][i], r->temp[2][i], r->temp[3][i]};
            float f[4];
            scatter4(in, p->mix_x, p->mix_y, f);
            r->temp[0][i] = f[0]; r->temp[1][i] = f[1]; r->temp[2][i] = f[2]; r->temp[3][i] = f[3];
        }
        for(int j = 0;j < NL;++j) line_write(q->late_in, p->late_in_len, offset, (size_t)j, r->temp[j], todo);
        base += todo; offset += todo;
    }
}

/* Modulation::calcDelays, reverb.cpp:1662-1681 */
static void calc_delays(oreverb *r, opipe *q, size_t todo)
{
    unsigned idx = q->mod_index;
    const unsigned step = q->p.mod_step;
    const float depth = q->p.mod_depth * 256.0f;
    for(size_t i = 0;i < todo;++i)
    {
        const float x = (float)(idx&MOD_FRACMASK) * (1.0f/MOD_FRACONE);
        const float lfo = !(idx&(MOD_FRACONE>>1))
            ? ((-16.0f * x * x) + (8.0f * x))
            : ((16.0f * x * x) + (-8.0f * x) + (-16.0f * x) + 8.0f);
        idx += step;
        const float v = (lfo+1.0f) * depth;
        r->moddelays[i] = (v > 0.0f) ? (unsigned)v : 0u;   /* float2uint */
    }
    q->mod_index = idx;
}

/* ReverbPipeline::processLate, reverb.cpp:1696-1811 */
static void process_late(oreverb *r, opipe *q, size_t offset, size_t n)
{
    const b200mix_reverb_params *p = &q->p;
    for(size_t base = 0;base < n;)
    {
        size_t todo = p->late_offset[0] < MAXUPD ? p->late_offset[0] : MAXUPD;
        if(n-base < todo) todo = n-base;
        calc_delays(r, q, todo);
        for(int j = 0;j < NL;++j)
        {
            const float *input = q->late_d + (size_t)j*p->late_len;
            const size_t m = p->late_len-1;
            const float midGain = p->t60_mid_gain[j];
            size_t tap = offset - p->late_offset[j];
            for(size_t i = 0;i < todo;++i)
            {
                const unsigned idelay = r->moddelays[i];
                const size_t delay = tap - (idelay>>8);
                const size_t doff = idelay & 255u;
                ++tap;
                const float out0 = input[(delay  ) & m], out1 = input[(delay-1) & m];
                const float out2 = input[(delay-2) & m], out3 = input[(delay-3) & m];
                /* gCubicTable.getCoeff0..3, core/cubic_tables.h:32-39 */
                const float out = out0*r->cubic[256+doff] + out1*r->cubic[doff]
                    + out2*r->cubic[256-doff] + out3*r->cubic[512-doff];
                r->temp[j][i] = out * midGain;
            }
            dual_biquad(&q->t60hf[j], &q->t60lf[j], r->temp[j], r->temp[j], todo);
        }
        const float fadeStep = 1.0f / (float)todo;
        for(int j = 0;j < NL;++j)
        {
            const float *input = q->late_in + (size_t)j*p->late_in_len;
            const size_t m = p->late_in_len-1;
            size_t t0 = offset - q->late_tap_cur[j], t1 = offset - p->late_tap[j];
            q->late_tap_cur[j] = p->late_tap[j];
            const float densityGain = p->density_gain;
            const float densityStep = (t0 != t1) ? densityGain*fadeStep : 0.0f;
            float fadeCount = 0.0f;
            for(size_t i = 0;i < todo;++i)
            {
                const float fade0 = densityGain - densityStep*fadeCount;
                const float fade1 = densityStep*fadeCount;
                fadeCount += 1.0f;
                r->temp[j][i] = input[(t0++) & m]*fade0 + input[(t1++) & m]*fade1 + r->temp[j][i];
            }
        }
        vec_allpass(q, r->temp, offset, p->mix_x, p->mix_y, todo);
        for(int j = 0;j < NL;++j) memcpy(r->late_out[j]+base, r->temp[j], sizeof(float)*todo);
        /* VectorScatterRev, reverb.cpp:1428-1443 */
        for(size_t i = 0;i < todo;++i)
        {
            const float in[4] = {r->temp[3][i], r->temp[2][i], r->temp[1][i], r->temp[0][i]};
            float f[4];
            scatter4(in, p->mix_x, p->mix_y, f);
            r->temp[0][i] = f[0]; r->temp[1][i] = f[1]; r->temp[2][i] = f[2]; r->temp[3][i] = f[3];
        }
        for(int j = 0;j < NL;++j) line_write(q->late_d, p->late_len, offset, (size_t)j, r->temp[j], todo);
        base += todo; offset += todo;
    }
}

/* EarlyA2B / LateA2B, reverb.cpp:102-122 */
#define INV_SQRT2 0.707106781186547524400844362104849039f
static const float EarlyA2B[NL][NL] = {
    { 0.5f,  0.5f,  0.5f,  0.5f }, { 0.5f, -0.5f,  0.5f, -0.5f },
    { 0.5f, -0.5f, -0.5f,  0.5f }, { 0.5f,  0.5f, -0.5f, -0.5f } };
static const float LateA2B[NL][NL] = {
    { 0.5f, 0.5f, 0.5f, 0.5f }, { INV_SQRT2, -INV_SQRT2, 0.0f, 0.0f },
    { 0.0f, 0.0f, -INV_SQRT2, INV_SQRT2 }, { 0.5f, 0.5f, -0.5f, -0.5f } };

/* MixOutAmbiUp's per-row work (reverb.cpp:618-634,658-699): DoMixRow, then the in-place
 * BandSplitter::processHfScale (core/filters/splitter.cpp:99-131), then MixSamples */
static void mix_row_up(oreverb *r, opipe *q, int which, int row, const float a2b[NL],
    float (*in)[B200MIX_LINE_SIZE], size_t n, float *cur, const float *tgt, oreverb_mix_fn mix, void *mixctx)
{
    float tmp[B200MIX_LINE_SIZE];
    for(size_t i = 0;i < n;++i) tmp[i] = 0.0f;
    for(int k = 0;k < NL;++k)
        if(fabsf(a2b[k]) > 0.00001f)                 /* GainSilenceThreshold */
            for(size_t i = 0;i < n;++i) tmp[i] = tmp[i] + in[k][i]*a2b[k];
    const float hfscale = q->p.order_scale[row ? 1 : 0];
    const float ap_coeff = q->p.splitter_coeff, lp_coeff = q->p.splitter_coeff*0.5f + 0.5f;
    float lp_z1 = q->split[which][row].lp_z1, lp_z2 = q->split[which][row].lp_z2;
    float ap_z1 = q->split[which][row].ap_z1;
    for(size_t i = 0;i < n;++i)
    {
        const float in0 = tmp[i];
        const float d0 = (in0 - lp_z1) * lp_coeff;
        const float lp_y0 = lp_z1 + d0;
        lp_z1 = lp_y0 + d0;
        const float d1 = (lp_y0 - lp_z2) * lp_coeff;
        const float lp_y1 = lp_z2 + d1;
        lp_z2 = lp_y1 + d1;
        const float ap_y = in0*ap_coeff + ap_z1;
        ap_z1 = in0 - ap_y*ap_coeff;
        tmp[i] = (ap_y-lp_y1)*hfscale + lp_y1;
    }
    q->split[which][row].lp_z1 = lp_z1; q->split[which][row].lp_z2 = lp_z2;
    q->split[which][row].ap_z1 = ap_z1;
    mix(mixctx, tmp, n, cur, tgt);
    (void)r;
}

/* mixOut -> MixOutPlain / MixOutAmbiUp, reverb.cpp:637-705 */
static void mix_out(oreverb *r, opipe *q, size_t n, oreverb_mix_fn mix, void *mixctx)
{
    if(q->p.upmix)
    {
        for(int j = 0;j < NL;++j)
            mix_row_up(r, q, 0, j, EarlyA2B[j], r->early_out, n, q->gcur[j], q->gtgt[j], mix, mixctx);
        for(int j = 0;j < NL;++j)
            mix_row_up(r, q, 1, j, LateA2B[j], r->late_out, n, q->gcur[4+j], q->gtgt[4+j], mix, mixctx);
        return;
    }
    for(int j = 0;j < NL;++j) mix(mixctx, r->early_out[j], n, q->gcur[j], q->gtgt[j]);
    for(int j = 0;j < NL;++j) mix(mixctx, r->late_out[j], n, q->gcur[4+j], q->gtgt[4+j]);
}

/* ReverbState::process, reverb.cpp:1813-1882 */
void oreverb_process(oreverb *r, size_t n, const float (*wet)[B200MIX_LINE_SIZE], uint32_t cw,
    oreverb_mix_fn mix, void *mixctx)
{
    const size_t offset = r->offset;
    opipe *old = &r->pipes[!r->cur], *q = &r->pipes[r->cur];
    const uint32_t numInput = cw < NL ? cw : NL;
    float tmp[B200MIX_LINE_SIZE];
    for(int c = 0;c < NL;++c)
    {
        for(size_t i = 0;i < n;++i) tmp[i] = 0.0f;
        for(uint32_t k = 0;k < numInput;++k)
            for(size_t i = 0;i < n;++i) tmp[i] = tmp[i] + wet[k][i]*B2A[c][k];
        line_write(r->main_d, r->main_len, offset, (size_t)c, tmp, n);
    }
    if(r->state < ST_FADING) r->state = ST_FADING;

    process_early(r, q, offset, n);
    process_late(r, q, offset, n);
    mix_out(r, q, n, mix, mixctx);

    if(r->state != ST_NORMAL)
    {
        if(r->state == ST_CLEANUP)
        {
            pipe_clear(old);
            r->state = ST_NORMAL;
        }
        else
        {
            if(n >= old->fade_samples)
            {
                memset(old->gtgt, 0, sizeof(old->gtgt));
                old->fade_samples = 0;
                r->state = ST_CLEANUP;
            }
            else
                old->fade_samples -= n;
            process_early(r, old, offset, n);
            process_late(r, old, offset, n);
            mix_out(r, old, n, mix, mixctx);
        }
    }
    r->offset = offset + n;
}
