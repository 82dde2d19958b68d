/* TEST INFRASTRUCTURE — CPU restatement of the reference's output limiter.
 *
 * Follows core/mastering.cpp: SlidingHold (:23-105), Compressor::Create (:108-166),
 * Compressor::gainCompressor (:177-259), Compressor::process (:261-379).  Pinned against the
 * compiled reference by tests/test_oracle_golden.py (limiter fixtures rendered by the
 * unmodified library) and tests/test_oracle_vs_ref.py. */
#include "limiter_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define LINE 1024u
#define MASK (LINE-1u)

typedef struct {
    float values[LINE];
    uint32_t expiries[LINE];
    uint32_t lower, upper, length;
} ohold;

struct olimiter {
    uint32_t auto_knee, auto_attack, auto_release, auto_postgain, auto_declip;
    uint32_t look_ahead, num_chans;
    float pre_gain, post_gain, threshold, slope, knee, attack, release;
    float side_chain[LINE*2], crest[LINE];
    ohold *hold;
    float (*delay)[LINE];
    float crest_coeff, gain_estimate, adapt_coeff;
    float last_peak_sq, last_rms_sq, last_release, last_attack, last_gain_dev;
};

static float lerpf(float a, float b, float mu) { return a + (b-a)*mu; }   /* altypes.hpp:1197 */
static float maxf(float a, float b) { return (a < b) ? b : a; }             /* std::max */
static float clampf(float v, float lo, float hi) { return (v < lo) ? lo : ((hi < v) ? hi : v); }

/* UpdateSlidingHold, core/mastering.cpp:46-89 */
static float hold_update(ohold *h, uint32_t i, float in)
{
    uint32_t lower = h->lower, upper = h->upper;
    if(i >= h->expiries[upper])
        upper = (upper + 1u) & MASK;
    if(in >= h->values[upper])
    {
        h->values[upper] = in;
        h->expiries[upper] = i + h->length;
        lower = upper;
    }
    else
    {
        for(;;)
        {
            int found = 0;
            do {
                if(!(in >= h->values[lower])) { found = 1; break; }
            } while(lower--);
            if(found) break;
            lower = MASK;
        }
        lower = (lower + 1u) & MASK;
        h->values[lower] = in;
        h->expiries[lower] = i + h->length;
    }
    h->lower = lower; h->upper = upper;
    return h->values[upper];
}

/* ShiftSlidingHold, core/mastering.cpp:91-105 */
static void hold_shift(ohold *h, uint32_t n)
{
    if(h->lower < h->upper)
    {
        for(uint32_t k = 0;k <= h->lower;++k) h->expiries[k] -= n;
        for(uint32_t k = h->upper;k < LINE;++k) h->expiries[k] -= n;
    }
    else
        for(uint32_t k = h->upper;k <= h->lower;++k) h->expiries[k] -= n;
}

olimiter *olimiter_create(const b200mix_limiter_desc *p, uint32_t num_chans, float sample_rate)
{
    olimiter *c = calloc(1, sizeof(*c));
    if(!c) return NULL;
    const uint32_t look_ahead = (uint32_t)clampf(roundf(p->look_ahead_time*sample_rate), 0.0f, (float)LINE-1.0f);
    const uint32_t hold = (uint32_t)clampf(roundf(p->hold_time*sample_rate), 0.0f, (float)LINE-1.0f);
    c->auto_knee = !!(p->auto_flags & B200MIX_LIM_AUTO_KNEE);
    c->auto_attack = !!(p->auto_flags & B200MIX_LIM_AUTO_ATTACK);
    c->auto_release = !!(p->auto_flags & B200MIX_LIM_AUTO_RELEASE);
    c->auto_postgain = !!(p->auto_flags & B200MIX_LIM_AUTO_POSTGAIN);
    c->auto_declip = c->auto_postgain && (p->auto_flags & B200MIX_LIM_AUTO_DECLIP);
    c->look_ahead = look_ahead;
    c->num_chans = num_chans;
    c->pre_gain = powf(10.0f, p->pre_gain_db / 20.0f);
    c->post_gain = (float)(log(10.0)/20.0 * (double)p->post_gain_db);
    c->threshold = (float)(log(10.0)/20.0 * (double)p->threshold_db);
    c->slope = 1.0f/maxf(1.0f, p->ratio) - 1.0f;
    {
        const double k = log(10.0)/20.0 * (double)p->knee_db;
        c->knee = (float)((0.0 < k) ? k : 0.0);
    }
    c->attack = maxf(1.0f, p->attack_time * sample_rate);
    c->release = maxf(1.0f, p->release_time * sample_rate);
    if(c->auto_knee) c->slope = -1.0f;
    if(look_ahead > 0)
    {
        if(hold > 1)
        {
            c->hold = calloc(1, sizeof(ohold));
            c->hold->values[0] = -INFINITY;
            c->hold->expiries[0] = hold;
            c->hold->length = hold;
        }
        c->delay = calloc(num_chans ? num_chans : 1, sizeof(float[LINE]));
    }
    c->crest_coeff = expf(-1.0f / (0.200f * sample_rate));
    c->gain_estimate = c->threshold * -0.5f * c->slope;
    c->adapt_coeff = expf(-1.0f / (2.0f * sample_rate));
    return c;
}

void olimiter_destroy(olimiter *l)
{
    if(!l) return;
    free(l->hold); free(l->delay); free(l);
}

uint32_t olimiter_look_ahead(const olimiter *l) { return l->look_ahead; }

/* Compressor::gainCompressor, core/mastering.cpp:177-259 */
static void gain_compressor(olimiter *c, uint32_t n)
{
    const float threshold = c->threshold, slope = c->slope, attack = c->attack, release = c->release;
    const float c_est = c->gain_estimate, a_adp = c->adapt_coeff;
    float post_gain = c->post_gain, knee = c->knee;
    float t_att = attack, t_rel = release - attack;
    float a_att = expf(-1.0f / t_att), a_rel = expf(-1.0f / t_rel);
    float y_1 = c->last_release, y_L = c->last_attack, c_dev = c->last_gain_dev;

    for(uint32_t i = 0;i < n;++i)
    {
        const float input = c->side_chain[i];
        const float look = c->side_chain[c->look_ahead + i];
        if(c->auto_knee) knee = maxf(0.0f, 2.5f*(c_dev + c_est));
        const float knee_h = 0.5f * knee;

        const float x_over = look - threshold;
        const float y_G = (x_over <= -knee_h) ? 0.0f
            : (fabsf(x_over) < knee_h) ? (x_over+knee_h) * (x_over+knee_h) / (2.0f * knee)
            : x_over;

        const float y2_crest = c->crest[i];
        if(c->auto_attack)
        {
            t_att = 2.0f*attack/y2_crest;
            a_att = expf(-1.0f / t_att);
        }
        if(c->auto_release)
        {
            t_rel = 2.0f*release/y2_crest - t_att;
            a_rel = expf(-1.0f / t_rel);
        }

        const float x_L = -slope * y_G;
        y_1 = maxf(x_L, lerpf(x_L, y_1, a_rel));
        y_L = lerpf(y_1, y_L, a_att);

        c_dev = lerpf(-(y_L+c_est), c_dev, a_adp);
        if(c->auto_postgain)
        {
            if(c->auto_declip) c_dev = maxf(c_dev, input - y_L - threshold - c_est);
            post_gain = -(c_dev + c_est);
        }
        c->side_chain[i] = expf(post_gain - y_L);
    }
    c->last_release = y_1; c->last_attack = y_L; c->last_gain_dev = c_dev;
}

/* Compressor::process, core/mastering.cpp:261-379 */
void olimiter_process(olimiter *c, uint32_t n, float (*inout)[1024])
{
    if(c->pre_gain != 1.0f)
        for(uint32_t ch = 0;ch < c->num_chans;++ch)
            for(uint32_t i = 0;i < n;++i) inout[ch][i] *= c->pre_gain;

    float *side = c->side_chain + c->look_ahead;
    for(uint32_t i = 0;i < n;++i) side[i] = 0.0f;
    for(uint32_t ch = 0;ch < c->num_chans;++ch)
        for(uint32_t i = 0;i < n;++i) side[i] = maxf(side[i], fabsf(inout[ch][i]));

    if(c->auto_attack || c->auto_release)
    {
        const float a_crest = c->crest_coeff;
        float y2_peak = c->last_peak_sq, y2_rms = c->last_rms_sq;
        for(uint32_t i = 0;i < n;++i)
        {
            const float x2 = clampf(side[i]*side[i], 0.000001f, 1000000.0f);
            y2_peak = maxf(x2, lerpf(x2, y2_peak, a_crest));
            y2_rms = lerpf(x2, y2_rms, a_crest);
            c->crest[i] = y2_peak / y2_rms;
        }
        c->last_peak_sq = y2_peak; c->last_rms_sq = y2_rms;
    }

    if(c->hold)
    {
        for(uint32_t i = 0;i < n;++i)
            side[i] = hold_update(c->hold, i, logf(maxf(0.000001f, side[i])));
        hold_shift(c->hold, n);
    }
    else
        for(uint32_t i = 0;i < n;++i) side[i] = logf(maxf(0.000001f, side[i]));

    gain_compressor(c, n);

    if(c->delay)
    {
        /* the rotate/swap dance of :331-358 is a FIFO of look_ahead samples per channel */
        const uint32_t la = c->look_ahead;
        float tmp[LINE];
        for(uint32_t ch = 0;ch < c->num_chans;++ch)
        {
            float *buf = inout[ch], *dl = c->delay[ch];
            if(n >= la)
            {
                memcpy(tmp, buf + (n-la), sizeof(float)*la);          /* the newest la inputs */
                memmove(buf + la, buf, sizeof(float)*(n-la));
                memcpy(buf, dl, sizeof(float)*la);
                memcpy(dl, tmp, sizeof(float)*la);
            }
            else
            {
                memcpy(tmp, buf, sizeof(float)*n);
                memcpy(buf, dl, sizeof(float)*n);
                memmove(dl, dl + n, sizeof(float)*(la-n));
                memcpy(dl + (la-n), tmp, sizeof(float)*n);
            }
        }
    }

    for(uint32_t ch = 0;ch < c->num_chans;++ch)
        for(uint32_t i = 0;i < n;++i) inout[ch][i] = c->side_chain[i] * inout[ch][i];

    memmove(c->side_chain, c->side_chain + n, sizeof(float)*c->look_ahead);
}
