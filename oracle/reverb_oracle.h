/* reverb_oracle.h — TEST INFRASTRUCTURE ONLY (see reverb_oracle.c). */
#ifndef REVERB_ORACLE_H
#define REVERB_ORACLE_H
#include <stddef.h>
#include "../include/b200mix.h"

typedef struct oreverb oreverb;
/* MixSamples(in, Dry, cur, tgt, Counter = n) supplied by the caller */
typedef void (*oreverb_mix_fn)(void *ctx, const float *in, size_t n, float *cur, const float *tgt);

oreverb *oreverb_create(const b200mix_reverb_params *p);
void oreverb_destroy(oreverb *r);
/* ReverbState::update as the mixer sees it; p = post-update values of the then-current pipeline */
void oreverb_update(oreverb *r, const b200mix_reverb_params *p, int full_update);
void oreverb_set_gains(oreverb *r, const float *gains, uint32_t cd);
void oreverb_process(oreverb *r, size_t n, const float (*wet)[B200MIX_LINE_SIZE], uint32_t cw,
    oreverb_mix_fn mix, void *mixctx);
#endif
