/* TEST INFRASTRUCTURE — CPU restatement of the reference's output limiter
 * (Compressor, core/mastering.h / core/mastering.cpp).  Only tests/, bench.py's baseline
 * legs and __graft_entry__.smoke() may use it. */
#ifndef LIMITER_ORACLE_H
#define LIMITER_ORACLE_H

#include <stddef.h>
#include <stdint.h>
#include "../include/b200mix.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct olimiter olimiter;

/* Compressor::Create (core/mastering.cpp:108-166) */
olimiter *olimiter_create(const b200mix_limiter_desc *desc, uint32_t num_chans, float sample_rate);
void olimiter_destroy(olimiter *l);
uint32_t olimiter_look_ahead(const olimiter *l);
/* Compressor::process (core/mastering.cpp:261-379) on num_chans lines of 1024 floats */
void olimiter_process(olimiter *l, uint32_t samples, float (*inout)[1024]);

#ifdef __cplusplus
}
#endif
#endif
