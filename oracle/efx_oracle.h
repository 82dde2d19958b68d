/* efx_oracle.h — CPU restatement of the EFX effects' process() (TEST INFRASTRUCTURE ONLY; see
 * efx_oracle.cpp).  C interface for almix_oracle.c. */
#ifndef EFX_ORACLE_H
#define EFX_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#include "../include/b200mix.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct oefx oefx;
/* deviceUpdate + update; *rc receives B200MIX_OK / an error */
oefx *oefx_create(const b200mix_efx_props *props, const b200mix_efx_target *target, int *rc);
/* update() on an existing state (same type); B200MIX_ERR_INVALID if the state must be re-created */
int oefx_update(oefx *e, const b200mix_efx_props *props, const b200mix_efx_target *target);
uint32_t oefx_type(const oefx *e);
/* process(): in = the slot's wet lines [nin][1024], out = the target mix [nout][1024] (accumulated) */
void oefx_process(oefx *e, size_t n, const float (*in)[1024], size_t nin, float (*out)[1024], size_t nout);
void oefx_free(oefx *e);
#ifdef __cplusplus
}
#endif
#endif
