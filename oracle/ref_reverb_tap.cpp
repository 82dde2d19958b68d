/* ref_reverb_tap.cpp — TEST INFRASTRUCTURE ONLY.
 * ReverbState lives in an anonymous namespace inside alc/effects/reverb.cpp, so the only
 * way to read the parameters ReverbState::update computed (without patching the reference)
 * is to compile that file into this translation unit and cast the live EffectState.
 * Nothing is copied: the include below reads the reference source where it lies.  The
 * duplicate ReverbStateFactory symbols stay local to this .so. */
#include "config.h"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <functional>
#include <numeric>
#include <span>
#include <variant>
#include <vector>

#include "opthelpers.h"
#define class struct
#define private public
#include "core/filters/splitter.h"
#undef private
#undef class
#define protected public
#include "core/filters/biquad.h"
#undef protected

#include "alc/effects/reverb.cpp"

#include "AL/alc.h"
#include "alc/context.hpp"
#include "../include/b200mix.h"

namespace {
void put_biquad(const BiquadFilter &f, float *out)
{
    out[0] = f.mCoeffs.mB0; out[1] = f.mCoeffs.mB1; out[2] = f.mCoeffs.mB2;
    out[3] = f.mCoeffs.mA1; out[4] = f.mCoeffs.mA2;
}
} // namespace

extern "C" {

/* Fills `out` from the CURRENT pipeline of the ReverbState on active slot `idx`;
 * gains is [8][dry_channels]: early lines 0-3 then late lines 0-3 (Target gains).
 * Returns 0, or <0 if the slot holds no reverb. */
int refh_reverb_params(ALCcontext *actx, int idx, b200mix_reverb_params *out, float *gains,
    int *pipeline_state)
{
    auto *ctx = static_cast<al::Context*>(actx);
    auto *arr = ctx->mActiveAuxSlots.load(std::memory_order_acquire);
    if(!arr || idx < 0 || size_t(idx) >= (arr->size()>>1)) return -1;
    auto *slot = (*arr)[size_t(idx)];
    if(slot->EffectType != EffectSlotType::Reverb) return -2;
    auto *st = static_cast<ReverbState*>(slot->mEffectState.get());
    auto &p = st->mPipelines[st->mCurrentPipeline];
    /* low byte: mPipelineState; bit 8: mCurrentPipeline (flips on every full update) */
    *pipeline_state = int(st->mPipelineState) | (int(st->mCurrentPipeline) << 8);
    auto *dev = static_cast<DeviceBase*>(ctx->mDevice);
    const auto cd = dev->Dry.Buffer.size();

    std::memset(out, 0, sizeof(*out));
    out->struct_size = sizeof(*out);
    out->main_len = uint32_t(st->mMainDelay.mLine.size()/NUM_LINES);
    out->late_in_len = uint32_t(p.mLateDelayIn.mLine.size()/NUM_LINES);
    out->early_ap_len = uint32_t(p.mEarly.Allpass.Delay.mLine.size()/NUM_LINES);
    out->early_len = uint32_t(p.mEarly.Delay.mLine.size()/NUM_LINES);
    out->late_ap_len = uint32_t(p.mLate.VecAp.Delay.mLine.size()/NUM_LINES);
    out->late_len = uint32_t(p.mLate.Delay.mLine.size()/NUM_LINES);
    for(auto j = 0_uz;j < NUM_LINES;++j)
    {
        out->early_tap[j] = uint32_t(p.mEarlyDelayTap[j][1]);
        out->late_tap[j] = uint32_t(p.mLateDelayTap[j][1]);
        out->early_ap_offset[j] = uint32_t(p.mEarly.Allpass.Offset[j]);
        out->early_offset[j] = uint32_t(p.mEarly.Offset[j]);
        out->late_offset[j] = uint32_t(p.mLate.Offset[j]);
        out->late_ap_offset[j] = uint32_t(p.mLate.VecAp.Offset[j]);
        out->t60_mid_gain[j] = p.mLate.T60[j].mMidGain;
        put_biquad(p.mLate.T60[j].mHFFilter, out->t60_hf[j]);
        put_biquad(p.mLate.T60[j].mLFFilter, out->t60_lf[j]);
        for(auto c = 0_uz;c < cd;++c)
        {
            gains[j*cd + c] = p.mEarly.Gains[j].Target[c];
            gains[(4+j)*cd + c] = p.mLate.Gains[j].Target[c];
        }
    }
    out->early_tap_coeff = p.mEarlyDelayCoeff[1];
    out->mix_x = p.mMixX; out->mix_y = p.mMixY;
    put_biquad(p.mFilter[0].Lp, out->filter_lp);
    put_biquad(p.mFilter[0].Hp, out->filter_hp);
    out->early_ap_coeff = p.mEarly.Allpass.Coeff;
    out->early_coeff = p.mEarly.Coeff;
    out->density_gain = p.mLate.DensityGain;
    out->mod_step = p.mLate.Mod.Step;
    out->mod_depth = p.mLate.Mod.Depth;
    out->late_ap_coeff = p.mLate.VecAp.Coeff;
    out->fade_samples = uint32_t(p.mFadeSampleCount);
    out->upmix = st->mUpmixOutput ? 1u : 0u;
    out->order_scale[0] = st->mOrderScales[0];
    out->order_scale[1] = st->mOrderScales[1];
    out->splitter_coeff = p.mAmbiSplitter[0][0].mCoeff;
    return 0;
}

} // extern "C"
