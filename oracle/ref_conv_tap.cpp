/* ref_conv_tap.cpp — TEST INFRASTRUCTURE ONLY.
 * ConvolutionState lives in an anonymous namespace inside alc/effects/convolution.cpp; to read the
 * output gains ConvolutionState::update computed (without patching the reference) that file is
 * compiled into this translation unit and the live EffectState is cast.  Nothing is copied: the
 * include below reads the reference source where it lies. */
#include "config.h"

#include <algorithm>
#include <array>
#include <cstddef>
#include <cstdint>
#include <cstring>

#include "alc/effects/convolution.cpp"

#include "AL/alc.h"
#include "alc/context.hpp"

extern "C" {

/* gains[c][MaxAmbiChannels] = ConvolutionState::mChans[c].Target of the convolution effect on
 * active slot idx; returns the channel (IR line) count or < 0. */
int refh_conv_gains(ALCcontext *actx, int idx, float *gains)
{
    auto *ctx = static_cast<al::Context*>(actx);
    auto *arr = ctx->mActiveAuxSlots.load(std::memory_order_acquire);
    if(!arr || idx < 0 || size_t(idx) >= (arr->size()>>1)) return -1;
    auto *slot = (*arr)[size_t(idx)];
    if(slot->EffectType != EffectSlotType::Convolution) return -2;
    auto *st = static_cast<ConvolutionState*>(slot->mEffectState.get());
    for(size_t c{0};c < st->mChans.size();++c)
        std::copy(st->mChans[c].Target.begin(), st->mChans[c].Target.end(), gains + c*MaxAmbiChannels);
    return static_cast<int>(st->mChans.size());
}

} // extern "C"
