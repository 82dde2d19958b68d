// pshift_emul.cpp — TEST INFRASTRUCTURE ONLY.  Runs the product's pitch-shifter frame arithmetic
// (openal-soft_b200/csrc/pshift.hpp: the very source the GPU kernel executes one warp per channel)
// on the host with one lane, sequenced the way k_efx_process sequences it, so that
// tests/test_pshift_host.py can hold it against the oracle's independent restatement
// (efx_oracle.cpp, scatter form, straight from alc/effects/pshifter.cpp) without a GPU.
#include <cmath>
#include <cstring>
#include <vector>

#include "../openal-soft_b200/csrc/pshift.hpp"

using namespace b200mix::pshift;

namespace {
struct Emul {
    uint32_t count{0}, pos{kSize - kStep}, shift_i{65536};
    float shift{1.0f};
    std::vector<float> fifo, accum, last, sum, win;
    std::vector<Cplx> tw, X;
    std::vector<float> bins;
};
}

extern "C" {

void *pshift_emul_create(void)
{
    auto *e = new Emul{};
    e->fifo.assign(kMaxLines*kSize, 0.0f); e->accum.assign(kMaxLines*kSize, 0.0f);
    e->last.assign(kBins, 0.0f); e->sum.assign(kBins, 0.0f);
    e->win.resize(kSize); e->tw.resize(kHalf); e->X.resize(kSize); e->bins.resize(2*kBins);
    for(uint32_t k = 0;k < kHalf;++k)
    {   // the tables k_efx_tables builds on the device
        const double a = 3.14159265358979323846 * double(k) / 512.0;
        e->tw[k] = Cplx{std::cos(a), std::sin(a)};
        const double v = std::sin((double(k) + 1.0) * (3.14159265358979323846 / 1025.0));
        e->win[k] = e->win[kSize - 1u - k] = float(v * v);
    }
    return e;
}
void pshift_emul_free(void *p) { delete static_cast<Emul*>(p); }
void pshift_emul_set(void *p, uint32_t shift_i) { auto *e = static_cast<Emul*>(p); e->shift_i = shift_i; e->shift = float(shift_i) * (1.0f/65536.0f); }

// in / out: [channels][1024]; out receives mBBuffer (before the output gains)
void pshift_emul_process(void *p, uint32_t n, uint32_t channels, const float *in, float *out)
{
    auto *e = static_cast<Emul*>(p);
    const Lanes L{0u, 1u};
    float *re = e->bins.data(), *im = re + kBins;
    for(uint32_t base = 0;base < n;)
    {
        const uint32_t todo = std::min(kStep - e->count, n - base);
        for(uint32_t c = 0;c < channels;++c)
            fifo_exchange(e->fifo.data() + c*kSize + e->pos + e->count, in + c*1024u + base, out + c*1024u + base, todo, L);
        e->count += todo; base += todo;
        if(e->count < kStep) break;
        e->count = 0u; e->pos = (e->pos + kStep) & (kSize - 1u);
        for(uint32_t c = 0;c < channels;++c)
        {
            float *fifo = e->fifo.data() + c*kSize, *accum = e->accum.data() + c*kSize;
            analyse_frame(e->X.data(), e->tw.data(), fifo, e->win.data(), e->pos, re, im, L);
            if(c == 0u)
            {
                bins_channel0(re, im, e->last.data(), e->shift, L);
                synthesise_bins<true>(e->X.data(), re, im, e->sum.data(), e->shift_i, L);
            }
            else
            {
                bins_channelN(re, im, e->last.data(), L);
                synthesise_bins<false>(e->X.data(), re, im, e->sum.data(), e->shift_i, L);
            }
            resynthesise_frame(e->X.data(), e->tw.data(), fifo, accum, e->win.data(), e->pos, L);
        }
    }
}

} // extern "C"
