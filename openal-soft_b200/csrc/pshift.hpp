// pshift.hpp — the pitch shifter's frame arithmetic (PshifterState::process, alc/effects/pshifter.cpp:207-472)
// as lane-strided loops.  On the GPU one warp runs one wet channel of one slot (lane 0..31, stride 32,
// sync = __syncwarp; k_efx_process in efx_kernels.cu sequences the channels); compiled for the host
// with one lane and stride 1 the same source is a serial program, which tests/test_pshift_host.py
// holds against the oracle's independent restatement without a GPU.
//
// Re-formulations against the reference, all exact in real arithmetic:
//  * the real FFT of 1024 floats (pffft, half-complex "ordered" output) is a 1024-point complex FFT
//    in double whose bins 0..512 are rounded to float; the inverse fills the conjugate half;
//  * the scatter of analysis bins k onto synthesis bins j = (k*shift + 0.5) >> 16 (sequential in k,
//    "dominant magnitude wins") is a gather: bin j walks its own contiguous k range in k order.
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define PS_HD __host__ __device__ __forceinline__
#else
#define PS_HD inline
#endif

namespace b200mix {
namespace pshift {

constexpr uint32_t kSize = 1024u, kHalf = 512u, kStep = 128u, kBins = kHalf + 1u;   // StftSize, StftHalfSize, StftStep
constexpr uint32_t kMaxLines = 9u;                      // NumLines: second-order ambisonics
constexpr float kPi = 3.14159265358979323846f, kInvPi = 0.318309886183790671538f;
constexpr float kExpectedCycles = kPi*2.0f / 8.0f;      // per-hop phase advance of one bin, OversampleFactor 8
constexpr float kScale = 3.0f / 8.0f / 1024.0f;         // pshifter.cpp:421

struct alignas(16) Cplx { double x, y; };

struct Lanes {
    uint32_t lane, stride;
    PS_HD void sync() const
    {
#if defined(__CUDA_ARCH__)
        __syncwarp();
#endif
    }
};

// float2int (common/alnumeric.h:195-222): truncation
PS_HD int f2i(float f)
{
    if(!(f < 2147483648.0f)) return 2147483647;
    if(!(f > -2147483648.0f)) return -2147483647 - 1;
    return static_cast<int>(f);
}
// wrap a phase expressed in units of pi into [-1, +1] (pshifter.cpp:290-292, :348-350, :391-393)
PS_HD float wrap_pi_units(float tmp)
{
    const int qpd = f2i(tmp);
    return tmp - static_cast<float>(qpd + (qpd%2));
}

// In-place radix-2 decimation-in-time FFT of 1024 complex doubles; tw[k] = exp(+i 2 pi k / 1024),
// k < 512; sign -1: forward (exp(-i..)), +1: backward, unscaled.
PS_HD void fft1024(Cplx *x, const Cplx *tw, const Lanes L, const double sign)
{
    for(uint32_t i = L.lane;i < kSize;i += L.stride)
    {
        uint32_t j = 0u;
        for(uint32_t b = 0;b < 10u;++b) j |= ((i >> b) & 1u) << (9u - b);
        if(i < j) { const Cplx a = x[i]; x[i] = x[j]; x[j] = a; }
    }
    L.sync();
    for(uint32_t s = 0;s < 10u;++s)
    {
        const uint32_t half = 1u << s;
        for(uint32_t b = L.lane;b < kHalf;b += L.stride)
        {
            const uint32_t j = b & (half - 1u), k = ((b >> s) << (s + 1u)) + j;
            const Cplx w = tw[j * (kHalf >> s)];
            const double wi = w.y * sign;
            const Cplx v = x[k + half];
            const Cplx t{v.x*w.x - v.y*wi, v.x*wi + v.y*w.x};
            const Cplx u = x[k];
            x[k + half] = Cplx{u.x - t.x, u.y - t.y};
            x[k] = Cplx{u.x + t.x, u.y + t.y};
        }
        L.sync();
    }
}

// FIFO step (pshifter.cpp:222-231): what the FIFO held comes out, the new input goes in.
PS_HD void fifo_exchange(float *fifo_at, const float *in, float *out, const uint32_t todo, const Lanes L)
{
    for(uint32_t i = L.lane;i < todo;i += L.stride)
    {
        out[i] = fifo_at[i];
        fifo_at[i] = in[i];
    }
    L.sync();
}

// Window the FIFO from `pos` on and transform (pshifter.cpp:246-255); mag / val receive the float
// spectrum of bins 0..512 (real, imaginary; bins 0 and 512 are real).
PS_HD void analyse_frame(Cplx *X, const Cplx *tw, const float *fifo, const float *win, const uint32_t pos,
    float *re, float *im, const Lanes L)
{
    for(uint32_t k = L.lane;k < kSize;k += L.stride)
        X[k] = Cplx{static_cast<double>(fifo[(pos + k) & (kSize - 1u)] * win[k]), 0.0};
    L.sync();
    fft1024(X, tw, L, -1.0);
    for(uint32_t k = L.lane;k < kBins;k += L.stride)
    {
        re[k] = static_cast<float>(X[k].x);
        im[k] = (k == 0u || k == kHalf) ? 0.0f : static_cast<float>(X[k].y);
    }
    L.sync();
}

// Channel 0, per analysis bin (pshifter.cpp:264-316): magnitude and the bin's true frequency (scaled
// by the shift) replace re / im in place; mLastPhase takes the new phase.
PS_HD void bins_channel0(float *mag_re, float *val_im, float *last_phase, const float shift, const Lanes L)
{
    for(uint32_t k = L.lane;k < kBins;k += L.stride)
    {
        const float re = mag_re[k], im = val_im[k];
        const float magnitude = ::hypotf(re, im);
        const float phase = ::atan2f(im, re);
        const float bin_offset = static_cast<float>(k & 7u);
        float tmp = (phase - last_phase[k]) - bin_offset*kExpectedCycles;
        last_phase[k] = phase;
        tmp *= kInvPi;
        tmp = wrap_pi_units(tmp);
        tmp *= 0.5f*8.0f;
        const float freqbin = static_cast<float>(k) + tmp;
        mag_re[k] = magnitude;
        val_im[k] = freqbin * shift;
    }
    L.sync();
}
// The other channels (pshifter.cpp:372-391): magnitude and the phase difference from channel 0.
PS_HD void bins_channelN(float *mag_re, float *val_im, const float *last_phase, const Lanes L)
{
    for(uint32_t k = L.lane;k < kBins;k += L.stride)
    {
        const float re = mag_re[k], im = val_im[k];
        mag_re[k] = ::hypotf(re, im);
        val_im[k] = ::atan2f(im, re) - last_phase[k];
    }
    L.sync();
}

// first analysis bin k with (k*shift_i + 0.5) >> 16 >= j
PS_HD uint32_t first_bin(const uint32_t j, const uint32_t shift_i)
{
    if(j == 0u) return 0u;
    const uint32_t a = (j << 16) - 32768u;
    return (a + shift_i - 1u) / shift_i;
}

// Synthesis bin j gathers its analysis bins in k order (pshifter.cpp:305-315 / :384-390), turns the
// result into a phase (channel 0: accumulates mSumPhase, :341-353; others: offset from it, :393-401)
// and writes the bin and its conjugate mirror into the transform buffer (:355-363 / :403-411).
template<bool Channel0>
PS_HD void synthesise_bins(Cplx *X, const float *mag, const float *val, float *sum_phase, const uint32_t shift_i,
    const Lanes L)
{
    for(uint32_t j = L.lane;j < kBins;j += L.stride)
    {
        uint32_t k = first_bin(j, shift_i);
        uint32_t kend = first_bin(j + 1u, shift_i);
        if(kend > kBins) kend = kBins;
        float smag = 0.0f, sval = 0.0f;
        for(;k < kend;++k)
        {
            const float m = mag[k];
            if(smag < m) sval = val[k];
            smag += m;
        }
        float phase;
        if(Channel0)
        {
            const float bin_offset = static_cast<float>(j & ~7u);
            float tmp = (sval - bin_offset) * kExpectedCycles;
            tmp = (tmp + sum_phase[j]) * kInvPi;
            tmp = wrap_pi_units(tmp);
            phase = tmp * kPi;
            sum_phase[j] = phase;
        }
        else
        {
            float tmp = sum_phase[j] + sval;
            tmp *= kInvPi;
            tmp = wrap_pi_units(tmp);
            phase = tmp * kPi;
        }
        const float re = smag * ::cosf(phase);                     // std::polar
        const float im = smag * ::sinf(phase);
        if(j == 0u || j == kHalf) X[j] = Cplx{static_cast<double>(re), 0.0};
        else
        {
            X[j] = Cplx{static_cast<double>(re), static_cast<double>(im)};
            X[kSize - j] = Cplx{static_cast<double>(re), -static_cast<double>(im)};
        }
    }
    L.sync();
}

// Back to time, window, overlap-add at `pos`, and the finished hop moves into the FIFO slot that
// was just consumed (pshifter.cpp:413-437).
PS_HD void resynthesise_frame(Cplx *X, const Cplx *tw, float *fifo, float *accum, const float *win, const uint32_t pos,
    const Lanes L)
{
    fft1024(X, tw, L, 1.0);
    for(uint32_t k = L.lane;k < kSize;k += L.stride)
    {
        const float v = win[k]*static_cast<float>(X[k].x)*kScale;
        const uint32_t q = (pos + k) & (kSize - 1u);
        accum[q] = accum[q] + v;
    }
    L.sync();
    for(uint32_t k = L.lane;k < kStep;k += L.stride)
    {
        fifo[pos + k] = accum[pos + k];
        accum[pos + k] = 0.0f;
    }
    L.sync();
}

} // namespace pshift
} // namespace b200mix
