// Host-side preparation of a convolution effect's impulse response (no GPU): the rate conversion
// ConvolutionState::deviceUpdate applies when the IR buffer's rate differs from the device's
// (alc/effects/convolution.cpp:356-361,417-425): PPhaseResampler (common/polyphase_resampler.cpp)
// — a Kaiser-windowed sinc (180 dB rejection) run as a polyphase up/down sampler in double —
// restated operation for operation; the result is rounded to float once, as the reference does
// when it stores the filter (:428-431).  tests/test_ir_resample.py pins it bit for bit against
// the reference's class.
#include "../../include/b200mix.h"

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <numeric>
#include <vector>

namespace {

constexpr double kPi = 3.14159265358979323846;

// modified Bessel function of the first kind, order 0, by its power series (:21-45)
double bessel_i0(double x)
{
    const double x2 = x/2.0;
    double term = 1.0, sum = 1.0, last_sum = 0.0;
    int k = 1;
    do {
        const double y = x2 / k;
        ++k;
        last_sum = sum;
        term *= y * y;
        sum += term;
    } while(sum != last_sum);
    return sum;
}

double sinc(double x)
{
    if(std::abs(x) < 1e-9) return 1.0;
    return std::sin(kPi*x) / (kPi*x);
}

double kaiser(double beta, double k, double i0_beta)
{
    if(!(k >= -1.0 && k <= 1.0)) return 0.0;
    return bessel_i0(beta * std::sqrt(1.0 - k*k)) / i0_beta;
}

} // namespace

extern "C" {

int64_t b200mix_resampled_ir_frames(uint32_t src_rate, uint32_t dst_rate, uint32_t frames)
{
    if(!src_rate || !dst_rate) return B200MIX_ERR_INVALID;
    return int64_t((uint64_t(frames)*dst_rate + (src_rate - 1u)) / src_rate);   // convolution.cpp:358-360
}

int b200mix_resample_ir(uint32_t src_rate, uint32_t dst_rate, const float *in, uint32_t in_frames,
    float *out, uint32_t out_frames)
{
    if(!src_rate || !dst_rate || (!in && in_frames) || (!out && out_frames)) return B200MIX_ERR_INVALID;
    if(src_rate == dst_rate)
    {
        // no resampler: the samples are copied (convolution.cpp:424-425)
        for(uint32_t i = 0;i < out_frames;++i) out[i] = i < in_frames ? in[i] : 0.0f;
        return B200MIX_OK;
    }
    // PPhaseResampler::init (:96-121)
    const uint32_t gcd = std::gcd(src_rate, dst_rate);
    const uint32_t P = dst_rate / gcd, Q = src_rate / gcd;
    const double cutoff = (P > Q) ? 0.47 / P : 0.47 / Q;
    const double width = (P > Q) ? 0.03 / P : 0.03 / Q;
    const double rejection = 180.0;
    const double w_t = 2.0 * kPi * width;
    const uint32_t order = uint32_t(std::ceil((rejection - 7.95) / (2.285 * w_t)));
    const uint32_t L = (order + 1u) / 2u;
    const double beta = 0.1102 * (rejection - 8.7);
    const double i0_beta = bessel_i0(beta);
    const size_t M = size_t(L)*2u + 1u;
    std::vector<double> f(M);
    for(size_t i = 0;i < M;++i)
    {
        const double x = double(i) - L;
        f[i] = kaiser(beta, x/L, i0_beta) * 2.0 * double(P) * cutoff * sinc(2.0 * cutoff * x);
    }

    // PPhaseResampler::process (:125-186) on the samples widened to double
    std::vector<double> src(in, in + in_frames);
    const size_t p = P, q = Q, n = in_frames;
    size_t l = L;
    for(uint32_t o = 0;o < out_frames;++o)
    {
        size_t j_s = l / p, j_f = l % p;
        l += q;
        double acc = 0.0;
        if(j_f < M)
        {
            size_t filt_len = (M - j_f - 1)/p + 1;
            if(j_s + 1 > n)
            {
                const size_t skip = std::min(j_s + 1 - n, filt_len);
                j_f += p*skip; j_s -= skip; filt_len -= skip;
            }
            if(filt_len != 0 && j_s + 1 <= n)
            {
                // newest sample first: src[j_s], src[j_s-1], ... paired with f[j_f], f[j_f+p], ...
                const size_t count = std::min(j_s + 1, filt_len);
                for(size_t k = 0;k < count;++k)
                {
                    acc = acc + f[j_f]*src[j_s - k];
                    j_f += p;
                }
            }
        }
        out[o] = float(acc);
    }
    return B200MIX_OK;
}

} // extern "C"
