// resampler_tables.cpp — see resampler_tables.hpp.
#include "resampler_tables.hpp"

#include <algorithm>
#include <cmath>
#include <numbers>
#include <stdexcept>

namespace b200mix {
namespace {

constexpr double kPi = std::numbers::pi;

// Zero-order modified Bessel function by its power series, summed until the
// term stops changing the sum (core/bsinc_tables.cpp:34-57).
double BesselI0(double x)
{
    const double halfx = x/2.0;
    double term = 1.0, sum = 1.0, prev;
    int k = 1;
    do {
        const double y = halfx / k;
        ++k;
        prev = sum;
        term *= y*y;
        sum += term;
    } while(sum != prev);
    return sum;
}

double Sinc(double x)
{
    constexpr double eps = 2.220446049250313e-16;
    if(!(x > eps || x < -eps)) return 1.0;
    return std::sin(kPi*x) / (kPi*x);
}

double Kaiser(double beta, double k, double i0beta)
{
    if(!(k >= -1.0 && k <= 1.0)) return 0.0;
    return BesselI0(beta * std::sqrt(1.0 - k*k)) / i0beta;
}

double KaiserWidth(double rejection, double order)
{
    if(rejection > 21.19)
        return (rejection-7.95) / (2.285 * kPi*2.0 * order);
    return 5.79 / (kPi*2.0) / order;
}

double KaiserBeta(double rejection)
{
    if(rejection > 50.0) return 0.1102 * (rejection-8.7);
    if(rejection >= 21.0)
        return 0.5842*std::pow(rejection-21.0, 0.4) + 0.07886*(rejection-21.0);
    return 0.0;
}

// SNES-style gaussian kernel sample (core/cubic_tables.cpp:24-33).
double GaussCoeff(double idx)
{
    const double k = 0.5 + idx;
    if(k > 512.0) return 0.0;
    const double s = std::sin(kPi*1.280/1024.0 * k);
    const double t = (std::cos(kPi*2.000/1023.0 * k) - 1.0) * 0.50;
    const double u = (std::cos(kPi*4.000/1023.0 * k) - 1.0) * 0.08;
    return s * (t + u + 1.0) / k;
}

void FillCubicDeltas(std::vector<float> &tab)
{
    auto at = [&tab](unsigned pi, unsigned k) -> float& { return tab[pi*8 + k]; };
    for(unsigned pi = 0;pi+1 < kCubicPhases;++pi)
        for(unsigned k = 0;k < 4;++k)
            at(pi, 4+k) = at(pi+1, k) - at(pi, k);
    const unsigned last = kCubicPhases-1;
    at(last, 4) = 0.0f - at(last, 0);
    at(last, 5) = at(0, 0) - at(last, 1);
    at(last, 6) = at(0, 1) - at(last, 2);
    at(last, 7) = at(0, 2) - at(last, 3);
}

} // namespace

BsincTable BuildBsincTable(double rejection, double order, double maxScale)
{
    BsincTable out;
    const double beta = KaiserBeta(rejection);
    const double scaleBase = KaiserWidth(rejection, order) / 2.0;
    const double scaleLimit = 1.0 / maxScale;
    const double base_a = (order+1.0) / 2.0;

    double a[kBsincScales];
    unsigned mraw[kBsincScales];
    size_t total = 0;
    for(unsigned si = 0;si < kBsincScales;++si)
    {
        const double scale = std::lerp(scaleBase, 1.0, double(si+1u)/double(kBsincScales));
        a[si] = std::min(base_a/scale, base_a*maxScale);
        auto ai = static_cast<unsigned>(a[si]);
        if(double(ai) != a[si]) ++ai; // ceil
        mraw[si] = ai*2u;
        total += size_t{4}*kBsincPhases*((mraw[si]+3u) & ~3u);
    }
    const unsigned ptsMax = (mraw[0]+3u) & ~3u;
    if(ptsMax > kMaxTaps) throw std::runtime_error{"bsinc: too many taps"};

    // filter[si][pi][tap] in f64
    std::vector<double> filter(size_t{kBsincScales}*kBsincPhases*kMaxTaps, 0.0);
    auto F = [&filter](unsigned si, unsigned pi, size_t i) -> double&
    { return filter[(size_t{si}*kBsincPhases + pi)*kMaxTaps + i]; };
    const double i0beta = BesselI0(beta);

    for(unsigned si = 0;si < kBsincScales;++si)
    {
        const unsigned m = mraw[si];
        const double l = std::floor(m*0.5) - 1.0;
        const size_t o = (ptsMax - m)/2u;
        const double scale = std::lerp(scaleBase, 1.0, double(si+1u)/double(kBsincScales));
        const double maxCutoff = (0.5 - scaleBase)*scale;
        const double width = scaleBase * std::max(scaleLimit, scale);
        const double cutoff2 = std::min(maxCutoff, (scale - width)*0.5) * 2.0;
        for(unsigned pi = 0;pi < kBsincPhases;++pi)
        {
            const double phase = l + double(pi)/double(kBsincPhases);
            for(unsigned i = 0;i < m;++i)
            {
                const double x = double(i) - phase;
                F(si, pi, o+i) = Kaiser(beta, x/a[si], i0beta) * cutoff2 * Sinc(cutoff2*x);
            }
        }
    }

    out.tab.assign(total, 0.0f);
    size_t idx = 0;
    for(unsigned si = 0;si < kBsincScales;++si)
    {
        const size_t m = (mraw[si]+3u) & ~3u;
        const size_t o = (ptsMax - m)/2u;
        // per phase: the filter, then its delta to the next phase (the last phase's
        // delta targets phase 0 shifted by one tap)
        for(unsigned pi = 0;pi < kBsincPhases;++pi)
        {
            for(size_t i = 0;i < m;++i)
                out.tab[idx++] = float(F(si, pi, o+i));
            if(pi+1 < kBsincPhases)
            {
                for(size_t i = 0;i < m;++i)
                    out.tab[idx++] = float(F(si, pi+1, o+i) - F(si, pi, o+i));
            }
            else
            {
                out.tab[idx++] = float(0.0 - F(si, pi, o));
                for(size_t i = 1;i < m;++i)
                    out.tab[idx++] = float(F(si, 0, o+i-1) - F(si, pi, o+i));
            }
        }
        // scale deltas and scale+phase deltas towards the next scale
        if(si+1 < kBsincScales)
        {
            for(unsigned pi = 0;pi < kBsincPhases;++pi)
            {
                for(size_t i = 0;i < m;++i)
                    out.tab[idx++] = float(F(si+1, pi, o+i) - F(si, pi, o+i));
                if(pi+1 < kBsincPhases)
                {
                    for(size_t i = 0;i < m;++i)
                        out.tab[idx++] = float((F(si+1, pi+1, o+i)-F(si+1, pi, o+i)) -
                            (F(si, pi+1, o+i)-F(si, pi, o+i)));
                }
                else
                {
                    out.tab[idx++] = float((0.0 - F(si+1, pi, o)) - (0.0 - F(si, pi, o)));
                    for(size_t i = 1;i < m;++i)
                        out.tab[idx++] = float((F(si+1, 0, o+i-1) - F(si+1, pi, o+i)) -
                            (F(si, 0, o+i-1) - F(si, pi, o+i)));
                }
            }
        }
        else
            idx += size_t{kBsincPhases}*m*2; // zeros
    }
    if(idx != total) throw std::runtime_error{"bsinc: table size mismatch"};

    out.scaleBase = float(scaleBase);
    out.scaleRange = float(1.0 / (1.0 - scaleBase));
    for(unsigned i = 0;i < kBsincScales;++i)
        out.m[i] = (mraw[i]+3u) & ~3u;
    out.filterOffset[0] = 0;
    for(unsigned i = 1;i < kBsincScales;++i)
        out.filterOffset[i] = out.filterOffset[i-1] + out.m[i-1]*4u*kBsincPhases;
    return out;
}

BsincState PrepareBsinc(const BsincTable &t, uint32_t increment)
{
    BsincState st;
    unsigned si = kBsincScales-1;
    float sf = 0.0f;
    if(increment > 65536u)
    {
        sf = 65536.0f/float(increment) - t.scaleBase;
        sf = std::max(0.0f, float(kBsincScales)*sf*t.scaleRange - 1.0f);
        si = static_cast<unsigned>(sf);
        sf -= float(si);
        // diagonally-symmetric curve reducing scale-transition ripple (alu.cpp:152-157)
        sf = 1.0f - std::sqrt(1.0f - sf*sf);
    }
    st.sf = sf;
    st.m = t.m[si];
    st.l = st.m/2u - 1u;
    st.offset = t.filterOffset[si];
    return st;
}

std::vector<float> BuildSplineTable()
{
    std::vector<float> tab(kCubicPhases*8, 0.0f);
    constexpr double third = 1.0/3.0, sixth = 1.0/6.0;
    for(unsigned pi = 0;pi < kCubicPhases;++pi)
    {
        const double mu = double(pi)/double(kCubicPhases);
        const double mu2 = mu*mu, mu3 = mu*mu2;
        tab[pi*8+0] = float(      -third*mu + 0.5*mu2  - sixth*mu3);
        tab[pi*8+1] = float(1.0 -    0.5*mu -     mu2  +   0.5*mu3);
        tab[pi*8+2] = float(             mu + 0.5*mu2  -   0.5*mu3);
        tab[pi*8+3] = float(      -sixth*mu            + sixth*mu3);
    }
    FillCubicDeltas(tab);
    return tab;
}

std::vector<float> BuildGaussianTable()
{
    std::vector<float> tab(kCubicPhases*8, 0.0f);
    const double indexScale = 512.0 / double(kCubicPhases*2);
    for(unsigned pi = 0;pi < kCubicPhases;++pi)
    {
        const double c0 = GaussCoeff(double(kCubicPhases + pi)*indexScale);
        const double c1 = GaussCoeff(double(pi)*indexScale);
        const double c2 = GaussCoeff(double(kCubicPhases - pi)*indexScale);
        const double c3 = GaussCoeff(double(kCubicPhases*2 - pi)*indexScale);
        const double scale = 1.0 / (c0 + c1 + c2 + c3);
        tab[pi*8+0] = float(c0*scale);
        tab[pi*8+1] = float(c1*scale);
        tab[pi*8+2] = float(c2*scale);
        tab[pi*8+3] = float(c3*scale);
    }
    FillCubicDeltas(tab);
    return tab;
}

std::vector<float> BuildCubicFilter()
{
    constexpr unsigned steps = 256;
    std::vector<float> f(steps*2 + 1, 0.0f);
    const double indexScale = 512.0 / double(steps*2);
    for(unsigned i = 0;i < steps/2 + 1;++i)
    {
        const double c0 = GaussCoeff(double(steps + i)*indexScale);
        const double c1 = GaussCoeff(double(i)*indexScale);
        const double c2 = GaussCoeff(double(steps - i)*indexScale);
        const double c3 = GaussCoeff(double(steps*2 - i)*indexScale);
        const double scale = 1.0 / (c0 + c1 + c2 + c3);
        f[steps + i] = float(c0*scale);
        f[i] = float(c1*scale);
        f[steps - i] = float(c2*scale);
        f[steps*2 - i] = float(c3*scale);
    }
    return f;
}

} // namespace b200mix
