/* oracle/tables.c — TEST INFRASTRUCTURE ONLY.
 * Plain-C restatement of the reference's resampler coefficient tables:
 *   bsinc12/24/48   core/bsinc_tables.cpp:34-375  (Kaiser-windowed sinc, 16 scales x 32 phases)
 *   spline/gaussian core/cubic_tables.cpp:22-106
 *   gCubicTable     core/cubic_tables.cpp:109-128 (reverb modulation taps)
 * Arithmetic is IEEE f64 with the same operation order, stored as f32.
 * Pinned bit-for-bit against the compiled reference in tests/test_oracle_tables.py.
 */
#include "almix_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define PI 3.14159265358979323846

/* core/bsinc_tables.cpp:34-57 */
static double bessel_i0(double x)
{
    const double x2 = x/2.0;
    double term = 1.0, sum = 1.0, last_sum;
    int k = 1;
    do {
        const double y = x2 / k;
        ++k;
        last_sum = sum;
        term *= y*y;
        sum += term;
    } while(sum != last_sum);
    return sum;
}

/* core/bsinc_tables.cpp:64-69 */
static double sinc(double x)
{
    if(!(x > 2.220446049250313e-16 || x < -2.220446049250313e-16))
        return 1.0;
    return sin(PI*x) / (PI*x);
}

/* core/bsinc_tables.cpp:86-91 */
static double kaiser(double beta, double k, double besseli_0_beta)
{
    if(!(k >= -1.0 && k <= 1.0))
        return 0.0;
    return bessel_i0(beta * sqrt(1.0 - k*k)) / besseli_0_beta;
}

/* libstdc++ std::lerp for doubles (the reference calls it through altypes.hpp:1193). */
static double lerp_f64(double a, double b, double t)
{
    if((a <= 0 && b >= 0) || (a >= 0 && b <= 0))
        return t*b + (1 - t)*a;
    if(t == 1) return b;
    const double x = a + t*(b - a);
    return ((t > 1) == (b > a)) ? (b < x ? x : b) : (b > x ? x : b);
}

/* core/bsinc_tables.cpp:96-113 */
static double kaiser_width(double rejection, double order)
{
    if(rejection > 21.19)
        return (rejection-7.95) / (2.285 * PI*2.0 * order);
    return 5.79 / (PI*2.0) / order;
}
static double kaiser_beta(double rejection)
{
    if(rejection > 50.0)
        return 0.1102 * (rejection-8.7);
    if(rejection >= 21.0)
        return 0.5842*pow(rejection-21.0, 0.4) + 0.07886*(rejection-21.0);
    return 0.0;
}

/* BSincHeader + BSincFilterArray + GenerateBSincTable, core/bsinc_tables.cpp:116-368 */
int oracle_build_bsinc(oracle_bsinc_table *t, double rejection, double order, double maxScale)
{
    enum { SC = ORACLE_BSINC_SCALES, PH = ORACLE_BSINC_PHASES };
    const double beta = kaiser_beta(rejection);
    const double scaleBase = kaiser_width(rejection, order) / 2.0;
    const double scaleLimit = 1.0 / maxScale;
    double a[SC];
    unsigned m_raw[SC];
    size_t total = 0;
    const double base_a = (order+1.0) / 2.0;
    for(unsigned si = 0;si < SC;++si)
    {
        const double scale = lerp_f64(scaleBase, 1.0, (double)(si+1u) / (double)SC);
        const double a0 = base_a/scale, a1 = base_a*maxScale;
        a[si] = (a0 < a1) ? a0 : a1;
        unsigned a_ = (unsigned)a[si];
        a_ += ((double)a_ != a[si]) ? 1u : 0u;
        m_raw[si] = a_ * 2u;
        total += (size_t)4 * PH * ((m_raw[si]+3u) & ~3u);
    }
    const unsigned pts_max = (m_raw[0]+3u) & ~3u;
    double (*filter)[PH][ORACLE_MAX_TAPS] = calloc(SC, sizeof(*filter));
    float *tab = calloc(total, sizeof(float));
    if(!filter || !tab) { free(filter); free(tab); return -1; }
    const double besseli_0_beta = bessel_i0(beta);

    for(unsigned si = 0;si < SC;++si)
    {
        const unsigned m = m_raw[si];
        const double l = floor(m*0.5) - 1.0;
        const size_t o = (pts_max - m) / 2u;
        const double scale = lerp_f64(scaleBase, 1.0, (double)(si+1u)/(double)SC);
        const double max_cutoff = (0.5 - scaleBase)*scale;
        const double width = scaleBase * ((scaleLimit > scale) ? scaleLimit : scale);
        const double c0 = (scale - width)*0.5;
        const double cutoff2 = ((max_cutoff < c0) ? max_cutoff : c0) * 2.0;
        for(unsigned pi = 0;pi < PH;++pi)
        {
            const double phase = l + (double)pi/(double)PH;
            for(unsigned i = 0;i < m;++i)
            {
                const double x = (double)i - phase;
                filter[si][pi][o+i] = kaiser(beta, x/a[si], besseli_0_beta) * cutoff2
                    * sinc(cutoff2*x);
            }
        }
    }

    size_t idx = 0;
    for(unsigned si = 0;si < SC;++si)
    {
        const size_t m = (m_raw[si]+3u) & ~3u;
        const size_t o = (pts_max - m) / 2u;
        for(unsigned pi = 0;pi < PH;++pi)
        {
            for(size_t i = 0;i < m;++i)
                tab[idx++] = (float)filter[si][pi][o+i];
            if(pi < PH-1)
            {
                for(size_t i = 0;i < m;++i)
                    tab[idx++] = (float)(filter[si][pi+1][o+i] - filter[si][pi][o+i]);
            }
            else
            {
                tab[idx++] = (float)(0.0 - filter[si][pi][o]);
                for(size_t i = 1;i < m;++i)
                    tab[idx++] = (float)(filter[si][0][o+i-1] - filter[si][pi][o+i]);
            }
        }
        if(si < SC-1)
        {
            for(unsigned pi = 0;pi < PH;++pi)
            {
                for(size_t i = 0;i < m;++i)
                    tab[idx++] = (float)(filter[si+1][pi][o+i] - filter[si][pi][o+i]);
                if(pi < PH-1)
                {
                    for(size_t i = 0;i < m;++i)
                        tab[idx++] = (float)((filter[si+1][pi+1][o+i]-filter[si+1][pi][o+i]) -
                            (filter[si][pi+1][o+i]-filter[si][pi][o+i]));
                }
                else
                {
                    tab[idx++] = (float)((0.0 - filter[si+1][pi][o]) - (0.0 - filter[si][pi][o]));
                    for(size_t i = 1;i < m;++i)
                        tab[idx++] = (float)((filter[si+1][0][o+i-1] - filter[si+1][pi][o+i]) -
                            (filter[si][0][o+i-1] - filter[si][pi][o+i]));
                }
            }
        }
        else
        {
            idx += (size_t)PH * m * 2; /* zero-filled (calloc) */
        }
    }
    free(filter);
    if(idx != total) { free(tab); return -2; }

    t->scaleBase = (float)scaleBase;
    t->scaleRange = (float)(1.0 / (1.0 - scaleBase));
    for(unsigned i = 0;i < SC;++i)
        t->m[i] = (m_raw[i]+3u) & ~3u;
    t->filterOffset[0] = 0;
    for(unsigned i = 1;i < SC;++i)
        t->filterOffset[i] = t->filterOffset[i-1] + t->m[i-1]*4u*PH;
    t->tab = tab;
    t->total = total;
    return 0;
}

/* core/cubic_tables.cpp:24-33 */
static double gauss_coeff(double idx)
{
    const double k = 0.5 + idx;
    if(k > 512.0) return 0.0;
    const double s = sin(PI*1.280/1024.0 * k);
    const double t = (cos(PI*2.000/1023.0 * k) - 1.0) * 0.50;
    const double u = (cos(PI*4.000/1023.0 * k) - 1.0) * 0.08;
    return s * (t + u + 1.0) / k;
}

static void cubic_deltas(float tab[ORACLE_CUBIC_PHASES][8])
{
    enum { PH = ORACLE_CUBIC_PHASES };
    for(unsigned pi = 0;pi < PH-1;++pi)
        for(unsigned k = 0;k < 4;++k)
            tab[pi][4+k] = tab[pi+1][k] - tab[pi][k];
    const unsigned pi = PH-1;
    tab[pi][4+0] = 0.0f - tab[pi][0];
    tab[pi][4+1] = tab[0][0] - tab[pi][1];
    tab[pi][4+2] = tab[0][1] - tab[pi][2];
    tab[pi][4+3] = tab[0][2] - tab[pi][3];
}

/* GaussianTable::GaussianTable, core/cubic_tables.cpp:37-72 */
void oracle_build_gaussian(float tab[ORACLE_CUBIC_PHASES][8])
{
    enum { PH = ORACLE_CUBIC_PHASES };
    const double IndexScale = 512.0 / (double)(PH*2);
    for(unsigned pi = 0;pi < PH;++pi)
    {
        const double c0 = gauss_coeff((double)(PH + pi)*IndexScale);
        const double c1 = gauss_coeff((double)pi*IndexScale);
        const double c2 = gauss_coeff((double)(PH - pi)*IndexScale);
        const double c3 = gauss_coeff((double)(PH*2 - pi)*IndexScale);
        const double scale = 1.0 / (c0 + c1 + c2 + c3);
        tab[pi][0] = (float)(c0*scale);
        tab[pi][1] = (float)(c1*scale);
        tab[pi][2] = (float)(c2*scale);
        tab[pi][3] = (float)(c3*scale);
    }
    cubic_deltas(tab);
}

/* SplineTable::SplineTable, core/cubic_tables.cpp:74-105 */
void oracle_build_spline(float tab[ORACLE_CUBIC_PHASES][8])
{
    enum { PH = ORACLE_CUBIC_PHASES };
    const double third = 1.0/3.0, sixth = 1.0/6.0;
    for(unsigned pi = 0;pi < PH;++pi)
    {
        const double mu = (double)pi / (double)PH;
        const double mu2 = mu*mu;
        const double mu3 = mu*mu2;
        tab[pi][0] = (float)(      -third*mu + 0.5*mu2  - sixth*mu3);
        tab[pi][1] = (float)(1.0 -    0.5*mu -     mu2  +   0.5*mu3);
        tab[pi][2] = (float)(             mu + 0.5*mu2  -   0.5*mu3);
        tab[pi][3] = (float)(      -sixth*mu            + sixth*mu3);
    }
    cubic_deltas(tab);
}

/* CubicFilter::CubicFilter, core/cubic_tables.cpp:109-128: 513 floats */
void oracle_build_cubic_filter(float filter[513])
{
    enum { STEPS = 256 };
    const double IndexScale = 512.0 / (double)(STEPS*2);
    for(unsigned i = 0;i < STEPS/2 + 1;++i)
    {
        const double c0 = gauss_coeff((double)(STEPS + i)*IndexScale);
        const double c1 = gauss_coeff((double)i*IndexScale);
        const double c2 = gauss_coeff((double)(STEPS - i)*IndexScale);
        const double c3 = gauss_coeff((double)(STEPS*2 - i)*IndexScale);
        const double scale = 1.0 / (c0 + c1 + c2 + c3);
        filter[STEPS + i] = (float)(c0*scale);
        filter[i] = (float)(c1*scale);
        filter[STEPS - i] = (float)(c2*scale);
        filter[STEPS*2 - i] = (float)(c3*scale);
    }
}

/* BsincPrepare, alc/alu.cpp:140-165 */
void oracle_bsinc_prepare(const oracle_bsinc_table *t, uint32_t increment, oracle_bsinc_state *st)
{
    unsigned si = ORACLE_BSINC_SCALES - 1;
    float sf = 0.0f;
    if(increment > 65536u)
    {
        sf = 65536.0f/(float)increment - t->scaleBase;
        sf = 16.0f*sf*t->scaleRange - 1.0f;
        if(!(sf > 0.0f)) sf = 0.0f;
        si = (unsigned)sf; /* float2uint: truncation, sf >= 0 and < 16 here */
        sf -= (float)si;
        sf = 1.0f - sqrtf(1.0f - sf*sf);
    }
    st->sf = sf;
    st->m = t->m[si];
    st->l = st->m/2u - 1u;
    st->filter = t->tab + t->filterOffset[si];
}
