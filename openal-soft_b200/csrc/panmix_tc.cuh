// panmix_tc.cuh — the dense ambisonic pan-mix on the 5th-generation tensor cores.
//
// A device that mixes above first order (config 4a: third-order B-Format output, 16 dry
// channels) sums every voice into every channel: Dry[c][i] += line_v[i] * gain_v,c — the
// reference's MixSamples 1->many (core/mixer/mixer_c.cpp:150-186) over all voices is the
// contraction  Dry[16 x 1024] = G[16 x V] . S[V x 1024], a true dense GEMM with K = voices.
// Past the gain fade (Counter <= 64 samples, core/voice.cpp:1093) the gains are constants, so
// samples 128..1023 of every line go through tcgen05.mma:
//
//   D[M = 128 samples][N = 16 channels] (+)= A[M x K = 8 voices] . B[K x N]      (kind::tf32)
//
//   A = the parked lines.  They lie with the samples (M) contiguous, but kind::tf32 takes its
//       shared-memory operands K-major only (an MN-major descriptor yields zeros — measured,
//       tools/ubench/umma_probe.cu), so the lines are transposed on their way into shared memory:
//       8-row x 16-byte core matrices, (m%8)*16 + (m/8)*SBO + (k/4)*LBO + (k%4)*4, with
//       LBO = 144 and SBO = 288 bytes so that the transposing scalar stores are conflict-free.
//   B = geff[entry][channel] (k_send_gains_prepare), K-major
//   D = fp32 accumulators in TENSOR MEMORY, 7 sample tiles x 16 columns, accumulated over ALL
//       the voices a CTA owns and read back once (tcgen05.ld) into the CTA's partial row.
//
// fp32 parity from tf32 tensor cores: both operands are split  x = hi + lo  with hi = x rounded
// to tf32 and lo = x - hi (exact), and three MMAs accumulate hi.hi + lo.hi + hi.lo; the dropped
// lo.lo term and lo's own truncation are ~2^-22 relative per product, an order below the parity
// budget (tests: at-size oracle comparison and ambi3 goldens within 1e-6).
// Samples 0..127 (the fades) stay with k_send_mix<16>'s first tile; both kernels write disjoint
// columns of the same per-chunk partial rows, k_reduce_rows sums the chunks in fixed order.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "async_ptx.cuh"
#include "effect_kernels.cuh"

namespace b200mix {

constexpr int kPmTiles = 7;                         // sample tiles 1..7 of 128 (896 samples)
constexpr int kPmK = 8;                             // voices per MMA (tf32: 32 bytes of K)
constexpr int kPmN = 16;                            // channels (padded)
constexpr int kPmLboA = 144, kPmSboA = 288;         // A core-matrix strides (padded: conflict-free stores)
constexpr int kPmTileBytes = 16*kPmSboA;            // one A tile: 16 groups of 8 samples x 2 K halves
constexpr int kPmStageBytes = 2*kPmTiles*kPmTileBytes + 2*kPmN*kPmK*4;     // A hi, A lo, B hi, B lo
constexpr int kPmStages = 2;
constexpr uint32_t kPmTmemCols = 128;               // 7 x 16 columns, power of two

struct PanMixTcParams {
    const uint32_t *slot_start;     // [2] entry range of the dry bus
    const SendEntry *entries;
    const uint32_t *sendinfo;
    const float *xscratch;          // [max_voices][1024]
    const float *dline;             // [max_voices][1024] lines of deferred (direct-filtered) voices
    const float *geff;              // [entries][cw] constant gain of every entry-channel
    uint32_t cw, chunks;
    float *partial;                 // [chunks][cw][1024]
};

// tf32 split of an fp32 value: hi = round-to-nearest at 10 mantissa bits, lo = x - hi truncated
__device__ __forceinline__ void tf32_split(float x, float &hi, float &lo)
{
    const uint32_t b = __float_as_uint(x);
    hi = __uint_as_float((b + 0x1000u) & 0xffffe000u);
    lo = __uint_as_float(__float_as_uint(x - hi) & 0xffffe000u);
}

// grid = chunks (the entry ranges k_send_mix uses), 128 threads, kPmStages*kPmStageBytes dynamic smem
__global__ void __launch_bounds__(128, 1) k_panmix_tc(const PanMixTcParams Q)
{
    extern __shared__ __align__(1024) unsigned char pm_smem[];
    __shared__ uint64_t bar_free[kPmStages];
    __shared__ uint32_t tmem_slot;
    const uint32_t t = threadIdx.x, warp = t >> 5, lane = t & 31u;

    uint32_t e0 = Q.slot_start[0], e1 = Q.slot_start[1];
    {
        const uint32_t per = (e1 - e0 + Q.chunks - 1u)/Q.chunks;
        e0 = min(e0 + blockIdx.x*per, e1);
        e1 = min(e0 + per, e1);
    }
    const uint32_t nkb = (e1 - e0 + uint32_t(kPmK) - 1u)/uint32_t(kPmK);
    float *out = Q.partial + size_t(blockIdx.x)*Q.cw*kLine;

    if(warp == 0) tmem_alloc<kPmTmemCols>(&tmem_slot);
    if(t == 0)
    {
        for(int s = 0;s < kPmStages;++s) mbar_init(&bar_free[s], 1u);
        mbar_fence_init();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;
    constexpr uint32_t idesc = umma_idesc_tf32(128u, uint32_t(kPmN), /*A K-major*/false, /*B K-major*/false);

    // this thread's voice of a K block and its sample groups (4 consecutive samples each)
    const uint32_t kk = t & 7u, gl = t >> 3;             // gl in [0,16)
    for(uint32_t kb = 0;kb < nkb;++kb)
    {
        const uint32_t st = kb % uint32_t(kPmStages);
        unsigned char *stage = pm_smem + size_t(st)*kPmStageBytes;
        if(kb >= uint32_t(kPmStages))                    // the MMAs that read this stage are done
            mbar_wait(&bar_free[st], ((kb / uint32_t(kPmStages)) - 1u) & 1u);
        // ---- A: 8 lines x 896 samples, split, transposed into K-major core matrices
        const uint32_t e = e0 + kb*uint32_t(kPmK) + kk;
        const bool live = e < e1;
        const SendEntry en = Q.entries[live ? e : e0];
        const float *line = ((Q.dline && (Q.sendinfo[en.voice] & kSiDeferred)) ? Q.dline : Q.xscratch)
            + size_t(en.voice)*kLine + 128;
        float4 v[14];
        #pragma unroll
        for(int i = 0;i < 14;++i)
            v[i] = live ? __ldg(reinterpret_cast<const float4*>(line) + gl + 16u*uint32_t(i))
                        : make_float4(0.f, 0.f, 0.f, 0.f);
        #pragma unroll
        for(int i = 0;i < 14;++i)
        {
            const uint32_t g = gl + 16u*uint32_t(i);     // [0, 224)
            float4 hi, lo;
            tf32_split(v[i].x, hi.x, lo.x); tf32_split(v[i].y, hi.y, lo.y);
            tf32_split(v[i].z, hi.z, lo.z); tf32_split(v[i].w, hi.w, lo.w);
            // samples m = 4*(g % 32) + j of tile g / 32, voice kk
            const uint32_t mg = (g & 31u) >> 1;          // group of 8 rows
            const uint32_t off = (g >> 5)*uint32_t(kPmTileBytes) + mg*uint32_t(kPmSboA)
                + ((g & 1u)*4u)*16u + (kk >> 2)*uint32_t(kPmLboA) + (kk & 3u)*4u;
            float *ph = reinterpret_cast<float*>(stage + off);
            float *pl = reinterpret_cast<float*>(stage + kPmTiles*kPmTileBytes + off);
            ph[0] = hi.x; ph[4] = hi.y; ph[8] = hi.z; ph[12] = hi.w;       // rows are 16 bytes apart
            pl[0] = lo.x; pl[4] = lo.y; pl[8] = lo.z; pl[12] = lo.w;
        }
        // ---- B: gains [n = channel][k = voice], K-major core matrices:
        //      (n % 8)*16 + (n / 8)*256 + (k / 4)*128 + (k % 4)*4
        {
            const uint32_t n = t & 15u, k = t >> 4;      // 16 x 8
            const uint32_t eb = e0 + kb*uint32_t(kPmK) + k;
            float g = 0.0f;
            if(eb < e1 && n < Q.cw) g = Q.geff[size_t(eb)*Q.cw + n];
            float hi, lo;
            tf32_split(g, hi, lo);
            const uint32_t off = (n & 7u)*16u + (n >> 3)*256u + (k >> 2)*128u + (k & 3u)*4u;
            unsigned char *bs = stage + 2*kPmTiles*kPmTileBytes;
            *reinterpret_cast<float*>(bs + off) = hi;
            *reinterpret_cast<float*>(bs + kPmN*kPmK*4 + off) = lo;
        }
        fence_proxy_async_smem();                        // generic-proxy stores -> tensor core reads
        __syncthreads();
        if(t == 0)
        {
            tc_fence_after_sync();
            const uint32_t sa = smem_u32(stage);
            const uint32_t sb = sa + 2u*kPmTiles*kPmTileBytes;
            const uint64_t bhi = umma_smem_desc(sb, 128u, 256u), blo = umma_smem_desc(sb + kPmN*kPmK*4, 128u, 256u);
            #pragma unroll
            for(int tile = 0;tile < kPmTiles;++tile)
            {
                const uint64_t ahi = umma_smem_desc(sa + uint32_t(tile)*kPmTileBytes, kPmLboA, kPmSboA);
                const uint64_t alo = umma_smem_desc(sa + uint32_t(kPmTiles + tile)*kPmTileBytes, kPmLboA, kPmSboA);
                const uint32_t d = tmem + uint32_t(tile)*uint32_t(kPmN);
                umma_tf32(d, ahi, bhi, idesc, kb != 0u);
                umma_tf32(d, alo, bhi, idesc, true);
                umma_tf32(d, ahi, blo, idesc, true);
            }
            umma_commit(&bar_free[st]);                  // arrives when these MMAs have completed
        }
    }
    // ---- all MMAs done: the last commit of every stage in use
    for(uint32_t s = 0;s < uint32_t(kPmStages);++s)
    {
        if(nkb <= s) continue;
        const uint32_t uses = (nkb - 1u - s)/uint32_t(kPmStages) + 1u;      // commits on this stage
        mbar_wait(&bar_free[s], (uses - 1u) & 1u);
    }
    tc_fence_after_sync();
    // ---- epilogue: TMEM -> registers -> the CTA's partial row (samples 128..1023)
    //      warp w reads lanes 32w..32w+31 of every tile = samples 128 + tile*128 + 32w + lane
    #pragma unroll 1
    for(int tile = 0;tile < kPmTiles;++tile)
    {
        float v[16];
        if(nkb)
            tmem_ld_32x16(tmem + ((warp*32u) << 16) + uint32_t(tile)*uint32_t(kPmN), v);
        else
        {
            #pragma unroll
            for(int c = 0;c < 16;++c) v[c] = 0.0f;
        }
        const uint32_t i = 128u + uint32_t(tile)*128u + warp*32u + lane;
        #pragma unroll
        for(int c = 0;c < 16;++c)
            if(uint32_t(c) < Q.cw) out[size_t(c)*kLine + i] = v[c];
    }
    tc_fence_before_sync();
    __syncthreads();
    if(warp == 0) tmem_dealloc<kPmTmemCols>(tmem);
}

} // namespace b200mix
