// shard_kernels.cuh — the per-update exchange of a voice-sharded device set as
// hand-written one-shot NVLink transfers: every rank stores its block straight into the
// receiver's memory (peer-mapped through CUDA IPC), publishes an epoch flag with release
// semantics at system scope, and the receiver sums the blocks in RANK ORDER (deterministic —
// unlike a ring/tree the result does not depend on arrival order).
//
//   RealOut reduce (alc/alu.cpp:2439-2443 is linear, so the ranks' RealOut blocks add up):
//       k_shard_push  on ranks != 0  ->  k_shard_sum on rank 0
//   Wet reduce-scatter (effects consume the SUMMED send input, alc/alu.cpp:2252-2256; slot s
//   is owned by rank s mod G):
//       k_shard_push (one target per owner)  ->  k_shard_sum_wet on every owner
//
// Blocks are double-buffered by epoch parity and a sender waits for the receiver's
// acknowledgement of epoch e-2 before it overwrites that parity, so a rank may run one update
// ahead of the root but never two.  Every wait has a time-out (a peer that died must not hang
// the GPU): the kernel then raises ShardCtl::status and the host reports B200MIX_ERR_CUDA.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace b200mix {

constexpr uint32_t kShardMaxWorld = 16;
constexpr unsigned long long kShardTimeoutNs = 4000000000ull;     // 4 s

struct alignas(256) ShardCtl {
    uint32_t real_flag[kShardMaxWorld];   // [src]   newest epoch src has delivered here (rank 0 only)
    uint32_t real_ack[kShardMaxWorld];    // [0]     newest epoch rank 0 has consumed from this rank
    uint32_t wet_flag[kShardMaxWorld];    // [src]   newest epoch src has delivered to this owner
    uint32_t wet_ack[kShardMaxWorld];     // [owner] newest epoch that owner has consumed from this rank
    uint32_t status;                      // != 0: a wait timed out (1 flag, 2 ack)
    uint32_t pad[63];
};
static_assert(sizeof(ShardCtl) == 512, "ShardCtl layout");

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v)
{ asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// waits until *flag has reached `epoch` (epochs only grow; compared as a signed difference)
__device__ __forceinline__ bool shard_wait(const uint32_t *flag, uint32_t epoch)
{
    const unsigned long long t0 = global_ns();
    while(int32_t(ld_acquire_sys(flag) - epoch) < 0)
    {
        if(global_ns() - t0 > kShardTimeoutNs) return false;
        __nanosleep(64);
    }
    return true;
}

struct ShardPushParams {
    const float *src;                 // this rank's block (RealOut, or the Wet buffers)
    uint32_t rank, world, epoch;
    uint32_t floats;                  // RealOut: floats of the block
    // wet: slots are dealt round-robin (slot s -> owner s mod world); slot_floats = cw*1024
    uint32_t wet, num_slots, slot_floats, owned_max;
    ShardCtl *own;
    char *peer[kShardMaxWorld];       // peer-mapped blocks (peer[rank] == own block)
    size_t off_data;                  // byte offset of recv[2][world][...] in a block
    size_t per_src_floats;            // floats one source occupies in a receive buffer
    uint32_t *counters;               // [gridDim.y] last-block counters (device-local, zero at rest)
};

// grid (chunks, targets): targets = 1 for the RealOut push (target rank 0), world for the wet
// push (blockIdx.y = owner; the own rank's row returns at once).
__global__ void __launch_bounds__(256) k_shard_push(const ShardPushParams P)
{
    const uint32_t tgt = P.wet ? blockIdx.y : 0u;
    if(tgt == P.rank) return;
    __shared__ int ok;
    if(threadIdx.x == 0)
    {
        // the receiver must have consumed the block two epochs back (same parity)
        const uint32_t *ack = P.wet ? &P.own->wet_ack[tgt] : &P.own->real_ack[0];
        ok = (P.epoch <= 2u) || shard_wait(ack, P.epoch - 2u);
        if(!ok) P.own->status = 2u;
    }
    __syncthreads();
    float *dst = reinterpret_cast<float*>(P.peer[tgt] + P.off_data)
        + (size_t(P.epoch & 1u)*P.world + P.rank)*P.per_src_floats;
    if(!P.wet)
    {
        const float4 *s4 = reinterpret_cast<const float4*>(P.src);
        float4 *d4 = reinterpret_cast<float4*>(dst);
        for(uint32_t i = blockIdx.x*blockDim.x + threadIdx.x;i < P.floats/4u;i += gridDim.x*blockDim.x)
            d4[i] = s4[i];
    }
    else
    {
        // owner tgt's slots: tgt, tgt + world, ... -> local index j
        const uint32_t f4 = P.slot_floats/4u;
        for(uint32_t j = 0, s = tgt;s < P.num_slots;++j, s += P.world)
        {
            const float4 *s4 = reinterpret_cast<const float4*>(P.src + size_t(s)*P.slot_floats);
            float4 *d4 = reinterpret_cast<float4*>(dst + size_t(j)*P.slot_floats);
            for(uint32_t i = blockIdx.x*blockDim.x + threadIdx.x;i < f4;i += gridDim.x*blockDim.x)
                d4[i] = s4[i];
        }
    }
    __threadfence_system();
    __syncthreads();
    if(threadIdx.x == 0)
    {
        uint32_t *cnt = P.counters + blockIdx.y;
        if(atomicAdd(cnt, 1u) == gridDim.x - 1u)
        {
            *cnt = 0u;
            // (the release store below orders everything the counter made visible before it)
            ShardCtl *tc = reinterpret_cast<ShardCtl*>(P.peer[tgt]);
            st_release_sys(P.wet ? &tc->wet_flag[P.rank] : &tc->real_flag[P.rank], P.epoch);
        }
    }
}

struct ShardSumParams {
    float *dst;                       // RealOut of rank 0 / this owner's Wet buffers (holds the own part)
    uint32_t rank, world, epoch;
    uint32_t floats;                  // RealOut block size
    uint32_t wet, num_slots, slot_floats, owned_max;
    ShardCtl *own;
    char *peer[kShardMaxWorld];
    size_t off_data, per_src_floats;
    uint32_t *counter;
};

// Sums the delivered blocks in rank order (the own contribution in its place) and
// acknowledges the epoch to every sender.
__global__ void __launch_bounds__(256) k_shard_sum(const ShardSumParams P)
{
    __shared__ int ok;
    if(threadIdx.x == 0) ok = 1;
    __syncthreads();
    if(threadIdx.x < P.world && threadIdx.x != P.rank)
    {
        const uint32_t *flag = P.wet ? &P.own->wet_flag[threadIdx.x] : &P.own->real_flag[threadIdx.x];
        if(!shard_wait(flag, P.epoch)) { ok = 0; P.own->status = 1u; }
    }
    __syncthreads();
    const float *recv = reinterpret_cast<const float*>(reinterpret_cast<char*>(P.own) + P.off_data)
        + size_t(P.epoch & 1u)*P.world*P.per_src_floats;
    if(ok)
    {
        const uint32_t owned = P.wet ? (P.num_slots > P.rank ? (P.num_slots - P.rank + P.world - 1u)/P.world : 0u) : 1u;
        const uint32_t f4 = (P.wet ? P.slot_floats : P.floats)/4u;
        for(uint32_t j = 0;j < owned;++j)
        {
            float4 *d4 = reinterpret_cast<float4*>(P.wet ? P.dst + size_t(P.rank + j*P.world)*P.slot_floats : P.dst);
            for(uint32_t i = blockIdx.x*blockDim.x + threadIdx.x;i < f4;i += gridDim.x*blockDim.x)
            {
                const float4 mine = d4[i];
                float4 tot = make_float4(0.f, 0.f, 0.f, 0.f);
                for(uint32_t r = 0;r < P.world;++r)
                {
                    float4 v = mine;
                    if(r != P.rank)       // delivered by a peer: bypass L1 (the lines change under us)
                        v = __ldcg(reinterpret_cast<const float4*>(recv + size_t(r)*P.per_src_floats
                            + size_t(j)*(P.wet ? P.slot_floats : 0u)) + i);
                    if(r == 0) tot = v;
                    else { tot.x += v.x; tot.y += v.y; tot.z += v.z; tot.w += v.w; }
                }
                d4[i] = tot;
            }
        }
    }
    __threadfence_system();
    __syncthreads();
    // the last block acknowledges: every block's loads of the delivered data are complete once its
    // fence + counter increment are visible.  One thread per peer — seven remote release stores
    // issued one after the other were ~20 us of rank 0's update on 8 GPUs.
    __shared__ int last;
    if(threadIdx.x == 0)
    {
        last = (atomicAdd(P.counter, 1u) == gridDim.x - 1u) ? 1 : 0;
        if(last) *P.counter = 0u;
    }
    __syncthreads();
    if(last && threadIdx.x < P.world && threadIdx.x != P.rank)
    {
        ShardCtl *pc = reinterpret_cast<ShardCtl*>(P.peer[threadIdx.x]);
        st_release_sys(P.wet ? &pc->wet_ack[P.rank] : &pc->real_ack[0], P.epoch);
    }
}

} // namespace b200mix
