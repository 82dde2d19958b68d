// Probe of the tcgen05.mma kind::tf32 operand layouts (no swizzle): one MMA, M=128 N=16 K=8, with
// shared-memory images and descriptors built on the host under several layout hypotheses; reports
// which hypothesis reproduces A.B.   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++20
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../openal-soft_b200/csrc/async_ptx.cuh"
using namespace b200mix;
#define CK(x) do { cudaError_t e = (x); if(e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while(0)

struct Probe { uint32_t a_lbo, a_sbo, b_lbo, b_sbo, idesc, a_bytes, b_bytes, mode; };

__global__ void __launch_bounds__(128) k_probe(const unsigned char *img_a, const unsigned char *img_b, Probe P, float *out)
{
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ uint64_t bar; __shared__ uint32_t slot;
    const uint32_t t = threadIdx.x, warp = t >> 5, lane = t & 31u;
    unsigned char *sa = sm, *sb = sm + 16384;
    for(uint32_t i = t;i < P.a_bytes;i += 128) sa[i] = img_a[i];
    for(uint32_t i = t;i < P.b_bytes;i += 128) sb[i] = img_b[i];
    if(warp == 0) tmem_alloc<32>(&slot);
    if(t == 0) { mbar_init(&bar, 1u); mbar_fence_init(); }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = slot;
    if(P.mode == 1u)
    {   // addressing check: store lane*100 + col, read back
        uint32_t v[16];
        for(int c = 0;c < 16;++c) v[c] = __float_as_uint(float((warp*32u + lane)*100u + c));
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
            :: "r"(tmem + ((warp*32u) << 16)), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
               "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    else if(t == 0)
    {
        const uint64_t ad = umma_smem_desc(smem_u32(sa), P.a_lbo, P.a_sbo), bd = umma_smem_desc(smem_u32(sb), P.b_lbo, P.b_sbo);
        umma_tf32(tmem, ad, bd, P.idesc, false);
        umma_commit(&bar);
    }
    if(P.mode != 1u) mbar_wait(&bar, 0u);
    tc_fence_after_sync();
    __syncthreads();
    float v[16];
    tmem_ld_32x16(tmem + ((warp*32u) << 16), v);
    for(int c = 0;c < 16;++c) out[(warp*32u + lane)*16u + c] = v[c];
    tc_fence_before_sync();
    __syncthreads();
    if(warp == 0) tmem_dealloc<32>(tmem);
}

int main()
{
    const int M = 128, N = 16, K = 8;
    std::vector<float> A(M*K), B(K*N), ref(M*N, 0.f);
    srand(3);
    for(auto &x : A) x = float(rand()%17 - 8);
    for(auto &x : B) x = float(rand()%13 - 6);
    for(int m = 0;m < M;++m) for(int n = 0;n < N;++n) { float s = 0; for(int k = 0;k < K;++k) s += A[m*K+k]*B[k*N+n]; ref[m*N+n] = s; }
    unsigned char *d_a, *d_b; float *d_out;
    CK(cudaMalloc(&d_a, 16384)); CK(cudaMalloc(&d_b, 4096)); CK(cudaMalloc(&d_out, M*N*4));
    CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
    std::vector<float> out(M*N);
    auto run = [&](const char *name, std::vector<unsigned char> &ia, std::vector<unsigned char> &ib, Probe P) -> int {
        CK(cudaMemcpy(d_a, ia.data(), ia.size(), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_b, ib.data(), ib.size(), cudaMemcpyHostToDevice));
        CK(cudaMemset(d_out, 0xff, M*N*4));
        P.a_bytes = uint32_t(ia.size()); P.b_bytes = uint32_t(ib.size());
        k_probe<<<1, 128, 32768>>>(d_a, d_b, P, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if(e != cudaSuccess) { printf("%-40s CUDA error %s\n", name, cudaGetErrorString(e)); return 1; }
        CK(cudaMemcpy(out.data(), d_out, M*N*4, cudaMemcpyDeviceToHost));
        int bad = 0, zero = 0;
        for(int i = 0;i < M*N;++i) { if(out[i] != ref[i]) ++bad; if(out[i] == 0.f) ++zero; }
        printf("%-40s mismatches %4d / %d   zeros %4d   D[0][0..3] = %g %g %g %g  (ref %g %g %g %g)  D[5][2]=%g (ref %g)\n", name, bad, M*N, zero,
            out[0], out[1], out[2], out[3], ref[0], ref[1], ref[2], ref[3], out[5*16+2], ref[5*16+2]);
        return 0;
    };
    std::vector<unsigned char> ia(4096), ib(512);
    auto put = [](std::vector<unsigned char> &img, size_t off, float v) { std::memcpy(img.data() + off, &v, 4); };
    // mode 1: TMEM addressing
    {
        Probe P{}; P.mode = 1;
        CK(cudaMemset(d_out, 0, M*N*4));
        k_probe<<<1, 128, 32768>>>(d_a, d_b, P, d_out);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(out.data(), d_out, M*N*4, cudaMemcpyDeviceToHost));
        int bad = 0; for(int m = 0;m < M;++m) for(int c = 0;c < 16;++c) if(out[m*16+c] != float(m*100 + c)) ++bad;
        printf("tmem st/ld addressing: mismatches %d\n", bad);
    }
    // B K-major interleave: (n%8)*16 + (n/8)*SBO + (k/4)*LBO + (k%4)*4, SBO=256 LBO=128
    auto fillB_K = [&](uint32_t lbo, uint32_t sbo) { std::fill(ib.begin(), ib.end(), 0); for(int n = 0;n < N;++n) for(int k = 0;k < K;++k) put(ib, (n%8)*16 + (n/8)*sbo + (k/4)*lbo + (k%4)*4, B[k*N+n]); };
    // B MN-major interleave: N contiguous: (n/4)*SBO + k*16 + (n%4)*4  (4 N-groups)
    auto fillB_MN = [&](uint32_t sbo) { std::fill(ib.begin(), ib.end(), 0); for(int n = 0;n < N;++n) for(int k = 0;k < K;++k) put(ib, (n/4)*sbo + k*16 + (n%4)*4, B[k*N+n]); };
    // A MN-major interleave: (m/4)*SBO + k*16 + (m%4)*4
    auto fillA_MN = [&](uint32_t sbo) { std::fill(ia.begin(), ia.end(), 0); for(int m = 0;m < M;++m) for(int k = 0;k < K;++k) put(ia, (m/4)*sbo + k*16 + (m%4)*4, A[m*K+k]); };
    // A K-major interleave: (m%8)*16 + (m/8)*SBO + (k/4)*LBO + (k%4)*4
    auto fillA_K = [&](uint32_t lbo, uint32_t sbo) { std::fill(ia.begin(), ia.end(), 0); for(int m = 0;m < M;++m) for(int k = 0;k < K;++k) put(ia, (m%8)*16 + (m/8)*sbo + (k/4)*lbo + (k%4)*4, A[m*K+k]); };

    // 1. everything K-major (the textbook case)
    fillA_K(128, 256); fillB_K(128, 256);
    run("A K-major, B K-major", ia, ib, Probe{128, 256, 128, 256, umma_idesc_tf32(128, 16, false, false), 0, 0, 0});
    run("A K-major, B K-major, lbo/sbo swapped", ia, ib, Probe{256, 128, 256, 128, umma_idesc_tf32(128, 16, false, false), 0, 0, 0});
    // 2. A MN-major (as k_panmix_tc)
    fillA_MN(128); fillB_K(128, 256);
    run("A MN-major sbo=128, B K-major", ia, ib, Probe{4096, 128, 128, 256, umma_idesc_tf32(128, 16, true, false), 0, 0, 0});
    run("A MN-major lbo=128 (swapped), B K", ia, ib, Probe{128, 4096, 128, 256, umma_idesc_tf32(128, 16, true, false), 0, 0, 0});
    run("A MN-major lbo=128 sbo=128", ia, ib, Probe{128, 128, 128, 256, umma_idesc_tf32(128, 16, true, false), 0, 0, 0});
    // 3. B MN-major
    fillA_K(128, 256); fillB_MN(128);
    run("A K-major, B MN-major sbo=128", ia, ib, Probe{128, 256, 4096, 128, umma_idesc_tf32(128, 16, false, true), 0, 0, 0});
    run("A K-major, B MN-major lbo=128", ia, ib, Probe{128, 256, 128, 4096, umma_idesc_tf32(128, 16, false, true), 0, 0, 0});
    return 0;
}
