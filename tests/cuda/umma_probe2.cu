// Standalone probe (GPU box only) for the CTA-pair (cta_group::2) building blocks: which rows of A / columns of B / rows
// of D each CTA of the pair owns, the MN-major SWIZZLE_128B activation tile with 128 points per CTA, operand rows
// written into the PEER's shared memory (st.shared::cluster) and handed to the leader's MMA thread through a
// cluster-scope mbarrier, and the multicast commit.  Exact integer GEMM against the host; prints PASS/FAIL.
//   D[o][n] = sum_k A[o][k] * B[n][k],  o < 256, n < 256, k < 64
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../../vanilla-nerf-model-compression-using-lsa-enhanced-nncodec_b200/csrc/ptx_sm100.cuh"

using namespace nerfq;

constexpr uint32_t kKGroup = 2048, kNGroup = 1024;     // MN-major tile of [K][128 points]
__device__ __forceinline__ uint32_t act_off(uint32_t k, uint32_t n) {
    return (k >> 3) * kKGroup + (n >> 6) * kNGroup + (k & 7u) * 128u + ((((n & 63u) >> 3) ^ (k & 7u)) << 4) + (n & 7u) * 2u;
}
__device__ __forceinline__ uint64_t desc_mn(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((kNGroup >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((kKGroup >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)SWZ_128B << 61;
    return d;
}

// remote != 0: every CTA writes the operand rows of the OTHER CTA's B tile through DSMEM
__global__ void __launch_bounds__(128, 1) probe2_kernel(const uint16_t* A, const uint16_t* B, float* D, int remote) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const uint32_t rank = cluster_ctarank();
    const int tid = threadIdx.x, warp = tid >> 5;
    constexpr uint32_t kA = 0, kB = 16384, kBars = 32768, kTptr = kBars + 64;
    auto bar = [&](int i) { return sbase + kBars + 8u * i; };
    // A: two SWIZZLE_64B stages of [128 rows][32 k]; this CTA's rows are 128*rank ..
    for (int st = 0; st < 2; ++st)
        for (int ch = 0; ch < 4; ++ch) {
            const uint4 q = *reinterpret_cast<const uint4*>(A + (size_t)(128 * rank + tid) * 64 + st * 32 + ch * 8);
            *reinterpret_cast<uint4*>(smem + kA + st * 8192 + sw64_offset(tid, ch)) = q;
        }
    if (tid == 0) {
        mbar_init(bar(0), 1);          // accumulator ready (multicast commit)
        mbar_init(bar(1), 8);          // operand ready: 4 warps x 2 CTAs (leader's copy is the one used)
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc2(sbase + kTptr, 256);
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after_sync();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + kTptr);
    // B tile of CTA `dst`: points 128*dst .. ; thread = point, writes its 64 channels (2 bytes each: slow, fine for a probe)
    {
        const uint32_t dst = remote ? rank ^ 1u : rank;
        const uint32_t base = mapa_u32(sbase + kB, dst);
        for (int k = 0; k < 64; ++k) {
            const uint16_t v = B[(size_t)(128 * dst + tid) * 64 + k];
            asm volatile("st.shared::cluster.u16 [%0], %1;" ::"r"(base + act_off(k, tid)), "h"(v) : "memory");
        }
        fence_proxy_async_all();
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive_cluster(mapa_u32(bar(1), 0));
    }
    if (rank == 0 && warp == 0) {
        mbar_wait_cluster(bar(1), 0);
        tc_fence_after_sync();
        if (elect_one()) {
            const uint32_t idesc = umma_idesc(256, 256, false) | (1u << 16);        // B MN-major
            const uint64_t a0 = umma_smem_desc(sbase + kA, 512, SWZ_64B), b0 = desc_mn(sbase + kB);
            for (int j = 0; j < 4; ++j)       // K = 16 each: A +32 B inside a stage / +8 KB per stage; B +2 k-groups
                umma_ss2(tmem, a0 + (((j >> 1) * 8192 + (j & 1) * 32) >> 4), b0 + ((j * 2 * kKGroup) >> 4), idesc, j ? 1u : 0u);
            umma_commit2_mc(bar(0), 3);
        }
        __syncwarp();
    }
    mbar_wait(bar(0), 0);
    tc_fence_after_sync();
    for (int c = 0; c < 256; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
        tmem_ld_wait();
        for (int i = 0; i < 32; ++i) D[(size_t)(128 * rank + tid) * 256 + c + i] = __uint_as_float(v[i]);
    }
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) tmem_dealloc2(tmem, 256);
}

// Second configuration ("points on lanes"): A = activations [256 points][64 k], K-major SWIZZLE_128B, each CTA holds ITS
// 128 rows; B = weights [256 channels][64 k] as two K-major SWIZZLE_64B stages of [128 rows][32 k], each CTA holds ITS
// 128 rows.  Expected: CTA r receives D rows 128 r .. (its points) x all 256 columns, column n = weight row n.
__global__ void __launch_bounds__(128, 1) probe2k_kernel(const uint16_t* A, const uint16_t* B, float* D) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const uint32_t rank = cluster_ctarank();
    const int tid = threadIdx.x, warp = tid >> 5;
    constexpr uint32_t kA = 0, kB = 16384, kBars = 32768, kTptr = kBars + 64;
    auto bar = [&](int i) { return sbase + kBars + 8u * i; };
    for (int ch = 0; ch < 8; ++ch) {          // A: one SWIZZLE_128B block [128 rows][128 B]
        const uint4 q = *reinterpret_cast<const uint4*>(A + (size_t)(128 * rank + tid) * 64 + ch * 8);
        *reinterpret_cast<uint4*>(smem + kA + sw128_offset(tid, ch)) = q;
    }
    for (int st = 0; st < 2; ++st)
        for (int ch = 0; ch < 4; ++ch) {
            const uint4 q = *reinterpret_cast<const uint4*>(B + (size_t)(128 * rank + tid) * 64 + st * 32 + ch * 8);
            *reinterpret_cast<uint4*>(smem + kB + st * 8192 + sw64_offset(tid, ch)) = q;
        }
    if (tid == 0) { mbar_init(bar(0), 1); mbar_init(bar(1), 8); mbar_fence_init(); }
    if (warp == 0) tmem_alloc2(sbase + kTptr, 256);
    fence_proxy_async_all();
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after_sync();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + kTptr);
    if (rank == 0 && warp == 0) {
        if (elect_one()) {
            const uint32_t idesc = umma_idesc(256, 256, false);
            const uint64_t a0 = umma_smem_desc(sbase + kA, 1024, SWZ_128B), b0 = umma_smem_desc(sbase + kB, 512, SWZ_64B);
            for (int j = 0; j < 4; ++j)
                umma_ss2(tmem, a0 + ((j * 32) >> 4), b0 + (((j >> 1) * 8192 + (j & 1) * 32) >> 4), idesc, j ? 1u : 0u);
            umma_commit2_mc(bar(0), 3);
        }
        __syncwarp();
    }
    mbar_wait(bar(0), 0);
    tc_fence_after_sync();
    for (int c = 0; c < 256; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
        tmem_ld_wait();
        for (int i = 0; i < 32; ++i) D[(size_t)(128 * rank + tid) * 256 + c + i] = __uint_as_float(v[i]);
    }
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) tmem_dealloc2(tmem, 256);
}

static uint16_t h16(float f) { __half h = __float2half(f); uint16_t u; memcpy(&u, &h, 2); return u; }

int main() {
    std::vector<uint16_t> A(256 * 64), B(256 * 64);
    std::vector<float> Af(256 * 64), Bf(256 * 64);
    srand(1);
    for (int i = 0; i < 256 * 64; ++i) { Af[i] = (float)(rand() % 9 - 4); Bf[i] = (float)(rand() % 7 - 3); A[i] = h16(Af[i]); B[i] = h16(Bf[i]); }
    uint16_t *dA, *dB; float* dD;
    cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dD, 256 * 256 * 4);
    cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
    const int smem = 34816;
    cudaFuncSetAttribute(probe2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int fails = 0;
    for (int remote = 0; remote < 2; ++remote) {
        cudaMemset(dD, 0, 256 * 256 * 4);
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(2); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaError_t err = cudaLaunchKernelEx(&cfg, probe2_kernel, (const uint16_t*)dA, (const uint16_t*)dB, dD, remote);
        if (err == cudaSuccess) err = cudaDeviceSynchronize();
        if (err != cudaSuccess) { printf("remote=%d: CUDA error %s\n", remote, cudaGetErrorString(err)); return 1; }
        std::vector<float> D(256 * 256);
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0; int first = -1;
        for (int o = 0; o < 256; ++o)
            for (int n = 0; n < 256; ++n) {
                float ref = 0;
                for (int k = 0; k < 64; ++k) ref += Af[o * 64 + k] * Bf[n * 64 + k];
                if (D[o * 256 + n] != ref) { if (first < 0) first = o * 256 + n; ++bad; }
            }
        printf("cta_group::2 M=256 N=256 K=64, operand tile written %s: %s (%d mismatches%s)\n", remote ? "by the peer CTA (DSMEM)" : "locally",
               bad ? "FAIL" : "PASS", bad, bad ? "" : "");
        if (bad) { printf("  first mismatch at o=%d n=%d: got %g\n", first / 256, first % 256, D[first]); ++fails; }
    }
    {
        cudaFuncSetAttribute(probe2k_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaMemset(dD, 0, 256 * 256 * 4);
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(2); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaError_t err = cudaLaunchKernelEx(&cfg, probe2k_kernel, (const uint16_t*)dA, (const uint16_t*)dB, dD);
        if (err == cudaSuccess) err = cudaDeviceSynchronize();
        if (err != cudaSuccess) { printf("K-major/K-major: CUDA error %s\n", cudaGetErrorString(err)); return 1; }
        std::vector<float> D(256 * 256);
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0, first = -1;
        for (int m = 0; m < 256; ++m)
            for (int n = 0; n < 256; ++n) {
                float ref = 0;
                for (int k = 0; k < 64; ++k) ref += Af[m * 64 + k] * Bf[n * 64 + k];
                if (D[m * 256 + n] != ref) { if (first < 0) first = m * 256 + n; ++bad; }
            }
        printf("cta_group::2 M=256 N=256 K=64, A K-major SW128 (points) x B K-major SW64 (weights): %s (%d mismatches)\n", bad ? "FAIL" : "PASS", bad);
        if (bad) { printf("  first mismatch at m=%d n=%d: got %g\n", first / 256, first % 256, D[first]); ++fails; }
    }
    return fails ? 1 : 0;
}
