// Probe (GPU box only): tcgen05.mma with an MN-major B operand (activations stored [channel][point], points
// contiguous) and a K-major A operand (weight stage, SWIZZLE_64B), i.e. the "channels on TMEM lanes"
// formulation D[o][n] = sum_k W[o][k] * X[n][k].  Checks descriptor conventions against an exact host GEMM.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../../vanilla-nerf-model-compression-using-lsa-enhanced-nncodec_b200/csrc/ptx_sm100.cuh"

using namespace nerfq;

__device__ __forceinline__ uint64_t desc_mn(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)SWZ_128B << 61;
    return d;
}

// W: [128][K] halves (row-major), X: [N][K] halves (row-major, logical), D: [128][N] float
__global__ void __launch_bounds__(128, 1) probe(const uint16_t* W, const uint16_t* X, float* D, int N, int K, int swap) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* a_s = smem;                           // K/32 stages of 128 rows x 64 B (SW64)
    uint8_t* b_s = smem + (K / 32) * 8192;         // MN-major: kg groups (8 k) x ng groups (64 n) atoms of 1024 B
    const int ngroups = N / 64;
    const uint32_t LBO = 1024, SBO = 1024 * ngroups;
    uint64_t* bars = reinterpret_cast<uint64_t*>(b_s + (K / 8) * SBO);
    uint32_t* tptr = reinterpret_cast<uint32_t*>(bars + 2);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int st = 0; st < K / 32; ++st)
        for (int ch = 0; ch < 4; ++ch) {
            uint4 q = *reinterpret_cast<const uint4*>(W + (size_t)tid * K + st * 32 + ch * 8);
            *reinterpret_cast<uint4*>(a_s + st * 8192 + sw64_offset(tid, ch)) = q;
        }
    // B: element (n, k) -> 2 bytes; written element-wise (slow, probe only)
    for (int idx = tid; idx < N * K; idx += 128) {
        const int n = idx / K, k = idx % K;
        const uint32_t off = (k / 8) * SBO + (n / 64) * LBO + (k % 8) * 128 + ((((n % 64) / 8) ^ (k % 8)) << 4) + (n % 8) * 2;
        *reinterpret_cast<uint16_t*>(b_s + off) = X[idx];
    }
    if (tid == 0) { mbar_init(smem_u32(bars), 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(smem_u32(tptr), 256);
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(tptr);
    if (tid == 0) {
        const uint32_t idesc = umma_idesc(128, N, false) | (1u << 16);   // B is MN-major
        for (int ks = 0; ks < K / 16; ++ks) {
            const uint32_t a_addr = smem_u32(a_s) + (ks >> 1) * 8192 + (ks & 1) * 32;
            const uint32_t b_addr = smem_u32(b_s) + 2 * ks * SBO;
            const uint64_t bd = swap ? desc_mn(b_addr, SBO, LBO) : desc_mn(b_addr, LBO, SBO);
            umma_ss(tmem, umma_smem_desc(a_addr, 512, SWZ_64B), bd, idesc, ks ? 1u : 0u);
        }
        umma_commit(smem_u32(bars));
    }
    mbar_wait(smem_u32(bars), 0);
    tc_fence_after_sync();
    for (int c = 0; c < N / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem + (uint32_t(warp * 32) << 16) + c * 32, v);
        tmem_ld_wait();
        for (int i = 0; i < 32; ++i) D[(size_t)tid * N + c * 32 + i] = __uint_as_float(v[i]);
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

static uint16_t f2h(float f) { __half h = __float2half(f); return *reinterpret_cast<uint16_t*>(&h); }

int main() {
    int fails = 0;
    const int cfgs[][2] = {{256, 64}, {256, 256}, {128, 128}, {64, 32}};
    for (auto& c : cfgs)
        for (int swap = 0; swap < 2; ++swap) {
            const int N = c[0], K = c[1];
            std::vector<float> Wf(128 * K), Xf(N * K), Dref(128 * N), Dh(128 * N);
            std::vector<uint16_t> Wh(128 * K), Xh(N * K);
            srand(7 + N + K);
            for (size_t i = 0; i < Wf.size(); ++i) { Wf[i] = float(rand() % 9 - 4); Wh[i] = f2h(Wf[i]); }
            for (size_t i = 0; i < Xf.size(); ++i) { Xf[i] = float(rand() % 7 - 3); Xh[i] = f2h(Xf[i]); }
            for (int o = 0; o < 128; ++o)
                for (int n = 0; n < N; ++n) {
                    float s = 0;
                    for (int k = 0; k < K; ++k) s += Wf[o * K + k] * Xf[n * K + k];
                    Dref[o * N + n] = s;
                }
            uint16_t *dW, *dX; float* dD;
            cudaMalloc(&dW, Wh.size() * 2); cudaMalloc(&dX, Xh.size() * 2); cudaMalloc(&dD, Dh.size() * 4);
            cudaMemcpy(dW, Wh.data(), Wh.size() * 2, cudaMemcpyHostToDevice);
            cudaMemcpy(dX, Xh.data(), Xh.size() * 2, cudaMemcpyHostToDevice);
            cudaMemset(dD, 0xff, Dh.size() * 4);
            const int smem = (K / 32) * 8192 + (K / 8) * 1024 * (N / 64) + 64 + 1024;
            cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            probe<<<1, 128, smem>>>(dW, dX, dD, N, K, swap);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("N=%d K=%d swap=%d: CUDA error %s\n", N, K, swap, cudaGetErrorString(e)); return 2; }
            cudaMemcpy(Dh.data(), dD, Dh.size() * 4, cudaMemcpyDeviceToHost);
            int bad = 0; double maxerr = 0;
            for (size_t i = 0; i < Dh.size(); ++i) { double d = fabs(Dh[i] - Dref[i]); if (!(d <= 1e-3)) ++bad; if (d > maxerr) maxerr = d; }
            printf("MN-major B: N=%d K=%d lbo/sbo %s: %s maxerr=%g bad=%d/%zu D[1][0..3]=%g %g %g %g ref=%g %g %g %g\n", N, K,
                   swap ? "swapped" : "as designed", bad ? "FAIL" : "PASS", maxerr, bad, Dh.size(), Dh[N], Dh[N + 1], Dh[N + 2], Dh[N + 3],
                   Dref[N], Dref[N + 1], Dref[N + 2], Dref[N + 3]);
            if (!swap) fails += bad != 0;
            cudaFree(dW); cudaFree(dX); cudaFree(dD);
        }
    return fails ? 1 : 0;
}
