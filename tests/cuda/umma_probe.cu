// Standalone probe (GPU box only): checks the tcgen05 operand-layout assumptions the MLP kernels rely
// on -- SWIZZLE_128B activation tiles written by threads, SWIZZLE_64B weight stages, K-advance by
// +32 bytes inside a swizzled row, the 32x32b TMEM load mapping -- against an exact integer GEMM on the
// host.  Prints one PASS/FAIL line per configuration.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../../vanilla-nerf-model-compression-using-lsa-enhanced-nncodec_b200/csrc/ptx_sm100.cuh"

using namespace nerfq;

struct Cfg {
    int n;         // MMA N
    int kblocks;   // number of 64-wide A blocks (K = 64 * kblocks)
    int bf16;
};

// A: [128][K] halves logical, B: [n][K] halves logical; D: [128][n] float
__global__ void __launch_bounds__(128, 1) probe_kernel(const uint16_t* A, const uint16_t* B, float* D, Cfg cfg) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t sbase = smem_u32(smem);
    const int K = 64 * cfg.kblocks;
    uint8_t* a_s = smem;                                  // kblocks x 16 KB
    uint8_t* b_s = smem + cfg.kblocks * 16384;            // (K/32) stages x n x 64 B
    uint64_t* bars = reinterpret_cast<uint64_t*>(b_s + (K / 32) * cfg.n * 64);
    uint32_t* tptr = reinterpret_cast<uint32_t*>(bars + 2);
    const int tid = threadIdx.x, warp = tid >> 5;
    // A: thread = row
    for (int kb = 0; kb < cfg.kblocks; ++kb)
        for (int ch = 0; ch < 8; ++ch) {
            uint4 q = *reinterpret_cast<const uint4*>(A + (size_t)tid * K + kb * 64 + ch * 8);
            *reinterpret_cast<uint4*>(a_s + kb * 16384 + sw128_offset(tid, ch)) = q;
        }
    for (int st = 0; st < K / 32; ++st)
        for (int item = tid; item < cfg.n * 4; item += 128) {
            const int n = item >> 2, ch = item & 3;
            uint4 q = *reinterpret_cast<const uint4*>(B + (size_t)n * K + st * 32 + ch * 8);
            *reinterpret_cast<uint4*>(b_s + st * cfg.n * 64 + sw64_offset(n, ch)) = q;
        }
    if (tid == 0) { mbar_init(smem_u32(bars), 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(smem_u32(tptr), 256);
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(tptr);
    if (tid == 0) {
        const uint32_t idesc = umma_idesc(128, cfg.n, cfg.bf16 != 0);
        for (int st = 0; st < K / 32; ++st) {
            const uint32_t a_addr = sbase + (st >> 1) * 16384 + (st & 1) * 64;
            const uint32_t b_addr = smem_u32(b_s) + st * cfg.n * 64;
            for (int j = 0; j < 2; ++j)
                umma_ss(tmem, umma_smem_desc(a_addr + j * 32, 1024, SWZ_128B), umma_smem_desc(b_addr + j * 32, 512, SWZ_64B),
                        idesc, (st | j) ? 1u : 0u);
        }
        umma_commit(smem_u32(bars));
    }
    mbar_wait(smem_u32(bars), 0);
    tc_fence_after_sync();
    for (int c = 0; c < cfg.n / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem + (uint32_t(warp * 32) << 16) + c * 32, v);
        tmem_ld_wait();
        for (int i = 0; i < 32; ++i) D[(size_t)tid * cfg.n + c * 32 + i] = __uint_as_float(v[i]);
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

static uint16_t f2h(float f, int bf16) {
    if (bf16) { __nv_bfloat16 h = __float2bfloat16(f); return *reinterpret_cast<uint16_t*>(&h); }
    __half h = __float2half(f); return *reinterpret_cast<uint16_t*>(&h);
}

int main() {
    Cfg cfgs[] = {{256, 1, 0}, {256, 4, 0}, {128, 4, 0}, {256, 4, 1}, {128, 2, 1}};
    int fails = 0;
    for (const Cfg& cfg : cfgs) {
        const int K = 64 * cfg.kblocks;
        std::vector<float> Af(128 * K), Bf(cfg.n * K), Dref(128 * cfg.n, 0.f), Dh(128 * cfg.n);
        std::vector<uint16_t> Ah(128 * K), Bh(cfg.n * K);
        srand(1234 + K + cfg.n);
        for (size_t i = 0; i < Af.size(); ++i) { Af[i] = float(rand() % 9 - 4); Ah[i] = f2h(Af[i], cfg.bf16); }
        for (size_t i = 0; i < Bf.size(); ++i) { Bf[i] = float(rand() % 7 - 3); Bh[i] = f2h(Bf[i], cfg.bf16); }
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < cfg.n; ++n) {
                float s = 0;
                for (int k = 0; k < K; ++k) s += Af[m * K + k] * Bf[n * K + k];
                Dref[m * cfg.n + n] = s;
            }
        uint16_t *dA, *dB; float* dD;
        cudaMalloc(&dA, Ah.size() * 2); cudaMalloc(&dB, Bh.size() * 2); cudaMalloc(&dD, Dh.size() * 4);
        cudaMemcpy(dA, Ah.data(), Ah.size() * 2, cudaMemcpyHostToDevice);
        cudaMemcpy(dB, Bh.data(), Bh.size() * 2, cudaMemcpyHostToDevice);
        cudaMemset(dD, 0xff, Dh.size() * 4);
        const int smem = cfg.kblocks * 16384 + (K / 32) * cfg.n * 64 + 64 + 1024;
        cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        probe_kernel<<<1, 128, smem>>>(dA, dB, dD, cfg);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("cfg n=%d K=%d bf16=%d: CUDA error %s\n", cfg.n, K, cfg.bf16, cudaGetErrorString(e)); return 2; }
        cudaMemcpy(Dh.data(), dD, Dh.size() * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0; int bad = 0;
        for (size_t i = 0; i < Dh.size(); ++i) { double d = fabs(Dh[i] - Dref[i]); if (!(d <= 1e-3)) { ++bad; } if (d > maxerr) maxerr = d; }
        printf("cfg n=%d K=%d bf16=%d: %s  maxerr=%g bad=%d/%zu  D[0][0..3]=%g %g %g %g ref=%g %g %g %g  D[37][5]=%g ref=%g\n", cfg.n, K,
               cfg.bf16, bad ? "FAIL" : "PASS", maxerr, bad, Dh.size(), Dh[0], Dh[1], Dh[2], Dh[3], Dref[0], Dref[1], Dref[2], Dref[3],
               Dh[37 * cfg.n + 5], Dref[37 * cfg.n + 5]);
        fails += bad != 0;
        cudaFree(dA); cudaFree(dB); cudaFree(dD);
    }
    return fails ? 1 : 0;
}
