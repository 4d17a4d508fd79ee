// Microbenchmark (GPU box only): sustained tcgen05.mma rate (cycles per M128 x N x K16 instruction) for the
// operand layouts the MLP kernels use.  One CTA per SM, one issuing thread, operands resident in shared memory.
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

// variant: 0: A K-major SW128, B K-major SW128     1: A K-major SW64, B MN-major SW128 (v2 hidden)
//          2: A K-major SW128, B K-major SW64 (v1)  3: A K-major SW64, B K-major SW128 (v2 encodings)
__global__ void __launch_bounds__(128, 1) rate_kernel(int variant, int N, int iters, int batch, unsigned long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t sbase = smem_u32(smem);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 196608);
    uint32_t* tptr = reinterpret_cast<uint32_t*>(bars + 4);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 196608 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // fp16 1.0
    if (tid == 0) { mbar_init(smem_u32(bars), 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(smem_u32(tptr), 512);
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(tptr);
    if (tid == 0) {
        const uint32_t a_base = sbase;            // 64 KB region for A
        const uint32_t b_base = sbase + 65536;    // 128 KB region for B
        uint32_t idesc = (1u << 4) | ((uint32_t(N) >> 3) << 17) | ((128u >> 4) << 24);
        if (variant == 1) idesc |= (1u << 16);
        uint32_t parity = 0;
        unsigned long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            for (int b = 0; b < batch; ++b) {
                const int k = (it * batch + b) & 15;       // cycle through 16 K-steps of a 256-deep operand
                uint64_t ad, bd;
                if (variant == 0) {
                    ad = umma_smem_desc(a_base + (k >> 2) * 16384 + (k & 3) * 32, 1024, SWZ_128B);
                    bd = umma_smem_desc(b_base + (k >> 2) * 32768 + (k & 3) * 32, 1024, SWZ_128B);
                } else if (variant == 1) {
                    ad = umma_smem_desc(a_base + (k >> 1) * 8192 + (k & 1) * 32, 512, SWZ_64B);
                    bd = desc_mn(b_base + 2 * k * 4096, 1024, 4096);
                } else if (variant == 2) {
                    ad = umma_smem_desc(a_base + (k >> 2) * 16384 + (k & 3) * 32, 1024, SWZ_128B);
                    bd = umma_smem_desc(b_base + (k >> 1) * 16384 + (k & 1) * 32, 512, SWZ_64B);
                } else {
                    ad = umma_smem_desc(a_base + (k >> 1) * 8192 + (k & 1) * 32, 512, SWZ_64B);
                    bd = umma_smem_desc(b_base + (k >> 2) * 32768 + (k & 3) * 32, 1024, SWZ_128B);
                }
                umma_ss(tmem + (b & 1) * 256, ad, bd, idesc, 1u);
            }
            umma_commit(smem_u32(bars));
            mbar_wait(smem_u32(bars), parity);
            parity ^= 1;
        }
        unsigned long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    unsigned long long* d_out;
    cudaMalloc(&d_out, sms * 8);
    const int smem = 196608 + 64 + 1024;
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const char* names[] = {"A K-major SW128 / B K-major SW128", "A K-major SW64  / B MN-major SW128", "A K-major SW128 / B K-major SW64 ",
                           "A K-major SW64  / B K-major SW128"};
    for (int grid : {1, sms})
        for (int variant = 0; variant < 4; ++variant)
            for (int N : {256, 128})
                for (int batch : {64, 8, 4}) {
                    const int iters = 4096 / batch;
                    rate_kernel<<<grid, 128, smem>>>(variant, N, iters, batch, d_out);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                    std::vector<unsigned long long> h(grid);
                    cudaMemcpy(h.data(), d_out, grid * 8, cudaMemcpyDeviceToHost);
                    unsigned long long mx = 0;
                    for (auto v : h) mx = v > mx ? v : mx;
                    printf("grid=%3d  %s  N=%3d  batch=%2d (commit+wait per batch): %.1f cycles/MMA\n", grid, names[variant], N, batch,
                           (double)mx / 4096.0);
                }
    return 0;
}
