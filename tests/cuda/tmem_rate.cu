// Microbenchmark (GPU box only): tcgen05.ld (TMEM -> registers) throughput with 4..16 reading warps, alone and
// while the tensor pipe runs back-to-back M128 x N256 x K16 MMAs; and the MMA rate under those reads.
#include <cuda_runtime.h>
#include <stdio.h>
#include <vector>

#include "../../vanilla-nerf-model-compression-using-lsa-enhanced-nncodec_b200/csrc/ptx_sm100.cuh"

using namespace nerfq;

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}

struct Args { int mma; int readers; int iters; int same_acc; unsigned long long* out; };

__global__ void __launch_bounds__(640, 1) tmem_rate_kernel(const Args a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t sbase = smem_u32(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    volatile uint32_t* flag = reinterpret_cast<volatile uint32_t*>(smem + 98304 + 64);
    uint32_t* tptr = reinterpret_cast<uint32_t*>(smem + 98304 + 128);
    const uint32_t bar0 = sbase + 98304;
    for (int i = tid; i < 98304 / 4; i += 640) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (tid == 0) { *flag = 0; mbar_init(bar0, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(smem_u32(tptr), 512);
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(tptr);
    unsigned long long cycles = 0, n_ld = 0;
    if (warp == 0) {
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
            const uint64_t a0 = umma_smem_desc(sbase, 1024, SWZ_128B);
            const uint64_t b0 = umma_smem_desc(sbase + 32768, 1024, SWZ_128B);
            const unsigned long long t0 = clock64();
            if (a.mma) {
                for (int it = 0; it < a.iters; ++it) {
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        umma_ss_c<1>(tmem, a0 + (((k >> 2) * 16384 + (k & 3) * 32) >> 4), b0 + (((k >> 2) * 32768 + (k & 3) * 32) >> 4), idesc);
                }
                umma_commit(bar0);
                mbar_wait(bar0, 0);
            } else {
                while (clock64() - t0 < 300000ull) {}
            }
            cycles = clock64() - t0;
            *flag = 1;
        }
    } else if (warp >= 4 && warp < 4 + a.readers) {
        const int e = warp - 4;
        // MMAs write columns 0..255; readers read the other accumulator (256..511) unless same_acc
        const uint32_t ta = tmem + (uint32_t((e & 3) * 32) << 16) + (a.same_acc ? 0u : 256u) + (e >> 2) * 64;
        uint32_t acc = 0;
        while (!*flag) {
#pragma unroll 1
            for (int r = 0; r < 8; ++r) {
                uint32_t v0[32], v1[32];
                tmem_ld32(ta, v0);
                tmem_ld32(ta + 32, v1);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) acc += v0[i] ^ v1[i];
            }
            n_ld += 16;
        }
        if (acc == 0x12345678u) a.out[0] = acc;
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
    if (lane == 0) {
        if (cycles) atomicAdd(a.out + 2 * blockIdx.x + 0, cycles);
        if (n_ld) atomicAdd(a.out + 2 * blockIdx.x + 1, n_ld);
    }
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    unsigned long long* d_out;
    cudaMalloc(&d_out, sms * 16);
    const int smem = 98304 + 256 + 1024;
    cudaFuncSetAttribute(tmem_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int mma : {0, 1})
        for (int same : {0, 1})
            for (int readers : {0, 4, 8, 16}) {
                if (!mma && (same || readers == 0)) continue;
                Args a{mma, readers, 1024, same, d_out};
                cudaMemset(d_out, 0, sms * 16);
                tmem_rate_kernel<<<sms, 640, smem>>>(a);
                cudaError_t err = cudaGetLastError();
                if (err == cudaSuccess) err = cudaDeviceSynchronize();
                if (err != cudaSuccess) { printf("error %s\n", cudaGetErrorString(err)); return 1; }
                std::vector<unsigned long long> h(2 * sms);
                cudaMemcpy(h.data(), d_out, sms * 16, cudaMemcpyDeviceToHost);
                unsigned long long mx = 0, ld = 0;
                for (int i = 0; i < sms; ++i) { if (h[2 * i] > mx) mx = h[2 * i]; ld += h[2 * i + 1]; }
                printf("mma=%d same_accumulator=%d readers=%2d: %7.1f cyc/MMA   tcgen05.ld %7.1f B/clk/SM (%.0f clk per x32 load per warp)\n", mma, same, readers,
                       mma ? (double)mx / (1024.0 * 8) : 0.0, (double)ld * 4096.0 / sms / (double)mx,
                       readers ? (double)mx / ((double)ld / sms / readers) : 0.0);
            }
    return 0;
}
