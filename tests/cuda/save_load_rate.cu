// Microbenchmark (GPU box only): how fast can 148 persistent CTAs READ the saved activations in the backward kernel's pattern
// when they do nothing else?  Same addresses as mlp3_bwd.cu: per group of 256 points the views job (slot 9, 128 channels) and
// 18 dgrad jobs (slots 8..0, low / high channel half); a thread owns one channel and 64 points = 4 chunks of 32 bytes, a warp
// reads 1 KB contiguous per chunk, the four warps of a point quarter 4 KB; 16 warps per CTA like the kernel's epilogue.
// Modes: no L2 prefetch / one prefetch per 128-byte line two jobs ahead (the kernel's default) / both 64-byte halves.
// If this runs far above the ~4.3 TB/s the backward reaches, the pattern is not what limits that kernel.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

constexpr size_t kSlotBytes = 16 * 256 * 32;
constexpr size_t kGroupBytes = 10 * kSlotBytes;

__device__ __forceinline__ uint32_t ld32(const void* p) {
    uint32_t a, b, c, d, e, f, g, h;
    asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p));
    return a ^ b ^ c ^ d ^ e ^ f ^ g ^ h;
}
__device__ __forceinline__ size_t job_base(int v, int cl) {      // v: 0 = views, 1..18 = dgrad jobs
    const int slot = v == 0 ? 9 : 8 - ((v - 1) >> 1);
    const int ch = v == 0 ? cl : 128 * ((v - 1) & 1) + cl;
    return (size_t)slot * kSlotBytes + (size_t)ch * 32;
}

template <int MODE, int DEPTH>
__global__ void __launch_bounds__(512, 1) load_kernel(const uint8_t* save, int n_groups, uint32_t* sink) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = warp & 3, pq = warp >> 2, cl = 32 * q + lane;
    uint32_t acc = 0;
    for (int g = blockIdx.x; g < n_groups; g += gridDim.x) {
        const uint8_t* base = save + (size_t)g * kGroupBytes;
        for (int v = 0; v < 19; ++v) {
            if (MODE >= 1 && v + 2 < 19) {      // the line of this warp's chunk lane/8 that holds channels 4*(lane%8)..+3
                const uint8_t* line = base + job_base(v + 2, 32 * q + 4 * (lane & 7)) + (size_t)(4 * pq + (lane >> 3)) * 8192;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(line) : "memory");
                if (MODE >= 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(line + 64) : "memory");
            }
            const uint8_t* row = base + job_base(v, cl) + (size_t)(4 * pq) * 8192;
            if (DEPTH == 4) {
                const uint32_t a = ld32(row), b = ld32(row + 8192), c = ld32(row + 16384), d = ld32(row + 24576);
                acc ^= a ^ b ^ c ^ d;
            } else {            // two in flight
                uint32_t a = ld32(row), b = ld32(row + 8192);
                acc ^= a;
                a = ld32(row + 16384);
                acc ^= b;
                b = ld32(row + 24576);
                acc ^= a ^ b;
            }
        }
    }
    if (acc == 0x12345678u) sink[0] = acc;
}

int main() {
    const int n_groups = 4096 * 192 / 256;
    uint8_t* buf;
    uint32_t* sink;
    cudaMalloc(&buf, (size_t)n_groups * kGroupBytes);
    cudaMalloc(&sink, 4);
    cudaMemset(buf, 1, (size_t)n_groups * kGroupBytes);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const double bytes = (double)n_groups * 256 * 4864;
    auto run = [&](int mode, int depth) {
        float ms = 0;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0 && depth == 2) load_kernel<0, 2><<<148, 512>>>(buf, n_groups, sink);
            if (mode == 1 && depth == 2) load_kernel<1, 2><<<148, 512>>>(buf, n_groups, sink);
            if (mode == 2 && depth == 2) load_kernel<2, 2><<<148, 512>>>(buf, n_groups, sink);
            if (mode == 0 && depth == 4) load_kernel<0, 4><<<148, 512>>>(buf, n_groups, sink);
            if (mode == 1 && depth == 4) load_kernel<1, 4><<<148, 512>>>(buf, n_groups, sink);
            if (mode == 2 && depth == 4) load_kernel<2, 4><<<148, 512>>>(buf, n_groups, sink);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
        }
        printf("prefetch mode %d, %d loads in flight per thread, 148 CTAs x 16 warps: %.3f ms for %.2f GB -> %.0f GB/s\n", mode, depth, ms,
               bytes / 1e9, bytes / ms / 1e6);
    };
    for (int depth : {2, 4})
        for (int mode = 0; mode < 3; ++mode) run(mode, depth);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
