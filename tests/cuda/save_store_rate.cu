// Microbenchmark (GPU box only): how fast can 148 persistent CTAs write the forward kernel's saved-activation pattern when
// they do nothing else?  Same addresses as mlp3_fwd.cu with `save`: per group of 256 points ten 128 KB slots laid out
// [16 point chunks][256 channels][16 points] fp16, written as warp-wide 32-byte-per-lane stores (1 KB contiguous per
// instruction), 4864 B per point; group g is owned by CTA g % gridDim.  Prints GB/s for a few store policies and warp counts.
// If this runs far above the ~4.0 TB/s the forward+save kernel reaches, the pattern is not what limits that kernel.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

constexpr size_t kSlotBytes = 16 * 256 * 32;
constexpr size_t kGroupBytes = 10 * kSlotBytes;

template <int POLICY>
__device__ __forceinline__ void st32(void* p, uint32_t v) {
    if (POLICY == 0)
        asm volatile("st.global.L2::evict_first.v8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"l"(p), "r"(v) : "memory");
    else
        asm volatile("st.global.v8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"l"(p), "r"(v) : "memory");
}

template <int POLICY>
__global__ void __launch_bounds__(1024, 1) store_kernel(uint8_t* save, int n_groups) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
    for (int g = blockIdx.x; g < n_groups; g += gridDim.x) {
        uint8_t* base = save + (size_t)g * kGroupBytes;
        // 10 slots x 16 point chunks x 8 channel groups of 32 (the views slot, 9, has 4)
        for (int it = warp; it < 10 * 16 * 8; it += n_warps) {
            const int slot = it / 128, pc = (it >> 3) & 15, cg = it & 7;
            if (slot == 9 && cg >= 4) continue;
            st32<POLICY>(base + (size_t)slot * kSlotBytes + ((size_t)(pc * 256 + cg * 32 + lane)) * 32, (uint32_t)it);
        }
    }
}

int main() {
    const int n_groups = 4096 * 192 / 256;
    uint8_t* buf;
    cudaMalloc(&buf, (size_t)n_groups * kGroupBytes);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const double bytes = (double)n_groups * 256 * 4864;
    for (int policy = 0; policy < 2; ++policy)
        for (int warps : {4, 8, 16, 32}) {
            for (int rep = 0; rep < 3; ++rep) {
                cudaEventRecord(e0);
                if (policy == 0) store_kernel<0><<<148, 32 * warps>>>(buf, n_groups);
                else store_kernel<1><<<148, 32 * warps>>>(buf, n_groups);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
            }
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            printf("policy %s, %2d warps/CTA, 148 CTAs: %.3f ms for %.2f GB -> %.0f GB/s (%.1f B/clk/SM at 1.9 GHz)\n",
                   policy == 0 ? "evict_first" : "default    ", warps, ms, bytes / 1e9, bytes / ms / 1e6, bytes / ms / 1e6 / 148 / 1.9);
        }
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
