// Microbenchmark (GPU box only): what bounds a fused-MLP CTA on sm_100a?
//   - the tcgen05.mma floor for the operand sources/layouts the MLP kernels can use
//     (A from shared memory or TMEM, B K-major or MN-major, 1-CTA M=128 and 2-CTA M=256), issued
//     from a fully unrolled loop so that the single issuing thread is not the limit;
//   - how that rate changes while other warps stream st.shared traffic (the epilogue writing operand
//     tiles) and while a bulk-copy ring pulls weight stages from L2 into shared memory;
//   - the st.shared and L2->smem bulk-copy rates on their own.
// Prints one line per experiment: cycles per MMA, st.shared bytes/clk/SM, bulk-copy bytes/clk/SM.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../../vanilla-nerf-model-compression-using-lsa-enhanced-nncodec_b200/csrc/ptx_sm100.cuh"

using namespace nerfq;

constexpr uint32_t kA = 0;                    // 32 KB : A operand, 128 rows x K=128 (2 SW128 blocks)
constexpr uint32_t kB = 32768;                // 64 KB : B operand, 256 rows x K=128 (or MN-major 128 x 256)
constexpr uint32_t kRing = 98304;             // 96 KB : bulk-copy slots
constexpr uint32_t kSts = 196608;             // 16 KB : st.shared targets, 4 x 512 B per warp
constexpr uint32_t kBars = 212992;            // barriers, flag, tmem pointer
constexpr uint32_t kSmem = kBars + 256 + 1024;

struct Args {
    int mode;        // -1: no MMA (spin), 0: SS K/K, 1: SS B MN-major, 2: TS (A in TMEM), 3: 2-CTA SS (M=256)
    int n;           // MMA N
    int iters;       // x16 MMAs
    int sts_warps;   // 0..8
    int bulk;        // issuing warps (0..3)
    int lanes;       // issuing lanes per warp
    int slots;       // slots per issuer
    int chunk;       // bulk copy bytes; bulk * lanes * slots * chunk <= 96 KB
    const uint8_t* src;          // >= 2 MB, L2 resident
    uint8_t* dst;                // bulk-store target: 3 MB per CTA
    unsigned long long* out;     // per CTA: cycles, sts count, bulk count
};

__device__ __forceinline__ uint64_t desc_mn(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)SWZ_128B << 61;
    return d;
}
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc) : "memory");
}
__device__ __forceinline__ void umma_ss2(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc) : "memory");
}
__device__ __forceinline__ void umma_commit2(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}

template <int kMode>
__device__ __forceinline__ void issue16(uint32_t tmem, uint64_t a0, uint64_t b0, uint32_t idesc, uint32_t cbar = 0) {
    // 8 distinct K=16 steps of the K=128 operands, twice
#pragma unroll
    for (int r = 0; r < 2; ++r) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (kMode == 0) {
                // A: block (k>>2) of 16 KB, +32 B per step; B (256 rows): block of 32 KB
                umma_ss_c<1>(tmem, a0 + (((k >> 2) * 16384 + (k & 3) * 32) >> 4), b0 + (((k >> 2) * 32768 + (k & 3) * 32) >> 4), idesc);
            } else if (kMode == 1) {
                // A: SW128 K-major as above; B: MN-major tile, 2 K-groups of 4 KB per step
                umma_ss_c<1>(tmem, a0 + (((k >> 2) * 16384 + (k & 3) * 32) >> 4), b0 + ((k * 8192) >> 4), idesc);
            } else if (kMode == 5) {
                // as mode 1, plus a commit after every 4 MMAs (what the MLP kernels do per 16 KB weight chunk)
                umma_ss_c<1>(tmem, a0 + (((k >> 2) * 16384 + (k & 3) * 32) >> 4), b0 + ((k * 8192) >> 4), idesc);
                if ((k & 3) == 3) umma_commit(cbar + 8 * (k >> 2));
            } else if (kMode == 4) {
                // A: SWIZZLE_64B K-major stages of [128 rows x 32 k] (8 KB), two K=16 steps per stage; B: MN-major tile
                umma_ss_c<1>(tmem, a0 + (((k >> 1) * 8192 + (k & 1) * 32) >> 4), b0 + ((k * 8192) >> 4), idesc);
            } else if (kMode == 2) {
                umma_ts(tmem, (uint32_t)a0 + k * 8, b0 + (((k >> 2) * 32768 + (k & 3) * 32) >> 4), idesc);
            } else {
                // 2-CTA: each CTA supplies 128 rows of A and 128 rows of B (N/2)
                umma_ss2(tmem, a0 + (((k >> 2) * 16384 + (k & 3) * 32) >> 4), b0 + (((k >> 2) * 16384 + (k & 3) * 32) >> 4), idesc);
            }
        }
    }
}

template <bool kCluster>
__global__ void __launch_bounds__(384, 1) rate2_kernel(const Args a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t sbase = smem_u32(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    volatile uint32_t* flag = reinterpret_cast<volatile uint32_t*>(smem + kBars + 224);
    uint32_t* tptr = reinterpret_cast<uint32_t*>(smem + kBars + 240);
    auto bar = [&](int i) { return sbase + kBars + 8u * i; };

    for (int i = tid; i < (int)(kSts / 4); i += 384) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // fp16 1.0
    if (tid == 0) {
        *flag = 0;
        mbar_init(bar(0), 1);
        for (int s = 0; s < 24; ++s) mbar_init(bar(1 + s), 1);
        mbar_fence_init();
    }
    const uint32_t rank = kCluster ? cluster_rank() : 0;
    if (warp == 0) {
        if (kCluster) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tptr)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            tmem_alloc(smem_u32(tptr), 512);
        }
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    if (kCluster) cluster_sync_all();
    tc_fence_after_sync();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(tptr);

    unsigned long long cycles = 0, n_sts = 0, n_bulk = 0;
    if (warp == 0) {
        if (lane == 0 && rank == 0) {
            const uint32_t idesc = (1u << 4) | ((uint32_t(a.n) >> 3) << 17) | (((a.mode == 3 ? 256u : 128u) >> 4) << 24) |
                                   ((a.mode == 1 || a.mode == 4 || a.mode == 5) ? (1u << 16) : 0u);
            const uint64_t a0 = (a.mode == 2) ? (uint64_t)(tmem + 256)
                                : (a.mode == 4 ? umma_smem_desc(sbase + kA, 512, SWZ_64B) : umma_smem_desc(sbase + kA, 1024, SWZ_128B));
            const uint64_t b0 = (a.mode == 1 || a.mode == 4 || a.mode == 5) ? desc_mn(sbase + kB, 1024, 4096) : umma_smem_desc(sbase + kB, 1024, SWZ_128B);
            const unsigned long long t0 = clock64();
            if (a.mode < 0) {
                while (clock64() - t0 < 400000ull) {}
            } else {
                for (int it = 0; it < a.iters; ++it) {
                    if constexpr (kCluster) {
                        issue16<3>(tmem, a0, b0, idesc);
                    } else {
                        if (a.mode == 0) issue16<0>(tmem, a0, b0, idesc);
                        else if (a.mode == 1) issue16<1>(tmem, a0, b0, idesc);
                        else if (a.mode == 4) issue16<4>(tmem, a0, b0, idesc);
                        else if (a.mode == 5) issue16<5>(tmem, a0, b0, idesc, bar(2));
                        else issue16<2>(tmem, a0, b0, idesc);
                    }
                }
                if constexpr (kCluster) umma_commit2(bar(0)); else umma_commit(bar(0));
                mbar_wait(bar(0), 0);
            }
            cycles = clock64() - t0;
            *flag = 1;
            if (kCluster) {
                uint32_t raddr;
                asm volatile("mapa.shared::cluster.u32 %0, %1, 1;" : "=r"(raddr) : "r"(sbase + kBars + 224));
                asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(raddr), "r"(1u) : "memory");
            }
        }
    } else if (warp >= 1 && warp <= 3 && a.bulk < 0) {
        // smem -> global bulk stores (what the forward kernel's activation saving does): keep `slots` groups in flight
        if (warp - 1 < -a.bulk && lane == 0) {
            uint8_t* dst = a.dst + ((size_t)blockIdx.x * 3 + (warp - 1)) * (1u << 20);
            uint32_t issued = 0;
            while (!*flag) {
                bulk_s2g(dst + (issued & 63) * 16384, sbase + kRing + (warp - 1) * 16384, a.chunk);
                bulk_commit();
                ++issued;
                if (a.slots == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                else if (a.slots == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                else asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
            }
            bulk_wait_all();
            n_bulk = issued;
        }
    } else if (warp >= 1 && warp <= 3) {
        if (warp - 1 < a.bulk && lane < a.lanes) {
            const int issuer = (warp - 1) * a.lanes + lane;
            const uint32_t base = sbase + kRing + issuer * a.slots * a.chunk;
            const uint32_t bar0 = bar(1 + issuer * a.slots);
            uint32_t issued = 0;
            while (!*flag) {
                const uint32_t slot = issued % a.slots, round = issued / a.slots;
                if (round >= 1) mbar_wait(bar0 + 8 * slot, (round - 1) & 1);
                mbar_arrive_expect_tx(bar0 + 8 * slot, a.chunk);
                bulk_g2s(base + slot * a.chunk, a.src + ((size_t)((issued * 7 + issuer) & 63) * 32768), a.chunk, bar0 + 8 * slot);
                ++issued;
            }
            for (uint32_t i = (issued >= (uint32_t)a.slots ? issued - a.slots : 0); i < issued; ++i)
                mbar_wait(bar0 + 8 * (i % a.slots), (i / a.slots) & 1);
            n_bulk = issued;
        }
    } else if (warp >= 4 && warp < 4 + a.sts_warps) {
        const uint32_t dst = sbase + kSts + (warp - 4) * 2048 + lane * 16;
        uint32_t v = tid;
        while (!*flag) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
                asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(dst + (i & 3) * 512), "r"(v + i) : "memory");
            n_sts += 16;
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (kCluster) cluster_sync_all();
    if (warp == 0) {
        if (kCluster) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
        else tmem_dealloc(tmem, 512);
    }
    {
        if (cycles) atomicAdd(a.out + 3 * blockIdx.x + 0, cycles);
        if (n_sts && lane == 0) atomicAdd(a.out + 3 * blockIdx.x + 1, n_sts);
        if (n_bulk) atomicAdd(a.out + 3 * blockIdx.x + 2, n_bulk);
    }
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    unsigned long long* d_out;
    cudaMalloc(&d_out, sms * 24);
    uint8_t* d_src;
    cudaMalloc(&d_src, 4 << 20);
    cudaMemset(d_src, 0, 4 << 20);
    uint8_t* d_dst;
    cudaMalloc(&d_dst, (size_t)sms * 3 << 20);
    cudaFuncSetAttribute(rate2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    cudaFuncSetAttribute(rate2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    const char* names[] = {"none (spin)", "SS A K / B K   ", "SS A K / B MN  ", "TS A tmem / B K", "2CTA SS M=256  ", "SS A K SW64/B MN", "B MN + commit/4 "};
    struct Exp { int mode, n, sts, bulk, lanes, slots, chunk; };
    std::vector<Exp> exps;
    // A-operand swizzle: 128-byte rows (64 k per row) against the 64-byte rows (32 k) of the MLP weight stages
    for (int mode : {1, 5})
        for (int n : {256, 128}) {
            exps.push_back({mode, n, 0, 0, 1, 1, 16384});
            exps.push_back({mode, n, 8, 0, 1, 1, 16384});
        }
    for (int grid : {sms}) {
        for (const Exp& e : exps) {
            Args a{e.mode, e.n, 512, e.sts, e.bulk, e.lanes, e.slots, e.chunk, d_src, d_dst, d_out};
            cudaMemset(d_out, 0, sms * 24);
            if (e.mode == 3) {
                cudaLaunchConfig_t cfg{};
                cfg.gridDim = dim3(grid & ~1);
                cfg.blockDim = dim3(384);
                cfg.dynamicSmemBytes = kSmem;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeClusterDimension;
                at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                cfg.attrs = at; cfg.numAttrs = 1;
                cudaLaunchKernelEx(&cfg, rate2_kernel<true>, a);
            } else {
                rate2_kernel<false><<<grid, 384, kSmem>>>(a);
            }
            cudaError_t err = cudaGetLastError();
            if (err != cudaSuccess) { printf("launch error %s (mode %d)\n", cudaGetErrorString(err), e.mode); return 1; }
            err = cudaDeviceSynchronize();
            if (err != cudaSuccess) { printf("error %s (mode %d)\n", cudaGetErrorString(err), e.mode); return 1; }
            std::vector<unsigned long long> h(3 * grid);
            cudaMemcpy(h.data(), d_out, grid * 24, cudaMemcpyDeviceToHost);
            unsigned long long mx = 0, sts = 0, blk = 0; int nc = 0;
            for (int i = 0; i < grid; ++i) { if (h[3 * i]) { ++nc; if (h[3 * i] > mx) mx = h[3 * i]; } sts += h[3 * i + 1]; blk += h[3 * i + 2]; }
            const int ctas = (e.mode == 3) ? (grid & ~1) : grid;
            const double n_mma = e.mode < 0 ? 0 : 512.0 * 16;
            printf("%s N=%3d sts_warps=%d bulk=%dw x %dl x %ds x %5d B: %7.1f cyc/MMA   st.shared %6.1f B/clk/SM   bulk %6.1f B/clk/SM   (cycles %llu)\n",
                   names[e.mode + 1], e.n, e.sts, e.bulk, e.lanes, e.slots, e.chunk, n_mma ? (double)mx / n_mma : 0.0,
                   (double)sts * 512.0 / ctas / (double)mx, (double)blk * e.chunk / ctas / (double)mx, mx);
        }
    }
    return 0;
}
