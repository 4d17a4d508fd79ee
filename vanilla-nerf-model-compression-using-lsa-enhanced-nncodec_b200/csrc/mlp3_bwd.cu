// Fused NeRF MLP backward into the LSA scales on tcgen05 tensor cores (sm_100a).
//
// Replaces torch autograd over NeRF.forward with ScaledLinear layers (utils.py:57-80, transforms.py:104-111) when
// only `weight_scaling` requires grad (framework/pytorch_model/__init__.py:1129-1145, run_nerf.py:756).  For a layer
//   y = s * (delta * (L x)) + b        (L = integer levels, s = LSA scale per output channel)
// the scale gradient is  ds[o] = sum_n dY[n,o] * (y[n,o] - b[o]) / s[o]  -- an elementwise product reduced over
// points, no weight-gradient GEMM -- and the input gradient is the dgrad GEMM  dX = (dY * delta * s) L.
//
// Same CTA organisation as the forward kernel (mlp3_layout.h): per group of 256 points the chain of 9 dgrad GEMMs
//   D[k][n] = sum_o L[o][k] * G[n][o]      (A = W^T chunks streamed from L2, B = gradient tile in shared memory)
// runs with input channels k on the TMEM lanes, so an epilogue thread owns ONE channel of the layer whose output the
// accumulator is the gradient of: it masks with the saved forward activation (ReLU), accumulates its channel's
// scale-gradient sums in registers (no cross-lane reduction), and writes the next GEMM's operand row.  Saved
// activations are read straight from the forward kernel's chunk-major HBM slots (coalesced 32-byte requests,
// L2-prefetched two jobs ahead).
// As in the forward kernel, the two 128-point halves of a group are in flight together (own accumulators, barriers and
// team of 8 epilogue warps; every weight chunk feeds both).
// Gradients are tiny (1e-7..1e-3): each half-group is multiplied by a power of two that brings max|d_raw| into [8,16)
// -- backpropagation is linear, the factor is exact and is divided out where the scale-gradient sums are flushed --
// which keeps fp16's 11-bit significand for the operands without its range problem.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdio.h>

#include "mlp3_common.cuh"

namespace nerfq {

struct Bwd3Params {
    const uint8_t* packed;
    const float* d_raw;      // [n_points, 4]
    const float* raw;        // [n_points, 4]   forward output (for the rgb / alpha head terms)
    const uint8_t* save;     // saved operand images from the forward pass (kSave3GroupBytes per group)
    long long* grad_tmp;     // [2436] scratch inside the packed buffer: sum_n dY (y - b) = s * ds in fixed point, zero on entry
    long long n_points;
    int n_groups;
    unsigned long long* dbg;     // tracing build only: 8 cycle counters per CTA
    Prog3Bwd prog;
};

constexpr uint32_t kS3BwdGa = kS3BwdPg + 4096;          // float ga[256]: alpha-head gradient per point
constexpr uint32_t kS3BwdMax = kS3BwdGa + 1024;         // per half (16 bytes each): uint gmax[2], float rinv, int group
static_assert(kS3BwdMax + 32 <= kS3Misc, "backward scratch must fit the encoding-tile region");

#ifndef NERFQ_BWD_PREFETCH
#define NERFQ_BWD_PREFETCH 1
#endif
#ifndef NERFQ_BWD_HH
// saved activations in flight per thread: 2 = two register slots, a chunk is requested two chunks ahead (default);
// 5 = four slots, a whole job ahead, also across the CTA's groups (measured equal: profiles/r02_ab_bwd_pinned_addresses_and_register_split.log)
#define NERFQ_BWD_HH 2
#endif
// register budgets of this kernel's control / epilogue warps (4*32*ctrl + 16*32*epi must not exceed the 61,440 registers of
// the launch allocation: 64 / 104 or 32 / 112)
#ifndef NERFQ_BWD_REGS_CTRL
#define NERFQ_BWD_REGS_CTRL NERFQ_REGS_CTRL3
#endif
#ifndef NERFQ_BWD_REGS_EPI
#define NERFQ_BWD_REGS_EPI NERFQ_REGS_EPI3
#endif
#ifndef NERFQ_BWD_PIN
#define NERFQ_BWD_PIN 0
#endif
#ifndef NERFQ_BWD_PF_DIST
#define NERFQ_BWD_PF_DIST 2       // how many jobs ahead a job's slice of saved activations is requested into L2
#endif
// the 64 bytes (L2 fill granularity) around p into L2
__device__ __forceinline__ void prefetch_l2_line(const void* p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p) : "memory");
}
// fire-and-forget add into global memory (L2 atomic unit) of v in fixed point, mlp3_layout.h kGradFixShift: the sum does
// not depend on the order in which CTAs arrive
__device__ __forceinline__ void red_global_add_fixed(long long* p, float v) {
    const long long q = __float2ll_rn(v * (float)(1ull << kGradFixShift));          // saturates; |v| < 32768 in range
    asm volatile("red.global.add.u64 [%0], %1;" ::"l"(p), "l"(q) : "memory");
}
// 32 bytes (16 saved activations of one channel) in one request; the L2 is asked to fetch the surrounding 256 bytes
struct H32 { uint4 a, b; };
__device__ __forceinline__ H32 ldg_nc_32B(const void* p) {
    H32 r;
#ifndef NERFQ_BWD_LD_POLICY
#define NERFQ_BWD_LD_POLICY 0
#endif
#if NERFQ_BWD_LD_POLICY == 0
    asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r.a.x), "=r"(r.a.y), "=r"(r.a.z), "=r"(r.a.w), "=r"(r.b.x), "=r"(r.b.y), "=r"(r.b.z), "=r"(r.b.w) : "l"(p));
#elif NERFQ_BWD_LD_POLICY == 1      // read-once data leaves the L2 first (the weight image all SMs stream from it stays)
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.L2::256B.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r.a.x), "=r"(r.a.y), "=r"(r.a.z), "=r"(r.a.w), "=r"(r.b.x), "=r"(r.b.y), "=r"(r.b.z), "=r"(r.b.w) : "l"(p));
#else
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r.a.x), "=r"(r.a.y), "=r"(r.a.z), "=r"(r.a.w), "=r"(r.b.x), "=r"(r.b.y), "=r"(r.b.z), "=r"(r.b.w) : "l"(p));
#endif
    return r;
}

template <bool kTrace>
__global__ void __launch_bounds__(kThreads3, 1) mlp3_backward_kernel(const __grid_constant__ Bwd3Params prm) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t sbase = opaque_u32(smem_u32(smem));
    const int warp = uniform_warp_idx();
    const int lane = threadIdx.x & 31;
    auto bar = [&](int i) { return sbase + kS3Bars + 8u * i; };

    {
        if (threadIdx.x < 8) reinterpret_cast<uint32_t*>(smem + kS3BwdMax)[threadIdx.x] = 0u;
    }
    // the CTA owns the SM (1 CTA/SM by shared-memory size) and allocates all 512 TMEM columns, so the allocation starts
    // at column 0; treating the base as a constant frees a register in every epilogue thread
    constexpr uint32_t tmem_base = 0;
    if (setup3(smem, sbase, warp) != tmem_base) __trap();
    const int first = blockIdx.x, stride = gridDim.x;
    const int n_iters = first < prm.n_groups ? (prm.n_groups - first + stride - 1) / stride : 0;

    if (warp >= kCtrlWarp0 && warp < kCtrlWarp0 + kCtrlWarps3) {
        const int cw = warp - kCtrlWarp0;
        asm volatile("setmaxnreg.dec.sync.aligned.u32 " NERFQ_BWD_REGS_CTRL ";");
        if (cw == 0 || cw == 2) {
            loader3(sbase, prm.packed + kOffBwd3Image, kBwd3Chunks, n_iters, cw >> 1);
        } else {
            if (n_iters > 0) issuer3<false, kTrace>(sbase, tmem_base, n_iters, (uint32_t)(cw >> 1), prm.dbg);
        }
    } else {
        // ================= epilogue warps =================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 " NERFQ_BWD_REGS_EPI ";");
        // the role dispatch above uses the shuffled (uniform-register) warp index; the per-thread geometry below is derived
        // from threadIdx so that the compiler rematerialises it from the special register rather than spilling it
        const int tw = (int)(threadIdx.x >> 5);
        const int e = tw - kEpiWarp0;
        // All 16 warps work on one job at a time (a job is latency- not issue-bound, so halving the points per warp
        // halves its duration; two 8-warp teams running both jobs of a step concurrently measured slower).  A warp owns
        // lane quarter q and point quarter pq (64 points = 4 chunks of 16) of both accumulators.
        const int q = tw & 3, pq = e >> 2;
        // half-group A (points 0..127: warps e < 8) or B: own accumulators, barriers, gradient scale and scratch
        const int team = e >> 3, le = e & 7;
        const uint32_t tmem_lane = tmem_base + (uint32_t(q * 32) << 16) + team * 256 + (pq & 1) * 64;
        const uint32_t act = sbase + kS3Act;
        const uint32_t pg_a = sbase + kS3BwdPg, ga_a = sbase + kS3BwdGa + 4 * (pq * 64), max_a = sbase + kS3BwdMax + 16 * team;
        const float2* g_sb = reinterpret_cast<const float2*>(prm.packed + kOffSB);
        const float* g_wa = reinterpret_cast<const float*>(prm.packed + kOffWAlpha);
        const float* g_wr = reinterpret_cast<const float*>(prm.packed + kOffWRgb);
        uint32_t ph_acc0 = 0, ph_acc1 = 0, ph_sf = 0;
        unsigned long long t_pro = 0, t_acc = 0, t_job = 0, t_views = 0;
        const bool tracing = kTrace && lane == 0 && e == 5;
        const int cl = 32 * q + lane;                 // this thread's channel within a half (its TMEM lane)

        auto publish = [&](int which) {
            fence_proxy_async_smem();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(which));
        };
        // saved activations of this thread: channel chh of slot `slot`, its 4 chunks of 16 points start at point chunk 4*pq
        auto saved_row = [&](int g, int slot, uint32_t chh) {
            return prm.save + (size_t)g * kSave3GroupBytes + (size_t)slot * kSave3SlotBytes + save3_offset(4 * pq, chh);
        };
        // one chunk of 16 points of this thread's channel.  d = gradient w.r.t. the layer output (fp32, from TMEM),
        // h = saved activations (8 x half2).  The elementwise work runs on packed halves:
        //   dh = fp16(d);  s1h += dh*h;  g = dh (masked where the unit was inactive);  s2h += g;  G row <- g
        // (the channel's delta*scale factor of the dgrad GEMM lives in the backward weight image, mlp3_layout.h)
        // The half2 partial sums cover 8 terms each and are folded into the fp32 sums per chunk (rounding errors are
        // unbiased and average out over the ~10^5 chunks a channel sees).
        auto chunk16 = [&](const uint32_t (&dpk)[8], const uint4 h0, const uint4 h1, bool relu, bool write, uint32_t row_addr,
                           uint32_t swz, int cc, float& s1, float& s2) {
            const uint32_t hw[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
            uint32_t pk[8];
            const __half2 zero2 = __float2half2_rn(0.0f);
            __half2 s1h = zero2, s2h = zero2;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const __half2 h2 = *reinterpret_cast<const __half2*>(&hw[i]);
                const __half2 dh = *reinterpret_cast<const __half2*>(&dpk[i]);
                s1h = __hfma2(dh, h2, s1h);
                __half2 g2 = dh;
                if (relu) g2 = __hmul2(dh, __hgt2(h2, zero2));       // mask: 1.0 where the unit was active
                s2h = __hadd2(s2h, g2);
                pk[i] = *reinterpret_cast<const uint32_t*>(&g2);
            }
            const float2 a1 = __half22float2(s1h), a2 = __half22float2(s2h);
            s1 += a1.x + a1.y;
            s2 += a2.x + a2.y;
            if (write) {
                // (row | swz) ^ x == row + (x ^ swz): the row starts on a 128-byte boundary (see the forward kernel)
#pragma unroll
                for (int k = 0; k < 2; ++k)
                    st_shared_v4((row_addr | swz) ^ (uint32_t)((cc * 2 + k) << 4), pk[4 * k], pk[4 * k + 1], pk[4 * k + 2], pk[4 * k + 3]);
            }
        };
        // chunk cc (0..3) of this thread's 64 points: 32 bytes, 8 KB apart (mlp3_layout.h, save3_offset)
        auto pair_off = [&](int cc, uint32_t) { return save3_offset(cc, 0); };
        // L2 prefetch of the saved activations, three jobs ahead inside a group; the first three slices of the CTA's next
        // group are requested when the current group's job loop ends (keeping "next group" out of the job loop's live
        // registers: a spilled value costs a ~1000-cycle local-memory load per job).  A job's 64 KB slice
        // (mlp3_layout.h, Prog3Bwd::slice_off) is 16 pieces of 4 KB (128 channels x 32 B), one per point chunk, 8 KB
        // apart: 512 lines, one per thread.
        const uint32_t pf_thread = (uint32_t)e * 8192u + (NERFQ_BWD_PREFETCH == 3 ? 0u : (uint32_t)lane * 128u);
        // (one request per 128-byte line.  The L2 fills 64 bytes per miss, so this covers half of every slice -- ncu r02 counted
        // exactly 50 % of the demand loads' sectors as L2 misses -- but requesting both halves, NERFQ_BWD_PREFETCH = 2, measured
        // 5 % SLOWER (0.953 vs 0.906 ms) and no prefetch at all 1 % slower: profiles/r02_ab_prefetch_and_saver_waits.log)
        auto prefetch_seq = [&](int g, int v) {      // v: index into [views, job 0 .. 17]
            const uint8_t* line = prm.save + (size_t)g * kSave3GroupBytes + (pf_thread + prm.prog.slice_off[v]);
            if (NERFQ_BWD_PREFETCH == 3) {
                // the warp's whole 4 KB piece (all 32 lines, both halves of each) as ONE request to the copy engine
                if (lane == 0) bulk_prefetch_l2(line, 4096);
                return;
            }
            if (NERFQ_BWD_PREFETCH >= 1) prefetch_l2_line(line);
            if (NERFQ_BWD_PREFETCH >= 2) prefetch_l2_line(line + 64);
        };
        float2 c_next = make_float2(1.f, 0.f);
        if (n_iters > 0) {
            publish(kB3ActHi + team);       // D_hi is free at kernel start
            c_next = __ldg(&g_sb[prm.prog.job[0].ch + cl]);
        }
#if NERFQ_BWD_HH == 5
        // a whole job ahead, across groups: the last job of a group requests the first job of the CTA's next group
        H32 hh[4];
        if (n_iters > 0) {
            const Job3 j0 = prm.prog.job[0];
            const uint8_t* hrow0 = saved_row(first, j0.slot, ((j0.flags & JB_HI_HALF) ? 128u : 0u) + cl);
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) hh[cc] = ldg_nc_32B(hrow0 + pair_off(cc, 0));
        }
#endif
        int it = 0;
        for (int g = first; g < prm.n_groups; g += stride, ++it) {
            // ================= prologue: head gradients, group scale, views-layer gradient =================
            unsigned long long tp0 = 0;
            if (tracing) tp0 = clock64();
            const int pt = team * (int)kHalfPts3 + le * 32 + lane;             // the first four warps of a team own one point each
            const long long gidx = (long long)g * kGroupPts + pt;
            float4 dr = make_float4(0.f, 0.f, 0.f, 0.f), rw = make_float4(0.f, 0.f, 0.f, 0.f);
            if (le < 4 && gidx < prm.n_points) {
                dr = *reinterpret_cast<const float4*>(prm.d_raw + 4 * gidx);
                rw = *reinterpret_cast<const float4*>(prm.raw + 4 * gidx);
            }
            if (it == 0) {         // later groups: prefetched by the previous group's last jobs
#pragma unroll
                for (int v = 0; v <= NERFQ_BWD_PF_DIST; ++v) prefetch_seq(g, v);
            }
            const uint32_t slot_max = max_a + 4 * (it & 1);
            if (le < 4) {
                float m = fmaxf(fmaxf(fabsf(dr.x), fabsf(dr.y)), fmaxf(fabsf(dr.z), fabsf(dr.w)));
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
                if (lane == 0) asm volatile("red.shared.max.u32 [%0], %1;" ::"r"(slot_max), "r"(__float_as_uint(m)) : "memory");
                // head scale gradients (unscaled): sum_n dr * (raw - b)
                float pr = dr.x * (rw.x - __ldg(&g_sb[kChRgb + 0]).y);
                float pgn = dr.y * (rw.y - __ldg(&g_sb[kChRgb + 1]).y);
                float pb = dr.z * (rw.z - __ldg(&g_sb[kChRgb + 2]).y);
                float pa = dr.w * (rw.w - __ldg(&g_sb[kChAlpha]).y);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    pr += __shfl_xor_sync(0xffffffffu, pr, o);
                    pgn += __shfl_xor_sync(0xffffffffu, pgn, o);
                    pb += __shfl_xor_sync(0xffffffffu, pb, o);
                    pa += __shfl_xor_sync(0xffffffffu, pa, o);
                }
                if (lane == 0) {
                    red_global_add_fixed(prm.grad_tmp + kChRgb + 0, pr);
                    red_global_add_fixed(prm.grad_tmp + kChRgb + 1, pgn);
                    red_global_add_fixed(prm.grad_tmp + kChRgb + 2, pb);
                    red_global_add_fixed(prm.grad_tmp + kChAlpha, pa);
                }
            }
            named_bar_sync3(1 + team, 32 * kTeamWarps3);
            float rscale = 0.0f, rinv = 0.0f;
            {
                uint32_t mbits;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(mbits) : "r"(slot_max) : "memory");
                const int ex = (mbits >> 23) & 0xff;
                if (ex >= 16 && ex <= 240) {
                    rscale = __int_as_float((257 - ex) << 23);       // max|d_raw| * rscale in [8, 16)
                    rinv = __int_as_float((ex - 3) << 23);
                }
            }
            if (le < 4) {
                const float er = __ldg(&g_sb[kChRgb + 0]).x, eg = __ldg(&g_sb[kChRgb + 1]).x, eb = __ldg(&g_sb[kChRgb + 2]).x;
                const float ea = __ldg(&g_sb[kChAlpha]).x;
                st_shared_v4(pg_a + 16 * pt, __float_as_uint(dr.x * rscale * er), __float_as_uint(dr.y * rscale * eg),
                             __float_as_uint(dr.z * rscale * eb), 0u);
                st_shared_f32(sbase + kS3BwdGa + 4 * pt, dr.w * rscale * ea);
            }
            // 1/scale is read back from shared memory where it is needed (once per job): a register would be spilled to
            // local memory, whose loads take ~1000 cycles with the L1 carved out for shared memory
            // (one slot is enough: it is rewritten after the next group's first barrier, which every warp reaches only
            // after its last job of this group)
            const uint32_t rinv_a = max_a + 8;
            if (le == 4 && lane == 0) {
                asm volatile("st.shared.u32 [%0], %1;" ::"r"(max_a + 4 * ((it & 1) ^ 1)), "r"(0u) : "memory");
                st_shared_f32(rinv_a, rinv);
                asm volatile("st.shared.u32 [%0], %1;" ::"r"(max_a + 12), "r"(g) : "memory");      // group index for the job loop (same reason)
            }
            named_bar_sync3(1 + team, 32 * kTeamWarps3);

            // ---- views-layer job: d hv[k][n] = sum_c g_c[n] * w_rgb[c][k], channels 0..127 ----
            if (tracing) { const unsigned long long t = clock64(); t_pro += t - tp0; tp0 = t; }
#if NERFQ_BWD_HH != 5
            H32 hh[2];             // saved activations of the dgrad chain's current job, chunks 0 and 1 (see the job loop)
            {
                const Job3 j0 = prm.prog.job[0];
                const uint8_t* hrow0 = saved_row(g, j0.slot, ((j0.flags & JB_HI_HALF) ? 128u : 0u) + cl);
                hh[0] = ldg_nc_32B(hrow0 + pair_off(0, 0));
                hh[1] = ldg_nc_32B(hrow0 + pair_off(1, 0));
            }
#endif
            {
                const uint32_t chh = cl;                                   // views hidden channel
                const float2 c = __ldg(&g_sb[kChViews + chh]);
                const float w0 = __ldg(&g_wr[chh]), w1 = __ldg(&g_wr[128 + chh]), w2 = __ldg(&g_wr[256 + chh]);
                const uint8_t* hrow = saved_row(g, 9, chh);
                const uint32_t row_addr = act + (chh >> 3) * kKGroup3 + pq * kNGroup3 + (chh & 7u) * 128u;
                const uint32_t swz = (chh & 7u) << 4;
                float s1 = 0.0f, s2 = 0.0f;
#pragma unroll 1
                for (int cc = 0; cc < 4; ++cc) {
                    const H32 hp = ldg_nc_32B(hrow + pair_off(cc, swz));
                    float d[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float4 gq = ld_shared_v4f(pg_a + 16 * (pq * 64 + cc * 16 + i));
                        d[i] = fmaf(gq.x, w0, fmaf(gq.y, w1, gq.z * w2));
                    }
                    uint32_t dpk[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) dpk[i] = cvt_pack_f16(d[2 * i], d[2 * i + 1]);
                    chunk16(dpk, hp.a, hp.b, true, true, row_addr, swz, cc, s1, s2);
                }
                // ds * s = sum dY (y - b) = s1 - b * (sum dY)
                publish(kB3ActLo + team);
                red_global_add_fixed(prm.grad_tmp + kChViews + chh, (s1 - c.y * s2) * ld_shared_f32(rinv_a));
            }

            // ================= dgrad chain =================
            if (tracing) t_views += clock64() - tp0;
#pragma unroll 1
            for (int j = 0; j < kBwd3Jobs; ++j) {
                const Job3 jb = prm.prog.job[j];
                const uint32_t f = jb.flags;
                const uint32_t hi = (f & JB_HI_HALF) ? 1u : 0u;
                const uint32_t chh = 128u * hi + cl;                       // channel within the layer
                const float2 c = c_next;
                const int gj = ld_shared_s32(max_a + 12);           // == g, from shared memory instead of a spill slot
                const uint8_t* hrow = saved_row(gj, jb.slot, chh);
                const uint32_t row_addr = act + (chh >> 3) * kKGroup3 + pq * kNGroup3 + (chh & 7u) * 128u;
                const uint32_t swz = (chh & 7u) << 4;
                // Saved activations: hh[] holds chunks 0 and 1 of THIS job, requested while the previous job (or the views job)
                // was still running -- the accumulator of a "hi" job is ready long before its job starts, so loads issued at
                // the job's start had their whole HBM latency exposed (ncu r02: 12 % of all warp samples on their first use).
                // Each consumed slot is refilled with the chunk two ahead; chunks 2 and 3 free the slots for the next job.
#if NERFQ_BWD_HH == 5
                const Job3 jn = prm.prog.job[j + 1 < kBwd3Jobs ? j + 1 : 0];
                const int gn = j + 1 < kBwd3Jobs ? gj : (gj + stride < prm.n_groups ? gj + stride : gj);
                const uint8_t* hrow_next = saved_row(gn, jn.slot, ((jn.flags & JB_HI_HALF) ? 128u : 0u) + cl);
#else
                const Job3 jn = prm.prog.job[j + 1 < kBwd3Jobs ? j + 1 : j];
                const uint8_t* hrow_next = saved_row(gj, jn.slot, ((jn.flags & JB_HI_HALF) ? 128u : 0u) + cl);
#endif
                if (j + 1 + NERFQ_BWD_PF_DIST <= kBwd3Jobs) prefetch_seq(gj, j + 1 + NERFQ_BWD_PF_DIST);      // NERFQ_BWD_PF_DIST jobs ahead, into L2 (one line per thread)
                c_next = __ldg(&g_sb[prm.prog.job[j + 1 < kBwd3Jobs ? j + 1 : 0].ch + cl]);        // in flight during this job; issued before
                                                                                                     // the wait like everything that does not need the accumulator
                unsigned long long tj0 = 0;
                if (tracing) tj0 = clock64();
                if (hi) { mbar_wait_acc(bar(kB3AccReady + 2 * team + 1), ph_acc1); ph_acc1 ^= 1; }
                else { mbar_wait_acc(bar(kB3AccReady + 2 * team), ph_acc0); ph_acc0 ^= 1; }
                tc_fence_after_sync();
                if (tracing) { const unsigned long long t = clock64(); t_acc += t - tj0; tj0 = t; }
#if NERFQ_BWD_PIN
                // keep the job's two addresses in registers: without this the compiler re-derives them from %tid in every chunk
                uint32_t ta = tmem_lane + (hi ? 128u : 0u), rs = row_addr | swz;
                asm volatile("" : "+r"(ta), "+r"(rs));
#else
                const uint32_t ta = tmem_lane + (hi ? 128u : 0u), rs = row_addr;
#endif
                const bool relu = f & JB_RELU, write = !(f & JB_NO_WRITE);
                float s1 = 0.0f, s2 = 0.0f;
                uint32_t va[16];
                if (f & JB_ADD_ALPHA) {        // d h8 also receives the alpha head's gradient: one job in 18, done as a
                                               // read-modify-write pass over the accumulator so the chunk loop stays lean
                    const float wa = __ldg(&g_wa[chh]);
#pragma unroll 1
                    for (int cc = 0; cc < 4; ++cc) {
                        tmem_ld16(ta + 16 * cc, va);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; i += 4) {
                            const float4 ga = ld_shared_v4f(ga_a + 4 * (cc * 16 + i));
                            va[i] = __float_as_uint(fmaf(ga.x, wa, __uint_as_float(va[i])));
                            va[i + 1] = __float_as_uint(fmaf(ga.y, wa, __uint_as_float(va[i + 1])));
                            va[i + 2] = __float_as_uint(fmaf(ga.z, wa, __uint_as_float(va[i + 2])));
                            va[i + 3] = __float_as_uint(fmaf(ga.w, wa, __uint_as_float(va[i + 3])));
                        }
                        tmem_st16(ta + 16 * cc, va);
                    }
                    tmem_st_wait();
                }
                tmem_ld16(ta, va);
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    tmem_ld_wait();
                    // fp16 copy of the gradients; the accumulator registers are then free for the next chunk's load,
                    // which is in flight while this chunk is processed
                    uint32_t dpk[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) dpk[i] = cvt_pack_f16(__uint_as_float(va[2 * i]), __uint_as_float(va[2 * i + 1]));
                    if (cc < 3) tmem_ld16(ta + 16 * (cc + 1), va);
                    if (cc == 0 && (f & JB_WAIT_SF)) { mbar_wait(bar(kB3StageFree + 2 * team + (q >> 1)), ph_sf); ph_sf ^= 1; }
#if NERFQ_BWD_HH == 5
                    const H32 hp = hh[cc];
                    hh[cc] = ldg_nc_32B(hrow_next + pair_off(cc, swz));          // unconditional: the row always exists
#else
                    const H32 hp = hh[cc & 1];
                    if (cc < 2) hh[cc & 1] = ldg_nc_32B(hrow + pair_off(cc + 2, swz));       // refill with chunk cc + 2
                    else if (j + 1 < kBwd3Jobs) hh[cc & 1] = ldg_nc_32B(hrow_next + pair_off(cc - 2, swz));      // next job's chunk cc - 2
#endif
                    chunk16(dpk, hp.a, hp.b, relu, write, rs, swz, cc, s1, s2);
                }
                if (write || hi) publish((hi ? kB3ActHi : kB3ActLo) + team);
                red_global_add_fixed(prm.grad_tmp + jb.ch + cl, (s1 - c.y * s2) * ld_shared_f32(rinv_a));     // after the hand-over
                if (tracing) t_job += clock64() - tj0;
            }
            if (g + stride < prm.n_groups) {
#pragma unroll
                for (int v = 0; v <= NERFQ_BWD_PF_DIST; ++v) prefetch_seq(g + stride, v);
            }
        }
        if (tracing && prm.dbg) {
            unsigned long long* o = prm.dbg + 8 * 148 + 32 * blockIdx.x;
            o[0] = t_pro; o[1] = t_views; o[2] = t_acc; o[3] = t_job;
        }
    }

    // ---- teardown ----
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// d_scale[i] += fix[i] / 2^kGradFixShift / s[i];  fix[i] = 0  (the scratch is left zeroed for the next launch).
// A scale of exactly 0 makes y - b = 0, so the kernel cannot recover sum dY (L x) delta from its saved activations:
// that channel's gradient is reported as 0 (finite) instead of 0/0 -- documented in include/nerfq.h.
__global__ void mlp3_backward_finalize_kernel(const uint8_t* packed, long long* __restrict__ fix, float* __restrict__ d_scale) {
    const float* scale = reinterpret_cast<const float*>(packed + kOffScale);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < kNumChannels) {
        const long long q = fix[i];
        fix[i] = 0;
        const float sc = scale[i];
        if (q != 0 && sc != 0.0f) d_scale[i] += (float)((double)q * (1.0 / (double)(1ull << kGradFixShift))) / sc;
    }
}

// ---- data-parallel finalize over peer memory ---------------------------------------------------------------------
// The all-reduce of the fixed-point sums fused into the finalize step: every rank copies its sums into a staging area
// that all ranks of the node map (symmetric memory over NVLink / NVSwitch), raises a flag in every peer's pad, waits for
// every peer's flag, and then reads ALL ranks' staged sums directly and adds them up.  Integer addition: every rank gets
// the same bits whatever the order, and the same bits as one rank on the concatenated batch.  One kernel, one block;
// replaces {NCCL all-reduce (0.032 ms at N = 8 plus two kernel boundaries: 0.069 ms exposed) ; finalize ; finalize}.
//
// Peer region layout (nerfq_dp_peer_bytes(), zero-initialised once, before the ranks first meet):
//   long long stage[2][2][2440]   [epoch parity][network][channel]     unsigned pad[64] at kPeerPadOff: pad[r] = last epoch
//                                                                       rank r has published
// Reuse: stage[e & 1] is rewritten by its owner at epoch e + 2, i.e. after the owner has seen every peer's flag of epoch
// e + 1, which a peer raises only after it finished reading epoch e.  Waits are bounded (a dead peer traps, not hangs).
constexpr int kPeerNetElems = 2440;
constexpr size_t kPeerPadOff = (size_t)2 * 2 * kPeerNetElems * sizeof(long long);
constexpr size_t kPeerBytes = kPeerPadOff + 64 * sizeof(unsigned);

__device__ __forceinline__ void st_release_sys_u32(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ long long ld_relaxed_sys_s64(const long long* p) {      // system-coherent: never served by a stale L1 line
    long long v;
    asm volatile("ld.relaxed.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(1024) mlp3_backward_finalize_peers_kernel(const uint8_t* packed0, const uint8_t* packed1,
                                                                             long long* __restrict__ fix, uint8_t* const* __restrict__ peers,
                                                                             int world, int rank, unsigned* __restrict__ epoch_p,
                                                                             float* __restrict__ d_scale) {
    const unsigned e = *epoch_p + 1u;                       // epochs start at 1; the pads start at 0
    const int par = (int)(e & 1u);
    const int n = 2 * kPeerNetElems;
    long long* stage = reinterpret_cast<long long*>(peers[rank]) + par * n;
    // (0) publish this rank's sums; the accumulators are left zeroed for the next backward
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        stage[i] = fix[i];
        fix[i] = 0;
    }
    __threadfence_system();
    __syncthreads();
    // (1) raise this rank's flag in every peer's pad, then wait for every peer's flag in ours
    if ((int)threadIdx.x < world) {
        st_release_sys_u32(reinterpret_cast<unsigned*>(peers[threadIdx.x] + kPeerPadOff) + rank, e);
        const unsigned* mine = reinterpret_cast<const unsigned*>(peers[rank] + kPeerPadOff) + threadIdx.x;
        unsigned spins = 0;
        while ((int)(ld_acquire_sys_u32(mine) - e) < 0) {
            if (++spins > (1u << 26)) {
                printf("nerfq: data-parallel finalize: rank %d gave up waiting for rank %d at epoch %u\n", rank, (int)threadIdx.x, e);
                __trap();
            }
        }
    }
    __syncthreads();
    // (2) add up all ranks' staged sums and convert.  The reads cross NVLink (~2 us each): a thread keeps 4 ranks x 5
    // elements = 20 of them in flight, so N = 8 costs two round trips, not forty.
    constexpr int kPerThread = (2 * kPeerNetElems + 1023) / 1024;       // 5
    long long q[kPerThread];
#pragma unroll
    for (int k = 0; k < kPerThread; ++k) q[k] = 0;
    for (int r0 = 0; r0 < world; r0 += 4) {
        long long v[4][kPerThread];
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
            const long long* src = reinterpret_cast<const long long*>(peers[min(r0 + rr, world - 1)]) + par * n;
#pragma unroll
            for (int k = 0; k < kPerThread; ++k) {
                const int i = (int)threadIdx.x + 1024 * k;
                v[rr][k] = (r0 + rr < world && i < n) ? ld_relaxed_sys_s64(src + i) : 0;
            }
        }
#pragma unroll
        for (int rr = 0; rr < 4; ++rr)
#pragma unroll
            for (int k = 0; k < kPerThread; ++k) q[k] += v[rr][k];
    }
#pragma unroll
    for (int k = 0; k < kPerThread; ++k) {
        const int i = (int)threadIdx.x + 1024 * k;
        if (i >= n) continue;
        const int net = i / kPeerNetElems, c = i - net * kPeerNetElems;
        if (c < kNumChannels && q[k] != 0) {
            const float sc = reinterpret_cast<const float*>((net ? packed1 : packed0) + kOffScale)[c];
            if (sc != 0.0f) d_scale[net * kNumChannels + c] += (float)((double)q[k] * (1.0 / (double)(1ull << kGradFixShift))) / sc;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) *epoch_p = e;
}

}  // namespace nerfq

static unsigned long long* g_trace3b = nullptr;
// Profiling aid (not part of include/nerfq.h): see nerfq_mlp_set_trace.
extern "C" void nerfq_mlp_set_trace_bwd(unsigned long long* buf) { g_trace3b = buf; }

extern "C" unsigned long long nerfq_mlp_grad_fix_bytes(void) { return nerfq::kGradTmp3Bytes; }

// The backward kernel proper: accumulates s*ds per channel into `grad_fix` (64-bit fixed point, see mlp3_layout.h).
extern "C" int nerfq_mlp_backward_partial(const void* packed, const float* d_raw, const float* raw, const void* save, long long n_points,
                                           long long* grad_fix, int max_ctas, cudaStream_t stream) {
    using namespace nerfq;
    if (n_points == 0) return 0;
    if (!packed || !d_raw || !raw || !save || !grad_fix || n_points < 0) return -1;
    static const Prog3Bwd prog = make_prog3_bwd();
    const int n_groups = (int)((n_points + kGroupPts - 1) / kGroupPts);
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (max_ctas > 0 && max_ctas < sms) sms = max_ctas;
    const int grid = n_groups < sms ? n_groups : sms;
    Bwd3Params prm{(const uint8_t*)packed, d_raw, raw, (const uint8_t*)save, grad_fix, n_points, n_groups, g_trace3b, prog};
    if (g_trace3b) {
        if (cudaFuncSetAttribute(mlp3_backward_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kS3Bytes) != cudaSuccess) return -2;
        mlp3_backward_kernel<true><<<grid, kThreads3, kS3Bytes, stream>>>(prm);
    } else {
        if (cudaFuncSetAttribute(mlp3_backward_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kS3Bytes) != cudaSuccess) return -2;
        mlp3_backward_kernel<false><<<grid, kThreads3, kS3Bytes, stream>>>(prm);
    }
    return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

extern "C" int nerfq_mlp_backward_finalize(const void* packed, long long* grad_fix, float* d_scale, cudaStream_t stream) {
    using namespace nerfq;
    if (!packed || !grad_fix || !d_scale) return -1;
    mlp3_backward_finalize_kernel<<<(kNumChannels + 255) / 256, 256, 0, stream>>>((const uint8_t*)packed, grad_fix, d_scale);
    return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

extern "C" unsigned long long nerfq_dp_peer_bytes(void) { return nerfq::kPeerBytes; }

extern "C" int nerfq_mlp_backward_finalize_peers(const void* packed_coarse, const void* packed_fine, long long* grad_fix2, void* const* peers,
                                                  int world, int rank, unsigned int* epoch, float* d_scale2, cudaStream_t stream) {
    using namespace nerfq;
    if (!packed_coarse || !grad_fix2 || !peers || !epoch || !d_scale2 || world < 1 || world > 64 || rank < 0 || rank >= world) return -1;
    mlp3_backward_finalize_peers_kernel<<<1, 1024, 0, stream>>>((const uint8_t*)packed_coarse, (const uint8_t*)(packed_fine ? packed_fine : packed_coarse),
                                                               grad_fix2, (uint8_t* const*)peers, world, rank, epoch, d_scale2);
    return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

// partial + finalize through the scratch array inside the packed buffer (zeroed by nerfq_pack_net and by every finalize)
extern "C" int nerfq_mlp_backward(void* packed, const float* d_raw, const float* raw, const void* save, long long n_points,
                                   float* d_scale, int max_ctas, cudaStream_t stream) {
    using namespace nerfq;
    if (n_points == 0) return 0;
    if (!packed || !d_scale) return -1;
    long long* fix = reinterpret_cast<long long*>(reinterpret_cast<uint8_t*>(packed) + kOffGradTmp3);
    const int rc = nerfq_mlp_backward_partial(packed, d_raw, raw, save, n_points, fix, max_ctas, stream);
    if (rc != 0) return rc;
    return nerfq_mlp_backward_finalize(packed, fix, d_scale, stream);
}
