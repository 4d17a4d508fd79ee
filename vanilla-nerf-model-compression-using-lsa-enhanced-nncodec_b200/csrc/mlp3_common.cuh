// Device-side pieces shared by the fused MLP kernels (mlp3_fwd.cu, mlp3_bwd.cu):
// CTA shape, shared-memory map, barrier indices, the weight loader and the MMA-issuing loop.
#pragma once
#include "mlp3_layout.h"
#include "ptx_sm100.cuh"

namespace nerfq {

// warps 0, 2: weight loaders (warp 0 owns TMEM)   warp 1: MMA issuer   warp 3: idle
// warps 4..19: epilogue; a warp may only touch TMEM lanes 32*(warp % 4).., so warp w owns lane quarter q = w & 3 and
// point quarter pq = (w - 4) >> 2.
// Registers: 20 warps are launched with 96 registers each; the control warp group then releases registers
// (setmaxnreg.dec) and the four epilogue warp groups claim them (setmaxnreg.inc) -- see kRegsCtrl3 / kRegsEpi3.
constexpr int kCtrlWarps3 = 4;
#define NERFQ_REGS_CTRL3 "56"
#define NERFQ_REGS_EPI3 "104"
constexpr int kEpiWarps3 = 16;
constexpr int kThreads3 = 32 * (kCtrlWarps3 + kEpiWarps3);
constexpr int kSlots3 = 4;

constexpr uint32_t kS3Act = 0;                                   // 128 KB activation / gradient tile
constexpr uint32_t kS3Ring = kS3Act + kAct3Bytes;                // 4 x 16 KB weight chunks
constexpr uint32_t kS3Enc = kS3Ring + kSlots3 * kChunk3Bytes;    // forward: 32 KB encodings; backward: scratch
constexpr uint32_t kS3Misc = kS3Enc + kEnc3Bytes;                // forward: float[256] alpha sums
constexpr uint32_t kS3Bars = kS3Misc + 1024;
constexpr uint32_t kS3TmemPtr = kS3Bars + 8 * 16;
constexpr uint32_t kS3Bytes = kS3TmemPtr + 16 + 1024;            // + slack for the manual 1 KB alignment

// backward: the 32 KB "Enc" region holds  float4 pg[256] (per-point head gradients)  and  float red[2436]
constexpr uint32_t kS3BwdPg = kS3Enc;
constexpr uint32_t kS3BwdRed = kS3Enc + 4096;

constexpr int kB3WFull = 0;       // [4]
constexpr int kB3WEmpty = 4;      // [4]
constexpr int kB3ActLo = 8;       // 16 arrivals: job for channels 0..127 done (operand written, D_lo drained)
constexpr int kB3ActHi = 9;
constexpr int kB3AccReady = 10;   // [2] tcgen05.commit
constexpr int kB3StageFree = 12;  // [2] tcgen05.commit

__device__ __forceinline__ uint64_t umma_desc_mn3(uint32_t saddr) {     // MN-major SWIZZLE_128B activation tile
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
    d |= static_cast<uint64_t>((kNGroup3 >> 4) & 0x3FFF) << 16;
    d |= static_cast<uint64_t>((kKGroup3 >> 4) & 0x3FFF) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(SWZ_128B) << 61;
    return d;
}
constexpr uint32_t kIdesc3BK = (1u << 4) | ((256u >> 3) << 17) | ((128u >> 4) << 24);   // f16 x f16 -> f32, M=128, N=256
constexpr uint32_t kIdesc3BMN = kIdesc3BK | (1u << 16);                                  // B operand MN-major

__device__ __forceinline__ void named_bar_sync3(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- one-time setup shared by both kernels; returns the TMEM base -----------------------------------
__device__ __forceinline__ uint32_t setup3(uint8_t* smem, uint32_t sbase, int warp, int act_arrivals = kEpiWarps3) {
    auto bar = [&](int i) { return sbase + kS3Bars + 8u * i; };
    if (threadIdx.x == 0) {
        for (int i = 0; i < kSlots3; ++i) { mbar_init(bar(kB3WFull + i), 1); mbar_init(bar(kB3WEmpty + i), 1); }
        mbar_init(bar(kB3ActLo), act_arrivals);
        mbar_init(bar(kB3ActHi), act_arrivals);
        mbar_init(bar(kB3AccReady + 0), 1);
        mbar_init(bar(kB3AccReady + 1), 1);
        mbar_init(bar(kB3StageFree + 0), 1);
        mbar_init(bar(kB3StageFree + 1), 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(sbase + kS3TmemPtr, 512);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    return *reinterpret_cast<volatile uint32_t*>(smem + kS3TmemPtr);
}

// ---- weight loader: the image is a stream of equal chunks in consumption order, repeated per group ----
// Called by a whole converged warp (so that addresses stay in uniform registers); one elected lane issues.
// One issuing thread sustains only ~30-37 B/clk of bulk copies (profiles/r01_umma_rate2_bulk_issue_parallelism.log)
// against the 32 B/clk the MMAs consume, so kLoaders3 warps share the stream: warp `which` copies the chunks with
// sequence number = which (mod kLoaders3).
constexpr int kLoaders3 = 2;
__device__ __forceinline__ void loader3(uint32_t sbase, const uint8_t* img, int n_chunks, int n_iters, int which) {
    auto bar = [&](int i) { return sbase + kS3Bars + 8u * i; };
    uint32_t seq = 0;
    for (int it = 0; it < n_iters; ++it) {
        const uint8_t* src = img;
        for (int c = 0; c < n_chunks; ++c, ++seq, src += kChunk3Bytes) {
            if ((int)(seq % kLoaders3) != which) continue;
            const uint32_t slot = seq & (kSlots3 - 1), par = (seq >> 2) & 1;
            mbar_wait(bar(kB3WEmpty) + 8 * slot, par ^ 1);
            if (elect_one()) {
                mbar_arrive_expect_tx(bar(kB3WFull) + 8 * slot, kChunk3Bytes);
                bulk_g2s(sbase + kS3Ring + slot * kChunk3Bytes, src, kChunk3Bytes, bar(kB3WFull) + 8 * slot);
            }
            __syncwarp();
        }
    }
}

// ---- MMA issuer: walks the half-step program once per group ------------------------------------------------
// Called by a whole converged warp: every lane waits, one elected lane issues the MMAs and commits, and all operands
// are warp-uniform so they live in uniform registers.  The tensor pipe queues only ~2 MMAs, so the code between the
// last MMA of a chunk and the first MMA of the next (commit, ring wait, descriptor bump) has to stay short.
// `first_lo_wait`: the forward kernel's first group waits for the initial encodings (later groups get them with the
// ActLo arrival the rgb step already consumed).
// kTrace: accumulate cycle counters {total, wait WFull, wait ActLo, wait ActHi} into dbg[8*blockIdx.x + 0..3].
template <int kHalves, bool kTrace = false>
__device__ __forceinline__ void issuer3(uint32_t sbase, uint32_t tmem_base, const Half3* __restrict__ prog, int n_iters,
                                        bool first_lo_wait, unsigned long long* dbg = nullptr) {
    unsigned long long t_w = 0, t_lo = 0, t_hi = 0, t_begin = 0;
    if (kTrace) t_begin = clock64();
    const uint32_t bar0 = sbase + kS3Bars;
    const uint64_t a_desc0 = umma_smem_desc(sbase + kS3Ring, 512, SWZ_64B);
    const uint64_t b_act0 = umma_desc_mn3(sbase + kS3Act);
    const uint64_t b_enc0 = umma_smem_desc(sbase + kS3Enc, 1024, SWZ_128B);
    uint32_t seq = 0, ph_lo = 0, ph_hi = 0;
    bool slot_probed = false;          // the previous chunk already saw this chunk's weights in the ring
    // one weight chunk: 4 MMAs; stage 1 of the B operand starts `stage` (16-byte units) after stage 0, the second
    // K=16 half of a stage `kstep` after the first.  The tensor pipe queues ~2 MMAs, so after the first two the issue
    // of the third blocks until the first retires: that time is used to probe the NEXT chunk's "slot full" barrier,
    // which takes the ~130-cycle barrier wait out of the gap between chunks whenever the loader is ahead.
    auto chunk = [&](uint32_t d_tmem, uint64_t b, uint32_t idesc, uint32_t kstep, uint32_t stage, uint32_t accumulate,
                     uint32_t commit_a, uint32_t commit_b) {
        const uint32_t slot = seq & (kSlots3 - 1), par = (seq >> 2) & 1;
        ++seq;
        if (!slot_probed) {
            unsigned long long t0 = 0;
            if (kTrace) t0 = clock64();
            mbar_wait(bar0 + 8 * (kB3WFull + slot), par);
            if (kTrace) t_w += clock64() - t0;
        }
        tc_fence_after_sync();
        const uint64_t ad = a_desc0 + slot * (kChunk3Bytes >> 4);
        if (elect_one()) {
            umma_ss(d_tmem, ad, b, idesc, accumulate);
            umma_ss(d_tmem, ad + 2, b + kstep, idesc, 1u);
        }
        __syncwarp();
        slot_probed = __all_sync(0xffffffffu, mbar_test_wait(bar0 + 8 * (kB3WFull + (seq & (kSlots3 - 1))), (seq >> 2) & 1));
        if (elect_one()) {
            umma_ss(d_tmem, ad + (kStage3Bytes >> 4), b + stage, idesc, 1u);
            umma_ss(d_tmem, ad + (kStage3Bytes >> 4) + 2, b + stage + kstep, idesc, 1u);
            umma_commit(bar0 + 8 * (kB3WEmpty + slot));
            if (commit_a) umma_commit(commit_a);
            if (commit_b) umma_commit(commit_b);
        }
        __syncwarp();
    };
    if (first_lo_wait) {
        mbar_wait(bar0 + 8 * kB3ActLo, ph_lo);
        ph_lo ^= 1;
        tc_fence_after_sync();
    }
    for (int it = 0; it < n_iters; ++it) {
#pragma unroll 1
        for (int h = 0; h < kHalves; ++h) {
            const Half3 hs = prog[h];
            const uint32_t f = hs.flags;
            const uint32_t hi = f & HS_ACC_HI;
            const uint32_t d_tmem = tmem_base + (hi ? 256u : 0u);
            const uint32_t acc_bar = bar0 + 8 * (kB3AccReady + (hi ? 1 : 0));
            if (f & HS_WAIT_LO) {
                unsigned long long t0 = 0;
                if (kTrace) t0 = clock64();
                mbar_wait(bar0 + 8 * kB3ActLo, ph_lo);
                ph_lo ^= 1;
                tc_fence_after_sync();
                if (kTrace) t_lo += clock64() - t0;
            }
            if (f & HS_WAIT_HI_AT0) {
                unsigned long long t0 = 0;
                if (kTrace) t0 = clock64();
                mbar_wait(bar0 + 8 * kB3ActHi, ph_hi);
                ph_hi ^= 1;
                tc_fence_after_sync();
                if (kTrace) t_hi += clock64() - t0;
            }
            const int n_act = hs.n_act, n_enc = hs.n_enc;
            uint64_t b = b_act0;
#pragma unroll 1
            for (int j = 0; j < n_act; ++j) {
                if (j == 2 && (f & HS_WAIT_HI_AT2)) {
                    unsigned long long t0 = 0;
                    if (kTrace) t0 = clock64();
                    mbar_wait(bar0 + 8 * kB3ActHi, ph_hi);
                    ph_hi ^= 1;
                    tc_fence_after_sync();
                    if (kTrace) t_hi += clock64() - t0;
                }
                const uint32_t sf = ((f & HS_SF) && j < 2) ? bar0 + 8 * (kB3StageFree + j) : 0u;
                const uint32_t done = (j + 1 == n_act && n_enc == 0) ? acc_bar : 0u;
                chunk(d_tmem, b, kIdesc3BMN, (2u * kKGroup3) >> 4, (4u * kKGroup3) >> 4, j > 0 ? 1u : 0u, sf, done);
                b += (8u * kKGroup3) >> 4;
            }
            if (n_enc) chunk(d_tmem, b_enc0, kIdesc3BK, 32u >> 4, 64u >> 4, n_act > 0 ? 1u : 0u, 0u, acc_bar);
        }
    }
    if (kTrace && dbg && (threadIdx.x & 31) == 0) {
        dbg[8 * blockIdx.x + 0] = clock64() - t_begin;
        dbg[8 * blockIdx.x + 1] = t_w;
        dbg[8 * blockIdx.x + 2] = t_lo;
        dbg[8 * blockIdx.x + 3] = t_hi;
    }
}

// ---- positional encoding (run_nerf_helpers.py:23-49), split in two halves of 5 levels --------------------
// levels [l0, l0+5) of gamma(p): out[6*i + c] = sin(2^(l0+i) p_c), out[6*i + 3 + c] = cos(2^(l0+i) p_c).
// sincosf once per coordinate, then angle doubling (error < 2^4 ulp, far below the fp16 operand rounding).
__device__ __forceinline__ void encode5(const float p[3], float scale, float* out /* 30 */) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float s, co;
        sincosf(p[c] * scale, &s, &co);
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            if (i > 0) {
                const float s2 = 2.0f * s * co;
                const float c2 = fmaf(-2.0f * s, s, 1.0f);
                s = s2; co = c2;
            }
            out[6 * i + c] = s;
            out[6 * i + 3 + c] = co;
        }
    }
}

// Write 32 fp16 columns [col0, col0+32) of row `row` of the K-major SWIZZLE_128B encoding tile (shared address `enc`).
__device__ __forceinline__ void write_enc32(uint32_t enc, int row, int col0, const float (&v)[32]) {
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
        st_shared_v4(enc + sw128_offset(row, (col0 >> 3) + ch), cvt_pack_f16(v[8 * ch + 0], v[8 * ch + 1]),
                     cvt_pack_f16(v[8 * ch + 2], v[8 * ch + 3]), cvt_pack_f16(v[8 * ch + 4], v[8 * ch + 5]),
                     cvt_pack_f16(v[8 * ch + 6], v[8 * ch + 7]));
    }
}

// gamma(x) half: role 0 -> columns 0..31 = [y, z, levels 0..4]; role 1 -> columns 32..63 = [levels 5..9, x, 0]
__device__ __forceinline__ void write_pe_half(uint32_t enc, int row, int role, const float p[3]) {
    float v[32];
    if (role == 0) {
        v[0] = p[1]; v[1] = p[2];
        encode5(p, 1.0f, v + 2);
    } else {
        encode5(p, 32.0f, v);
        v[30] = p[0]; v[31] = 0.0f;
    }
    write_enc32(enc, row, role * 32, v);
}

// gamma(d): 27 values (d, then 4 levels) + 5 zeros into columns 0..31
__device__ __forceinline__ void write_dir_enc(uint32_t enc, int row, const float d[3]) {
    float v[32];
    v[0] = d[0]; v[1] = d[1]; v[2] = d[2];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float s, co;
        sincosf(d[c], &s, &co);
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            if (l > 0) {
                const float s2 = 2.0f * s * co;
                const float c2 = fmaf(-2.0f * s, s, 1.0f);
                s = s2; co = c2;
            }
            v[3 + 6 * l + c] = s;
            v[3 + 6 * l + 3 + c] = co;
        }
    }
#pragma unroll
    for (int i = 27; i < 32; ++i) v[i] = 0.0f;
    write_enc32(enc, row, 0, v);
}

// Sum each of 32 per-lane columns over the 32 lanes of the warp; lane j returns column j.
__device__ __forceinline__ float column_reduce32_3(float (&p)[32], int lane) {
    float q16[16], q8[8], q4[4], q2[2];
    {
        const bool hi = lane & 16;
#pragma unroll
        for (int i = 0; i < 16; ++i) q16[i] = (hi ? p[i + 16] : p[i]) + __shfl_xor_sync(0xffffffffu, hi ? p[i] : p[i + 16], 16);
    }
    {
        const bool hi = lane & 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) q8[i] = (hi ? q16[i + 8] : q16[i]) + __shfl_xor_sync(0xffffffffu, hi ? q16[i] : q16[i + 8], 8);
    }
    {
        const bool hi = lane & 4;
#pragma unroll
        for (int i = 0; i < 4; ++i) q4[i] = (hi ? q8[i + 4] : q8[i]) + __shfl_xor_sync(0xffffffffu, hi ? q8[i] : q8[i + 4], 4);
    }
    {
        const bool hi = lane & 2;
#pragma unroll
        for (int i = 0; i < 2; ++i) q2[i] = (hi ? q4[i + 2] : q4[i]) + __shfl_xor_sync(0xffffffffu, hi ? q4[i] : q4[i + 2], 2);
    }
    const bool hi = lane & 1;
    return (hi ? q2[1] : q2[0]) + __shfl_xor_sync(0xffffffffu, hi ? q2[0] : q2[1], 1);
}

// Sum each of 16 per-lane columns over the 32 lanes of the warp; lanes 2j and 2j+1 both return column j.
__device__ __forceinline__ float column_reduce16_3(float (&p)[16], int lane) {
    float q8[8], q4[4], q2[2];
    {
        const bool hi = lane & 16;
#pragma unroll
        for (int i = 0; i < 8; ++i) q8[i] = (hi ? p[i + 8] : p[i]) + __shfl_xor_sync(0xffffffffu, hi ? p[i] : p[i + 8], 16);
    }
    {
        const bool hi = lane & 8;
#pragma unroll
        for (int i = 0; i < 4; ++i) q4[i] = (hi ? q8[i + 4] : q8[i]) + __shfl_xor_sync(0xffffffffu, hi ? q8[i] : q8[i + 4], 8);
    }
    {
        const bool hi = lane & 4;
#pragma unroll
        for (int i = 0; i < 2; ++i) q2[i] = (hi ? q4[i + 2] : q4[i]) + __shfl_xor_sync(0xffffffffu, hi ? q4[i] : q4[i + 2], 4);
    }
    const bool hi = lane & 2;
    const float r = (hi ? q2[1] : q2[0]) + __shfl_xor_sync(0xffffffffu, hi ? q2[0] : q2[1], 2);
    return r + __shfl_xor_sync(0xffffffffu, r, 1);
}

}  // namespace nerfq
