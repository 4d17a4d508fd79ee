// Fused positional-encoding + NeRF MLP forward, "channels on TMEM lanes" formulation (net_layout2.h).
//
// Replaces, for one network and M = n_rays*S sample points:
//   pts = o + d*z                                   run_nerf.py:408,430
//   embed_fn / embeddirs_fn / cat                   run_nerf.py:52-59, run_nerf_helpers.py:18-67
//   NeRF.forward with ScaledLinear layers           utils.py:57-80, transforms.py:104-111
// Output: raw[M,4] = (rgb logits, sigma) as run_network returns it (run_nerf.py:61-63).
//
// One persistent CTA per SM iterates over groups of 256 points.  Per layer the tensor cores compute
//   D[o][n] = sum_k L[o][k] * X[n][k]       (L = integer weight levels, fp16; X = activations, fp16; D fp32 in TMEM)
// as two accumulators of 128 output channels x 256 points.  Weight stages (8 KB) stream from L2 through a
// 6-slot ring of bulk async copies; the activation tile stays in shared memory and is rewritten in place.
// 16 epilogue warps each own 32 output channels x 128 points: a thread keeps the dequantisation constants of
// ITS channel in two registers and applies  y = acc * (delta * s[o]) + b[o], ReLU, fp16 conversion, storing
// 8 points per 16-byte shared-memory store into the next layer's operand tile.  The epilogue of the first
// 128 channels overlaps the MMAs of the second 128 and vice versa, so the tensor pipe only waits at the
// ends of a group.  The alpha head is reduced on CUDA cores in the L7 epilogue; the rgb head is one more
// (3-of-128-row) MMA.  With `save` set every operand tile is streamed to HBM for the backward pass.
#include <cuda_runtime.h>
#include <stdlib.h>

#include "mlp_common.cuh"
#include "net_layout2.h"

namespace nerfq {

struct Fwd2Params {
    const uint8_t* packed;
    const float* rays;       // [n_rays, 11]
    const float* z;          // [n_rays * S]
    float* raw;              // [n_rays * S, 4]
    uint8_t* save;           // nullable; kSave2PairBytes per group of 256 points
    long long n_points;
    int samples_per_ray;
    int n_groups;
    int debug_flags;         // bit 0: skip the layer epilogues (tensor-pipe ceiling measurement, results invalid)
};

constexpr int kEpiWarps2 = 16;
constexpr int kThreads2 = 32 * (kCtrlWarps + kEpiWarps2);
constexpr int kSlots2 = 4;                      // ring slots of two weight stages (16 KB) each
constexpr int kSlot2Bytes = 2 * kStage2Bytes;

constexpr uint32_t kS2Act = 0;
constexpr uint32_t kS2Enc = kS2Act + kActBytes;
constexpr uint32_t kS2Ring = kS2Enc + kEncBytes;
constexpr uint32_t kS2Out = kS2Ring + kSlots2 * kSlot2Bytes;   // float[256]: alpha-head partial sums
constexpr uint32_t kS2Bars = kS2Out + 256 * 4;
constexpr uint32_t kS2TmemPtr = kS2Bars + 8 * 32;
constexpr uint32_t kS2BytesFwd = kS2TmemPtr + 16 + 1024;

constexpr int kB2WFull = 0;       // [kSlots2]
constexpr int kB2WEmpty = 4;      // [kSlots2]
constexpr int kB2LoReady = 12;    // channels 0..127 of the operand tile written (or encodings), D_lo drained
constexpr int kB2HiReady = 13;    // channels 128..255 written, D_hi drained
constexpr int kB2AccReady = 14;   // [2]
constexpr int kB2StageFree = 16;  // [4]  K stage j of the operand tile no longer read by this layer's MMAs

__device__ __constant__ Step2 kFwd2Dev[kFwd2Steps] = NERFQ_FWD2_TABLE;

__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t saddr) {     // MN-major SWIZZLE_128B activation tile
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
    d |= static_cast<uint64_t>((kNGroupBytes >> 4) & 0x3FFF) << 16;
    d |= static_cast<uint64_t>((kKGroupBytes >> 4) & 0x3FFF) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(SWZ_128B) << 61;
    return d;
}
constexpr uint32_t kIdescBK = (1u << 4) | ((256u >> 3) << 17) | ((128u >> 4) << 24);   // f16 x f16 -> f32, M=128, N=256
constexpr uint32_t kIdescBMN = kIdescBK | (1u << 16);                                   // B operand MN-major

// Sum each of 32 per-lane columns over the 32 lanes of the warp; lane j returns column j.
__device__ __forceinline__ float column_reduce32_f(float (&p)[32], int lane) {
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

// Epilogue of one layer for one warp: 32 channels (lane = channel) x 128 points (4 chunks of 32).
template <bool kRelu, bool kAlpha>
__device__ __forceinline__ void epilogue2(uint32_t tmem_addr, uint8_t* act, uint32_t ch, int n_base, float es, float b, float wa,
                                          float* out_s, int lane) {
    uint32_t v[2][32];
    tmem_ld32(tmem_addr, v[0]);
    tmem_ld_wait();
    const uint32_t row_off = (ch >> 3) * kKGroupBytes + (ch & 7u) * 128u;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        if (c + 1 < 4) tmem_ld32(tmem_addr + 32 * (c + 1), v[(c + 1) & 1]);
        uint32_t* cur = v[c & 1];
        float y[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            y[i] = fmaf(__uint_as_float(cur[i]), es, b);
            if (kRelu) y[i] = fmaxf(y[i], 0.0f);
        }
        const int n0 = n_base + c * 32;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint4 q;
            q.x = pack_half2(y[8 * k + 0], y[8 * k + 1]);
            q.y = pack_half2(y[8 * k + 2], y[8 * k + 3]);
            q.z = pack_half2(y[8 * k + 4], y[8 * k + 5]);
            q.w = pack_half2(y[8 * k + 6], y[8 * k + 7]);
            const int n = n0 + 8 * k;
            *reinterpret_cast<uint4*>(act + row_off + (n >> 6) * kNGroupBytes + ((((n & 63) >> 3) ^ (ch & 7u)) << 4)) = q;
        }
        if (kAlpha) {
#pragma unroll
            for (int i = 0; i < 32; ++i) y[i] *= wa;
            const float s = column_reduce32_f(y, lane);
            atomicAdd(&out_s[n0 + lane], s);
        }
        if (c + 1 < 4) tmem_ld_wait();
    }
}

// Debug timeline (debug_flags bit 3): block 0 appends (tag, clock) pairs to prm.save, 4096 entries per traced warp.
struct Trace {
    unsigned long long* buf;
    int n;
    __device__ __forceinline__ void mark(int tag) {
        if (buf && n < 2047) { buf[2 * n] = tag; buf[2 * n + 1] = clock64(); ++n; }
    }
};

template <bool kSave>
__global__ void __launch_bounds__(kThreads2, 1) mlp2_forward_kernel(const Fwd2Params prm) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t sbase = smem_u32(smem);
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    float* out_s = reinterpret_cast<float*>(smem + kS2Out);
    auto bar = [&](int i) { return sbase + kS2Bars + 8u * i; };

    for (int i = threadIdx.x; i < 256; i += kThreads2) out_s[i] = 0.0f;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kSlots2; ++i) { mbar_init(bar(kB2WFull + i), 1); mbar_init(bar(kB2WEmpty + i), 1); }
        mbar_init(bar(kB2LoReady), 8);
        mbar_init(bar(kB2HiReady), 8);
        mbar_init(bar(kB2AccReady + 0), 1);
        mbar_init(bar(kB2AccReady + 1), 1);
        for (int j = 0; j < 4; ++j) mbar_init(bar(kB2StageFree + j), 1);
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc(sbase + kS2TmemPtr, 512);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + kS2TmemPtr);
    const int first = blockIdx.x, stride = gridDim.x;
    Trace tr{nullptr, 0};
    if (!kSave && (prm.debug_flags & 8) && blockIdx.x == 0 && lane == 0 && (warp == 1 || warp == 4 || warp == 8))
        tr.buf = reinterpret_cast<unsigned long long*>(prm.save) + (warp == 1 ? 0 : warp == 4 ? 1 : 2) * 4096;

    if (warp == 0) {
        // ================= weight loader =================
        // The image is stored in consumption order.  Stages travel in chunks of two (16 KB, one bulk copy, one
        // barrier round trip).  Control threads run dependent scalar code at ~5 cycles per instruction and a
        // satisfied mbarrier wait alone costs ~100 cycles, so the step program is unrolled at compile time
        // (all table look-ups, offsets and branch conditions fold) and each handshake covers 4 MMAs.
        if (lane == 0) {
            const uint8_t* img = prm.packed + kOffFwd2Image;
            const bool no_copy = prm.debug_flags & 4;
            uint32_t seq = 0;
            for (int g = first; g < prm.n_groups; g += stride) {
                uint32_t off = 0;
#pragma unroll
                for (int s = 0; s < kFwd2Steps; ++s) {
                    const Step2 st = fwd2_step(s);
                    const int nst = st.kh + st.kp;
#pragma unroll
                    for (int mh = 0; mh < 2; ++mh) {
                        if (mh < st.halves) {
#pragma unroll
                            for (int j0 = 0; j0 < nst; j0 += 2) {
                                const uint32_t bytes = (nst - j0 >= 2 ? 2 : 1) * kStage2Bytes;
                                const uint32_t slot = seq & (kSlots2 - 1), par = (seq >> 2) & 1;
                                ++seq;
                                mbar_wait(bar(kB2WEmpty) + 8 * slot, par ^ 1);
                                if (no_copy) {
                                    mbar_arrive(bar(kB2WFull) + 8 * slot);
                                } else {
                                    mbar_arrive_expect_tx(bar(kB2WFull) + 8 * slot, bytes);
                                    bulk_g2s(sbase + kS2Ring + slot * kSlot2Bytes, img + off, bytes, bar(kB2WFull) + 8 * slot);
                                }
                                off += bytes;
                            }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            uint32_t seq = 0, n_lo = 0, n_hi = 0;
            const uint64_t a_desc0 = umma_smem_desc(sbase + kS2Ring, 512, SWZ_64B);
            const uint64_t b_act0 = umma_desc_mn(sbase + kS2Act);
            const uint64_t b_enc0 = umma_smem_desc(sbase + kS2Enc, 1024, SWZ_128B);
            const bool no_mma = prm.debug_flags & 2;
            for (int g = first; g < prm.n_groups; g += stride) {
#pragma unroll
                for (int s = 0; s < kFwd2Steps; ++s) {
                    const Step2 st = fwd2_step(s);
                    const int nst = st.kh + st.kp;
                    tr.mark(100 + s);
#pragma unroll
                    for (int mh = 0; mh < 2; ++mh) {
                        if (mh < st.halves) {
                            const int acc = (mh == 1 || st.dst == 1) ? 1 : 0;
                            const uint32_t d_tmem = tmem_base + acc * 256;
                            // first chunk of this step that touches D_hi or channels >= 128 of the operand tile
                            const int hi_j0 = (st.dst == 1) ? 0 : (st.kh > 4 ? 4 : -1);
#pragma unroll
                            for (int j0 = 0; j0 < nst; j0 += 2) {
                                // Every wait below is matched by exactly one arrival that itself sits behind a wait on
                                // this thread's commits, so a barrier can never run two phases ahead of its waiter.
                                // (Step 0 of later groups needs no LoReady: the encodings were written, and D_lo drained,
                                // before the LoReady phase the rgb step consumed.  The rgb step needs no HiReady: D_hi has
                                // been free since the phase the views step consumed.)
                                if (mh == 0 && j0 == 0 && (s > 0 || g == first)) {
                                    mbar_wait(bar(kB2LoReady), n_lo++ & 1);
                                    tc_fence_after_sync();
                                    tr.mark(200 + s);
                                }
                                if (s != 10 && ((mh == 0 && j0 == hi_j0) || (mh == 1 && j0 == 0 && hi_j0 < 0))) {
                                    mbar_wait(bar(kB2HiReady), n_hi++ & 1);
                                    tc_fence_after_sync();
                                    tr.mark(300 + s);
                                }
                                const uint32_t slot = seq & (kSlots2 - 1), par = (seq >> 2) & 1;
                                ++seq;
                                mbar_wait(bar(kB2WFull) + 8 * slot, par);
                                tc_fence_after_sync();
                                tr.mark(400 + mh * 10 + (j0 >> 1));
                                const uint64_t a_desc = a_desc0 + slot * (kSlot2Bytes >> 4);     // descriptor addresses are in 16-byte units
#pragma unroll
                                for (int jj = 0; jj < 2; ++jj) {
                                    const int j = j0 + jj;
                                    if (j < nst) {
                                        if (!no_mma) {
                                            const uint64_t ad = a_desc + jj * (kStage2Bytes >> 4);
                                            if (j < st.kh) {
                                                const uint64_t bd = b_act0 + j * ((4 * kKGroupBytes) >> 4);
                                                if (j == 0) umma_ss_c<0>(d_tmem, ad, bd, kIdescBMN); else umma_ss_c<1>(d_tmem, ad, bd, kIdescBMN);
                                                umma_ss_c<1>(d_tmem, ad + 2, bd + ((2 * kKGroupBytes) >> 4), kIdescBMN);
                                            } else {
                                                const uint64_t bd = b_enc0 + (j - st.kh) * (64 >> 4);
                                                if (j == 0) umma_ss_c<0>(d_tmem, ad, bd, kIdescBK); else umma_ss_c<1>(d_tmem, ad, bd, kIdescBK);
                                                umma_ss_c<1>(d_tmem, ad + 2, bd + 2, kIdescBK);
                                            }
                                        }
                                        if (s >= 1 && s <= 8 && mh == 1 && j < 4) {
                                            if (no_mma) mbar_arrive(bar(kB2StageFree + j)); else umma_commit(bar(kB2StageFree + j));
                                        }
                                    }
                                }
                                if (no_mma) mbar_arrive(bar(kB2WEmpty) + 8 * slot); else umma_commit(bar(kB2WEmpty) + 8 * slot);
                            }
                            if (no_mma) mbar_arrive(bar(kB2AccReady + acc)); else umma_commit(bar(kB2AccReady + acc));
                        }
                    }
                }
            }
        }
    } else if (warp >= kCtrlWarps) {
        // ================= epilogue warps =================
        const int e = warp - kCtrlWarps;
        const int q = warp & 3, mh = (e >> 2) & 1, ph = e >> 3;
        const uint32_t ch = 128 * mh + 32 * q + lane;           // this thread's channel within the layer
        const uint32_t tmem_lane = tmem_base + (uint32_t(q * 32) << 16);
        uint8_t* act = smem + kS2Act;
        uint8_t* enc = smem + kS2Enc;
        const float2* g_sb = reinterpret_cast<const float2*>(prm.packed + kOffSB);
        const float* g_wa = reinterpret_cast<const float*>(prm.packed + kOffWAlpha);
        const int n_base = ph * 128;
        // shared-memory / global pieces this warp owns in every activation image (for the save stream)
        const uint32_t piece_off = ((128 * mh + 32 * q) >> 3) * kKGroupBytes + ph * 2 * kNGroupBytes;
        uint32_t n_acc = 0, n_sf = 0;

        auto publish = [&](int which) {
            fence_proxy_async_smem();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(which));
        };
        auto store_pieces = [&](uint8_t* dst_img) {      // after the warp's stores: stream its 4 x 2 KB to HBM
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
#pragma unroll
                for (int gk = 0; gk < 4; ++gk)
                    bulk_s2g(dst_img + piece_off + gk * kKGroupBytes, sbase + kS2Act + piece_off + gk * kKGroupBytes, 2 * kNGroupBytes);
                bulk_commit();
            }
        };

        if (mh == 0) {
            const int pt = ph * 128 + q * 32 + lane;          // the point this thread encodes
            float p[3], vd[3];
            long long gidx = 0;
            auto load_point = [&](int g) {
                gidx = (long long)g * kPairPoints + pt;
                const long long gc = gidx < prm.n_points ? gidx : prm.n_points - 1;
                const long long ray = gc / prm.samples_per_ray;
                const float zz = __ldg(prm.z + gc);
                const float* r = prm.rays + ray * 11;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    p[k] = fmaf(__ldg(r + 3 + k), zz, __ldg(r + k));
                    vd[k] = __ldg(r + 8 + k);
                }
            };
            if (first < prm.n_groups) {
                load_point(first);
                write_pts_encoding(enc, pt, p);
                publish(kB2LoReady);
            }
            for (int g = first; g < prm.n_groups; g += stride) {
                uint8_t* save_g = kSave ? prm.save + (size_t)g * kSave2PairBytes : nullptr;
                for (int s = 0; s < kFwd2Steps; ++s) {
                    if (s <= 9) {
                        const Step2 st = kFwd2Dev[s];
                        const float2 c = __ldg(&g_sb[st.ch + ch]);
                        const float wa = (s == 7) ? __ldg(&g_wa[ch]) : 0.0f;
                        tr.mark(500 + s);
                        mbar_wait(bar(kB2AccReady + 0), n_acc++ & 1);
                        tc_fence_after_sync();
                        tr.mark(600 + s);
                        if (s >= 1 && s <= 8) mbar_wait(bar(kB2StageFree + q), n_sf++ & 1);
                        tr.mark(700 + s);
                        if (s == 6) write_dir_encoding(enc, pt, vd);       // gamma(x) is dead after L5
                        if (kSave) {
                            if (lane == 0) bulk_wait_read_all();
                            __syncwarp();
                        }
                        const uint32_t ta = tmem_lane + n_base;
                        if (prm.debug_flags & 1) { /* skip */ }
                        else if (s == 7) epilogue2<true, true>(ta, act, ch, n_base, c.x, c.y, wa, out_s, lane);
                        else if (s == 8) epilogue2<false, false>(ta, act, ch, n_base, c.x, c.y, 0.f, out_s, lane);
                        else epilogue2<true, false>(ta, act, ch, n_base, c.x, c.y, 0.f, out_s, lane);
                        if (kSave) store_pieces(save_g + (size_t)s * kActBytes);
                        if (s == 9) {
                            const int next = g + stride;
                            if (next < prm.n_groups) {
                                load_point(next);
                                write_pts_encoding(enc, pt, p);   // the direction stage of this group has completed
                            }
                        }
                        tr.mark(800 + s);
                        publish(kB2LoReady);
                        tr.mark(900 + s);
                    }
                }
            }
        } else {
            publish(kB2HiReady);      // D_hi is free at kernel start
            for (int g = first; g < prm.n_groups; g += stride) {
                uint8_t* save_g = kSave ? prm.save + (size_t)g * kSave2PairBytes : nullptr;
                for (int s = 0; s < kFwd2Steps; ++s) {
                    if (s <= 8) {
                        const Step2 st = kFwd2Dev[s];
                        const float2 c = __ldg(&g_sb[st.ch + ch]);
                        const float wa = (s == 7) ? __ldg(&g_wa[ch]) : 0.0f;
                        tr.mark(500 + s);
                        mbar_wait(bar(kB2AccReady + 1), n_acc++ & 1);
                        tc_fence_after_sync();
                        tr.mark(600 + s);
                        if (kSave) {
                            if (lane == 0) bulk_wait_read_all();
                            __syncwarp();
                        }
                        const uint32_t ta = tmem_lane + 256 + n_base;
                        if (prm.debug_flags & 1) { /* skip */ }
                        else if (s == 7) epilogue2<true, true>(ta, act, ch, n_base, c.x, c.y, wa, out_s, lane);
                        else if (s == 8) epilogue2<false, false>(ta, act, ch, n_base, c.x, c.y, 0.f, out_s, lane);
                        else epilogue2<true, false>(ta, act, ch, n_base, c.x, c.y, 0.f, out_s, lane);
                        if (kSave) store_pieces(save_g + (size_t)s * kActBytes);
                        tr.mark(800 + s);
                        publish(kB2HiReady);
                        tr.mark(900 + s);
                    } else if (s == 10) {
                        mbar_wait(bar(kB2AccReady + 1), n_acc++ & 1);
                        tc_fence_after_sync();
                        if (q == 0) {
                            // rgb head: lanes 0..2 of the accumulator hold the three logit rows for this warp's 128 points;
                            // sigma comes from the alpha partial sums accumulated in the L7 epilogues
                            const float2 c = __ldg(&g_sb[kChRgb + (lane < 3 ? lane : 0)]);
                            const long long g0 = (long long)g * kPairPoints + n_base;
#pragma unroll 1
                            for (int cc = 0; cc < 4; ++cc) {
                                uint32_t v[32];
                                tmem_ld32(tmem_lane + 256 + n_base + cc * 32, v);
                                tmem_ld_wait();
                                if (lane < 3) {
#pragma unroll
                                    for (int i = 0; i < 32; ++i) {
                                        const long long gi = g0 + cc * 32 + i;
                                        if (gi < prm.n_points) prm.raw[4 * gi + lane] = fmaf(__uint_as_float(v[i]), c.x, c.y);
                                    }
                                }
                            }
                            const float2 ca = __ldg(&g_sb[kChAlpha]);
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const int pl = n_base + i * 32 + lane;
                                const float sg = fmaf(out_s[pl], ca.x, ca.y);
                                out_s[pl] = 0.0f;
                                const long long gi = (long long)g * kPairPoints + pl;
                                if (gi < prm.n_points) prm.raw[4 * gi + 3] = sg;
                            }
                        }
                        publish(kB2HiReady);
                    }
                }
            }
        }
        if (kSave && lane == 0) bulk_wait_all();
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

}  // namespace nerfq

extern "C" int nerfq_mlp2_forward(const void* packed, const float* rays, const float* z, long long n_rays, int samples_per_ray,
                                  float* raw, void* save, int max_ctas, cudaStream_t stream) {
    using namespace nerfq;
    if (n_rays == 0) return 0;
    if (!packed || !rays || !z || !raw || n_rays < 0 || samples_per_ray <= 0) return -1;
    const long long n_points = n_rays * samples_per_ray;
    const int n_groups = (int)((n_points + kPairPoints - 1) / kPairPoints);
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (max_ctas > 0 && max_ctas < sms) sms = max_ctas;
    const int grid = n_groups < sms ? n_groups : sms;
    const char* dbg = getenv("NERFQ_DEBUG_FLAGS");
    Fwd2Params prm{(const uint8_t*)packed, rays, z, raw, (uint8_t*)save, n_points, samples_per_ray, n_groups, dbg ? atoi(dbg) : 0};
    if (save && !(prm.debug_flags & 8)) {
        if (cudaFuncSetAttribute(mlp2_forward_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kS2BytesFwd) != cudaSuccess) return -2;
        mlp2_forward_kernel<true><<<grid, kThreads2, kS2BytesFwd, stream>>>(prm);
    } else {
        if (cudaFuncSetAttribute(mlp2_forward_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kS2BytesFwd) != cudaSuccess) return -2;
        mlp2_forward_kernel<false><<<grid, kThreads2, kS2BytesFwd, stream>>>(prm);
    }
    return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

extern "C" unsigned long long nerfq_mlp2_save_bytes(long long n_points) {
    using namespace nerfq;
    const long long n_groups = (n_points + kPairPoints - 1) / kPairPoints;
    return (unsigned long long)n_groups * kSave2PairBytes;
}
