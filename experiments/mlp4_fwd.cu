// Fused positional-encoding + NeRF MLP forward, CTA-pair schedule (tcgen05 cta_group::2, sm_100a).
//
// Same function, inputs, outputs, weight image and saved-activation layout as mlp3_fwd.cu (which documents the
// reference lines replaced); what changes is how the work is laid over the SMs.  The single-CTA kernel keeps one
// 256-point group per SM, whose accumulators fill TMEM: while a layer's epilogue runs, the tensor pipe has nothing to
// do (measured 53 % active).  Here a CLUSTER OF TWO CTAs shares two groups, A and B:
//
//   * every MMA is M = 256 x N = 256 x K = 16 over both SMs: CTA r supplies weight rows 128 r .. (its half of the
//     16 KB chunks -- the per-SM weight stream is unchanged) and the operand columns of ITS 128 points of the group,
//     and receives output channels 128 r .. of all 256 points in its own TMEM (256 columns per group, two groups);
//   * the issuer (one thread of CTA 0) alternates  layer t of A, layer t of B, layer t+1 of A, ...  and the 16
//     epilogue warps of each CTA alternate the same way one phase behind, so the epilogue of one group runs under the
//     MMAs of the other;
//   * an epilogue thread owns one output channel (TMEM lane) and 64 points of a group; the fp16 row it produces is an
//     operand row of the next layer, which lives in the shared memory of the CTA that owns those POINTS: half of the
//     rows are written locally, half into the peer through distributed shared memory (st.shared::cluster);
//   * hand-over barriers are cluster-scope: "operand rows of group X written and accumulator X drained" counts the
//     32 epilogue warps of both CTAs on CTA 0's barrier; tcgen05.commit multicasts "accumulator ready" and "weight
//     slot free" to both CTAs; CTA 1 forwards its "weight slot full" to CTA 0.
//   tests/cuda/umma_probe2.cu checks these building blocks against an exact integer GEMM.
#include <cuda_runtime.h>
#include <stdlib.h>

#include "mlp3_common.cuh"
#include "mlp4_fwd.h"

namespace nerfq {

constexpr int kThreads4 = kThreads3;
constexpr uint32_t kAct4Bytes = 256 * 128 * 2;        // MN-major operand tile of one group in one CTA: 256 channels x 128 points
constexpr uint32_t kEnc4Bytes = 128 * 128;            // K-major encoding tile: 128 points x 64 columns
constexpr uint32_t kKGroup4 = 2048;                   // 8 channels x 128 points x 2 B
constexpr uint32_t kS4Act = 0;                                     // [2 groups]
constexpr uint32_t kS4Ring = kS4Act + 2 * kAct4Bytes;              // 4 x 16 KB weight chunks
constexpr uint32_t kS4Enc = kS4Ring + kSlots3 * kChunk3Bytes;      // [2 groups]
constexpr uint32_t kS4Alpha = kS4Enc + 2 * kEnc4Bytes;             // float[2][256]: this CTA's partial alpha-head sums
constexpr uint32_t kS4Bars = kS4Alpha + 2 * 256 * 4;
constexpr uint32_t kS4TmemPtr = kS4Bars + 8 * 16;
constexpr uint32_t kS4Bytes = kS4TmemPtr + 16;
static_assert(kS4Bytes <= 232448, "shared memory budget");

constexpr int kB4WFull = 0;       // [4] own loader's bulk copies
constexpr int kB4WEmpty = 4;      // [4] multicast commit
constexpr int kB4PeerFull = 8;    // [4] CTA 0 only: CTA 1's slot is full
constexpr int kB4Act = 12;        // [2] CTA 0 only: 32 arrivals (16 epilogue warps x 2 CTAs)
constexpr int kB4AccReady = 14;   // [2] multicast commit

enum : uint32_t { J4_RELU = 1, J4_ALPHA = 2, J4_FINAL = 4, J4_DIR_BEFORE = 8, J4_PE_AFTER = 16, J4_RANK0_ONLY = 32, J4_SHIP_ALPHA = 64 };
struct Pass4 {
    uint16_t chunk0[2];      // first chunk (index into the forward weight image) of this layer for CTA rank 0 / 1
    uint8_t n_act, n_enc;    // chunks contracted against the activation tile / the encoding tile
    uint16_t flags;
    int16_t ch;              // channel base of the layer's epilogue constants
    int16_t slot;            // saved-activation slot, -1 none
};
struct Prog4Fwd { Pass4 pass[kFwd3Steps]; };

static Prog4Fwd make_prog4_fwd() {
    Prog4Fwd p{};
    int base = 0;
    for (int s = 0; s < kFwd3Steps; ++s) {
        const Step3& st = kFwd3[s];
        Pass4& e = p.pass[s];
        const int per_half = (st.kh + st.kp) / 2;
        e.chunk0[0] = (uint16_t)base;
        e.chunk0[1] = (uint16_t)(st.halves == 2 ? base + per_half : base);      // single-half layers: rank 1's rows are unused
        e.n_act = (uint8_t)(st.kh / 2);
        e.n_enc = (uint8_t)(st.kp / 2);
        uint32_t f = 0;
        if (st.relu) f |= J4_RELU;
        if (s == 7) f |= J4_ALPHA;
        if (s == 8) f |= J4_SHIP_ALPHA;
        if (s == 10) f |= J4_FINAL;
        if (s == 6) f |= J4_DIR_BEFORE;
        if (s == 9) f |= J4_PE_AFTER;
        if (st.halves == 1) f |= J4_RANK0_ONLY;
        e.flags = (uint16_t)f;
        e.ch = st.ch;
        e.slot = (int16_t)(s <= 9 ? s : -1);
        base += st.halves * per_half;
    }
    return p;
}

struct Fwd4Params {
    const uint8_t* packed;
    const float* rays;       // [n_rays, 11]
    const float* z;          // [n_rays * S]
    float* raw;              // [n_rays * S, 4]
    uint8_t* save;           // nullable
    long long n_points;
    int samples_per_ray;
    int n_groups;
    unsigned long long* dbg;     // tracing build: 8 cycle counters per CTA
    Prog4Fwd prog;
};

__device__ __forceinline__ uint64_t umma_desc_mn4(uint32_t saddr) {     // MN-major SWIZZLE_128B tile, 128 points per CTA
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
    d |= static_cast<uint64_t>((kNGroup3 >> 4) & 0x3FFF) << 16;
    d |= static_cast<uint64_t>((kKGroup4 >> 4) & 0x3FFF) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(SWZ_128B) << 61;
    return d;
}
constexpr uint32_t kIdesc4BK = umma_idesc(256, 256, false);
constexpr uint32_t kIdesc4BMN = kIdesc4BK | (1u << 16);

template <bool kSave, bool kTrace>
__global__ void __launch_bounds__(kThreads4, 1) mlp4_forward_kernel(const __grid_constant__ Fwd4Params prm) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = opaque_u32(smem_u32(smem));
    const int warp = uniform_warp_idx();
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    auto bar = [&](int i) { return sbase + kS4Bars + 8u * i; };
    if (sbase & 1023u) __trap();

    for (int i = threadIdx.x; i < 512; i += kThreads4) reinterpret_cast<float*>(smem + kS4Alpha)[i] = 0.0f;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kSlots3; ++i) {
            mbar_init(bar(kB4WFull + i), 1);
            mbar_init(bar(kB4WEmpty + i), 1);
            mbar_init(bar(kB4PeerFull + i), 1);
        }
        for (int x = 0; x < 2; ++x) {
            mbar_init(bar(kB4Act + x), 2 * kEpiWarps3);
            mbar_init(bar(kB4AccReady + x), 1);
        }
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc2(sbase + kS4TmemPtr, 512);
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after_sync();
    constexpr uint32_t tmem_base = 0;         // the pair allocates all 512 columns of both SMs
    if (*reinterpret_cast<volatile uint32_t*>(smem + kS4TmemPtr) != tmem_base) __trap();

    // pair p of the cluster's iteration `it`: groups 2p and 2p+1 (256 points each; the second may not exist)
    const int n_pairs = (prm.n_groups + 1) >> 1;
    const int cid = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    const int n_iters = cid < n_pairs ? (n_pairs - cid + n_clusters - 1) / n_clusters : 0;

    if (warp == 0 || warp == 2) {
        // ================= weight loaders (both CTAs): this CTA's half of every layer, once per group =================
        const int which = warp >> 1;
        const uint8_t* img = prm.packed + kOffFwd3Image;
        uint32_t seq = 0;
        for (int it = 0; it < n_iters; ++it) {
#pragma unroll 1
            for (int s = 0; s < kFwd3Steps; ++s) {
                const Pass4 ps = prm.prog.pass[s];
                const int n = ps.n_act + ps.n_enc;
                const uint8_t* src0 = img + (size_t)ps.chunk0[rank] * kChunk3Bytes;
#pragma unroll 1
                for (int c2 = 0; c2 < 2 * n; ++c2, ++seq) {          // group A's pass, then group B's: the same chunks again
                    if ((int)(seq % kLoaders3) != which) continue;
                    const int c = c2 < n ? c2 : c2 - n;
                    const uint32_t slot = seq & (kSlots3 - 1), par = (seq >> 2) & 1;
                    mbar_wait(bar(kB4WEmpty) + 8 * slot, par ^ 1);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(bar(kB4WFull) + 8 * slot, kChunk3Bytes);
                        bulk_g2s(sbase + kS4Ring + slot * kChunk3Bytes, src0 + (size_t)c * kChunk3Bytes, kChunk3Bytes, bar(kB4WFull) + 8 * slot);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 1 && rank == 1) {
        // ================= CTA 1: tell the issuer in CTA 0 that this CTA's half of a chunk has landed =================
        const uint32_t peer_full0 = mapa_u32(bar(kB4PeerFull), 0);
        uint32_t seq = 0;
        for (int it = 0; it < n_iters; ++it) {
#pragma unroll 1
            for (int s = 0; s < kFwd3Steps; ++s) {
                const Pass4 ps = prm.prog.pass[s];
                const int n2 = 2 * (ps.n_act + ps.n_enc);
#pragma unroll 1
                for (int c = 0; c < n2; ++c, ++seq) {
                    const uint32_t slot = seq & (kSlots3 - 1), par = (seq >> 2) & 1;
                    mbar_wait(bar(kB4WFull) + 8 * slot, par);
                    if (elect_one()) mbar_arrive_remote(peer_full0 + 8 * slot);      // no data of this thread to release
                    __syncwarp();
                }
            }
        }
    } else if (warp == 1) {
        // ================= CTA 0: MMA issuer for the pair =================
        const uint32_t bar0 = sbase + kS4Bars;
        const uint64_t a_desc0 = umma_smem_desc(sbase + kS4Ring, 512, SWZ_64B);
        uint32_t seq = 0, ph_act[2] = {0, 0};
        unsigned long long t_begin = 0, t_wf = 0, t_pf = 0, t_act = 0;
        if (kTrace) t_begin = clock64();
        auto chunk = [&](uint32_t d_tmem, uint64_t b, uint32_t idesc, uint32_t kstep, uint32_t stage, uint32_t accumulate, uint32_t done_bar) {
            const uint32_t slot = seq & (kSlots3 - 1), par = (seq >> 2) & 1;
            ++seq;
            unsigned long long t0 = 0, t1 = 0;
            if (kTrace) t0 = clock64();
            mbar_wait(bar0 + 8 * (kB4WFull + slot), par);
            if (kTrace) t1 = clock64();
            // the peer's half was written by its bulk-copy engine for its own tensor core: nothing for this thread to acquire
            mbar_wait(bar0 + 8 * (kB4PeerFull + slot), par);
            if (kTrace) { const unsigned long long t2 = clock64(); t_wf += t1 - t0; t_pf += t2 - t1; }
            tc_fence_after_sync();
            if (elect_one()) {
                const uint64_t ad = a_desc0 + slot * (kChunk3Bytes >> 4);
                umma_ss2(d_tmem, ad, b, idesc, accumulate);
                umma_ss2(d_tmem, ad + 2, b + kstep, idesc, 1u);
                umma_ss2(d_tmem, ad + (kStage3Bytes >> 4), b + stage, idesc, 1u);
                umma_ss2(d_tmem, ad + (kStage3Bytes >> 4) + 2, b + stage + kstep, idesc, 1u);
                umma_commit2_mc(bar0 + 8 * (kB4WEmpty + slot), 3);
                if (done_bar) umma_commit2_mc(done_bar, 3);
            }
            __syncwarp();
        };
        for (int it = 0; it < n_iters; ++it) {
#pragma unroll 1
            for (int s = 0; s < kFwd3Steps; ++s) {
                const Pass4 ps = prm.prog.pass[s];
                const int n_act = ps.n_act, n_enc = ps.n_enc;
#pragma unroll 1
                for (int x = 0; x < 2; ++x) {
                    // operand rows of group x written by both CTAs (or its encodings, for the first layer) and its
                    // accumulator drained
                    unsigned long long t0 = 0;
                    if (kTrace) t0 = clock64();
                    mbar_wait_cluster(bar0 + 8 * (kB4Act + x), ph_act[x]);
                    ph_act[x] ^= 1;
                    tc_fence_after_sync();
                    if (kTrace) t_act += clock64() - t0;
                    const uint32_t d_tmem = tmem_base + 256u * x;
                    const uint32_t acc_bar = bar0 + 8 * (kB4AccReady + x);
                    uint64_t b = umma_desc_mn4(sbase + kS4Act + x * kAct4Bytes);
#pragma unroll 1
                    for (int j = 0; j < n_act; ++j) {
                        chunk(d_tmem, b, kIdesc4BMN, (2u * kKGroup4) >> 4, (4u * kKGroup4) >> 4, j > 0 ? 1u : 0u,
                              (j + 1 == n_act && n_enc == 0) ? acc_bar : 0u);
                        b += (8u * kKGroup4) >> 4;
                    }
                    if (n_enc) chunk(d_tmem, umma_smem_desc(sbase + kS4Enc + x * kEnc4Bytes, 1024, SWZ_128B), kIdesc4BK, 32u >> 4, 64u >> 4,
                                     n_act > 0 ? 1u : 0u, acc_bar);
                }
            }
        }
        if (kTrace && prm.dbg && lane == 0) {
            unsigned long long* o = prm.dbg + 8 * blockIdx.x;
            o[0] = clock64() - t_begin; o[1] = t_wf; o[2] = t_pf; o[3] = t_act;
        }
    } else if (warp >= kCtrlWarps3) {
        // ================= epilogue warps (both CTAs) =================
        const int e = warp - kCtrlWarps3;
        const int q = warp & 3, pq = e >> 2;            // TMEM lane quarter; 64-point quarter of the group's 256 points
        const uint32_t chl = 32u * q + lane;             // this thread's channel within the CTA's 128
        const uint32_t row = 128u * rank + chl;          // ... = operand row (K index) of the next layer
        const uint32_t tmem_lane = tmem_base + (uint32_t(q * 32) << 16) + pq * 64;
        const uint32_t dst_rank = (uint32_t)pq >> 1;     // the CTA that owns these 64 points
        const float2* g_sb = reinterpret_cast<const float2*>(prm.packed + kOffSB);
        const float* g_wa = reinterpret_cast<const float*>(prm.packed + kOffWAlpha);
        const bool dst_local = dst_rank == rank;                                 // (distributed shared memory moves ~20 B/clk/SM: only
        const uint32_t act_dst = dst_local ? sbase + kS4Act : mapa_u32(sbase + kS4Act, dst_rank);      //  the peer's half takes that path)
        const uint32_t alpha_sa = sbase + kS4Alpha;
        const uint32_t act_bar0 = mapa_u32(bar(kB4Act), 0);
        // encodings: two threads per point; threads of warps 0..7 serve group A's 128 local points, 8..15 group B's
        const int xe = e >> 3, lp = ((e >> 1) & 3) * 32 + lane, role = e & 1;
        const uint32_t enc_mine = sbase + kS4Enc + xe * kEnc4Bytes;
        uint32_t ph_acc[2] = {0, 0};
        unsigned long long t_acc = 0, t_job = 0, t_pub = 0;
        const bool tracing = kTrace && e == 5 && lane == 0;

        auto publish = [&](int x) {
            fence_proxy_async_all();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(act_bar0 + 8 * x);
        };
        float p[3], vd[3], vd_next[3];
        auto load_point = [&](int pair, float* vdir) {        // the point whose encodings this thread (co-)writes
            const long long gidx = ((long long)(2 * pair + xe) * kGroupPts) + rank * 128 + lp;
            const long long gc = gidx < prm.n_points ? gidx : prm.n_points - 1;
            const long long ray = gc / prm.samples_per_ray;
            const float zz = __ldg(prm.z + gc);
            const float* r = prm.rays + ray * 11;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                p[k] = fmaf(__ldg(r + 3 + k), zz, __ldg(r + k));
                vdir[k] = __ldg(r + 8 + k);
            }
        };
        if (n_iters > 0) {
            load_point(cid, vd);
            write_pe_half(enc_mine, lp, role, p);
            publish(0);
            publish(1);
        }
        for (int it = 0; it < n_iters; ++it) {
            const int pair = cid + it * n_clusters;
            const bool more = it + 1 < n_iters;
            if (more) load_point(pair + n_clusters, vd_next);
#pragma unroll 1
            for (int s = 0; s < kFwd3Steps; ++s) {
                const Pass4 ps = prm.prog.pass[s];
                const uint32_t f = ps.flags;
                const bool active = !(f & J4_RANK0_ONLY) || rank == 0;
                float2 c = make_float2(0.f, 0.f);
                float wa = 0.0f;
                if (active && !(f & J4_FINAL)) c = __ldg(&g_sb[ps.ch + ((f & J4_RANK0_ONLY) ? chl : row)]);
                if (f & J4_ALPHA) wa = __ldg(&g_wa[row]);
#pragma unroll 1
                for (int x = 0; x < 2; ++x) {
                    const int g = 2 * pair + x;
                    unsigned long long t0 = 0;
                    if (tracing) t0 = clock64();
                    mbar_wait(bar(kB4AccReady + x), ph_acc[x]);
                    ph_acc[x] ^= 1;
                    tc_fence_after_sync();
                    if (tracing) { const unsigned long long t1 = clock64(); t_acc += t1 - t0; t0 = t1; }
                    const uint32_t ta = tmem_lane + 256u * x;
                    const long long g0 = (long long)g * kGroupPts + pq * 64;

                    if (f & J4_SHIP_ALPHA) {
                        // all of this CTA's L7 epilogue warps have added their partial alpha sums (their arrivals precede
                        // this accumulator); CTA 1 parks its partials in raw[..][3], where CTA 0's final job picks them up
                        if (rank == 1 && e < 8) {
                            const int pt = e * 32 + lane;
                            const long long gi = (long long)g * kGroupPts + pt;
                            const float v = ld_shared_f32(alpha_sa + 4 * (x * 256 + pt));
                            st_shared_f32(alpha_sa + 4 * (x * 256 + pt), 0.0f);
                            if (gi < prm.n_points) prm.raw[4 * gi + 3] = v;
                        }
                    }
                    if (f & J4_FINAL) {
                        // rgb head: lanes 0..2 of CTA 0's accumulator hold the three logit rows
                        if (rank == 0 && q == 0) {
                            const float2 cr = __ldg(&g_sb[kChRgb + (lane < 3 ? lane : 0)]);
                            const float2 ca = __ldg(&g_sb[kChAlpha]);
#pragma unroll 1
                            for (int cc = 0; cc < 2; ++cc) {
                                uint32_t v[32];
                                tmem_ld32(ta + cc * 32, v);
                                tmem_ld_wait();
                                if (lane < 3) {
#pragma unroll
                                    for (int i = 0; i < 32; ++i) {
                                        const long long gi = g0 + cc * 32 + i;
                                        if (gi < prm.n_points) prm.raw[4 * gi + lane] = fmaf(__uint_as_float(v[i]), cr.x, cr.y);
                                    }
                                }
                                const int pl = pq * 64 + cc * 32 + lane;
                                const long long gi = g0 + cc * 32 + lane;
                                float sum = ld_shared_f32(alpha_sa + 4 * (x * 256 + pl));
                                st_shared_f32(alpha_sa + 4 * (x * 256 + pl), 0.0f);
                                if (gi < prm.n_points) {
                                    sum += __ldcg(prm.raw + 4 * gi + 3);          // CTA 1's half of the channels
                                    prm.raw[4 * gi + 3] = fmaf(sum, ca.x, ca.y);
                                }
                            }
                        }
                        publish(x);
                        continue;
                    }
                    if ((f & J4_DIR_BEFORE) && x == xe && role == 0) write_dir_enc(enc_mine, lp, vd);       // gamma(x) is dead after L5

                    if (active) {
                        // ---- 4 chunks of 16 points: TMEM -> y = acc*es + b -> (ReLU) -> fp16 -> operand row in the owner of the points ----
                        const bool relu = f & J4_RELU;
                        const uint32_t row_addr = act_dst + x * kAct4Bytes + (row >> 3) * kKGroup4 + (pq & 1) * kNGroup3 + (row & 7u) * 128u;
                        const uint32_t swz = (row & 7u) << 4;
                        uint8_t* save_ch = nullptr;
                        if (kSave && g < prm.n_groups)
                            save_ch = prm.save + (size_t)g * kSave3GroupBytes + (size_t)ps.slot * kSave3SlotBytes + save3_offset(pq * 4, row);
                        uint32_t va[16], vb[16];
                        auto process = [&](const uint32_t (&v)[16], int cc) {
                            float y[16];
                            uint32_t pk[8];
#pragma unroll
                            for (int i = 0; i < 16; ++i) y[i] = fmaf(__uint_as_float(v[i]), c.x, c.y);
                            if (relu) {
#pragma unroll
                                for (int i = 0; i < 8; ++i) pk[i] = cvt_pack_f16_relu(y[2 * i], y[2 * i + 1]);
                            } else {
#pragma unroll
                                for (int i = 0; i < 8; ++i) pk[i] = cvt_pack_f16(y[2 * i], y[2 * i + 1]);
                            }
                            if (dst_local) {
#pragma unroll
                                for (int k = 0; k < 2; ++k)
                                    st_shared_v4(row_addr + ((uint32_t)((cc * 2 + k) << 4) ^ swz), pk[4 * k], pk[4 * k + 1], pk[4 * k + 2], pk[4 * k + 3]);
                            } else {
#pragma unroll
                                for (int k = 0; k < 2; ++k)
                                    st_cluster_v4(row_addr + ((uint32_t)((cc * 2 + k) << 4) ^ swz), pk[4 * k], pk[4 * k + 1], pk[4 * k + 2], pk[4 * k + 3]);
                            }
                            if (kSave && save_ch) st_global_v8(save_ch + save3_offset(cc, 0), pk[0], pk[1], pk[2], pk[3], pk[4], pk[5], pk[6], pk[7]);
                            if (f & J4_ALPHA) {        // L7 is a ReLU layer: the head sees max(y, 0)
#pragma unroll
                                for (int i = 0; i < 16; ++i) y[i] = fmaxf(y[i], 0.0f) * wa;
                                const float sm = column_reduce16_3(y, lane);
                                if (!(lane & 1)) red_shared_add_f32(alpha_sa + 4 * (x * 256 + pq * 64 + cc * 16 + (lane >> 1)), sm);
                            }
                        };
                        tmem_ld16(ta, va);
                        tmem_ld_wait();
                        tmem_ld16(ta + 16, vb);
                        process(va, 0);
                        tmem_ld_wait();
                        tmem_ld16(ta + 32, va);
                        process(vb, 1);
                        tmem_ld_wait();
                        tmem_ld16(ta + 48, vb);
                        process(va, 2);
                        tmem_ld_wait();
                        process(vb, 3);
                    }
                    if ((f & J4_PE_AFTER) && x == xe && more) {          // the direction stage of this group has been accumulated
                        write_pe_half(enc_mine, lp, role, p);
#pragma unroll
                        for (int k = 0; k < 3; ++k) vd[k] = vd_next[k];
                    }
                    if (tracing) { const unsigned long long t1 = clock64(); t_job += t1 - t0; t0 = t1; }
                    publish(x);
                    if (tracing) t_pub += clock64() - t0;
                }
            }
        }
        if (tracing && prm.dbg) {
            unsigned long long* o = prm.dbg + 8 * blockIdx.x;
            o[4] = t_acc; o[5] = t_job; o[6] = t_pub;
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) tmem_dealloc2(tmem_base, 512);
}

}  // namespace nerfq

static unsigned long long* g_trace4 = nullptr;
// Profiling aid (not part of include/nerfq.h): 8 cycle counters per CTA {issuer: total, wait own slot, wait peer slot,
// wait operand; epilogue warp 5: wait accumulator, job, hand-over} from the tracing instantiation.
extern "C" void nerfq_mlp4_set_trace(unsigned long long* buf) { g_trace4 = buf; }

// The single-CTA kernel (mlp3_fwd.cu) is the default; NERFQ_MLP_FWD=4 selects this pair schedule, NERFQ_MLP_FWD=5 the
// points-on-lanes pair schedule of mlp5_fwd.cu for calls without `save` (both experiments, both slower).  Measured on B200
// (profiles/r01_mlp4_pair_schedule_trace.log): results identical, but 1.53 ms against 0.85 ms -- the operand rows that
// cross the pair (32 KB per layer and group each way) move at distributed-shared-memory speed, and the cluster-scope
// release that hands them over compiles to a GPU-scope MEMBAR: ~2.0 k cycles per hand-over and 2.6 k per epilogue job
// instead of ~1.7 k, more than the overlap with the other group's MMAs wins back.
extern "C" int nerfq_mlp_forward(const void* packed, const float* rays, const float* z, long long n_rays, int samples_per_ray,
                                  float* raw, void* save, int max_ctas, cudaStream_t stream) {
    using namespace nerfq;
    if (n_rays == 0) return 0;
    if (!packed || !rays || !z || !raw || n_rays < 0 || samples_per_ray <= 0) return -1;
    static const int which = [] { const char* e = getenv("NERFQ_MLP_FWD"); return e ? e[0] - '0' : 0; }();
    if (which == 5 && !save && !mlp3_forward_tracing()) return mlp5_forward_launch(packed, rays, z, n_rays, samples_per_ray, raw, max_ctas, stream);
    const bool use_v4 = which == 4;
    if (!use_v4 || mlp3_forward_tracing()) return mlp3_forward_launch(packed, rays, z, n_rays, samples_per_ray, raw, save, max_ctas, stream);
    static const Prog4Fwd prog = make_prog4_fwd();
    const long long n_points = n_rays * samples_per_ray;
    const int n_groups = (int)((n_points + kGroupPts - 1) / kGroupPts);
    const int n_pairs = (n_groups + 1) / 2;
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (max_ctas > 0 && max_ctas < sms) sms = max_ctas;
    int clusters = sms / 2;
    if (clusters < 1) clusters = 1;
    if (n_pairs < clusters) clusters = n_pairs;
    Fwd4Params prm{(const uint8_t*)packed, rays, z, raw, (uint8_t*)save, n_points, samples_per_ray, n_groups, g_trace4, prog};
    auto launch = [&](auto kernel) -> int {
        if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kS4Bytes) != cudaSuccess) return -2;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(2 * clusters);
        cfg.blockDim = dim3(kThreads4);
        cfg.dynamicSmemBytes = kS4Bytes;
        cfg.stream = stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        return cudaLaunchKernelEx(&cfg, kernel, prm) == cudaSuccess ? 0 : -3;
    };
    if (g_trace4) return save ? launch(mlp4_forward_kernel<true, true>) : launch(mlp4_forward_kernel<false, true>);
    return save ? launch(mlp4_forward_kernel<true, false>) : launch(mlp4_forward_kernel<false, false>);
}
