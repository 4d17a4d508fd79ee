// Fused positional-encoding + NeRF MLP forward for RENDERING (no saved activations): CTA pairs, points on TMEM lanes.
//
// Same function, inputs, outputs and weight image as mlp3_fwd.cu (which documents the reference lines replaced).
// The single-CTA kernel keeps one 256-point group per SM and its tensor pipe idles while that group's epilogue runs
// (54 % active); the first pair schedule (mlp4_fwd.cu, channels on lanes) kept two groups in flight but had to move half
// of every layer's output across the pair through distributed shared memory, which cost more than the overlap won.
// Here the GEMM is transposed,
//     D[n][o] = sum_k X[n][k] * L[o][k]        A = activations (K-major, M = points), B = weight chunk (K-major, N = channels),
// as one M = 256 x N = 256 x K = 16 MMA over both SMs of a cluster (tcgen05 cta_group::2): CTA r supplies the rows of ITS
// 128 points and the weight rows of output channels 128 r .. (its half of every 16 KB chunk, the per-SM weight stream is
// unchanged), and receives, for its 128 points, ALL 256 output channels in its own TMEM.  A point's next-layer operand
// row is therefore written entirely by its own CTA: nothing crosses the pair but barrier signals.  Two groups (A, B) of
// 256 points share the cluster; the issuer alternates  layer t of A, layer t of B, layer t+1 of A ...  and the epilogue
// warps follow one phase behind, so one group's epilogue runs under the other group's MMAs.
//
// Epilogue: a thread owns one POINT (TMEM lane) and, per job, 64 output channels (one 128-byte row of a K block of the
// operand tile); the per-channel constants {delta*scale, bias} are warp-uniform loads.  The alpha head becomes an
// in-thread dot product (no cross-lane reduction); its four partial sums per point meet in fixed-point integer atomics,
// so results stay bit-reproducible.  tests/cuda/umma_probe2.cu checks the operand / accumulator ownership used here.
#include <cuda_runtime.h>
#include <stdlib.h>

#include "mlp3_common.cuh"
#include "mlp4_fwd.h"

namespace nerfq {

constexpr int kThreads5 = kThreads3;
constexpr uint32_t kBlock5Bytes = 128 * 128;                       // one K block of the operand tile: 128 points x 64 channels
constexpr uint32_t kAct5Bytes = 4 * kBlock5Bytes;                  // 256 channels
constexpr uint32_t kS5Act = 0;                                     // [2 groups]
constexpr uint32_t kS5Ring = kS5Act + 2 * kAct5Bytes;              // 4 x 16 KB weight chunks
constexpr uint32_t kS5Enc = kS5Ring + kSlots3 * kChunk3Bytes;      // [2 groups] 128 points x 64 columns
constexpr uint32_t kS5Alpha = kS5Enc + 2 * kBlock5Bytes;           // int[2][128]: alpha-head sums in 2^-20 logit units
constexpr uint32_t kS5Bars = kS5Alpha + 2 * 128 * 4;
constexpr uint32_t kS5TmemPtr = kS5Bars + 8 * 24;
constexpr uint32_t kS5Bytes = kS5TmemPtr + 16;
static_assert(kS5Bytes <= 232448, "shared memory budget");

constexpr int kB5WFull = 0;       // [4] own loader's bulk copies
constexpr int kB5WEmpty = 4;      // [4] multicast commit
constexpr int kB5PeerFull = 8;    // [4] CTA 0 only: CTA 1's slot is full
constexpr int kB5Act = 12;        // [2] CTA 0: 16 own epilogue warps + 1 arrival forwarded from CTA 1
constexpr int kB5LocalAct = 14;   // [2] CTA 1: its 16 epilogue warps
constexpr int kB5AccReady = 16;   // [2] multicast commit

enum : uint32_t { J5_RELU = 1, J5_ALPHA = 2, J5_FINAL = 4, J5_DIR_BEFORE = 8, J5_PE_AFTER = 16, J5_HALF = 32 };
struct Pass5 {
    uint16_t chunk0[2];      // first chunk (index into the forward weight image) of this layer for CTA rank 0 / 1
    uint8_t n_act, n_enc;    // chunks contracted against the activation tile / the encoding tile
    uint16_t flags;
    int16_t ch;              // channel base of the layer's epilogue constants
    int16_t pad;
};
struct Prog5Fwd { Pass5 pass[kFwd3Steps]; };

static Prog5Fwd make_prog5_fwd() {
    Prog5Fwd p{};
    int base = 0;
    for (int s = 0; s < kFwd3Steps; ++s) {
        const Step3& st = kFwd3[s];
        Pass5& e = p.pass[s];
        const int per_half = (st.kh + st.kp) / 2;
        e.chunk0[0] = (uint16_t)base;
        e.chunk0[1] = (uint16_t)(st.halves == 2 ? base + per_half : base);      // 128-output layers: columns 128.. are unused
        e.n_act = (uint8_t)(st.kh / 2);
        e.n_enc = (uint8_t)(st.kp / 2);
        uint32_t f = 0;
        if (st.relu) f |= J5_RELU;
        if (s == 7) f |= J5_ALPHA;
        if (s == 10) f |= J5_FINAL;
        if (s == 6) f |= J5_DIR_BEFORE;
        if (s == 9) f |= J5_PE_AFTER;
        if (st.halves == 1) f |= J5_HALF;
        e.flags = (uint16_t)f;
        e.ch = st.ch;
        base += st.halves * per_half;
    }
    return p;
}

struct Fwd5Params {
    const uint8_t* packed;
    const float* rays;       // [n_rays, 11]
    const float* z;          // [n_rays * S]
    float* raw;              // [n_rays * S, 4]
    long long n_points;
    int samples_per_ray;
    int n_groups;
    unsigned long long* dbg;     // tracing build: 8 cycle counters per CTA
    Prog5Fwd prog;
};

constexpr uint32_t kIdesc5 = umma_idesc(256, 256, false);          // A and B K-major

// Epilogue constants of the network being evaluated, in CONSTANT memory: a thread owns a point here, so {delta*scale, bias}
// differ per accumulator column and are warp-uniform -- through the constant cache a warp-uniform read costs one issue slot,
// as global or shared loads each one writes back 512 bytes of registers through the load-store unit (measured: 2 k cycles
// per job).  Refreshed from the packed buffer by the launch function (device-to-device copy on the same stream).
__constant__ float2 c_sb5[kNumChannels + 4];
__constant__ float c_wa5[256];

template <bool kTrace>
__global__ void __launch_bounds__(kThreads5, 1) mlp5_forward_kernel(const __grid_constant__ Fwd5Params prm) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = opaque_u32(smem_u32(smem));
    const int warp = uniform_warp_idx();
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    auto bar = [&](int i) { return sbase + kS5Bars + 8u * i; };
    if (sbase & 1023u) __trap();

    for (int i = threadIdx.x; i < 256; i += kThreads5) reinterpret_cast<int*>(smem + kS5Alpha)[i] = 0;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kSlots3; ++i) {
            mbar_init(bar(kB5WFull + i), 1);
            mbar_init(bar(kB5WEmpty + i), 1);
            mbar_init(bar(kB5PeerFull + i), 1);
        }
        for (int x = 0; x < 2; ++x) {
            mbar_init(bar(kB5Act + x), kEpiWarps3 + 1);
            mbar_init(bar(kB5LocalAct + x), kEpiWarps3);
            mbar_init(bar(kB5AccReady + x), 1);
        }
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc2(sbase + kS5TmemPtr, 512);
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after_sync();
    constexpr uint32_t tmem_base = 0;         // the pair allocates all 512 columns of both SMs
    if (*reinterpret_cast<volatile uint32_t*>(smem + kS5TmemPtr) != tmem_base) __trap();

    // pair p of the cluster's iteration `it`: groups 2p and 2p+1 (256 points each: 128 per CTA; the second may not exist)
    const int n_pairs = (prm.n_groups + 1) >> 1;
    const int cid = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    const int n_iters = cid < n_pairs ? (n_pairs - cid + n_clusters - 1) / n_clusters : 0;

    if (warp == 0 || warp == 2) {
        // ================= weight loaders (both CTAs): this CTA's 128 output channels of every layer, once per group =================
        const int which = warp >> 1;
        const uint8_t* img = prm.packed + kOffFwd3Image;
        uint32_t seq = 0;
        for (int it = 0; it < n_iters; ++it) {
#pragma unroll 1
            for (int s = 0; s < kFwd3Steps; ++s) {
                const Pass5 ps = prm.prog.pass[s];
                const int n = ps.n_act + ps.n_enc;
                const uint8_t* src0 = img + (size_t)ps.chunk0[rank] * kChunk3Bytes;
#pragma unroll 1
                for (int c2 = 0; c2 < 2 * n; ++c2, ++seq) {          // group A's pass, then group B's: the same chunks again
                    if ((int)(seq % kLoaders3) != which) continue;
                    const int c = c2 < n ? c2 : c2 - n;
                    const uint32_t slot = seq & (kSlots3 - 1), par = (seq >> 2) & 1;
                    mbar_wait(bar(kB5WEmpty) + 8 * slot, par ^ 1);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(bar(kB5WFull) + 8 * slot, kChunk3Bytes);
                        bulk_g2s(sbase + kS5Ring + slot * kChunk3Bytes, src0 + (size_t)c * kChunk3Bytes, kChunk3Bytes, bar(kB5WFull) + 8 * slot);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 1 && rank == 1) {
        // ================= CTA 1: tell the issuer in CTA 0 that this CTA's half of a chunk has landed =================
        const uint32_t peer_full0 = mapa_u32(bar(kB5PeerFull), 0);
        uint32_t seq = 0;
        for (int it = 0; it < n_iters; ++it) {
#pragma unroll 1
            for (int s = 0; s < kFwd3Steps; ++s) {
                const Pass5 ps = prm.prog.pass[s];
                const int n2 = 2 * (ps.n_act + ps.n_enc);
#pragma unroll 1
                for (int c = 0; c < n2; ++c, ++seq) {
                    const uint32_t slot = seq & (kSlots3 - 1), par = (seq >> 2) & 1;
                    mbar_wait(bar(kB5WFull) + 8 * slot, par);
                    if (elect_one()) mbar_arrive_remote(peer_full0 + 8 * slot);      // no data of this thread to release
                    __syncwarp();
                }
            }
        }
    } else if (warp == 3 && rank == 1) {
        // ================= CTA 1: forward "my 16 epilogue warps finished group x's job" to CTA 0 =================
        // One cluster-scope release per job instead of sixteen (it compiles to a GPU-scope MEMBAR); cumulativity carries the
        // epilogue warps' shared-memory writes, which this warp acquired through the local barrier.
        const uint32_t act0 = mapa_u32(bar(kB5Act), 0);
        uint32_t ph[2] = {0, 0};
        const int n_jobs = n_iters * kFwd3Steps + (n_iters > 0 ? 1 : 0);        // + the initial encodings
#pragma unroll 1
        for (int i = 0; i < n_jobs; ++i) {
#pragma unroll 1
            for (int x = 0; x < 2; ++x) {
                mbar_wait(bar(kB5LocalAct + x), ph[x]);
                ph[x] ^= 1;
                if (elect_one()) mbar_arrive_cluster(act0 + 8 * x);
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ================= CTA 0: MMA issuer for the pair =================
        const uint32_t bar0 = sbase + kS5Bars;
        const uint64_t b_desc0 = umma_smem_desc(sbase + kS5Ring, 512, SWZ_64B);
        uint32_t seq = 0, ph_act[2] = {0, 0};
        unsigned long long t_begin = 0, t_wf = 0, t_pf = 0, t_act = 0;
        if (kTrace) t_begin = clock64();
        // one weight chunk = 64 k: four MMAs; A advances 32 bytes inside its 128-byte rows, B 32 bytes inside a stage / 8 KB per stage
        auto chunk = [&](uint32_t d_tmem, uint64_t a, uint32_t accumulate, uint32_t done_bar) {
            const uint32_t slot = seq & (kSlots3 - 1), par = (seq >> 2) & 1;
            ++seq;
            unsigned long long t0 = 0, t1 = 0;
            if (kTrace) t0 = clock64();
            mbar_wait(bar0 + 8 * (kB5WFull + slot), par);
            if (kTrace) t1 = clock64();
            mbar_wait(bar0 + 8 * (kB5PeerFull + slot), par);       // written by the peer's copy engine for its own tensor core
            if (kTrace) { const unsigned long long t2 = clock64(); t_wf += t1 - t0; t_pf += t2 - t1; }
            tc_fence_after_sync();
            if (elect_one()) {
                const uint64_t bd = b_desc0 + slot * (kChunk3Bytes >> 4);
                umma_ss2(d_tmem, a, bd, kIdesc5, accumulate);
                umma_ss2(d_tmem, a + 2, bd + 2, kIdesc5, 1u);
                umma_ss2(d_tmem, a + 4, bd + (kStage3Bytes >> 4), kIdesc5, 1u);
                umma_ss2(d_tmem, a + 6, bd + (kStage3Bytes >> 4) + 2, kIdesc5, 1u);
                umma_commit2_mc(bar0 + 8 * (kB5WEmpty + slot), 3);
                if (done_bar) umma_commit2_mc(done_bar, 3);
            }
            __syncwarp();
        };
        for (int it = 0; it < n_iters; ++it) {
#pragma unroll 1
            for (int s = 0; s < kFwd3Steps; ++s) {
                const Pass5 ps = prm.prog.pass[s];
                const int n_act = ps.n_act, n_enc = ps.n_enc;
#pragma unroll 1
                for (int x = 0; x < 2; ++x) {
                    // group x's operand rows written in both CTAs (or its encodings, for the first layer), accumulator drained
                    unsigned long long t0 = 0;
                    if (kTrace) t0 = clock64();
                    mbar_wait_cluster(bar0 + 8 * (kB5Act + x), ph_act[x]);
                    ph_act[x] ^= 1;
                    tc_fence_after_sync();
                    if (kTrace) t_act += clock64() - t0;
                    const uint32_t d_tmem = tmem_base + 256u * x;
                    const uint32_t acc_bar = bar0 + 8 * (kB5AccReady + x);
                    uint64_t a = umma_smem_desc(sbase + kS5Act + x * kAct5Bytes, 1024, SWZ_128B);
#pragma unroll 1
                    for (int j = 0; j < n_act; ++j) {
                        chunk(d_tmem, a, j > 0 ? 1u : 0u, (j + 1 == n_act && n_enc == 0) ? acc_bar : 0u);
                        a += kBlock5Bytes >> 4;
                    }
                    if (n_enc) chunk(d_tmem, umma_smem_desc(sbase + kS5Enc + x * kBlock5Bytes, 1024, SWZ_128B), n_act > 0 ? 1u : 0u, acc_bar);
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
        const int q = warp & 3, cq = e >> 2;             // TMEM lane quarter (32 points); quarter of the 256 output channels
        const uint32_t n = 32u * q + lane;               // this thread's point within the CTA's 128
        const uint32_t tmem_lane = tmem_base + (uint32_t(q * 32) << 16) + cq * 64;
        const float2* g_sb = reinterpret_cast<const float2*>(prm.packed + kOffSB);
        const float* g_wa = reinterpret_cast<const float*>(prm.packed + kOffWAlpha);
        const uint32_t alpha_sa = sbase + kS5Alpha;
        const uint32_t act_bar = rank == 0 ? bar(kB5Act) : bar(kB5LocalAct);
        // encodings: two threads per point; threads of warps 0..7 serve group A's 128 local points, 8..15 group B's
        const int xe = e >> 3, lp = ((e >> 1) & 3) * 32 + lane, role = e & 1;
        const uint32_t enc_mine = sbase + kS5Enc + xe * kBlock5Bytes;
        uint32_t ph_acc[2] = {0, 0};
        unsigned long long t_acc = 0, t_job = 0, t_pub = 0;
        const bool tracing = kTrace && e == 5 && lane == 0;

        auto publish = [&](int x) {
            fence_proxy_async_smem();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(act_bar + 8 * x);
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
                const Pass5 ps = prm.prog.pass[s];
                const uint32_t f = ps.flags;
                const bool active = !(f & J5_HALF) || cq < 2;          // 128-output layers fill columns 0..127 only
                const int cst = ps.ch + 64 * cq;                       // first of this warp's 64 channels in c_sb5
                const float alpha_es = (f & J5_ALPHA) ? __ldg(&g_sb[kChAlpha]).x * 1048576.0f : 0.0f;
#pragma unroll 1
                for (int x = 0; x < 2; ++x) {
                    const int g = 2 * pair + x;
                    unsigned long long t0 = 0;
                    if (tracing) t0 = clock64();
                    mbar_wait(bar(kB5AccReady + x), ph_acc[x]);
                    ph_acc[x] ^= 1;
                    tc_fence_after_sync();
                    if (tracing) { const unsigned long long t1 = clock64(); t_acc += t1 - t0; t0 = t1; }
                    const uint32_t ta = tmem_lane + 256u * x;
                    const long long gi = (long long)g * kGroupPts + rank * 128 + n;       // this thread's point

                    if (f & J5_FINAL) {
                        // rgb head: columns 0..2 of the accumulator; sigma from the alpha sums of the L7 jobs
                        if (cq == 0) {
                            uint32_t v[16];
                            tmem_ld16(ta, v);
                            tmem_ld_wait();
                            const float2 c0 = __ldg(&g_sb[kChRgb + 0]), c1 = __ldg(&g_sb[kChRgb + 1]), c2 = __ldg(&g_sb[kChRgb + 2]);
                            const float2 ca = __ldg(&g_sb[kChAlpha]);
                            const int sum = ld_shared_s32(alpha_sa + 4 * (x * 128 + n));
                            st_shared_f32(alpha_sa + 4 * (x * 128 + n), 0.0f);
                            if (gi < prm.n_points) {
                                float4 o;
                                o.x = fmaf(__uint_as_float(v[0]), c0.x, c0.y);
                                o.y = fmaf(__uint_as_float(v[1]), c1.x, c1.y);
                                o.z = fmaf(__uint_as_float(v[2]), c2.x, c2.y);
                                o.w = fmaf((float)sum, 1.0f / 1048576.0f, ca.y);
                                *reinterpret_cast<float4*>(prm.raw + 4 * gi) = o;
                            }
                        }
                        publish(x);
                        continue;
                    }
                    if ((f & J5_DIR_BEFORE) && x == xe && role == 0) write_dir_enc(enc_mine, lp, vd);       // gamma(x) is dead after L5

                    if (active) {
                        // ---- 4 chunks of 16 channels of this thread's point: TMEM -> y = acc*es + b -> (ReLU) -> fp16 -> its operand row ----
                        const bool relu = f & J5_RELU;
                        const uint32_t row_addr = sbase + kS5Act + x * kAct5Bytes + cq * kBlock5Bytes + n * 128u;
                        const uint32_t swz = (n & 7u) << 4;
                        int alpha_acc = 0;
                        uint32_t va[16], vb[16];
                        auto process = [&](const uint32_t (&v)[16], int cc) {
                            uint32_t pk[8];
                            float dot = 0.0f;
#pragma unroll
                            for (int i = 0; i < 16; i += 2) {
                                const float2 ca = c_sb5[cst + 16 * cc + i], cb = c_sb5[cst + 16 * cc + i + 1];      // warp-uniform
                                float y0 = fmaf(__uint_as_float(v[i]), ca.x, ca.y);
                                float y1 = fmaf(__uint_as_float(v[i + 1]), cb.x, cb.y);
                                pk[i >> 1] = relu ? cvt_pack_f16_relu(y0, y1) : cvt_pack_f16(y0, y1);       // ReLU inside the conversion
                                if (f & J5_ALPHA) {
                                    y0 = fmaxf(y0, 0.0f);
                                    y1 = fmaxf(y1, 0.0f);
                                    dot = fmaf(y0, c_wa5[64 * cq + 16 * cc + i], dot);
                                    dot = fmaf(y1, c_wa5[64 * cq + 16 * cc + i + 1], dot);
                                }
                            }
#pragma unroll
                            for (int k = 0; k < 2; ++k)
                                st_shared_v4(row_addr + ((uint32_t)((cc * 2 + k) << 4) ^ swz), pk[4 * k], pk[4 * k + 1], pk[4 * k + 2], pk[4 * k + 3]);
                            if (f & J5_ALPHA) alpha_acc += __float2int_rn(dot * alpha_es);       // 16 terms of the sigma logit, 2^-20 units
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
                        if (f & J5_ALPHA) red_shared_add_s32(alpha_sa + 4 * (x * 128 + n), alpha_acc);
                    }
                    if ((f & J5_PE_AFTER) && x == xe && more) {          // the direction stage of this group has been accumulated
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

unsigned long long* g_trace5 = nullptr;

// Launch for the dispatcher in mlp4_fwd.cu.
int mlp5_forward_launch(const void* packed, const float* rays, const float* z, long long n_rays, int samples_per_ray, float* raw,
                        int max_ctas, cudaStream_t stream) {
    static const Prog5Fwd prog = make_prog5_fwd();
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
    if (cudaMemcpyToSymbolAsync(c_sb5, (const uint8_t*)packed + kOffSB, sizeof(float2) * kNumChannels, 0, cudaMemcpyDeviceToDevice, stream) != cudaSuccess ||
        cudaMemcpyToSymbolAsync(c_wa5, (const uint8_t*)packed + kOffWAlpha, sizeof(float) * 256, 0, cudaMemcpyDeviceToDevice, stream) != cudaSuccess)
        return -3;
    Fwd5Params prm{(const uint8_t*)packed, rays, z, raw, n_points, samples_per_ray, n_groups, g_trace5, prog};
    auto kernel = g_trace5 ? mlp5_forward_kernel<true> : mlp5_forward_kernel<false>;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kS5Bytes) != cudaSuccess) return -2;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(kThreads5);
    cfg.dynamicSmemBytes = kS5Bytes;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, prm) == cudaSuccess ? 0 : -3;
}

}  // namespace nerfq

// Profiling aid (not part of include/nerfq.h): 8 cycle counters per CTA {issuer: total, wait own slot, wait peer slot,
// wait operand; epilogue warp 5: wait accumulator, job, hand-over} from the tracing instantiation.
extern "C" void nerfq_mlp5_set_trace(unsigned long long* buf) { nerfq::g_trace5 = buf; }
