// Fused positional-encoding + NeRF MLP forward on tcgen05 tensor cores (sm_100a).
//
// Replaces, for one network and M = n_rays*S sample points:
//   pts = o + d*z                                   run_nerf.py:408,430
//   embed_fn / embeddirs_fn / cat                   run_nerf.py:52-59, run_nerf_helpers.py:18-67
//   NeRF.forward with ScaledLinear layers           utils.py:57-80, transforms.py:104-111
// Output: raw[M,4] = (rgb logits, sigma) as run_network returns it (run_nerf.py:61-63).
//
// One persistent CTA per SM iterates over groups of 256 points (mlp3_layout.h), processed as two independent halves of
// 128 points that are in flight together: per layer and half the tensor cores compute
//   D[o][n] = sum_k L[o][k] * X[n][k]      (L = integer weight levels, fp16; X = activations, fp16; D fp32 in TMEM)
// as two accumulators of 128 output channels x 128 points (four accumulators = all 512 TMEM columns); 16 KB weight
// chunks stream from L2 through a 4-slot ring of bulk async copies and each chunk feeds half A's MMAs, then half B's;
// the activation tile stays in shared memory and is rewritten in place.  A half has its own team of 8 epilogue warps and
// its own barriers, so while one half's epilogue runs (the tensor pipe has nothing to do for THAT half: the next layer
// needs its output) the pipe works on the other half.  Each epilogue
// warp owns 32 output channels (its TMEM lanes) x 64 points of its half's accumulators: a thread keeps the
// dequantisation constants of ITS channel in registers and applies  y = acc * (delta * s[o]) + b[o], ReLU, fp16
// conversion, storing 8 points per 16-byte shared-memory store into the next layer's operand tile.  The epilogue of
// channels 0..127 overlaps the MMAs of channels 128..255 and vice versa.  The alpha head is reduced on CUDA cores
// in the L7 epilogue; the rgb head is one more (3-of-128-row) MMA.  With `save` set the fp16 activations are also
// written to HBM for the backward pass in a chunk-major layout that makes every warp store 1 KB contiguous
// (mlp3_layout.h): half of a job's stores go straight from the registers that feed the shared-memory store, the other
// half is read back from the tile after the hand-over, which spreads the stores over the time the warps would wait
// for the next accumulator (the SM's store path is the limit while a job runs; kSaveInJob).  (Letting the bulk-copy
// engine read the tile back out of shared memory competes with the MMA operand reads: 11 B/clk/SM,
// profiles/r01_umma_rate2_bulk_store_vs_mma.log.)
#include <cuda_runtime.h>

#include <type_traits>

#include "mlp3_common.cuh"

namespace nerfq {

struct Fwd3Params {
    const uint8_t* packed;
    const float* rays;       // [n_rays, 11]
    const float* z;          // [n_rays * S]
    float* raw;              // [n_rays * S, 4]
    uint8_t* save;           // nullable; kSave3GroupBytes per group of 256 points
    long long n_points;
    int samples_per_ray;
    int n_groups;
    unsigned long long* dbg;     // tracing build only: 8 cycle counters per CTA
    int dbg_flags;               // tracing build only, ablations (results are then wrong, only the timing is of interest):
                                 // 1 skip TMEM loads, 2 skip operand stores, 4 skip conversion math, 8 publish without the
                                 // async-proxy fence, 16 skip the alpha-head reduction, 32 skip the encodings; bits 8.. = timed job
    Prog3Fwd prog;
};

#ifndef NERFQ_PREWAIT_SETUP
#define NERFQ_PREWAIT_SETUP 1
#endif
#ifndef NERFQ_PE_SPLIT
#define NERFQ_PE_SPLIT 1
#endif
#ifndef NERFQ_EPI_F32X2
#define NERFQ_EPI_F32X2 1
#endif
constexpr float kAlphaFix = 262144.0f;       // 2^18: fixed-point unit of the alpha-head sum (int32: +-8192 logit units)
#ifndef NERFQ_SAVE_IN_JOB
#define NERFQ_SAVE_IN_JOB 2
#endif
constexpr int kSaveInJob = NERFQ_SAVE_IN_JOB;      // chunks (of 4) whose saved activations are stored from registers inside the job; measured 1: 1.04, 2: 0.97, 3: 0.99, 4: 1.04 ms

template <bool kSave, bool kTrace>
__global__ void __launch_bounds__(kThreads3, 1) mlp3_forward_kernel(const __grid_constant__ Fwd3Params prm) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t sbase = opaque_u32(smem_u32(smem));
    const int warp = uniform_warp_idx();
    const int lane = threadIdx.x & 31;
    float* out_s = reinterpret_cast<float*>(smem + kS3Misc);
    auto bar = [&](int i) { return sbase + kS3Bars + 8u * i; };

    for (int i = threadIdx.x; i < 256; i += kThreads3) out_s[i] = 0.0f;
    // the CTA owns the SM (1 CTA/SM by shared-memory size) and allocates all 512 TMEM columns, so the allocation starts
    // at column 0; treating the base as a constant frees a register in every epilogue thread
    constexpr uint32_t tmem_base = 0;
    if (setup3(smem, sbase, warp) != tmem_base) __trap();
    const int first = blockIdx.x, stride = gridDim.x;
    const int n_iters = first < prm.n_groups ? (prm.n_groups - first + stride - 1) / stride : 0;

    if (warp >= kCtrlWarp0 && warp < kCtrlWarp0 + kCtrlWarps3) {
        const int cw = warp - kCtrlWarp0;
        // the control warp group hands registers to the four epilogue warp groups (64 / 112 per thread)
        asm volatile("setmaxnreg.dec.sync.aligned.u32 " NERFQ_REGS_CTRL3 ";");
        if (cw == 0 || cw == 2) {
            loader3(sbase, prm.packed + kOffFwd3Image, kFwd3Chunks, n_iters, cw >> 1);
        } else {
            if (n_iters > 0) issuer3<true, kTrace>(sbase, tmem_base, n_iters, (uint32_t)(cw >> 1), prm.dbg);
        }
    } else {
        // ================= epilogue warps =================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 " NERFQ_REGS_EPI3 ";");
        const int e = warp - kEpiWarp0;
        const int q = warp & 3, pq = e >> 2;
        const int team = e >> 3;                  // half-group A (points 0..127) or B (128..255): own accumulators and barriers
        const uint32_t tmem_lane = tmem_base + (uint32_t(q * 32) << 16) + team * 256 + (pq & 1) * 64;
        const uint32_t act = sbase + kS3Act;      // shared-space addresses
        const uint32_t enc = sbase + kS3Enc;
        const uint32_t out_sa = sbase + kS3Misc;
        const float2* g_sb = reinterpret_cast<const float2*>(prm.packed + kOffSB);
        const float* g_wa = reinterpret_cast<const float*>(prm.packed + kOffWAlpha);
        // the point whose encodings this thread (co-)writes: two threads per point for gamma(x)
        const int pt = (e >> 1) * 32 + lane, role = e & 1;
        uint32_t ph_acc0 = 0, ph_acc1 = 0, ph_sf = 0;
        unsigned long long t_acc = 0, t_sf = 0, t_job = 0, t_math = 0, t_sel = 0, t_sel_math = 0;
        const int j_sel = prm.dbg_flags >> 8;
        const bool tracing = kTrace && e == 5 && lane == 0;
        const bool ab_ld = kTrace && (prm.dbg_flags & 1), ab_st = kTrace && (prm.dbg_flags & 2), ab_math = kTrace && (prm.dbg_flags & 4);
        const bool ab_fence = kTrace && (prm.dbg_flags & 8), ab_alpha = kTrace && (prm.dbg_flags & 16), ab_pe = kTrace && (prm.dbg_flags & 32);

        auto publish = [&](int which) {
            if (!ab_fence) fence_proxy_async_smem();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(which));
        };
        float p[3], vd[3], vd_next[3];
        auto load_point = [&](int g, float* vdir) {
            const long long gidx = (long long)g * kGroupPts + pt;
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
        auto load_consts = [&](int j, float2& c, float& wa) {      // epilogue constants of job j for this thread's channel
            const Job3 jn = prm.prog.job[j];
            c = make_float2(0.f, 0.f);
            wa = 0.0f;
            if (!(jn.flags & JB_FINAL)) c = __ldg(&g_sb[jn.ch + 32 * q + lane]);
            // alpha head weight of this channel as fixed-point logit units: level * (delta*scale of the head) * kAlphaFix
            if (jn.flags & JB_ALPHA)
                wa = __ldg(&g_wa[((jn.flags & JB_HI_HALF) ? 128 : 0) + 32 * q + lane]) * __ldg(&g_sb[kChAlpha]).x * kAlphaFix;
        };

        float2 c_next = make_float2(0.f, 0.f);
        float wa_next = 0.0f;
        if (n_iters > 0) {
            load_point(first, vd);
            write_pe_half(enc, pt, role, p);
            publish(kB3ActLo + team);        // initial encodings
            publish(kB3ActHi + team);        // D_hi is free at kernel start
            load_consts(0, c_next, wa_next);
        }
        for (int g = first; g < prm.n_groups; g += stride) {
            uint8_t* save_g = kSave ? prm.save + (size_t)g * kSave3GroupBytes : nullptr;
            // the next group's point is fetched now so that its encoding can be written the moment the tile is free
            if (g + stride < prm.n_groups) load_point(g + stride, vd_next);
#pragma unroll 1
            for (int j = 0; j < kFwd3Jobs; ++j) {
                const Job3 jb = prm.prog.job[j];
                const uint32_t f = jb.flags;
                const uint32_t hi = (f & JB_HI_HALF) ? 1u : 0u;
                // per-thread geometry re-derived from the thread index inside the loop (a volatile read the compiler cannot
                // hoist): hoisted address parts were spilled, and a local-memory reload per job costs hundreds of cycles here
                uint32_t tid_j;
                asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid_j));
                const uint32_t q_j = (tid_j >> 5) & 3u, pq_j = ((tid_j >> 5) - (uint32_t)kEpiWarp0) >> 2;
                const uint32_t ch = 128u * hi + 32u * q_j + (tid_j & 31u);   // this thread's channel within the layer
                const float2 c = c_next;
                const float wa = wa_next;
#if NERFQ_PE_SPLIT
                // Encoding work taken OFF the dependency chain.  A warp that gets here has finished job (5, hi): every MMA of L5, the
                // last reader of this half's gamma(x), has completed, and the accumulator this job waits for cannot be ready before
                // the issuer has seen that job's hand-over and run two more weight chunks -- a guaranteed gap of > 1000 cycles in
                // which the even warps write gamma(d) into columns 0..31 and the odd warps the NEXT group's gamma(x) columns 32..63
                // (the views layer multiplies those columns by zero weights: any finite value may stand there).
                if ((f & JB_DIR_BEFORE) && !ab_pe) {
                    const int row = (int)((((tid_j >> 5) - (uint32_t)kEpiWarp0) >> 1) * 32u + (tid_j & 31u));
                    if (role == 0) write_dir_enc(enc, row, vd);
                    else if (g + stride < prm.n_groups) write_pe_half(enc, row, 1, p);
                }
#endif
#if NERFQ_PREWAIT_SETUP
                // everything that does not depend on the accumulator is issued BEFORE the wait for it: the next job's constants, the
                // operand row and saved-activation addresses (a "low" job finds a gap here; after the wait these would sit on the chain)
                const uint32_t ta = tmem_lane + ((f & JB_ACC_HI) ? 128u : 0u);
                load_consts(j + 1 < kFwd3Jobs ? j + 1 : 0, c_next, wa_next);    // in flight during this job
                const uint32_t row_addr = act + (ch >> 3) * kKGroup3 + pq_j * kNGroup3 + (ch & 7u) * 128u;
                const uint32_t swz = (ch & 7u) << 4;
                // the row starts on a 128-byte boundary (tile 1 KB-aligned, all terms multiples of 128): row + (x ^ swz) == (row | swz) ^ x
                // for the 16-byte slots x = 0..7 << 4 -- one instruction per store address instead of two
                const uint32_t st_base = row_addr | swz;
                uint8_t* save_ch = kSave ? save_g + (size_t)jb.slot * kSave3SlotBytes + save3_offset(0, ch) : nullptr;
#endif
                unsigned long long t0 = 0;
                if (tracing) t0 = clock64();
                if (f & JB_ACC_HI) { mbar_wait_acc(bar(kB3AccReady + 2 * team + 1), ph_acc1); ph_acc1 ^= 1; }
                else { mbar_wait_acc(bar(kB3AccReady + 2 * team), ph_acc0); ph_acc0 ^= 1; }
                tc_fence_after_sync();
                if (tracing) { const unsigned long long t1 = clock64(); t_acc += t1 - t0; t0 = t1; }
#if !NERFQ_PREWAIT_SETUP
                const uint32_t ta = tmem_lane + ((f & JB_ACC_HI) ? 128u : 0u);
                load_consts(j + 1 < kFwd3Jobs ? j + 1 : 0, c_next, wa_next);    // in flight during this job
#endif

                if (f & JB_FINAL) {
                    // rgb head: lanes 0..2 of the accumulator hold the three logit rows; sigma comes from the alpha
                    // partial sums accumulated in the L7 epilogues.
                    if (q == 0) {
                        const float2 cr = __ldg(&g_sb[kChRgb + (lane < 3 ? lane : 0)]);
                        const float2 ca = __ldg(&g_sb[kChAlpha]);
                        const long long g0 = (long long)g * kGroupPts + pq * 64;
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
                            const float sg = fmaf((float)ld_shared_s32(out_sa + 4 * pl), 1.0f / kAlphaFix, ca.y);
                            st_shared_f32(out_sa + 4 * pl, 0.0f);
                            const long long gi = g0 + cc * 32 + lane;
                            if (gi < prm.n_points) prm.raw[4 * gi + 3] = sg;
                        }
                    }
                    publish(kB3ActHi + team);
                    if (tracing) { const unsigned long long dt = clock64() - t0; t_job += dt; if (j == j_sel) t_sel += dt; }
                    continue;
                }

#if !NERFQ_PE_SPLIT
                if (f & JB_DIR_BEFORE) {        // gamma(x) is dead once L5 has been accumulated
                    if (role == 0 && !ab_pe) write_dir_enc(enc, (int)((((tid_j >> 5) - (uint32_t)kEpiWarp0) >> 1) * 32u + (tid_j & 31u)), vd);
                }
#endif
                // ---- 4 chunks of 16 points: TMEM -> y = acc*es + b -> (ReLU) -> fp16 -> operand tile ----
                // The load of chunk i+1 is in flight while chunk i is converted and stored.
                const bool relu = f & JB_RELU;
#if !NERFQ_PREWAIT_SETUP
                const uint32_t row_addr = act + (ch >> 3) * kKGroup3 + pq_j * kNGroup3 + (ch & 7u) * 128u;
                const uint32_t swz = (ch & 7u) << 4;
                const uint32_t st_base = row_addr | swz;
                uint8_t* save_ch = kSave ? save_g + (size_t)jb.slot * kSave3SlotBytes + save3_offset(0, ch) : nullptr;
#endif
                uint32_t va[16], vb[16];
                // 0: scalar fma everywhere; 1: packed pairs in the kernel without save (measured on one box, profiles/r02_ab_pe_fast_f32x2.log:
                // forward 0.683 -> 0.666 ms, but forward + save 0.916 -> 0.932 ms -- there the stores' register hazards decide); 2: both
                constexpr bool kEpiF32x2 = NERFQ_EPI_F32X2 == 2 || (NERFQ_EPI_F32X2 == 1 && !kSave);
                const uint64_t cx2 = pack_f32x2(c.x, c.x), cy2 = pack_f32x2(c.y, c.y);
                // (relu_c / alpha_c are compile-time tags: the job's kind is decided ONCE, before its four chunks, instead of by
                // two branches inside every chunk -- ncu r02 had 15 % of the epilogue's stall samples in branch resolution)
                auto process = [&](const uint32_t (&v)[16], int cc, auto relu_c, auto alpha_c) {
                    float y[16];
                    uint32_t pk[8];
                    if (kEpiF32x2) {
                        // two accumulator values per instruction (fma.rn.f32x2, SASS FFMA2; same IEEE results as the scalar form):
                        // a job is bound by the length of its instruction stream, not by the FP32 pipe
#pragma unroll
                        for (int i = 0; i < 8; ++i) fma_f32x2(y[2 * i], y[2 * i + 1], v[2 * i], v[2 * i + 1], cx2, cy2);
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i) y[i] = fmaf(__uint_as_float(v[i]), c.x, c.y);
                    }
                    if (ab_math) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) pk[i] = v[2 * i] ^ v[2 * i + 1];
                    } else if (decltype(relu_c)::value) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) pk[i] = cvt_pack_f16_relu(y[2 * i], y[2 * i + 1]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) pk[i] = cvt_pack_f16(y[2 * i], y[2 * i + 1]);
                    }
                    if (cc == 0) {
                        // first store of the job: the last readers of these channels must be done
                        if (f & JB_WAIT_SF) {
                            unsigned long long t1 = 0;
                            if (tracing) t1 = clock64();
                            mbar_wait(bar(kB3StageFree + 2 * team + (q >> 1)), ph_sf);
                            ph_sf ^= 1;
                            if (tracing) t_sf += clock64() - t1;
                        }
                    }
                    if (!ab_st) {
#pragma unroll
                        for (int k = 0; k < 2; ++k)
                            st_shared_v4(st_base ^ (uint32_t)((cc * 2 + k) << 4), pk[4 * k], pk[4 * k + 1], pk[4 * k + 2], pk[4 * k + 3]);
                    }
                    // the same 16 values go to the saved-activation slot: the warp's 32 channels are adjacent, 1 KB per store
                    if (kSave && cc < kSaveInJob) st_global_v8(save_ch + save3_offset(pq * 4 + cc, 0), pk[0], pk[1], pk[2], pk[3], pk[4], pk[5], pk[6], pk[7]);
                    if (decltype(alpha_c)::value && !ab_alpha) {
                        // Alpha head over this warp's 32 channels (L7 is a ReLU layer: the head sees max(y, 0)), in fixed
                        // point: each term  max(y,0) * level * (delta*scale) of the sigma logit is rounded to 2^-18, the
                        // warp sum is an integer shuffle butterfly, and the eight partial sums of a point (4 warps x 2 halves) meet
                        // in native integer shared-memory atomics.  Integer addition is associative, so sigma -- and with it
                        // every pixel -- is bit-reproducible (a float butterfly + float atomics, a compare-and-swap loop
                        // on sm_100, cost the same and were not).  Range +-8192 logit units (kAlphaFix; two's-complement partial
                        // sums may wrap, only the total must fit), rounding error 1.8e-5 rms over the 256 terms.
                        int term[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) term[i] = __float2int_rn(fmaxf(y[i], 0.0f) * wa);
                        // transpose-and-add butterfly over the warp's 32 channels: lanes 2l and 2l+1 end up with point l's sum
                        const int mine = column_reduce16_i3(term, lane);
                        if (!(lane & 1)) red_shared_add_s32(out_sa + 4 * (pq * 64 + cc * 16 + (lane >> 1)), mine);
                    }
                };
                // The load of chunk i+1 is in flight while chunk i is converted and stored.
                auto run4 = [&](auto relu_c, auto alpha_c) {
                    if (ab_ld) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) va[i] = vb[i] = (uint32_t)(i + j);
                        process(va, 0, relu_c, alpha_c);
                        process(vb, 1, relu_c, alpha_c);
                        process(va, 2, relu_c, alpha_c);
                        process(vb, 3, relu_c, alpha_c);
                    } else {
                        tmem_ld16(ta, va);
                        tmem_ld_wait();
                        tmem_ld16(ta + 16, vb);
                        process(va, 0, relu_c, alpha_c);
                        tmem_ld_wait();
                        tmem_ld16(ta + 32, va);
                        process(vb, 1, relu_c, alpha_c);
                        tmem_ld_wait();
                        tmem_ld16(ta + 48, vb);
                        process(va, 2, relu_c, alpha_c);
                        tmem_ld_wait();
                        process(vb, 3, relu_c, alpha_c);
                    }
                };
                if (f & JB_ALPHA) run4(std::true_type{}, std::true_type{});            // L7 (ReLU) + alpha head
                else if (relu) run4(std::true_type{}, std::false_type{});
                else run4(std::false_type{}, std::false_type{});                       // feature layer: no activation
                if (tracing) { const unsigned long long dt = clock64() - t0; t_math += dt; if (j == j_sel) t_sel_math += dt; }
                if (f & JB_PE_AFTER) {          // the direction stage of this group has been accumulated
                    if (g + stride < prm.n_groups) {
#if NERFQ_PE_SPLIT
                        // columns 0..31 held gamma(d) until the views layer was accumulated: the two warps of a point write 16 of them each
                        if (!ab_pe) write_pe_lo16(enc, (int)((((tid_j >> 5) - (uint32_t)kEpiWarp0) >> 1) * 32u + (tid_j & 31u)), role, p);
#else
                        if (!ab_pe) write_pe_half(enc, (int)((((tid_j >> 5) - (uint32_t)kEpiWarp0) >> 1) * 32u + (tid_j & 31u)), role, p);
#endif
#pragma unroll
                        for (int k = 0; k < 3; ++k) vd[k] = vd_next[k];
                    }
                }
                publish((hi ? kB3ActHi : kB3ActLo) + team);
                if (kSave && kSaveInJob < 4) {
                    // The SM's store path to L2 (~30 B/clk) is the limit while a job runs, and idle while the warps wait for
                    // the next accumulator: the last chunks' saved activations are read back from the operand tile (only this
                    // thread rewrites these rows, at the next layer) and stored after the hand-over.
#pragma unroll
                    for (int cc = kSaveInJob; cc < 4; ++cc) {
                        const uint4 v0 = ld_shared_v4(st_base ^ (uint32_t)((cc * 2) << 4));
                        const uint4 v1 = ld_shared_v4(st_base ^ (uint32_t)((cc * 2 + 1) << 4));
                        st_global_v8(save_ch + save3_offset(pq * 4 + cc, 0), v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w);
                    }
                }
                if (tracing) { const unsigned long long dt = clock64() - t0; t_job += dt; if (j == j_sel) t_sel += dt; }
            }
        }
        if (tracing && prm.dbg) {
            if (e == 5) { prm.dbg[8 * blockIdx.x + 4] = t_acc; prm.dbg[8 * blockIdx.x + 5] = t_sf; prm.dbg[8 * blockIdx.x + 6] = t_job; prm.dbg[8 * blockIdx.x + 7] = t_math; prm.dbg[8 * 148 + 32 * blockIdx.x] = t_sel; prm.dbg[8 * 148 + 32 * blockIdx.x + 1] = t_sel_math; }
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

}  // namespace nerfq

static unsigned long long* g_trace3 = nullptr;
static int g_trace3_flags = 0;
// Profiling aid (not part of include/nerfq.h): when set, launches use the tracing instantiation of the kernels, which
// writes 8 cycle counters per CTA into this device buffer (see profiles/trace_v3.py).
extern "C" void nerfq_mlp_set_trace(unsigned long long* buf, int flags) { g_trace3 = buf; g_trace3_flags = flags; }

extern "C" int nerfq_mlp_forward(const void* packed, const float* rays, const float* z, long long n_rays, int samples_per_ray,
                                  float* raw, void* save, int max_ctas, cudaStream_t stream) {
    using namespace nerfq;
    if (n_rays == 0) return 0;
    if (!packed || !rays || !z || !raw || n_rays < 0 || samples_per_ray <= 0) return -1;
    static const Prog3Fwd prog = make_prog3_fwd();
    const long long n_points = n_rays * samples_per_ray;
    const int n_groups = (int)((n_points + kGroupPts - 1) / kGroupPts);
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (max_ctas > 0 && max_ctas < sms) sms = max_ctas;
    const int grid = n_groups < sms ? n_groups : sms;
    Fwd3Params prm{(const uint8_t*)packed, rays, z, raw, (uint8_t*)save, n_points, samples_per_ray, n_groups, g_trace3, g_trace3_flags, prog};
    auto launch = [&](auto kernel) -> int {
        if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kS3Bytes) != cudaSuccess) return -2;
        kernel<<<grid, kThreads3, kS3Bytes, stream>>>(prm);
        return cudaGetLastError() == cudaSuccess ? 0 : -3;
    };
    if (g_trace3) return save ? launch(mlp3_forward_kernel<true, true>) : launch(mlp3_forward_kernel<false, true>);
    return save ? launch(mlp3_forward_kernel<true, false>) : launch(mlp3_forward_kernel<false, false>);
}

extern "C" unsigned long long nerfq_mlp_save_bytes(long long n_points) {
    using namespace nerfq;
    const long long n_groups = (n_points + kGroupPts - 1) / kGroupPts;
    return (unsigned long long)n_groups * kSave3GroupBytes;
}
