// Fused positional-encoding + NeRF MLP forward on tcgen05 tensor cores (sm_100a), third generation.
//
// Replaces, for one network and M = n_rays*S sample points:
//   pts = o + d*z                                   run_nerf.py:408,430
//   embed_fn / embeddirs_fn / cat                   run_nerf.py:52-59, run_nerf_helpers.py:18-67
//   NeRF.forward with ScaledLinear layers           utils.py:57-80, transforms.py:104-111
// Output: raw[M,4] = (rgb logits, sigma) as run_network returns it (run_nerf.py:61-63).
//
// One persistent CTA per SM iterates over groups of 256 points (mlp3_layout.h).  Per layer the tensor cores compute
//   D[o][n] = sum_k L[o][k] * X[n][k]      (L = integer weight levels, fp16; X = activations, fp16; D fp32 in TMEM)
// as two accumulators of 128 output channels x 256 points; 16 KB weight chunks stream from L2 through a 4-slot
// ring of bulk async copies, the activation tile stays in shared memory and is rewritten in place.  Each of the
// 16 epilogue warps owns 32 output channels (its TMEM lanes) x 64 points of every accumulator: a thread keeps the
// dequantisation constants of ITS channel in registers and applies  y = acc * (delta * s[o]) + b[o], ReLU, fp16
// conversion, storing 8 points per 16-byte shared-memory store into the next layer's operand tile.  The epilogue of
// channels 0..127 overlaps the MMAs of channels 128..255 and vice versa.  The alpha head is reduced on CUDA cores
// in the L7 epilogue; the rgb head is one more (3-of-128-row) MMA.  With `save` set every operand tile is also
// streamed to HBM (bulk stores, one per 1 KB piece a warp owns) for the backward pass.
#include <cuda_runtime.h>

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
    Prog3Fwd prog;
};

template <bool kSave>
__global__ void __launch_bounds__(kThreads3, 1) mlp3_forward_kernel(const __grid_constant__ Fwd3Params prm) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t sbase = smem_u32(smem);
    const int warp = uniform_warp_idx();
    const int lane = threadIdx.x & 31;
    float* out_s = reinterpret_cast<float*>(smem + kS3Misc);
    auto bar = [&](int i) { return sbase + kS3Bars + 8u * i; };

    for (int i = threadIdx.x; i < 256; i += kThreads3) out_s[i] = 0.0f;
    const uint32_t tmem_base = setup3(smem, sbase, warp);
    const int first = blockIdx.x, stride = gridDim.x;
    const int n_iters = first < prm.n_groups ? (prm.n_groups - first + stride - 1) / stride : 0;

    if (warp == 0) {
        loader3(sbase, prm.packed + kOffFwd3Image, kFwd3Chunks, n_iters);
    } else if (warp == 1) {
        if (n_iters > 0) issuer3<kFwd3Jobs>(sbase, tmem_base, prm.prog.half, n_iters, true);
    } else if (warp >= kCtrlWarps3) {
        // ================= epilogue warps =================
        const int e = warp - kCtrlWarps3;
        const int q = e & 3, pq = e >> 2;
        const uint32_t tmem_lane = tmem_base + (uint32_t(q * 32) << 16) + pq * 64;
        uint8_t* act = smem + kS3Act;
        uint8_t* enc = smem + kS3Enc;
        const float2* g_sb = reinterpret_cast<const float2*>(prm.packed + kOffSB);
        const float* g_wa = reinterpret_cast<const float*>(prm.packed + kOffWAlpha);
        // the point whose encodings this thread (co-)writes: two threads per point for gamma(x)
        const int pt = (e >> 1) * 32 + lane, role = e & 1;
        uint32_t ph_acc0 = 0, ph_acc1 = 0, ph_sf = 0;

        auto publish = [&](int which) {
            fence_proxy_async_smem();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(which));
        };
        float p[3], vd[3];
        auto load_point = [&](int g) {
            const long long gidx = (long long)g * kGroupPts + pt;
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

        if (n_iters > 0) {
            load_point(first);
            write_pe_half(enc, pt, role, p);
            publish(kB3ActLo);        // initial encodings
            publish(kB3ActHi);        // D_hi is free at kernel start
        }
        for (int g = first; g < prm.n_groups; g += stride) {
            uint8_t* save_g = kSave ? prm.save + (size_t)g * kSave3GroupBytes : nullptr;
#pragma unroll 1
            for (int j = 0; j < kFwd3Jobs; ++j) {
                const Job3 jb = prm.prog.job[j];
                const uint32_t f = jb.flags;
                const uint32_t hi = (f & JB_HI_HALF) ? 1u : 0u;
                const uint32_t ch = 128u * hi + 32u * q + lane;              // this thread's channel within the layer
                float2 c = make_float2(0.f, 0.f);
                float wa = 0.0f;
                if (!(f & JB_FINAL)) c = __ldg(&g_sb[jb.ch + 32 * q + lane]);
                if (f & JB_ALPHA) wa = __ldg(&g_wa[ch]);
                if (f & JB_ACC_HI) { mbar_wait(bar(kB3AccReady + 1), ph_acc1); ph_acc1 ^= 1; }
                else { mbar_wait(bar(kB3AccReady + 0), ph_acc0); ph_acc0 ^= 1; }
                tc_fence_after_sync();
                const uint32_t ta = tmem_lane + ((f & JB_ACC_HI) ? 256u : 0u);

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
                            const float sg = fmaf(out_s[pl], ca.x, ca.y);
                            out_s[pl] = 0.0f;
                            const long long gi = g0 + cc * 32 + lane;
                            if (gi < prm.n_points) prm.raw[4 * gi + 3] = sg;
                        }
                    }
                    publish(kB3ActHi);
                    continue;
                }

                if (f & JB_DIR_BEFORE) {        // gamma(x) is dead once L5 has been accumulated
                    if (role == 0) write_dir_enc(enc, pt, vd);
                }
                // ---- 2 chunks of 32 points: TMEM -> y = acc*es + b -> (ReLU) -> fp16 ----
                const float lo_clamp = (f & JB_RELU) ? 0.0f : -3.0e38f;
                uint32_t pk[2][16];
                {
                    uint32_t v0[32], v1[32];
                    auto process = [&](const uint32_t (&v)[32], int cc) {
                        float y[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) y[i] = fmaxf(fmaf(__uint_as_float(v[i]), c.x, c.y), lo_clamp);
#pragma unroll
                        for (int i = 0; i < 16; ++i) pk[cc][i] = pack_half2(y[2 * i], y[2 * i + 1]);
                        if (f & JB_ALPHA) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) y[i] *= wa;
                            const float s = column_reduce32_3(y, lane);
                            atomicAdd(&out_s[pq * 64 + cc * 32 + lane], s);
                        }
                    };
                    tmem_ld32(ta, v0);
                    tmem_ld_wait();
                    tmem_ld32(ta + 32, v1);
                    process(v0, 0);
                    tmem_ld_wait();
                    process(v1, 1);
                }
                // ---- write the operand tile (channels of this half), after its last readers are done ----
                if (f & JB_WAIT_SF) { mbar_wait(bar(kB3StageFree + (q >> 1)), ph_sf); ph_sf ^= 1; }
                if (kSave) {
                    if (lane == 0) bulk_wait_read_all();
                    __syncwarp();
                }
                const uint32_t row_off = (ch >> 3) * kKGroup3 + pq * kNGroup3 + (ch & 7u) * 128u;
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint4 qv = make_uint4(pk[cc][4 * k], pk[cc][4 * k + 1], pk[cc][4 * k + 2], pk[cc][4 * k + 3]);
                        *reinterpret_cast<uint4*>(act + row_off + ((((uint32_t)(cc * 4 + k)) ^ (ch & 7u)) << 4)) = qv;
                    }
                }
                if (kSave) {
                    // this warp's 4 pieces of 1 KB (8 channels x 64 points each) -> HBM image of slot jb.slot
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        const uint32_t piece = ((128u * hi + 32u * q) >> 3) * kKGroup3 + pq * kNGroup3;
                        uint8_t* dst = save_g + (size_t)jb.slot * kAct3Bytes + piece;
#pragma unroll
                        for (int gk = 0; gk < 4; ++gk)
                            bulk_s2g(dst + gk * kKGroup3, sbase + kS3Act + piece + gk * kKGroup3, kNGroup3);
                        bulk_commit();
                    }
                }
                if (f & JB_PE_AFTER) {          // the direction stage of this group has been accumulated
                    const int next = g + stride;
                    if (next < prm.n_groups) {
                        load_point(next);
                        write_pe_half(enc, pt, role, p);
                    }
                }
                publish(hi ? kB3ActHi : kB3ActLo);
            }
        }
        if (kSave && lane == 0) bulk_wait_all();
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

}  // namespace nerfq

extern "C" int nerfq_mlp3_forward(const void* packed, const float* rays, const float* z, long long n_rays, int samples_per_ray,
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
    Fwd3Params prm{(const uint8_t*)packed, rays, z, raw, (uint8_t*)save, n_points, samples_per_ray, n_groups, prog};
    if (save) {
        if (cudaFuncSetAttribute(mlp3_forward_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kS3Bytes) != cudaSuccess) return -2;
        mlp3_forward_kernel<true><<<grid, kThreads3, kS3Bytes, stream>>>(prm);
    } else {
        if (cudaFuncSetAttribute(mlp3_forward_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kS3Bytes) != cudaSuccess) return -2;
        mlp3_forward_kernel<false><<<grid, kThreads3, kS3Bytes, stream>>>(prm);
    }
    return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

extern "C" unsigned long long nerfq_mlp3_save_bytes(long long n_points) {
    using namespace nerfq;
    const long long n_groups = (n_points + kGroupPts - 1) / kGroupPts;
    return (unsigned long long)n_groups * kSave3GroupBytes;
}
