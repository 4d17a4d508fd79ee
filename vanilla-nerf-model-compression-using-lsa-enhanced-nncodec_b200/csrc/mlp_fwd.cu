// Fused positional-encoding + NeRF MLP forward on tcgen05 tensor cores (sm_100a).
//
// Replaces, for one network and M = n_rays*S sample points:
//   pts = o + d*z                                   run_nerf.py:408,430
//   embed_fn / embeddirs_fn / cat                   run_nerf.py:52-59, run_nerf_helpers.py:18-67
//   NeRF.forward with ScaledLinear layers           utils.py:57-80, transforms.py:104-111
// Output: raw[M,4] = (rgb logits, sigma) as run_network returns it (run_nerf.py:61-63).
//
// One persistent CTA per SM processes pairs of 128-point tiles.  Per tile the 12 tensor-core steps
// of net_layout.h run D[128 x N] (fp32, TMEM) = A[128 x K] (fp16, shared, written by the epilogue
// warps) x W^T (fp16 integer levels, streamed from L2 through a 4-slot ring by bulk async copies).
// The epilogue applies  y = acc * (delta * s[o]) + b[o]  (on-the-fly dequantisation with the LSA
// scale), ReLU, converts to fp16 and writes the next layer's operand tile in place.  The 1-wide
// alpha head and the 3-wide rgb head are evaluated on CUDA cores inside the epilogues of L7 and of
// the views layer.  With `save` set, every operand tile is also streamed to HBM (bulk store) for
// the backward pass.
#include <cuda_runtime.h>
#include <stdio.h>

#include "mlp_common.cuh"

namespace nerfq {

struct FwdParams {
    const uint8_t* packed;   // packed network (net_layout.h)
    const float* rays;       // [n_rays, 11]  o(3) d(3) near far viewdir(3)
    const float* z;          // [n_rays * S]
    float* raw;              // [n_rays * S, 4]
    uint8_t* save;           // nullable: saved operand tiles, kSaveTileBytes per tile
    long long n_points;
    int samples_per_ray;
    int n_pairs;
};

enum EpiKind : int { EPI_STD = 0, EPI_WRITE_PTS = 1, EPI_WRITE_DIR = 2, EPI_FINAL = 3 };
__device__ __constant__ int kEpiKind[kFwdSteps] = {EPI_STD, EPI_STD, EPI_STD, EPI_STD, EPI_STD, EPI_WRITE_PTS,
                                                    EPI_STD, EPI_STD, EPI_STD, EPI_STD, EPI_WRITE_DIR, EPI_FINAL};
// channel base of the layer finished by each step; save slot written by each step (-1: none)
__device__ __constant__ int kEpiCh[kFwdSteps] = {0, 256, 512, 768, 1024, -1, 1280, 1536, 1792, kChFeature, -1, kChViews};
__device__ __constant__ int kEpiSlot[kFwdSteps] = {0, 1, 2, 3, 4, -1, 5, 6, 7, 8, -1, 9};
__device__ __constant__ MmaStep kFwdDev[kFwdSteps] = NERFQ_FWD_STEP_TABLE;

// ---------------------------------------------------------------------------------------------
// epilogue of one 256-wide (or 128-wide) layer for one row
// ---------------------------------------------------------------------------------------------
template <bool kRelu, bool kAlpha, bool kRgb, int kChunks, bool kWrite>
__device__ __forceinline__ void layer_epilogue(uint32_t tmem_row, uint8_t* abuf, int row, const float2* __restrict__ sb,
                                               const float* __restrict__ wvec, float* acc_out) {
    uint32_t v[2][32];
    tmem_ld32(tmem_row, v[0]);
    tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
        if (c + 1 < kChunks) tmem_ld32(tmem_row + 32 * (c + 1), v[(c + 1) & 1]);
        uint32_t* cur = v[c & 1];
        uint32_t packed[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float4 p = *reinterpret_cast<const float4*>(&sb[c * 32 + 2 * i]);
            float y0 = fmaf(__uint_as_float(cur[2 * i]), p.x, p.y);
            float y1 = fmaf(__uint_as_float(cur[2 * i + 1]), p.z, p.w);
            if (kRelu) { y0 = fmaxf(y0, 0.0f); y1 = fmaxf(y1, 0.0f); }
            if (kAlpha) {
                const float2 w = *reinterpret_cast<const float2*>(&wvec[c * 32 + 2 * i]);
                acc_out[0] = fmaf(y0, w.x, acc_out[0]);
                acc_out[0] = fmaf(y1, w.y, acc_out[0]);
            }
            if (kRgb) {
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float2 w = *reinterpret_cast<const float2*>(&wvec[k * 128 + c * 32 + 2 * i]);
                    acc_out[k] = fmaf(y0, w.x, acc_out[k]);
                    acc_out[k] = fmaf(y1, w.y, acc_out[k]);
                }
            }
            packed[i] = pack_half2(y0, y1);
        }
        if (kWrite) {
            uint8_t* blk = abuf + (c >> 1) * kABlockBytes;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint4 q = make_uint4(packed[4 * k], packed[4 * k + 1], packed[4 * k + 2], packed[4 * k + 3]);
                *reinterpret_cast<uint4*>(blk + sw128_offset(row, (c & 1) * 4 + k)) = q;
            }
        }
        if (c + 1 < kChunks) tmem_ld_wait();
    }
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
template <bool kPingPong>
__global__ void __launch_bounds__(kThreads, 1) mlp_forward_kernel(const FwdParams prm) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t sbase = smem_u32(smem);
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    float2* sb = reinterpret_cast<float2*>(smem + kSmemSB);
    float* w_alpha = reinterpret_cast<float*>(smem + kSmemWAlpha);
    float* w_rgb = reinterpret_cast<float*>(smem + kSmemWRgb);
    auto bar = [&](int i) { return sbase + kSmemBars + 8u * i; };

    // ---- one-time setup --------------------------------------------------------------------
    {
        const float2* g_sb = reinterpret_cast<const float2*>(prm.packed + kOffSB);
        for (int i = threadIdx.x; i < kNumChannels; i += kThreads) sb[i] = g_sb[i];
        const float* g_wa = reinterpret_cast<const float*>(prm.packed + kOffWAlpha);
        for (int i = threadIdx.x; i < 256 + 384; i += kThreads) w_alpha[i] = g_wa[i];   // w_rgb follows w_alpha
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < kSlots; ++i) { mbar_init(bar(kBarWFull + i), 1); mbar_init(bar(kBarWEmpty + i), 1); }
        for (int t = 0; t < 2; ++t) { mbar_init(bar(kBarActReady + t), kEpiWarpsPerTile); mbar_init(bar(kBarAccReady + t), 1); }
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc(sbase + kSmemTmemPtr, 512);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + kSmemTmemPtr);

    const int first_pair = blockIdx.x;
    const int pair_stride = gridDim.x;

    if (warp == 0) {
        // ================= weight loader =================
        if (lane == 0) {
            const uint8_t* img = prm.packed + kOffFwdImage;
            uint32_t seq = 0;
            for (int pair = first_pair; pair < prm.n_pairs; pair += pair_stride) {
                uint32_t off = 0;
                for (int s = 0; s < kFwdSteps; ++s) {
                    const int nst = kFwdDev[s].stages;
                    const uint32_t bytes = kFwdDev[s].n * kStageRowBytes;
                    for (int rep = 0; rep < (kPingPong ? 2 : 1); ++rep) {
                        for (int i = 0; i < nst; ++i, ++seq) {
                            const uint32_t slot = seq % kSlots, par = (seq / kSlots) & 1;
                            mbar_wait(bar(kBarWEmpty + slot), par ^ 1);
                            mbar_arrive_expect_tx(bar(kBarWFull + slot), bytes);
                            bulk_g2s(sbase + kSmemRing + slot * kSlotBytes, img + off + i * bytes, bytes, bar(kBarWFull + slot));
                        }
                    }
                    off += nst * bytes;
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            uint32_t seq = 0, n_act[2] = {0, 0};
            for (int pair = first_pair; pair < prm.n_pairs; pair += pair_stride) {
                for (int s = 0; s < kFwdSteps; ++s) {
                    const MmaStep st = kFwdDev[s];
                    const uint32_t idesc = umma_idesc(128, st.n, false);
                    if (!kPingPong) {
                        mbar_wait(bar(kBarActReady + 0), n_act[0]++ & 1);
                        mbar_wait(bar(kBarActReady + 1), n_act[1]++ & 1);
                        tc_fence_after_sync();
                        for (int i = 0; i < st.stages; ++i, ++seq) {
                            const uint32_t slot = seq % kSlots, par = (seq / kSlots) & 1;
                            mbar_wait(bar(kBarWFull + slot), par);
                            tc_fence_after_sync();
                            const uint32_t b_addr = sbase + kSmemRing + slot * kSlotBytes;
#pragma unroll
                            for (int t = 0; t < 2; ++t) {
                                const uint32_t a_addr = sbase + kSmemABuf + t * kABufBytes + (st.a_blk0 + (i >> 1)) * kABlockBytes + (i & 1) * 64;
#pragma unroll
                                for (int j = 0; j < 2; ++j)
                                    umma_ss(tmem_base + t * 256, umma_smem_desc(a_addr + j * 32, 1024, SWZ_128B),
                                            umma_smem_desc(b_addr + j * 32, 512, SWZ_64B), idesc, (i | j | st.accumulate) ? 1u : 0u);
                            }
                            umma_commit(bar(kBarWEmpty + slot));
                        }
                        umma_commit(bar(kBarAccReady + 0));
                        umma_commit(bar(kBarAccReady + 1));
                    } else {
                        for (int t = 0; t < 2; ++t) {
                            mbar_wait(bar(kBarActReady + t), n_act[t]++ & 1);
                            tc_fence_after_sync();
                            for (int i = 0; i < st.stages; ++i, ++seq) {
                                const uint32_t slot = seq % kSlots, par = (seq / kSlots) & 1;
                                mbar_wait(bar(kBarWFull + slot), par);
                                tc_fence_after_sync();
                                const uint32_t b_addr = sbase + kSmemRing + slot * kSlotBytes;
                                const uint32_t a_addr = sbase + kSmemABuf + t * kABufBytes + (st.a_blk0 + (i >> 1)) * kABlockBytes + (i & 1) * 64;
#pragma unroll
                                for (int j = 0; j < 2; ++j)
                                    umma_ss(tmem_base + t * 256, umma_smem_desc(a_addr + j * 32, 1024, SWZ_128B),
                                            umma_smem_desc(b_addr + j * 32, 512, SWZ_64B), idesc, (i | j | st.accumulate) ? 1u : 0u);
                                umma_commit(bar(kBarWEmpty + slot));
                            }
                            umma_commit(bar(kBarAccReady + t));
                        }
                    }
                }
            }
        }
    } else if (warp >= kCtrlWarps) {
        // ================= epilogue warps =================
        const int t = (warp - kCtrlWarps) / kEpiWarpsPerTile;
        const int q = warp & 3;
        const int row = q * 32 + lane;
        uint8_t* abuf = smem + kSmemABuf + t * kABufBytes;
        const uint32_t abuf_s = sbase + kSmemABuf + t * kABufBytes;
        const uint32_t tmem_row = tmem_base + (uint32_t(q * 32) << 16) + t * 256;
        const bool saving = prm.save != nullptr;
        const bool store_leader = (row == 0);
        uint32_t n_acc = 0;

        float p[3], vd[3];
        long long g = 0;
        auto load_point = [&](int pair) {
            const long long tile = 2ll * pair + t;
            g = tile * kTileM + row;
            const long long gc = g < prm.n_points ? g : prm.n_points - 1;
            const long long ray = gc / prm.samples_per_ray;
            const float zz = __ldg(prm.z + gc);
            const float* r = prm.rays + ray * 11;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                p[k] = fmaf(__ldg(r + 3 + k), zz, __ldg(r + k));
                vd[k] = __ldg(r + 8 + k);
            }
        };
        auto publish = [&]() {   // operand tile written / accumulator drained -> MMA issuer may proceed
            fence_proxy_async_smem();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(kBarActReady + t));
        };

        if (first_pair < prm.n_pairs) {
            load_point(first_pair);
            write_pts_encoding(abuf, row, p);
            publish();
        }
        for (int pair = first_pair; pair < prm.n_pairs; pair += pair_stride) {
            float sigma_acc = 0.0f;
            uint8_t* save_tile = saving ? prm.save + (size_t)(2ll * pair + t) * kSaveTileBytes : nullptr;
            for (int s = 0; s < kFwdSteps; ++s) {
                mbar_wait(bar(kBarAccReady + t), n_acc++ & 1);
                tc_fence_after_sync();
                const int kind = kEpiKind[s];
                if (saving) {   // the previous bulk store must have finished reading the tile before it is overwritten
                    if (store_leader) bulk_wait_read_all();
                    named_bar_sync(1 + t, 32 * kEpiWarpsPerTile);
                }
                if (kind == EPI_STD) {
                    const float2* sbl = sb + kEpiCh[s];
                    if (s == 8) layer_epilogue<true, true, false, 8, true>(tmem_row, abuf, row, sbl, w_alpha, &sigma_acc);
                    else if (s == 9) layer_epilogue<false, false, false, 8, true>(tmem_row, abuf, row, sbl, nullptr, nullptr);
                    else layer_epilogue<true, false, false, 8, true>(tmem_row, abuf, row, sbl, nullptr, nullptr);
                    if (saving) {
                        fence_proxy_async_smem();
                        named_bar_sync(1 + t, 32 * kEpiWarpsPerTile);
                        if (store_leader) {
                            bulk_s2g(save_tile + (size_t)kEpiSlot[s] * kABufBytes, abuf_s, kABufBytes);
                            bulk_commit();
                        }
                    }
                    publish();
                } else if (kind == EPI_WRITE_PTS) {
                    write_pts_encoding(abuf, row, p);
                    publish();
                } else if (kind == EPI_WRITE_DIR) {
                    write_dir_encoding(abuf, row, vd);
                    publish();
                } else {
                    float rgb_acc[3] = {0.0f, 0.0f, 0.0f};
                    if (saving) {
                        layer_epilogue<true, false, true, 4, true>(tmem_row, abuf, row, sb + kChViews, w_rgb, rgb_acc);
                        fence_proxy_async_smem();
                        named_bar_sync(1 + t, 32 * kEpiWarpsPerTile);
                        if (store_leader) {
                            bulk_s2g(save_tile + (size_t)kSaveSlotsFull * kABufBytes, abuf_s, 2 * kABlockBytes);
                            bulk_commit();
                            bulk_wait_read_all();
                        }
                        named_bar_sync(1 + t, 32 * kEpiWarpsPerTile);
                    } else {
                        layer_epilogue<true, false, true, 4, false>(tmem_row, abuf, row, sb + kChViews, w_rgb, rgb_acc);
                    }
                    if (g < prm.n_points) {
                        float4 o;
                        o.x = fmaf(rgb_acc[0], sb[kChRgb + 0].x, sb[kChRgb + 0].y);
                        o.y = fmaf(rgb_acc[1], sb[kChRgb + 1].x, sb[kChRgb + 1].y);
                        o.z = fmaf(rgb_acc[2], sb[kChRgb + 2].x, sb[kChRgb + 2].y);
                        o.w = fmaf(sigma_acc, sb[kChAlpha].x, sb[kChAlpha].y);
                        *reinterpret_cast<float4*>(prm.raw + 4 * g) = o;
                    }
                    const int next = pair + pair_stride;
                    if (next < prm.n_pairs) {
                        load_point(next);
                        write_pts_encoding(abuf, row, p);
                        publish();
                    }
                }
            }
        }
        if (saving && store_leader) bulk_wait_all();
    }

    // ---- teardown ----------------------------------------------------------------------------
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

}  // namespace nerfq

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" int nerfq_mlp_forward(const void* packed, const float* rays, const float* z, long long n_rays,
                                 int samples_per_ray, float* raw, void* save, int pingpong, int max_ctas, cudaStream_t stream) {
    using namespace nerfq;
    if (n_rays == 0) return 0;
    if (!packed || !rays || !z || !raw || n_rays < 0 || samples_per_ray <= 0) return -1;
    const long long n_points = n_rays * samples_per_ray;
    const long long n_tiles = (n_points + kTileM - 1) / kTileM;
    const int n_pairs = (int)((n_tiles + 1) / 2);
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (max_ctas > 0 && max_ctas < sms) sms = max_ctas;
    const int grid = n_pairs < sms ? n_pairs : sms;
    FwdParams prm{(const uint8_t*)packed, rays, z, raw, (uint8_t*)save, n_points, samples_per_ray, n_pairs};
    cudaError_t e;
    if (pingpong) {
        e = cudaFuncSetAttribute(mlp_forward_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytesFwd);
        if (e != cudaSuccess) return -2;
        mlp_forward_kernel<true><<<grid, kThreads, kSmemBytesFwd, stream>>>(prm);
    } else {
        e = cudaFuncSetAttribute(mlp_forward_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytesFwd);
        if (e != cudaSuccess) return -2;
        mlp_forward_kernel<false><<<grid, kThreads, kSmemBytesFwd, stream>>>(prm);
    }
    e = cudaGetLastError();
    return e == cudaSuccess ? 0 : -3;
}

extern "C" unsigned long long nerfq_mlp_save_bytes(long long n_points) {
    using namespace nerfq;
    const long long n_tiles = (n_points + kTileM - 1) / kTileM;
    const long long n_pairs = (n_tiles + 1) / 2;
    return (unsigned long long)(2 * n_pairs) * kSaveTileBytes;
}
