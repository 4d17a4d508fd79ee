// Fused NeRF MLP backward into the LSA scales on tcgen05 tensor cores (sm_100a).
//
// Replaces torch autograd over NeRF.forward with ScaledLinear layers (utils.py:57-80,
// transforms.py:104-111) when only `weight_scaling` requires grad
// (framework/pytorch_model/__init__.py:1129-1145, run_nerf.py:756).  For a layer
//   y = s * (delta * (L x)) + b      (L = integer levels, s = LSA scale per output channel)
// the scale gradient is  ds[o] = sum_n dY[n,o] * (y[n,o] - b[o]) / s[o]  -- an elementwise product
// reduced over points, no weight-gradient GEMM -- and the input gradient is the dgrad GEMM
//   dX = (dY * delta * s) L.
// Per 128-point tile the chain of 9 dgrad GEMMs (net_layout.h kBwd) runs on the tensor cores with
// fp16 operands and fp32 accumulation in TMEM.  Gradients are tiny (1e-7..1e-3), so every row (point) is
// first multiplied by a power of two that brings max|d_raw| into [8,16) -- backpropagation is linear per
// row, the factor is exact and is divided out again where the scale-gradient products are formed; this
// keeps fp16's 11-bit significand (8x finer than bf16) without its range problem.  The epilogue of
// each step reads the accumulator, masks it with the saved forward activation (ReLU), accumulates
// the scale-gradient partial sums (warp butterfly transpose-reduce -> shared atomics) and writes the
// next GEMM's operand tile in place.  Saved activations (written by the forward kernel) are
// streamed by bulk async copies directly into the operand-tile blocks the tensor cores have just
// finished reading, so they need no staging buffer.
#include <cuda_runtime.h>

#include "mlp_common.cuh"

namespace nerfq {

struct BwdParams {
    const uint8_t* packed;
    const float* d_raw;      // [n_points, 4]
    const float* raw;        // [n_points, 4]   forward output (for the rgb / alpha head terms)
    const uint8_t* save;     // saved operand tiles from the forward pass
    float* d_scale;          // [2436], accumulated with atomics (caller zeroes)
    long long n_points;
    int n_pairs;
};

__device__ __constant__ MmaStep kBwdDev[kBwdSteps] = NERFQ_BWD_STEP_TABLE;
constexpr int kBarTileDone = 24;   // [2]

// Sum each of 32 per-lane columns over the 32 lanes of the warp; lane j returns column j.
__device__ __forceinline__ float column_reduce32(float (&p)[32], int lane) {
    float q16[16];
    {
        const bool hi = lane & 16;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float send = hi ? p[i] : p[i + 16];
            const float keep = hi ? p[i + 16] : p[i];
            q16[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
    }
    float q8[8];
    {
        const bool hi = lane & 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float send = hi ? q16[i] : q16[i + 8];
            const float keep = hi ? q16[i + 8] : q16[i];
            q8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
    }
    float q4[4];
    {
        const bool hi = lane & 4;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float send = hi ? q8[i] : q8[i + 4];
            const float keep = hi ? q8[i + 4] : q8[i];
            q4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
    }
    float q2[2];
    {
        const bool hi = lane & 2;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const float send = hi ? q4[i] : q4[i + 2];
            const float keep = hi ? q4[i + 2] : q4[i];
            q2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
        }
    }
    const bool hi = lane & 1;
    const float send = hi ? q2[0] : q2[1];
    const float keep = hi ? q2[1] : q2[0];
    return keep + __shfl_xor_sync(0xffffffffu, send, 1);
}

// One 32-column chunk of a backward epilogue.
//   d[32]   : gradient w.r.t. the layer output (post-activation), fp32
//   abuf    : operand tile holding the saved activation chunk (fp16) -> overwritten with dY*eff_scale (fp16)
//   rinv    : inverse of this row's power-of-two gradient scale
template <bool kRelu, bool kWriteA>
__device__ __forceinline__ void bwd_chunk(float (&d)[32], uint8_t* blk, int row, int chunk_in_blk, const float2* __restrict__ sb,
                                          float* __restrict__ red, int lane, float rinv) {
    uint32_t hraw[16];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint4 q = *reinterpret_cast<const uint4*>(blk + sw128_offset(row, chunk_in_blk * 4 + k));
        hraw[4 * k] = q.x; hraw[4 * k + 1] = q.y; hraw[4 * k + 2] = q.z; hraw[4 * k + 3] = q.w;
    }
    float p[32];
    uint32_t packed[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const float2 h = __half22float2(*reinterpret_cast<const __half2*>(&hraw[i]));
        const float4 c = *reinterpret_cast<const float4*>(&sb[2 * i]);   // es0, b0, es1, b1
        float dy0 = d[2 * i], dy1 = d[2 * i + 1];
        if (kRelu) {
            dy0 = h.x > 0.0f ? dy0 : 0.0f;
            dy1 = h.y > 0.0f ? dy1 : 0.0f;
        }
        p[2 * i] = (dy0 * rinv) * (h.x - c.y);
        p[2 * i + 1] = (dy1 * rinv) * (h.y - c.w);
        packed[i] = pack_half2(dy0 * c.x, dy1 * c.z);
    }
    if (kWriteA) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint4 q = make_uint4(packed[4 * k], packed[4 * k + 1], packed[4 * k + 2], packed[4 * k + 3]);
            *reinterpret_cast<uint4*>(blk + sw128_offset(row, chunk_in_blk * 4 + k)) = q;
        }
    }
    const float colsum = column_reduce32(p, lane);
    atomicAdd(red + lane, colsum);
}

__global__ void __launch_bounds__(kThreads, 1) mlp_backward_kernel(const BwdParams prm) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t sbase = smem_u32(smem);
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    float2* sb = reinterpret_cast<float2*>(smem + kSmemSB);
    float* w_alpha = reinterpret_cast<float*>(smem + kSmemWAlpha);
    float* w_rgb = reinterpret_cast<float*>(smem + kSmemWRgb);
    float* red = reinterpret_cast<float*>(smem + kSmemRed);
    auto bar = [&](int i) { return sbase + kSmemBars + 8u * i; };

    {
        const float2* g_sb = reinterpret_cast<const float2*>(prm.packed + kOffSB);
        for (int i = threadIdx.x; i < kNumChannels; i += kThreads) { sb[i] = g_sb[i]; red[i] = 0.0f; }
        const float* g_wa = reinterpret_cast<const float*>(prm.packed + kOffWAlpha);
        for (int i = threadIdx.x; i < 256 + 384; i += kThreads) w_alpha[i] = g_wa[i];
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < kSlots; ++i) { mbar_init(bar(kBarWFull + i), 1); mbar_init(bar(kBarWEmpty + i), 1); }
        for (int t = 0; t < 2; ++t) {
            mbar_init(bar(kBarActReady + t), kEpiWarpsPerTile);
            mbar_init(bar(kBarAccReady + t), 1);
            mbar_init(bar(kBarTileDone + t), kEpiWarpsPerTile);
            for (int b = 0; b < 4; ++b) mbar_init(bar(kBarHFull + 4 * t + b), 1);
        }
        for (int b = 0; b < 4; ++b) mbar_init(bar(kBarBlkFree + b), 1);
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
        // ================= weight loader (W^T stages) =================
        if (lane == 0) {
            const uint8_t* img = prm.packed + kOffBwdImage;
            uint32_t seq = 0;
            for (int pair = first_pair; pair < prm.n_pairs; pair += pair_stride) {
                uint32_t off = 0;
                for (int s = 0; s < kBwdSteps; ++s) {
                    const int nst = kBwdDev[s].stages;
                    const uint32_t bytes = kBwdDev[s].n * kStageRowBytes;
                    for (int i = 0; i < nst; ++i, ++seq) {
                        const uint32_t slot = seq % kSlots, par = (seq / kSlots) & 1;
                        mbar_wait(bar(kBarWEmpty + slot), par ^ 1);
                        mbar_arrive_expect_tx(bar(kBarWFull + slot), bytes);
                        bulk_g2s(sbase + kSmemRing + slot * kSlotBytes, img + off + i * bytes, bytes, bar(kBarWFull + slot));
                    }
                    off += nst * bytes;
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            uint32_t seq = 0, n_act = 0;
            const uint32_t idesc = umma_idesc(128, 256, false);
            for (int pair = first_pair; pair < prm.n_pairs; pair += pair_stride) {
                for (int s = 0; s < kBwdSteps; ++s) {
                    const MmaStep st = kBwdDev[s];
                    mbar_wait(bar(kBarActReady + 0), n_act & 1);
                    mbar_wait(bar(kBarActReady + 1), n_act & 1);
                    ++n_act;
                    tc_fence_after_sync();
                    if (st.stages < 8) {   // blocks this step does not read are free for the activation loader at once
                        for (int b = st.stages / 2; b < 4; ++b) umma_commit(bar(kBarBlkFree + b));
                    }
                    for (int i = 0; i < st.stages; ++i, ++seq) {
                        const uint32_t slot = seq % kSlots, par = (seq / kSlots) & 1;
                        mbar_wait(bar(kBarWFull + slot), par);
                        tc_fence_after_sync();
                        const uint32_t b_addr = sbase + kSmemRing + slot * kSlotBytes;
#pragma unroll
                        for (int t = 0; t < 2; ++t) {
                            const uint32_t a_addr = sbase + kSmemABuf + t * kABufBytes + (i >> 1) * kABlockBytes + (i & 1) * 64;
#pragma unroll
                            for (int j = 0; j < 2; ++j)
                                umma_ss(tmem_base + t * 256, umma_smem_desc(a_addr + j * 32, 1024, SWZ_128B),
                                        umma_smem_desc(b_addr + j * 32, 512, SWZ_64B), idesc, (i | j) ? 1u : 0u);
                        }
                        umma_commit(bar(kBarWEmpty + slot));
                        if (i & 1) umma_commit(bar(kBarBlkFree + (i >> 1)));
                    }
                    umma_commit(bar(kBarAccReady + 0));
                    umma_commit(bar(kBarAccReady + 1));
                }
            }
        }
    } else if (warp == 3) {
        // ================= saved-activation loader =================
        if (lane == 0) {
            uint32_t n_free = 0, n_done = 0;
            for (int pair = first_pair; pair < prm.n_pairs; pair += pair_stride) {
                const uint8_t* tile_src[2] = {prm.save + (size_t)(2ll * pair) * kSaveTileBytes,
                                              prm.save + (size_t)(2ll * pair + 1) * kSaveTileBytes};
                for (int t = 0; t < 2; ++t) {
                    if (pair != first_pair) mbar_wait(bar(kBarTileDone + t), (n_done & 1));
                    for (int b = 0; b < 2; ++b) {   // views hidden (128 wide) -> blocks 0,1
                        mbar_arrive_expect_tx(bar(kBarHFull + 4 * t + b), kABlockBytes);
                        bulk_g2s(sbase + kSmemABuf + t * kABufBytes + b * kABlockBytes,
                                 tile_src[t] + (size_t)kSaveSlotsFull * kABufBytes + b * kABlockBytes, kABlockBytes,
                                 bar(kBarHFull + 4 * t + b));
                    }
                }
                if (pair != first_pair) ++n_done;
                for (int s = 0; s < kBwdSteps; ++s) {
                    const int slot = 8 - s;
                    for (int b = 0; b < 4; ++b) {
                        mbar_wait(bar(kBarBlkFree + b), n_free & 1);
                        for (int t = 0; t < 2; ++t) {
                            mbar_arrive_expect_tx(bar(kBarHFull + 4 * t + b), kABlockBytes);
                            bulk_g2s(sbase + kSmemABuf + t * kABufBytes + b * kABlockBytes,
                                     tile_src[t] + (size_t)slot * kABufBytes + b * kABlockBytes, kABlockBytes,
                                     bar(kBarHFull + 4 * t + b));
                        }
                    }
                    ++n_free;
                }
            }
        }
    } else if (warp >= kCtrlWarps) {
        // ================= epilogue warps =================
        const int t = (warp - kCtrlWarps) / kEpiWarpsPerTile;
        const int q = warp & 3;
        const int row = q * 32 + lane;
        uint8_t* abuf = smem + kSmemABuf + t * kABufBytes;
        const uint32_t tmem_row = tmem_base + (uint32_t(q * 32) << 16) + t * 256;
        uint32_t n_acc = 0, n_h[4] = {0, 0, 0, 0};

        auto publish = [&]() {
            fence_proxy_async_smem();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(kBarActReady + t));
        };
        auto wait_h = [&](int b) { mbar_wait(bar(kBarHFull + 4 * t + b), n_h[b]++ & 1); };

        for (int pair = first_pair; pair < prm.n_pairs; pair += pair_stride) {
            const long long g = (2ll * pair + t) * kTileM + row;
            float4 dr = make_float4(0.f, 0.f, 0.f, 0.f), rw = make_float4(0.f, 0.f, 0.f, 0.f);
            if (g < prm.n_points) {
                dr = *reinterpret_cast<const float4*>(prm.d_raw + 4 * g);
                rw = *reinterpret_cast<const float4*>(prm.raw + 4 * g);
            }
            // ---- per-row power-of-two gradient scale: max|d_raw| -> [8,16) ----
            float rinv = 0.0f;
            {
                const float rowmax = fmaxf(fmaxf(fabsf(dr.x), fabsf(dr.y)), fmaxf(fabsf(dr.z), fabsf(dr.w)));
                const int ex = (__float_as_int(rowmax) >> 23) & 0xff;
                float rscale = 0.0f;
                if (ex >= 16 && ex <= 240) {
                    rscale = __int_as_float((257 - ex) << 23);
                    rinv = __int_as_float((ex - 3) << 23);
                }
                dr.x *= rscale; dr.y *= rscale; dr.z *= rscale; dr.w *= rscale;
            }
            // ---- heads: rgb_linear (3 x 128) and alpha_linear (1 x 256) on CUDA cores ----
            {
                float pr = dr.x * rinv * (rw.x - sb[kChRgb + 0].y);
                float pg = dr.y * rinv * (rw.y - sb[kChRgb + 1].y);
                float pb = dr.z * rinv * (rw.z - sb[kChRgb + 2].y);
                float pa = dr.w * rinv * (rw.w - sb[kChAlpha].y);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    pr += __shfl_xor_sync(0xffffffffu, pr, o);
                    pg += __shfl_xor_sync(0xffffffffu, pg, o);
                    pb += __shfl_xor_sync(0xffffffffu, pb, o);
                    pa += __shfl_xor_sync(0xffffffffu, pa, o);
                }
                if (lane == 0) {
                    atomicAdd(red + kChRgb + 0, pr);
                    atomicAdd(red + kChRgb + 1, pg);
                    atomicAdd(red + kChRgb + 2, pb);
                    atomicAdd(red + kChAlpha, pa);
                }
            }
            const float gr = dr.x * sb[kChRgb + 0].x, gg = dr.y * sb[kChRgb + 1].x, gb = dr.z * sb[kChRgb + 2].x;
            const float ga = dr.w * sb[kChAlpha].x;
            // ---- prep: views layer gradient from the rgb head; operand for the first dgrad GEMM ----
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                if ((c & 1) == 0) wait_h(c >> 1);
                float d[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int k = c * 32 + i;
                    d[i] = gr * w_rgb[k] + gg * w_rgb[128 + k] + gb * w_rgb[256 + k];
                }
                bwd_chunk<true, true>(d, abuf + (c >> 1) * kABlockBytes, row, c & 1, sb + kChViews + c * 32,
                                      red + kChViews + c * 32, lane, rinv);
            }
            publish();
            // ---- dgrad chain ----
            for (int s = 0; s < kBwdSteps; ++s) {
                mbar_wait(bar(kBarAccReady + t), n_acc++ & 1);
                tc_fence_after_sync();
                const int ch = (s == 0) ? kChFeature : 256 * (8 - s);
                const bool last = (s == kBwdSteps - 1);
#pragma unroll 1
                for (int c = 0; c < 8; ++c) {
                    if ((c & 1) == 0) wait_h(c >> 1);
                    uint32_t v[32];
                    tmem_ld32(tmem_row + c * 32, v);
                    tmem_ld_wait();
                    float d[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) d[i] = __uint_as_float(v[i]);
                    if (s == 1) {   // d h8 also receives the alpha head's gradient
#pragma unroll
                        for (int i = 0; i < 32; ++i) d[i] = fmaf(ga, w_alpha[c * 32 + i], d[i]);
                    }
                    uint8_t* blk = abuf + (c >> 1) * kABlockBytes;
                    if (s == 0) bwd_chunk<false, true>(d, blk, row, c & 1, sb + ch + c * 32, red + ch + c * 32, lane, rinv);
                    else if (!last) bwd_chunk<true, true>(d, blk, row, c & 1, sb + ch + c * 32, red + ch + c * 32, lane, rinv);
                    else bwd_chunk<true, false>(d, blk, row, c & 1, sb + ch + c * 32, red + ch + c * 32, lane, rinv);
                }
                if (!last) {
                    publish();
                } else {
                    tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar(kBarTileDone + t));
                }
            }
        }
    }

    // ---- teardown: flush the per-CTA partial scale gradients ----
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
    const float* scale = reinterpret_cast<const float*>(prm.packed + kOffScale);
    for (int i = threadIdx.x; i < kNumChannels; i += kThreads) {
        const float v = red[i];
        if (v != 0.0f) atomicAdd(prm.d_scale + i, v / scale[i]);
    }
}

}  // namespace nerfq

extern "C" int nerfq_mlp_backward(const void* packed, const float* d_raw, const float* raw, const void* save, long long n_points,
                                  float* d_scale, int max_ctas, cudaStream_t stream) {
    using namespace nerfq;
    if (n_points == 0) return 0;
    if (!packed || !d_raw || !raw || !save || !d_scale || n_points < 0) return -1;
    const long long n_tiles = (n_points + kTileM - 1) / kTileM;
    const int n_pairs = (int)((n_tiles + 1) / 2);
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (max_ctas > 0 && max_ctas < sms) sms = max_ctas;
    const int grid = n_pairs < sms ? n_pairs : sms;
    BwdParams prm{(const uint8_t*)packed, d_raw, raw, (const uint8_t*)save, d_scale, n_points, n_pairs};
    if (cudaFuncSetAttribute(mlp_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytesBwd) != cudaSuccess) return -2;
    mlp_backward_kernel<<<grid, kThreads, kSmemBytesBwd, stream>>>(prm);
    return cudaGetLastError() == cudaSuccess ? 0 : -3;
}
