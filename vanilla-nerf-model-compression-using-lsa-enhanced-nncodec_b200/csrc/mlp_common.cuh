// Shared device-side pieces of the fused MLP kernels (forward and backward):
// shared-memory map, warp roles, positional-encoding writers.
#pragma once
#include "net_layout.h"
#include "ptx_sm100.cuh"

namespace nerfq {

// ---- CTA shape -----------------------------------------------------------------------------
// warp 0: weight loader (bulk copies)   warp 1: MMA issuer   warp 2: TMEM owner   warp 3: aux loader
// warps 4..7: epilogue of tile 0        warps 8..11: epilogue of tile 1
// (an epilogue warp may only touch TMEM lanes 32*(warp%4)..+31, hence the 4-aligned start)
constexpr int kCtrlWarps = 4;
constexpr int kEpiWarpsPerTile = 4;
constexpr int kTilesPerCta = 2;
constexpr int kThreads = 32 * (kCtrlWarps + kEpiWarpsPerTile * kTilesPerCta);
constexpr int kSlots = 4;                  // weight ring depth
constexpr int kSlotBytes = 256 * kStageRowBytes;

// ---- shared memory map (offsets from a 1024-aligned base) -----------------------------------
constexpr uint32_t kSmemABuf = 0;                                   // 2 x 64 KB operand tiles
constexpr uint32_t kSmemRing = kSmemABuf + kTilesPerCta * kABufBytes;
constexpr uint32_t kSmemSB = kSmemRing + kSlots * kSlotBytes;       // float2[2436]
constexpr uint32_t kSmemWAlpha = kSmemSB + 8 * kNumChannels;        // float[256]
constexpr uint32_t kSmemWRgb = kSmemWAlpha + 4 * 256;               // float[384]
constexpr uint32_t kSmemBars = (kSmemWRgb + 4 * 384 + 15) / 16 * 16;
constexpr uint32_t kNumBars = 32;
constexpr uint32_t kSmemTmemPtr = kSmemBars + 8 * kNumBars;
constexpr uint32_t kSmemRed = kSmemTmemPtr + 16;                    // backward: float[2436] partial d_scale
constexpr uint32_t kSmemBytesFwd = kSmemRed + 1024;                 // + slack for manual 1 KB alignment
constexpr uint32_t kSmemBytesBwd = kSmemRed + 4 * kNumChannels + 1024;

// barrier indices
constexpr int kBarWFull = 0;      // [kSlots]
constexpr int kBarWEmpty = 4;     // [kSlots]
constexpr int kBarActReady = 8;   // [2]  operand tile written + accumulator drained (4 warp arrivals)
constexpr int kBarAccReady = 10;  // [2]  accumulator complete (tcgen05.commit)
constexpr int kBarBlkFree = 12;   // [4]  backward: operand block b consumed by the MMAs of this step
constexpr int kBarHFull = 16;     // [2][4] backward: saved-activation block landed in the operand tile

// ---- positional encodings written straight into operand tiles --------------------------------
// gamma(p) (run_nerf_helpers.py:23-49): column 3+6l+c = sin(2^l p_c), 3+6l+3+c = cos(2^l p_c).
// sin/cos are evaluated with sincosf at l = 0 and l = 5 and by angle doubling in between, which keeps
// the error below 2^4 ulp(1) ~ 1e-6, far under the fp16 rounding of the operand (2.4e-4).
template <int L>
__device__ __forceinline__ void encode3(const float p[3], float* out /* 3 + 6L */) {
    out[0] = p[0]; out[1] = p[1]; out[2] = p[2];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float s, co;
#pragma unroll
        for (int l = 0; l < L; ++l) {
            if (l == 0) {
                sincosf(p[c], &s, &co);
            } else if (l == 5) {
                sincosf(p[c] * 32.0f, &s, &co);
            } else {
                float s2 = 2.0f * s * co;
                float c2 = fmaf(-2.0f * s, s, 1.0f);
                s = s2; co = c2;
            }
            out[3 + 6 * l + c] = s;
            out[3 + 6 * l + 3 + c] = co;
        }
    }
}

// Write gamma(p) (63 values + one zero) as fp16 into block 0 of a SWIZZLE_128B operand tile.
__device__ __forceinline__ void write_pts_encoding(uint8_t* abuf, int row, const float p[3]) {
    float v[64];
    encode3<10>(p, v);
    v[63] = 0.0f;
#pragma unroll
    for (int ch = 0; ch < 8; ++ch) {
        uint4 q;
        q.x = pack_half2(v[8 * ch + 0], v[8 * ch + 1]);
        q.y = pack_half2(v[8 * ch + 2], v[8 * ch + 3]);
        q.z = pack_half2(v[8 * ch + 4], v[8 * ch + 5]);
        q.w = pack_half2(v[8 * ch + 6], v[8 * ch + 7]);
        *reinterpret_cast<uint4*>(abuf + sw128_offset(row, ch)) = q;
    }
}

// Write gamma(d) (27 values + 5 zeros) as fp16 into the first 32 columns of block 0.
__device__ __forceinline__ void write_dir_encoding(uint8_t* abuf, int row, const float d[3]) {
    float v[32];
    encode3<4>(d, v);
#pragma unroll
    for (int i = 27; i < 32; ++i) v[i] = 0.0f;
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
        uint4 q;
        q.x = pack_half2(v[8 * ch + 0], v[8 * ch + 1]);
        q.y = pack_half2(v[8 * ch + 2], v[8 * ch + 3]);
        q.z = pack_half2(v[8 * ch + 4], v[8 * ch + 5]);
        q.w = pack_half2(v[8 * ch + 6], v[8 * ch + 7]);
        *reinterpret_cast<uint4*>(abuf + sw128_offset(row, ch)) = q;
    }
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

}  // namespace nerfq
